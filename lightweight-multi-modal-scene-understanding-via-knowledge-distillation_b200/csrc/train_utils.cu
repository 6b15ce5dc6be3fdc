// Training-step helpers that keep the step free of host round trips (sm_100a):
//   * confusion matrix of SegmentationMetrics.update   (reference src/training/trainer.py:18-26,
//     a per-pixel Python loop after a .cpu() copy there)
//   * AdamW over one flat parameter buffer              (torch.optim.AdamW, trainer.py:56,90;
//     the reference steps 96 small tensors)
#include "kdf_common.cuh"

namespace kdf {

constexpr int CM_MAX_K = 8;

template <typename TL>
__global__ void __launch_bounds__(256)
confusion_kernel(const TL *__restrict__ logits, const int64_t *__restrict__ labels, int B, int K, int64_t HW,
                 int64_t ignore_index, unsigned long long *__restrict__ conf) {
    __shared__ unsigned int hist[CM_MAX_K * CM_MAX_K];
    for (int i = threadIdx.x; i < K * K; i += blockDim.x) hist[i] = 0u;
    __syncthreads();
    const int64_t npix = (int64_t)B * HW;
    const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += nthreads) {
        const int64_t t = labels[i];
        if (t == ignore_index || t < 0 || t >= K) continue;
        const int64_t b = i / HW, hw = i - b * HW;
        const TL *z = logits + b * K * HW + hw;
        float best = to_float<TL>(z[0]);
        int arg = 0;
        for (int k = 1; k < K; ++k) {                 // torch.argmax: first maximal index wins
            const float v = to_float<TL>(z[(int64_t)k * HW]);
            if (v > best || (v != v && best == best)) { best = v; arg = k; }   // NaN counts as max like torch
        }
        atomicAdd(&hist[(int)t * K + arg], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < K * K; i += blockDim.x)
        if (hist[i]) atomicAdd(conf + i, (unsigned long long)hist[i]);
}

// torch.optim.AdamW (amsgrad=False, maximize=False), single-tensor formulation:
//   p *= 1 - lr*wd ; m = lerp(m, g, 1-b1) ; v = b2*v + (1-b2) g^2
//   p -= (lr / (1-b1^t)) * m / (sqrt(v)/sqrt(1-b2^t) + eps)
__global__ void __launch_bounds__(256)
adamw_flat_kernel(float *__restrict__ p, const float *__restrict__ g, float *__restrict__ m, float *__restrict__ v,
                  int64_t n, const float *__restrict__ hyper, float beta1, float beta2, float eps, float wd,
                  float grad_scale, __nv_bfloat16 *__restrict__ p16) {
    __shared__ float sh[3];
    if (threadIdx.x == 0) {
        const double lr = (double)hyper[0];
        const double step = (double)hyper[1];
        const double bc1 = 1.0 - pow((double)beta1, step);
        const double bc2 = 1.0 - pow((double)beta2, step);
        sh[0] = (float)(lr / bc1);            // step_size
        sh[1] = (float)sqrt(bc2);             // bias_correction2_sqrt
        sh[2] = (float)(1.0 - lr * (double)wd);
    }
    __syncthreads();
    const float step_size = sh[0], bc2s = sh[1], decay = sh[2];
    const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t n4 = n >> 2;
    for (int64_t i = tid; i < n4; i += nthreads) {
        float4 pp = reinterpret_cast<float4 *>(p)[i];
        float4 gg = reinterpret_cast<const float4 *>(g)[i];
        float4 mm = reinterpret_cast<float4 *>(m)[i];
        float4 vv = reinterpret_cast<float4 *>(v)[i];
        float *P = &pp.x, *G = &gg.x, *M = &mm.x, *V = &vv.x;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float gr = G[q] * grad_scale;
            P[q] *= decay;
            M[q] = M[q] + (1.f - beta1) * (gr - M[q]);
            V[q] = V[q] * beta2 + (1.f - beta2) * gr * gr;
            const float denom = sqrtf(V[q]) / bc2s + eps;
            P[q] -= step_size * (M[q] / denom);
        }
        reinterpret_cast<float4 *>(p)[i] = pp;
        if (p16) {                                    // bf16 shadow of the parameters: the tensor-core layers read it as it is
            uint2 u;
            u.x = pack_bf16(pp.x, pp.y);
            u.y = pack_bf16(pp.z, pp.w);
            reinterpret_cast<uint2 *>(p16)[i] = u;
        }
        reinterpret_cast<float4 *>(m)[i] = mm;
        reinterpret_cast<float4 *>(v)[i] = vv;
    }
    for (int64_t i = (n4 << 2) + tid; i < n; i += nthreads) {
        const float gr = g[i] * grad_scale;
        float pv = p[i] * decay;
        const float mv = m[i] + (1.f - beta1) * (gr - m[i]);
        const float vv = v[i] * beta2 + (1.f - beta2) * gr * gr;
        pv -= step_size * (mv / (sqrtf(vv) / bc2s + eps));
        p[i] = pv; m[i] = mv; v[i] = vv;
        if (p16) p16[i] = __float2bfloat16_rn(pv);
    }
}

}  // namespace kdf

using namespace kdf;

extern "C" {

int kdf_confusion_matrix(const void *logits, const int64_t *labels, int B, int K, int64_t HW,
                         int dtype_logits, int64_t ignore_index, int64_t *conf, void *stream) {
    KDF_CHECK_ARG(B >= 0 && HW >= 0, "confusion: bad sizes");
    KDF_CHECK_ARG(K >= 1 && K <= CM_MAX_K, "confusion: K=%d outside [1,%d]", K, CM_MAX_K);
    KDF_CHECK_ARG(logits && labels && conf, "confusion: null pointer");
    KDF_CHECK_ARG(dtype_logits == KDF_F32 || dtype_logits == KDF_BF16, "confusion: bad dtype");
    const int64_t npix = (int64_t)B * HW;
    if (npix == 0) return KDF_OK;
    int64_t blocks = (npix + 1023) / 1024;
    if (blocks > sm_count() * 4) blocks = sm_count() * 4;
    cudaStream_t st = as_stream(stream);
    if (dtype_logits == KDF_F32)
        confusion_kernel<float><<<(int)blocks, 256, 0, st>>>(reinterpret_cast<const float *>(logits), labels, B, K, HW,
                                                           ignore_index, reinterpret_cast<unsigned long long *>(conf));
    else
        confusion_kernel<__nv_bfloat16><<<(int)blocks, 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16 *>(logits), labels,
                                                                   B, K, HW, ignore_index,
                                                                   reinterpret_cast<unsigned long long *>(conf));
    KDF_LAUNCH_CHECK();
    return KDF_OK;
}

int kdf_adamw_flat(float *param, const float *grad, float *exp_avg, float *exp_avg_sq, int64_t n,
                   const float *hyper, float beta1, float beta2, float eps, float weight_decay,
                   float grad_scale, void *param_bf16, void *stream) {
    KDF_CHECK_ARG(n >= 0, "adamw: negative size");
    KDF_CHECK_ARG((reinterpret_cast<uintptr_t>(param_bf16) & 7) == 0, "adamw: the bf16 shadow must be 8-byte aligned");
    KDF_CHECK_ARG(param && grad && exp_avg && exp_avg_sq && hyper, "adamw: null pointer");
    const uintptr_t al = reinterpret_cast<uintptr_t>(param) | reinterpret_cast<uintptr_t>(grad) |
                         reinterpret_cast<uintptr_t>(exp_avg) | reinterpret_cast<uintptr_t>(exp_avg_sq);
    KDF_CHECK_ARG((al & 15) == 0, "adamw: buffers must be 16-byte aligned");
    if (n == 0) return KDF_OK;
    int64_t blocks = (n / 4 + 255) / 256;
    if (blocks > sm_count() * 8) blocks = sm_count() * 8;
    if (blocks < 1) blocks = 1;
    adamw_flat_kernel<<<(int)blocks, 256, 0, as_stream(stream)>>>(param, grad, exp_avg, exp_avg_sq, n, hyper, beta1,
                                                                beta2, eps, weight_decay, grad_scale,
                                                                static_cast<__nv_bfloat16 *>(param_bf16));
    KDF_LAUNCH_CHECK();
    return KDF_OK;
}

}  // extern "C"
