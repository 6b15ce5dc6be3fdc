// Knowledge-distillation loss, forward + backward in one pass (sm_100a).
//
//   ce  : nn.CrossEntropyLoss(ignore_index, weight)      reference src/training/trainer.py:55,88
//   kl  : T^2 * KL(softmax(z_t/T) || softmax(z_s/T)), per-pixel mean     SURVEY.md 8c (not in the reference)
//   mse : feature-mimic MSE over up to two taps                         SURVEY.md 8c (not in the reference)
//
// A tiny pre-kernel histograms the valid labels (giving the CE normaliser
// sum_i w[y_i] that every CE gradient needs); the main kernel then streams the logits
// and the mimic taps exactly once, writes d_logits / d_feats in the same pass
// and reduces the loss terms with warp shuffles -> per-CTA partials -> a fixed
// order final sum by the last CTA (bitwise deterministic).
// HBM-bound: K*s*3 + 8 bytes per pixel plus 3*s bytes per tap element.
#include <stddef.h>
#include <stdlib.h>

#include "kdf_common.cuh"

namespace kdf {

constexpr int KD_MAX_K = 8;
constexpr int KD_THREADS = 256;
constexpr int KD_MAX_BLOCKS = 2048;
constexpr int KD_NACC = 6;   // ce_num, kl, tap0, tap1, n_valid, (unused)

struct KdWorkspace {
    unsigned int class_count[KD_MAX_K];   // valid labels per class (integer -> deterministic)
    unsigned int done;                    // CTA completion ticket
    unsigned int pad[7];
    float partial[KD_MAX_BLOCKS][KD_NACC];
};

struct KdParams {
    const void *zs, *zt;
    const int64_t *labels;
    const float *cw;
    int B, K;
    int64_t HW;
    float T, alpha, beta, grad_scale;
    int64_t ignore_index;
    const void *s0, *t0, *s1, *t1;
    void *d0, *d1;
    int64_t n0, n1;
    void *dz;
    float *scalars;
    KdWorkspace *ws;
};

// Histogram of the valid labels (the header of the workspace is zeroed first).
// The CE normaliser is then sum_k w[k]*count[k]: integer counting keeps it
// independent of the summation order.
__global__ void __launch_bounds__(KD_THREADS)
kd_label_count_kernel(const int64_t *__restrict__ labels, int K, int64_t n, int64_t ignore_index,
                      KdWorkspace *ws) {
    __shared__ unsigned int hist[KD_MAX_K];
    if (threadIdx.x < KD_MAX_K) hist[threadIdx.x] = 0u;
    __syncthreads();
    const int64_t nthreads = (int64_t)gridDim.x * KD_THREADS;
    for (int64_t i = (int64_t)blockIdx.x * KD_THREADS + threadIdx.x; i < n; i += nthreads) {
        const int64_t y = labels[i];
        if (y != ignore_index && y >= 0 && y < K) atomicAdd(&hist[y], 1u);
    }
    __syncthreads();
    if (threadIdx.x < K && hist[threadIdx.x]) atomicAdd(&ws->class_count[threadIdx.x], hist[threadIdx.x]);
}

__device__ __forceinline__ void kd_normaliser(const KdWorkspace *ws, const float *cw, int K,
                                              float &wsum, float &n_valid) {
    wsum = 0.f; n_valid = 0.f;
    for (int k = 0; k < K; ++k) {
        const float c = (float)__ldcg(&ws->class_count[k]);
        wsum += (cw ? cw[k] : 1.f) * c;
        n_valid += c;
    }
}

// 16 bytes of tap elements <-> floats
template <typename TF> struct TapVec;
template <> struct TapVec<float> {
    static constexpr int N = 4;
    static __device__ __forceinline__ void unpack(const uint4 &u, float *v) {
        v[0] = __uint_as_float(u.x); v[1] = __uint_as_float(u.y); v[2] = __uint_as_float(u.z); v[3] = __uint_as_float(u.w);
    }
    static __device__ __forceinline__ uint4 pack(const float *v) {
        return make_uint4(__float_as_uint(v[0]), __float_as_uint(v[1]), __float_as_uint(v[2]), __float_as_uint(v[3]));
    }
};
template <> struct TapVec<__nv_bfloat16> {
    static constexpr int N = 8;
    static __device__ __forceinline__ void unpack(const uint4 &u, float *v) {
        v[0] = bf16_lo(u.x); v[1] = bf16_hi(u.x); v[2] = bf16_lo(u.y); v[3] = bf16_hi(u.y);
        v[4] = bf16_lo(u.z); v[5] = bf16_hi(u.z); v[6] = bf16_lo(u.w); v[7] = bf16_hi(u.w);
    }
    static __device__ __forceinline__ uint4 pack(const float *v) {
        return make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
    }
};

// One mimic tap: sum (s-t)^2 and d = coef*(s-t), 16-byte accesses, U independent load pairs in flight per thread.
template <typename TF>
__device__ __forceinline__ float kd_tap(const TF *__restrict__ s, const TF *__restrict__ t, TF *__restrict__ d,
                                        int64_t n, float coef, int64_t tid, int64_t nthreads) {
    // coef = grad_scale * beta * 2 / n
    constexpr int V = TapVec<TF>::N, U = 4;
    float acc = 0.f;
    const int64_t nv = n / V;
    const uint4 *s4 = reinterpret_cast<const uint4 *>(s), *t4 = reinterpret_cast<const uint4 *>(t);
    uint4 *d4 = reinterpret_cast<uint4 *>(d);
    for (int64_t i = tid; i < nv; i += nthreads * U) {
        uint4 a[U], b[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t j = i + (int64_t)u * nthreads;
            if (j < nv) { a[u] = ldg_stream_u4(s4 + j); b[u] = ldg_stream_u4(t4 + j); }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t j = i + (int64_t)u * nthreads;
            if (j < nv) {
                float x[V], y[V];
                TapVec<TF>::unpack(a[u], x);
                TapVec<TF>::unpack(b[u], y);
#pragma unroll
                for (int q = 0; q < V; ++q) {
                    const float e = x[q] - y[q];
                    acc = fmaf(e, e, acc);
                    x[q] = coef * e;
                }
                d4[j] = TapVec<TF>::pack(x);
            }
        }
    }
    for (int64_t j = nv * V + tid; j < n; j += nthreads) {        // tail (numel % V)
        const float e = to_float<TF>(s[j]) - to_float<TF>(t[j]);
        acc += e * e;
        d[j] = from_float<TF>(coef * e);
    }
    return acc;
}

template <typename TL, typename TF>
__global__ void __launch_bounds__(KD_THREADS, 3)
kd_loss_kernel(KdParams p) {
    __shared__ float red[KD_THREADS / 32];
    __shared__ bool is_last;
    const int64_t tid = (int64_t)blockIdx.x * KD_THREADS + threadIdx.x;
    const int64_t nthreads = (int64_t)gridDim.x * KD_THREADS;
    const TL *zs = reinterpret_cast<const TL *>(p.zs);
    const TL *zt = reinterpret_cast<const TL *>(p.zt);
    TL *dz = reinterpret_cast<TL *>(p.dz);
    const int K = p.K;
    const int64_t HW = p.HW, npix = (int64_t)p.B * HW;
    float m0 = 0.f, m1 = 0.f;
    float wsum, n_valid;
    kd_normaliser(p.ws, p.cw, K, wsum, n_valid);
    const float inv_wsum = 1.f / wsum;                       // 0/0 -> NaN like torch when nothing is valid
    const float invT = 1.f / p.T;
    const float ce_coef = p.grad_scale * (1.f - p.alpha) * inv_wsum;
    const float kl_coef = (zt != nullptr) ? p.grad_scale * p.alpha * p.T / (float)npix : 0.f;

    float ce_num = 0.f, kl = 0.f;
    for (int64_t i = tid; i < npix; i += nthreads) {
        const int64_t b = i / HW, hw = i - b * HW;
        const int64_t base = b * K * HW + hw;
        float s[KD_MAX_K], t[KD_MAX_K];
        float smax = -INFINITY;
#pragma unroll
        for (int k = 0; k < KD_MAX_K; ++k)
            if (k < K) { s[k] = to_float<TL>(zs[base + k * HW]); smax = fmaxf(smax, s[k]); }
        // --- cross entropy (T = 1)
        float se = 0.f;
#pragma unroll
        for (int k = 0; k < KD_MAX_K; ++k) if (k < K) se += expf(s[k] - smax);
        const float lse = smax + logf(se);
        const int64_t y = p.labels[i];
        const bool valid = (y != p.ignore_index) && y >= 0 && y < K;
        const float w = valid ? (p.cw ? p.cw[y] : 1.f) : 0.f;
        float g[KD_MAX_K];
#pragma unroll
        for (int k = 0; k < KD_MAX_K; ++k)
            if (k < K) {
                const float pk = expf(s[k] - lse);
                g[k] = ce_coef * w * (pk - ((int64_t)k == y ? 1.f : 0.f));
                if ((int64_t)k == y && valid) ce_num += w * (lse - s[k]);
            }
        // --- temperature-scaled KL(teacher || student)
        if (zt != nullptr) {
            float tmax = -INFINITY;
#pragma unroll
            for (int k = 0; k < KD_MAX_K; ++k)
                if (k < K) { t[k] = to_float<TL>(zt[base + k * HW]); tmax = fmaxf(tmax, t[k]); }
            float ses = 0.f, set = 0.f;
#pragma unroll
            for (int k = 0; k < KD_MAX_K; ++k)
                if (k < K) { ses += expf((s[k] - smax) * invT); set += expf((t[k] - tmax) * invT); }
            const float lses = logf(ses), lset = logf(set);
#pragma unroll
            for (int k = 0; k < KD_MAX_K; ++k)
                if (k < K) {
                    const float lps = (s[k] - smax) * invT - lses;     // log softmax(z_s/T)
                    const float lpt = (t[k] - tmax) * invT - lset;     // log softmax(z_t/T)
                    const float pt = expf(lpt);
                    kl += pt * (lpt - lps);
                    g[k] += kl_coef * (expf(lps) - pt);
                }
        }
#pragma unroll
        for (int k = 0; k < KD_MAX_K; ++k)
            if (k < K) dz[base + k * HW] = from_float<TL>(g[k]);
    }

    // --- feature-mimic taps
    if (p.n0 > 0)
        m0 = kd_tap<TF>(reinterpret_cast<const TF *>(p.s0), reinterpret_cast<const TF *>(p.t0),
                        reinterpret_cast<TF *>(p.d0), p.n0, p.grad_scale * p.beta * 2.f / (float)p.n0, tid, nthreads);
    if (p.n1 > 0)
        m1 = kd_tap<TF>(reinterpret_cast<const TF *>(p.s1), reinterpret_cast<const TF *>(p.t1),
                        reinterpret_cast<TF *>(p.d1), p.n1, p.grad_scale * p.beta * 2.f / (float)p.n1, tid, nthreads);

    // --- reduction: warp shuffles -> CTA partial -> last CTA sums in block order
    ce_num = block_sum<KD_THREADS>(ce_num, red);
    kl = block_sum<KD_THREADS>(kl, red);
    m0 = block_sum<KD_THREADS>(m0, red);
    m1 = block_sum<KD_THREADS>(m1, red);
    if (threadIdx.x == 0) {
        float *pp = p.ws->partial[blockIdx.x];
        pp[0] = ce_num; pp[1] = kl; pp[2] = m0; pp[3] = m1;
        __threadfence();
        const unsigned ticket = atomicAdd(&p.ws->done, 1u);
        is_last = (ticket == gridDim.x - 1);
    }
    __syncthreads();
    if (is_last) {
        __threadfence();
        float a[4] = {0.f, 0.f, 0.f, 0.f};
        // fixed order: thread j sums blocks j, j+256, ...; then a block reduction
        for (int blk = threadIdx.x; blk < (int)gridDim.x; blk += KD_THREADS) {
            const volatile float *pp = p.ws->partial[blk];
#pragma unroll
            for (int q = 0; q < 4; ++q) a[q] += pp[q];
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) a[q] = block_sum<KD_THREADS>(a[q], red);
        if (threadIdx.x == 0) {
            const float ce = a[0] * inv_wsum;
            const float klv = (zt != nullptr) ? p.T * p.T * a[1] / (float)npix : 0.f;
            const float t0 = p.n0 > 0 ? a[2] / (float)p.n0 : 0.f;
            const float t1 = p.n1 > 0 ? a[3] / (float)p.n1 : 0.f;
            const float mse = t0 + t1;
            p.scalars[0] = (1.f - p.alpha) * ce + p.alpha * klv + p.beta * mse;
            p.scalars[1] = ce;
            p.scalars[2] = klv;
            p.scalars[3] = mse;
            p.scalars[4] = wsum;
            p.scalars[5] = t0;
            p.scalars[6] = t1;
            p.scalars[7] = n_valid;
        }
    }
}

}  // namespace kdf

using namespace kdf;

extern "C" {

size_t kdf_kd_loss_workspace_bytes(void) { return sizeof(KdWorkspace); }

int kdf_kd_label_count(const int64_t *labels, int B, int K, int64_t HW, int64_t ignore_index, void *workspace, void *stream) {
    KDF_CHECK_ARG(B > 0 && HW > 0, "kd_label_count: empty batch");
    KDF_CHECK_ARG(K >= 1 && K <= KD_MAX_K, "kd_label_count: K=%d outside [1,%d]", K, KD_MAX_K);
    KDF_CHECK_ARG(labels && workspace, "kd_label_count: null pointer");
    cudaStream_t st = as_stream(stream);
    KdWorkspace *ws = reinterpret_cast<KdWorkspace *>(workspace);
    const int64_t npix = (int64_t)B * HW;
    KDF_CUDA(cudaMemsetAsync(ws, 0, offsetof(KdWorkspace, partial), st));
    int64_t nb = (npix + KD_THREADS * 4 - 1) / (KD_THREADS * 4);
    if (nb > sm_count() * 4) nb = sm_count() * 4;
    kd_label_count_kernel<<<(int)nb, KD_THREADS, 0, st>>>(labels, K, npix, ignore_index, ws);
    KDF_LAUNCH_CHECK();
    return KDF_OK;
}

static int kd_loss_impl(const void *s_logits, const void *t_logits, const int64_t *labels,
                        const float *class_w, int B, int K, int64_t HW, int dtype_logits,
                        float T, float alpha, float beta, int64_t ignore_index,
                        const void *s_feat0, const void *t_feat0, void *d_feat0, int64_t numel0,
                        const void *s_feat1, const void *t_feat1, void *d_feat1, int64_t numel1,
                        int dtype_feat, float grad_scale,
                        void *d_logits, float *scalars, void *workspace, void *stream, bool counts_ready) {
    KDF_CHECK_ARG(B > 0 && HW > 0, "kd_loss: empty batch");
    KDF_CHECK_ARG(K >= 1 && K <= KD_MAX_K, "kd_loss: K=%d outside [1,%d]", K, KD_MAX_K);
    KDF_CHECK_ARG(s_logits && labels && d_logits && scalars && workspace, "kd_loss: null pointer");
    KDF_CHECK_ARG(T > 0.f, "kd_loss: temperature must be positive");
    KDF_CHECK_ARG(dtype_logits == KDF_F32 || dtype_logits == KDF_BF16, "kd_loss: bad logits dtype");
    KDF_CHECK_ARG(dtype_feat == KDF_F32 || dtype_feat == KDF_BF16, "kd_loss: bad feature dtype");
    if (numel0 > 0) KDF_CHECK_ARG(s_feat0 && t_feat0 && d_feat0, "kd_loss: tap 0 null pointer");
    if (numel1 > 0) KDF_CHECK_ARG(s_feat1 && t_feat1 && d_feat1, "kd_loss: tap 1 null pointer");
    const uintptr_t al = reinterpret_cast<uintptr_t>(s_feat0) | reinterpret_cast<uintptr_t>(t_feat0) |
                         reinterpret_cast<uintptr_t>(d_feat0) | reinterpret_cast<uintptr_t>(s_feat1) |
                         reinterpret_cast<uintptr_t>(t_feat1) | reinterpret_cast<uintptr_t>(d_feat1);
    KDF_CHECK_ARG((al & 15) == 0, "kd_loss: feature taps must be 16-byte aligned");
    cudaStream_t st = as_stream(stream);
    KdWorkspace *ws = reinterpret_cast<KdWorkspace *>(workspace);
    const int64_t npix = (int64_t)B * HW;
    if (!counts_ready) {
        if (int e = kdf_kd_label_count(labels, B, K, HW, ignore_index, workspace, stream)) return e;
    }

    KdParams p;
    p.zs = s_logits; p.zt = t_logits; p.labels = labels; p.cw = class_w;
    p.B = B; p.K = K; p.HW = HW; p.T = T; p.alpha = alpha; p.beta = beta; p.grad_scale = grad_scale;
    p.ignore_index = ignore_index;
    p.s0 = s_feat0; p.t0 = t_feat0; p.d0 = d_feat0; p.n0 = numel0 > 0 ? numel0 : 0;
    p.s1 = s_feat1; p.t1 = t_feat1; p.d1 = d_feat1; p.n1 = numel1 > 0 ? numel1 : 0;
    p.dz = d_logits; p.scalars = scalars; p.ws = ws;

    // persistent grid: exactly the resident CTAs (3 per SM, one wave), fewer when the work is small
    const int64_t work = npix + (p.n0 + p.n1) / 16;
    int blocks = sm_count() * 3;
    const int64_t need = (work + KD_THREADS - 1) / KD_THREADS;
    if (need < blocks) blocks = (int)(need < 1 ? 1 : need);
    if (blocks > KD_MAX_BLOCKS) blocks = KD_MAX_BLOCKS;
    if (dtype_logits == KDF_F32 && dtype_feat == KDF_F32)        kd_loss_kernel<float, float><<<blocks, KD_THREADS, 0, st>>>(p);
    else if (dtype_logits == KDF_F32 && dtype_feat == KDF_BF16)  kd_loss_kernel<float, __nv_bfloat16><<<blocks, KD_THREADS, 0, st>>>(p);
    else if (dtype_logits == KDF_BF16 && dtype_feat == KDF_F32)  kd_loss_kernel<__nv_bfloat16, float><<<blocks, KD_THREADS, 0, st>>>(p);
    else                                                         kd_loss_kernel<__nv_bfloat16, __nv_bfloat16><<<blocks, KD_THREADS, 0, st>>>(p);
    KDF_LAUNCH_CHECK();
    return KDF_OK;
}

int kdf_kd_loss_fwd_bwd(const void *s_logits, const void *t_logits, const int64_t *labels,
                        const float *class_w, int B, int K, int64_t HW, int dtype_logits,
                        float T, float alpha, float beta, int64_t ignore_index,
                        const void *s_feat0, const void *t_feat0, void *d_feat0, int64_t numel0,
                        const void *s_feat1, const void *t_feat1, void *d_feat1, int64_t numel1,
                        int dtype_feat, float grad_scale,
                        void *d_logits, float *scalars, void *workspace, void *stream) {
    return kd_loss_impl(s_logits, t_logits, labels, class_w, B, K, HW, dtype_logits, T, alpha, beta, ignore_index, s_feat0, t_feat0,
                        d_feat0, numel0, s_feat1, t_feat1, d_feat1, numel1, dtype_feat, grad_scale, d_logits, scalars, workspace,
                        stream, false);
}

int kdf_kd_loss_fwd_bwd_counted(const void *s_logits, const void *t_logits, const int64_t *labels,
                                const float *class_w, int B, int K, int64_t HW, int dtype_logits,
                                float T, float alpha, float beta, int64_t ignore_index,
                                const void *s_feat0, const void *t_feat0, void *d_feat0, int64_t numel0,
                                const void *s_feat1, const void *t_feat1, void *d_feat1, int64_t numel1,
                                int dtype_feat, float grad_scale,
                                void *d_logits, float *scalars, void *workspace, void *stream) {
    return kd_loss_impl(s_logits, t_logits, labels, class_w, B, K, HW, dtype_logits, T, alpha, beta, ignore_index, s_feat0, t_feat0,
                        d_feat0, numel0, s_feat1, t_feat1, d_feat1, numel1, dtype_feat, grad_scale, d_logits, scalars, workspace,
                        stream, true);
}

}  // extern "C"
