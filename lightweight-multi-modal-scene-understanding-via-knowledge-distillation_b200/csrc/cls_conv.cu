// The classifier of the BEV-resolution head: nn.Conv2d(32, num_classes, 1) with bias over bf16 pixel rows (reference
// src/models/fusion_module.py:162-173), written as the planar [B, K, H, W] logits the loss reads.
//
// K = 2 output channels is no library shape: the implicit-GEMM path spends 24 us per model on 8 MB, 50 us on the backward,
// and the loss then copies the channels-last logits into planes.  Here a thread owns a pixel: 64 bytes in, K dot products of
// 32 against taps held in registers (rounded to bf16 as the autocast convolution does), K coalesced 2-byte stores into the
// planes.  The backward is one kernel: data gradient rows, and per-thread partial sums of the weight / bias gradients over
// a persistent grid, reduced by warp shuffles and one round of atomics.
#include <stdlib.h>

#include "kdf_common.cuh"

namespace kdf {

constexpr int CLS_C = 32;           // input channels

__device__ __forceinline__ void cls_load_row(const __nv_bfloat16 *p, float (&v)[CLS_C]) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const uint4 u = __ldg(reinterpret_cast<const uint4 *>(p) + q);
        v[8 * q + 0] = bf16_lo(u.x); v[8 * q + 1] = bf16_hi(u.x); v[8 * q + 2] = bf16_lo(u.y); v[8 * q + 3] = bf16_hi(u.y);
        v[8 * q + 4] = bf16_lo(u.z); v[8 * q + 5] = bf16_hi(u.z); v[8 * q + 6] = bf16_lo(u.w); v[8 * q + 7] = bf16_hi(u.w);
    }
}
__device__ __forceinline__ float cls_bf16r(float v) { return bf16_lo(pack_bf16(v, v)); }

template <int K>
__global__ void __launch_bounds__(256)
cls_conv_fwd_kernel(const __nv_bfloat16 *__restrict__ x, const float *__restrict__ w /* [K][32] */, const float *__restrict__ bias,
                    __nv_bfloat16 *__restrict__ out /* [B][K][HW] */, int64_t M, int HW) {
    float wr[K][CLS_C], br[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
        br[k] = bias ? __ldg(bias + k) : 0.f;
#pragma unroll
        for (int c = 0; c < CLS_C; ++c) wr[k][c] = cls_bf16r(__ldg(w + k * CLS_C + c));
    }
    for (int64_t m = (int64_t)blockIdx.x * 256 + threadIdx.x; m < M; m += (int64_t)gridDim.x * 256) {
        float v[CLS_C];
        cls_load_row(x + m * CLS_C, v);
        const int64_t b = m / HW, p = m - b * HW;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            float a0 = 0.f, a1 = 0.f;                       // two chains
#pragma unroll
            for (int c = 0; c < CLS_C; c += 2) { a0 = fmaf(wr[k][c], v[c], a0); a1 = fmaf(wr[k][c + 1], v[c + 1], a1); }
            out[(b * K + k) * HW + p] = __float2bfloat16_rn(a0 + a1 + br[k]);
        }
    }
}

template <int K>
__global__ void __launch_bounds__(256)
cls_conv_bwd_kernel(const __nv_bfloat16 *__restrict__ x, const __nv_bfloat16 *__restrict__ dl /* [B][K][HW] */,
                    const float *__restrict__ w, __nv_bfloat16 *__restrict__ dx /* nullable [M][32] */,
                    float *__restrict__ dW /* [K][32] */, float *__restrict__ db /* [K] */, int64_t M, int HW) {
    __shared__ float red[8][K * CLS_C + K];
    float wr[K][CLS_C], aw[K][CLS_C], ab[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
        ab[k] = 0.f;
#pragma unroll
        for (int c = 0; c < CLS_C; ++c) { wr[k][c] = cls_bf16r(__ldg(w + k * CLS_C + c)); aw[k][c] = 0.f; }
    }
    for (int64_t m = (int64_t)blockIdx.x * 256 + threadIdx.x; m < M; m += (int64_t)gridDim.x * 256) {
        float v[CLS_C], g[K];
        cls_load_row(x + m * CLS_C, v);
        const int64_t b = m / HW, p = m - b * HW;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            g[k] = __bfloat162float(dl[(b * K + k) * HW + p]);
            ab[k] += g[k];
#pragma unroll
            for (int c = 0; c < CLS_C; ++c) aw[k][c] = fmaf(g[k], v[c], aw[k][c]);
        }
        if (dx) {
            uint32_t o[CLS_C / 2];
#pragma unroll
            for (int c = 0; c < CLS_C; c += 2) {
                float d0 = 0.f, d1 = 0.f;
#pragma unroll
                for (int k = 0; k < K; ++k) { d0 = fmaf(g[k], wr[k][c], d0); d1 = fmaf(g[k], wr[k][c + 1], d1); }
                o[c / 2] = pack_bf16(d0, d1);
            }
#pragma unroll
            for (int q = 0; q < 4; ++q)
                reinterpret_cast<uint4 *>(dx + m * CLS_C)[q] = make_uint4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]);
        }
    }
    // warp shuffles, then the 8 warps through shared memory, then one atomic per gradient and CTA
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < K; ++k) {
#pragma unroll
        for (int c = 0; c < CLS_C; ++c) {
            float s = aw[k][c];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if (lane == 0) red[warp][k * CLS_C + c] = s;
        }
        float s = ab[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) red[warp][K * CLS_C + k] = s;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < K * CLS_C + K; i += 256) {
        float s = 0.f;
#pragma unroll
        for (int wi = 0; wi < 8; ++wi) s += red[wi][i];
        if (i < K * CLS_C) atomicAdd(dW + i, s);
        else if (db) atomicAdd(db + (i - K * CLS_C), s);
    }
}

}  // namespace kdf

using namespace kdf;

extern "C" {

static int cls_check(const char *who, int64_t M, int Cin, int K, int HW) {
    KDF_CHECK_ARG(M >= 0 && HW > 0 && M % HW == 0, "%s: M must be frames x H*W", who);
    KDF_CHECK_ARG(Cin == CLS_C, "%s: built for %d input channels (got %d)", who, CLS_C, Cin);
    KDF_CHECK_ARG(K >= 1 && K <= 4, "%s: 1..4 classes supported (got %d)", who, K);
    return KDF_OK;
}

int kdf_cls_conv_fwd(const void *x_bf16, const float *weight, const float *bias, int64_t M, int Cin, int K, int HW,
                     void *logits_bf16, void *stream) {
    if (int e = cls_check("cls_conv_fwd", M, Cin, K, HW)) return e;
    if (M == 0) return KDF_OK;
    KDF_CHECK_ARG(x_bf16 && weight && logits_bf16, "cls_conv_fwd: null pointer");
    KDF_CHECK_ARG((reinterpret_cast<uintptr_t>(x_bf16) & 15) == 0, "cls_conv_fwd: rows must be 16-byte aligned");
    int64_t blocks = (M + 255) / 256;
    if (blocks > (int64_t)sm_count() * 8) blocks = (int64_t)sm_count() * 8;
    cudaStream_t st = as_stream(stream);
    const __nv_bfloat16 *x = reinterpret_cast<const __nv_bfloat16 *>(x_bf16);
    __nv_bfloat16 *out = reinterpret_cast<__nv_bfloat16 *>(logits_bf16);
    switch (K) {
        case 1: cls_conv_fwd_kernel<1><<<(int)blocks, 256, 0, st>>>(x, weight, bias, out, M, HW); break;
        case 2: cls_conv_fwd_kernel<2><<<(int)blocks, 256, 0, st>>>(x, weight, bias, out, M, HW); break;
        case 3: cls_conv_fwd_kernel<3><<<(int)blocks, 256, 0, st>>>(x, weight, bias, out, M, HW); break;
        default: cls_conv_fwd_kernel<4><<<(int)blocks, 256, 0, st>>>(x, weight, bias, out, M, HW); break;
    }
    KDF_LAUNCH_CHECK();
    return KDF_OK;
}

int kdf_cls_conv_bwd(const void *x_bf16, const void *dlogits_bf16, const float *weight, int64_t M, int Cin, int K, int HW,
                     void *dx_bf16, float *grad_weight, float *grad_bias, void *stream) {
    if (int e = cls_check("cls_conv_bwd", M, Cin, K, HW)) return e;
    KDF_CHECK_ARG(grad_weight, "cls_conv_bwd: null pointer");
    cudaStream_t st = as_stream(stream);
    KDF_CUDA(cudaMemsetAsync(grad_weight, 0, sizeof(float) * K * CLS_C, st));
    if (grad_bias) KDF_CUDA(cudaMemsetAsync(grad_bias, 0, sizeof(float) * K, st));
    if (M == 0) return KDF_OK;
    KDF_CHECK_ARG(x_bf16 && dlogits_bf16 && weight, "cls_conv_bwd: null pointer");
    KDF_CHECK_ARG(((reinterpret_cast<uintptr_t>(x_bf16) | reinterpret_cast<uintptr_t>(dx_bf16)) & 15) == 0, "cls_conv_bwd: rows must be 16-byte aligned");
    int64_t blocks = (M + 255) / 256;
    if (blocks > (int64_t)sm_count() * 2) blocks = (int64_t)sm_count() * 2;        // persistent: per-CTA partial sums -> few atomics
    const __nv_bfloat16 *x = reinterpret_cast<const __nv_bfloat16 *>(x_bf16), *dl = reinterpret_cast<const __nv_bfloat16 *>(dlogits_bf16);
    __nv_bfloat16 *dx = reinterpret_cast<__nv_bfloat16 *>(dx_bf16);
    switch (K) {
        case 1: cls_conv_bwd_kernel<1><<<(int)blocks, 256, 0, st>>>(x, dl, weight, dx, grad_weight, grad_bias, M, HW); break;
        case 2: cls_conv_bwd_kernel<2><<<(int)blocks, 256, 0, st>>>(x, dl, weight, dx, grad_weight, grad_bias, M, HW); break;
        case 3: cls_conv_bwd_kernel<3><<<(int)blocks, 256, 0, st>>>(x, dl, weight, dx, grad_weight, grad_bias, M, HW); break;
        default: cls_conv_bwd_kernel<4><<<(int)blocks, 256, 0, st>>>(x, dl, weight, dx, grad_weight, grad_bias, M, HW); break;
    }
    KDF_LAUNCH_CHECK();
    return KDF_OK;
}

}  // extern "C"
