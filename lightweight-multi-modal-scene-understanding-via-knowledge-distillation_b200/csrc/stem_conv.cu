// The camera stem: 3x3 convolution, stride 2, padding 1, 3 -> 32 channels, no bias (reference
// src/models/camera_encoder.py:63-67), straight from the fp32 NCHW image the loader delivers to bf16 pixel-major rows.
//
// Through the library this layer costs four passes per model and step -- image -> channels-last copy, fp32 -> bf16 copy,
// a channel-padding pass and the implicit-GEMM kernel (0.12 ms per model at 32 x 256 x 256 frames, 0.09 ms more for its
// weight gradient) -- for 59 MB of traffic.  K = 27 is no tensor-core shape; it is 864 multiply-adds per output pixel on
// the CUDA cores:
//   * a CTA owns a tile of ST_R output rows x ST_XT output columns; its input patch (3 planes x 2*ST_R+1 rows x
//     2*ST_XT+1 columns, zero padded) is staged in shared memory ONCE, rounded to bf16 like the autocast convolution
//     does, as duplicated fp32 pairs (v, v): a thread's `fma.rn.f32x2` then multiplies one input by the taps of its TWO
//     output channels with no per-use unpacking;
//   * a thread owns a channel pair (27 packed taps in registers, bf16-rounded) and 4 adjacent output pixels: 9 shared
//     loads feed 12 packed FMAs per (input plane, tap row); lanes 0-15 / 16-31 of a warp are the 16 channel pairs of two
//     pixel groups, so the loads are broadcasts and an output pixel's 32 channels leave as one 64-byte store;
//   * persistent CTAs; epilogue = the statistics of the BatchNorm that follows (training) or its folded form +
//     ReLU6 (inference), exactly as in the depthwise kernels.
// The weight gradient walks the same tiles with the roles swapped (27 packed accumulators per thread).
#include <stdlib.h>

#include "kdf_common.cuh"

namespace kdf {

constexpr int ST_CO = 32;          // output channels
constexpr int ST_CI = 3;           // image planes
constexpr int ST_P = 4;            // output pixels per thread (along x)
constexpr int ST_XT = 64;          // output columns per tile = 16 pixel groups
constexpr int ST_R = 4;            // output rows per tile
constexpr int ST_IC = 2 * ST_XT + 1, ST_IR = 2 * ST_R + 1;
constexpr int ST_ICP = ST_IC + 1;                         // row pitch in pairs: even, so that a pixel group's 9 pairs start 16-byte aligned
constexpr int ST_TILE = ST_CI * ST_IR * ST_ICP;           // staged values per tile (pairs of 8 bytes)
constexpr int ST_TAPS = ST_CI * 9;

typedef unsigned long long u64;

__device__ __forceinline__ u64 st_pk2(float lo, float hi) {
    u64 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void st_upk2(u64 v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ u64 st_fma2(u64 a, u64 b, u64 c) {
    u64 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ u64 st_add2(u64 a, u64 b) {
    u64 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ float st_bf16r(float v) { return bf16_lo(pack_bf16(v, v)); }      // round to bf16, back to fp32

struct StemTile {
    int b, oy0, ox0;
};
__device__ __forceinline__ StemTile stem_tile(int t, int n_xt, int n_rb) {
    StemTile s;
    const int xt = t % n_xt, rest = t / n_xt;
    s.ox0 = xt * ST_XT;
    s.oy0 = (rest % n_rb) * ST_R;
    s.b = rest / n_rb;
    return s;
}

// the 9 pairs of a pixel group in one row of the staged patch: four 16-byte loads + one 8-byte load
__device__ __forceinline__ void stem_row9(const u64 *row, u64 (&d)[2 * ST_P + 1]) {
#pragma unroll
    for (int i = 0; i < ST_P; ++i) {
        const ulonglong2 v = *reinterpret_cast<const ulonglong2 *>(row + 2 * i);
        d[2 * i] = v.x;
        d[2 * i + 1] = v.y;
    }
    d[2 * ST_P] = row[2 * ST_P];
}

// The tile's input patch: fetched into registers one tile AHEAD (the loads fly while the current tile is computed), then
// written to shared memory as (v, v) pairs of the bf16-rounded value; zero outside the image.
constexpr int ST_PRE = (ST_TILE + 255) / 256;             // staged values per thread
__device__ __forceinline__ void stem_fetch(const float *__restrict__ img, float (&pre)[ST_PRE], const StemTile &s, int H, int W) {
    const int iy0 = 2 * s.oy0 - 1, ix0 = 2 * s.ox0 - 1;
#pragma unroll
    for (int k = 0; k < ST_PRE; ++k) {
        const int i = threadIdx.x + k * 256;
        const int c = i % ST_ICP, rest = i / ST_ICP, r = rest % ST_IR, ci = rest / ST_IR;
        const int iy = iy0 + r, ix = ix0 + c;
        pre[k] = 0.f;
        if (i < ST_TILE && (unsigned)iy < (unsigned)H && (unsigned)ix < (unsigned)W)
            pre[k] = __ldg(img + (((int64_t)s.b * ST_CI + ci) * H + iy) * W + ix);
    }
}
__device__ __forceinline__ void stem_store(u64 *tile, const float (&pre)[ST_PRE]) {
#pragma unroll
    for (int k = 0; k < ST_PRE; ++k) {
        const int i = threadIdx.x + k * 256;
        const float v = st_bf16r(pre[k]);
        if (i < ST_TILE) tile[i] = st_pk2(v, v);
    }
}

enum { ST_PLAIN = 0, ST_STATS = 1, ST_POST = 2 };

template <int MODE>
__global__ void __launch_bounds__(256, 2)
stem_conv_fwd_kernel(const float *__restrict__ img, const float *__restrict__ w /* [32][3][3][3] */, __nv_bfloat16 *__restrict__ out,
                     int B, int H, int W, int OH, int OW, double *__restrict__ stats /* [2][32] */,
                     const float *__restrict__ post_scale, const float *__restrict__ post_shift, int post_act) {
    __shared__ __align__(16) u64 tile[ST_TILE];
    __shared__ float sred[MODE == ST_STATS ? 256 * 4 : 1];
    const int tid = threadIdx.x, cp = tid & 15, pg = tid >> 4;
    u64 wr[ST_TAPS];
#pragma unroll
    for (int k = 0; k < ST_TAPS; ++k) wr[k] = st_pk2(st_bf16r(__ldg(w + (2 * cp) * ST_TAPS + k)), st_bf16r(__ldg(w + (2 * cp + 1) * ST_TAPS + k)));
    float psc0 = 1.f, psc1 = 1.f, psh0 = 0.f, psh1 = 0.f;
    if (MODE == ST_POST) {
        psc0 = __ldg(post_scale + 2 * cp); psc1 = __ldg(post_scale + 2 * cp + 1);
        psh0 = __ldg(post_shift + 2 * cp); psh1 = __ldg(post_shift + 2 * cp + 1);
    }
    u64 s_sum = 0ull, s_sq = 0ull;
    const int n_xt = (OW + ST_XT - 1) / ST_XT, n_rb = (OH + ST_R - 1) / ST_R;
    const int n_tiles = B * n_rb * n_xt;
    float pre[ST_PRE];
    if ((int)blockIdx.x < n_tiles) stem_fetch(img, pre, stem_tile(blockIdx.x, n_xt, n_rb), H, W);
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const StemTile s = stem_tile(t, n_xt, n_rb);
        __syncthreads();                                   // the previous tile has been consumed
        stem_store(tile, pre);
        __syncthreads();
        if (t + (int)gridDim.x < n_tiles) stem_fetch(img, pre, stem_tile(t + gridDim.x, n_xt, n_rb), H, W);
        const int ox = s.ox0 + ST_P * pg;
#pragma unroll 1
        for (int r = 0; r < ST_R; ++r) {
            const int oy = s.oy0 + r;
            if (oy >= OH) break;
            u64 acc[ST_P] = {0ull, 0ull, 0ull, 0ull};
#pragma unroll
            for (int ci = 0; ci < ST_CI; ++ci)
#pragma unroll
                for (int ky = 0; ky < 3; ++ky) {
                    u64 d[2 * ST_P + 1];
                    stem_row9(tile + (ci * ST_IR + 2 * r + ky) * ST_ICP + 2 * ST_P * pg, d);
#pragma unroll
                    for (int kx = 0; kx < 3; ++kx)
#pragma unroll
                        for (int px = 0; px < ST_P; ++px) acc[px] = st_fma2(wr[ci * 9 + ky * 3 + kx], d[2 * px + kx], acc[px]);
                }
            __nv_bfloat16 *dst = out + (((int64_t)s.b * OH + oy) * OW + ox) * ST_CO + 2 * cp;
#pragma unroll
            for (int px = 0; px < ST_P; ++px) {
                if (ox + px < OW) {
                    float a0, a1;
                    st_upk2(acc[px], a0, a1);
                    uint32_t o = pack_bf16(a0, a1);
                    if (MODE == ST_POST) {                 // on the value as it would have been stored
                        float y0 = fmaf(bf16_lo(o), psc0, psh0), y1 = fmaf(bf16_hi(o), psc1, psh1);
                        if (post_act == 1) { y0 = fmaxf(y0, 0.f); y1 = fmaxf(y1, 0.f); }
                        else if (post_act == 2) { y0 = fminf(fmaxf(y0, 0.f), 6.f); y1 = fminf(fmaxf(y1, 0.f), 6.f); }
                        o = pack_bf16(y0, y1);
                    }
                    *reinterpret_cast<uint32_t *>(dst + px * ST_CO) = o;
                    if (MODE == ST_STATS) {
                        const u64 v = st_pk2(bf16_lo(o), bf16_hi(o));
                        s_sum = st_add2(s_sum, v);
                        s_sq = st_fma2(v, v, s_sq);
                    }
                }
            }
        }
    }
    if (MODE == ST_STATS) {                                // 16 pixel groups per channel pair -> one fp64 atomic per channel and CTA
        st_upk2(s_sum, sred[tid * 4 + 0], sred[tid * 4 + 1]);
        st_upk2(s_sq, sred[tid * 4 + 2], sred[tid * 4 + 3]);
        __syncthreads();
        if (tid < 64) {
            const int c2 = tid >> 2, e = tid & 3;          // channel pair, (sum0, sum1, sq0, sq1)
            float v = 0.f;
#pragma unroll
            for (int g = 0; g < 16; ++g) v += sred[(g * 16 + c2) * 4 + e];
            atomicAdd(stats + (e >> 1) * ST_CO + 2 * c2 + (e & 1), (double)v);
        }
    }
}

// dW[co][ci][ky][kx] = sum over (b, oy, ox) of g(b,oy,ox,co) * bf16(img(b,ci,2oy+ky-1,2ox+kx-1))
__global__ void __launch_bounds__(256, 2)
stem_conv_wgrad_kernel(const float *__restrict__ img, const __nv_bfloat16 *__restrict__ gout, float *__restrict__ dw /* [32][27] */,
                       int B, int H, int W, int OH, int OW) {
    extern __shared__ __align__(16) uint8_t st_smem[];
    u64 *tile = reinterpret_cast<u64 *>(st_smem);
    const int tid = threadIdx.x, cp = tid & 15, pg = tid >> 4;
    u64 acc[ST_TAPS];
#pragma unroll
    for (int k = 0; k < ST_TAPS; ++k) acc[k] = 0ull;
    const int n_xt = (OW + ST_XT - 1) / ST_XT, n_rb = (OH + ST_R - 1) / ST_R;
    const int n_tiles = B * n_rb * n_xt;
    float pre[ST_PRE];
    if ((int)blockIdx.x < n_tiles) stem_fetch(img, pre, stem_tile(blockIdx.x, n_xt, n_rb), H, W);
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const StemTile s = stem_tile(t, n_xt, n_rb);
        __syncthreads();
        stem_store(tile, pre);
        __syncthreads();
        if (t + (int)gridDim.x < n_tiles) stem_fetch(img, pre, stem_tile(t + gridDim.x, n_xt, n_rb), H, W);
        const int ox = s.ox0 + ST_P * pg;
#pragma unroll 1
        for (int r = 0; r < ST_R; ++r) {
            const int oy = s.oy0 + r;
            if (oy >= OH) break;
            const __nv_bfloat16 *gp = gout + (((int64_t)s.b * OH + oy) * OW + ox) * ST_CO + 2 * cp;
            u64 g2[ST_P];
#pragma unroll
            for (int px = 0; px < ST_P; ++px) {
                const uint32_t u = (ox + px < OW) ? __ldg(reinterpret_cast<const uint32_t *>(gp + px * ST_CO)) : 0u;
                g2[px] = st_pk2(bf16_lo(u), bf16_hi(u));
            }
#pragma unroll
            for (int ci = 0; ci < ST_CI; ++ci)
#pragma unroll
                for (int ky = 0; ky < 3; ++ky) {
                    u64 d[2 * ST_P + 1];
                    stem_row9(tile + (ci * ST_IR + 2 * r + ky) * ST_ICP + 2 * ST_P * pg, d);
#pragma unroll
                    for (int kx = 0; kx < 3; ++kx)
#pragma unroll
                        for (int px = 0; px < ST_P; ++px) acc[ci * 9 + ky * 3 + kx] = st_fma2(g2[px], d[2 * px + kx], acc[ci * 9 + ky * 3 + kx]);
                }
        }
    }
    __syncthreads();                                       // the tile memory becomes the reduction scratch [pg][cp][27][2]
    float *red = reinterpret_cast<float *>(st_smem);
#pragma unroll
    for (int k = 0; k < ST_TAPS; ++k) st_upk2(acc[k], red[(tid * ST_TAPS + k) * 2], red[(tid * ST_TAPS + k) * 2 + 1]);
    __syncthreads();
    for (int i = tid; i < 16 * ST_TAPS * 2; i += 256) {
        const int c2 = i / (ST_TAPS * 2), rest = i - c2 * ST_TAPS * 2, k = rest >> 1, h = rest & 1;
        float v = 0.f;
#pragma unroll
        for (int g = 0; g < 16; ++g) v += red[((g * 16 + c2) * ST_TAPS + k) * 2 + h];
        atomicAdd(dw + (2 * c2 + h) * ST_TAPS + k, v);
    }
}

static int stem_check(const char *who, int B, int H, int W) {
    KDF_CHECK_ARG(B >= 0 && H > 0 && W > 0, "%s: bad sizes", who);
    KDF_CHECK_ARG((int64_t)B * ((H + 1) / 2) * ((W + 1) / 2) < (1ll << 31), "%s: image batch too large for 32-bit tile indexing", who);
    return KDF_OK;
}

static int stem_blocks(int B, int OH, int OW) {
    const int64_t tiles = (int64_t)B * ((OH + ST_R - 1) / ST_R) * ((OW + ST_XT - 1) / ST_XT);
    const int64_t cap = (int64_t)sm_count() * 2;
    if (tiles <= cap) return (int)(tiles < 1 ? 1 : tiles);
    const int64_t per = (tiles + cap - 1) / cap;                      // whole rounds of the persistent grid
    return (int)((tiles + per - 1) / per);
}

}  // namespace kdf

using namespace kdf;

extern "C" {

int kdf_stem_conv_fwd(const float *image, const float *weight, int B, int H, int W,
                      const float *post_scale, const float *post_shift, int post_act, void *out_bf16, double *stats, void *stream) {
    if (int e = stem_check("stem_conv_fwd", B, H, W)) return e;
    KDF_CHECK_ARG(!(stats && post_scale), "stem_conv_fwd: statistics and a folded BatchNorm exclude each other");
    KDF_CHECK_ARG((post_scale == nullptr) == (post_shift == nullptr) && post_act >= 0 && post_act <= 2, "stem_conv_fwd: bad epilogue");
    cudaStream_t st = as_stream(stream);
    if (stats) KDF_CUDA(cudaMemsetAsync(stats, 0, sizeof(double) * 2 * ST_CO, st));
    if (B == 0) return KDF_OK;
    KDF_CHECK_ARG(image && weight && out_bf16, "stem_conv_fwd: null pointer");
    KDF_CHECK_ARG((reinterpret_cast<uintptr_t>(out_bf16) & 3) == 0, "stem_conv_fwd: output must be 4-byte aligned");
    const int OH = (H - 1) / 2 + 1, OW = (W - 1) / 2 + 1;
    const int blocks = stem_blocks(B, OH, OW);
    __nv_bfloat16 *out = reinterpret_cast<__nv_bfloat16 *>(out_bf16);
    if (stats) stem_conv_fwd_kernel<ST_STATS><<<blocks, 256, 0, st>>>(image, weight, out, B, H, W, OH, OW, stats, nullptr, nullptr, 0);
    else if (post_scale) stem_conv_fwd_kernel<ST_POST><<<blocks, 256, 0, st>>>(image, weight, out, B, H, W, OH, OW, nullptr, post_scale, post_shift, post_act);
    else stem_conv_fwd_kernel<ST_PLAIN><<<blocks, 256, 0, st>>>(image, weight, out, B, H, W, OH, OW, nullptr, nullptr, nullptr, 0);
    KDF_LAUNCH_CHECK();
    return KDF_OK;
}

int kdf_stem_conv_bwd_weight(const float *image, const void *grad_out_bf16, int B, int H, int W, float *grad_weight, void *stream) {
    if (int e = stem_check("stem_conv_bwd_weight", B, H, W)) return e;
    KDF_CHECK_ARG(grad_weight, "stem_conv_bwd_weight: null pointer");
    cudaStream_t st = as_stream(stream);
    KDF_CUDA(cudaMemsetAsync(grad_weight, 0, sizeof(float) * ST_CO * ST_TAPS, st));
    if (B == 0) return KDF_OK;
    KDF_CHECK_ARG(image && grad_out_bf16, "stem_conv_bwd_weight: null pointer");
    KDF_CHECK_ARG((reinterpret_cast<uintptr_t>(grad_out_bf16) & 3) == 0, "stem_conv_bwd_weight: gradient must be 4-byte aligned");
    const int OH = (H - 1) / 2 + 1, OW = (W - 1) / 2 + 1;
    const size_t tile_bytes = sizeof(u64) * ST_TILE, red_bytes = sizeof(float) * 256 * ST_TAPS * 2;
    const size_t smem = tile_bytes > red_bytes ? tile_bytes : red_bytes;
    KDF_CUDA(cudaFuncSetAttribute(stem_conv_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    stem_conv_wgrad_kernel<<<stem_blocks(B, OH, OW), 256, smem, st>>>(image, reinterpret_cast<const __nv_bfloat16 *>(grad_out_bf16),
                                                                      grad_weight, B, H, W, OH, OW);
    KDF_LAUNCH_CHECK();
    return KDF_OK;
}

}  // extern "C"
