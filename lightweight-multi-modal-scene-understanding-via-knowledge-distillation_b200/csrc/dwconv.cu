// Depthwise 3x3 convolution (padding 1, stride 1 or 2, no bias) over pixel-major (NHWC) maps, sm_100a.
//
// The camera branch's inverted-residual blocks (reference src/models/camera_encoder.py:9-51), the FPN-lite
// smoothing and the segmentation head (fusion_module.py:20-34) run a depthwise 3x3 between two 1x1
// convolutions.  It is a 9-tap stencil per channel -- 18 FLOP per output element against 2*s bytes in and
// out -- i.e. HBM traffic, provided the kernel issues few enough instructions per element to keep up with it:
// at 6.5 TB/s an SM has about 12 issue slots per bf16 output element.  All kernels here are built the same way:
//   * a thread owns 4 channels (its 9x4 taps stay in registers as 18 packed fp32 pairs) and a strip of TWO adjacent
//     output columns, and slides down the rows of a row block: every input chunk is loaded once per thread (4 chunks
//     per row serve 2 outputs x 3 rows), unpacked once, and used in up to 9 multiply-adds;
//   * the arithmetic is packed: `fma.rn.f32x2` (FFMA2) does two channels per instruction with exactly the
//     rounding of two scalar FMAs -- 4.5 instructions per output element instead of 9;
//   * the row loop is fully unrolled (compile-time row block), so the three output rows in flight rotate through
//     registers without moves, and the loads of the next two rows are in flight while a row is consumed.
// forward (stride 1 / 2; + statistics of the BatchNorm that follows, or the folded BatchNorm + activation in inference),
// stride-1 data gradient = the forward with flipped taps, stride-2 data gradient in gather form (2x2 input patches
// sliding down the gradient rows), weight gradients with the same sliding window (36 partial sums per thread,
// block reduction, 16-byte vector reductions).
#include <stdlib.h>

#include "kdf_common.cuh"
#include "tma_common.cuh"

namespace kdf {

// A thread owns 4 channels (8-byte accesses for bf16, 16-byte for fp32) so that its 9x4 taps live in registers.
constexpr int DW_V = 4;
constexpr int DW_R1 = 16;        // stride 1: output rows per row block (18 input rows)
constexpr int DW_R2 = 8;         // stride 2: output rows per row block (17 input rows)
constexpr int DW_RD = 4;         // stride-2 data gradient: 2x2 patches per item along y (5 gradient rows)
// rows loaded ahead of the one being consumed (fp32 chunks are twice as wide: one row ahead keeps the kernel in registers)
template <typename T> struct DwAhead { static constexpr int PF = sizeof(T) == 2 ? 2 : 1; };

typedef unsigned long long u64;  // two fp32 values in a 64-bit register pair (lower channel in the low word)

__device__ __forceinline__ u64 pk2(float lo, float hi) {
    u64 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void upk2(u64 v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) {
    u64 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ u64 add2(u64 a, u64 b) {
    u64 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}

template <typename T> struct DwChunk;      // 4 consecutive channels of one pixel
template <> struct DwChunk<float> {
    typedef uint4 raw_t;
    static __device__ __forceinline__ raw_t ld(const float *p) { return __ldg(reinterpret_cast<const uint4 *>(p)); }
    static __device__ __forceinline__ raw_t zero() { return make_uint4(0u, 0u, 0u, 0u); }
    static __device__ __forceinline__ void unpack(const raw_t &u, u64 (&v)[2]) {
        v[0] = pk2(__uint_as_float(u.x), __uint_as_float(u.y));
        v[1] = pk2(__uint_as_float(u.z), __uint_as_float(u.w));
    }
    static __device__ __forceinline__ void st(float *p, const float *v) { *reinterpret_cast<float4 *>(p) = make_float4(v[0], v[1], v[2], v[3]); }
    static __device__ __forceinline__ void round(float *) {}                      // storage rounding: none for fp32
};
template <> struct DwChunk<__nv_bfloat16> {
    typedef uint2 raw_t;
    static __device__ __forceinline__ raw_t ld(const __nv_bfloat16 *p) { return __ldg(reinterpret_cast<const uint2 *>(p)); }
    static __device__ __forceinline__ raw_t zero() { return make_uint2(0u, 0u); }
    static __device__ __forceinline__ void unpack(const raw_t &u, u64 (&v)[2]) {
        v[0] = pk2(bf16_lo(u.x), bf16_hi(u.x));
        v[1] = pk2(bf16_lo(u.y), bf16_hi(u.y));
    }
    static __device__ __forceinline__ void st(__nv_bfloat16 *p, const float *v) {
        *reinterpret_cast<uint2 *>(p) = make_uint2(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]));
    }
    // the values as they are stored: two packed conversions + four shifts (a scalar cvt.rn.bf16.f32 per element compiles
    // to F2F on the conversion pipe, 8x the issue cost of an FMA)
    static __device__ __forceinline__ void round(float *v) {
        const uint32_t a = pack_bf16(v[0], v[1]), b = pack_bf16(v[2], v[3]);
        v[0] = bf16_lo(a); v[1] = bf16_hi(a); v[2] = bf16_lo(b); v[3] = bf16_hi(b);
    }
};

// the 9 taps of this thread's 4 channels as channel pairs: wr[k][p] = (w[(c0+2p)*9 + kk], w[(c0+2p+1)*9 + kk]), kk = FLIP ? 8-k : k
template <bool FLIP>
__device__ __forceinline__ void dw_load_taps(const float *__restrict__ w, int c0, u64 (&wr)[9][2]) {
#pragma unroll
    for (int p = 0; p < 2; ++p)
#pragma unroll
        for (int k = 0; k < 9; ++k) {
            const int kk = FLIP ? 8 - k : k;
            wr[k][p] = pk2(__ldg(w + (c0 + 2 * p) * 9 + kk), __ldg(w + (c0 + 2 * p + 1) * 9 + kk));
        }
}

enum { DW_PLAIN = 0, DW_STATS = 1, DW_POST = 2 };

// what happens to a finished output chunk: (inference) folded BatchNorm + activation on the value as it would have been
// stored, the store, (training) sums of the stored values for the BatchNorm that follows
template <typename T, int MODE>
__device__ __forceinline__ void dw_finish(const u64 (&a2)[2], T *dst, const float (&psc)[DW_V], const float (&psh)[DW_V], int post_act,
                                          u64 (&s_sum)[2], u64 (&s_sq)[2]) {
    typedef DwChunk<T> K;
    float a[DW_V];
    upk2(a2[0], a[0], a[1]);
    upk2(a2[1], a[2], a[3]);
    if (MODE == DW_POST) {
        K::round(a);
#pragma unroll
        for (int q = 0; q < DW_V; ++q) {
            float y = fmaf(a[q], psc[q], psh[q]);
            if (post_act == 1) y = fmaxf(y, 0.f);
            else if (post_act == 2) y = fminf(fmaxf(y, 0.f), 6.f);
            a[q] = y;
        }
    }
    K::st(dst, a);
    if (MODE == DW_STATS) {
        K::round(a);                                                       // the values stored
#pragma unroll
        for (int p = 0; p < 2; ++p) {
            const u64 v = pk2(a[2 * p], a[2 * p + 1]);
            s_sum[p] = add2(s_sum[p], v);
            s_sq[p] = fma2(v, v, s_sq[p]);
        }
    }
}

// per-CTA reduction of the statistics partials: threads tid = g (mod cg) share a channel group
__device__ __forceinline__ void dw_flush_stats(float *sred, const u64 (&s_sum)[2], const u64 (&s_sq)[2], int cg, int C,
                                               double *__restrict__ stats) {
#pragma unroll
    for (int p = 0; p < 2; ++p) {
        upk2(s_sum[p], sred[threadIdx.x * 2 * DW_V + 2 * p], sred[threadIdx.x * 2 * DW_V + 2 * p + 1]);
        upk2(s_sq[p], sred[threadIdx.x * 2 * DW_V + DW_V + 2 * p], sred[threadIdx.x * 2 * DW_V + DW_V + 2 * p + 1]);
    }
    __syncthreads();
    const int g0 = (blockIdx.x * blockDim.x) % cg;   // channel group of thread 0 of this CTA
    for (int i = threadIdx.x; i < cg * 2 * DW_V; i += blockDim.x) {
        const int gg = i / (2 * DW_V), e = i - gg * 2 * DW_V;
        int t0 = gg - g0;                             // first thread of the CTA whose channel group is gg, then every cg-th
        if (t0 < 0) t0 += cg;
        float v = 0.f;
        for (int t = t0; t < (int)blockDim.x; t += cg) v += sred[t * 2 * DW_V + e];
        const int which = e / DW_V, q = e - which * DW_V;
        if (v != 0.f) atomicAdd(stats + which * C + gg * DW_V + q, (double)v);
    }
}

// ----------------------------------------------------------------------------- stride 1: forward, and data gradient with FLIP
// Input row j of the block (iy = oy0 - 1 + j) feeds output rows r = j - ky; the three rows in flight live in acc[r % 3].
template <typename T, bool FLIP, int MODE>
__global__ void __launch_bounds__(256, 2)
dwconv3x3_s1_kernel(const T *__restrict__ in, const float *__restrict__ w /* [C][9] */, T *__restrict__ out,
                    int B, int H, int W, int C, double *__restrict__ stats /* MODE 1: [2][C] */,
                    const float *__restrict__ post_scale /* MODE 2: [C] */, const float *__restrict__ post_shift, int post_act) {
    typedef DwChunk<T> K;
    typedef typename K::raw_t raw_t;
    constexpr int R = DW_R1, NJ = R + 2, DW_PF = DwAhead<T>::PF;
    __shared__ float sred[MODE == DW_STATS ? 256 * 2 * DW_V : 1];
    const int cg = C / DW_V;
    const int XP = (W + 1) >> 1;
    const int nrb = (H + R - 1) / R;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = idx < XP * cg;
    const int xp = live ? idx / cg : 0, g = live ? idx - xp * cg : 0;
    u64 wr[9][2];
    dw_load_taps<FLIP>(w, g * DW_V, wr);
    u64 s_sum[2] = {0ull, 0ull}, s_sq[2] = {0ull, 0ull};
    float psc[DW_V], psh[DW_V];
#pragma unroll
    for (int q = 0; q < DW_V; ++q) {
        psc[q] = MODE == DW_POST ? __ldg(post_scale + g * DW_V + q) : 1.f;
        psh[q] = MODE == DW_POST ? __ldg(post_shift + g * DW_V + q) : 0.f;
    }
    const int x0 = 2 * xp;
    const bool ok0 = x0 >= 1, ok2 = x0 + 1 < W, ok3 = x0 + 2 < W;
    const int rs = W * C;                                                  // one map row; a frame's map is < 2^31 elements (host check)
    for (int rb = blockIdx.y; live && rb < B * nrb; rb += gridDim.y) {
        const int b = rb / nrb, oy0 = (rb - b * nrb) * R;
        const int64_t o_in = ((int64_t)b * H + (oy0 - 1)) * rs + (int64_t)(x0 - 1) * C + g * DW_V;     // chunk 0 of input row j = 0
        T *q0 = out + ((int64_t)b * H + oy0) * rs + (int64_t)x0 * C + g * DW_V;
        raw_t raw[NJ][4];
        auto ld_row = [&](int j) {
            const bool yok = (unsigned)(oy0 - 1 + j) < (unsigned)H;
            const int64_t o = o_in + j * rs;
            raw[j][0] = (yok && ok0) ? K::ld(in + o) : K::zero();
            raw[j][1] = yok ? K::ld(in + o + C) : K::zero();
            raw[j][2] = (yok && ok2) ? K::ld(in + o + 2 * C) : K::zero();
            raw[j][3] = (yok && ok3) ? K::ld(in + o + 3 * C) : K::zero();
        };
#pragma unroll
        for (int j = 0; j < DW_PF; ++j) ld_row(j);
        u64 acc[3][2][2];
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            if (j + DW_PF < NJ) ld_row(j + DW_PF);
            u64 v[4][2];
#pragma unroll
            for (int c = 0; c < 4; ++c) K::unpack(raw[j][c], v[c]);
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) {
                const int r = j - ky;                                     // compile-time after unrolling
                if (r >= 0 && r < R) {
                    const int s = r % 3;
                    if (ky == 0) { acc[s][0][0] = 0ull; acc[s][0][1] = 0ull; acc[s][1][0] = 0ull; acc[s][1][1] = 0ull; }
#pragma unroll
                    for (int kx = 0; kx < 3; ++kx)
#pragma unroll
                        for (int xo = 0; xo < 2; ++xo)
#pragma unroll
                            for (int p = 0; p < 2; ++p) acc[s][xo][p] = fma2(wr[ky * 3 + kx][p], v[xo + kx][p], acc[s][xo][p]);
                }
            }
            if (j >= 2) {
                const int r = j - 2, s = r % 3;
                if (oy0 + r < H) {
                    dw_finish<T, MODE>(acc[s][0], q0 + r * rs, psc, psh, post_act, s_sum, s_sq);
                    if (ok2) dw_finish<T, MODE>(acc[s][1], q0 + r * rs + C, psc, psh, post_act, s_sum, s_sq);
                }
            }
        }
    }
    if (MODE == DW_STATS) dw_flush_stats(sred, s_sum, s_sq, cg, C, stats);
}

// ----------------------------------------------------------------------------- stride 1, bf16: the input tile arrives by TMA
// The register kernel above waits on its own global loads half of the time (12 warps per SM, two rows ahead).  Here one
// thread asks the TMA unit for the whole input tile of the CTA -- DW_RT + 2 rows x (XT + 2) columns x CT channels, as
// NB boxes of DW_RB rows, each with its own mbarrier -- and the zero padding is the hardware's out-of-bounds fill (the box
// starts at x = -1 / y = -1).  The threads read their four chunks per row from shared memory (consecutive lanes =
// consecutive 8-byte chunks: conflict-free), so nothing in the arithmetic loop waits on DRAM and the rows ahead cost no
// registers; several CTAs per SM keep the next tiles in flight while one computes.
constexpr int DW_RT = 16;        // output rows per tile
constexpr int DW_RB = 2;         // rows per TMA box

template <bool FLIP, int MODE, int CT>
__global__ void __launch_bounds__(256, 2)
dwconv3x3_s1_tma_kernel(const __grid_constant__ CUtensorMap tm_in, const float *__restrict__ w, __nv_bfloat16 *__restrict__ out,
                        int H, int W, int C, int nxt, double *__restrict__ stats,
                        const float *__restrict__ post_scale, const float *__restrict__ post_shift, int post_act) {
    typedef __nv_bfloat16 T;
    typedef DwChunk<T> K;
    constexpr int G = CT / DW_V, XPT = 256 / G, XT = 2 * XPT, COLS = XT + 2;
    constexpr int R = DW_RT, NJ = R + 2, NB = NJ / DW_RB;
    constexpr uint32_t ROW_BYTES = COLS * CT * 2, BOX_BYTES = DW_RB * ROW_BYTES;
    extern __shared__ __align__(128) uint8_t dw_smem[];
    __shared__ uint64_t bars[NB];
    __shared__ float sred[MODE == DW_STATS ? 256 * 2 * DW_V : 1];
    const int tid = threadIdx.x;
    const int xt = blockIdx.x % nxt, slab = blockIdx.x / nxt;
    const int nrb = (H + R - 1) / R;
    const int b = blockIdx.y / nrb, oy0 = (blockIdx.y - b * nrb) * R;
    uint8_t *tile = dw_smem + ((128u - (tc::smem_u32(dw_smem) & 127u)) & 127u);
    if (tid == 0) {
#pragma unroll
        for (int k = 0; k < NB; ++k) tc::mbar_init(&bars[k], 1);
        tc::mbar_fence_init();
    }
    __syncthreads();
    if (tid == 0) {
#pragma unroll
        for (int k = 0; k < NB; ++k) {
            tma::mbar_expect_tx(&bars[k], BOX_BYTES);
            tma::load_4d(tile + k * BOX_BYTES, &tm_in, slab * CT, xt * XT - 1, oy0 - 1 + k * DW_RB, b, &bars[k]);
        }
    }
    const int xp = tid / G, g = tid - xp * G;
    const int c0 = slab * CT + g * DW_V;
    const int x0 = xt * XT + 2 * xp;
    const bool live = x0 < W, ok2 = x0 + 1 < W;
    u64 s_sum[2] = {0ull, 0ull}, s_sq[2] = {0ull, 0ull};
    if (live) {
        u64 wr[9][2];
        dw_load_taps<FLIP>(w, c0, wr);
        float psc[DW_V], psh[DW_V];
#pragma unroll
        for (int q = 0; q < DW_V; ++q) {
            psc[q] = MODE == DW_POST ? __ldg(post_scale + c0 + q) : 1.f;
            psh[q] = MODE == DW_POST ? __ldg(post_shift + c0 + q) : 0.f;
        }
        const int rs = W * C;
        T *q0 = out + ((int64_t)b * H + oy0) * rs + (int64_t)x0 * C + c0;
        const uint8_t *trow = tile + (2 * xp * CT + g * DW_V) * 2;
        u64 acc[3][2][2];
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            if (j % DW_RB == 0) tc::mbar_wait(&bars[j / DW_RB], 0);
            u64 v[4][2];
#pragma unroll
            for (int c = 0; c < 4; ++c) K::unpack(*reinterpret_cast<const uint2 *>(trow + j * ROW_BYTES + c * CT * 2), v[c]);
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) {
                const int r = j - ky;                                     // compile-time after unrolling
                if (r >= 0 && r < R) {
                    const int s = r % 3;
                    if (ky == 0) { acc[s][0][0] = 0ull; acc[s][0][1] = 0ull; acc[s][1][0] = 0ull; acc[s][1][1] = 0ull; }
#pragma unroll
                    for (int kx = 0; kx < 3; ++kx)
#pragma unroll
                        for (int xo = 0; xo < 2; ++xo)
#pragma unroll
                            for (int p = 0; p < 2; ++p) acc[s][xo][p] = fma2(wr[ky * 3 + kx][p], v[xo + kx][p], acc[s][xo][p]);
                }
            }
            if (j >= 2) {
                const int r = j - 2, s = r % 3;
                if (oy0 + r < H) {
                    dw_finish<T, MODE>(acc[s][0], q0 + r * rs, psc, psh, post_act, s_sum, s_sq);
                    if (ok2) dw_finish<T, MODE>(acc[s][1], q0 + r * rs + C, psc, psh, post_act, s_sum, s_sq);
                }
            }
        }
    }
    if (MODE == DW_STATS) {                          // thread tid holds channel group g = tid % G of this slab
#pragma unroll
        for (int p = 0; p < 2; ++p) {
            upk2(s_sum[p], sred[tid * 2 * DW_V + 2 * p], sred[tid * 2 * DW_V + 2 * p + 1]);
            upk2(s_sq[p], sred[tid * 2 * DW_V + DW_V + 2 * p], sred[tid * 2 * DW_V + DW_V + 2 * p + 1]);
        }
        __syncthreads();
        for (int i = tid; i < G * 2 * DW_V; i += 256) {
            const int gg = i / (2 * DW_V), e = i - gg * 2 * DW_V;
            float v = 0.f;
#pragma unroll 4
            for (int t = gg; t < 256; t += G) v += sred[t * 2 * DW_V + e];
            const int which = e / DW_V, q = e - which * DW_V;
            if (v != 0.f) atomicAdd(stats + which * C + slab * CT + gg * DW_V + q, (double)v);
        }
    }
}

// ----------------------------------------------------------------------------- stride 2: forward
// Output columns ox0, ox0+1 read input columns 2*ox0-1 .. 2*ox0+3 (5 chunks per row); input row j of the block
// (iy = 2*oy0 - 1 + j) feeds output row j/2 (ky = 0) and j/2-1 (ky = 2) when j is even, (j-1)/2 (ky = 1) when odd.
template <typename T, int MODE>
__global__ void __launch_bounds__(256, 2)
dwconv3x3_s2_kernel(const T *__restrict__ in, const float *__restrict__ w, T *__restrict__ out,
                    int B, int H, int W, int C, int OH, int OW, double *__restrict__ stats,
                    const float *__restrict__ post_scale, const float *__restrict__ post_shift, int post_act) {
    typedef DwChunk<T> K;
    typedef typename K::raw_t raw_t;
    constexpr int R = DW_R2, NJ = 2 * R + 1, DW_PF = DwAhead<T>::PF;
    __shared__ float sred[MODE == DW_STATS ? 256 * 2 * DW_V : 1];
    const int cg = C / DW_V;
    const int XP = (OW + 1) >> 1;
    const int nrb = (OH + R - 1) / R;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = idx < XP * cg;
    const int xp = live ? idx / cg : 0, g = live ? idx - xp * cg : 0;
    u64 wr[9][2];
    dw_load_taps<false>(w, g * DW_V, wr);
    u64 s_sum[2] = {0ull, 0ull}, s_sq[2] = {0ull, 0ull};
    float psc[DW_V], psh[DW_V];
#pragma unroll
    for (int q = 0; q < DW_V; ++q) {
        psc[q] = MODE == DW_POST ? __ldg(post_scale + g * DW_V + q) : 1.f;
        psh[q] = MODE == DW_POST ? __ldg(post_shift + g * DW_V + q) : 0.f;
    }
    const int ox0 = 2 * xp, ix0 = 2 * ox0 - 1;
    bool okc[5];
#pragma unroll
    for (int c = 0; c < 5; ++c) okc[c] = (unsigned)(ix0 + c) < (unsigned)W;
    const bool okx1 = ox0 + 1 < OW;
    const int rs = W * C, ors = OW * C;
    for (int rb = blockIdx.y; live && rb < B * nrb; rb += gridDim.y) {
        const int b = rb / nrb, oy0 = (rb - b * nrb) * R;
        const int64_t o_in = ((int64_t)b * H + (2 * oy0 - 1)) * rs + (int64_t)ix0 * C + g * DW_V;
        T *q0 = out + ((int64_t)b * OH + oy0) * ors + (int64_t)ox0 * C + g * DW_V;
        raw_t raw[NJ][5];
        auto ld_row = [&](int j) {
            const bool yok = (unsigned)(2 * oy0 - 1 + j) < (unsigned)H;
            const int64_t o = o_in + j * rs;
#pragma unroll
            for (int c = 0; c < 5; ++c) raw[j][c] = (yok && okc[c]) ? K::ld(in + o + c * C) : K::zero();
        };
#pragma unroll
        for (int j = 0; j < DW_PF; ++j) ld_row(j);
        u64 acc[2][2][2];
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            if (j + DW_PF < NJ) ld_row(j + DW_PF);
            u64 v[5][2];
#pragma unroll
            for (int c = 0; c < 5; ++c) K::unpack(raw[j][c], v[c]);
#pragma unroll
            for (int ky = 2; ky >= 0; --ky) {                            // the finishing row (ky = 2) first, then the starting one
                if (((j - ky) & 1) != 0) continue;
                const int r = (j - ky) / 2;
                if (r >= 0 && r < R && j - ky >= 0) {
                    const int s = r & 1;
                    if (ky == 0) { acc[s][0][0] = 0ull; acc[s][0][1] = 0ull; acc[s][1][0] = 0ull; acc[s][1][1] = 0ull; }
#pragma unroll
                    for (int kx = 0; kx < 3; ++kx)
#pragma unroll
                        for (int xo = 0; xo < 2; ++xo)
#pragma unroll
                            for (int p = 0; p < 2; ++p) acc[s][xo][p] = fma2(wr[ky * 3 + kx][p], v[2 * xo + kx][p], acc[s][xo][p]);
                }
            }
            if (j >= 2 && (j & 1) == 0) {
                const int r = j / 2 - 1, s = r & 1;
                if (oy0 + r < OH) {
                    dw_finish<T, MODE>(acc[s][0], q0 + r * ors, psc, psh, post_act, s_sum, s_sq);
                    if (okx1) dw_finish<T, MODE>(acc[s][1], q0 + r * ors + C, psc, psh, post_act, s_sum, s_sq);
                }
            }
        }
    }
    if (MODE == DW_STATS) dw_flush_stats(sred, s_sum, s_sq, cg, C, stats);
}

// ----------------------------------------------------------------------------- stride-2 data gradient (gather, 2x2 input patches)
// out(oy,ox) reads in(2*oy+ky-1, 2*ox+kx-1).  For the input patch rows {2a, 2a+1} x cols {2c, 2c+1}:
//   even row 2a   <- (oy=a,   ky=1);          odd row 2a+1 <- (oy=a, ky=2) and (oy=a+1, ky=0); same along x.
// A thread owns one patch column and DW_RD patches along y: DW_RD+1 gradient rows x 2 chunks, each loaded once.
template <typename T>
__global__ void __launch_bounds__(256, 2)
dwconv3x3_dgrad2_kernel(const T *__restrict__ gout, const float *__restrict__ w, T *__restrict__ gin,
                        int B, int H, int W, int C, int OH, int OW) {
    typedef DwChunk<T> K;
    typedef typename K::raw_t raw_t;
    constexpr int R = DW_RD;
    const int cg = C / DW_V;
    const int PH = (H + 1) / 2, PW = (W + 1) / 2;
    const int nrb = (PH + R - 1) / R;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= PW * cg) return;
    const int c = idx / cg, g = idx - c * cg;
    u64 wr[9][2];
    dw_load_taps<false>(w, g * DW_V, wr);
    const bool okg1 = c + 1 < OW, okx1 = 2 * c + 1 < W;
    const int grs = OW * C, rs = W * C;
    for (int rb = blockIdx.y; rb < B * nrb; rb += gridDim.y) {
        const int b = rb / nrb, a0 = (rb - b * nrb) * R;
        const int64_t o_g = ((int64_t)b * OH + a0) * grs + (int64_t)c * C + g * DW_V;
        T *q0 = gin + ((int64_t)b * H + 2 * a0) * rs + (int64_t)(2 * c) * C + g * DW_V;
        raw_t raw[R + 1][2];
#pragma unroll
        for (int j = 0; j <= R; ++j) {
            const bool yok = a0 + j < OH && c < OW;
            raw[j][0] = yok ? K::ld(gout + o_g + j * grs) : K::zero();
            raw[j][1] = (yok && okg1) ? K::ld(gout + o_g + j * grs + C) : K::zero();
        }
        u64 v0[2][2], v1[2][2];                                          // gradient rows a and a+1: [dx][pair]
        K::unpack(raw[0][0], v0[0]);
        K::unpack(raw[0][1], v0[1]);
#pragma unroll
        for (int j = 0; j < R; ++j) {
            K::unpack(raw[j + 1][0], v1[0]);
            K::unpack(raw[j + 1][1], v1[1]);
            u64 o[2][2][2];
#pragma unroll
            for (int p = 0; p < 2; ++p) {
                o[0][0][p] = fma2(wr[4][p], v0[0][p], 0ull);
                o[0][1][p] = fma2(wr[3][p], v0[1][p], fma2(wr[5][p], v0[0][p], 0ull));
                o[1][0][p] = fma2(wr[1][p], v1[0][p], fma2(wr[7][p], v0[0][p], 0ull));
                o[1][1][p] = fma2(wr[0][p], v1[1][p], fma2(wr[2][p], v1[0][p], fma2(wr[6][p], v0[1][p], fma2(wr[8][p], v0[0][p], 0ull))));
            }
            const int iy = 2 * (a0 + j);
#pragma unroll
            for (int dy = 0; dy < 2; ++dy)
#pragma unroll
                for (int dx = 0; dx < 2; ++dx) {
                    if (iy + dy < H && (dx == 0 || okx1)) {
                        float a[DW_V];
                        upk2(o[dy][dx][0], a[0], a[1]);
                        upk2(o[dy][dx][1], a[2], a[3]);
                        K::st(q0 + (2 * j + dy) * rs + dx * C, a);
                    }
                }
#pragma unroll
            for (int dx = 0; dx < 2; ++dx)
#pragma unroll
                for (int p = 0; p < 2; ++p) v0[dx][p] = v1[dx][p];
        }
    }
}

// ----------------------------------------------------------------------------- weight gradient
// dw[c][ky][kx] = sum over (b, oy, ox) of gout(b,oy,ox,c) * in(b, s*oy+ky-1, s*ox+kx-1, c).
// Same sliding window as the forward: a thread owns 4 channels and two gradient columns, walks down a row block, keeps the
// gradient rows that still meet coming input rows unpacked in registers, and accumulates its 36 partial sums as 18
// packed pairs; block reduction through shared memory, 16-byte vector reductions to global.
template <typename T, int STRIDE>
__global__ void __launch_bounds__(256, 2)
dwconv3x3_wgrad_kernel(const T *__restrict__ in, const T *__restrict__ gout, float *__restrict__ dw /* [C][9] */,
                       int B, int H, int W, int C, int OH, int OW) {
    typedef DwChunk<T> K;
    typedef typename K::raw_t raw_t;
    extern __shared__ float red[];                      // [blockDim/cg][cg][36]
    constexpr int R = STRIDE == 1 ? DW_R1 : DW_R2;
    constexpr int NJ = STRIDE == 1 ? R + 2 : 2 * R + 1;
    constexpr int NC = STRIDE == 1 ? 4 : 5;             // input chunks per row
    constexpr int NS = STRIDE == 1 ? 3 : 2;             // gradient rows in flight
    constexpr int DW_PF = DwAhead<T>::PF;
    const int cg = C / DW_V;
    const int g = threadIdx.x % cg, lane_row = threadIdx.x / cg, rows_per_cta = blockDim.x / cg;
    u64 acc[9][2];
#pragma unroll
    for (int k = 0; k < 9; ++k) { acc[k][0] = 0ull; acc[k][1] = 0ull; }
    const int XP = (OW + 1) >> 1;
    const int nrb = (OH + R - 1) / R;
    const int nitems = B * nrb * XP;                                     // < 2^31 (checked by the host wrapper)
    const int rs = W * C, grs = OW * C;
    for (int it = blockIdx.x * rows_per_cta + lane_row; it < nitems; it += gridDim.x * rows_per_cta) {
        const int xp = it % XP, rbi = it / XP;
        const int b = rbi / nrb, oy0 = (rbi - b * nrb) * R;
        const int ox0 = 2 * xp, ix0 = ox0 * STRIDE - 1, iy0 = oy0 * STRIDE - 1;
        bool okc[NC];
#pragma unroll
        for (int c = 0; c < NC; ++c) okc[c] = (unsigned)(ix0 + c) < (unsigned)W;
        const bool okx1 = ox0 + 1 < OW;
        const int64_t o_in = ((int64_t)b * H + iy0) * rs + (int64_t)ix0 * C + g * DW_V;
        const int64_t o_g = ((int64_t)b * OH + oy0) * grs + (int64_t)ox0 * C + g * DW_V;
        raw_t raw[NJ][NC], graw[R][2];
        auto ld_row = [&](int j) {
            const bool yok = (unsigned)(iy0 + j) < (unsigned)H;
            const int64_t o = o_in + j * rs;
#pragma unroll
            for (int c = 0; c < NC; ++c) raw[j][c] = (yok && okc[c]) ? K::ld(in + o + c * C) : K::zero();
        };
        auto ld_g = [&](int r) {
            const bool yok = oy0 + r < OH;
            graw[r][0] = yok ? K::ld(gout + o_g + r * grs) : K::zero();
            graw[r][1] = (yok && okx1) ? K::ld(gout + o_g + r * grs + C) : K::zero();
        };
        // gradient row r first meets input row j = r * STRIDE
#pragma unroll
        for (int j = 0; j < DW_PF; ++j) {
            ld_row(j);
            if (j % STRIDE == 0 && j / STRIDE < R) ld_g(j / STRIDE);
        }
        u64 gv[NS][2][2];
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            if (j + DW_PF < NJ) {
                ld_row(j + DW_PF);
                if ((j + DW_PF) % STRIDE == 0 && (j + DW_PF) / STRIDE < R) ld_g((j + DW_PF) / STRIDE);
            }
            u64 v[NC][2];
#pragma unroll
            for (int c = 0; c < NC; ++c) K::unpack(raw[j][c], v[c]);
            if (j % STRIDE == 0 && j / STRIDE < R) {
                const int r = j / STRIDE;
                K::unpack(graw[r][0], gv[r % NS][0]);
                K::unpack(graw[r][1], gv[r % NS][1]);
            }
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) {
                if ((j - ky) % STRIDE != 0 || j - ky < 0) continue;
                const int r = (j - ky) / STRIDE;
                if (r < R) {
#pragma unroll
                    for (int kx = 0; kx < 3; ++kx)
#pragma unroll
                        for (int xo = 0; xo < 2; ++xo)
#pragma unroll
                            for (int p = 0; p < 2; ++p)
                                acc[ky * 3 + kx][p] = fma2(gv[r % NS][xo][p], v[STRIDE * xo + kx][p], acc[ky * 3 + kx][p]);
                }
            }
        }
    }
#pragma unroll
    for (int k = 0; k < 9; ++k)
#pragma unroll
        for (int p = 0; p < 2; ++p) {
            float lo, hi;
            upk2(acc[k][p], lo, hi);
            red[(lane_row * cg + g) * 36 + k * DW_V + 2 * p] = lo;
            red[(lane_row * cg + g) * 36 + k * DW_V + 2 * p + 1] = hi;
        }
    __syncthreads();
    // a channel group's 4 x 9 gradients are 36 consecutive floats of dw: 9 16-byte vector reductions per group
    for (int i = threadIdx.x; i < cg * 9; i += blockDim.x) {
        const int gg = i / 9, j = i - gg * 9;
        float v[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const int o = j * 4 + t, q = o / 9, k = o - q * 9;              // dw[(gg*4 + q)*9 + k]
            float acc_ = 0.f;
            for (int r = 0; r < rows_per_cta; ++r) acc_ += red[(r * cg + gg) * 36 + k * DW_V + q];
            v[t] = acc_;
        }
        asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1, %2, %3, %4};"
                     ::"l"(dw + gg * 36 + j * 4), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]) : "memory");
    }
}

// ----------------------------------------------------------------------------- stride 2, bf16: the input tile arrives by TMA
// DW_RT2 output rows x XT output columns per CTA read 2*DW_RT2+1 input rows x 2*XT+1 input columns (the box is one column
// wider to stay even); boxes of DW_RB2 rows.  Same thread layout as the stride-1 kernel.
constexpr int DW_RT2 = 4;        // output rows per tile
constexpr int DW_RB2 = 3;        // input rows per TMA box

template <int MODE, int CT>
__global__ void __launch_bounds__(256, 2)
dwconv3x3_s2_tma_kernel(const __grid_constant__ CUtensorMap tm_in, const float *__restrict__ w, __nv_bfloat16 *__restrict__ out,
                        int OH, int OW, int C, int nxt, double *__restrict__ stats,
                        const float *__restrict__ post_scale, const float *__restrict__ post_shift, int post_act) {
    typedef __nv_bfloat16 T;
    typedef DwChunk<T> K;
    constexpr int G = CT / DW_V, XPT = 256 / G, XT = 2 * XPT, COLS = 2 * XT + 2;
    constexpr int R = DW_RT2, NJ = 2 * R + 1, NB = NJ / DW_RB2;
    constexpr uint32_t ROW_BYTES = COLS * CT * 2, BOX_BYTES = DW_RB2 * ROW_BYTES;
    extern __shared__ __align__(128) uint8_t dw_smem[];
    __shared__ uint64_t bars[NB];
    __shared__ float sred[MODE == DW_STATS ? 256 * 2 * DW_V : 1];
    const int tid = threadIdx.x;
    const int xt = blockIdx.x % nxt, slab = blockIdx.x / nxt;
    const int nrb = (OH + R - 1) / R;
    const int b = blockIdx.y / nrb, oy0 = (blockIdx.y - b * nrb) * R;
    uint8_t *tile = dw_smem + ((128u - (tc::smem_u32(dw_smem) & 127u)) & 127u);
    if (tid == 0) {
#pragma unroll
        for (int k = 0; k < NB; ++k) tc::mbar_init(&bars[k], 1);
        tc::mbar_fence_init();
    }
    __syncthreads();
    if (tid == 0) {
#pragma unroll
        for (int k = 0; k < NB; ++k) {
            tma::mbar_expect_tx(&bars[k], BOX_BYTES);
            tma::load_4d(tile + k * BOX_BYTES, &tm_in, slab * CT, 2 * xt * XT - 1, 2 * oy0 - 1 + k * DW_RB2, b, &bars[k]);
        }
    }
    const int xp = tid / G, g = tid - xp * G;
    const int c0 = slab * CT + g * DW_V;
    const int ox0 = xt * XT + 2 * xp;
    const bool live = ox0 < OW, okx1 = ox0 + 1 < OW;
    u64 s_sum[2] = {0ull, 0ull}, s_sq[2] = {0ull, 0ull};
    if (live) {
        u64 wr[9][2];
        dw_load_taps<false>(w, c0, wr);
        float psc[DW_V], psh[DW_V];
#pragma unroll
        for (int q = 0; q < DW_V; ++q) {
            psc[q] = MODE == DW_POST ? __ldg(post_scale + c0 + q) : 1.f;
            psh[q] = MODE == DW_POST ? __ldg(post_shift + c0 + q) : 0.f;
        }
        const int ors = OW * C;
        T *q0 = out + ((int64_t)b * OH + oy0) * ors + (int64_t)ox0 * C + c0;
        const uint8_t *trow = tile + (4 * xp * CT + g * DW_V) * 2;
        u64 acc[2][2][2];
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            if (j % DW_RB2 == 0) tc::mbar_wait(&bars[j / DW_RB2], 0);
            u64 v[5][2];
#pragma unroll
            for (int c = 0; c < 5; ++c) K::unpack(*reinterpret_cast<const uint2 *>(trow + j * ROW_BYTES + c * CT * 2), v[c]);
#pragma unroll
            for (int ky = 2; ky >= 0; --ky) {
                if (((j - ky) & 1) != 0) continue;
                const int r = (j - ky) / 2;
                if (r >= 0 && r < R && j - ky >= 0) {
                    const int s = r & 1;
                    if (ky == 0) { acc[s][0][0] = 0ull; acc[s][0][1] = 0ull; acc[s][1][0] = 0ull; acc[s][1][1] = 0ull; }
#pragma unroll
                    for (int kx = 0; kx < 3; ++kx)
#pragma unroll
                        for (int xo = 0; xo < 2; ++xo)
#pragma unroll
                            for (int p = 0; p < 2; ++p) acc[s][xo][p] = fma2(wr[ky * 3 + kx][p], v[2 * xo + kx][p], acc[s][xo][p]);
                }
            }
            if (j >= 2 && (j & 1) == 0) {
                const int r = j / 2 - 1, s = r & 1;
                if (oy0 + r < OH) {
                    dw_finish<T, MODE>(acc[s][0], q0 + r * ors, psc, psh, post_act, s_sum, s_sq);
                    if (okx1) dw_finish<T, MODE>(acc[s][1], q0 + r * ors + C, psc, psh, post_act, s_sum, s_sq);
                }
            }
        }
    }
    if (MODE == DW_STATS) {
#pragma unroll
        for (int p = 0; p < 2; ++p) {
            upk2(s_sum[p], sred[tid * 2 * DW_V + 2 * p], sred[tid * 2 * DW_V + 2 * p + 1]);
            upk2(s_sq[p], sred[tid * 2 * DW_V + DW_V + 2 * p], sred[tid * 2 * DW_V + DW_V + 2 * p + 1]);
        }
        __syncthreads();
        for (int i = tid; i < G * 2 * DW_V; i += 256) {
            const int gg = i / (2 * DW_V), e = i - gg * 2 * DW_V;
            float v = 0.f;
#pragma unroll 4
            for (int t = gg; t < 256; t += G) v += sred[t * 2 * DW_V + e];
            const int which = e / DW_V, q = e - which * DW_V;
            if (v != 0.f) atomicAdd(stats + which * C + slab * CT + gg * DW_V + q, (double)v);
        }
    }
}

// ----------------------------------------------------------------------------- weight gradient, bf16: both tiles arrive by TMA
// The input tile (with its halo) in boxes of rows, the gradient tile as one box; rows / columns past the maps are zero
// (hardware fill), so the loop has no predicates at all.  36 partial sums per thread, reduced over the CTA's column pairs
// through the (by then free) tile memory, one 16-byte vector reduction per 4 gradients.
template <int STRIDE, int CT>
__global__ void __launch_bounds__(256, 2)
dwconv3x3_wgrad_tma_kernel(const __grid_constant__ CUtensorMap tm_in, const __grid_constant__ CUtensorMap tm_g,
                           float *__restrict__ dw /* [C][9] */, int OH, int nxt) {
    typedef __nv_bfloat16 T;
    typedef DwChunk<T> K;
    constexpr int G = CT / DW_V, XPT = 256 / G, XT = 2 * XPT;
    constexpr int R = STRIDE == 1 ? 8 : DW_RT2, NJ = STRIDE == 1 ? R + 2 : 2 * R + 1;
    constexpr int RB = STRIDE == 1 ? DW_RB : DW_RB2, NB = NJ / RB;
    constexpr int COLS = STRIDE == 1 ? XT + 2 : 2 * XT + 2, NC = STRIDE == 1 ? 4 : 5, NS = STRIDE == 1 ? 3 : 2;
    constexpr uint32_t ROW_BYTES = COLS * CT * 2, BOX_BYTES = RB * ROW_BYTES, G_ROW_BYTES = XT * CT * 2, G_BYTES = R * G_ROW_BYTES;
    static_assert(NJ % RB == 0, "whole boxes");
    static_assert(NJ * ROW_BYTES + G_BYTES >= 256 * 36 * 4, "the reduction scratch aliases the tiles");
    extern __shared__ __align__(128) uint8_t dw_smem[];
    __shared__ uint64_t bars[NB];
    const int tid = threadIdx.x;
    const int xt = blockIdx.x % nxt, slab = blockIdx.x / nxt;
    const int nrb = (OH + R - 1) / R;
    const int b = blockIdx.y / nrb, oy0 = (blockIdx.y - b * nrb) * R;
    uint8_t *tile = dw_smem + ((128u - (tc::smem_u32(dw_smem) & 127u)) & 127u);
    uint8_t *gtile = tile + NJ * ROW_BYTES;
    if (tid == 0) {
#pragma unroll
        for (int k = 0; k < NB; ++k) tc::mbar_init(&bars[k], 1);
        tc::mbar_fence_init();
    }
    __syncthreads();
    if (tid == 0) {
        tma::mbar_expect_tx(&bars[0], BOX_BYTES + G_BYTES);
        tma::load_4d(gtile, &tm_g, slab * CT, xt * XT, oy0, b, &bars[0]);
#pragma unroll
        for (int k = 0; k < NB; ++k) {
            if (k > 0) tma::mbar_expect_tx(&bars[k], BOX_BYTES);
            tma::load_4d(tile + k * BOX_BYTES, &tm_in, slab * CT, STRIDE * xt * XT - 1, STRIDE * oy0 - 1 + k * RB, b, &bars[k]);
        }
    }
    const int xp = tid / G, g = tid - xp * G;
    u64 acc[9][2];
#pragma unroll
    for (int k = 0; k < 9; ++k) { acc[k][0] = 0ull; acc[k][1] = 0ull; }
    const uint8_t *trow = tile + (2 * STRIDE * xp * CT + g * DW_V) * 2;
    const uint8_t *grow = gtile + (2 * xp * CT + g * DW_V) * 2;
    u64 gv[NS][2][2];
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
        if (j % RB == 0) tc::mbar_wait(&bars[j / RB], 0);
        u64 v[NC][2];
#pragma unroll
        for (int c = 0; c < NC; ++c) K::unpack(*reinterpret_cast<const uint2 *>(trow + j * ROW_BYTES + c * CT * 2), v[c]);
        if (j % STRIDE == 0 && j / STRIDE < R) {                          // gradient row r first meets input row j = r * STRIDE
            const int r = j / STRIDE;
            K::unpack(*reinterpret_cast<const uint2 *>(grow + r * G_ROW_BYTES), gv[r % NS][0]);
            K::unpack(*reinterpret_cast<const uint2 *>(grow + r * G_ROW_BYTES + CT * 2), gv[r % NS][1]);
        }
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
            if ((j - ky) % STRIDE != 0 || j - ky < 0) continue;
            const int r = (j - ky) / STRIDE;
            if (r < R) {
#pragma unroll
                for (int kx = 0; kx < 3; ++kx)
#pragma unroll
                    for (int xo = 0; xo < 2; ++xo)
#pragma unroll
                        for (int p = 0; p < 2; ++p)
                            acc[ky * 3 + kx][p] = fma2(gv[r % NS][xo][p], v[STRIDE * xo + kx][p], acc[ky * 3 + kx][p]);
            }
        }
    }
    __syncthreads();                                                       // every thread is done with the tiles
    float *red = reinterpret_cast<float *>(tile);                          // [xp][g][36]
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        float a0, a1, a2, a3;
        upk2(acc[k][0], a0, a1);
        upk2(acc[k][1], a2, a3);
        *reinterpret_cast<float4 *>(red + tid * 36 + k * DW_V) = make_float4(a0, a1, a2, a3);
    }
    __syncthreads();
    // a channel group's 4 x 9 gradients are 36 consecutive floats of dw: 9 16-byte vector reductions per group
    for (int i = tid; i < G * 9; i += 256) {
        const int gg = i / 9, j = i - gg * 9;
        float v[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const int o = j * 4 + t, q = o / 9, k = o - q * 9;              // dw[(c0 + q)*9 + k]
            float a = 0.f;
#pragma unroll 4
            for (int x = 0; x < XPT; ++x) a += red[(x * G + gg) * 36 + k * DW_V + q];
            v[t] = a;
        }
        asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1, %2, %3, %4};"
                     ::"l"(dw + (slab * CT + gg * DW_V) * 9 + j * 4), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]) : "memory");
    }
}

static int dw_check(const char *who, int dtype, int B, int H, int W, int C, int stride) {
    KDF_CHECK_ARG(B >= 0 && H > 0 && W > 0, "%s: bad sizes", who);
    KDF_CHECK_ARG(dtype == KDF_F32 || dtype == KDF_BF16, "%s: bad dtype %d", who, dtype);
    KDF_CHECK_ARG(stride == 1 || stride == 2, "%s: stride %d not supported (1, 2)", who, stride);
    KDF_CHECK_ARG(C > 0 && C % DW_V == 0 && C / DW_V <= 256, "%s: C=%d must be a multiple of %d, at most %d", who, C, DW_V, 256 * DW_V);
    return KDF_OK;
}

// threads per CTA: the largest multiple of the channel-group count <= 256, so that a thread keeps its channel group
static int dw_block(int cg) { return (256 / cg) * cg; }

// grid.y for `items` row blocks next to grid.x = gx: about 8 CTAs per SM in total, and a count that divides the row blocks
// evenly (a thread walks items / gy of them with its taps in registers)
static int dw_grid_y(int gx, int items) {
    int gy = (sm_count() * 8 + gx - 1) / gx;
    if (gy >= items) return items < 1 ? 1 : items;
    const int per = (items + gy - 1) / gy;
    return (items + per - 1) / per;
}

}  // namespace kdf

using namespace kdf;

extern "C" {

static int dwconv_fwd_impl(const void *in, const float *weight, int dtype, int B, int H, int W, int C, int stride,
                           int flip, void *out, double *stats, const float *post_scale, const float *post_shift, int post_act,
                           void *stream) {
    if (int e = dw_check("dwconv3x3_fwd", dtype, B, H, W, C, stride)) return e;
    KDF_CHECK_ARG(!(flip && stride != 1), "dwconv3x3_fwd: flipped taps are the stride-1 data gradient only");
    KDF_CHECK_ARG(!(flip && (stats || post_scale)), "dwconv3x3_fwd: the data gradient has no epilogue");
    KDF_CHECK_ARG(!(stats && post_scale), "dwconv3x3_fwd: statistics and a folded BatchNorm exclude each other");
    if (B == 0) {
        if (stats) KDF_CUDA(cudaMemsetAsync(stats, 0, sizeof(double) * 2 * C, as_stream(stream)));
        return KDF_OK;
    }
    KDF_CHECK_ARG(in && weight && out, "dwconv3x3_fwd: null pointer");
    KDF_CHECK_ARG(((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15) == 0, "dwconv3x3_fwd: maps must be 16-byte aligned");
    const int OH = (H - 1) / stride + 1, OW = (W - 1) / stride + 1;
    const int cg = C / DW_V, nt = dw_block(cg);
    KDF_CHECK_ARG((int64_t)OW * cg < (1ll << 30) && (int64_t)B * OH < (1ll << 30) && (int64_t)H * W * C < (1ll << 31),
                  "dwconv3x3_fwd: map too large for 32-bit indexing");
    cudaStream_t st = as_stream(stream);
    if (stats) KDF_CUDA(cudaMemsetAsync(stats, 0, sizeof(double) * 2 * C, st));
    // bf16, stride 1, channel slabs of 64 or 32: the TMA-staged kernel (one tile per CTA)
    if (dtype == KDF_BF16 && stride == 1 && C % 32 == 0 && (int64_t)B * ((H + DW_RT - 1) / DW_RT) <= 65535) {
        const int CT = C % 64 == 0 ? 64 : 32;
        const int XT = 2 * (256 / (CT / DW_V)), COLS = XT + 2;
        CUtensorMap tm;
        KDF_CHECK_ARG(tma::make_nhwc_map(&tm, in, B, H, W, C, CT, COLS, DW_RB), "dwconv3x3_fwd: cuTensorMapEncodeTiled failed");
        const int nxt = (W + XT - 1) / XT;
        const dim3 tgrid((unsigned)(nxt * (C / CT)), (unsigned)(B * ((H + DW_RT - 1) / DW_RT)));
        const size_t smem = (size_t)(DW_RT + 2) * COLS * CT * 2 + 128;
        const int mode = stats ? DW_STATS : post_scale ? DW_POST : DW_PLAIN;
#define KDF_DWT(F, M, CTV)                                                                                                        \
    do {                                                                                                                          \
        KDF_CUDA(cudaFuncSetAttribute(dwconv3x3_s1_tma_kernel<F, M, CTV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        dwconv3x3_s1_tma_kernel<F, M, CTV><<<tgrid, 256, smem, st>>>(tm, weight, (__nv_bfloat16 *)out, H, W, C, nxt, stats, post_scale, \
                                                                     post_shift, post_act);                                       \
    } while (0)
#define KDF_DWT_CT(F, M) do { if (CT == 64) KDF_DWT(F, M, 64); else KDF_DWT(F, M, 32); } while (0)
        if (flip) KDF_DWT_CT(true, DW_PLAIN);
        else if (mode == DW_STATS) KDF_DWT_CT(false, DW_STATS);
        else if (mode == DW_POST) KDF_DWT_CT(false, DW_POST);
        else KDF_DWT_CT(false, DW_PLAIN);
#undef KDF_DWT_CT
#undef KDF_DWT
        KDF_LAUNCH_CHECK();
        return KDF_OK;
    }
    if (dtype == KDF_BF16 && stride == 2 && C % 32 == 0 && (int64_t)B * ((OH + DW_RT2 - 1) / DW_RT2) <= 65535) {
        const int CT = C % 64 == 0 ? 64 : 32;
        const int XT = 2 * (256 / (CT / DW_V)), COLS = 2 * XT + 2;
        CUtensorMap tm;
        KDF_CHECK_ARG(tma::make_nhwc_map(&tm, in, B, H, W, C, CT, COLS, DW_RB2), "dwconv3x3_fwd: cuTensorMapEncodeTiled failed");
        const int nxt = (OW + XT - 1) / XT;
        const dim3 tgrid((unsigned)(nxt * (C / CT)), (unsigned)(B * ((OH + DW_RT2 - 1) / DW_RT2)));
        const size_t smem = (size_t)(2 * DW_RT2 + 1) * COLS * CT * 2 + 128;
#define KDF_DWT(M, CTV)                                                                                                           \
    do {                                                                                                                          \
        KDF_CUDA(cudaFuncSetAttribute(dwconv3x3_s2_tma_kernel<M, CTV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        dwconv3x3_s2_tma_kernel<M, CTV><<<tgrid, 256, smem, st>>>(tm, weight, (__nv_bfloat16 *)out, OH, OW, C, nxt, stats, post_scale, \
                                                                  post_shift, post_act);                                          \
    } while (0)
#define KDF_DWT_CT(M) do { if (CT == 64) KDF_DWT(M, 64); else KDF_DWT(M, 32); } while (0)
        if (stats) KDF_DWT_CT(DW_STATS);
        else if (post_scale) KDF_DWT_CT(DW_POST);
        else KDF_DWT_CT(DW_PLAIN);
#undef KDF_DWT_CT
#undef KDF_DWT
        KDF_LAUNCH_CHECK();
        return KDF_OK;
    }
    const int R = stride == 1 ? DW_R1 : DW_R2;
    const int row_blocks = B * ((OH + R - 1) / R);
    const int gx = (((OW + 1) / 2) * cg + nt - 1) / nt;
    const dim3 grid((unsigned)gx, (unsigned)dw_grid_y(gx, row_blocks));
#define KDF_DW1(T, F, M) dwconv3x3_s1_kernel<T, F, M><<<grid, nt, 0, st>>>((const T *)in, weight, (T *)out, B, H, W, C, stats, post_scale, post_shift, post_act)
#define KDF_DW2(T, M) dwconv3x3_s2_kernel<T, M><<<grid, nt, 0, st>>>((const T *)in, weight, (T *)out, B, H, W, C, OH, OW, stats, post_scale, post_shift, post_act)
#define KDF_DW(T)                                                                                     \
    do {                                                                                              \
        if (stride == 1) {                                                                            \
            if (flip) KDF_DW1(T, true, DW_PLAIN);                                                     \
            else if (stats) KDF_DW1(T, false, DW_STATS);                                              \
            else if (post_scale) KDF_DW1(T, false, DW_POST);                                          \
            else KDF_DW1(T, false, DW_PLAIN);                                                         \
        } else {                                                                                      \
            if (stats) KDF_DW2(T, DW_STATS);                                                          \
            else if (post_scale) KDF_DW2(T, DW_POST);                                                 \
            else KDF_DW2(T, DW_PLAIN);                                                                \
        }                                                                                             \
    } while (0)
    if (dtype == KDF_F32) KDF_DW(float); else KDF_DW(__nv_bfloat16);
#undef KDF_DW
#undef KDF_DW2
#undef KDF_DW1
    KDF_LAUNCH_CHECK();
    return KDF_OK;
}

int kdf_dwconv3x3_fwd(const void *in, const float *weight, int dtype, int B, int H, int W, int C, int stride,
                      int flip, void *out, double *stats, void *stream) {
    return dwconv_fwd_impl(in, weight, dtype, B, H, W, C, stride, flip, out, stats, nullptr, nullptr, 0, stream);
}

int kdf_dwconv3x3_affine_fwd(const void *in, const float *weight, int dtype, int B, int H, int W, int C, int stride,
                             const float *post_scale, const float *post_shift, int act, void *out, void *stream) {
    KDF_CHECK_ARG(post_scale && post_shift, "dwconv3x3_affine_fwd: null pointer");
    KDF_CHECK_ARG(act >= 0 && act <= 2, "dwconv3x3_affine_fwd: bad activation %d", act);
    return dwconv_fwd_impl(in, weight, dtype, B, H, W, C, stride, 0, out, nullptr, post_scale, post_shift, act, stream);
}

int kdf_dwconv3x3_bwd_data(const void *grad_out, const float *weight, int dtype, int B, int H, int W, int C, int stride,
                           void *grad_in, void *stream) {
    if (stride == 1) return kdf_dwconv3x3_fwd(grad_out, weight, dtype, B, H, W, C, 1, 1, grad_in, nullptr, stream);
    if (int e = dw_check("dwconv3x3_bwd_data", dtype, B, H, W, C, stride)) return e;
    if (B == 0) return KDF_OK;
    KDF_CHECK_ARG(grad_out && weight && grad_in, "dwconv3x3_bwd_data: null pointer");
    KDF_CHECK_ARG(((reinterpret_cast<uintptr_t>(grad_out) | reinterpret_cast<uintptr_t>(grad_in)) & 15) == 0,
                  "dwconv3x3_bwd_data: maps must be 16-byte aligned");
    const int OH = (H - 1) / 2 + 1, OW = (W - 1) / 2 + 1;
    const int cg = C / DW_V, nt = dw_block(cg);
    KDF_CHECK_ARG((int64_t)W * cg < (1ll << 30) && (int64_t)B * H < (1ll << 30) && (int64_t)H * W * C < (1ll << 31),
                  "dwconv3x3_bwd_data: map too large for 32-bit indexing");
    const int PH = (H + 1) / 2, PW = (W + 1) / 2;
    const int gx = (PW * cg + nt - 1) / nt;
    const dim3 grid((unsigned)gx, (unsigned)dw_grid_y(gx, B * ((PH + DW_RD - 1) / DW_RD)));
    cudaStream_t st = as_stream(stream);
    if (dtype == KDF_F32) dwconv3x3_dgrad2_kernel<float><<<grid, nt, 0, st>>>((const float *)grad_out, weight, (float *)grad_in, B, H, W, C, OH, OW);
    else dwconv3x3_dgrad2_kernel<__nv_bfloat16><<<grid, nt, 0, st>>>((const __nv_bfloat16 *)grad_out, weight, (__nv_bfloat16 *)grad_in, B, H, W, C, OH, OW);
    KDF_LAUNCH_CHECK();
    return KDF_OK;
}

int kdf_dwconv3x3_bwd_weight(const void *in, const void *grad_out, int dtype, int B, int H, int W, int C, int stride,
                             float *grad_weight, void *stream) {
    if (int e = dw_check("dwconv3x3_bwd_weight", dtype, B, H, W, C, stride)) return e;
    KDF_CHECK_ARG(grad_weight, "dwconv3x3_bwd_weight: null pointer");
    cudaStream_t st = as_stream(stream);
    KDF_CUDA(cudaMemsetAsync(grad_weight, 0, sizeof(float) * 9 * C, st));
    if (B == 0) return KDF_OK;
    KDF_CHECK_ARG(in && grad_out, "dwconv3x3_bwd_weight: null pointer");
    KDF_CHECK_ARG(((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(grad_out)) & 15) == 0,
                  "dwconv3x3_bwd_weight: maps must be 16-byte aligned");
    const int OH = (H - 1) / stride + 1, OW = (W - 1) / stride + 1;
    {
        const int RT = stride == 1 ? 8 : DW_RT2;
        if (dtype == KDF_BF16 && C % 32 == 0 && (int64_t)B * ((OH + RT - 1) / RT) <= 65535) {
            const int CT = C % 64 == 0 ? 64 : 32;
            const int XT = 2 * (256 / (CT / DW_V)), COLS = stride == 1 ? XT + 2 : 2 * XT + 2;
            const int NJ = stride == 1 ? RT + 2 : 2 * RT + 1, RB = stride == 1 ? DW_RB : DW_RB2;
            CUtensorMap tm_in, tm_g;
            KDF_CHECK_ARG(tma::make_nhwc_map(&tm_in, in, B, H, W, C, CT, COLS, RB) && tma::make_nhwc_map(&tm_g, grad_out, B, OH, OW, C, CT, XT, RT),
                          "dwconv3x3_bwd_weight: cuTensorMapEncodeTiled failed");
            const int nxt = (OW + XT - 1) / XT;
            const dim3 tgrid((unsigned)(nxt * (C / CT)), (unsigned)(B * ((OH + RT - 1) / RT)));
            const size_t smem = (size_t)NJ * COLS * CT * 2 + (size_t)RT * XT * CT * 2 + 128;
#define KDF_DWWT(S, CTV)                                                                                                          \
    do {                                                                                                                          \
        KDF_CUDA(cudaFuncSetAttribute(dwconv3x3_wgrad_tma_kernel<S, CTV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        dwconv3x3_wgrad_tma_kernel<S, CTV><<<tgrid, 256, smem, st>>>(tm_in, tm_g, grad_weight, OH, nxt);                         \
    } while (0)
            if (stride == 1) { if (CT == 64) KDF_DWWT(1, 64); else KDF_DWWT(1, 32); }
            else { if (CT == 64) KDF_DWWT(2, 64); else KDF_DWWT(2, 32); }
#undef KDF_DWWT
            KDF_LAUNCH_CHECK();
            return KDF_OK;
        }
    }
    const int cg = C / DW_V;
    const int nt = dw_block(cg), rows = nt / cg;
    const int R = stride == 1 ? DW_R1 : DW_R2;
    const int64_t nitems = (int64_t)B * ((OH + R - 1) / R) * ((OW + 1) / 2);
    KDF_CHECK_ARG(nitems < (1ll << 30) && (int64_t)H * W * C < (1ll << 31), "dwconv3x3_bwd_weight: map too large for 32-bit indexing");
    int64_t blocks = (nitems + rows - 1) / rows;
    if (blocks > (int64_t)sm_count() * 4) {                              // whole rounds of the grid over the items
        const int64_t per = (blocks + (int64_t)sm_count() * 4 - 1) / ((int64_t)sm_count() * 4);
        blocks = (blocks + per - 1) / per;
    }
    const size_t smem = sizeof(float) * (size_t)nt * 36;
#define KDF_DWW(T, S)                                                                                          \
    do {                                                                                                       \
        KDF_CUDA(cudaFuncSetAttribute(dwconv3x3_wgrad_kernel<T, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        dwconv3x3_wgrad_kernel<T, S><<<(int)blocks, nt, smem, st>>>((const T *)in, (const T *)grad_out, grad_weight, B, H, W, C, OH, OW); \
    } while (0)
    if (dtype == KDF_F32) { if (stride == 2) KDF_DWW(float, 2); else KDF_DWW(float, 1); }
    else { if (stride == 2) KDF_DWW(__nv_bfloat16, 2); else KDF_DWW(__nv_bfloat16, 1); }
#undef KDF_DWW
    KDF_LAUNCH_CHECK();
    return KDF_OK;
}

}  // extern "C"
