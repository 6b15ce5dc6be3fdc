// Fused point-MLP layers on the 5th-gen tensor cores (tcgen05 + TMEM), sm_100a.
//
// The reference's point MLP (src/models/lidar_encoder.py:25-35,66) is Conv1d(k=1)+BatchNorm1d+ReLU
// three times over every point of the sweep.  Executed layer by layer it moves ~9 passes over a
// [B*N, 128] activation per layer (GEMM out, statistics, normalise, and the same again backwards).
// Here one layer is ONE kernel that keeps only the pre-BatchNorm outputs z in HBM:
//
//   prologue : global -> registers -> previous layer's BatchNorm-apply + ReLU (or, for the first
//              tensor-core layer, the whole 4->64 first layer recomputed from the raw fp32 point)
//              -> bf16 -> 128-byte-swizzled shared memory operand tile (128 points x K)
//   MMA      : tcgen05.mma  D[128 x 128, fp32, TMEM] = A[128 x K] . W^T      (one elected thread)
//   epilogue : tcgen05.ld -> bf16 -> swizzled staging tile -> coalesced 16-byte stores of z, and the
//              per-channel sum / sum of squares of exactly the values stored (this layer's BatchNorm
//              statistics), accumulated in registers over the CTA's tiles, fp64 atomics at the end.
//
// Persistent CTAs (one per SM), double-buffered operand tile and accumulator: the MMA of tile i runs
// while the global loads of tile i+1 are in flight and tile i-1 drains through the epilogue.
#include <stdlib.h>

#include "kdf_common.cuh"
#include "tc_common.cuh"
#include "tma_common.cuh"

namespace kdf {

constexpr int PM_ROWS = 128;          // points per tile = MMA M
constexpr int PM_N = 128;             // output channels of the tensor-core layers
// Threads per CTA are a template parameter NT (256 or 512): the CUDA-core prologue / epilogue work is latency-bound
// and 16 warps hide more of it, except for the first-layer-recompute forward kernel (measured: 0.46 ms at 256
// threads, 0.60 ms at 512).  Derived per kernel:
//   PM_CGROUPS = NT/128   TMEM epilogue: warp w reads lanes 32*(w%4).., column group w/4
//   PM_OROWS   = NT/16    rows per pass of the 16-chunks-per-row (256-byte row) phases
#define KDF_PM_DERIVED(NT)                                   \
    constexpr int PM_THREADS = NT;                           \
    constexpr int PM_CGROUPS = PM_THREADS / 128;             \
    constexpr int PM_OROWS = PM_THREADS / 16;                \
    constexpr int PM_OPASSES = PM_ROWS / PM_OROWS;           \
    (void)PM_CGROUPS; (void)PM_OROWS; (void)PM_OPASSES

struct MlpFwdArgs {
    const void *input;                // MODE 0: points f32 [M,4];  MODE 1: z_prev bf16 [M,KIN]
    int64_t M;
    const float *pro_a, *pro_b;       // MODE 0: q f32 [64,4], r f32 [64];  MODE 1: scale, shift f32 [KIN]
    const __nv_bfloat16 *W;           // [PM_N, KIN] row-major
    __nv_bfloat16 *z_out;             // [M, PM_N]
    double *stats;                    // [2][PM_N]  sum, sum of squares (accumulated)
};

template <int KIN>
struct MlpSmem {
    static constexpr int PANELS = KIN / 64;
    static constexpr int A_BYTES = PM_ROWS * KIN * 2;
    static constexpr int W_BYTES = PM_N * KIN * 2;
    static constexpr int STAGE_BYTES = PM_ROWS * PM_N * 2;
    static constexpr int OFF_W = 0;
    static constexpr int OFF_A0 = OFF_W + W_BYTES;
    static constexpr int OFF_A1 = OFF_A0 + A_BYTES;
    static constexpr int OFF_STAGE = OFF_A1 + A_BYTES;
    static constexpr int OFF_MISC = OFF_STAGE + STAGE_BYTES;     // barriers, tmem address, coefficient tables
    static constexpr int MISC_BYTES = 64 + 4 * (64 * 4 + 64 + 2 * 128);
    static constexpr int TOTAL = OFF_MISC + MISC_BYTES + 1024;    // + slack for 1024-byte alignment
};

// MODE 0: a1[c] = relu(q[c,:] . (x,y,z,i) + r[c])  for 8 channels c0..c0+7 of one point
// packed fp32 pairs: `fma.rn.f32x2` does two channels per instruction with the rounding of two scalar FMAs, and ReLU is taken
// on the packed bf16 pair after the conversion (0 is a bf16 number and the rounding is monotonic: max(round(x), 0) ==
// round(max(x, 0))) -- the prologues below are bit-identical to their scalar form at 55-70 % of its instructions
typedef unsigned long long pm_u64;
__device__ __forceinline__ pm_u64 pm_pk2(float lo, float hi) {
    pm_u64 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ pm_u64 pm_fma2(pm_u64 a, pm_u64 b, pm_u64 c) {
    pm_u64 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ pm_u64 pm_add2(pm_u64 a, pm_u64 b) {
    pm_u64 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ pm_u64 pm_unpack2(uint32_t w) { return pm_pk2(bf16_lo(w), bf16_hi(w)); }
__device__ __forceinline__ void pm_store8(float *dst, const pm_u64 (&v)[4]) {          // 4 pairs -> 8 consecutive floats
#pragma unroll
    for (int i = 0; i < 4; ++i) asm("mov.b64 {%0, %1}, %2;" : "=f"(dst[2 * i]), "=f"(dst[2 * i + 1]) : "l"(v[i]));
}
__device__ __forceinline__ uint32_t pm_gt0_mask(uint32_t a) {              // 0xFFFF per bf16 half of `a` that is > 0
    uint32_t d;
    asm("set.gt.u32.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(0u));
    return d;
}
__device__ __forceinline__ uint32_t pm_pack_relu(pm_u64 v) {               // bf16x2(relu(lo), relu(hi))
    float lo, hi;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
    uint32_t r;
    const uint32_t p = pack_bf16(lo, hi);
    asm("max.bf16x2 %0, %1, %2;" : "=r"(r) : "r"(p), "r"(0u));
    return r;
}
__device__ __forceinline__ uint4 first_layer_chunk(const float4 &p, const float *q, const float *r) {
    const pm_u64 px = pm_pk2(p.x, p.x), py = pm_pk2(p.y, p.y), pz = pm_pk2(p.z, p.z), pw = pm_pk2(p.w, p.w);
    uint32_t o[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {                                          // channels 2i, 2i+1 of the chunk
        pm_u64 acc = pm_pk2(r[2 * i], r[2 * i + 1]);
        acc = pm_fma2(pm_pk2(q[8 * i + 0], q[8 * i + 4]), px, acc);
        acc = pm_fma2(pm_pk2(q[8 * i + 1], q[8 * i + 5]), py, acc);
        acc = pm_fma2(pm_pk2(q[8 * i + 2], q[8 * i + 6]), pz, acc);
        acc = pm_fma2(pm_pk2(q[8 * i + 3], q[8 * i + 7]), pw, acc);
        o[i] = pm_pack_relu(acc);
    }
    return make_uint4(o[0], o[1], o[2], o[3]);
}
__device__ __forceinline__ uint4 affine_relu_chunk(const uint4 &u, const float *sc, const float *sh) {
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
    uint32_t o[4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
        o[i] = pm_pack_relu(pm_fma2(pm_pk2(bf16_lo(w[i]), bf16_hi(w[i])), pm_pk2(sc[2 * i], sc[2 * i + 1]), pm_pk2(sh[2 * i], sh[2 * i + 1])));
    return make_uint4(o[0], o[1], o[2], o[3]);
}

// MINB = 2: two co-resident CTAs per SM (layers 1+2: 85 KB of shared memory, 256 TMEM columns and <= 128 registers each),
// so that one CTA's barrier / TMEM round trips overlap the other's CUDA-core phases.
template <int MODE, int KIN, int NT, int MINB = 1>
__global__ void __launch_bounds__(NT, MINB)
mlp_layer_fwd_kernel(MlpFwdArgs a) {
    KDF_PM_DERIVED(NT);
    using L = MlpSmem<KIN>;
    constexpr int CHUNKS_PER_ROW = KIN / 8;                       // 16-byte chunks of one operand row
    constexpr int ROWS_PER_PASS = PM_THREADS / CHUNKS_PER_ROW;    // 16 (KIN=128) or 32 (KIN=64)
    constexpr int PASSES = PM_ROWS / ROWS_PER_PASS;               // 8 or 4
    constexpr uint32_t IDESC = tc::make_idesc(PM_ROWS, PM_N, 0, 0);

    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = tc::align_smem_1024(smem_raw);
    uint8_t *sW = smem + L::OFF_W, *sStage = smem + L::OFF_STAGE;
    auto sA = [smem](int b) -> uint8_t * { return smem + (b ? L::OFF_A1 : L::OFF_A0); };   // offsets from ONE shared base: LDS/STS, not generic
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + L::OFF_MISC);          // [2]
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + L::OFF_MISC + 16);
    float *coef = reinterpret_cast<float *>(smem + L::OFF_MISC + 64);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t n_tiles = (a.M + PM_ROWS - 1) / PM_ROWS;

    // ---- one-time setup: coefficient tables, weights -> swizzled smem, barriers, TMEM
    if (MODE == 0) {
        for (int i = tid; i < 64 * 4; i += PM_THREADS) coef[i] = a.pro_a[i];
        for (int i = tid; i < 64; i += PM_THREADS) coef[256 + i] = a.pro_b[i];
    } else {
        for (int i = tid; i < KIN; i += PM_THREADS) { coef[i] = a.pro_a[i]; coef[KIN + i] = a.pro_b[i]; }
    }
    for (int idx = tid; idx < PM_N * CHUNKS_PER_ROW; idx += PM_THREADS) {
        const int n = idx / CHUNKS_PER_ROW, ch = idx % CHUNKS_PER_ROW;
        const uint4 w = *reinterpret_cast<const uint4 *>(a.W + (int64_t)n * KIN + ch * 8);
        *reinterpret_cast<uint4 *>(sW + (ch >> 3) * (PM_N * tc::ROW_BYTES) + tc::sw128_offset(n, ch & 7)) = w;
    }
    if (tid == 0) {
        tc::mbar_init(&bars[0], 1);
        tc::mbar_init(&bars[1], 1);
        tc::mbar_fence_init();
    }
    if (warp == 0) tc::tmem_alloc(tmem_slot, 256);
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;

    // fixed per-thread operand chunk (so its coefficients stay in registers for the whole kernel)
    const int pch = tid % CHUNKS_PER_ROW, prow0 = tid / CHUNKS_PER_ROW;
    float c0[MODE == 0 ? 32 : 8], c1[8];
    if (MODE == 0) {
#pragma unroll
        for (int j = 0; j < 32; ++j) c0[j] = coef[pch * 32 + j];          // q rows of channels 8*pch..8*pch+7
#pragma unroll
        for (int j = 0; j < 8; ++j) c1[j] = coef[256 + pch * 8 + j];
    } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) { c0[j] = coef[pch * 8 + j]; c1[j] = coef[KIN + pch * 8 + j]; }
    }
    // fixed per-thread output chunk for the store / statistics phase: 16 chunks per 256-byte row
    const int och = tid & 15, orow0 = tid >> 4;
    float s_sum[8], s_sq[8];                          // (scalar on purpose: as packed pairs the 256-thread variant spills, 0.34 -> 0.39 ms)
#pragma unroll
    for (int j = 0; j < 8; ++j) { s_sum[j] = 0.f; s_sq[j] = 0.f; }

    uint4 raw[PASSES];
    auto load_tile = [&](int64_t tile) {
        const int64_t r0 = tile * PM_ROWS;
#pragma unroll
        for (int p = 0; p < PASSES; ++p) {
            const int64_t row = r0 + prow0 + p * ROWS_PER_PASS;
            if (row < a.M) {
                if (MODE == 0) raw[p] = ldg_stream_u4(reinterpret_cast<const uint4 *>(a.input) + row);
                else raw[p] = ldg_stream_u4(reinterpret_cast<const uint4 *>(reinterpret_cast<const __nv_bfloat16 *>(a.input) + row * KIN + pch * 8));
            }
        }
    };
    // DRAM -> L2 for a whole tile with one instruction: the register loads of that tile, one iteration later, then see
    // L2 latency instead of HBM latency (bytes in flight are bounded by registers, ~100 KB per SM otherwise)
    auto l2_prefetch = [&](int64_t tile) {
        if (tile >= n_tiles) return;
        const int64_t r0 = tile * PM_ROWS;
        const int64_t rows = (a.M - r0 < PM_ROWS) ? (a.M - r0) : PM_ROWS;
        if (MODE == 0) tc::prefetch_l2(reinterpret_cast<const uint4 *>(a.input) + r0, (uint32_t)(rows * 16));
        else tc::prefetch_l2(reinterpret_cast<const __nv_bfloat16 *>(a.input) + r0 * KIN, (uint32_t)(rows * KIN * 2));
    };
    auto stage_tile = [&](int64_t tile, uint8_t *dst) {
        const int64_t r0 = tile * PM_ROWS;
#pragma unroll
        for (int p = 0; p < PASSES; ++p) {
            const int r = prow0 + p * ROWS_PER_PASS;
            uint4 v = make_uint4(0u, 0u, 0u, 0u);                          // rows past M contribute exact zeros
            if (r0 + r < a.M) {
                if (MODE == 0) {
                    const float4 pt = make_float4(__uint_as_float(raw[p].x), __uint_as_float(raw[p].y),
                                                  __uint_as_float(raw[p].z), __uint_as_float(raw[p].w));
                    v = first_layer_chunk(pt, c0, c1);
                } else {
                    v = affine_relu_chunk(raw[p], c0, c1);
                }
            }
            *reinterpret_cast<uint4 *>(dst + (pch >> 3) * (PM_ROWS * tc::ROW_BYTES) + tc::sw128_offset(r, pch & 7)) = v;
        }
        tc::fence_async_smem();
    };
    auto issue_mma = [&](int buf) {                                        // one thread
        const uint32_t a_base = tc::smem_u32(sA(buf)), w_base = tc::smem_u32(sW);
        const uint32_t d = tmem_base + (uint32_t)buf * PM_N;
#pragma unroll
        for (int k = 0; k < KIN / 16; ++k) {
            const uint32_t koff = (uint32_t)(k >> 2) * (PM_ROWS * tc::ROW_BYTES) + (uint32_t)(k & 3) * 32u;
            const uint32_t woff = (uint32_t)(k >> 2) * (PM_N * tc::ROW_BYTES) + (uint32_t)(k & 3) * 32u;
            tc::mma_bf16(d, tc::desc_kmajor(a_base + koff), tc::desc_kmajor(w_base + woff), IDESC, k > 0);
        }
        tc::mma_commit(&bars[buf]);
    };
    auto epilogue = [&](int64_t tile, int buf, uint32_t parity) {
        tc::mbar_wait(&bars[buf], parity);
        tc::fence_after_sync();
        // warp w: TMEM lanes 32*(w%4).., columns 64*(w/4)..; thread = one output row
        constexpr int COLS_W = PM_N / PM_CGROUPS;                           // columns per warp
        const int row = (warp & 3) * 32 + lane;
        const uint32_t taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)buf * PM_N + (uint32_t)(warp >> 2) * COLS_W;
#pragma unroll
        for (int half = 0; half < COLS_W / 32; ++half) {
            uint32_t r[32];
            tc::tmem_ld32(taddr + half * 32, r);
            tc::tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int chunk = (warp >> 2) * (COLS_W / 8) + half * 4 + j; // 16-byte chunk of the 256-byte output row
                const uint4 v = make_uint4(pack_bf16(__uint_as_float(r[8 * j + 0]), __uint_as_float(r[8 * j + 1])),
                                           pack_bf16(__uint_as_float(r[8 * j + 2]), __uint_as_float(r[8 * j + 3])),
                                           pack_bf16(__uint_as_float(r[8 * j + 4]), __uint_as_float(r[8 * j + 5])),
                                           pack_bf16(__uint_as_float(r[8 * j + 6]), __uint_as_float(r[8 * j + 7])));
                *reinterpret_cast<uint4 *>(sStage + row * 256 + ((chunk ^ (row & 7)) << 4)) = v;
            }
        }
        tc::fence_before_sync();
        __syncthreads();
        // coalesced stores + statistics of exactly the stored bf16 values
        const int64_t r0 = tile * PM_ROWS;
#pragma unroll
        for (int p = 0; p < PM_OPASSES; ++p) {
            const int r = orow0 + p * PM_OROWS;
            if (r0 + r < a.M) {
                const uint4 v = *reinterpret_cast<const uint4 *>(sStage + r * 256 + ((och ^ (r & 7)) << 4));
                *reinterpret_cast<uint4 *>(a.z_out + (r0 + r) * PM_N + och * 8) = v;
                const float f[8] = {bf16_lo(v.x), bf16_hi(v.x), bf16_lo(v.y), bf16_hi(v.y),
                                    bf16_lo(v.z), bf16_hi(v.z), bf16_lo(v.w), bf16_hi(v.w)};
#pragma unroll
                for (int j = 0; j < 8; ++j) { s_sum[j] += f[j]; s_sq[j] = fmaf(f[j], f[j], s_sq[j]); }
            }
        }
    };

    // ---- software pipeline over this CTA's tiles
    int64_t tile = blockIdx.x;
    if (tile < n_tiles) {
        load_tile(tile);
        stage_tile(tile, sA(0));
    }
    __syncthreads();
    int it = 0;
    int64_t prev_tile = -1;
    for (; tile < n_tiles; tile += gridDim.x, ++it) {
        const int buf = it & 1;
        if (tid == 0) {
            tc::fence_after_sync();
            issue_mma(buf);
        }
        const int64_t next = tile + gridDim.x;
        if (next < n_tiles) load_tile(next);                               // global loads in flight during the epilogue
        if (tid == 32) l2_prefetch(next + gridDim.x);                      // and the tile after it on its way into L2
        if (it > 0) epilogue(prev_tile, buf ^ 1, (uint32_t)(((it - 1) >> 1) & 1));
        if (next < n_tiles) stage_tile(next, sA(buf ^ 1));                 // A[buf^1] was consumed by the MMA just waited on
        __syncthreads();
        prev_tile = tile;
    }
    if (it > 0) {
        epilogue(prev_tile, (it - 1) & 1, (uint32_t)(((it - 1) >> 1) & 1));
    }
    __syncthreads();

    // ---- statistics: reduce the 16 threads that share an output chunk, then 2*128 fp64 atomics per CTA
    float *red = reinterpret_cast<float *>(sStage);                       // [PM_OROWS row lanes][16 chunks][16]
#pragma unroll
    for (int j = 0; j < 8; ++j) { red[(orow0 * 16 + och) * 16 + j] = s_sum[j]; red[(orow0 * 16 + och) * 16 + 8 + j] = s_sq[j]; }
    __syncthreads();
    if (tid < 16 * 16) {
        const int ch = tid >> 4, j = tid & 15;                             // chunk, (sum|sq, element)
        float v = 0.f;
        for (int r = 0; r < PM_OROWS; ++r) v += red[(r * 16 + ch) * 16 + j];
        const int col = ch * 8 + (j & 7);
        atomicAdd(a.stats + (j >> 3) * PM_N + col, (double)v);
    }
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem_base, 256);
}


// ============================================================================= all three layers, running statistics
// With running statistics (eval mode: the frozen teacher of the distillation step) no batch statistic separates
// the layers, so the whole point MLP is ONE kernel per 128-point tile:
//   prologue  : fp32 point -> layer 1 (folded BatchNorm) + ReLU -> bf16 operand tile A1 [128 x 64]
//   MMA 1     : z2 = A1 . W2^T                     (TMEM columns 0..127)
//   epilogue 1: tcgen05.ld -> round to bf16 (what the layer-by-layer kernels store) -> BatchNorm-2 apply + ReLU
//               -> bf16 -> operand tile A2 [128 x 128] in shared memory: z2 never exists in HBM
//   MMA 2     : z3 = A2 . W3^T                     (TMEM columns 128..255)
//   epilogue 2: tcgen05.ld -> bf16 -> staging -> coalesced 16-byte stores of z3 (no statistics in eval mode)
// 16 B in, C*2 B out per point instead of 16 + 3*C*2.  The arithmetic is that of mlp_layer_fwd_kernel<0> followed
// by <1>, operation for operation, so z3 is bit-identical to the two-kernel path.
// Pipeline: MMA 1 of tile i runs while the points of tile i+1 are loaded and tile i-1 drains through epilogue 2.
struct MlpEvalArgs {
    const void *points;               // f32 [M,4]
    int64_t M;
    const float *q, *r;               // folded first layer: a1 = relu(q . p + r), q f32 [64,4], r f32 [64]
    const __nv_bfloat16 *W2;          // [128, 64]
    const float *sc2, *sh2;           // BatchNorm-2 (running statistics) as scale / shift f32 [128]
    const __nv_bfloat16 *W3;          // [128, 128]
    __nv_bfloat16 *z_out;             // [M, 128] pre-BatchNorm-3 rows
};

struct MlpEvalSmem {
    static constexpr int OFF_W2 = 0;                                   // 1 panel  [128 n][64 k]
    static constexpr int OFF_W3 = OFF_W2 + PM_N * 64 * 2;              // 2 panels [128 n][128 k]
    // Two CTAs share an SM (99 KB and 256 TMEM columns each): one CTA's TMEM epilogues overlap the other's MMAs and loads.
    // That needs the A1 tile single-buffered (it is re-staged after MMA 1 has been waited for) and the staging tile of the
    // output rows ALIASED onto the A2 operand tile (free once MMA 2 has completed, which is what epilogue 2 waits for).
    static constexpr int OFF_A1_0 = OFF_W3 + PM_N * 128 * 2;
    static constexpr int OFF_A2 = OFF_A1_0 + PM_ROWS * 64 * 2;
    static constexpr int OFF_STAGE = OFF_A2;
    static constexpr int OFF_MISC = OFF_A2 + PM_ROWS * 128 * 2;
    static constexpr int MISC_BYTES = 64 + 4 * (64 * 4 + 64 + 2 * 128);
    static constexpr int TOTAL = OFF_MISC + MISC_BYTES + 1024;
};

template <int NT>
__global__ void __launch_bounds__(NT, 2)
mlp_eval3_kernel(MlpEvalArgs a) {
    KDF_PM_DERIVED(NT);
    using L = MlpEvalSmem;
    constexpr int ROWS_PER_PASS = PM_THREADS / 8;                 // 8 chunks per 64-wide A1 row
    constexpr int PASSES = PM_ROWS / ROWS_PER_PASS;
    constexpr uint32_t IDESC = tc::make_idesc(PM_ROWS, PM_N, 0, 0);
    constexpr uint32_t PANEL = PM_ROWS * tc::ROW_BYTES;

    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = tc::align_smem_1024(smem_raw);
    uint8_t *sW2 = smem + L::OFF_W2, *sW3 = smem + L::OFF_W3;
    uint8_t *sA1 = smem + L::OFF_A1_0;
    uint8_t *sA2 = smem + L::OFF_A2, *sStage = smem + L::OFF_STAGE;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + L::OFF_MISC);          // [0] MMA 1 done, [1] MMA 2 done
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + L::OFF_MISC + 16);
    float *coef = reinterpret_cast<float *>(smem + L::OFF_MISC + 64);           // q[256] r[64] sc2[128] sh2[128]
    const float *tsc = coef + 320, *tsh = coef + 448;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t n_tiles = (a.M + PM_ROWS - 1) / PM_ROWS;

    for (int i = tid; i < 64 * 4; i += PM_THREADS) coef[i] = a.q[i];
    for (int i = tid; i < 64; i += PM_THREADS) coef[256 + i] = a.r[i];
    for (int i = tid; i < 128; i += PM_THREADS) { coef[320 + i] = a.sc2[i]; coef[448 + i] = a.sh2[i]; }
    for (int idx = tid; idx < PM_N * 8; idx += PM_THREADS) {
        const int n = idx >> 3, ch = idx & 7;
        *reinterpret_cast<uint4 *>(sW2 + tc::sw128_offset(n, ch)) = *reinterpret_cast<const uint4 *>(a.W2 + (int64_t)n * 64 + ch * 8);
    }
    for (int idx = tid; idx < PM_N * 16; idx += PM_THREADS) {
        const int n = idx >> 4, ch = idx & 15;
        *reinterpret_cast<uint4 *>(sW3 + (ch >> 3) * (PM_N * tc::ROW_BYTES) + tc::sw128_offset(n, ch & 7)) =
            *reinterpret_cast<const uint4 *>(a.W3 + (int64_t)n * 128 + ch * 8);
    }
    if (tid == 0) {
        tc::mbar_init(&bars[0], 1);
        tc::mbar_init(&bars[1], 1);
        tc::mbar_fence_init();
    }
    if (warp == 0) tc::tmem_alloc(tmem_slot, 256);
    tc::fence_async_smem();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;

    const int pch = tid & 7, prow0 = tid >> 3;
    float c0[32], c1[8];
#pragma unroll
    for (int j = 0; j < 32; ++j) c0[j] = coef[pch * 32 + j];
#pragma unroll
    for (int j = 0; j < 8; ++j) c1[j] = coef[256 + pch * 8 + j];
    const int och = tid & 15, orow0 = tid >> 4;

    uint4 raw[PASSES];
    auto load_tile = [&](int64_t tile) {
        const int64_t r0 = tile * PM_ROWS;
#pragma unroll
        for (int p = 0; p < PASSES; ++p) {
            const int64_t row = r0 + prow0 + p * ROWS_PER_PASS;
            if (row < a.M) raw[p] = ldg_stream_u4(reinterpret_cast<const uint4 *>(a.points) + row);
        }
    };
    auto l2_prefetch = [&](int64_t tile) {
        if (tile >= n_tiles) return;
        const int64_t r0 = tile * PM_ROWS;
        const int64_t rows = (a.M - r0 < PM_ROWS) ? (a.M - r0) : PM_ROWS;
        tc::prefetch_l2(reinterpret_cast<const uint4 *>(a.points) + r0, (uint32_t)(rows * 16));
    };
    auto stage_tile = [&](int64_t tile, uint8_t *dst) {
        const int64_t r0 = tile * PM_ROWS;
#pragma unroll
        for (int p = 0; p < PASSES; ++p) {
            const int r = prow0 + p * ROWS_PER_PASS;
            uint4 v = make_uint4(0u, 0u, 0u, 0u);
            if (r0 + r < a.M) {
                const float4 pt = make_float4(__uint_as_float(raw[p].x), __uint_as_float(raw[p].y),
                                              __uint_as_float(raw[p].z), __uint_as_float(raw[p].w));
                v = first_layer_chunk(pt, c0, c1);
            }
            *reinterpret_cast<uint4 *>(dst + tc::sw128_offset(r, pch)) = v;
        }
        tc::fence_async_smem();
    };
    auto issue_mma1 = [&]() {                                              // one thread
        const uint32_t a_base = tc::smem_u32(sA1), w_base = tc::smem_u32(sW2);
#pragma unroll
        for (int k = 0; k < 4; ++k)
            tc::mma_bf16(tmem_base, tc::desc_kmajor(a_base + (uint32_t)k * 32u), tc::desc_kmajor(w_base + (uint32_t)k * 32u), IDESC, k > 0);
        tc::mma_commit(&bars[0]);
    };
    auto issue_mma2 = [&]() {                                              // one thread
        const uint32_t a_base = tc::smem_u32(sA2), w_base = tc::smem_u32(sW3);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const uint32_t koff = (uint32_t)(k >> 2) * PANEL + (uint32_t)(k & 3) * 32u;
            tc::mma_bf16(tmem_base + PM_N, tc::desc_kmajor(a_base + koff), tc::desc_kmajor(w_base + koff), IDESC, k > 0);
        }
        tc::mma_commit(&bars[1]);
    };
    constexpr int COLS_W = PM_N / PM_CGROUPS;                              // accumulator columns per warp
    const int erow = (warp & 3) * 32 + lane;                               // this thread's accumulator row
    const uint32_t lane_bits = (uint32_t)((warp & 3) * 32) << 16;
    auto epilogue1 = [&](uint32_t parity) {                                // z2 -> BatchNorm-2 + ReLU -> A2
        tc::mbar_wait(&bars[0], parity);
        tc::fence_after_sync();
#pragma unroll
        for (int half = 0; half < COLS_W / 32; ++half) {
            uint32_t r[32];
            tc::tmem_ld32(tmem_base + lane_bits + (uint32_t)(warp >> 2) * COLS_W + half * 32, r);
            tc::tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int chunk = (warp >> 2) * (COLS_W / 8) + half * 4 + j;
                const int col = chunk * 8;
                // coefficients as four 16-byte broadcast loads per chunk (not sixteen 4-byte ones)
                const float4 sa = *reinterpret_cast<const float4 *>(tsc + col), sb = *reinterpret_cast<const float4 *>(tsc + col + 4);
                const float4 ha = *reinterpret_cast<const float4 *>(tsh + col), hb = *reinterpret_cast<const float4 *>(tsh + col + 4);
                const float cs[8] = {sa.x, sa.y, sa.z, sa.w, sb.x, sb.y, sb.z, sb.w};
                const float ch_[8] = {ha.x, ha.y, ha.z, ha.w, hb.x, hb.y, hb.z, hb.w};
                uint32_t o[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const uint32_t zz = pack_bf16(__uint_as_float(r[8 * j + 2 * e]), __uint_as_float(r[8 * j + 2 * e + 1]));   // z2 as stored
                    o[e] = pm_pack_relu(pm_fma2(pm_unpack2(zz), pm_pk2(cs[2 * e], cs[2 * e + 1]), pm_pk2(ch_[2 * e], ch_[2 * e + 1])));
                }
                *reinterpret_cast<uint4 *>(sA2 + (uint32_t)(chunk >> 3) * PANEL + tc::sw128_offset(erow, chunk & 7)) = make_uint4(o[0], o[1], o[2], o[3]);
            }
        }
        tc::fence_async_smem();
        tc::fence_before_sync();
    };
    auto epilogue2 = [&](int64_t tile, uint32_t parity) {                  // z3 -> bf16 rows
        tc::mbar_wait(&bars[1], parity);
        tc::fence_after_sync();
#pragma unroll
        for (int half = 0; half < COLS_W / 32; ++half) {
            uint32_t r[32];
            tc::tmem_ld32(tmem_base + lane_bits + PM_N + (uint32_t)(warp >> 2) * COLS_W + half * 32, r);
            tc::tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int chunk = (warp >> 2) * (COLS_W / 8) + half * 4 + j;
                const uint4 v = make_uint4(pack_bf16(__uint_as_float(r[8 * j + 0]), __uint_as_float(r[8 * j + 1])),
                                           pack_bf16(__uint_as_float(r[8 * j + 2]), __uint_as_float(r[8 * j + 3])),
                                           pack_bf16(__uint_as_float(r[8 * j + 4]), __uint_as_float(r[8 * j + 5])),
                                           pack_bf16(__uint_as_float(r[8 * j + 6]), __uint_as_float(r[8 * j + 7])));
                *reinterpret_cast<uint4 *>(sStage + erow * 256 + ((chunk ^ (erow & 7)) << 4)) = v;
            }
        }
        tc::fence_before_sync();
        __syncthreads();
        const int64_t r0 = tile * PM_ROWS;
#pragma unroll
        for (int p = 0; p < PM_OPASSES; ++p) {
            const int r = orow0 + p * PM_OROWS;
            if (r0 + r < a.M)
                *reinterpret_cast<uint4 *>(a.z_out + (r0 + r) * PM_N + och * 8) =
                    *reinterpret_cast<const uint4 *>(sStage + r * 256 + ((och ^ (r & 7)) << 4));
        }
    };

    int64_t tile = blockIdx.x;
    if (tile < n_tiles) {
        load_tile(tile);
        stage_tile(tile, sA1);
    }
    __syncthreads();
    int it = 0;
    int64_t prev_tile = -1;
    for (; tile < n_tiles; tile += gridDim.x, ++it) {
        if (tid == 0) {
            tc::fence_after_sync();
            issue_mma1();
        }
        const int64_t next = tile + gridDim.x;
        if (next < n_tiles) load_tile(next);
        if (tid == 32) l2_prefetch(next + gridDim.x);
        if (it > 0) {
            epilogue2(prev_tile, (uint32_t)((it - 1) & 1));                // waits for MMA 2 of the previous tile: A2 is free, the rows go out
            __syncthreads();                                               // ... through the staging tile that IS the A2 tile: drained before epilogue 1 refills it
        }
        epilogue1((uint32_t)(it & 1));                                     // waits for MMA 1: the A1 tile is free for the next tile's rows
        if (next < n_tiles) stage_tile(next, sA1);
        __syncthreads();
        if (tid == 0) {
            tc::fence_after_sync();
            issue_mma2();
        }
        prev_tile = tile;
    }
    if (it > 0) epilogue2(prev_tile, (uint32_t)((it - 1) & 1));
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem_base, 256);
}


// ============================================================================= fused backward layer
// One tensor-core layer of the point MLP, backwards, as ONE kernel.  For a 128-point tile it forms, from
// the gradient dy w.r.t. this layer's BatchNorm output (ReLU already folded in) and the stored pre-BatchNorm
// rows z, the BatchNorm backward       dz = gs*dy + ga + gb*z            (per-channel coefficients that chain
// through the batch mean / variance, prepared by the host from the two column sums of the previous kernel),
// re-creates the layer's INPUT activation in the prologue exactly like the forward kernel did, and runs
//   dgrad : dA[128 pts x KIN]  = dz . W            (A = dz tile K-major, B = W tile MN-major)
//   wgrad : dW[128 x KIN]     += dz^T . A          (both operands are the MN-major view of the same tiles;
//                                                   the accumulator stays in TMEM for all tiles of the CTA)
// Epilogue, MODE 1 (layer 3 -> 2): dy_prev = dA * (a_prev > 0) stored as bf16 together with its two column
// sums (sum dy_prev, sum dy_prev * z_prev) -- the inputs of the next layer's BatchNorm backward.
// Epilogue, MODE 0 (layer 2 -> 1): the first layer is linear in the raw point, so nothing per-point is
// stored: only S0[c] = sum dy1[.,c] and T[c,:] = sum dy1[.,c] * point  (64 x 5 numbers) leave the kernel.
struct MlpBwdArgs {
    const __nv_bfloat16 *dy;          // [M,128]
    const __nv_bfloat16 *z;           // [M,128]
    const float *gs, *ga, *gb;        // [128]
    const void *input;                // MODE 1: z_prev bf16 [M,128];  MODE 0: points f32 [M,4]
    const float *pro_a, *pro_b;       // as in the forward kernel
    const __nv_bfloat16 *W;           // [128, KIN] row-major
    int64_t M;
    __nv_bfloat16 *dy_prev;           // MODE 1: [M,128]
    double *sums;                     // MODE 1: [2][128];  MODE 0: [5][64] = S0, T[:,x], T[:,y], T[:,z], T[:,i]
    float *dW;                        // [128, KIN] fp32, accumulated with atomics (zeroed by the host wrapper)
    const int32_t *row_cell;          // nullable [M]: rows with row_cell < 0 have dy == 0 by contract and their dy rows are
                                      // not read as data (the producer kdf_bev_bwd_affine may then skip writing them)
    // cell-sorted rows (SHARE kernels): dy is not materialised; row_cell holds GLOBAL cell ids and
    // dy[row][c] = bit c of bits[row] ? share[row_cell[row]][c] : 0   (kdf_bev_bwd_share)
    const __nv_bfloat16 *share;       // [cells,128]
    const uint8_t *bits;              // [M,16]
};

template <int KIN>
struct MlpBwdSmem {
    static constexpr int W_BYTES = PM_N * KIN * 2;                // rows n, KIN/64 panels
    static constexpr int D_BYTES = PM_ROWS * PM_N * 2;            // dz tile: rows = points, 2 panels (also the staging tile)
    static constexpr int A_BYTES = PM_ROWS * KIN * 2;             // input-activation tile
    static constexpr int OFF_W = 0;
    static constexpr int OFF_D0 = OFF_W + W_BYTES;
    static constexpr int OFF_D1 = OFF_D0 + D_BYTES;
    static constexpr int OFF_A0 = OFF_D1 + D_BYTES;
    static constexpr int OFF_A1 = OFF_A0 + A_BYTES;
    static constexpr int OFF_PTS = OFF_A1 + A_BYTES;               // MODE 0: the tile's raw points, 2 x 128 float4
    static constexpr int OFF_MISC = OFF_PTS + 2 * PM_ROWS * 16;
    static constexpr int MISC_BYTES = 64 + 4 * (3 * 128 + 64 * 4 + 64 + 2 * 128);
    static constexpr int TOTAL = OFF_MISC + MISC_BYTES + 1024;
};

__device__ __forceinline__ void unpack8(const uint4 &u, float (&f)[8]) {
    f[0] = bf16_lo(u.x); f[1] = bf16_hi(u.x); f[2] = bf16_lo(u.y); f[3] = bf16_hi(u.y);
    f[4] = bf16_lo(u.z); f[5] = bf16_hi(u.z); f[6] = bf16_lo(u.w); f[7] = bf16_hi(u.w);
}
__device__ __forceinline__ uint4 pack8(const float (&v)[8]) {
    return make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
}

// Per-channel coefficient tables of 128 floats are read as "the 8 channels of 16-byte chunk d" by threads whose d is their
// lane id modulo 16: stored linearly that is two LDS.128 at a 32-byte lane stride, a 4-way bank conflict each (ncu: 2.7e6
// excess wavefronts per such load, 3e7 in the layer-3 backward).  Stored as two halves -- channels 8d..8d+3 of every chunk
// first, 8d+4..8d+7 behind them -- the same loads walk 16 bytes per lane: conflict-free.
__device__ __forceinline__ int coef_slot(int c) { return ((c & 4) << 4) + ((c >> 3) << 2) + (c & 3); }
__device__ __forceinline__ void coef_load8(const float *tab, int d, float (&v)[8]) {
    const float4 lo = *reinterpret_cast<const float4 *>(tab + d * 4), hi = *reinterpret_cast<const float4 *>(tab + 64 + d * 4);
    v[0] = lo.x; v[1] = lo.y; v[2] = lo.z; v[3] = lo.w; v[4] = hi.x; v[5] = hi.y; v[6] = hi.z; v[7] = hi.w;
}

template <int MODE, int KIN, int NT, bool SHARE = false>
__global__ void __launch_bounds__(NT, 1)
mlp_layer_bwd_kernel(MlpBwdArgs a) {
    static_assert(!SHARE || MODE == 1, "the share form is the layer-3 backward");
    KDF_PM_DERIVED(NT);
    using L = MlpBwdSmem<KIN>;
    constexpr int ACH = KIN / 8;                                  // 16-byte chunks per activation row
    constexpr int A_ROWS_PER_PASS = PM_THREADS / ACH;             // 16 or 32
    constexpr int A_PASSES = PM_ROWS / A_ROWS_PER_PASS;           // 8 or 4
    constexpr int D_PASSES = PM_OPASSES;                          // dz tile: 16 chunks per row, PM_OROWS rows per pass
    constexpr uint32_t PANEL = PM_ROWS * tc::ROW_BYTES;           // 16384: bytes between 64-column panels
    constexpr uint32_t IDESC_D = tc::make_idesc(PM_ROWS, KIN, 0, 1);   // dgrad: A K-major, B MN-major
    constexpr uint32_t IDESC_W = tc::make_idesc(PM_N, KIN, 1, 1);      // wgrad: both MN-major
    constexpr uint32_t TMEM_COLS = (3 * KIN > 256) ? 512 : 256;

    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = tc::align_smem_1024(smem_raw);
    uint8_t *sW = smem + L::OFF_W;
    auto sD = [smem](int b) -> uint8_t * { return smem + (b ? L::OFF_D1 : L::OFF_D0); };   // offsets from ONE shared base: LDS/STS, not generic
    auto sA = [smem](int b) -> uint8_t * { return smem + (b ? L::OFF_A1 : L::OFF_A0); };
    float4 *sPts = reinterpret_cast<float4 *>(smem + L::OFF_PTS);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + L::OFF_MISC);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + L::OFF_MISC + 16);
    float *cgs = reinterpret_cast<float *>(smem + L::OFF_MISC + 64), *cga = cgs + 128, *cgb = cga + 128;
    float *coef = cgb + 128;                                      // MODE 0: q[256], r[64];  MODE 1: scale[128], shift[128]

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t n_tiles = (a.M + PM_ROWS - 1) / PM_ROWS;

    for (int i = tid; i < 128; i += PM_THREADS) { cgs[coef_slot(i)] = a.gs[i]; cga[coef_slot(i)] = a.ga[i]; cgb[coef_slot(i)] = a.gb[i]; }
    if (MODE == 0) {                                              // q rows padded to 36 floats per 8-channel chunk: conflict-free LDS.128
        for (int i = tid; i < 64 * 4; i += PM_THREADS) coef[(i >> 5) * 36 + (i & 31)] = a.pro_a[i];
        for (int i = tid; i < 64; i += PM_THREADS) coef[288 + i] = a.pro_b[i];
    } else {
        for (int i = tid; i < KIN; i += PM_THREADS) { coef[coef_slot(i)] = a.pro_a[i]; coef[KIN + coef_slot(i)] = a.pro_b[i]; }
    }
    for (int idx = tid; idx < PM_N * ACH; idx += PM_THREADS) {
        const int n = idx / ACH, ch = idx % ACH;
        const uint4 w = *reinterpret_cast<const uint4 *>(a.W + (int64_t)n * KIN + ch * 8);
        *reinterpret_cast<uint4 *>(sW + (ch >> 3) * PANEL + tc::sw128_offset(n, ch & 7)) = w;
    }
    if (tid == 0) {
        tc::mbar_init(&bars[0], 1);
        tc::mbar_init(&bars[1], 1);
        tc::mbar_fence_init();
    }
    if (warp == 0) tc::tmem_alloc(tmem_slot, TMEM_COLS);
    tc::fence_async_smem();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;

    // fixed per-thread chunks: dz tile (dch, rows drow0 + 16p); activation tile (ach, rows arow0 + A_ROWS_PER_PASS*p)
    const int dch = tid & 15, drow0 = tid >> 4;
    const int ach = tid % ACH, arow0 = tid / ACH;

    uint4 raw_dy[D_PASSES], raw_z[D_PASSES], raw_in[A_PASSES];
    int raw_cell[D_PASSES];
    // SHARE: the share loads of a tile need its cell ids one tile earlier than everything else.  They travel through shared
    // memory by cp.async (no register waits on them: a loop-carried register copy of a fresh load stalled every thread for a
    // full memory latency per tile); sCell[k & 1] holds the ids of this CTA's k-th tile (the point tile is unused in mode 1).
    int32_t *sCell = reinterpret_cast<int32_t *>(smem + L::OFF_PTS);
    auto request_cells = [&](int64_t tile, int k) {
        if (tid < PM_ROWS) {
            const int64_t row = tile * PM_ROWS + tid;
            int32_t *dst = sCell + (k & 1) * PM_ROWS + tid;
            if (tile < n_tiles && row < a.M)
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(tc::smem_u32(dst)), "l"(a.row_cell + row) : "memory");
            else
                *dst = -1;
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    auto load_tile = [&](int64_t tile, int k = 0) {
        const int64_t r0 = tile * PM_ROWS;
        if (SHARE) {
#pragma unroll
            for (int p = 0; p < D_PASSES; ++p) {
                const int64_t row = r0 + drow0 + p * PM_OROWS;
                raw_cell[p] = sCell[(k & 1) * PM_ROWS + drow0 + p * PM_OROWS];
                if (raw_cell[p] >= 0) {                                    // (implies row < M)
                    // the share chunk and the tie byte are read at staging time (a 128-row tile meets ~5 cells: L1 hits);
                    // here they are only pulled towards L1 -- holding them in registers for a tile spilled the kernel
                    asm volatile("prefetch.global.L1 [%0];" ::"l"(a.share + (int64_t)raw_cell[p] * PM_N + dch * 8));
                    asm volatile("prefetch.global.L1 [%0];" ::"l"(a.bits + row * (PM_N / 8) + dch));
                }
                if (row < a.M) raw_z[p] = ldg_stream_u4(reinterpret_cast<const uint4 *>(a.z + row * PM_N + dch * 8));
            }
            request_cells(tile + gridDim.x, k + 1);
        } else {
#pragma unroll
            for (int p = 0; p < D_PASSES; ++p) {
                const int64_t row = r0 + drow0 + p * PM_OROWS;
                raw_cell[p] = 0;
                if (row < a.M) {
                    if (a.row_cell) raw_cell[p] = __ldg(a.row_cell + row);
                    raw_dy[p] = ldg_stream_u4(reinterpret_cast<const uint4 *>(a.dy + row * PM_N + dch * 8));
                    raw_z[p] = ldg_stream_u4(reinterpret_cast<const uint4 *>(a.z + row * PM_N + dch * 8));
                }
            }
        }
#pragma unroll
        for (int p = 0; p < A_PASSES; ++p) {
            const int64_t row = r0 + arow0 + p * A_ROWS_PER_PASS;
            if (row < a.M) {
                if (MODE == 0) raw_in[p] = ldg_stream_u4(reinterpret_cast<const uint4 *>(a.input) + row);
                else raw_in[p] = *reinterpret_cast<const uint4 *>(reinterpret_cast<const __nv_bfloat16 *>(a.input) + row * KIN + ach * 8);
            }
        }
    };
    auto l2_prefetch = [&](int64_t tile) {                                 // see the forward kernel
        if (tile >= n_tiles) return;
        const int64_t r0 = tile * PM_ROWS;
        const int64_t rows = (a.M - r0 < PM_ROWS) ? (a.M - r0) : PM_ROWS;
        if (SHARE) { if (((rows * (PM_N / 8)) & ~15) > 0) tc::prefetch_l2(a.bits + r0 * (PM_N / 8), (uint32_t)((rows * (PM_N / 8)) & ~15)); }
        else tc::prefetch_l2(a.dy + r0 * PM_N, (uint32_t)(rows * PM_N * 2));
        tc::prefetch_l2(a.z + r0 * PM_N, (uint32_t)(rows * PM_N * 2));
        // the cell ids too: without it they are the one load of the tile that still sees DRAM latency (measured: the kernel
        // ran 1.45 ms with row_cell against 1.32 ms without; 1.36 ms with this prefetch).  Skipping the dy rows of points
        // outside the grid (never written, 38 % of the rows) with per-row prefetches and loads that wait for the cell id
        // was tried and is slower (1.68 ms): the bulk prefetch of the whole tile stays.
        if (a.row_cell && ((rows * 4) & ~15) > 0) tc::prefetch_l2(a.row_cell + r0, (uint32_t)((rows * 4) & ~15));
        if (MODE == 0) tc::prefetch_l2(reinterpret_cast<const uint4 *>(a.input) + r0, (uint32_t)(rows * 16));
        else tc::prefetch_l2(reinterpret_cast<const __nv_bfloat16 *>(a.input) + r0 * KIN, (uint32_t)(rows * KIN * 2));
    };
    auto stage_tile = [&](int64_t tile, int buf) {
        const int64_t r0 = tile * PM_ROWS;
        {
            float gs[8], ga[8], gb[8];
            coef_load8(cgs, dch, gs); coef_load8(cga, dch, ga); coef_load8(cgb, dch, gb);
#pragma unroll
            for (int p = 0; p < D_PASSES; ++p) {
                const int r = drow0 + p * PM_OROWS;
                uint4 v = make_uint4(0u, 0u, 0u, 0u);
                if (r0 + r < a.M) {
                    float g[8], zz[8];
                    if (SHARE) {                                          // the row's share of its cell's gradient: only where it sits at the extreme
                        uint4 sh = make_uint4(0u, 0u, 0u, 0u);
                        uint32_t tb = 0u;
                        if (raw_cell[p] >= 0) {
                            sh = __ldg(reinterpret_cast<const uint4 *>(a.share + (int64_t)raw_cell[p] * PM_N + dch * 8));
                            tb = __ldg(a.bits + (r0 + r) * (PM_N / 8) + dch);
                        }
                        unpack8(sh, g);
#pragma unroll
                        for (int j = 0; j < 8; ++j) g[j] = (tb >> j) & 1u ? g[j] : 0.f;
                    } else {
                        unpack8(raw_cell[p] >= 0 ? raw_dy[p] : make_uint4(0u, 0u, 0u, 0u), g);
                    }
                    unpack8(raw_z[p], zz);
                    uint32_t o4[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {                          // dz = gs*dy + (gb*z + ga), two channels per instruction
                        const pm_u64 t2 = pm_fma2(pm_pk2(gb[2 * j], gb[2 * j + 1]), pm_pk2(zz[2 * j], zz[2 * j + 1]), pm_pk2(ga[2 * j], ga[2 * j + 1]));
                        const pm_u64 d2 = pm_fma2(pm_pk2(gs[2 * j], gs[2 * j + 1]), pm_pk2(g[2 * j], g[2 * j + 1]), t2);
                        float lo, hi;
                        asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(d2));
                        o4[j] = pack_bf16(lo, hi);
                    }
                    v = make_uint4(o4[0], o4[1], o4[2], o4[3]);
                }
                *reinterpret_cast<uint4 *>(sD(buf) + (dch >> 3) * PANEL + tc::sw128_offset(r, dch & 7)) = v;
            }
        }
        {
            float c0[MODE == 0 ? 32 : 8], c1[8];
            if (MODE == 0) {
#pragma unroll
                for (int j = 0; j < 32; ++j) c0[j] = coef[ach * 36 + j];
#pragma unroll
                for (int j = 0; j < 8; ++j) c1[j] = coef[288 + ach * 8 + j];
            } else {
                float t0[8], t1[8];
                coef_load8(coef, ach, t0); coef_load8(coef + KIN, ach, t1);
#pragma unroll
                for (int j = 0; j < 8; ++j) { c0[j] = t0[j]; c1[j] = t1[j]; }
            }
#pragma unroll
            for (int p = 0; p < A_PASSES; ++p) {
                const int r = arow0 + p * A_ROWS_PER_PASS;
                uint4 v = make_uint4(0u, 0u, 0u, 0u);
                if (r0 + r < a.M) {
                    if (MODE == 0) {
                        const float4 pt = make_float4(__uint_as_float(raw_in[p].x), __uint_as_float(raw_in[p].y),
                                                      __uint_as_float(raw_in[p].z), __uint_as_float(raw_in[p].w));
                        v = first_layer_chunk(pt, c0, c1);
                        if (ach == 0) sPts[buf * PM_ROWS + r] = pt;
                    } else {
                        v = affine_relu_chunk(raw_in[p], c0, c1);
                    }
                }
                *reinterpret_cast<uint4 *>(sA(buf) + (ach >> 3) * PANEL + tc::sw128_offset(r, ach & 7)) = v;
            }
        }
        tc::fence_async_smem();
    };
    auto issue_mma = [&](int buf, bool first) {                            // one thread
        const uint32_t d_base = tc::smem_u32(sD(buf)), a_base = tc::smem_u32(sA(buf)), w_base = tc::smem_u32(sW);
        const uint32_t acc_d = tmem_base + (uint32_t)buf * KIN, acc_w = tmem_base + 2u * KIN;
#pragma unroll
        for (int k = 0; k < PM_N / 16; ++k) {                              // dgrad: K = the 128 gradient channels
            const uint32_t koff = (uint32_t)(k >> 2) * PANEL + (uint32_t)(k & 3) * 32u;
            tc::mma_bf16(acc_d, tc::desc_kmajor(d_base + koff), tc::desc_mnmajor(w_base + (uint32_t)k * 2048u, PANEL), IDESC_D, k > 0);
        }
#pragma unroll
        for (int k = 0; k < PM_ROWS / 16; ++k) {                           // wgrad: K = the 128 points of the tile
            tc::mma_bf16(acc_w, tc::desc_mnmajor(d_base + (uint32_t)k * 2048u, PANEL), tc::desc_mnmajor(a_base + (uint32_t)k * 2048u, PANEL),
                         IDESC_W, !(first && k == 0));
        }
        tc::mma_commit(&bars[buf]);
    };

    // accumulators of the column sums, fixed per thread for the whole kernel
    constexpr int NACC = MODE == 1 ? 1 : 5;                                // MODE 1 accumulates channel pairs (acc2)
    float acc[NACC];
#pragma unroll
    for (int j = 0; j < NACC; ++j) acc[j] = 0.f;
    pm_u64 acc2[8] = {0ull, 0ull, 0ull, 0ull, 0ull, 0ull, 0ull, 0ull};     // MODE 1: sum dy_prev (4 pairs), sum dy_prev * z_prev (4 pairs)
    const int och = tid & 15, orow0 = tid >> 4;                            // MODE 1 store phase
    const int ccol = tid & 63, crq = tid >> 6;                             // MODE 0 column-owner phase
    constexpr int CRQ_ROWS = PM_ROWS / (PM_THREADS / 64);                  // rows per column-owner group

    auto epilogue = [&](int64_t tile, int buf, uint32_t parity) {
        const int64_t r0 = tile * PM_ROWS;
        uint8_t *sStage = sD(buf);                                         // the dz tile is dead once the MMAs have completed
        uint4 zk[MODE == 1 ? D_PASSES : 1];
        if (MODE == 1) {                                                   // z_prev chunks of the store phase (L2 hits), issued early
#pragma unroll
            for (int p = 0; p < D_PASSES; ++p) {
                const int64_t row = r0 + orow0 + p * PM_OROWS;
                zk[p] = make_uint4(0u, 0u, 0u, 0u);
                if (row < a.M) zk[p] = *reinterpret_cast<const uint4 *>(reinterpret_cast<const __nv_bfloat16 *>(a.input) + row * KIN + och * 8);
            }
        }
        tc::mbar_wait(&bars[buf], parity);
        tc::fence_after_sync();
        const int row = (warp & 3) * 32 + lane;
        if (MODE == 1) {
            constexpr int COLS_W = KIN / PM_CGROUPS;
            const uint32_t taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)buf * KIN + (uint32_t)(warp >> 2) * COLS_W;
#pragma unroll
            for (int half = 0; half < COLS_W / 32; ++half) {
                uint32_t r[32];
                tc::tmem_ld32(taddr + half * 32, r);
                tc::tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int chunk = (warp >> 2) * (COLS_W / 8) + half * 4 + j;   // 16-byte chunk of the 256-byte row
                    const uint4 av = *reinterpret_cast<const uint4 *>(sA(buf) + (chunk >> 3) * PANEL + tc::sw128_offset(row, chunk & 7));
                    // ReLU mask on the packed pairs: convert, then AND with `activation > 0` (rounding a masked-out value is moot)
                    uint4 o;
                    o.x = pack_bf16(__uint_as_float(r[8 * j + 0]), __uint_as_float(r[8 * j + 1])) & pm_gt0_mask(av.x);
                    o.y = pack_bf16(__uint_as_float(r[8 * j + 2]), __uint_as_float(r[8 * j + 3])) & pm_gt0_mask(av.y);
                    o.z = pack_bf16(__uint_as_float(r[8 * j + 4]), __uint_as_float(r[8 * j + 5])) & pm_gt0_mask(av.z);
                    o.w = pack_bf16(__uint_as_float(r[8 * j + 6]), __uint_as_float(r[8 * j + 7])) & pm_gt0_mask(av.w);
                    *reinterpret_cast<uint4 *>(sStage + row * 256 + ((chunk ^ (row & 7)) << 4)) = o;
                }
            }
        } else {
            constexpr int COLS_W = KIN / PM_CGROUPS;                        // 16 with 16 warps, 32 with 8
            const uint32_t taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)buf * KIN + (uint32_t)(warp >> 2) * COLS_W;
            uint32_t r[COLS_W];
            tc::tmem_ld<COLS_W>(taddr, r);
            tc::tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < COLS_W / 8; ++j) {
                const int chunk = (warp >> 2) * (COLS_W / 8) + j;           // 8 channels of the 64-channel activation row
                const uint4 av = *reinterpret_cast<const uint4 *>(sA(buf) + tc::sw128_offset(row, chunk));
                float act[8];
                unpack8(av, act);
#pragma unroll
                for (int h = 0; h < 2; ++h) {                               // fp32 staging: 16 chunks of 4 floats per row
                    const int c4 = chunk * 2 + h;
                    float4 o;
                    o.x = act[4 * h + 0] > 0.f ? __uint_as_float(r[8 * j + 4 * h + 0]) : 0.f;
                    o.y = act[4 * h + 1] > 0.f ? __uint_as_float(r[8 * j + 4 * h + 1]) : 0.f;
                    o.z = act[4 * h + 2] > 0.f ? __uint_as_float(r[8 * j + 4 * h + 2]) : 0.f;
                    o.w = act[4 * h + 3] > 0.f ? __uint_as_float(r[8 * j + 4 * h + 3]) : 0.f;
                    *reinterpret_cast<float4 *>(sStage + row * 256 + ((c4 ^ (row & 15)) << 4)) = o;
                }
            }
        }
        tc::fence_before_sync();
        __syncthreads();
        if (MODE == 1) {
#pragma unroll
            for (int p = 0; p < D_PASSES; ++p) {
                const int r = orow0 + p * PM_OROWS;
                if (r0 + r < a.M) {
                    const uint4 v = *reinterpret_cast<const uint4 *>(sStage + r * 256 + ((och ^ (r & 7)) << 4));
                    *reinterpret_cast<uint4 *>(a.dy_prev + (r0 + r) * PM_N + och * 8) = v;
                    const uint32_t fw[4] = {v.x, v.y, v.z, v.w}, zw[4] = {zk[p].x, zk[p].y, zk[p].z, zk[p].w};
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const pm_u64 f2 = pm_unpack2(fw[j]);
                        acc2[j] = pm_add2(acc2[j], f2);
                        acc2[4 + j] = pm_fma2(f2, pm_unpack2(zw[j]), acc2[4 + j]);
                    }
                }
            }
        } else {
            const int rows = (int)((a.M - r0 < PM_ROWS) ? (a.M - r0) : PM_ROWS);
            const int rbeg = crq * CRQ_ROWS, rend = (rbeg + CRQ_ROWS < rows) ? rbeg + CRQ_ROWS : rows;
            for (int r = rbeg; r < rend; ++r) {
                const float v = *reinterpret_cast<const float *>(sStage + r * 256 + (((ccol >> 2) ^ (r & 15)) << 4) + (ccol & 3) * 4);
                const float4 pt = sPts[buf * PM_ROWS + r];
                acc[0] += v;
                acc[1] = fmaf(v, pt.x, acc[1]); acc[2] = fmaf(v, pt.y, acc[2]);
                acc[3] = fmaf(v, pt.z, acc[3]); acc[4] = fmaf(v, pt.w, acc[4]);
            }
        }
        __syncthreads();                                                   // the staging tile is the next dz tile
    };

    // ---- software pipeline over this CTA's tiles (same schedule as the forward kernel)
    int64_t tile = blockIdx.x;
    if (SHARE) {
        request_cells(tile, 0);
        asm volatile("cp.async.wait_all;" ::: "memory");
        __syncthreads();
    }
    if (tile < n_tiles) {
        load_tile(tile, 0);
        stage_tile(tile, 0);
    }
    if (SHARE) asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();
    int it = 0;
    int64_t prev_tile = -1;
    for (; tile < n_tiles; tile += gridDim.x, ++it) {
        const int buf = it & 1;
        if (tid == 0) {
            tc::fence_after_sync();
            issue_mma(buf, it == 0);
        }
        const int64_t next = tile + gridDim.x;
        if (next < n_tiles) load_tile(next, it + 1);
        if (tid == 32) l2_prefetch(next + gridDim.x);
        if (it > 0) epilogue(prev_tile, buf ^ 1, (uint32_t)(((it - 1) >> 1) & 1));
        if (next < n_tiles) stage_tile(next, buf ^ 1);
        if (SHARE) asm volatile("cp.async.wait_all;" ::: "memory");
        __syncthreads();
        prev_tile = tile;
    }
    if (it > 0) epilogue(prev_tile, (it - 1) & 1, (uint32_t)(((it - 1) >> 1) & 1));

    // ---- column sums -> fp64 atomics
    float *red = reinterpret_cast<float *>(sD(0));
    if (MODE == 1) {
        {
            const pm_u64 lo4[4] = {acc2[0], acc2[1], acc2[2], acc2[3]}, hi4[4] = {acc2[4], acc2[5], acc2[6], acc2[7]};
            pm_store8(red + (orow0 * 16 + och) * 16, lo4);
            pm_store8(red + (orow0 * 16 + och) * 16 + 8, hi4);
        }
        __syncthreads();
        if (tid < 16 * 16) {
            const int ch = tid >> 4, j = tid & 15;
            float v = 0.f;
            for (int r = 0; r < PM_OROWS; ++r) v += red[(r * 16 + ch) * 16 + j];
            atomicAdd(a.sums + (j >> 3) * PM_N + ch * 8 + (j & 7), (double)v);
        }
    } else {
#pragma unroll
        for (int j = 0; j < 5; ++j) red[(crq * 64 + ccol) * 5 + j] = acc[j];
        __syncthreads();
        for (int i = tid; i < 64 * 5; i += PM_THREADS) {
            const int c = i / 5, j = i % 5;
            float v = 0.f;
            for (int q = 0; q < PM_THREADS / 64; ++q) v += red[(q * 64 + c) * 5 + j];
            atomicAdd(a.sums + j * 64 + c, (double)v);
        }
    }
    // ---- weight gradient of this CTA's tiles: TMEM -> fp32 atomics (every MMA completed: the last epilogue waited)
    if (it > 0) {
        tc::fence_after_sync();
        const int n = (warp & 3) * 32 + lane;
        constexpr int COLS_PER_WARP = KIN / PM_CGROUPS, LDW = COLS_PER_WARP < 32 ? COLS_PER_WARP : 32;
#pragma unroll
        for (int h = 0; h < COLS_PER_WARP / LDW; ++h) {
            const int col = (warp >> 2) * COLS_PER_WARP + h * LDW;
            uint32_t r[LDW];
            tc::tmem_ld<LDW>(tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + 2u * KIN + (uint32_t)col, r);
            tc::tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < LDW; j += 4)                                // 16-byte vector reductions: 4x fewer L2 atomic operations
                tc::red_add_v4(a.dW + n * KIN + col + j, __uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem_base, TMEM_COLS);
}


// ============================================================================= layers 2+1 backward, TMA-staged
// mlp_layer_bwd_kernel<0> spends a third of its instructions forming dz = gs*dy + ga + gb*z element by element before
// the GEMMs.  The GEMMs are linear and gs/ga/gb are per-channel, so the BatchNorm backward folds into the WEIGHTS:
//   dgrad : dA = dy.(diag(gs) W) + z.(diag(gb) W) + ga.W                        (two GEMMs + a constant row)
//   wgrad : dW = diag(gs).(dy^T a1) + diag(gb).(z^T a1) + ga (x) sum_rows(a1)   (two accumulators, combined at the end)
// and the raw bf16 dy / z tiles become tensor-core operands as they are: they arrive by TMA (SWIZZLE_128B boxes, two
// tiles in flight, no registers, no instructions), nothing is re-rounded, and the only prologue work left is the
// recompute of the first layer from the raw point.  Epilogue as before (ReLU mask, sum dy1, sum dy1 p^T), staged in bf16.
struct MlpBwd0TmaSmem {
    static constexpr int PANEL = PM_ROWS * 128;                      // 16384
    static constexpr int OFF_WS = 0;                                 // diag(gs) W2  [128 n x 64 k] bf16
    static constexpr int OFF_WB = OFF_WS + PANEL;                    // diag(gb) W2
    static constexpr int OFF_DY0 = OFF_WB + PANEL;                   // dy tile, 2 panels, x 2 stages
    static constexpr int OFF_DY1 = OFF_DY0 + 2 * PANEL;
    static constexpr int OFF_Z0 = OFF_DY1 + 2 * PANEL;               // z tile, 2 panels, x 2 stages
    static constexpr int OFF_Z1 = OFF_Z0 + 2 * PANEL;
    static constexpr int OFF_A0 = OFF_Z1 + 2 * PANEL;                // a1 tile, 1 panel, x 2 stages
    static constexpr int OFF_A1 = OFF_A0 + PANEL;
    static constexpr int OFF_STAGE = OFF_A1 + PANEL;                 // dy1 staging, bf16 [128 x 64]
    static constexpr int OFF_PTS = OFF_STAGE + PANEL;                // raw points, 3 x 128 float4 (tile i-1 is summed while tile i+1 is staged)
    static constexpr int OFF_MISC = OFF_PTS + 3 * PM_ROWS * 16;
    // 4 mbarriers + tmem slot, then floats: ga[128] gs[128] gb[128] q[8*36] r[64] cvec[64] asum[64]
    static constexpr int MISC_BYTES = 64 + 4 * (3 * 128 + 8 * 36 + 64 + 64 + 64);
    static constexpr int TOTAL = OFF_MISC + MISC_BYTES + 1024;
};

constexpr int B0T_THREADS = 512;

__global__ void __launch_bounds__(B0T_THREADS, 1)
mlp_layer_bwd0_tma_kernel(MlpBwdArgs a, const __grid_constant__ CUtensorMap tm_dy, const __grid_constant__ CUtensorMap tm_z) {
    using L = MlpBwd0TmaSmem;
    constexpr int KIN = 64;
    constexpr uint32_t PANEL = L::PANEL;
    constexpr uint32_t IDESC_D = tc::make_idesc(PM_ROWS, KIN, 0, 1);       // dgrad: A K-major, B MN-major
    constexpr uint32_t IDESC_W = tc::make_idesc(PM_N, KIN, 1, 1);          // wgrad: both MN-major
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = tc::align_smem_1024(smem_raw);
    uint8_t *sWs = smem + L::OFF_WS, *sWb = smem + L::OFF_WB;
    auto sDy = [smem](int b) -> uint8_t * { return smem + (b ? L::OFF_DY1 : L::OFF_DY0); };
    auto sZ = [smem](int b) -> uint8_t * { return smem + (b ? L::OFF_Z1 : L::OFF_Z0); };
    auto sA = [smem](int b) -> uint8_t * { return smem + (b ? L::OFF_A1 : L::OFF_A0); };
    uint8_t *sStage = smem + L::OFF_STAGE;
    float4 *sPts = reinterpret_cast<float4 *>(smem + L::OFF_PTS);
    uint64_t *bar_raw = reinterpret_cast<uint64_t *>(smem + L::OFF_MISC);            // [2]
    uint64_t *bar_mma = bar_raw + 2;                                                 // [2]
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + L::OFF_MISC + 48);
    float *cga = reinterpret_cast<float *>(smem + L::OFF_MISC + 64), *cgs = cga + 128, *cgb = cgs + 128;
    float *coef = cgb + 128;                                         // q: 8 chunks x 36, r at 288
    float *cvec = coef + 8 * 36 + 64, *asum = cvec + 64;
    float *red = reinterpret_cast<float *>(sStage);                  // end-of-kernel reductions reuse the staging tile (4096 floats)

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t n_tiles = (a.M + PM_ROWS - 1) / PM_ROWS;
    auto issue_tile = [&](int64_t tile, int buf) {                   // one thread: dy and z tiles, 2 boxes each
        tma::mbar_expect_tx(&bar_raw[buf], 4 * PANEL);
        const int r0 = (int)(tile * PM_ROWS);
        tma::load_2d(sDy(buf), &tm_dy, 0, r0, &bar_raw[buf]);
        tma::load_2d(sDy(buf) + PANEL, &tm_dy, 64, r0, &bar_raw[buf]);
        tma::load_2d(sZ(buf), &tm_z, 0, r0, &bar_raw[buf]);
        tma::load_2d(sZ(buf) + PANEL, &tm_z, 64, r0, &bar_raw[buf]);
    };
    if (tid == 0) {
        tc::mbar_init(&bar_raw[0], 1); tc::mbar_init(&bar_raw[1], 1);
        tc::mbar_init(&bar_mma[0], 1); tc::mbar_init(&bar_mma[1], 1);
        tc::mbar_fence_init();
        tc::fence_async_smem();
        if ((int64_t)blockIdx.x < n_tiles) issue_tile(blockIdx.x, 0);
        if ((int64_t)blockIdx.x + gridDim.x < n_tiles) issue_tile((int64_t)blockIdx.x + gridDim.x, 1);
    }
    for (int i = tid; i < 128; i += B0T_THREADS) { cgs[i] = a.gs[i]; cga[i] = a.ga[i]; cgb[i] = a.gb[i]; }
    for (int i = tid; i < 64 * 4; i += B0T_THREADS) coef[(i >> 5) * 36 + (i & 31)] = a.pro_a[i];
    for (int i = tid; i < 64; i += B0T_THREADS) coef[288 + i] = a.pro_b[i];
    // the two scaled copies of W2 (rounded once, from the bf16 weight the forward used)
    for (int idx = tid; idx < PM_N * 8; idx += B0T_THREADS) {
        const int n = idx >> 3, ch = idx & 7;
        float w[8], ws[8], wb[8];
        unpack8(*reinterpret_cast<const uint4 *>(a.W + (int64_t)n * KIN + ch * 8), w);
        const float s = __ldg(a.gs + n), b = __ldg(a.gb + n);
#pragma unroll
        for (int j = 0; j < 8; ++j) { ws[j] = s * w[j]; wb[j] = b * w[j]; }
        *reinterpret_cast<uint4 *>(sWs + tc::sw128_offset(n, ch)) = pack8(ws);
        *reinterpret_cast<uint4 *>(sWb + tc::sw128_offset(n, ch)) = pack8(wb);
    }
    if (tid < 64) {                                                  // the constant row of the data gradient: ga . W2
        float v = 0.f;
        for (int n = 0; n < PM_N; ++n) v = fmaf(__ldg(a.ga + n), __bfloat162float(a.W[(int64_t)n * KIN + tid]), v);
        cvec[tid] = v;
    }
    if (warp == 0) tc::tmem_alloc(tmem_slot, 256);
    tc::fence_async_smem();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;

    // activation tile: 8 chunks per row, 64 rows per pass, 2 passes
    const int ach = tid & 7, arow0 = tid >> 3;
    float c0[32], c1[8];
#pragma unroll
    for (int j = 0; j < 32; ++j) c0[j] = coef[ach * 36 + j];
#pragma unroll
    for (int j = 0; j < 8; ++j) c1[j] = coef[288 + ach * 8 + j];
    float a_sum[8];                                                  // column sums of a1 over this CTA's rows (for ga (x) sum a1)
#pragma unroll
    for (int j = 0; j < 8; ++j) a_sum[j] = 0.f;
    uint4 rpt[2];
    auto load_pts = [&](int64_t tile) {
#pragma unroll
        for (int p = 0; p < 2; ++p) {
            const int64_t row = tile * PM_ROWS + arow0 + p * 64;
            if (tile < n_tiles && row < a.M) rpt[p] = ldg_stream_u4(reinterpret_cast<const uint4 *>(a.input) + row);
        }
    };
    auto stage_a = [&](int64_t tile, int buf, int slot) {
        const int64_t r0 = tile * PM_ROWS;
#pragma unroll
        for (int p = 0; p < 2; ++p) {
            const int r = arow0 + p * 64;
            uint4 v = make_uint4(0u, 0u, 0u, 0u);
            if (r0 + r < a.M) {
                const float4 pt = make_float4(__uint_as_float(rpt[p].x), __uint_as_float(rpt[p].y),
                                              __uint_as_float(rpt[p].z), __uint_as_float(rpt[p].w));
                v = first_layer_chunk(pt, c0, c1);
                float f[8];
                unpack8(v, f);
#pragma unroll
                for (int j = 0; j < 8; ++j) a_sum[j] += f[j];
                if (ach == 0) sPts[slot * PM_ROWS + r] = pt;
            }
            *reinterpret_cast<uint4 *>(sA(buf) + tc::sw128_offset(r, ach)) = v;
        }
        tc::fence_async_smem();
    };
    auto issue_mma = [&](int buf, bool first, uint32_t raw_parity) {     // one thread
        tc::mbar_wait(&bar_raw[buf], raw_parity);                        // dy / z of this tile have landed
        tc::fence_after_sync();
        const uint32_t dy = tc::smem_u32(sDy(buf)), z = tc::smem_u32(sZ(buf)), act = tc::smem_u32(sA(buf));
        const uint32_t ws = tc::smem_u32(sWs), wb = tc::smem_u32(sWb);
        const uint32_t acc_d = tmem_base + (uint32_t)buf * KIN, acc_ws = tmem_base + 2u * KIN, acc_wb = tmem_base + 3u * KIN;
#pragma unroll
        for (int k = 0; k < PM_N / 16; ++k) {                            // dgrad: K = the 128 gradient channels, twice
            const uint32_t koff = (uint32_t)(k >> 2) * PANEL + (uint32_t)(k & 3) * 32u;
            tc::mma_bf16(acc_d, tc::desc_kmajor(dy + koff), tc::desc_mnmajor(ws + (uint32_t)k * 2048u, PANEL), IDESC_D, k > 0);
        }
#pragma unroll
        for (int k = 0; k < PM_N / 16; ++k) {
            const uint32_t koff = (uint32_t)(k >> 2) * PANEL + (uint32_t)(k & 3) * 32u;
            tc::mma_bf16(acc_d, tc::desc_kmajor(z + koff), tc::desc_mnmajor(wb + (uint32_t)k * 2048u, PANEL), IDESC_D, true);
        }
#pragma unroll
        for (int k = 0; k < PM_ROWS / 16; ++k) {                         // wgrad: K = the 128 points of the tile, two accumulators
            tc::mma_bf16(acc_ws, tc::desc_mnmajor(dy + (uint32_t)k * 2048u, PANEL), tc::desc_mnmajor(act + (uint32_t)k * 2048u, PANEL),
                         IDESC_W, !(first && k == 0));
        }
#pragma unroll
        for (int k = 0; k < PM_ROWS / 16; ++k) {
            tc::mma_bf16(acc_wb, tc::desc_mnmajor(z + (uint32_t)k * 2048u, PANEL), tc::desc_mnmajor(act + (uint32_t)k * 2048u, PANEL),
                         IDESC_W, !(first && k == 0));
        }
        tc::mma_commit(&bar_mma[buf]);
    };

    float acc[5];
#pragma unroll
    for (int j = 0; j < 5; ++j) acc[j] = 0.f;
    const int ccol = tid & 63, crq = tid >> 6;                           // column-owner phase: 8 groups x 16 rows
    // part 1: accumulator -> ReLU mask -> bf16 staging tile;  part 2 (after a CTA barrier): column sums from the staging tile.
    // Part 2 shares its barrier interval with the staging of the NEXT tile's activation (two barriers per tile instead of three:
    // 43 % of this kernel's stall samples sat at CTA barriers), which is why the raw points are triple-buffered.
    auto epilogue1 = [&](int64_t tile, int buf, uint32_t parity, int64_t refill_tile) {
        tc::mbar_wait(&bar_mma[buf], parity);
        tc::fence_after_sync();
        // the tensor cores are done with this stage's dy / z tiles: refill them right away
        if (tid == 0 && refill_tile < n_tiles) issue_tile(refill_tile, buf);
        const int row = (warp & 3) * 32 + lane;
        {
            const int col0 = (warp >> 2) * 16;                           // 16 warps: 4 column groups of 16
            uint32_t r[16];
            tc::tmem_ld16(tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)buf * KIN + (uint32_t)col0, r);
            tc::tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int chunk = (col0 >> 3) + j;
                const uint4 av = *reinterpret_cast<const uint4 *>(sA(buf) + tc::sw128_offset(row, chunk));
                const uint32_t aw[4] = {av.x, av.y, av.z, av.w};
                uint32_t o[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {                            // (accumulator + constant row) on pairs, ReLU mask as an AND
                    const pm_u64 v2 = pm_add2(pm_pk2(__uint_as_float(r[8 * j + 2 * e]), __uint_as_float(r[8 * j + 2 * e + 1])),
                                              pm_pk2(cvec[chunk * 8 + 2 * e], cvec[chunk * 8 + 2 * e + 1]));
                    float lo, hi;
                    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v2));
                    o[e] = pack_bf16(lo, hi) & pm_gt0_mask(aw[e]);
                }
                *reinterpret_cast<uint4 *>(sStage + tc::sw128_offset(row, chunk)) = make_uint4(o[0], o[1], o[2], o[3]);
            }
        }
        tc::fence_before_sync();
    };
    auto epilogue2 = [&](int64_t tile, int slot) {
        const int64_t r0 = tile * PM_ROWS;
        const int rows = (int)((a.M - r0 < PM_ROWS) ? (a.M - r0) : PM_ROWS);
        const int rbeg = crq * 16, rend = (rbeg + 16 < rows) ? rbeg + 16 : rows;
        for (int r = rbeg; r < rend; ++r) {
            const __nv_bfloat16 hv = *reinterpret_cast<const __nv_bfloat16 *>(sStage + tc::sw128_offset(r, ccol >> 3) + (ccol & 7) * 2);
            const float v = __bfloat162float(hv);
            const float4 pt = sPts[slot * PM_ROWS + r];
            acc[0] += v;
            acc[1] = fmaf(v, pt.x, acc[1]); acc[2] = fmaf(v, pt.y, acc[2]);
            acc[3] = fmaf(v, pt.z, acc[3]); acc[4] = fmaf(v, pt.w, acc[4]);
        }
    };

    // ---- pipeline: MMA(i) | points(i+1) in flight | epilogue(i-1) + TMA refill | a1(i+1)
    int64_t tile = blockIdx.x;
    if (tile < n_tiles) {
        load_pts(tile);
        stage_a(tile, 0, 0);
    }
    __syncthreads();
    int it = 0;
    int64_t prev_tile = -1;
    for (; tile < n_tiles; tile += gridDim.x, ++it) {
        const int buf = it & 1;
        if (tid == 0) issue_mma(buf, it == 0, (uint32_t)((it >> 1) & 1));
        const int64_t next = tile + gridDim.x;
        load_pts(next);
        if (it > 0) epilogue1(prev_tile, buf ^ 1, (uint32_t)(((it - 1) >> 1) & 1), next);  // refills stage buf^1 with tile i+1
        __syncthreads();                                                 // staging tile written; the a1 tile of tile i-1 has been read
        if (it > 0) epilogue2(prev_tile, (it + 2) % 3);                  // points of tile i-1: slot (i-1) mod 3
        if (next < n_tiles) stage_a(next, buf ^ 1, (it + 1) % 3);
        __syncthreads();
        prev_tile = tile;
    }
    if (it > 0) {
        epilogue1(prev_tile, (it - 1) & 1, (uint32_t)(((it - 1) >> 1) & 1), n_tiles);
        __syncthreads();
        epilogue2(prev_tile, (it + 2) % 3);
        __syncthreads();
    }

    // ---- column sums -> fp64 atomics
#pragma unroll
    for (int j = 0; j < 5; ++j) red[tid * 5 + j] = acc[j];
    __syncthreads();
    for (int i = tid; i < 64 * 5; i += B0T_THREADS) {
        const int c = i / 5, j = i % 5;
        float v = 0.f;
        for (int q = 0; q < B0T_THREADS / 64; ++q) v += red[(q * 64 + c) * 5 + j];
        atomicAdd(a.sums + j * 64 + c, (double)v);
    }
    __syncthreads();
    // ---- sum_rows a1 of this CTA (threads sharing a chunk: tid & 7), then the weight gradient
#pragma unroll
    for (int j = 0; j < 8; ++j) red[tid * 8 + j] = a_sum[j];
    __syncthreads();
    if (tid < 64) {
        const int ch = tid >> 3, j = tid & 7;
        float v = 0.f;
        for (int t = ch; t < B0T_THREADS; t += 8) v += red[t * 8 + j];
        asum[tid] = v;
    }
    __syncthreads();
    if (it > 0) {
        tc::fence_after_sync();
        const int n = (warp & 3) * 32 + lane;
        const int col0 = (warp >> 2) * 16;
        uint32_t rs[16], rb[16];
        tc::tmem_ld16(tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + 2u * KIN + (uint32_t)col0, rs);
        tc::tmem_ld16(tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + 3u * KIN + (uint32_t)col0, rb);
        tc::tmem_ld_wait();
        const float gs = cgs[n], gb = cgb[n], ga = cga[n];
        float o[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) o[j] = fmaf(gs, __uint_as_float(rs[j]), fmaf(gb, __uint_as_float(rb[j]), ga * asum[col0 + j]));
#pragma unroll
        for (int j = 0; j < 16; j += 4) tc::red_add_v4(a.dW + n * KIN + col0 + j, o[j], o[j + 1], o[j + 2], o[j + 3]);
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem_base, 256);
}

// mean / invstd / folded scale,shift / running statistics from fp64 column sums
__global__ void bn_finalize_kernel(const double *stats, int64_t M, int C, const float *gamma, const float *beta,
                                   const float *pre_bias, float eps, float momentum, float *running_mean,
                                   float *running_var, float *mean, float *invstd, float *scale, float *shift) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const double invM = 1.0 / (double)M;
    const double mu = stats[c] * invM;
    double var = stats[C + c] * invM - mu * mu;
    if (var < 0.0) var = 0.0;
    const float is = (float)(1.0 / sqrt(var + (double)eps));
    mean[c] = (float)mu;
    invstd[c] = is;
    const float s = (gamma ? gamma[c] : 1.f) * is;
    scale[c] = s;
    shift[c] = (beta ? beta[c] : 0.f) - (float)mu * s;
    if (running_mean) {
        const double unb = M > 1 ? var * ((double)M / (double)(M - 1)) : var;
        const float bias = pre_bias ? pre_bias[c] : 0.f;
        running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * ((float)mu + bias);
        running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unb;
    }
}


// ----------------------------------------------------------------------------- per-channel algebra between the layer kernels
// (one tiny launch each instead of dozens of 64/128-element torch ops)

// Layer 1 is linear in the raw point, z1 = W1.p (+b1, cancelled by the batch statistics): its BatchNorm statistics
// follow from the 14 point moments.  Outputs the folded first layer q = scale1*W1, r = shift1 the prologues use, and
// mean / invstd / scale for the backward; advances the running statistics like nn.BatchNorm1d.
__global__ void mlp_l1_stats_kernel(const double *__restrict__ m14, int64_t M, const float *__restrict__ W1 /* [64,4] */,
                                    const float *__restrict__ b1, const float *__restrict__ gamma, const float *__restrict__ beta,
                                    float eps, float momentum, float *running_mean, float *running_var,
                                    float *__restrict__ q, float *__restrict__ r, float *__restrict__ mean,
                                    float *__restrict__ invstd, float *__restrict__ scale) {
    const int c = threadIdx.x;
    if (c >= 64) return;
    const double invM = 1.0 / (double)M;
    double mu[4], cov[4][4];
    const int idx[4][4] = {{4, 5, 6, 7}, {5, 8, 9, 10}, {6, 9, 11, 12}, {7, 10, 12, 13}};
    for (int i = 0; i < 4; ++i) mu[i] = m14[i] * invM;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) cov[i][j] = m14[idx[i][j]] * invM - mu[i] * mu[j];
    double w[4];
    for (int i = 0; i < 4; ++i) w[i] = (double)W1[c * 4 + i];
    double mn = 0.0, var = 0.0;
    for (int i = 0; i < 4; ++i) {
        mn += w[i] * mu[i];
        double t = 0.0;
        for (int j = 0; j < 4; ++j) t += cov[i][j] * w[j];
        var += w[i] * t;
    }
    if (var < 0.0) var = 0.0;
    const double is = 1.0 / sqrt(var + (double)eps);
    const double sc = (double)gamma[c] * is;
    const double sh = (double)beta[c] - mn * sc;
    mean[c] = (float)mn;
    invstd[c] = (float)is;
    scale[c] = (float)sc;
    r[c] = (float)sh;
    for (int i = 0; i < 4; ++i) q[c * 4 + i] = (float)(sc * w[i]);
    if (running_mean) {
        const double unb = M > 1 ? var * ((double)M / (double)(M - 1)) : var;
        running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)(mn + (double)b1[c]);
        running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unb;
    }
}

// BatchNorm backward through the batch statistics as per-channel coefficients: dz = gs*dy + ga + gb*z, plus
// dgamma / dbeta, from S0 = sum dy and S1 = sum dy*z.
__global__ void bn_bwd_coeffs_kernel(const double *__restrict__ sums /* [2][C] */, int C, int64_t M,
                                     const float *__restrict__ mean, const float *__restrict__ invstd,
                                     const float *__restrict__ scale, float *__restrict__ gs, float *__restrict__ ga,
                                     float *__restrict__ gb, float *__restrict__ dgamma, float *__restrict__ dbeta) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const double S0 = sums[c], S1 = sums[C + c], mn = mean[c], is = invstd[c], sc = scale[c];
    const double dg = is * (S1 - mn * S0);
    const double b = -(sc * is * dg) / (double)M;
    const double a = -(sc * S0) / (double)M - b * mn;
    gs[c] = (float)sc; ga[c] = (float)a; gb[c] = (float)b;
    dgamma[c] = (float)dg; dbeta[c] = (float)S0;
}

// Layer 1 backward in closed form from S0 = sum dy1, T = sum dy1 p^T (sums [5][64]) and the point moments:
// sum dy1*z1 = W1[c,:].T[c,:];  dW1 = gs*T + ga*sum p^T + gb*W1.(sum p p^T).
__global__ void mlp_l1_bwd_kernel(const double *__restrict__ sums /* [5][64] */, const double *__restrict__ m14, int64_t M,
                                  const float *__restrict__ W1, const float *__restrict__ mean, const float *__restrict__ invstd,
                                  const float *__restrict__ scale, float *__restrict__ dW1 /* [64,4] */,
                                  float *__restrict__ dgamma, float *__restrict__ dbeta) {
    const int c = threadIdx.x;
    if (c >= 64) return;
    const int idx[4][4] = {{4, 5, 6, 7}, {5, 8, 9, 10}, {6, 9, 11, 12}, {7, 10, 12, 13}};
    double w[4], T[4];
    for (int i = 0; i < 4; ++i) { w[i] = (double)W1[c * 4 + i]; T[i] = sums[(1 + i) * 64 + c]; }
    const double S0 = sums[c];
    double S1 = 0.0;
    for (int i = 0; i < 4; ++i) S1 += w[i] * T[i];
    const double mn = mean[c], is = invstd[c], sc = scale[c];
    const double dg = is * (S1 - mn * S0);
    const double b = -(sc * is * dg) / (double)M;
    const double a = -(sc * S0) / (double)M - b * mn;
    for (int j = 0; j < 4; ++j) {
        double wpp = 0.0;
        for (int i = 0; i < 4; ++i) wpp += w[i] * m14[idx[i][j]];
        dW1[c * 4 + j] = (float)(sc * T[j] + a * m14[j] + b * wpp);
    }
    dgamma[c] = (float)dg;
    dbeta[c] = (float)S0;
}

}  // namespace kdf

using namespace kdf;

extern "C" {

int kdf_mlp_layer_fwd(int mode, const void *input, int64_t M, const float *pro_a, const float *pro_b,
                      const void *W_bf16, int Kin, int Nout, void *z_out, double *stats, void *stream) {
    KDF_CHECK_ARG(mode == 0 || mode == 1, "mlp_layer_fwd: bad mode %d", mode);
    KDF_CHECK_ARG(M >= 0, "mlp_layer_fwd: negative M");
    KDF_CHECK_ARG(Nout == PM_N, "mlp_layer_fwd: Nout must be %d", PM_N);
    KDF_CHECK_ARG((mode == 0 && Kin == 64) || (mode == 1 && Kin == 128), "mlp_layer_fwd: unsupported (mode, Kin) = (%d, %d)", mode, Kin);
    KDF_CHECK_ARG(pro_a && pro_b && W_bf16 && stats, "mlp_layer_fwd: null pointer");
    cudaStream_t st = as_stream(stream);
    KDF_CUDA(cudaMemsetAsync(stats, 0, sizeof(double) * 2 * PM_N, st));
    if (M == 0) return KDF_OK;
    KDF_CHECK_ARG(input && z_out, "mlp_layer_fwd: null pointer");
    KDF_CHECK_ARG(((reinterpret_cast<uintptr_t>(input) | reinterpret_cast<uintptr_t>(z_out) | reinterpret_cast<uintptr_t>(W_bf16)) & 15) == 0,
                  "mlp_layer_fwd: buffers must be 16-byte aligned");
    MlpFwdArgs a{input, M, pro_a, pro_b, reinterpret_cast<const __nv_bfloat16 *>(W_bf16),
                 reinterpret_cast<__nv_bfloat16 *>(z_out), stats};
    const int64_t n_tiles = (M + PM_ROWS - 1) / PM_ROWS;
    int blocks = sm_count();
    if (n_tiles < blocks) blocks = (int)n_tiles;
    if (mode == 0) {
        const int smem = MlpSmem<64>::TOTAL;
        // two co-resident CTAs per SM (measured 0.447 -> 0.416 ms against one)
        blocks = 2 * sm_count();
        if (n_tiles < blocks) blocks = (int)n_tiles;
        KDF_CUDA(cudaFuncSetAttribute(mlp_layer_fwd_kernel<0, 64, 256, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        mlp_layer_fwd_kernel<0, 64, 256, 2><<<blocks, 256, smem, st>>>(a);
    } else {
        const int smem = MlpSmem<128>::TOTAL;
        KDF_CUDA(cudaFuncSetAttribute(mlp_layer_fwd_kernel<1, 128, 512>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        mlp_layer_fwd_kernel<1, 128, 512><<<blocks, 512, smem, st>>>(a);
    }
    KDF_LAUNCH_CHECK();
    return KDF_OK;
}

int kdf_mlp_eval3_fwd(const float *points, int64_t M, const float *q, const float *r, const void *W2_bf16,
                      const float *scale2, const float *shift2, const void *W3_bf16, void *z3_out, void *stream) {
    KDF_CHECK_ARG(M >= 0, "mlp_eval3_fwd: negative M");
    KDF_CHECK_ARG(q && r && W2_bf16 && scale2 && shift2 && W3_bf16, "mlp_eval3_fwd: null pointer");
    if (M == 0) return KDF_OK;
    KDF_CHECK_ARG(points && z3_out, "mlp_eval3_fwd: null pointer");
    KDF_CHECK_ARG(((reinterpret_cast<uintptr_t>(points) | reinterpret_cast<uintptr_t>(z3_out) | reinterpret_cast<uintptr_t>(W2_bf16) |
                    reinterpret_cast<uintptr_t>(W3_bf16)) & 15) == 0, "mlp_eval3_fwd: buffers must be 16-byte aligned");
    cudaStream_t st = as_stream(stream);
    MlpEvalArgs a{points, M, q, r, reinterpret_cast<const __nv_bfloat16 *>(W2_bf16), scale2, shift2,
                  reinterpret_cast<const __nv_bfloat16 *>(W3_bf16), reinterpret_cast<__nv_bfloat16 *>(z3_out)};
    const int64_t n_tiles = (M + PM_ROWS - 1) / PM_ROWS;
    int blocks = 2 * sm_count();                                        // two co-resident CTAs per SM
    if (n_tiles < blocks) blocks = (int)n_tiles;
    const int smem = MlpEvalSmem::TOTAL;
    KDF_CUDA(cudaFuncSetAttribute(mlp_eval3_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    mlp_eval3_kernel<256><<<blocks, 256, smem, st>>>(a);
    KDF_LAUNCH_CHECK();
    return KDF_OK;
}

int kdf_mlp_layer_bwd(int mode, const void *dy, const void *z, const float *gs, const float *ga, const float *gb,
                      const void *input, int64_t M, const float *pro_a, const float *pro_b, const void *W_bf16,
                      int Kin, void *dy_prev, double *sums, float *dW, const int32_t *row_cell, void *stream) {
    KDF_CHECK_ARG(mode == 0 || mode == 1, "mlp_layer_bwd: bad mode %d", mode);
    KDF_CHECK_ARG(M >= 0, "mlp_layer_bwd: negative M");
    KDF_CHECK_ARG((mode == 0 && Kin == 64) || (mode == 1 && Kin == 128), "mlp_layer_bwd: unsupported (mode, Kin) = (%d, %d)", mode, Kin);
    KDF_CHECK_ARG(gs && ga && gb && pro_a && pro_b && W_bf16 && sums && dW, "mlp_layer_bwd: null pointer");
    cudaStream_t st = as_stream(stream);
    KDF_CUDA(cudaMemsetAsync(sums, 0, sizeof(double) * (mode == 1 ? 2 * PM_N : 5 * 64), st));
    KDF_CUDA(cudaMemsetAsync(dW, 0, sizeof(float) * PM_N * Kin, st));
    if (M == 0) return KDF_OK;
    KDF_CHECK_ARG(dy && z && input && (mode == 0 || dy_prev), "mlp_layer_bwd: null pointer");
    KDF_CHECK_ARG(((reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(z) | reinterpret_cast<uintptr_t>(input) |
                    reinterpret_cast<uintptr_t>(dy_prev) | reinterpret_cast<uintptr_t>(W_bf16)) & 15) == 0,
                  "mlp_layer_bwd: buffers must be 16-byte aligned");
    MlpBwdArgs a{reinterpret_cast<const __nv_bfloat16 *>(dy), reinterpret_cast<const __nv_bfloat16 *>(z), gs, ga, gb, input,
                 pro_a, pro_b, reinterpret_cast<const __nv_bfloat16 *>(W_bf16), M,
                 reinterpret_cast<__nv_bfloat16 *>(dy_prev), sums, dW, row_cell, nullptr, nullptr};
    const int64_t n_tiles = (M + PM_ROWS - 1) / PM_ROWS;
    int blocks = sm_count();
    if (n_tiles < blocks) blocks = (int)n_tiles;
    // mode 0: TMA-staged kernel; the register-staged one serves callers that mask rows by cell id (row_cell)
    CUtensorMap tm_dy, tm_z;
    if (mode == 0 && row_cell == nullptr && tma::make_row_map(&tm_dy, dy, M, PM_N) && tma::make_row_map(&tm_z, z, M, PM_N)) {
        KDF_CUDA(cudaFuncSetAttribute(mlp_layer_bwd0_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, MlpBwd0TmaSmem::TOTAL));
        mlp_layer_bwd0_tma_kernel<<<blocks, B0T_THREADS, MlpBwd0TmaSmem::TOTAL, st>>>(a, tm_dy, tm_z);
        KDF_LAUNCH_CHECK();
        return KDF_OK;
    }
    if (mode == 0) {
        const int smem = MlpBwdSmem<64>::TOTAL;
        KDF_CUDA(cudaFuncSetAttribute(mlp_layer_bwd_kernel<0, 64, 512>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        mlp_layer_bwd_kernel<0, 64, 512><<<blocks, 512, smem, st>>>(a);
    } else {
        const int smem = MlpBwdSmem<128>::TOTAL;
        KDF_CUDA(cudaFuncSetAttribute(mlp_layer_bwd_kernel<1, 128, 512>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        mlp_layer_bwd_kernel<1, 128, 512><<<blocks, 512, smem, st>>>(a);
    }
    KDF_LAUNCH_CHECK();
    return KDF_OK;
}

// Layer-3 backward over cell-sorted rows: the gradient rows are formed from (row_cell = global cell id per row,
// share bf16 [cells,128], bits u8 [M,16]) as kdf_bev_bwd_share leaves them; everything else as kdf_mlp_layer_bwd mode 1.
int kdf_mlp_layer_bwd_share(const int32_t *row_cell, const void *share, const void *bits, const void *z,
                            const float *gs, const float *ga, const float *gb, const void *z_prev, int64_t M,
                            const float *pro_a, const float *pro_b, const void *W_bf16, void *dy_prev, double *sums, float *dW,
                            void *stream) {
    KDF_CHECK_ARG(M >= 0, "mlp_layer_bwd_share: negative M");
    KDF_CHECK_ARG(gs && ga && gb && pro_a && pro_b && W_bf16 && sums && dW, "mlp_layer_bwd_share: null pointer");
    cudaStream_t st = as_stream(stream);
    KDF_CUDA(cudaMemsetAsync(sums, 0, sizeof(double) * 2 * PM_N, st));
    KDF_CUDA(cudaMemsetAsync(dW, 0, sizeof(float) * PM_N * 128, st));
    if (M == 0) return KDF_OK;
    KDF_CHECK_ARG(row_cell && share && bits && z && z_prev && dy_prev, "mlp_layer_bwd_share: null pointer");
    KDF_CHECK_ARG(((reinterpret_cast<uintptr_t>(share) | reinterpret_cast<uintptr_t>(bits) | reinterpret_cast<uintptr_t>(z) |
                    reinterpret_cast<uintptr_t>(z_prev) | reinterpret_cast<uintptr_t>(dy_prev) | reinterpret_cast<uintptr_t>(W_bf16)) & 15) == 0,
                  "mlp_layer_bwd_share: buffers must be 16-byte aligned");
    MlpBwdArgs a{nullptr, reinterpret_cast<const __nv_bfloat16 *>(z), gs, ga, gb, z_prev, pro_a, pro_b,
                 reinterpret_cast<const __nv_bfloat16 *>(W_bf16), M, reinterpret_cast<__nv_bfloat16 *>(dy_prev), sums, dW, row_cell,
                 reinterpret_cast<const __nv_bfloat16 *>(share), reinterpret_cast<const uint8_t *>(bits)};
    const int64_t n_tiles = (M + PM_ROWS - 1) / PM_ROWS;
    int blocks = sm_count();
    if (n_tiles < blocks) blocks = (int)n_tiles;
    const int smem = MlpBwdSmem<128>::TOTAL;
    KDF_CUDA(cudaFuncSetAttribute(mlp_layer_bwd_kernel<1, 128, 512, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    mlp_layer_bwd_kernel<1, 128, 512, true><<<blocks, 512, smem, st>>>(a);
    KDF_LAUNCH_CHECK();
    return KDF_OK;
}

int kdf_bn_finalize(const double *stats, int64_t M, int C, const float *gamma, const float *beta, const float *pre_bias,
                    float eps, float momentum, float *running_mean, float *running_var,
                    float *mean, float *invstd, float *scale, float *shift, void *stream) {
    KDF_CHECK_ARG(M > 0 && C > 0, "bn_finalize: bad sizes");
    KDF_CHECK_ARG(stats && mean && invstd && scale && shift, "bn_finalize: null pointer");
    KDF_CHECK_ARG((running_mean == nullptr) == (running_var == nullptr), "bn_finalize: running stats come in pairs");
    bn_finalize_kernel<<<(C + 127) / 128, 128, 0, as_stream(stream)>>>(stats, M, C, gamma, beta, pre_bias, eps, momentum,
                                                                      running_mean, running_var, mean, invstd, scale, shift);
    KDF_LAUNCH_CHECK();
    return KDF_OK;
}

int kdf_mlp_l1_stats(const double *moments14, int64_t M, const float *W1, const float *b1, const float *gamma,
                     const float *beta, float eps, float momentum, float *running_mean, float *running_var,
                     float *q, float *r, float *mean, float *invstd, float *scale, void *stream) {
    KDF_CHECK_ARG(M > 0, "mlp_l1_stats: M must be positive");
    KDF_CHECK_ARG(moments14 && W1 && gamma && beta && q && r && mean && invstd && scale, "mlp_l1_stats: null pointer");
    KDF_CHECK_ARG((running_mean == nullptr) == (running_var == nullptr) && (running_mean == nullptr || b1 != nullptr),
                  "mlp_l1_stats: running statistics come in pairs and need the conv bias");
    mlp_l1_stats_kernel<<<1, 64, 0, as_stream(stream)>>>(moments14, M, W1, b1, gamma, beta, eps, momentum, running_mean,
                                                          running_var, q, r, mean, invstd, scale);
    KDF_LAUNCH_CHECK();
    return KDF_OK;
}

int kdf_bn_bwd_coeffs(const double *sums, int C, int64_t M, const float *mean, const float *invstd, const float *scale,
                      float *gs, float *ga, float *gb, float *dgamma, float *dbeta, void *stream) {
    KDF_CHECK_ARG(M > 0 && C > 0, "bn_bwd_coeffs: bad sizes");
    KDF_CHECK_ARG(sums && mean && invstd && scale && gs && ga && gb && dgamma && dbeta, "bn_bwd_coeffs: null pointer");
    bn_bwd_coeffs_kernel<<<(C + 127) / 128, 128, 0, as_stream(stream)>>>(sums, C, M, mean, invstd, scale, gs, ga, gb, dgamma, dbeta);
    KDF_LAUNCH_CHECK();
    return KDF_OK;
}

int kdf_mlp_l1_bwd(const double *sums5x64, const double *moments14, int64_t M, const float *W1, const float *mean,
                   const float *invstd, const float *scale, float *dW1, float *dgamma, float *dbeta, void *stream) {
    KDF_CHECK_ARG(M > 0, "mlp_l1_bwd: M must be positive");
    KDF_CHECK_ARG(sums5x64 && moments14 && W1 && mean && invstd && scale && dW1 && dgamma && dbeta, "mlp_l1_bwd: null pointer");
    mlp_l1_bwd_kernel<<<1, 64, 0, as_stream(stream)>>>(sums5x64, moments14, M, W1, mean, invstd, scale, dW1, dgamma, dbeta);
    KDF_LAUNCH_CHECK();
    return KDF_OK;
}

}  // extern "C"
