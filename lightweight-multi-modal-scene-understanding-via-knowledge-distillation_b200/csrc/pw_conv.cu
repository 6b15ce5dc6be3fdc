// Pointwise (1x1) convolutions of the camera branch, the FPN, the fusion projections and the head as ONE
// tensor-core kernel per layer (tcgen05 + TMEM + TMA), sm_100a.
//
// The reference runs every 1x1 convolution as conv2d -> BatchNorm2d -> ReLU(6) (camera_encoder.py:19-51,
// fusion_module.py:8-34, 162-173): a GEMM over pixel rows followed by a statistics pass, a normalise pass and an
// activation pass over its output, and the same again in front of the next convolution.  With channels-last maps a
// 1x1 convolution is  Z[M, N] = A[M, K] . W[N, K]^T  over M = B*H*W pixel rows, and all of those passes fold into
// the GEMM kernel:
//
//   load     : the raw bf16 rows of the layer's INPUT arrive by TMA (cp.async.bulk.tensor.2d, SWIZZLE_128B boxes of
//              128 rows x 64 channels) straight into the operand layout of the tensor cores, through a ring of
//              16 KB stages that runs ahead of the math by up to eight panels -- no registers, no instructions;
//   prologue : (optional) the PREVIOUS layer's BatchNorm-apply + ReLU/ReLU6, applied in place to the landed panel
//              -- the normalised activation of the producer never exists in HBM;
//   MMA      : tcgen05.mma 128 x NC x 16 per step into a double-buffered fp32 accumulator in TMEM (one elected
//              thread); the weights [NC x K] stay resident in shared memory for all tiles of the CTA;
//   epilogue : tcgen05.ld -> (training) bf16 rows of the pre-BatchNorm output + the per-channel sum / sum of
//              squares of exactly the stored values (this layer's batch statistics: no statistics pass), or
//              (inference) the folded BatchNorm + activation (+ shortcut) applied before the store -- through a
//              swizzled staging slab so that every global store is a full 128-byte line.
//
// HBM traffic per layer = rows in + rows out, once.  Wide outputs (N = 384, 768) are split into chunks of NC <= 256
// accumulator columns over blockIdx.y; the chunks of one row tile run at the same time on different SMs, so the
// second read of the tile is an L2 hit.
#include <stdlib.h>

#include "kdf_common.cuh"
#include "tc_common.cuh"
#include "tma_common.cuh"

namespace kdf {

constexpr int PW_ROWS = 128;                    // pixel rows per tile = MMA M
constexpr int PW_THREADS = 256;
constexpr int PW_PANEL = PW_ROWS * 128;         // bytes of one 64-channel panel of a row tile
constexpr int PW_MAX_STAGES = 8;
constexpr int PW_MAX_SLABS = 4;                 // NC <= 256 = 4 slabs of 64 output channels

struct PwArgs {
    int64_t M;
    int K, N, NC;                     // K % 64 == 0, NC % 32 == 0, NC <= 256, N % NC == 0
    int stages;                       // ring depth (2..8)
    const __nv_bfloat16 *W;           // [N, K] row-major
    const float *pro_scale, *pro_shift;   // PRO: [K]
    int pro_act;                      // 0 none, 1 relu, 2 relu6
    const float *epi_scale, *epi_shift;   // EPI 1: [N]
    int epi_act;
    const __nv_bfloat16 *residual;    // EPI 1, nullable: [M, N]
    __nv_bfloat16 *out;               // [M, N]
    double *stats;                    // EPI 0, nullable: [2][N] (accumulated)
};

struct PwSmem {
    // offsets from the 1024-aligned base, all multiples of 1024: W panels | ring | 2 staging slabs | misc
    int off_ring, off_stage, off_misc, total;
    __host__ __device__ PwSmem(int K, int NC, int stages) {
        off_ring = NC * K * 2;
        off_stage = off_ring + stages * PW_PANEL;
        off_misc = off_stage + 2 * PW_PANEL;
        // misc: 18 mbarriers (8 full, 8 empty, 2 accumulator) + tmem slot in 256 B, then pro scale/shift [K] and epi scale/shift [NC]
        total = off_misc + 256 + 4 * (2 * K + 2 * NC) + 1024;
    }
};

__device__ __forceinline__ float pw_act(float v, int act) {
    if (act == 1) return fmaxf(v, 0.f);
    if (act == 2) return fminf(fmaxf(v, 0.f), 6.f);
    return v;
}

template <bool PRO, int EPI>
__global__ void __launch_bounds__(PW_THREADS, 1)
pw_conv_fwd_kernel(PwArgs a, const __grid_constant__ CUtensorMap tmA) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = tc::align_smem_1024(smem_raw);
    const PwSmem L(a.K, a.NC, a.stages);
    uint8_t *sW = smem;
    uint8_t *sRing = smem + L.off_ring;
    uint8_t *sStage = smem + L.off_stage;
    uint64_t *bar_full = reinterpret_cast<uint64_t *>(smem + L.off_misc);       // [8]
    uint64_t *bar_empty = bar_full + PW_MAX_STAGES;                              // [8]
    uint64_t *bar_acc = bar_empty + PW_MAX_STAGES;                               // [2]
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + L.off_misc + 160);
    float *t_psc = reinterpret_cast<float *>(smem + L.off_misc + 256), *t_psh = t_psc + a.K;
    float *t_esc = t_psh + a.K, *t_esh = t_esc + a.NC;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int K = a.K, N = a.N, NC = a.NC, S = a.stages;
    const int KP = K >> 6;                                             // 64-channel panels per row tile
    const int n0 = blockIdx.y * NC;                                    // first output channel of this CTA's chunk
    const int n_slabs = (NC + 63) >> 6;
    const int64_t n_tiles = (a.M + PW_ROWS - 1) / PW_ROWS;
    const int64_t my_tiles = ((int64_t)blockIdx.x < n_tiles) ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const int64_t total_panels = my_tiles * KP;
    const uint32_t idesc = tc::make_idesc(PW_ROWS, NC, 0, 0);
    uint32_t tmem_cols = 32;
    while ((int)tmem_cols < 2 * NC) tmem_cols <<= 1;

    auto issue_tma = [&](int64_t g) {                                  // one thread: panel g of this CTA's sequence
        const int slot = (int)(g % S);
        const int64_t tile = (int64_t)blockIdx.x + (g / KP) * gridDim.x;
        const int pn = (int)(g % KP);
        tma::mbar_expect_tx(&bar_full[slot], PW_PANEL);
        tma::load_2d(sRing + slot * PW_PANEL, &tmA, pn * 64, (int)(tile * PW_ROWS), &bar_full[slot]);
    };

    // ---- one-time setup
    if (tid == 0) {
        for (int s = 0; s < PW_MAX_STAGES; ++s) { tc::mbar_init(&bar_full[s], 1); tc::mbar_init(&bar_empty[s], 1); }
        tc::mbar_init(&bar_acc[0], 1);
        tc::mbar_init(&bar_acc[1], 1);
        tc::mbar_fence_init();
        tc::fence_async_smem();
        for (int64_t g = 0; g < S && g < total_panels; ++g) issue_tma(g);   // the ring starts filling before anything else
    }
    if (PRO) for (int i = tid; i < K; i += PW_THREADS) { t_psc[i] = a.pro_scale[i]; t_psh[i] = a.pro_shift[i]; }
    if (EPI == 1) for (int i = tid; i < NC; i += PW_THREADS) { t_esc[i] = a.epi_scale[n0 + i]; t_esh[i] = a.epi_shift[n0 + i]; }
    {   // weights of this chunk -> swizzled panels [KP][NC rows][128 B]
        const int cpr = K >> 3;                                        // 16-byte chunks per weight row
        for (int idx = tid; idx < NC * cpr; idx += PW_THREADS) {
            const int n = idx / cpr, ch = idx - n * cpr;
            const uint4 w = *reinterpret_cast<const uint4 *>(a.W + (int64_t)(n0 + n) * K + ch * 8);
            *reinterpret_cast<uint4 *>(sW + (ch >> 3) * (NC * tc::ROW_BYTES) + tc::sw128_offset(n, ch & 7)) = w;
        }
    }
    __syncwarp();
    if (warp == 0) tc::tmem_alloc(tmem_slot, tmem_cols);
    tc::fence_async_smem();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;

    // fixed per-thread roles: prologue / store phases own 16-byte chunk `pch` of rows prow0 + 32 p
    const int pch = tid & 7, prow0 = tid >> 3;
    // TMEM epilogue: warp w reads lanes 32 (w & 3) .. of the 32 accumulator columns 64 slab + 32 (w >> 2) ..
    const int erow = (warp & 3) * 32 + lane, ehalf = warp >> 2;
    const uint32_t lane_bits = (uint32_t)((warp & 3) * 32) << 16;

    float s_sum[PW_MAX_SLABS][8], s_sq[PW_MAX_SLABS][8];
    if (EPI == 0) {
#pragma unroll
        for (int s = 0; s < PW_MAX_SLABS; ++s)
#pragma unroll
            for (int j = 0; j < 8; ++j) { s_sum[s][j] = 0.f; s_sq[s][j] = 0.f; }
    }
    const bool want_stats = (EPI == 0) && a.stats != nullptr;

    auto mma_panel = [&](int slot, int pn, int buf) {                  // one thread
        const uint32_t a_base = tc::smem_u32(sRing + slot * PW_PANEL);
        const uint32_t w_base = tc::smem_u32(sW + pn * (NC * tc::ROW_BYTES));
        const uint32_t d = tmem_base + (uint32_t)(buf * NC);
#pragma unroll
        for (int k = 0; k < 4; ++k)
            tc::mma_bf16(d, tc::desc_kmajor(a_base + (uint32_t)k * 32u), tc::desc_kmajor(w_base + (uint32_t)k * 32u), idesc, pn > 0 || k > 0);
        tc::mma_commit(&bar_empty[slot]);                              // the tensor cores are done with this panel
    };
    auto refill = [&](int64_t g_done) {                                // one thread: re-arm the slot of a panel whose MMAs were issued earlier
        if (g_done < 0 || g_done + S >= total_panels) return;
        tc::mbar_wait(&bar_empty[g_done % S], (uint32_t)((g_done / S) & 1));
        issue_tma(g_done + S);
    };

    int slab_parity = 0;                                               // staging buffers alternate across slabs AND tiles
    auto epilogue = [&](int64_t tile, int buf, uint32_t parity) {
        const int64_t r0 = tile * PW_ROWS;
        tc::mbar_wait(&bar_acc[buf], parity);
        tc::fence_after_sync();
#pragma unroll
        for (int s = 0; s < PW_MAX_SLABS; ++s) {
            if (s < n_slabs) {
                uint8_t *stg = sStage + slab_parity * PW_PANEL;
                const int c0 = s * 64 + ehalf * 32;                    // first accumulator column of this warp in the chunk
                if (c0 < NC) {
                    uint32_t r[32];
                    tc::tmem_ld32(tmem_base + lane_bits + (uint32_t)(buf * NC + c0), r);
                    tc::tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        float v[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e) v[e] = __uint_as_float(r[8 * j + e]);
                        if (EPI == 1) {
                            const float4 sa = *reinterpret_cast<const float4 *>(t_esc + c0 + 8 * j), sb = *reinterpret_cast<const float4 *>(t_esc + c0 + 8 * j + 4);
                            const float4 ha = *reinterpret_cast<const float4 *>(t_esh + c0 + 8 * j), hb = *reinterpret_cast<const float4 *>(t_esh + c0 + 8 * j + 4);
                            const float cs[8] = {sa.x, sa.y, sa.z, sa.w, sb.x, sb.y, sb.z, sb.w};
                            const float ch_[8] = {ha.x, ha.y, ha.z, ha.w, hb.x, hb.y, hb.z, hb.w};
#pragma unroll
                            for (int e = 0; e < 8; ++e) v[e] = pw_act(fmaf(v[e], cs[e], ch_[e]), a.epi_act);
                        }
                        const int chunk = ehalf * 4 + j;               // 16-byte chunk of the 128-byte slab row
                        *reinterpret_cast<uint4 *>(stg + erow * 128 + ((chunk ^ (erow & 7)) << 4)) =
                            make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
                    }
                }
                tc::fence_before_sync();
                __syncthreads();
                const int col = s * 64 + pch * 8;                      // this thread's 8 output channels inside the chunk
                if (col < NC) {
#pragma unroll
                    for (int p = 0; p < 4; ++p) {
                        const int rr = prow0 + 32 * p;
                        if (r0 + rr < a.M) {
                            uint4 v = *reinterpret_cast<const uint4 *>(stg + rr * 128 + ((pch ^ (rr & 7)) << 4));
                            const int64_t off = (r0 + rr) * N + n0 + col;
                            if (EPI == 1 && a.residual) {
                                const uint4 q = *reinterpret_cast<const uint4 *>(a.residual + off);
                                v.x = pack_bf16(bf16_lo(v.x) + bf16_lo(q.x), bf16_hi(v.x) + bf16_hi(q.x));
                                v.y = pack_bf16(bf16_lo(v.y) + bf16_lo(q.y), bf16_hi(v.y) + bf16_hi(q.y));
                                v.z = pack_bf16(bf16_lo(v.z) + bf16_lo(q.z), bf16_hi(v.z) + bf16_hi(q.z));
                                v.w = pack_bf16(bf16_lo(v.w) + bf16_lo(q.w), bf16_hi(v.w) + bf16_hi(q.w));
                            }
                            *reinterpret_cast<uint4 *>(a.out + off) = v;
                            if (want_stats) {
                                const float f[8] = {bf16_lo(v.x), bf16_hi(v.x), bf16_lo(v.y), bf16_hi(v.y),
                                                    bf16_lo(v.z), bf16_hi(v.z), bf16_lo(v.w), bf16_hi(v.w)};
#pragma unroll
                                for (int j = 0; j < 8; ++j) { s_sum[s][j] += f[j]; s_sq[s][j] = fmaf(f[j], f[j], s_sq[s][j]); }
                            }
                        }
                    }
                }
                slab_parity ^= 1;
            }
        }
    };

    // ---- main loop over this CTA's row tiles
    int64_t g = 0;                                                     // running panel index
    int64_t prev_tile = -1;
    int it = 0;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
        const int buf = it & 1;
        for (int pn = 0; pn < KP; ++pn, ++g) {
            const int slot = (int)(g % S);
            const uint32_t ph = (uint32_t)((g / S) & 1);
            if (PRO) {
                tc::mbar_wait(&bar_full[slot], ph);                    // every thread: the panel has landed
                uint8_t *pan = sRing + slot * PW_PANEL;
                const float *sc = t_psc + pn * 64 + pch * 8, *sh = t_psh + pn * 64 + pch * 8;
                const float4 sa = *reinterpret_cast<const float4 *>(sc), sb = *reinterpret_cast<const float4 *>(sc + 4);
                const float4 ha = *reinterpret_cast<const float4 *>(sh), hb = *reinterpret_cast<const float4 *>(sh + 4);
                const float cs[8] = {sa.x, sa.y, sa.z, sa.w, sb.x, sb.y, sb.z, sb.w};
                const float ch_[8] = {ha.x, ha.y, ha.z, ha.w, hb.x, hb.y, hb.z, hb.w};
#pragma unroll
                for (int p = 0; p < 4; ++p) {
                    uint4 *cell = reinterpret_cast<uint4 *>(pan + tc::sw128_offset(prow0 + 32 * p, pch));
                    const uint4 u = *cell;
                    float v[8] = {bf16_lo(u.x), bf16_hi(u.x), bf16_lo(u.y), bf16_hi(u.y), bf16_lo(u.z), bf16_hi(u.z), bf16_lo(u.w), bf16_hi(u.w)};
#pragma unroll
                    for (int e = 0; e < 8; ++e) v[e] = pw_act(fmaf(v[e], cs[e], ch_[e]), a.pro_act);
                    *cell = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
                }
                tc::fence_async_smem();                                // generic-proxy writes -> visible to the tensor cores
                tc::fence_before_sync();
                __syncthreads();
                if (tid == 0) {
                    tc::fence_after_sync();
                    mma_panel(slot, pn, buf);
                    if (pn == KP - 1) tc::mma_commit(&bar_acc[buf]);
                    refill(g - 1);
                }
            } else if (tid == 0) {
                tc::mbar_wait(&bar_full[slot], ph);
                tc::fence_after_sync();
                mma_panel(slot, pn, buf);
                if (pn == KP - 1) tc::mma_commit(&bar_acc[buf]);
                refill(g - 1);
            }
        }
        __syncwarp();                                                  // warp 0 reconverges after the single-thread issue loop
        if (it > 0) epilogue(prev_tile, buf ^ 1, (uint32_t)(((it - 1) >> 1) & 1));
        prev_tile = tile;
    }
    if (it > 0) epilogue(prev_tile, (it - 1) & 1, (uint32_t)(((it - 1) >> 1) & 1));
    __syncthreads();

    // ---- batch statistics: reduce the 32 threads that share a channel group, then 2 * NC fp64 atomics per CTA
    if (want_stats) {
        float *red = reinterpret_cast<float *>(sStage);                // [32 row lanes][8 chunks][16]
#pragma unroll
        for (int s = 0; s < PW_MAX_SLABS; ++s) {
            if (s < n_slabs) {
#pragma unroll
                for (int j = 0; j < 8; ++j) { red[(prow0 * 8 + pch) * 16 + j] = s_sum[s][j]; red[(prow0 * 8 + pch) * 16 + 8 + j] = s_sq[s][j]; }
                __syncthreads();
                if (tid < 128) {
                    const int ch = tid >> 4, j = tid & 15;             // chunk, (sum | sq, element)
                    const int col = s * 64 + ch * 8 + (j & 7);
                    if (col < NC) {
                        float v = 0.f;
                        for (int r = 0; r < 32; ++r) v += red[(r * 8 + ch) * 16 + j];
                        atomicAdd(a.stats + (int64_t)(j >> 3) * N + n0 + col, (double)v);
                    }
                }
                __syncthreads();
            }
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem_base, tmem_cols);
}

// NC: the widest chunk of output channels (<= 256 accumulator columns, a divisor of N) whose weights fit the budget
static int pw_pick_chunk(int K, int N) {
    const int cands[] = {256, 192, 128, 96, 64, 32};
    for (int c : cands)
        if (c <= N && N % c == 0 && (int64_t)c * K * 2 <= 96 * 1024) return c;
    return 0;
}

}  // namespace kdf

extern "C" int kdf_pw_conv_fwd(const void *x, int64_t M, int K, int N, const void *W,
                               const float *pro_scale, const float *pro_shift, int pro_act,
                               const float *epi_scale, const float *epi_shift, int epi_act, const void *residual,
                               void *out, double *stats, void *stream) {
    using namespace kdf;
    KDF_CHECK_ARG(M >= 0 && M < (1ll << 31), "pw_conv_fwd: M=%lld out of range", (long long)M);
    KDF_CHECK_ARG(K >= 64 && K % 64 == 0 && K <= 1024, "pw_conv_fwd: K=%d must be a multiple of 64 in [64, 1024]", K);
    KDF_CHECK_ARG(N >= 32 && N % 32 == 0, "pw_conv_fwd: N=%d must be a multiple of 32", N);
    KDF_CHECK_ARG((pro_scale == nullptr) == (pro_shift == nullptr) && (epi_scale == nullptr) == (epi_shift == nullptr),
                  "pw_conv_fwd: scale / shift come in pairs");
    KDF_CHECK_ARG(!(epi_scale && stats), "pw_conv_fwd: statistics are taken of the raw output (no epilogue affine)");
    KDF_CHECK_ARG(!(residual && !epi_scale), "pw_conv_fwd: the shortcut rides in the affine epilogue");
    KDF_CHECK_ARG(pro_act >= 0 && pro_act <= 2 && epi_act >= 0 && epi_act <= 2, "pw_conv_fwd: bad activation code");
    cudaStream_t st = as_stream(stream);
    if (stats) KDF_CUDA(cudaMemsetAsync(stats, 0, sizeof(double) * 2 * N, st));
    if (M == 0) return KDF_OK;
    KDF_CHECK_ARG(x && W && out, "pw_conv_fwd: null pointer");
    const int NC = pw_pick_chunk(K, N);
    KDF_CHECK_ARG(NC > 0, "pw_conv_fwd: no chunking of N=%d fits (K=%d)", N, K);
    // ring depth from what is left of 227 KB after the weights, the staging slabs and the tables
    const int fixed = PwSmem(K, NC, 0).total;
    int stages = (227 * 1024 - fixed) / PW_PANEL;
    if (stages > PW_MAX_STAGES) stages = PW_MAX_STAGES;
    KDF_CHECK_ARG(stages >= 2, "pw_conv_fwd: K=%d N=%d leaves no room for the panel ring", K, N);
    PwArgs a;
    a.M = M; a.K = K; a.N = N; a.NC = NC; a.stages = stages;
    a.W = static_cast<const __nv_bfloat16 *>(W);
    a.pro_scale = pro_scale; a.pro_shift = pro_shift; a.pro_act = pro_act;
    a.epi_scale = epi_scale; a.epi_shift = epi_shift; a.epi_act = epi_act;
    a.residual = static_cast<const __nv_bfloat16 *>(residual);
    a.out = static_cast<__nv_bfloat16 *>(out);
    a.stats = stats;
    CUtensorMap tm;
    KDF_CHECK_ARG(tma::make_row_map(&tm, x, M, K), "pw_conv_fwd: cuTensorMapEncodeTiled failed (x must be 16-byte aligned)");
    const int smem = PwSmem(K, NC, stages).total;
    const int n_chunks = N / NC;
    const int64_t n_tiles = (M + PW_ROWS - 1) / PW_ROWS;
    int64_t gx = sm_count() / n_chunks;
    if (gx < 1) gx = 1;
    if (gx > n_tiles) gx = n_tiles;
    const dim3 grid((unsigned)gx, (unsigned)n_chunks);
#define KDF_PW_LAUNCH(PRO, EPI)                                                                                        \
    do {                                                                                                               \
        KDF_CUDA(cudaFuncSetAttribute(pw_conv_fwd_kernel<PRO, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); \
        pw_conv_fwd_kernel<PRO, EPI><<<grid, PW_THREADS, smem, st>>>(a, tm);                                           \
    } while (0)
    if (pro_scale) {
        if (epi_scale) KDF_PW_LAUNCH(true, 1); else KDF_PW_LAUNCH(true, 0);
    } else {
        if (epi_scale) KDF_PW_LAUNCH(false, 1); else KDF_PW_LAUNCH(false, 0);
    }
#undef KDF_PW_LAUNCH
    KDF_LAUNCH_CHECK();
    return KDF_OK;
}
