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
//   MMA      : tcgen05.mma 128 x NC x 16 per step into one of 2..4 fp32 accumulators in TMEM; the weights
//              [NC x K] stay resident in shared memory for all tiles of the CTA;
//   epilogue : tcgen05.ld -> (training) bf16 rows of the pre-BatchNorm output + the per-channel sum / sum of
//              squares of exactly the stored values (this layer's batch statistics: no statistics pass), or
//              (inference) the folded BatchNorm + activation (+ shortcut) applied before the store -- through a
//              swizzled staging slab so that every global store is a full 128-byte line.
//
// Warp-specialised: one TMA warp, one MMA warp, n_tr "transform" warpgroups (the prologue) and n_ep epilogue
// warpgroups meet only at mbarriers (full -> ready -> empty per ring stage, acc_full / acc_empty per accumulator);
// an epilogue warpgroup owns one 64-channel slab of one row tile at a time, so up to four slabs drain concurrently
// while the tensor cores fill the next accumulator and the ring loads the tiles after it.  The first, lock-step
// version of this kernel (all 8 warps through load-wait / transform / MMA / epilogue phases separated by CTA
// barriers) reached 0.12-0.41 of the HBM peak: every phase exposed its own latency.
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
constexpr int PW_PANEL = PW_ROWS * 128;         // bytes of one 64-channel panel of a row tile
constexpr int PW_MAX_STAGES = 8;
constexpr int PW_MAX_WG = 4;                    // transform + epilogue warpgroups
constexpr int PW_MAX_THREADS = PW_MAX_WG * 128 + 64;

struct PwArgs {
    int64_t M;
    int K, N, NC;                     // K % 64 == 0, NC in {32, 64, 128, 256}, N % NC == 0
    int stages;                       // ring depth (2..8)
    int n_ep, n_tr;                   // epilogue / transform warpgroups (n_tr > 0 iff PRO)
    int nb;                           // accumulators in TMEM (2..4)
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
    // offsets from the 1024-aligned base, all multiples of 1024: W panels | ring | one staging slab per epilogue WG | misc
    int off_ring, off_stage, off_misc, total;
    __host__ __device__ PwSmem(int K, int NC, int stages, int n_ep) {
        off_ring = NC * K * 2;
        off_stage = off_ring + stages * PW_PANEL;
        off_misc = off_stage + n_ep * PW_PANEL;
        // misc: 32 mbarriers (8 full, 8 ready, 8 empty, 4 acc_full, 4 acc_empty) + tmem slot in 512 B, then the
        // prologue scale/shift [K] and the epilogue scale/shift [NC]
        total = off_misc + 512 + 4 * (2 * K + 2 * NC) + 1024;
    }
};

// (a, b) = (a, b) * (s0, s1) + (h0, h1) as one packed fp32 FMA (same rounding as two scalar ones)
__device__ __forceinline__ void pw_fma2(float &a, float &b, float s0, float s1, float h0, float h1) {
    unsigned long long x, sc, sh, r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(x) : "f"(a), "f"(b));
    asm("mov.b64 %0, {%1, %2};" : "=l"(sc) : "f"(s0), "f"(s1));
    asm("mov.b64 %0, {%1, %2};" : "=l"(sh) : "f"(h0), "f"(h1));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(x), "l"(sc), "l"(sh));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(r));
}
__device__ __forceinline__ uint32_t pw_max2(uint32_t a, uint32_t b) {
    uint32_t r;
    asm("max.bf16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    return r;
}
__device__ __forceinline__ uint32_t pw_min2(uint32_t a, uint32_t b) {
    uint32_t r;
    asm("min.bf16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    return r;
}
__device__ __forceinline__ float pw_act(float v, int act) {
    if (act == 1) return fmaxf(v, 0.f);
    if (act == 2) return fminf(fmaxf(v, 0.f), 6.f);
    return v;
}
__device__ __forceinline__ void pw_mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc::smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void pw_wg_sync(int id) {                   // the 128 threads of one warpgroup
    asm volatile("bar.sync %0, 128;" ::"r"(id) : "memory");
}

template <bool PRO, int EPI>
__global__ void __launch_bounds__(PW_MAX_THREADS, 1)
pw_conv_fwd_kernel(PwArgs a, const __grid_constant__ CUtensorMap tmA) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = tc::align_smem_1024(smem_raw);
    const PwSmem L(a.K, a.NC, a.stages, a.n_ep);
    uint8_t *sW = smem;
    uint8_t *sRing = smem + L.off_ring;
    uint8_t *sStage = smem + L.off_stage;
    uint64_t *bar_full = reinterpret_cast<uint64_t *>(smem + L.off_misc);       // [8] TMA landed
    uint64_t *bar_ready = bar_full + PW_MAX_STAGES;                              // [8] PRO: transformed
    uint64_t *bar_empty = bar_ready + PW_MAX_STAGES;                             // [8] MMAs of the panel completed
    uint64_t *bar_accf = bar_empty + PW_MAX_STAGES;                              // [4] accumulator complete
    uint64_t *bar_acce = bar_accf + 4;                                           // [4] accumulator drained
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + L.off_misc + 256 + 8);
    float *t_psc = reinterpret_cast<float *>(smem + L.off_misc + 512), *t_psh = t_psc + a.K;
    float *t_esc = t_psh + a.K, *t_esh = t_esc + a.NC;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int K = a.K, N = a.N, NC = a.NC, S = a.stages, NB = a.nb;
    const int n_ep = a.n_ep, n_tr = a.n_tr;
    const int KP = K >> 6;                                             // 64-channel panels per row tile
    const int n0 = blockIdx.y * NC;                                    // first output channel of this CTA's chunk
    const int n_slabs = (NC + 63) >> 6;
    const int64_t n_tiles = (a.M + PW_ROWS - 1) / PW_ROWS;
    const int64_t my_tiles = ((int64_t)blockIdx.x < n_tiles) ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const int64_t total_panels = my_tiles * KP;
    uint32_t tmem_cols = 32;
    while ((int)tmem_cols < NB * NC) tmem_cols <<= 1;

    // ---- one-time setup (all threads)
    if (tid == 0) {
        for (int s = 0; s < PW_MAX_STAGES; ++s) {
            tc::mbar_init(&bar_full[s], 1);
            tc::mbar_init(&bar_ready[s], 128);
            tc::mbar_init(&bar_empty[s], 1);
        }
        for (int b = 0; b < 4; ++b) { tc::mbar_init(&bar_accf[b], 1); tc::mbar_init(&bar_acce[b], (uint32_t)n_slabs); }
        tc::mbar_fence_init();
        tc::fence_async_smem();
        // the ring starts filling before anything else (the weights are copied while the first panels are on their way)
        int pn = 0, row = (int)blockIdx.x * PW_ROWS;
        for (int g = 0; g < S && g < total_panels; ++g) {
            tma::mbar_expect_tx(&bar_full[g], PW_PANEL);
            tma::load_2d(sRing + g * PW_PANEL, &tmA, pn * 64, row, &bar_full[g]);
            if (++pn == KP) { pn = 0; row += (int)gridDim.x * PW_ROWS; }
        }
    }
    const int nthreads = blockDim.x;
    if (PRO) for (int i = tid; i < K; i += nthreads) { t_psc[i] = a.pro_scale[i]; t_psh[i] = a.pro_shift[i]; }
    if (EPI == 1) for (int i = tid; i < NC; i += nthreads) { t_esc[i] = a.epi_scale[n0 + i]; t_esh[i] = a.epi_shift[n0 + i]; }
    {   // weights of this chunk -> swizzled panels [KP][NC rows][128 B]
        const int cpr = K >> 3;                                        // 16-byte chunks per weight row
        for (int idx = tid; idx < NC * cpr; idx += nthreads) {
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

    const int wg = warp >> 2;
    if (wg < n_ep) {
        // ================================================================== epilogue warpgroup `wg`
        // unit u = (local tile, slab); this warpgroup drains units wg, wg + n_ep, ...  n_slabs divides n_ep (host-checked),
        // so its slab index is fixed and one set of statistics accumulators covers it.
        const int t = tid & 127;
        const int erow = (warp & 3) * 32 + lane;                       // accumulator row = TMEM lane
        const uint32_t lane_bits = (uint32_t)((warp & 3) * 32) << 16;
        const int pch = t & 7, prow0 = t >> 3;                         // store phase: chunk pch of rows prow0 + 16 p
        uint8_t *stg = sStage + wg * PW_PANEL;
        const bool want_stats = (EPI == 0) && a.stats != nullptr;
        float s_sum[8], s_sq[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) { s_sum[j] = 0.f; s_sq[j] = 0.f; }
        // n_slabs divides n_ep: this warpgroup drains slab `s` of local tiles tl0, tl0 + tstep, ... (indices kept incrementally:
        // a 64-bit division per unit cost more issue slots than the 16 KB the unit moves)
        const int s = wg % n_slabs, tl0 = wg / n_slabs, tstep = n_ep / n_slabs;
        const int ntl = (int)my_tiles, Mi = (int)a.M;
        const int col = s * 64 + pch * 8;                              // this thread's 8 output channels inside the chunk
        const bool two = s * 64 + 32 < NC, col_ok = col < NC;
        const uint32_t st_w = tc::smem_u32(stg) + (uint32_t)erow * 128u;   // staging row of this thread's accumulator row
        const uint32_t st_r = tc::smem_u32(stg) + (uint32_t)prow0 * 128u + (uint32_t)((pch ^ (prow0 & 7)) << 4);
        int buf = tl0 % NB;
        uint32_t phase = (uint32_t)((tl0 / NB) & 1);
        int row0 = ((int)blockIdx.x + tl0 * (int)gridDim.x) * PW_ROWS;
        const int row_step = tstep * (int)gridDim.x * PW_ROWS;
        for (int tl = tl0; tl < ntl; tl += tstep, row0 += row_step) {
            tc::mbar_wait(&bar_accf[buf], phase);
            tc::fence_after_sync();
            // both 32-column halves of the slab in flight before the one wait (two serialised TMEM round trips otherwise)
            uint32_t r[2][32];
            const uint32_t tcol = tmem_base + lane_bits + (uint32_t)(buf * NC + s * 64);
            tc::tmem_ld32(tcol, r[0]);
            if (two) tc::tmem_ld32(tcol + 32, r[1]);
            tc::tmem_ld_wait();
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                if (h == 0 || two) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        float v[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e) v[e] = __uint_as_float(r[h][8 * j + e]);
                        if (EPI == 1) {                                // folded BatchNorm as four packed FMAs (two channels each)
                            const int c0 = s * 64 + h * 32 + 8 * j;
                            const float4 sa = *reinterpret_cast<const float4 *>(t_esc + c0), sb = *reinterpret_cast<const float4 *>(t_esc + c0 + 4);
                            const float4 ha = *reinterpret_cast<const float4 *>(t_esh + c0), hb = *reinterpret_cast<const float4 *>(t_esh + c0 + 4);
                            pw_fma2(v[0], v[1], sa.x, sa.y, ha.x, ha.y);
                            pw_fma2(v[2], v[3], sa.z, sa.w, ha.z, ha.w);
                            pw_fma2(v[4], v[5], sb.x, sb.y, hb.x, hb.y);
                            pw_fma2(v[6], v[7], sb.z, sb.w, hb.z, hb.w);
                        }
                        const uint32_t chunk = (uint32_t)(h * 4 + j);  // 16-byte chunk of the 128-byte slab row
                        uint32_t p0 = pack_bf16(v[0], v[1]), p1 = pack_bf16(v[2], v[3]), p2 = pack_bf16(v[4], v[5]), p3 = pack_bf16(v[6], v[7]);
                        if (EPI == 1) {
                            // the activation on the packed bf16 pairs: 0 and 6 are bf16 numbers and the rounding is monotonic, so
                            // clamp(round(x)) == round(clamp(x)) -- one instruction per two channels instead of two per channel
                            if (a.epi_act >= 1) { p0 = pw_max2(p0, 0u); p1 = pw_max2(p1, 0u); p2 = pw_max2(p2, 0u); p3 = pw_max2(p3, 0u); }
                            if (a.epi_act == 2) { p0 = pw_min2(p0, 0x40C040C0u); p1 = pw_min2(p1, 0x40C040C0u); p2 = pw_min2(p2, 0x40C040C0u); p3 = pw_min2(p3, 0x40C040C0u); }
                        }
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(st_w + ((chunk ^ (uint32_t)(erow & 7)) << 4)),
                                     "r"(p0), "r"(p1), "r"(p2), "r"(p3) : "memory");
                    }
                }
            }
            tc::fence_before_sync();
            pw_wg_sync(1 + wg);                                        // slab staged; every TMEM read of this unit has completed
            if (t == 0) pw_mbar_arrive(&bar_acce[buf]);                // (n_slabs arrivals release the accumulator)
            if (col_ok) {
                const int rows_left = Mi - row0;                       // rows of this tile that exist
                __nv_bfloat16 *obase = a.out + (int64_t)row0 * N + n0 + col;
                const __nv_bfloat16 *rbase = (EPI == 1 && a.residual) ? a.residual + (int64_t)row0 * N + n0 + col : nullptr;
#pragma unroll
                for (int p = 0; p < 8; ++p) {
                    const int rr = prow0 + 16 * p;
                    if (rr < rows_left) {
                        uint4 v;
                        // rows prow0 + 16 p share (rr & 7) with prow0: one staging address, stepped by 16 rows
                        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(st_r + (uint32_t)(p * 16 * 128)));
                        const int off = rr * N;
                        if (EPI == 1 && rbase) {
                            const uint4 q = *reinterpret_cast<const uint4 *>(rbase + off);
                            v.x = pack_bf16(bf16_lo(v.x) + bf16_lo(q.x), bf16_hi(v.x) + bf16_hi(q.x));
                            v.y = pack_bf16(bf16_lo(v.y) + bf16_lo(q.y), bf16_hi(v.y) + bf16_hi(q.y));
                            v.z = pack_bf16(bf16_lo(v.z) + bf16_lo(q.z), bf16_hi(v.z) + bf16_hi(q.z));
                            v.w = pack_bf16(bf16_lo(v.w) + bf16_lo(q.w), bf16_hi(v.w) + bf16_hi(q.w));
                        }
                        *reinterpret_cast<uint4 *>(obase + off) = v;
                        if (want_stats) {
                            const float f[8] = {bf16_lo(v.x), bf16_hi(v.x), bf16_lo(v.y), bf16_hi(v.y),
                                                bf16_lo(v.z), bf16_hi(v.z), bf16_lo(v.w), bf16_hi(v.w)};
#pragma unroll
                            for (int j = 0; j < 8; ++j) { s_sum[j] += f[j]; s_sq[j] = fmaf(f[j], f[j], s_sq[j]); }
                        }
                    }
                }
            }
            pw_wg_sync(1 + wg);                                        // the slab is the next unit's staging buffer
            buf += tstep;
            while (buf >= NB) { buf -= NB; phase ^= 1u; }
        }
        const int64_t n_units = (tl0 < ntl) ? 1 : 0;                   // did this warpgroup drain anything
        if (want_stats && n_units) {
            float *red = reinterpret_cast<float *>(stg);               // [16 row lanes][8 chunks][16]
#pragma unroll
            for (int j = 0; j < 8; ++j) { red[(prow0 * 8 + pch) * 16 + j] = s_sum[j]; red[(prow0 * 8 + pch) * 16 + 8 + j] = s_sq[j]; }
            pw_wg_sync(1 + wg);
            const int ch = t >> 4, j = t & 15;                         // chunk, (sum | sq, element)
            const int col = s * 64 + ch * 8 + (j & 7);
            if (col < NC) {
                float v = 0.f;
                for (int r = 0; r < 16; ++r) v += red[(r * 8 + ch) * 16 + j];
                atomicAdd(a.stats + (int64_t)(j >> 3) * N + n0 + col, (double)v);
            }
        }
    } else if (wg < n_ep + n_tr) {
        // ================================================================== transform warpgroup (PRO only)
        if (PRO) {
            const int j_tr = wg - n_ep;
            const int t = tid & 127;
            const int pch = t & 7, prow0 = t >> 3;
            int slot = j_tr % S, pn = j_tr % KP;
            uint32_t ph = (uint32_t)((j_tr / S) & 1);
            for (int64_t g = j_tr; g < total_panels; g += n_tr) {
                tc::mbar_wait(&bar_full[slot], ph);
                uint8_t *pan = sRing + slot * PW_PANEL;
                const float *sc = t_psc + pn * 64 + pch * 8, *sh = t_psh + pn * 64 + pch * 8;
                const float4 sa = *reinterpret_cast<const float4 *>(sc), sb = *reinterpret_cast<const float4 *>(sc + 4);
                const float4 ha = *reinterpret_cast<const float4 *>(sh), hb = *reinterpret_cast<const float4 *>(sh + 4);
                const float cs[8] = {sa.x, sa.y, sa.z, sa.w, sb.x, sb.y, sb.z, sb.w};
                const float ch_[8] = {ha.x, ha.y, ha.z, ha.w, hb.x, hb.y, hb.z, hb.w};
#pragma unroll
                for (int p = 0; p < 8; ++p) {
                    uint4 *cell = reinterpret_cast<uint4 *>(pan + tc::sw128_offset(prow0 + 16 * p, pch));
                    const uint4 u = *cell;
                    float v[8] = {bf16_lo(u.x), bf16_hi(u.x), bf16_lo(u.y), bf16_hi(u.y), bf16_lo(u.z), bf16_hi(u.z), bf16_lo(u.w), bf16_hi(u.w)};
#pragma unroll
                    for (int e = 0; e < 8; ++e) v[e] = pw_act(fmaf(v[e], cs[e], ch_[e]), a.pro_act);
                    *cell = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
                }
                tc::fence_async_smem();                                // generic-proxy writes -> visible to the tensor cores
                pw_mbar_arrive(&bar_ready[slot]);
                slot += n_tr;
                while (slot >= S) { slot -= S; ph ^= 1u; }
                pn += n_tr;
                while (pn >= KP) pn -= KP;
            }
        }
    } else if (warp == 4 * (n_ep + n_tr)) {
        // ================================================================== TMA warp
        if (lane == 0) {
            // (pulling panels DRAM -> L2 ahead of the ring with cp.async.bulk.prefetch.tensor was measured: 0.45 -> 0.50 ms over
            // the 16 layer shapes -- the prefetches compete with the loads they are meant to help)
            // panels 0 .. S-1 were issued during the set-up; every further load re-arms a slot the MMAs have released
            int slot = 0, pn = S % KP, row = ((int)blockIdx.x + (S / KP) * (int)gridDim.x) * PW_ROWS;
            uint32_t ph = 0;                                           // parity of the PREVIOUS use of the slot
            for (int64_t g = S; g < total_panels; ++g) {
                tc::mbar_wait(&bar_empty[slot], ph);
                tma::mbar_expect_tx(&bar_full[slot], PW_PANEL);
                tma::load_2d(sRing + slot * PW_PANEL, &tmA, pn * 64, row, &bar_full[slot]);
                if (++slot == S) { slot = 0; ph ^= 1u; }
                if (++pn == KP) { pn = 0; row += (int)gridDim.x * PW_ROWS; }
            }
        }
    } else if (warp == 4 * (n_ep + n_tr) + 1) {
        // ================================================================== MMA warp
        if (lane == 0) {
            const uint32_t idesc = tc::make_idesc(PW_ROWS, NC, 0, 0);
            int slot = 0, buf = 0;
            uint32_t ph = 0, aph = 1;                                  // ring parity; parity of the previous use of the accumulator
            for (int64_t tl = 0; tl < my_tiles; ++tl) {
                if (tl >= NB) tc::mbar_wait(&bar_acce[buf], aph);      // drained by the epilogue
                tc::fence_after_sync();
                const uint32_t d = tmem_base + (uint32_t)(buf * NC);
                for (int pn = 0; pn < KP; ++pn) {
                    tc::mbar_wait(PRO ? &bar_ready[slot] : &bar_full[slot], ph);
                    tc::fence_after_sync();
                    const uint32_t a_base = tc::smem_u32(sRing + slot * PW_PANEL);
                    const uint32_t w_base = tc::smem_u32(sW + pn * (NC * tc::ROW_BYTES));
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        tc::mma_bf16(d, tc::desc_kmajor(a_base + (uint32_t)k * 32u), tc::desc_kmajor(w_base + (uint32_t)k * 32u), idesc, pn > 0 || k > 0);
                    tc::mma_commit(&bar_empty[slot]);                  // the tensor cores are done with this panel
                    if (++slot == S) { slot = 0; ph ^= 1u; }
                }
                tc::mma_commit(&bar_accf[buf]);
                if (++buf == NB) { buf = 0; aph ^= 1u; }
            }
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem_base, tmem_cols);
}

// NC: the widest chunk of output channels (a divisor of N out of {256, 128, 64, 32}, at most `cap`) whose weights fit
static int pw_pick_chunk(int K, int N, int cap) {
    const int cands[] = {256, 128, 64, 32};
    for (int c : cands)
        if (c <= cap && c <= N && N % c == 0 && (int64_t)c * K * 2 <= 96 * 1024) return c;
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
    KDF_CHECK_ARG(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(W) | reinterpret_cast<uintptr_t>(out) |
                    reinterpret_cast<uintptr_t>(residual)) & 15) == 0, "pw_conv_fwd: x, W, out and residual must be 16-byte aligned");
    // with a prologue at least one warpgroup transforms, so at most two drain slabs: chunks of <= 128 channels then
    int NC = pw_pick_chunk(K, N, pro_scale ? 128 : 256);
    KDF_CHECK_ARG(NC > 0, "pw_conv_fwd: no chunking of N=%d fits (K=%d)", N, K);
    // few row tiles (the 32x32 maps): narrower chunks give the SMs more, smaller work items -- 256 tiles on 148 SMs is two
    // waves with the second 27 % full; the extra reads of the tile by the other chunks' CTAs are L2 hits
    // (not with a prologue: every chunk's CTA would transform all K panels again)
    while (!pro_scale && NC > 64 && ((M + PW_ROWS - 1) / PW_ROWS) * (N / NC) < 3 * (int64_t)sm_count()) NC /= 2;
    const int KP = K / 64, n_slabs = (NC + 63) / 64;
    // warpgroups: without a prologue all four drain accumulators; with one they are split by the work per row tile
    // (KP panels to transform against n_slabs slabs to drain, a slab costing about 1.25 panels).  n_ep is a multiple of
    // n_slabs, so that every epilogue warpgroup keeps one slab index (one set of statistics accumulators).
    int n_tr = 0, n_ep = PW_MAX_WG;
    if (pro_scale) {
        n_tr = (int)(PW_MAX_WG * (KP * 0.8) / (KP * 0.8 + n_slabs) + 0.5);
        if (n_tr < 1) n_tr = 1;
        if (n_tr > PW_MAX_WG - n_slabs) n_tr = PW_MAX_WG - n_slabs;
        n_ep = (PW_MAX_WG - n_tr) / n_slabs * n_slabs;
        n_tr = PW_MAX_WG - n_ep;
    }
    KDF_CHECK_ARG(n_ep >= 1 && n_ep % n_slabs == 0 && (pro_scale == nullptr || n_tr >= 1), "pw_conv_fwd: internal: warpgroup split");
    int nb = 512 / NC;
    if (nb > 4) nb = 4;
    // ring depth from what is left of 227 KB after the weights, the staging slabs and the tables
    const int fixed = PwSmem(K, NC, 0, n_ep).total;
    int stages = (227 * 1024 - fixed) / PW_PANEL;
    if (stages > PW_MAX_STAGES) stages = PW_MAX_STAGES;
    KDF_CHECK_ARG(stages >= 2, "pw_conv_fwd: K=%d N=%d leaves no room for the panel ring", K, N);
    PwArgs a;
    a.M = M; a.K = K; a.N = N; a.NC = NC; a.stages = stages; a.n_ep = n_ep; a.n_tr = n_tr; a.nb = nb;
    a.W = static_cast<const __nv_bfloat16 *>(W);
    a.pro_scale = pro_scale; a.pro_shift = pro_shift; a.pro_act = pro_act;
    a.epi_scale = epi_scale; a.epi_shift = epi_shift; a.epi_act = epi_act;
    a.residual = static_cast<const __nv_bfloat16 *>(residual);
    a.out = static_cast<__nv_bfloat16 *>(out);
    a.stats = stats;
    CUtensorMap tm;
    KDF_CHECK_ARG(tma::make_row_map(&tm, x, M, K), "pw_conv_fwd: cuTensorMapEncodeTiled failed (x must be 16-byte aligned)");
    const int smem = PwSmem(K, NC, stages, n_ep).total;
    const int n_chunks = N / NC;
    const int64_t n_tiles = (M + PW_ROWS - 1) / PW_ROWS;
    int64_t gx = sm_count() / n_chunks;
    if (gx < 1) gx = 1;
    if (gx > n_tiles) gx = n_tiles;
    const dim3 grid((unsigned)gx, (unsigned)n_chunks);
    const int threads = (n_ep + n_tr) * 128 + 64;
#define KDF_PW_LAUNCH(PRO, EPI)                                                                                        \
    do {                                                                                                               \
        KDF_CUDA(cudaFuncSetAttribute(pw_conv_fwd_kernel<PRO, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); \
        pw_conv_fwd_kernel<PRO, EPI><<<grid, threads, smem, st>>>(a, tm);                                              \
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
