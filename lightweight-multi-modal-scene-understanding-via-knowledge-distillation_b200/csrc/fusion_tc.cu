// Weighted camera-LiDAR fusion on the 5th-gen tensor cores (tcgen05 + TMEM), bf16 rows, C = 128, sm_100a.
//
// Same contract as fusion.cu's fp32-FMA kernels (reference src/models/fusion_module.py:107-136, executed
// inline at :248-253): BatchNorm-apply + ReLU of both 1x1 projection blocks, concat, the 256->128->2
// attention MLP, the 2-way softmax and the blend -- one kernel per direction.  The block moves 3*C*2 bytes
// per pixel forward and 5*C*2 backward against 0.13 / 0.4 MFLOP per pixel: HBM-bound once the contractions
// run on the tensor cores, which is what this file does (the FMA version is compute-bound at 2-3 % of the
// HBM roofline).
//
// Forward, per 128-pixel tile:
//   global -> registers -> BN-apply + ReLU -> bf16 -> Y tile [128 px x 256] (4 swizzled 64-column panels)
//   tcgen05.mma   hidden[128 x 128] (TMEM, fp32) = Y . W1^T            (W1 resident in smem as bf16)
//   epilogue A    (thread = pixel) relu(hidden + b1) . w2^T + b2 -> softmax -> attention weights
//   epilogue B    (coalesced) out = Ycam * w0 + Ylid * w1 -> bf16 stores
// Backward, per tile: the same Y tile and hidden recompute, then
//   epilogue 1    softmax / blend backward per pixel, d hidden (bf16) -> smem; column sums for db1, dW2 by
//                 warp transpose-reductions (no shared-memory pass)
//   tcgen05.mma   dY[128 x 256]  = dH . W1          (A K-major, B = the MN-major view of the resident W1 tile)
//   tcgen05.mma   dW1[128 x 256] += dH^T . Y        (both MN-major views; accumulates in TMEM across tiles)
//   epilogue 2    ReLU mask + blend path, BN-affine backward, bf16 gradient rows, d scale / d shift sums.
#include <stdlib.h>

#include "kdf_common.cuh"
#include "tc_common.cuh"
#include "tma_common.cuh"

namespace kdf {

constexpr int FT_THREADS = 256;
constexpr int FT_ROWS = 128;
constexpr int FT_C = 128;
constexpr int FT_K2 = 256;
constexpr uint32_t FT_PANEL = FT_ROWS * tc::ROW_BYTES;            // 16384 bytes between 64-column panels

struct FusionTcArgs {
    const __nv_bfloat16 *cam, *lid;
    int64_t M;
    const float *csc, *csh, *lsc, *lsh, *w1, *b1, *w2, *b2;
    __nv_bfloat16 *out;
    float *attn;                                                  // forward: written; backward: read
    const __nv_bfloat16 *gout;
    __nv_bfloat16 *gcam, *glid;
    float *gaff, *gw1, *gb1, *gw2, *gb2;
};

struct FtSmem {
    static constexpr int OFF_W1 = 0;                              // [128 j][256 k] bf16, 4 panels
    static constexpr int OFF_Y = OFF_W1 + 4 * FT_PANEL;           // [128 px][256 k] bf16, 4 panels
    static constexpr int OFF_MISC_FWD = OFF_Y + 4 * FT_PANEL;
    static constexpr int OFF_G = OFF_Y + 4 * FT_PANEL;            // backward only: upstream gradient tile, 2 panels
    static constexpr int OFF_DH = OFF_G + 2 * FT_PANEL;           //                d hidden tile, 2 panels
    static constexpr int OFF_MISC_BWD = OFF_DH + 2 * FT_PANEL;
    // misc: 2 mbarriers, tmem slot, then float tables b1[128] w2[256] aff[512] rowS[512] dS[256]
    static constexpr int MISC_BYTES = 64 + 4 * (128 + 256 + 512 + 512 + 256);
    static constexpr int TOTAL_FWD = OFF_MISC_FWD + MISC_BYTES + 1024;
    static constexpr int TOTAL_BWD = OFF_MISC_BWD + MISC_BYTES + 1024;
};

__device__ __forceinline__ void ft_unpack8(const uint4 &u, float (&f)[8]) {
    f[0] = bf16_lo(u.x); f[1] = bf16_hi(u.x); f[2] = bf16_lo(u.y); f[3] = bf16_hi(u.y);
    f[4] = bf16_lo(u.z); f[5] = bf16_hi(u.z); f[6] = bf16_lo(u.w); f[7] = bf16_hi(u.w);
}
__device__ __forceinline__ uint4 ft_pack8(const float (&v)[8]) {
    return make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
}
// relu(x*scale + shift) of one 16-byte chunk, coefficients from shared memory
__device__ __forceinline__ uint4 ft_affine_relu(const uint4 &raw, const float *sc, const float *sh) {
    float v[8];
    ft_unpack8(raw, v);
    // the 8 coefficients of a chunk are 32-byte aligned in the table: four 16-byte loads, not sixteen 4-byte ones
    const float4 s0 = *reinterpret_cast<const float4 *>(sc), s1 = *reinterpret_cast<const float4 *>(sc + 4);
    const float4 h0 = *reinterpret_cast<const float4 *>(sh), h1 = *reinterpret_cast<const float4 *>(sh + 4);
    v[0] = fmaxf(fmaf(v[0], s0.x, h0.x), 0.f); v[1] = fmaxf(fmaf(v[1], s0.y, h0.y), 0.f);
    v[2] = fmaxf(fmaf(v[2], s0.z, h0.z), 0.f); v[3] = fmaxf(fmaf(v[3], s0.w, h0.w), 0.f);
    v[4] = fmaxf(fmaf(v[4], s1.x, h1.x), 0.f); v[5] = fmaxf(fmaf(v[5], s1.y, h1.y), 0.f);
    v[6] = fmaxf(fmaf(v[6], s1.z, h1.z), 0.f); v[7] = fmaxf(fmaf(v[7], s1.w, h1.w), 0.f);
    return ft_pack8(v);
}

// Sum over the 32 lanes of a warp of column `lane` of a [32 lanes x 32 values] register block
// (recursive halving: 31 shuffles instead of 32 full reductions).
__device__ __forceinline__ float warp_transpose_sum(float (&v)[32], int lane) {
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
        const bool hi = (lane & o) != 0;
#pragma unroll
        for (int i = 0; i < o; ++i) {
            const float send = hi ? v[i] : v[i + o];
            const float keep = hi ? v[i + o] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
        }
    }
    return v[0];
}

// One-time setup shared by both kernels: coefficient tables, W1 -> bf16 swizzled panels, barriers, TMEM.
__device__ __forceinline__ void ft_setup(const FusionTcArgs &a, uint8_t *sW1, float *tb1, float *tw2, float *taff,
                                         uint64_t *bars, uint32_t *tmem_slot, uint32_t tmem_cols) {
    const int tid = threadIdx.x, nthr = blockDim.x;
    for (int i = tid; i < 128; i += nthr) {
        tb1[i] = a.b1[i];
        taff[i] = a.csc[i]; taff[128 + i] = a.csh[i]; taff[256 + i] = a.lsc[i]; taff[384 + i] = a.lsh[i];
    }
    for (int i = tid; i < 256; i += nthr) tw2[i] = a.w2[i];
    for (int idx = tid; idx < FT_C * (FT_K2 / 8); idx += nthr) {
        const int j = idx >> 5, ch = idx & 31;
        const float4 lo = __ldg(reinterpret_cast<const float4 *>(a.w1 + (int64_t)j * FT_K2 + ch * 8));
        const float4 hi = __ldg(reinterpret_cast<const float4 *>(a.w1 + (int64_t)j * FT_K2 + ch * 8 + 4));
        const float v[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
        *reinterpret_cast<uint4 *>(sW1 + (ch >> 3) * FT_PANEL + tc::sw128_offset(j, ch & 7)) = ft_pack8(v);
    }
    if (tid == 0) {
        tc::mbar_init(&bars[0], 1);
        tc::mbar_init(&bars[1], 1);
        tc::mbar_fence_init();
    }
    if ((tid >> 5) == 0) tc::tmem_alloc(tmem_slot, tmem_cols);
    tc::fence_async_smem();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
}

// hidden[128 x 128] = Y[128 x 256] . W1^T : both operands K-major, 16 k-steps
__device__ __forceinline__ void ft_mma_hidden(uint32_t y_base, uint32_t w_base, uint32_t acc, uint64_t *bar) {
    constexpr uint32_t IDESC = tc::make_idesc(FT_ROWS, FT_C, 0, 0);
#pragma unroll
    for (int k = 0; k < FT_K2 / 16; ++k) {
        const uint32_t off = (uint32_t)(k >> 2) * FT_PANEL + (uint32_t)(k & 3) * 32u;
        tc::mma_bf16(acc, tc::desc_kmajor(y_base + off), tc::desc_kmajor(w_base + off), IDESC, k > 0);
    }
    tc::mma_commit(bar);
}

// ============================================================================= forward, TMA-staged tiles
// The pre-BatchNorm rows of a tile arrive by TMA
// (cp.async.bulk.tensor.2d, SWIZZLE_128B tensor maps over the [M,128] row tensors, 4 boxes of 64 columns x 128
// rows) DIRECTLY in the panel layout the tensor cores read, are BatchNorm-applied + ReLU'd in place, and two
// operand tiles are kept in flight: the loads of tile i+1 land while tile i is transformed, multiplied and
// blended -- no registers hold data in flight, so the CUDA-core phases run 16 warps wide.
constexpr int FTM_THREADS = 512;

struct FtTmaSmem {
    static constexpr int OFF_W1 = 0;
    static constexpr int OFF_Y0 = OFF_W1 + 4 * FT_PANEL;
    static constexpr int OFF_Y1 = OFF_Y0 + 4 * FT_PANEL;
    static constexpr int OFF_MISC = OFF_Y1 + 4 * FT_PANEL;
    // 3 mbarriers + tmem slot, then float tables b1[128] w2[256] aff[512] rowS[256] part[1024]
    static constexpr int MISC_BYTES = 64 + 4 * (128 + 256 + 512 + 256 + 1024);
    static constexpr int TOTAL = OFF_MISC + MISC_BYTES + 1024;
};

__global__ void __launch_bounds__(FTM_THREADS, 1)
fusion_weighted_fwd_tma_kernel(FusionTcArgs a, const __grid_constant__ CUtensorMap tm_cam, const __grid_constant__ CUtensorMap tm_lid) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = tc::align_smem_1024(smem_raw);
    uint8_t *sW1 = smem + FtTmaSmem::OFF_W1;
    auto sY = [smem](int b) -> uint8_t * { return smem + (b ? FtTmaSmem::OFF_Y1 : FtTmaSmem::OFF_Y0); };   // one shared base: LDS/STS
    uint8_t *misc = smem + FtTmaSmem::OFF_MISC;
    uint64_t *bar_raw = reinterpret_cast<uint64_t *>(misc);           // [2]
    uint64_t *bar_mma = bar_raw + 2;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(misc + 32);
    float *tb1 = reinterpret_cast<float *>(misc + 64), *tw2 = tb1 + 128, *taff = tw2 + 256, *rowS = taff + 512, *part = rowS + 256;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t n_tiles = (a.M + FT_ROWS - 1) / FT_ROWS;
    auto issue_tile = [&](int64_t tile, int buf) {                     // one thread: 4 boxes of 64 columns x 128 rows
        tma::mbar_expect_tx(&bar_raw[buf], 4 * FT_PANEL);
        const int r0 = (int)(tile * FT_ROWS);
        tma::load_2d(sY(buf) + 0 * FT_PANEL, &tm_cam, 0, r0, &bar_raw[buf]);
        tma::load_2d(sY(buf) + 1 * FT_PANEL, &tm_cam, 64, r0, &bar_raw[buf]);
        tma::load_2d(sY(buf) + 2 * FT_PANEL, &tm_lid, 0, r0, &bar_raw[buf]);
        tma::load_2d(sY(buf) + 3 * FT_PANEL, &tm_lid, 64, r0, &bar_raw[buf]);
    };
    // barriers first, so that the first two tiles are already in flight during the rest of the setup
    if (tid == 0) {
        tc::mbar_init(&bar_raw[0], 1);
        tc::mbar_init(&bar_raw[1], 1);
        tc::mbar_init(bar_mma, 1);
        tc::mbar_fence_init();
        tc::fence_async_smem();
        if ((int64_t)blockIdx.x < n_tiles) issue_tile(blockIdx.x, 0);
        if ((int64_t)blockIdx.x + gridDim.x < n_tiles) issue_tile((int64_t)blockIdx.x + gridDim.x, 1);
    }
    for (int i = tid; i < 128; i += FTM_THREADS) {
        tb1[i] = a.b1[i];
        taff[i] = a.csc[i]; taff[128 + i] = a.csh[i]; taff[256 + i] = a.lsc[i]; taff[384 + i] = a.lsh[i];
    }
    for (int i = tid; i < 256; i += FTM_THREADS) tw2[i] = a.w2[i];
    for (int idx = tid; idx < FT_C * (FT_K2 / 8); idx += FTM_THREADS) {
        const int j = idx >> 5, ch = idx & 31;
        const float4 lo = __ldg(reinterpret_cast<const float4 *>(a.w1 + (int64_t)j * FT_K2 + ch * 8));
        const float4 hi = __ldg(reinterpret_cast<const float4 *>(a.w1 + (int64_t)j * FT_K2 + ch * 8 + 4));
        const float v[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
        *reinterpret_cast<uint4 *>(sW1 + (ch >> 3) * FT_PANEL + tc::sw128_offset(j, ch & 7)) = ft_pack8(v);
    }
    if (warp == 0) tc::tmem_alloc(tmem_slot, 128);
    tc::fence_async_smem();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    const float bb0 = __ldg(a.b2), bb1 = __ldg(a.b2 + 1);
    // transform mapping: 32 chunks per pixel row (16 camera + 16 LiDAR), 16 rows per pass
    const int tch = tid & 31, trow0 = tid >> 5;
    const int tpanel = tch >> 3, tc8 = tch & 7;                        // panel 0,1 camera; 2,3 LiDAR
    const float *t_sc = taff + (tpanel < 2 ? 0 : 256) + (tpanel & 1) * 64 + tc8 * 8;
    const float *t_sh = t_sc + 128;
    const int och = tid & 15, orow0 = tid >> 4;                        // blend mapping: 16 chunks per output row, 32 rows per pass

    int it = 0;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
        const int buf = it & 1;
        const int64_t r0 = tile * FT_ROWS;
        uint8_t *y = sY(buf);
        tc::mbar_wait(&bar_raw[buf], (uint32_t)((it >> 1) & 1));
        // ---- BatchNorm-apply + ReLU in place (rows past M are forced to zero: TMA zero-fills them, relu(shift) would not be zero)
#pragma unroll
        for (int p = 0; p < 8; ++p) {
            const int r = trow0 + p * 16;
            uint4 *ptr4 = reinterpret_cast<uint4 *>(y + tpanel * FT_PANEL + tc::sw128_offset(r, tc8));
            *ptr4 = (r0 + r < a.M) ? ft_affine_relu(*ptr4, t_sc, t_sh) : make_uint4(0u, 0u, 0u, 0u);
        }
        tc::fence_async_smem();
        __syncthreads();
        if (tid == 0) {
            tc::fence_after_sync();
            ft_mma_hidden(tc::smem_u32(y), tc::smem_u32(sW1), tmem_base, bar_mma);
        }
        // ---- epilogue A: attention logits per pixel (four threads share a pixel: 32 hidden units each)
        tc::mbar_wait(bar_mma, (uint32_t)(it & 1));
        tc::fence_after_sync();
        const int row = (warp & 3) * 32 + lane, cgp = warp >> 2;
        {
            float p0 = 0.f, p1 = 0.f;
            uint32_t r[32];
            const int col0 = cgp * 32;
            tc::tmem_ld32(tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)col0, r);
            tc::tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
                const float4 b4 = *reinterpret_cast<const float4 *>(tb1 + col0 + j);
                const float4 u4 = *reinterpret_cast<const float4 *>(tw2 + col0 + j), v4 = *reinterpret_cast<const float4 *>(tw2 + 128 + col0 + j);
                const float h0 = fmaxf(__uint_as_float(r[j]) + b4.x, 0.f), h1 = fmaxf(__uint_as_float(r[j + 1]) + b4.y, 0.f);
                const float h2 = fmaxf(__uint_as_float(r[j + 2]) + b4.z, 0.f), h3 = fmaxf(__uint_as_float(r[j + 3]) + b4.w, 0.f);
                p0 = fmaf(h0, u4.x, p0); p0 = fmaf(h1, u4.y, p0); p0 = fmaf(h2, u4.z, p0); p0 = fmaf(h3, u4.w, p0);
                p1 = fmaf(h0, v4.x, p1); p1 = fmaf(h1, v4.y, p1); p1 = fmaf(h2, v4.z, p1); p1 = fmaf(h3, v4.w, p1);
            }
            part[(row * 4 + cgp) * 2] = p0;
            part[(row * 4 + cgp) * 2 + 1] = p1;
        }
        tc::fence_before_sync();
        __syncthreads();
        if (tid < FT_ROWS) {
            const float4 pa = *reinterpret_cast<const float4 *>(part + tid * 8), pb = *reinterpret_cast<const float4 *>(part + tid * 8 + 4);
            const float s0 = ((pa.x + pa.z) + (pb.x + pb.z)) + bb0, s1 = ((pa.y + pa.w) + (pb.y + pb.w)) + bb1;
            const float mx = fmaxf(s0, s1);
            const float e0 = expf(s0 - mx), e1 = expf(s1 - mx);
            const float inv = 1.f / (e0 + e1);
            rowS[2 * tid] = e0 * inv;
            rowS[2 * tid + 1] = e1 * inv;
            if (r0 + tid < a.M) *reinterpret_cast<float2 *>(a.attn + 2 * (r0 + tid)) = make_float2(e0 * inv, e1 * inv);
        }
        __syncthreads();
        // ---- epilogue B: blend, coalesced 16-byte stores
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            const int r = orow0 + p * 32;
            if (r0 + r < a.M) {
                const uint32_t off = (och >> 3) * FT_PANEL + tc::sw128_offset(r, och & 7);
                float yc[8], yl[8], o[8];
                ft_unpack8(*reinterpret_cast<const uint4 *>(y + off), yc);
                ft_unpack8(*reinterpret_cast<const uint4 *>(y + 2 * FT_PANEL + off), yl);
                const float w0 = rowS[2 * r], w1 = rowS[2 * r + 1];
#pragma unroll
                for (int j = 0; j < 8; ++j) o[j] = yc[j] * w0 + yl[j] * w1;
                *reinterpret_cast<uint4 *>(a.out + (r0 + r) * FT_C + och * 8) = ft_pack8(o);
            }
        }
        __syncthreads();                                                  // every read of this operand tile is done
        if (tid == 0 && tile + 2 * (int64_t)gridDim.x < n_tiles) {
            tc::fence_async_smem();                                        // generic-proxy accesses before the async-proxy refill
            issue_tile(tile + 2 * (int64_t)gridDim.x, buf);
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem_base, 128);
}

// ============================================================================= backward
// NT = 256 or 512 threads: the CUDA-core phases (staging, the two epilogues, the coalesced store phase) are latency-bound,
// 16 warps hide more of it; per-thread tile state halves with NT = 512 (4 row passes, 32 accumulator columns per phase).
template <int NT>
__global__ void __launch_bounds__(NT, 1)
fusion_weighted_bwd_tc_kernel(FusionTcArgs a) {
    constexpr int RG = NT / 16;                   // row groups of the 16-chunks-per-row phases
    constexpr int RP = FT_ROWS / RG;              // row passes per tile
    constexpr int CG = NT / 128;                  // accumulator column groups (warp >> 2)
    constexpr int E1_COLS = FT_C / CG;            // hidden columns per warp in epilogue 1 (64 or 32)
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = tc::align_smem_1024(smem_raw);
    uint8_t *sW1 = smem + FtSmem::OFF_W1, *sY = smem + FtSmem::OFF_Y, *sG = smem + FtSmem::OFF_G, *sDH = smem + FtSmem::OFF_DH;
    uint8_t *misc = smem + FtSmem::OFF_MISC_BWD;
    uint64_t *bars = reinterpret_cast<uint64_t *>(misc);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(misc + 16);
    float *tb1 = reinterpret_cast<float *>(misc + 64), *tw2 = tb1 + 128, *taff = tw2 + 256, *rowS = taff + 512, *dS = rowS + 512;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int och = tid & 15, orow0 = tid >> 4;
    const int64_t n_tiles = (a.M + FT_ROWS - 1) / FT_ROWS;
    ft_setup(a, sW1, tb1, tw2, taff, bars, tmem_slot, 512);
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t lane_bits = (uint32_t)((warp & 3) * 32) << 16;
    constexpr uint32_t IDESC_DG = tc::make_idesc(FT_ROWS, FT_K2, 0, 1);    // dY = dH . W1   (B = MN-major view of W1)
    constexpr uint32_t IDESC_WG = tc::make_idesc(FT_C, FT_K2, 1, 1);       // dW1 += dH^T . Y

    // accumulators that live for the whole kernel
    float g_sc[16], g_sh[16];                    // d scale / d shift of (cam | lid) channels och*8 .. +8
#pragma unroll
    for (int j = 0; j < 16; ++j) { g_sc[j] = 0.f; g_sh[j] = 0.f; }
    float g_b1[E1_COLS / 32], g_w2a[E1_COLS / 32], g_w2b[E1_COLS / 32];       // column = (warp>>2)*E1_COLS + half*32 + lane
#pragma unroll
    for (int h = 0; h < E1_COLS / 32; ++h) { g_b1[h] = 0.f; g_w2a[h] = 0.f; g_w2b[h] = 0.f; }
    float g_b2a = 0.f, g_b2b = 0.f;

    int it = 0;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
        const int64_t r0 = tile * FT_ROWS;
        // ---- loads + staging: Y (4 panels), G (2 panels); per-pixel dots d0 = G.Ycam, d1 = G.Ylid
        if (tid == 64) {                                                   // the next tile on its way into L2
            const int64_t nt_ = tile + gridDim.x;
            if (nt_ < n_tiles) {
                const int64_t nr0 = nt_ * FT_ROWS;
                const uint32_t bytes = (uint32_t)(((a.M - nr0 < FT_ROWS) ? (a.M - nr0) : FT_ROWS) * FT_C * 2);
                tc::prefetch_l2(a.cam + nr0 * FT_C, bytes);
                tc::prefetch_l2(a.lid + nr0 * FT_C, bytes);
                tc::prefetch_l2(a.gout + nr0 * FT_C, bytes);
            }
        }
        uint4 raw_c[RP], raw_l[RP];
        {
            uint4 raw_g[RP];
#pragma unroll
            for (int p = 0; p < RP; ++p) {
                const int64_t row = r0 + orow0 + p * RG;
                if (row < a.M) {
                    raw_c[p] = ldg_stream_u4(reinterpret_cast<const uint4 *>(a.cam + row * FT_C + och * 8));
                    raw_l[p] = ldg_stream_u4(reinterpret_cast<const uint4 *>(a.lid + row * FT_C + och * 8));
                    raw_g[p] = ldg_stream_u4(reinterpret_cast<const uint4 *>(a.gout + row * FT_C + och * 8));
                }
            }
#pragma unroll
            for (int p = 0; p < RP; ++p) {
                const int r = orow0 + p * RG;
                uint4 yc = make_uint4(0u, 0u, 0u, 0u), yl = yc, gv = yc;
                if (r0 + r < a.M) {
                    yc = ft_affine_relu(raw_c[p], taff + och * 8, taff + 128 + och * 8);
                    yl = ft_affine_relu(raw_l[p], taff + 256 + och * 8, taff + 384 + och * 8);
                    gv = raw_g[p];
                }
                const uint32_t off = (och >> 3) * FT_PANEL + tc::sw128_offset(r, och & 7);
                *reinterpret_cast<uint4 *>(sY + off) = yc;
                *reinterpret_cast<uint4 *>(sY + 2 * FT_PANEL + off) = yl;
                *reinterpret_cast<uint4 *>(sG + off) = gv;
                float fc[8], fl[8], fg[8];
                ft_unpack8(yc, fc); ft_unpack8(yl, fl); ft_unpack8(gv, fg);
                float d0 = 0.f, d1 = 0.f;
#pragma unroll
                for (int j = 0; j < 8; ++j) { d0 = fmaf(fg[j], fc[j], d0); d1 = fmaf(fg[j], fl[j], d1); }
#pragma unroll
                for (int o = 8; o >= 1; o >>= 1) {                        // the 16 lanes that share the pixel
                    d0 += __shfl_xor_sync(0xffffffffu, d0, o);
                    d1 += __shfl_xor_sync(0xffffffffu, d1, o);
                }
                if (och == 0) { dS[2 * r] = d0; dS[2 * r + 1] = d1; }
            }
        }
        tc::fence_async_smem();
        __syncthreads();
        if (tid == 0) {
            tc::fence_after_sync();
            ft_mma_hidden(tc::smem_u32(sY), tc::smem_u32(sW1), tmem_base, &bars[0]);
        }
        // ---- epilogue 1: softmax / blend backward per pixel, d hidden -> smem, column sums
        const int row = (warp & 3) * 32 + lane;
        float w0 = 0.f, w1 = 0.f;
        if (r0 + row < a.M) {
            const float2 at = __ldg(reinterpret_cast<const float2 *>(a.attn + 2 * (r0 + row)));
            w0 = at.x; w1 = at.y;
        }
        const float d0 = dS[2 * row], d1 = dS[2 * row + 1];
        const float dot = w0 * d0 + w1 * d1;
        const float da0 = w0 * (d0 - dot), da1 = w1 * (d1 - dot);         // softmax backward
        if (warp < 4) {
            rowS[4 * row] = w0; rowS[4 * row + 1] = w1;
            g_b2a += da0; g_b2b += da1;
        }
        tc::mbar_wait(&bars[0], (uint32_t)(it & 1));
        tc::fence_after_sync();
#pragma unroll
        for (int half = 0; half < E1_COLS / 32; ++half) {
            const int col0 = (warp >> 2) * E1_COLS + half * 32;
            uint32_t r[32];
            tc::tmem_ld32(tmem_base + lane_bits + (uint32_t)col0, r);
            tc::tmem_ld_wait();
            float t[32];
            // d hidden (bf16 operand tile) and its column sum
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const float hp = __uint_as_float(r[j]) + tb1[col0 + j];
                t[j] = hp > 0.f ? fmaf(da0, tw2[col0 + j], da1 * tw2[128 + col0 + j]) : 0.f;
            }
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
                const int chunk = (col0 >> 3) + jj;                        // 16-byte chunk of the 128-wide d hidden row
                const float v[8] = {t[8 * jj], t[8 * jj + 1], t[8 * jj + 2], t[8 * jj + 3], t[8 * jj + 4], t[8 * jj + 5], t[8 * jj + 6], t[8 * jj + 7]};
                *reinterpret_cast<uint4 *>(sDH + (chunk >> 3) * FT_PANEL + tc::sw128_offset(row, chunk & 7)) = ft_pack8(v);
            }
            g_b1[half] += warp_transpose_sum(t, lane);
#pragma unroll
            for (int j = 0; j < 32; ++j) t[j] = da0 * fmaxf(__uint_as_float(r[j]) + tb1[col0 + j], 0.f);
            g_w2a[half] += warp_transpose_sum(t, lane);
#pragma unroll
            for (int j = 0; j < 32; ++j) t[j] = da1 * fmaxf(__uint_as_float(r[j]) + tb1[col0 + j], 0.f);
            g_w2b[half] += warp_transpose_sum(t, lane);
        }
        tc::fence_async_smem();
        tc::fence_before_sync();
        __syncthreads();
        // ---- dgrad + wgrad on the tensor cores
        if (tid == 0) {
            tc::fence_after_sync();
            const uint32_t dh_base = tc::smem_u32(sDH), w_base = tc::smem_u32(sW1), y_base = tc::smem_u32(sY);
#pragma unroll
            for (int k = 0; k < FT_C / 16; ++k) {                          // K = the 128 hidden units
                const uint32_t koff = (uint32_t)(k >> 2) * FT_PANEL + (uint32_t)(k & 3) * 32u;
                tc::mma_bf16(tmem_base, tc::desc_kmajor(dh_base + koff), tc::desc_mnmajor(w_base + (uint32_t)k * 2048u, FT_PANEL),
                             IDESC_DG, k > 0);
            }
#pragma unroll
            for (int k = 0; k < FT_ROWS / 16; ++k) {                       // K = the 128 pixels of the tile
                tc::mma_bf16(tmem_base + 256u, tc::desc_mnmajor(dh_base + (uint32_t)k * 2048u, FT_PANEL),
                             tc::desc_mnmajor(y_base + (uint32_t)k * 2048u, FT_PANEL), IDESC_WG, !(it == 0 && k == 0));
            }
            tc::mma_commit(&bars[1]);
        }
        tc::mbar_wait(&bars[1], (uint32_t)(it & 1));
        tc::fence_after_sync();
        // ---- epilogue 2 (thread = pixel; the first half of the warps the camera half, the rest the LiDAR half): dY + blend path, ReLU mask
        {
            constexpr int E2_WG = CG / 2;                                   // warp groups per half (1 or 2)
            const int half = (warp >> 2) / E2_WG, cq = (warp >> 2) % E2_WG;
            const float wsel = half ? rowS[4 * row + 1] : rowS[4 * row];
#pragma unroll
            for (int cc = 0; cc < 4 / E2_WG; ++cc) {
                const int c4 = cq * (4 / E2_WG) + cc;
                uint32_t r[32];
                tc::tmem_ld32(tmem_base + lane_bits + (uint32_t)(half * 128 + c4 * 32), r);
                tc::tmem_ld_wait();
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    const int chunk = c4 * 4 + jj;
                    const uint32_t off = (chunk >> 3) * FT_PANEL + tc::sw128_offset(row, chunk & 7);
                    uint8_t *yp = sY + (uint32_t)half * 2 * FT_PANEL + off;
                    float y[8], g[8], o[8];
                    ft_unpack8(*reinterpret_cast<const uint4 *>(yp), y);
                    ft_unpack8(*reinterpret_cast<const uint4 *>(sG + off), g);
#pragma unroll
                    for (int e = 0; e < 8; ++e) o[e] = y[e] > 0.f ? fmaf(g[e], wsel, __uint_as_float(r[8 * jj + e])) : 0.f;
                    *reinterpret_cast<uint4 *>(yp) = ft_pack8(o);          // in place: this thread owns the chunk
                }
            }
        }
        tc::fence_before_sync();
        __syncthreads();
        // ---- coalesced phase: BN-affine backward, gradient rows, d scale / d shift
#pragma unroll
        for (int p = 0; p < RP; ++p) {
            const int r = orow0 + p * RG;
            if (r0 + r < a.M) {
                const uint32_t off = (och >> 3) * FT_PANEL + tc::sw128_offset(r, och & 7);
                float gy[8], x[8], o[8];
                ft_unpack8(*reinterpret_cast<const uint4 *>(sY + off), gy);
                ft_unpack8(raw_c[p], x);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    g_sc[j] = fmaf(gy[j], x[j], g_sc[j]);
                    g_sh[j] += gy[j];
                    o[j] = gy[j] * taff[och * 8 + j];
                }
                *reinterpret_cast<uint4 *>(a.gcam + (r0 + r) * FT_C + och * 8) = ft_pack8(o);
                ft_unpack8(*reinterpret_cast<const uint4 *>(sY + 2 * FT_PANEL + off), gy);
                ft_unpack8(raw_l[p], x);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    g_sc[8 + j] = fmaf(gy[j], x[j], g_sc[8 + j]);
                    g_sh[8 + j] += gy[j];
                    o[j] = gy[j] * taff[256 + och * 8 + j];
                }
                *reinterpret_cast<uint4 *>(a.glid + (r0 + r) * FT_C + och * 8) = ft_pack8(o);
            }
        }
        __syncthreads();
    }

    // ---- flush: per-thread sums -> shared-memory reduction over the 16 row groups -> atomics
    float *red = reinterpret_cast<float *>(sY);                            // [RG row groups][16 chunks][32]
#pragma unroll
    for (int j = 0; j < 16; ++j) { red[(orow0 * 16 + och) * 32 + j] = g_sc[j]; red[(orow0 * 16 + och) * 32 + 16 + j] = g_sh[j]; }
    __syncthreads();
    for (int i = tid; i < 16 * 32; i += NT) {
        const int ch = i >> 5, j = i & 31;
        float v = 0.f;
        for (int g = 0; g < RG; ++g) v += red[(g * 16 + ch) * 32 + j];
        // j: 0-7 d cam scale, 8-15 d lid scale, 16-23 d cam shift, 24-31 d lid shift of channel ch*8 + (j & 7)
        const int which = (j < 8) ? 0 : (j < 16) ? 2 : (j < 24) ? 1 : 3;
        atomicAdd(a.gaff + which * FT_C + ch * 8 + (j & 7), v);
    }
#pragma unroll
    for (int half = 0; half < E1_COLS / 32; ++half) {
        const int col = (warp >> 2) * E1_COLS + half * 32 + lane;
        atomicAdd(a.gb1 + col, g_b1[half]);
        atomicAdd(a.gw2 + col, g_w2a[half]);
        atomicAdd(a.gw2 + FT_C + col, g_w2b[half]);
    }
    if (warp < 4) {
        const float s0 = warp_sum(g_b2a), s1 = warp_sum(g_b2b);
        if (lane == 0) { atomicAdd(a.gb2, s0); atomicAdd(a.gb2 + 1, s1); }
    }
    if (it > 0) {                                                          // dW1: TMEM columns 256..511, row = hidden unit
        tc::fence_after_sync();
        const int j = (warp & 3) * 32 + lane;
#pragma unroll
        for (int c4 = 0; c4 < 8 / CG; ++c4) {
            const int col = (warp >> 2) * (FT_K2 / CG) + c4 * 32;
            uint32_t r[32];
            tc::tmem_ld32(tmem_base + lane_bits + 256u + (uint32_t)col, r);
            tc::tmem_ld_wait();
            float *dst = a.gw1 + (int64_t)j * FT_K2 + col;                 // 32 consecutive floats of row j: 8 vector reductions
#pragma unroll
            for (int e = 0; e < 32; e += 4) tc::red_add_v4(dst + e, __uint_as_float(r[e]), __uint_as_float(r[e + 1]), __uint_as_float(r[e + 2]), __uint_as_float(r[e + 3]));
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem_base, 512);
}

// ----------------------------------------------------------------------------- launchers (called from fusion.cu's ABI entry points)
bool fusion_tc_enabled() {
    return true;
}

int fusion_weighted_fwd_tc(const void *cam_pre, const void *lid_pre, int64_t M,
                           const float *csc, const float *csh, const float *lsc, const float *lsh,
                           const float *w1, const float *b1, const float *w2, const float *b2,
                           void *out, float *attn, cudaStream_t st) {
    KDF_CHECK_ARG(((reinterpret_cast<uintptr_t>(cam_pre) | reinterpret_cast<uintptr_t>(lid_pre) | reinterpret_cast<uintptr_t>(out) |
                    reinterpret_cast<uintptr_t>(w1)) & 15) == 0 && (reinterpret_cast<uintptr_t>(attn) & 7) == 0,
                  "fusion_weighted_fwd: buffers must be 16-byte aligned");
    FusionTcArgs a{};
    a.cam = reinterpret_cast<const __nv_bfloat16 *>(cam_pre); a.lid = reinterpret_cast<const __nv_bfloat16 *>(lid_pre);
    a.M = M; a.csc = csc; a.csh = csh; a.lsc = lsc; a.lsh = lsh; a.w1 = w1; a.b1 = b1; a.w2 = w2; a.b2 = b2;
    a.out = reinterpret_cast<__nv_bfloat16 *>(out); a.attn = attn;
    const int64_t n_tiles = (M + FT_ROWS - 1) / FT_ROWS;
    const int blocks = (int)(n_tiles < sm_count() ? n_tiles : sm_count());
    CUtensorMap tm_cam, tm_lid;
    KDF_CHECK_ARG(M < (1ll << 31) && tma::make_row_map(&tm_cam, cam_pre, M, FT_C) && tma::make_row_map(&tm_lid, lid_pre, M, FT_C),
                  "fusion_weighted_fwd: cuTensorMapEncodeTiled failed");
    KDF_CUDA(cudaFuncSetAttribute(fusion_weighted_fwd_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FtTmaSmem::TOTAL));
    fusion_weighted_fwd_tma_kernel<<<blocks, FTM_THREADS, FtTmaSmem::TOTAL, st>>>(a, tm_cam, tm_lid);
    KDF_LAUNCH_CHECK();
    return KDF_OK;
}

int fusion_weighted_bwd_tc(const void *grad_out, const void *cam_pre, const void *lid_pre, int64_t M,
                           const float *csc, const float *csh, const float *lsc, const float *lsh,
                           const float *w1, const float *b1, const float *w2, const float *attn,
                           void *gcam, void *glid, float *gaff, float *gw1, float *gb1, float *gw2, float *gb2,
                           cudaStream_t st) {
    KDF_CHECK_ARG(((reinterpret_cast<uintptr_t>(cam_pre) | reinterpret_cast<uintptr_t>(lid_pre) | reinterpret_cast<uintptr_t>(grad_out) |
                    reinterpret_cast<uintptr_t>(gcam) | reinterpret_cast<uintptr_t>(glid) | reinterpret_cast<uintptr_t>(w1)) & 15) == 0 &&
                  (reinterpret_cast<uintptr_t>(attn) & 7) == 0,
                  "fusion_weighted_bwd: buffers must be 16-byte aligned");
    FusionTcArgs a{};
    a.cam = reinterpret_cast<const __nv_bfloat16 *>(cam_pre); a.lid = reinterpret_cast<const __nv_bfloat16 *>(lid_pre);
    a.gout = reinterpret_cast<const __nv_bfloat16 *>(grad_out);
    a.M = M; a.csc = csc; a.csh = csh; a.lsc = lsc; a.lsh = lsh; a.w1 = w1; a.b1 = b1; a.w2 = w2; a.b2 = nullptr;
    a.attn = const_cast<float *>(attn);
    a.gcam = reinterpret_cast<__nv_bfloat16 *>(gcam); a.glid = reinterpret_cast<__nv_bfloat16 *>(glid);
    a.gaff = gaff; a.gw1 = gw1; a.gb1 = gb1; a.gw2 = gw2; a.gb2 = gb2;
    const int64_t n_tiles = (M + FT_ROWS - 1) / FT_ROWS;
    const int blocks = (int)(n_tiles < sm_count() ? n_tiles : sm_count());
    KDF_CUDA(cudaFuncSetAttribute(fusion_weighted_bwd_tc_kernel<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, FtSmem::TOTAL_BWD));
    fusion_weighted_bwd_tc_kernel<512><<<blocks, 512, FtSmem::TOTAL_BWD, st>>>(a);
    KDF_LAUNCH_CHECK();
    return KDF_OK;
}

}  // namespace kdf
