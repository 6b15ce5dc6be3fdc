// LiDAR point cloud -> BEV grid projection for sm_100a.
//
// Replaces the index math and scatter of SpatialLiDAREncoder.forward_vectorized
// (reference src/models/lidar_encoder.py:42-55, 69-99) without its temporaries:
// the reference materialises six [B,N] fp32 tensors, an int64 [B,N,2] index, a
// boolean-mask gather of the [Nv,C] features and an int64 [Nv,C] expanded index;
// here a sweep is
//
//   index   : one coalesced float4 read per point -> cell id (bit-exact index
//             math) + warp-aggregated atomic occupancy count / arrival rank
//   scan    : per-frame exclusive scan of the occupancy -> segment offsets
//   fill    : counting-sort scatter of point ids into cell order
//   reduce  : one warp per cell streams that cell's feature rows (each row is a
//             contiguous, fully coalesced C*sizeof(T) read) and keeps the running
//             max (+ tie count) or sum in registers; every grid row is written once.
//
// Features are therefore read exactly once and never touched by atomics; the only
// atomics are one 4-byte add per (warp, distinct cell) in `index`.  Outputs are
// bitwise deterministic for max (order-independent).  HBM-bound: 16 B/point +
// C*s bytes per valid point + one grid write (kernel table).
#include <float.h>

#include <stdlib.h>
#include "kdf_common.cuh"

namespace kdf {

struct BevGeom {
    float x0, xspan, y0, yspan, sx, sy;   // sx = W-1, sy = H-1 as fp32 (grid_tensor)
    int H, W;
    int range_view;                       // 0: bird's-eye-view cells (the reference's grid); 1: spherical range image
    float fov_down, fov_span;             // range view: lowest elevation and vertical field of view, radians
};

// lidar_encoder.py:47-53,69-71 -- every operation rounded to fp32 separately
// (no FMA contraction, IEEE division), exactly like the eager ATen ops.
__device__ __forceinline__ int bev_cell_of(float x, float y, const BevGeom &g) {
    const float xn = __fdiv_rn(__fsub_rn(x, g.x0), g.xspan);
    const float yn = __fdiv_rn(__fsub_rn(y, g.y0), g.yspan);
    const bool valid = (xn >= 0.f) && (xn <= 1.f) && (yn >= 0.f) && (yn <= 1.f);   // NaN -> false
    if (!valid) return -1;
    int col = (int)__fmul_rn(xn, g.sx);      // .long(): truncation toward zero
    int row = (int)__fmul_rn(yn, g.sy);
    col = min(max(col, 0), g.W - 1);
    row = min(max(row, 0), g.H - 1);
    return row * g.W + col;
}

// Range-view (spherical) cell of a point: the other projection the distillation pipeline names next to the BEV grid.  The
// reference has no range-view code (parity unpinned); the convention is the usual one of
// range-image LiDAR networks: yaw = -atan2(y, x), pitch = asin(z / depth),
//   col = floor(0.5 * (yaw / pi + 1) * W),  row = floor((1 - (pitch - fov_down) / fov) * H),  both clamped to the image;
// points at the origin, with non-finite coordinates or outside the vertical field of view are invalid (-1).
__device__ __forceinline__ int range_cell_of(float x, float y, float z, const BevGeom &g) {
    const float depth = sqrtf(fmaf(x, x, fmaf(y, y, z * z)));
    if (!(depth > 0.f) || !(depth < 3.0e38f)) return -1;                 // origin, NaN, inf
    const float pitch = asinf(z / depth);
    const float t = (pitch - g.fov_down) / g.fov_span;
    if (!(t >= 0.f) || !(t <= 1.f)) return -1;
    const float yaw = -atan2f(y, x);
    int col = (int)floorf(0.5f * (yaw * 0.318309886183790672f + 1.0f) * (float)g.W);
    int row = (int)floorf((1.0f - t) * (float)g.H);
    col = min(max(col, 0), g.W - 1);
    row = min(max(row, 0), g.H - 1);
    return row * g.W + col;
}
// RV is a template parameter of the index kernels: as a run-time branch in their inner loops the choice cost the BEV cell-id
// kernel 17 % (0.024 -> 0.028 ms)
template <bool RV>
__device__ __forceinline__ int cell_of(float x, float y, float z, const BevGeom &g) {
    return RV ? range_cell_of(x, y, z, g) : bev_cell_of(x, y, g);
}

// ----------------------------------------------------------------------------- index
template <bool VEC4, bool RV>
__global__ void __launch_bounds__(256)
bev_index_kernel(const float *__restrict__ points, int64_t total, int64_t N, int stride, BevGeom g,
                 int32_t *__restrict__ cell_out, int32_t *__restrict__ rank_out, int32_t *__restrict__ count) {
    const int lane = threadIdx.x & 31;
    const int HW = g.H * g.W;
    const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
    // warp-uniform trip count: every lane takes part in the match/shuffle below
    const int64_t first = (int64_t)blockIdx.x * blockDim.x + (threadIdx.x & ~31);
    for (int64_t base = first; base < total; base += nthreads) {
        const int64_t i = base + lane;
        int cell = -1;
        int64_t key = -1;
        if (i < total) {
            float x, y, z = 0.f;
            if (VEC4) {
                const float4 p = ldg_stream_f4(reinterpret_cast<const float4 *>(points) + i);
                x = p.x; y = p.y; z = p.z;
            } else {
                x = __ldg(points + i * stride);
                y = __ldg(points + i * stride + 1);
                if (RV) z = __ldg(points + i * stride + 2);
            }
            cell = cell_of<RV>(x, y, z, g);
            cell_out[i] = cell;
            if (cell >= 0) key = (i / N) * HW + cell;
        }
        // warp-aggregated atomic: one add per distinct (frame, cell) in the warp
        const unsigned peers = __match_any_sync(0xffffffffu, key);
        const int leader = __ffs(peers) - 1;
        int r = 0;
        if (key >= 0 && lane == leader) r = atomicAdd(count + key, __popc(peers));
        r = __shfl_sync(0xffffffffu, r, leader);
        if (key >= 0 && rank_out) rank_out[i] = r + __popc(peers & ((1u << lane) - 1u));
    }
}

// Cell ids + occupancy only (no arrival ranks): a CTA walks a slice of one frame, counts into a SHARED-memory
// histogram with non-returning atomics and flushes its non-empty bins with one global add each -- HW adds per
// CTA instead of one per (warp, distinct cell).  The slice length is chosen by the host so that the grid still
// fills the SMs (launch_index).
template <bool RV>
__global__ void __launch_bounds__(512)
bev_index_hist_kernel(const float4 *__restrict__ points, int64_t N, int64_t slice, BevGeom g,
                      int32_t *__restrict__ cell_out, int32_t *__restrict__ count) {
    extern __shared__ int hist[];
    const int HW = g.H * g.W;
    const int b = blockIdx.y;
    for (int i = threadIdx.x; i < HW; i += 512) hist[i] = 0;
    __syncthreads();
    const int64_t beg = (int64_t)blockIdx.x * slice, end = (beg + slice < N) ? beg + slice : N;
    const float4 *pb = points + (int64_t)b * N;
    int32_t *cb = cell_out + (int64_t)b * N;
    for (int64_t i0 = beg + threadIdx.x; i0 < end; i0 += 4 * 512) {           // 4 independent points in flight per thread
        float4 p[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int64_t i = i0 + u * 512;
            if (i < end) p[u] = ldg_stream_f4(pb + i);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int64_t i = i0 + u * 512;
            if (i < end) {
                const int cell = cell_of<RV>(p[u].x, p[u].y, p[u].z, g);
                cb[i] = cell;
                if (cell >= 0) atomicAdd(&hist[cell], 1);
            }
        }
    }
    __syncthreads();
    int32_t *cnt = count + (int64_t)b * HW;
    for (int i = threadIdx.x; i < HW; i += 512) {
        const int h = hist[i];
        if (h) atomicAdd(cnt + i, h);
    }
}

// ----------------------------------------------------------------------------- scan
// One CTA per frame: offsets[b, 0..HW] = exclusive prefix of count[b, :].
__global__ void __launch_bounds__(1024)
bev_scan_kernel(const int32_t *__restrict__ count, int32_t *__restrict__ offsets, int HW) {
    __shared__ int warp_tot[32];
    __shared__ int carry_s;
    const int b = blockIdx.x, t = threadIdx.x, lane = t & 31, wid = t >> 5;
    const int32_t *cnt = count + (int64_t)b * HW;
    int32_t *off = offsets + (int64_t)b * (HW + 1);
    if (t == 0) carry_s = 0;
    __syncthreads();
    for (int c0 = 0; c0 < HW; c0 += 1024) {
        const int c = c0 + t;
        const int v = (c < HW) ? cnt[c] : 0;
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int n = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += n;
        }
        if (lane == 31) warp_tot[wid] = incl;
        __syncthreads();
        if (wid == 0) {
            int w = warp_tot[lane], wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int n = __shfl_up_sync(0xffffffffu, wi, o);
                if (lane >= o) wi += n;
            }
            warp_tot[lane] = wi - w;           // exclusive prefix of the warp totals
        }
        __syncthreads();
        const int carry = carry_s;
        const int excl = carry + warp_tot[wid] + incl - v;
        if (c < HW) off[c] = excl;
        __syncthreads();
        if (t == 1023) carry_s = excl + v;
        __syncthreads();
    }
    if (t == 0) off[HW] = carry_s;
}

// ----------------------------------------------------------------------------- counting sort without global atomics
// The warp-aggregated global atomics of bev_index_kernel are bound by the L2 atomic rate (~50 G adds/s: 67 us for
// 5.4 M points).  For the cell ORDERING a frame is cut into chunks of BEV_CHUNK points, one CTA per chunk:
//   count : cell id per point + arrival rank inside the chunk from SHARED-memory atomics on a per-chunk histogram,
//           which is then written out [B][chunk][HW]
//   scan  : per frame, per cell: running sum over the chunks (in place: histogram -> chunk base), occupancy, offsets
//   fill  : order[offsets[cell] + chunk_base[cell] + rank] = point id, bases staged in shared memory
// No global atomic is left; cell ids, occupancy and offsets are identical to the atomic path, the order inside a cell
// differs (both are valid: every reduction over a cell is order-independent).
constexpr int BEV_CHUNK = 8192;

template <bool RV>
__global__ void __launch_bounds__(256)
bev_chunk_count_kernel(const float4 *__restrict__ points, int64_t N, BevGeom g, int32_t *__restrict__ cell_out,
                       int32_t *__restrict__ rank_out, int32_t *__restrict__ chist, int nchunk,
                       int32_t *__restrict__ cout /* nullable [B][nchunk]: points outside the grid, ranked too */) {
    extern __shared__ int hist[];
    __shared__ int n_out;
    const int HW = g.H * g.W;
    const int chunk = blockIdx.x, b = blockIdx.y;
    for (int i = threadIdx.x; i < HW; i += 256) hist[i] = 0;
    if (threadIdx.x == 0) n_out = 0;
    __syncthreads();
    const int64_t beg = (int64_t)chunk * BEV_CHUNK, end = (beg + BEV_CHUNK < N) ? beg + BEV_CHUNK : N;
    const float4 *pb = points + (int64_t)b * N;
    int32_t *cb = cell_out + (int64_t)b * N, *rb = rank_out + (int64_t)b * N;
    for (int64_t i0 = beg + threadIdx.x; i0 < end; i0 += 4 * 256) {           // 4 independent points in flight per thread
        float4 p[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int64_t i = i0 + u * 256;
            if (i < end) p[u] = ldg_stream_f4(pb + i);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int64_t i = i0 + u * 256;
            if (i < end) {
                const int cell = cell_of<RV>(p[u].x, p[u].y, p[u].z, g);
                cb[i] = cell;
                if (cell >= 0) rb[i] = atomicAdd(&hist[cell], 1);
                else if (cout) rb[i] = atomicAdd(&n_out, 1);
            }
        }
    }
    __syncthreads();
    int32_t *out = chist + ((int64_t)b * nchunk + chunk) * HW;
    for (int i = threadIdx.x; i < HW; i += 256) out[i] = hist[i];
    if (cout && threadIdx.x == 0) cout[(int64_t)b * nchunk + chunk] = n_out;
}

__global__ void __launch_bounds__(1024)
bev_chunk_scan_kernel(int32_t *__restrict__ chist /* in: histograms, out: chunk bases */, int32_t *__restrict__ count,
                      int32_t *__restrict__ offsets, int HW, int nchunk, int32_t *__restrict__ cout /* nullable, in place */) {
    __shared__ int warp_tot[32];
    __shared__ int carry_s;
    const int b = blockIdx.x, t = threadIdx.x, lane = t & 31, wid = t >> 5;
    int32_t *cnt = count + (int64_t)b * HW;
    int32_t *off = offsets + (int64_t)b * (HW + 1);
    int32_t *hb = chist + (int64_t)b * nchunk * HW;
    if (t == 0) carry_s = 0;
    __syncthreads();
    for (int c0 = 0; c0 < HW; c0 += 1024) {
        const int c = c0 + t;
        int v = 0;
        if (c < HW) {
            for (int k0 = 0; k0 < nchunk; k0 += 8) {                           // 8 independent loads, then the running sum
                int h[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) h[u] = (k0 + u < nchunk) ? hb[(int64_t)(k0 + u) * HW + c] : 0;
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    if (k0 + u < nchunk) hb[(int64_t)(k0 + u) * HW + c] = v;
                    v += h[u];
                }
            }
            cnt[c] = v;
        }
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int n = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += n;
        }
        if (lane == 31) warp_tot[wid] = incl;
        __syncthreads();
        if (wid == 0) {
            int w = warp_tot[lane], wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int n = __shfl_up_sync(0xffffffffu, wi, o);
                if (lane >= o) wi += n;
            }
            warp_tot[lane] = wi - w;
        }
        __syncthreads();
        const int carry = carry_s;
        const int excl = carry + warp_tot[wid] + incl - v;
        if (c < HW) off[c] = excl;
        __syncthreads();
        if (t == 1023) carry_s = excl + v;
        __syncthreads();
    }
    if (t == 0) off[HW] = carry_s;
    if (cout && t == 0) {                                                    // outside points of chunk k start behind those of chunks < k
        int run = 0;
        for (int k = 0; k < nchunk; ++k) {
            const int v = cout[(int64_t)b * nchunk + k];
            cout[(int64_t)b * nchunk + k] = run;
            run += v;
        }
    }
}

__global__ void __launch_bounds__(256)
bev_chunk_fill_kernel(const int32_t *__restrict__ cell, const int32_t *__restrict__ rank, const int32_t *__restrict__ cbase,
                      const int32_t *__restrict__ offsets, int32_t *__restrict__ order, int64_t N, int HW, int nchunk) {
    extern __shared__ int base[];
    const int chunk = blockIdx.x, b = blockIdx.y;
    const int32_t *cbp = cbase + ((int64_t)b * nchunk + chunk) * HW, *off = offsets + (int64_t)b * (HW + 1);
    for (int i = threadIdx.x; i < HW; i += 256) base[i] = cbp[i] + off[i];
    __syncthreads();
    const int64_t beg = (int64_t)chunk * BEV_CHUNK, end = (beg + BEV_CHUNK < N) ? beg + BEV_CHUNK : N;
    const int32_t *cb = cell + (int64_t)b * N, *rb = rank + (int64_t)b * N;
    int32_t *ob = order + (int64_t)b * N;
    for (int64_t i0 = beg + threadIdx.x; i0 < end; i0 += 4 * 256) {
        int c[4], r[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int64_t i = i0 + u * 256;
            c[u] = (i < end) ? __ldg(cb + i) : -1;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) r[u] = (c[u] >= 0) ? __ldg(rb + i0 + u * 256) : 0;
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (c[u] >= 0) ob[base[c[u]] + r[u]] = (int32_t)(i0 + u * 256);
    }
}

// The same pass writing the POINTS in cell order: frame b's rows [offsets[c], offsets[c+1]) are the points of
// cell c, the points outside the grid follow from offsets[HW] on.  cell_sorted holds the global cell id b*HW + c of every
// sorted row (-1 outside).  Everything downstream of the point MLP then works on contiguous row segments.
__global__ void __launch_bounds__(256)
bev_chunk_permute_kernel(const float4 *__restrict__ points, const int32_t *__restrict__ cell, const int32_t *__restrict__ rank,
                         const int32_t *__restrict__ cbase, const int32_t *__restrict__ cout, const int32_t *__restrict__ offsets,
                         float4 *__restrict__ sorted, int32_t *__restrict__ cell_sorted, int32_t *__restrict__ order /* nullable */,
                         int64_t N, int HW, int nchunk) {
    extern __shared__ int base[];
    const int chunk = blockIdx.x, b = blockIdx.y;
    const int32_t *cbp = cbase + ((int64_t)b * nchunk + chunk) * HW, *off = offsets + (int64_t)b * (HW + 1);
    for (int i = threadIdx.x; i < HW; i += 256) base[i] = cbp[i] + off[i];
    __syncthreads();
    const int out_base = off[HW] + cout[(int64_t)b * nchunk + chunk];
    const int64_t beg = (int64_t)chunk * BEV_CHUNK, end = (beg + BEV_CHUNK < N) ? beg + BEV_CHUNK : N;
    const int32_t *cb = cell + (int64_t)b * N, *rb = rank + (int64_t)b * N;
    const float4 *pb = points + (int64_t)b * N;
    float4 *sb = sorted + (int64_t)b * N;
    int32_t *csb = cell_sorted + (int64_t)b * N, *ob = order ? order + (int64_t)b * N : nullptr;
    for (int64_t i0 = beg + threadIdx.x; i0 < end; i0 += 4 * 256) {
        int c[4], r[4];
        float4 p[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int64_t i = i0 + u * 256;
            if (i < end) { c[u] = __ldg(cb + i); r[u] = __ldg(rb + i); p[u] = ldg_stream_f4(pb + i); }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int64_t i = i0 + u * 256;
            if (i < end) {
                const int pos = (c[u] >= 0 ? base[c[u]] : out_base) + r[u];
                sb[pos] = p[u];
                csb[pos] = c[u] >= 0 ? b * HW + c[u] : -1;
                if (ob) ob[pos] = (int32_t)i;
            }
        }
    }
}

// ----------------------------------------------------------------------------- fill
__global__ void __launch_bounds__(256)
bev_fill_kernel(const int32_t *__restrict__ cell, const int32_t *__restrict__ rank,
                const int32_t *__restrict__ offsets, int32_t *__restrict__ order,
                int64_t total, int64_t N, int HW) {
    const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += nthreads) {
        const int c = cell[i];
        if (c < 0) continue;
        const int64_t b = i / N;
        const int pos = offsets[b * (HW + 1) + c] + rank[i];
        order[b * N + pos] = (int32_t)(i - b * N);
    }
}

// ----------------------------------------------------------------------------- reduce
// One warp per (frame, cell); lane l owns channels [c0 + 4l, c0 + 4l + 4).
template <typename T, int REDUCE>
__global__ void __launch_bounds__(256)
bev_reduce_kernel(const T *__restrict__ feats, const int32_t *__restrict__ order,
                  const int32_t *__restrict__ offsets, T *__restrict__ grid, int32_t *__restrict__ ties,
                  int64_t n_cells, int64_t N, int C, int HW) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t cid = warp0; cid < n_cells; cid += nwarps) {
        const int64_t b = cid / HW;
        const int c = (int)(cid - b * HW);
        const int beg = offsets[b * (HW + 1) + c];
        const int n = offsets[b * (HW + 1) + c + 1] - beg;
        const int32_t *ord = order + b * N + beg;
        const T *fb = feats + b * N * C;
        for (int c0 = 0; c0 < C; c0 += 128) {
            const int ch = c0 + lane * 4;
            const bool act = ch < C;
            float m[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};   // running max, or sum for mean
            int k[4] = {0, 0, 0, 0};
            if (REDUCE == KDF_REDUCE_MEAN) m[0] = m[1] = m[2] = m[3] = 0.f;
            for (int j0 = 0; j0 < n; j0 += 32) {
                const int mine = (j0 + lane < n) ? ord[j0 + lane] : 0;
                const int lim = min(32, n - j0);
                for (int jj = 0; jj < lim; jj += 4) {
                    float4 v[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int p = __shfl_sync(0xffffffffu, mine, (jj + u) & 31);
                        if (jj + u < lim && act) v[u] = Vec4<T>::load_stream(fb + (int64_t)p * C + ch);
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        if (jj + u < lim && act) {
                            const float f[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                if (REDUCE == KDF_REDUCE_MAX) {
                                    if (f[q] > m[q]) { m[q] = f[q]; k[q] = 1; }
                                    else if (f[q] == m[q]) { k[q]++; }
                                } else {
                                    m[q] += f[q];
                                }
                            }
                        }
                    }
                }
            }
            if (act) {
                float4 o;
                if (n == 0) {
                    o = make_float4(0.f, 0.f, 0.f, 0.f);       // include_self=False into zeros
                } else if (REDUCE == KDF_REDUCE_MEAN) {
                    const float d = (float)n;
                    o = make_float4(m[0] / d, m[1] / d, m[2] / d, m[3] / d);
                } else {
                    o = make_float4(m[0], m[1], m[2], m[3]);
                }
                Vec4<T>::store(grid + cid * C + ch, o);
                if (REDUCE == KDF_REDUCE_MAX && ties)
                    *reinterpret_cast<int4 *>(ties + cid * C + ch) = make_int4(k[0], k[1], k[2], k[3]);
            }
        }
    }
}

// ----------------------------------------------------------------------------- backward
// One warp per point row; rows of the same cell hit grid/ties/grad_grid in L2.
template <typename T, int REDUCE>
__global__ void __launch_bounds__(256)
bev_bwd_kernel(const T *__restrict__ grad_grid, const T *__restrict__ feats, const T *__restrict__ grid,
               const int32_t *__restrict__ ties, const int32_t *__restrict__ count,
               const int32_t *__restrict__ cell, T *__restrict__ grad_feats,
               int64_t total, int64_t N, int C, int HW) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    constexpr int U = 4;
    for (int64_t i0 = warp0 * U; i0 < total; i0 += nwarps * U) {
        int cl[U];
#pragma unroll
        for (int u = 0; u < U; ++u) cl[u] = (i0 + u < total) ? __ldg(cell + i0 + u) : -2;
        for (int c0 = 0; c0 < C; c0 += 128) {
            const int ch = c0 + lane * 4;
            if (ch >= C) continue;
            float4 f[U], g[U], m[U];
            int4 t[U];
            int cnt[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (cl[u] >= 0) {
                    const int64_t b = (i0 + u) / N;
                    const int64_t cid = b * HW + cl[u];
                    g[u] = Vec4<T>::load(grad_grid + cid * C + ch);
                    if (REDUCE == KDF_REDUCE_MAX) {
                        f[u] = Vec4<T>::load_stream(feats + (i0 + u) * C + ch);
                        m[u] = Vec4<T>::load(grid + cid * C + ch);
                        t[u] = *reinterpret_cast<const int4 *>(ties + cid * C + ch);
                    } else {
                        cnt[u] = __ldg(count + cid);
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (cl[u] == -2) continue;
                float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
                if (cl[u] >= 0) {
                    if (REDUCE == KDF_REDUCE_MAX) {
                        // ATen: N_to_distribute = (self == result) + #(src == result); self is 0
                        o.x = (f[u].x == m[u].x) ? g[u].x / (float)(t[u].x + (m[u].x == 0.f)) : 0.f;
                        o.y = (f[u].y == m[u].y) ? g[u].y / (float)(t[u].y + (m[u].y == 0.f)) : 0.f;
                        o.z = (f[u].z == m[u].z) ? g[u].z / (float)(t[u].z + (m[u].z == 0.f)) : 0.f;
                        o.w = (f[u].w == m[u].w) ? g[u].w / (float)(t[u].w + (m[u].w == 0.f)) : 0.f;
                    } else {
                        const float d = (float)cnt[u];
                        o = make_float4(g[u].x / d, g[u].y / d, g[u].z / d, g[u].w / d);
                    }
                }
                Vec4<T>::store(grad_feats + (i0 + u) * C + ch, o);
            }
        }
    }
}


// ----------------------------------------------------------------------------- reduce / backward, wide-load variants
// Same contract as the kernels above for the common shapes where one row is LPR 16-byte lanes
// (LPR in {8,16,32}: C = 64/128/256 bf16 or 32/64/128 fp32).  A warp owns a cell; its 32 lanes
// are split into RPL = 32/LPR row groups so that every lane always moves 16 bytes, U row groups
// are in flight per iteration and the point ids of the next iteration are prefetched while the
// current rows are consumed: enough bytes in flight to hide HBM latency at 32 warps/SM.
template <typename T> struct Raw16;
template <> struct Raw16<float> {
    static constexpr int VEC = 4;
    static __device__ __forceinline__ void unpack(const uint4 &u, float *v) {
        v[0] = __uint_as_float(u.x); v[1] = __uint_as_float(u.y); v[2] = __uint_as_float(u.z); v[3] = __uint_as_float(u.w);
    }
    static __device__ __forceinline__ uint4 pack(const float *v) {
        return make_uint4(__float_as_uint(v[0]), __float_as_uint(v[1]), __float_as_uint(v[2]), __float_as_uint(v[3]));
    }
};
template <> struct Raw16<__nv_bfloat16> {
    static constexpr int VEC = 8;
    static __device__ __forceinline__ void unpack(const uint4 &u, float *v) {
        v[0] = bf16_lo(u.x); v[1] = bf16_hi(u.x); v[2] = bf16_lo(u.y); v[3] = bf16_hi(u.y);
        v[4] = bf16_lo(u.z); v[5] = bf16_hi(u.z); v[6] = bf16_lo(u.w); v[7] = bf16_hi(u.w);
    }
    static __device__ __forceinline__ uint4 pack(const float *v) {
        return make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
    }
};

template <typename T, int LPR, int REDUCE>
__global__ void __launch_bounds__(256)
bev_reduce_wide_kernel(const T *__restrict__ feats, const int32_t *__restrict__ order,
                       const int32_t *__restrict__ offsets, T *__restrict__ grid, int32_t *__restrict__ ties,
                       int64_t n_cells, int64_t N, int HW) {
    constexpr int VEC = Raw16<T>::VEC, C = LPR * VEC, RPL = 32 / LPR, U = 4, STEP = RPL * U;
    const int lane = threadIdx.x & 31, sub = lane / LPR, ch = (lane % LPR) * VEC;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t cid = warp0; cid < n_cells; cid += nwarps) {
        const int64_t b = cid / HW;
        const int c = (int)(cid - b * HW);
        const int beg = __ldg(offsets + b * (HW + 1) + c);
        const int n = __ldg(offsets + b * (HW + 1) + c + 1) - beg;
        const int32_t *ord = order + b * N + beg;
        const T *fb = feats + b * N * C + ch;
        float m[VEC];
        int k[VEC];
#pragma unroll
        for (int q = 0; q < VEC; ++q) { m[q] = (REDUCE == KDF_REDUCE_MAX) ? -INFINITY : 0.f; k[q] = 0; }
        int idn[U];
#pragma unroll
        for (int u = 0; u < U; ++u) { const int j = u * RPL + sub; idn[u] = j < n ? __ldg(ord + j) : -1; }
        for (int j0 = 0; j0 < n; j0 += STEP) {
            uint4 raw[U];
            int id[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                id[u] = idn[u];
                if (id[u] >= 0) raw[u] = ldg_stream_u4(reinterpret_cast<const uint4 *>(fb + (int64_t)id[u] * C));
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {                      // ids of the next iteration, while the rows fly
                const int j = j0 + STEP + u * RPL + sub;
                idn[u] = j < n ? __ldg(ord + j) : -1;
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (id[u] >= 0) {
                    float f[VEC];
                    Raw16<T>::unpack(raw[u], f);
#pragma unroll
                    for (int q = 0; q < VEC; ++q) {
                        if (REDUCE == KDF_REDUCE_MAX) {
                            if (f[q] > m[q]) { m[q] = f[q]; k[q] = 1; }
                            else if (f[q] == m[q]) { k[q]++; }
                        } else {
                            m[q] += f[q];
                        }
                    }
                }
            }
        }
        // merge the RPL row groups (lanes that own the same channels)
#pragma unroll
        for (int o = LPR; o < 32; o <<= 1) {
#pragma unroll
            for (int q = 0; q < VEC; ++q) {
                const float om = __shfl_xor_sync(0xffffffffu, m[q], o);
                const int ok = __shfl_xor_sync(0xffffffffu, k[q], o);
                if (REDUCE == KDF_REDUCE_MAX) {
                    if (om > m[q]) { m[q] = om; k[q] = ok; }
                    else if (om == m[q]) { k[q] += ok; }
                } else {
                    m[q] += om;
                }
            }
        }
        if (sub == 0) {
            float o[VEC];
#pragma unroll
            for (int q = 0; q < VEC; ++q)
                o[q] = (n == 0) ? 0.f : (REDUCE == KDF_REDUCE_MEAN ? m[q] / (float)n : m[q]);
            *reinterpret_cast<uint4 *>(grid + cid * C + ch) = Raw16<T>::pack(o);
            if (REDUCE == KDF_REDUCE_MAX && ties) {
                int *tp = ties + cid * C + ch;
#pragma unroll
                for (int q = 0; q < VEC; q += 4)
                    *reinterpret_cast<int4 *>(tp + q) = make_int4(n ? k[q] : 0, n ? k[q + 1] : 0, n ? k[q + 2] : 0, n ? k[q + 3] : 0);
            }
        }
    }
}

// Cell-major backward: a warp loads its cell's grad / max / tie rows once, then streams the
// cell's feature rows and writes their gradient rows.  A second, point-major sweep zeroes the
// rows of the points that fell outside the grid (they are in no cell list).
template <typename T, int LPR, int REDUCE>
__global__ void __launch_bounds__(256)
bev_bwd_wide_kernel(const T *__restrict__ grad_grid, const T *__restrict__ feats, const T *__restrict__ grid,
                    const int32_t *__restrict__ ties, const int32_t *__restrict__ order,
                    const int32_t *__restrict__ offsets, const int32_t *__restrict__ cell,
                    T *__restrict__ grad_feats, int64_t n_cells, int64_t N, int HW, int64_t total) {
    constexpr int VEC = Raw16<T>::VEC, C = LPR * VEC, RPL = 32 / LPR, U = 4, STEP = RPL * U;
    const int lane = threadIdx.x & 31, sub = lane / LPR, ch = (lane % LPR) * VEC;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t cid = warp0; cid < n_cells; cid += nwarps) {
        const int64_t b = cid / HW;
        const int c = (int)(cid - b * HW);
        const int beg = __ldg(offsets + b * (HW + 1) + c);
        const int n = __ldg(offsets + b * (HW + 1) + c + 1) - beg;
        if (n == 0) continue;
        const int32_t *ord = order + b * N + beg;
        float g[VEC], mx[VEC];
        Raw16<T>::unpack(*reinterpret_cast<const uint4 *>(grad_grid + cid * C + ch), g);
        if (REDUCE == KDF_REDUCE_MAX) {
            Raw16<T>::unpack(*reinterpret_cast<const uint4 *>(grid + cid * C + ch), mx);
#pragma unroll
            for (int q = 0; q < VEC; ++q) {
                // ATen: N_to_distribute = (self == result) + #(src == result), self is the zero init
                const int t = __ldg(ties + cid * C + ch + q) + (mx[q] == 0.f ? 1 : 0);
                g[q] = g[q] / (float)t;
            }
        } else {
#pragma unroll
            for (int q = 0; q < VEC; ++q) g[q] = g[q] / (float)n;
        }
        int idn[U];
#pragma unroll
        for (int u = 0; u < U; ++u) { const int j = u * RPL + sub; idn[u] = j < n ? __ldg(ord + j) : -1; }
        for (int j0 = 0; j0 < n; j0 += STEP) {
            uint4 raw[U];
            int id[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                id[u] = idn[u];
                if (REDUCE == KDF_REDUCE_MAX && id[u] >= 0)
                    raw[u] = ldg_stream_u4(reinterpret_cast<const uint4 *>(feats + (b * N + id[u]) * C + ch));
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int j = j0 + STEP + u * RPL + sub;
                idn[u] = j < n ? __ldg(ord + j) : -1;
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (id[u] >= 0) {
                    float o[VEC];
                    if (REDUCE == KDF_REDUCE_MAX) {
                        float f[VEC];
                        Raw16<T>::unpack(raw[u], f);
#pragma unroll
                        for (int q = 0; q < VEC; ++q) o[q] = (f[q] == mx[q]) ? g[q] : 0.f;
                    } else {
#pragma unroll
                        for (int q = 0; q < VEC; ++q) o[q] = g[q];
                    }
                    *reinterpret_cast<uint4 *>(grad_feats + (b * N + id[u]) * C + ch) = Raw16<T>::pack(o);
                }
            }
        }
    }
    // rows of points outside the grid
    const int64_t g0 = (((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / LPR), gn = ((int64_t)gridDim.x * blockDim.x) / LPR;
    const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
    for (int64_t i = g0; i < total; i += gn)
        if (__ldg(cell + i) < 0) *reinterpret_cast<uint4 *>(grad_feats + i * C + ch) = zero;
}


// ----------------------------------------------------------------------------- projection of a BN+ReLU'd layer (fused MLP path)
// The fused point-MLP keeps only the pre-BatchNorm output z3 of the last layer in HBM (bf16); the feature the
// reference scatters is a3 = relu(z3*scale + shift) (lidar_encoder.py:32-34 then :85-96).  BatchNorm-apply,
// ReLU and the rounding to bf16 are all monotonic, so the per-cell maximum of a3 is a3 of the per-cell
// EXTREME of z3 (maximum where scale >= 0, minimum where scale < 0): the reduction runs on the raw packed
// bf16 rows -- one XOR (sign flip for negative scales) and one packed max per two channels -- and the
// affine+ReLU is applied once per cell.  Forward stores both the extreme (grid_z) and its activation (grid).
// Backward: the gradient of a cell goes, split evenly, to the rows whose stored z equals the extreme (the
// arg-max rows), if the activation is positive; the two column sums BatchNorm's backward needs follow per
// CELL (S0 += k*share, S1 += k*share*z_ext), not per row.
// Persistent warps walk cells with a one-cell-ahead prefetch of the cell's offsets and first point ids, so
// the offsets -> ids -> rows dependency chain of one cell overlaps the row traffic of the previous one.
__device__ __forceinline__ float bf16_round(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }

__device__ __forceinline__ uint32_t bf16x2_max(uint32_t a, uint32_t b) {
    uint32_t d;
    asm("max.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
}
__device__ __forceinline__ uint32_t bf16x2_eq_mask(uint32_t a, uint32_t b) {          // 0xFFFF per equal half
    uint32_t d;
    asm("set.eq.u32.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
}

struct CellMeta { int beg, n; int64_t rowbase; };
__device__ __forceinline__ CellMeta cell_meta(const int32_t *__restrict__ offsets, int64_t cid, int64_t n_cells, int64_t N, int HW) {
    CellMeta m{0, 0, 0};
    if (cid < n_cells) {
        const int64_t b = cid / HW;
        const int c = (int)(cid - b * HW);
        m.beg = __ldg(offsets + b * (HW + 1) + c);
        m.n = __ldg(offsets + b * (HW + 1) + c + 1) - m.beg;
        m.rowbase = b * N;
    }
    return m;
}

template <int LPR, bool ORD>
__global__ void __launch_bounds__(256, 4)
bev_reduce_affine_kernel(const __nv_bfloat16 *__restrict__ z, const float *__restrict__ scale, const float *__restrict__ shift,
                         const int32_t *__restrict__ order, const int32_t *__restrict__ offsets,
                         __nv_bfloat16 *__restrict__ grid, __nv_bfloat16 *__restrict__ grid_z, int64_t n_cells, int64_t N, int HW) {
    constexpr int C = LPR * 8, RPL = 32 / LPR, U = 4, STEP = RPL * U;
    const int lane = threadIdx.x & 31, sub = lane / LPR, ch = (lane % LPR) * 8;
    uint32_t flip[4];
#pragma unroll
    for (int q = 0; q < 4; ++q)
        flip[q] = (__ldg(scale + ch + 2 * q) < 0.f ? 0x8000u : 0u) | (__ldg(scale + ch + 2 * q + 1) < 0.f ? 0x80000000u : 0u);
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;

    int64_t cid = warp0;
    CellMeta cur = cell_meta(offsets, cid, n_cells, N, HW);
    int idn[U];
#pragma unroll
    for (int u = 0; u < U; ++u) { const int j = u * RPL + sub; idn[u] = j < cur.n ? (ORD ? __ldg(order + cur.rowbase + cur.beg + j) : cur.beg + j) : -1; }
    CellMeta nxt = cell_meta(offsets, cid + nwarps, n_cells, N, HW);
    while (cid < n_cells) {
        // one cell ahead: first ids of the next cell (its offsets arrived during the previous cell), offsets of the one after
        int idn2[U];
#pragma unroll
        for (int u = 0; u < U; ++u) { const int j = u * RPL + sub; idn2[u] = j < nxt.n ? (ORD ? __ldg(order + nxt.rowbase + nxt.beg + j) : nxt.beg + j) : -1; }
        const CellMeta nxt2 = cell_meta(offsets, cid + 2 * nwarps, n_cells, N, HW);

        const int32_t *ord = ORD ? order + cur.rowbase + cur.beg : nullptr;   // !ORD: the rows are already in cell order
        const __nv_bfloat16 *zb = z + cur.rowbase * C + ch;
        uint32_t m[4] = {0xFF80FF80u, 0xFF80FF80u, 0xFF80FF80u, 0xFF80FF80u};           // -inf, -inf
        for (int j0 = 0; j0 < cur.n; j0 += STEP) {
            uint4 raw[U];
            int id[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                id[u] = idn[u];
                if (id[u] >= 0) raw[u] = ldg_stream_u4(reinterpret_cast<const uint4 *>(zb + (int64_t)id[u] * C));
            }
#pragma unroll
            for (int u = 0; u < U; ++u) { const int j = j0 + STEP + u * RPL + sub; idn[u] = j < cur.n ? (ORD ? __ldg(ord + j) : cur.beg + j) : -1; }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (id[u] >= 0) {
                    m[0] = bf16x2_max(m[0], raw[u].x ^ flip[0]);
                    m[1] = bf16x2_max(m[1], raw[u].y ^ flip[1]);
                    m[2] = bf16x2_max(m[2], raw[u].z ^ flip[2]);
                    m[3] = bf16x2_max(m[3], raw[u].w ^ flip[3]);
                }
            }
        }
#pragma unroll
        for (int o = LPR; o < 32; o <<= 1) {
#pragma unroll
            for (int q = 0; q < 4; ++q) m[q] = bf16x2_max(m[q], __shfl_xor_sync(0xffffffffu, m[q], o));
        }
        if (sub == 0) {
            uint4 ze = make_uint4(0u, 0u, 0u, 0u), av = make_uint4(0u, 0u, 0u, 0u);
            if (cur.n > 0) {
                ze = make_uint4(m[0] ^ flip[0], m[1] ^ flip[1], m[2] ^ flip[2], m[3] ^ flip[3]);
                float f[8];
                Raw16<__nv_bfloat16>::unpack(ze, f);
                const float4 s0 = __ldg(reinterpret_cast<const float4 *>(scale + ch)), s1 = __ldg(reinterpret_cast<const float4 *>(scale + ch + 4));
                const float4 t0 = __ldg(reinterpret_cast<const float4 *>(shift + ch)), t1 = __ldg(reinterpret_cast<const float4 *>(shift + ch + 4));
                f[0] = fmaxf(fmaf(f[0], s0.x, t0.x), 0.f); f[1] = fmaxf(fmaf(f[1], s0.y, t0.y), 0.f);
                f[2] = fmaxf(fmaf(f[2], s0.z, t0.z), 0.f); f[3] = fmaxf(fmaf(f[3], s0.w, t0.w), 0.f);
                f[4] = fmaxf(fmaf(f[4], s1.x, t1.x), 0.f); f[5] = fmaxf(fmaf(f[5], s1.y, t1.y), 0.f);
                f[6] = fmaxf(fmaf(f[6], s1.z, t1.z), 0.f); f[7] = fmaxf(fmaf(f[7], s1.w, t1.w), 0.f);
                av = Raw16<__nv_bfloat16>::pack(f);
            }
            *reinterpret_cast<uint4 *>(grid + cid * C + ch) = av;
            if (grid_z) *reinterpret_cast<uint4 *>(grid_z + cid * C + ch) = ze;
        }
        cid += nwarps;
        cur = nxt;
        nxt = nxt2;
#pragma unroll
        for (int u = 0; u < U; ++u) idn[u] = idn2[u];
    }
}

template <int LPR, int MINB>
__global__ void __launch_bounds__(256, MINB)
bev_bwd_affine_kernel(const __nv_bfloat16 *__restrict__ grad_grid, const __nv_bfloat16 *__restrict__ z,
                      const __nv_bfloat16 *__restrict__ grid, const __nv_bfloat16 *__restrict__ grid_z,
                      const int32_t *__restrict__ order, const int32_t *__restrict__ offsets, const int32_t *__restrict__ cell,
                      __nv_bfloat16 *__restrict__ dy, double *__restrict__ sums /* [2][C] */,
                      int64_t n_cells, int64_t N, int HW, int64_t total) {
    using T = __nv_bfloat16;
    constexpr int C = LPR * 8, RPL = 32 / LPR, U = 4, STEP = RPL * U;
    __shared__ float red[8][2][LPR * 8];                                   // [warp][S0|S1][q * LPR + lane]: channel 8*lane + q (lanes hit distinct banks)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, sub = lane / LPR, ch = (lane % LPR) * 8;
    // per-warp column sums live in shared memory (touched once per cell by the sub-0 lanes, each its own slots)
    const int rl = lane % LPR;                                             // this lane's column of the per-warp sums
    if (sub == 0) {
#pragma unroll
        for (int q = 0; q < 8; ++q) { red[warp][0][q * LPR + rl] = 0.f; red[warp][1][q * LPR + rl] = 0.f; }
    }
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;

    int64_t cid = warp0;
    CellMeta cur = cell_meta(offsets, cid, n_cells, N, HW);
    int idn[U];
#pragma unroll
    for (int u = 0; u < U; ++u) { const int j = u * RPL + sub; idn[u] = j < cur.n ? __ldg(order + cur.rowbase + cur.beg + j) : -1; }
    CellMeta nxt = cell_meta(offsets, cid + nwarps, n_cells, N, HW);
    while (cid < n_cells) {
        int idn2[U];
#pragma unroll
        for (int u = 0; u < U; ++u) { const int j = u * RPL + sub; idn2[u] = j < nxt.n ? __ldg(order + nxt.rowbase + nxt.beg + j) : -1; }
        const CellMeta nxt2 = cell_meta(offsets, cid + 2 * nwarps, n_cells, N, HW);
        if (cur.n > 0) {
            const int32_t *ord = order + cur.rowbase + cur.beg;
            const T *zb = z + cur.rowbase * C + ch;
            T *db = dy + cur.rowbase * C + ch;
            const uint4 ze = __ldg(reinterpret_cast<const uint4 *>(grid_z + cid * C + ch));
            const uint4 gv = __ldg(reinterpret_cast<const uint4 *>(grad_grid + cid * C + ch));
            const uint4 av = __ldg(reinterpret_cast<const uint4 *>(grid + cid * C + ch));
            // pass A: how many rows sit at the extreme, per channel
            int k[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) k[q] = 0;
            int ida[U];
#pragma unroll
            for (int u = 0; u < U; ++u) ida[u] = idn[u];
            for (int j0 = 0; j0 < cur.n; j0 += STEP) {
                uint4 raw[U];
                int id[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    id[u] = ida[u];
                    if (id[u] >= 0) raw[u] = __ldg(reinterpret_cast<const uint4 *>(zb + (int64_t)id[u] * C));   // L1-allocating: pass B re-reads
                }
#pragma unroll
                for (int u = 0; u < U; ++u) { const int j = j0 + STEP + u * RPL + sub; ida[u] = j < cur.n ? __ldg(ord + j) : -1; }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    if (id[u] >= 0) {
                        const uint32_t e0 = bf16x2_eq_mask(raw[u].x, ze.x), e1 = bf16x2_eq_mask(raw[u].y, ze.y);
                        const uint32_t e2 = bf16x2_eq_mask(raw[u].z, ze.z), e3 = bf16x2_eq_mask(raw[u].w, ze.w);
                        k[0] += e0 & 1; k[1] += e0 >> 31; k[2] += e1 & 1; k[3] += e1 >> 31;
                        k[4] += e2 & 1; k[5] += e2 >> 31; k[6] += e3 & 1; k[7] += e3 >> 31;
                    }
                }
            }
#pragma unroll
            for (int o = LPR; o < 32; o <<= 1) {
#pragma unroll
                for (int q = 0; q < 8; ++q) k[q] += __shfl_xor_sync(0xffffffffu, k[q], o);
            }
            // the share every arg-max row receives (bf16, as stored), and this cell's contribution to the column sums
            float g[8], a3[8], zf[8], sh[8];
            Raw16<T>::unpack(gv, g);
            Raw16<T>::unpack(av, a3);
            Raw16<T>::unpack(ze, zf);
#pragma unroll
            for (int q = 0; q < 8; ++q) sh[q] = (a3[q] > 0.f && k[q] > 0) ? g[q] / (float)k[q] : 0.f;
            const uint4 share = Raw16<T>::pack(sh);                         // rounded to bf16 as stored ...
            Raw16<T>::unpack(share, sh);                                    // ... and the sums use exactly those values
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                if (sub == 0) {
                    const float tot = sh[q] * (float)k[q];
                    red[warp][0][q * LPR + rl] += tot;
                    red[warp][1][q * LPR + rl] = fmaf(tot, zf[q], red[warp][1][q * LPR + rl]);
                }
            }
            // pass B: write the gradient rows
#pragma unroll
            for (int u = 0; u < U; ++u) ida[u] = idn[u];
            for (int j0 = 0; j0 < cur.n; j0 += STEP) {
                uint4 raw[U];
                int id[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    id[u] = ida[u];
                    if (id[u] >= 0) raw[u] = __ldg(reinterpret_cast<const uint4 *>(zb + (int64_t)id[u] * C));
                }
#pragma unroll
                for (int u = 0; u < U; ++u) { const int j = j0 + STEP + u * RPL + sub; ida[u] = j < cur.n ? __ldg(ord + j) : -1; }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    if (id[u] >= 0) {
                        uint4 o;
                        o.x = share.x & bf16x2_eq_mask(raw[u].x, ze.x);
                        o.y = share.y & bf16x2_eq_mask(raw[u].y, ze.y);
                        o.z = share.z & bf16x2_eq_mask(raw[u].z, ze.z);
                        o.w = share.w & bf16x2_eq_mask(raw[u].w, ze.w);
                        *reinterpret_cast<uint4 *>(db + (int64_t)id[u] * C) = o;
                    }
                }
            }
        }
        cid += nwarps;
        cur = nxt;
        nxt = nxt2;
#pragma unroll
        for (int u = 0; u < U; ++u) idn[u] = idn2[u];
    }
    // rows of points outside the grid (skipped when the consumer masks them by cell id itself)
    if (cell != nullptr) {
        const int64_t g0 = (((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / LPR), gn = ((int64_t)gridDim.x * blockDim.x) / LPR;
        const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
        for (int64_t i = g0; i < total; i += gn)
            if (__ldg(cell + i) < 0) *reinterpret_cast<uint4 *>(dy + i * C + ch) = zero;
    }
    // S0 / S1: the 8 warps through shared memory, then fp64 atomics (only sub 0 accumulated)
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) {
        const int c = i % C, idx = (c & 7) * LPR + (c >> 3);
        float v = 0.f;
        for (int w = 0; w < 8; ++w) v += red[w][i / C][idx];
        atomicAdd(sums + i, (double)v);
    }
}

// ----------------------------------------------------------------------------- cell-sorted rows (f2): shares + tie bits, no dy rows
// With the rows of a cell contiguous (kdf_bev_build_sorted) the gradient rows need not exist: per cell this writes the
// share every row at the extreme receives (bf16 [cells, C]) and per row ONE BYTE per 8 channels saying which of them sit at
// the extreme (u8 [rows, C/8]); the layer-3 backward forms dy = bit ? share[cell] : 0 in its prologue (a 128-row tile meets
// ~5 cells, so the share rows are L1 hits).  C*s/8 + ... bytes per point instead of a C*s row written and read back.
template <int LPR, int MINB>
__global__ void __launch_bounds__(256, MINB)
bev_bwd_share_kernel(const __nv_bfloat16 *__restrict__ grad_grid, const __nv_bfloat16 *__restrict__ z,
                     const __nv_bfloat16 *__restrict__ grid, const __nv_bfloat16 *__restrict__ grid_z,
                     const int32_t *__restrict__ offsets, __nv_bfloat16 *__restrict__ share_out, uint8_t *__restrict__ bits,
                     double *__restrict__ sums /* [2][C] */, int64_t n_cells, int64_t N, int HW) {
    using T = __nv_bfloat16;
    constexpr int C = LPR * 8, RPL = 32 / LPR, U = 8, STEP = RPL * U;   // 16 rows of a C = 128 cell in flight per warp
    __shared__ float red[8][2][LPR * 8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, sub = lane / LPR, ch = (lane % LPR) * 8;
    const int rl = lane % LPR;
    if (sub == 0) {
#pragma unroll
        for (int q = 0; q < 8; ++q) { red[warp][0][q * LPR + rl] = 0.f; red[warp][1][q * LPR + rl] = 0.f; }
    }
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    int64_t cid = warp0;
    CellMeta cur = cell_meta(offsets, cid, n_cells, N, HW);
    CellMeta nxt = cell_meta(offsets, cid + nwarps, n_cells, N, HW);
    while (cid < n_cells) {
        const CellMeta nxt2 = cell_meta(offsets, cid + 2 * nwarps, n_cells, N, HW);
        if (cur.n > 0) {
            const T *zb = z + (cur.rowbase + cur.beg) * C + ch;
            uint8_t *bb = bits + (cur.rowbase + cur.beg) * LPR + rl;
            const uint4 ze = __ldg(reinterpret_cast<const uint4 *>(grid_z + cid * C + ch));
            const uint4 gv = __ldg(reinterpret_cast<const uint4 *>(grad_grid + cid * C + ch));
            const uint4 av = __ldg(reinterpret_cast<const uint4 *>(grid + cid * C + ch));
            int k[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) k[q] = 0;
            for (int j0 = 0; j0 < cur.n; j0 += STEP) {
                uint4 raw[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int j = j0 + u * RPL + sub;
                    if (j < cur.n) raw[u] = ldg_stream_u4(reinterpret_cast<const uint4 *>(zb + (int64_t)j * C));
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int j = j0 + u * RPL + sub;
                    if (j < cur.n) {
                        const uint32_t e0 = bf16x2_eq_mask(raw[u].x, ze.x), e1 = bf16x2_eq_mask(raw[u].y, ze.y);
                        const uint32_t e2 = bf16x2_eq_mask(raw[u].z, ze.z), e3 = bf16x2_eq_mask(raw[u].w, ze.w);
                        k[0] += e0 & 1; k[1] += e0 >> 31; k[2] += e1 & 1; k[3] += e1 >> 31;
                        k[4] += e2 & 1; k[5] += e2 >> 31; k[6] += e3 & 1; k[7] += e3 >> 31;
                        // the row's tie bits for these 8 channels (bit q = channel ch + q)
                        bb[(int64_t)j * LPR] = (uint8_t)((e0 & 1) | ((e0 >> 31) << 1) | ((e1 & 1) << 2) | ((e1 >> 31) << 3) |
                                                         ((e2 & 1) << 4) | ((e2 >> 31) << 5) | ((e3 & 1) << 6) | ((e3 >> 31) << 7));
                    }
                }
            }
#pragma unroll
            for (int o = LPR; o < 32; o <<= 1) {
#pragma unroll
                for (int q = 0; q < 8; ++q) k[q] += __shfl_xor_sync(0xffffffffu, k[q], o);
            }
            if (sub == 0) {
                float g[8], a3[8], zf[8], sh[8];
                Raw16<T>::unpack(gv, g);
                Raw16<T>::unpack(av, a3);
                Raw16<T>::unpack(ze, zf);
#pragma unroll
                for (int q = 0; q < 8; ++q) sh[q] = (a3[q] > 0.f && k[q] > 0) ? g[q] / (float)k[q] : 0.f;
                const uint4 share = Raw16<T>::pack(sh);                     // rounded to bf16 as the rows would have stored it ...
                *reinterpret_cast<uint4 *>(share_out + cid * C + ch) = share;
                Raw16<T>::unpack(share, sh);                                // ... and the sums use exactly those values
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const float tot = sh[q] * (float)k[q];
                    red[warp][0][q * LPR + rl] += tot;
                    red[warp][1][q * LPR + rl] = fmaf(tot, zf[q], red[warp][1][q * LPR + rl]);
                }
            }
        }
        cid += nwarps;
        cur = nxt;
        nxt = nxt2;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) {
        const int c = i % C, idx = (c & 7) * LPR + (c >> 3);
        float v = 0.f;
        for (int w = 0; w < 8; ++w) v += red[w][i / C][idx];
        atomicAdd(sums + i, (double)v);
    }
}

// ----------------------------------------------------------------------------- the reference op without tie counts in the forward
// `scatter_reduce_(amax, include_self=False)` (lidar_encoder.py:85-96) for rows of LPR 16-byte lanes, built like the
// two kernels above: the forward is a pure maximum (one `max.bf16x2` per two bf16 channels / one FMNMX per fp32
// channel -- the tie counts ATen's backward needs are NOT formed here, which is what made `bev_reduce_wide_kernel`
// ALU-bound), and the backward counts the rows at the maximum in a first sweep and writes the shares in a second
// one that is served by L1/L2.  Persistent warps with the one-cell-ahead prefetch of offsets and first point ids.
template <typename T> struct MaxAcc;
template <> struct MaxAcc<__nv_bfloat16> {
    uint32_t m[4];
    __device__ __forceinline__ void init() { m[0] = m[1] = m[2] = m[3] = 0xFF80FF80u; }
    __device__ __forceinline__ void take(const uint4 &r) {
        m[0] = bf16x2_max(m[0], r.x); m[1] = bf16x2_max(m[1], r.y); m[2] = bf16x2_max(m[2], r.z); m[3] = bf16x2_max(m[3], r.w);
    }
    __device__ __forceinline__ void merge(int o) {
#pragma unroll
        for (int q = 0; q < 4; ++q) m[q] = bf16x2_max(m[q], __shfl_xor_sync(0xffffffffu, m[q], o));
    }
    __device__ __forceinline__ uint4 raw() const { return make_uint4(m[0], m[1], m[2], m[3]); }
};
template <> struct MaxAcc<float> {
    float m[4];
    __device__ __forceinline__ void init() { m[0] = m[1] = m[2] = m[3] = -INFINITY; }
    __device__ __forceinline__ void take(const uint4 &r) {
        m[0] = fmaxf(m[0], __uint_as_float(r.x)); m[1] = fmaxf(m[1], __uint_as_float(r.y));
        m[2] = fmaxf(m[2], __uint_as_float(r.z)); m[3] = fmaxf(m[3], __uint_as_float(r.w));
    }
    __device__ __forceinline__ void merge(int o) {
#pragma unroll
        for (int q = 0; q < 4; ++q) m[q] = fmaxf(m[q], __shfl_xor_sync(0xffffffffu, m[q], o));
    }
    __device__ __forceinline__ uint4 raw() const {
        return make_uint4(__float_as_uint(m[0]), __float_as_uint(m[1]), __float_as_uint(m[2]), __float_as_uint(m[3]));
    }
};

template <typename T, int LPR>
__global__ void __launch_bounds__(256, 4)
bev_reduce_max_kernel(const T *__restrict__ feats, const int32_t *__restrict__ order, const int32_t *__restrict__ offsets,
                      T *__restrict__ grid, int64_t n_cells, int64_t N, int HW) {
    constexpr int VEC = Raw16<T>::VEC, C = LPR * VEC, RPL = 32 / LPR, U = 4, STEP = RPL * U;
    const int lane = threadIdx.x & 31, sub = lane / LPR, ch = (lane % LPR) * VEC;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;

    int64_t cid = warp0;
    CellMeta cur = cell_meta(offsets, cid, n_cells, N, HW);
    int idn[U];
#pragma unroll
    for (int u = 0; u < U; ++u) { const int j = u * RPL + sub; idn[u] = j < cur.n ? __ldg(order + cur.rowbase + cur.beg + j) : -1; }
    CellMeta nxt = cell_meta(offsets, cid + nwarps, n_cells, N, HW);
    while (cid < n_cells) {
        int idn2[U];
#pragma unroll
        for (int u = 0; u < U; ++u) { const int j = u * RPL + sub; idn2[u] = j < nxt.n ? __ldg(order + nxt.rowbase + nxt.beg + j) : -1; }
        const CellMeta nxt2 = cell_meta(offsets, cid + 2 * nwarps, n_cells, N, HW);

        const int32_t *ord = order + cur.rowbase + cur.beg;
        const T *fb = feats + cur.rowbase * C + ch;
        MaxAcc<T> acc;
        acc.init();
        for (int j0 = 0; j0 < cur.n; j0 += STEP) {
            uint4 raw[U];
            int id[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                id[u] = idn[u];
                if (id[u] >= 0) raw[u] = ldg_stream_u4(reinterpret_cast<const uint4 *>(fb + (int64_t)id[u] * C));
            }
#pragma unroll
            for (int u = 0; u < U; ++u) { const int j = j0 + STEP + u * RPL + sub; idn[u] = j < cur.n ? __ldg(ord + j) : -1; }
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (id[u] >= 0) acc.take(raw[u]);
        }
#pragma unroll
        for (int o = LPR; o < 32; o <<= 1) acc.merge(o);
        if (sub == 0)
            *reinterpret_cast<uint4 *>(grid + cid * C + ch) = cur.n > 0 ? acc.raw() : make_uint4(0u, 0u, 0u, 0u);   // empty cells are 0
        cid += nwarps;
        cur = nxt;
        nxt = nxt2;
#pragma unroll
        for (int u = 0; u < U; ++u) idn[u] = idn2[u];
    }
}

template <typename T> struct EqRows;
template <> struct EqRows<__nv_bfloat16> {
    static __device__ __forceinline__ void count(const uint4 &r, const uint4 &mx, int *k) {
        const uint32_t e0 = bf16x2_eq_mask(r.x, mx.x), e1 = bf16x2_eq_mask(r.y, mx.y);
        const uint32_t e2 = bf16x2_eq_mask(r.z, mx.z), e3 = bf16x2_eq_mask(r.w, mx.w);
        k[0] += e0 & 1; k[1] += e0 >> 31; k[2] += e1 & 1; k[3] += e1 >> 31;
        k[4] += e2 & 1; k[5] += e2 >> 31; k[6] += e3 & 1; k[7] += e3 >> 31;
    }
    static __device__ __forceinline__ uint4 select(const uint4 &r, const uint4 &mx, const uint4 &share) {
        return make_uint4(share.x & bf16x2_eq_mask(r.x, mx.x), share.y & bf16x2_eq_mask(r.y, mx.y),
                          share.z & bf16x2_eq_mask(r.z, mx.z), share.w & bf16x2_eq_mask(r.w, mx.w));
    }
};
template <> struct EqRows<float> {
    static __device__ __forceinline__ void count(const uint4 &r, const uint4 &mx, int *k) {
        k[0] += __uint_as_float(r.x) == __uint_as_float(mx.x); k[1] += __uint_as_float(r.y) == __uint_as_float(mx.y);
        k[2] += __uint_as_float(r.z) == __uint_as_float(mx.z); k[3] += __uint_as_float(r.w) == __uint_as_float(mx.w);
    }
    static __device__ __forceinline__ uint4 select(const uint4 &r, const uint4 &mx, const uint4 &share) {
        return make_uint4(__uint_as_float(r.x) == __uint_as_float(mx.x) ? share.x : 0u, __uint_as_float(r.y) == __uint_as_float(mx.y) ? share.y : 0u,
                          __uint_as_float(r.z) == __uint_as_float(mx.z) ? share.z : 0u, __uint_as_float(r.w) == __uint_as_float(mx.w) ? share.w : 0u);
    }
};

template <typename T, int LPR, int MINB>
__global__ void __launch_bounds__(256, MINB)
bev_bwd_max_kernel(const T *__restrict__ grad_grid, const T *__restrict__ feats, const T *__restrict__ grid,
                   const int32_t *__restrict__ order, const int32_t *__restrict__ offsets, const int32_t *__restrict__ cell,
                   T *__restrict__ grad_feats, int64_t n_cells, int64_t N, int HW, int64_t total) {
    constexpr int VEC = Raw16<T>::VEC, C = LPR * VEC, RPL = 32 / LPR, U = 4, STEP = RPL * U;
    const int lane = threadIdx.x & 31, sub = lane / LPR, ch = (lane % LPR) * VEC;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;

    int64_t cid = warp0;
    CellMeta cur = cell_meta(offsets, cid, n_cells, N, HW);
    int idn[U];
#pragma unroll
    for (int u = 0; u < U; ++u) { const int j = u * RPL + sub; idn[u] = j < cur.n ? __ldg(order + cur.rowbase + cur.beg + j) : -1; }
    CellMeta nxt = cell_meta(offsets, cid + nwarps, n_cells, N, HW);
    while (cid < n_cells) {
        int idn2[U];
#pragma unroll
        for (int u = 0; u < U; ++u) { const int j = u * RPL + sub; idn2[u] = j < nxt.n ? __ldg(order + nxt.rowbase + nxt.beg + j) : -1; }
        const CellMeta nxt2 = cell_meta(offsets, cid + 2 * nwarps, n_cells, N, HW);
        if (cur.n > 0) {
            const int32_t *ord = order + cur.rowbase + cur.beg;
            const T *fb = feats + cur.rowbase * C + ch;
            T *db = grad_feats + cur.rowbase * C + ch;
            const uint4 mx = __ldg(reinterpret_cast<const uint4 *>(grid + cid * C + ch));
            const uint4 gv = __ldg(reinterpret_cast<const uint4 *>(grad_grid + cid * C + ch));
            // sweep A: rows at the maximum, per channel
            int k[VEC];
#pragma unroll
            for (int q = 0; q < VEC; ++q) k[q] = 0;
            int ida[U];
#pragma unroll
            for (int u = 0; u < U; ++u) ida[u] = idn[u];
            for (int j0 = 0; j0 < cur.n; j0 += STEP) {
                uint4 raw[U];
                int id[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    id[u] = ida[u];
                    if (id[u] >= 0) raw[u] = __ldg(reinterpret_cast<const uint4 *>(fb + (int64_t)id[u] * C));     // L1-allocating: sweep B re-reads
                }
#pragma unroll
                for (int u = 0; u < U; ++u) { const int j = j0 + STEP + u * RPL + sub; ida[u] = j < cur.n ? __ldg(ord + j) : -1; }
#pragma unroll
                for (int u = 0; u < U; ++u)
                    if (id[u] >= 0) EqRows<T>::count(raw[u], mx, k);
            }
#pragma unroll
            for (int o = LPR; o < 32; o <<= 1) {
#pragma unroll
                for (int q = 0; q < VEC; ++q) k[q] += __shfl_xor_sync(0xffffffffu, k[q], o);
            }
            // ATen: N_to_distribute = (self == result) + #(src == result), self being the zero-initialised output
            float g[VEC], m[VEC];
            Raw16<T>::unpack(gv, g);
            Raw16<T>::unpack(mx, m);
#pragma unroll
            for (int q = 0; q < VEC; ++q) g[q] = g[q] / (float)(k[q] + (m[q] == 0.f ? 1 : 0));
            const uint4 share = Raw16<T>::pack(g);
            // sweep B: the gradient rows
#pragma unroll
            for (int u = 0; u < U; ++u) ida[u] = idn[u];
            for (int j0 = 0; j0 < cur.n; j0 += STEP) {
                uint4 raw[U];
                int id[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    id[u] = ida[u];
                    if (id[u] >= 0) raw[u] = __ldg(reinterpret_cast<const uint4 *>(fb + (int64_t)id[u] * C));
                }
#pragma unroll
                for (int u = 0; u < U; ++u) { const int j = j0 + STEP + u * RPL + sub; ida[u] = j < cur.n ? __ldg(ord + j) : -1; }
#pragma unroll
                for (int u = 0; u < U; ++u)
                    if (id[u] >= 0) *reinterpret_cast<uint4 *>(db + (int64_t)id[u] * C) = EqRows<T>::select(raw[u], mx, share);
            }
        }
        cid += nwarps;
        cur = nxt;
        nxt = nxt2;
#pragma unroll
        for (int u = 0; u < U; ++u) idn[u] = idn2[u];
    }
    // rows of points outside the grid
    const int64_t g0 = (((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / LPR), gn = ((int64_t)gridDim.x * blockDim.x) / LPR;
    const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
    for (int64_t i = g0; i < total; i += gn)
        if (__ldg(cell + i) < 0) *reinterpret_cast<uint4 *>(grad_feats + i * C + ch) = zero;
}

// ----------------------------------------------------------------------------- host side
static int check_geom(int B, int64_t N, int H, int W, float xspan, float yspan) {
    KDF_CHECK_ARG(B >= 0 && N >= 0, "bev: negative B or N");
    KDF_CHECK_ARG(H > 0 && W > 0 && (int64_t)H * W < (1 << 30), "bev: bad grid %dx%d", H, W);
    KDF_CHECK_ARG(N < (int64_t)INT32_MAX, "bev: N must fit int32");
    (void)xspan; (void)yspan;
    return KDF_OK;
}

static int grid_for(int64_t work_items, int per_block, int max_waves = 16) {
    int64_t blocks = (work_items + per_block - 1) / per_block;
    const int64_t cap = (int64_t)sm_count() * 8 * max_waves;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

static int launch_index(const float *points, int B, int64_t N, int stride, const BevGeom &g,
                        int32_t *cell, int32_t *rank, int32_t *count, cudaStream_t st) {
    const int64_t total = (int64_t)B * N;
    KDF_CUDA(cudaMemsetAsync(count, 0, sizeof(int32_t) * (size_t)B * g.H * g.W, st));
    if (total == 0) return KDF_OK;
    const bool vec4 = (stride == 4) && ((reinterpret_cast<uintptr_t>(points) & 15) == 0);
    const size_t hist_bytes = sizeof(int) * (size_t)g.H * g.W;
    if (vec4 && rank == nullptr && hist_bytes <= 96 * 1024 && B <= 65535) {
        // slices of >= 8192 points, about four 512-thread CTAs per SM over the whole batch
        constexpr int per_sm = 4;
        int64_t slice = (total + (int64_t)sm_count() * per_sm - 1) / ((int64_t)sm_count() * per_sm);
        if (slice < 8192) slice = 8192;
        slice = (slice + 2047) / 2048 * 2048;
        const int64_t nslice = (N + slice - 1) / slice;
        if (nslice <= 65535) {
            if (hist_bytes > 48 * 1024) {
                KDF_CUDA(cudaFuncSetAttribute(bev_index_hist_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hist_bytes));
                KDF_CUDA(cudaFuncSetAttribute(bev_index_hist_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hist_bytes));
            }
            if (g.range_view)
                bev_index_hist_kernel<true><<<dim3((unsigned)nslice, (unsigned)B), 512, hist_bytes, st>>>(
                    reinterpret_cast<const float4 *>(points), N, slice, g, cell, count);
            else
                bev_index_hist_kernel<false><<<dim3((unsigned)nslice, (unsigned)B), 512, hist_bytes, st>>>(
                    reinterpret_cast<const float4 *>(points), N, slice, g, cell, count);
            KDF_LAUNCH_CHECK();
            return KDF_OK;
        }
    }
    const int blocks = grid_for(total, 256, 4);
    if (g.range_view) {
        if (vec4) bev_index_kernel<true, true><<<blocks, 256, 0, st>>>(points, total, N, stride, g, cell, rank, count);
        else      bev_index_kernel<false, true><<<blocks, 256, 0, st>>>(points, total, N, stride, g, cell, rank, count);
    } else {
        if (vec4) bev_index_kernel<true, false><<<blocks, 256, 0, st>>>(points, total, N, stride, g, cell, rank, count);
        else      bev_index_kernel<false, false><<<blocks, 256, 0, st>>>(points, total, N, stride, g, cell, rank, count);
    }
    KDF_LAUNCH_CHECK();
    return KDF_OK;
}

static int wide_lpr(int dtype, int C) {
    const int lpr = C / (dtype == KDF_F32 ? 4 : 8);
    const int rem = C % (dtype == KDF_F32 ? 4 : 8);
    return (rem == 0 && (lpr == 8 || lpr == 16 || lpr == 32)) ? lpr : 0;
}

static int launch_reduce(const void *feats, int dtype, const int32_t *order, const int32_t *offsets,
                         int B, int64_t N, int C, int HW, int reduce, void *grid, int32_t *ties, cudaStream_t st) {
    const int64_t n_cells = (int64_t)B * HW;
    const int blocks = grid_for(n_cells * 32, 256, 64);
    const int lpr = wide_lpr(dtype, C);
    if (lpr && reduce == KDF_REDUCE_MAX && ties == nullptr) {        // pure maximum: persistent warps, several cells each
        int64_t pb = (n_cells + 7) / 8;
        if (pb > (int64_t)sm_count() * 8) pb = (int64_t)sm_count() * 8;
#define KDF_RM(T, L) bev_reduce_max_kernel<T, L><<<(int)pb, 256, 0, st>>>(reinterpret_cast<const T *>(feats), order, offsets, \
                                                                         reinterpret_cast<T *>(grid), n_cells, N, HW)
        if (dtype == KDF_F32) { if (lpr == 8) KDF_RM(float, 8); else if (lpr == 16) KDF_RM(float, 16); else KDF_RM(float, 32); }
        else { if (lpr == 8) KDF_RM(__nv_bfloat16, 8); else if (lpr == 16) KDF_RM(__nv_bfloat16, 16); else KDF_RM(__nv_bfloat16, 32); }
#undef KDF_RM
        KDF_LAUNCH_CHECK();
        return KDF_OK;
    }
#define KDF_RW(T, L, R)                                                                                 \
    bev_reduce_wide_kernel<T, L, R><<<blocks, 256, 0, st>>>(reinterpret_cast<const T *>(feats), order, offsets, \
                                                            reinterpret_cast<T *>(grid), ties, n_cells, N, HW)
#define KDF_RW_L(T, R)                                                          \
    do {                                                                        \
        if (lpr == 8) KDF_RW(T, 8, R); else if (lpr == 16) KDF_RW(T, 16, R); else KDF_RW(T, 32, R); \
    } while (0)
#define KDF_REDUCE_LAUNCH(T, R)                                                                    \
    bev_reduce_kernel<T, R><<<blocks, 256, 0, st>>>(reinterpret_cast<const T *>(feats), order, offsets, \
                                                   reinterpret_cast<T *>(grid), ties, n_cells, N, C, HW)
    if (lpr) {
        if (dtype == KDF_F32) { if (reduce == KDF_REDUCE_MAX) KDF_RW_L(float, KDF_REDUCE_MAX); else KDF_RW_L(float, KDF_REDUCE_MEAN); }
        else { if (reduce == KDF_REDUCE_MAX) KDF_RW_L(__nv_bfloat16, KDF_REDUCE_MAX); else KDF_RW_L(__nv_bfloat16, KDF_REDUCE_MEAN); }
    } else if (dtype == KDF_F32) {
        if (reduce == KDF_REDUCE_MAX) KDF_REDUCE_LAUNCH(float, KDF_REDUCE_MAX);
        else                          KDF_REDUCE_LAUNCH(float, KDF_REDUCE_MEAN);
    } else {
        if (reduce == KDF_REDUCE_MAX) KDF_REDUCE_LAUNCH(__nv_bfloat16, KDF_REDUCE_MAX);
        else                          KDF_REDUCE_LAUNCH(__nv_bfloat16, KDF_REDUCE_MEAN);
    }
#undef KDF_REDUCE_LAUNCH
#undef KDF_RW_L
#undef KDF_RW
    KDF_LAUNCH_CHECK();
    return KDF_OK;
}

}  // namespace kdf

using namespace kdf;

extern "C" {

int kdf_bev_index(const float *points, int B, int64_t N, int point_stride,
                  float x0, float xspan, float y0, float yspan, int H, int W,
                  int32_t *cell, int32_t *rank, int32_t *count, void *stream) {
    if (int e = check_geom(B, N, H, W, xspan, yspan)) return e;
    KDF_CHECK_ARG(point_stride >= 2, "bev: point_stride must be >= 2");
    KDF_CHECK_ARG(((points && cell) || (int64_t)B * N == 0) && (count || B == 0), "bev: null pointer");
    BevGeom g{x0, xspan, y0, yspan, (float)(W - 1), (float)(H - 1), H, W};
    return launch_index(points, B, N, point_stride, g, cell, rank, count, as_stream(stream));
}

static int range_geom(BevGeom &g, int B, int64_t N, int H, int W, float fov_up, float fov_down, int point_stride) {
    if (int e = check_geom(B, N, H, W, 1.f, 1.f)) return e;
    KDF_CHECK_ARG(point_stride >= 3, "range view: points need (x, y, z): point_stride must be >= 3");
    KDF_CHECK_ARG(fov_up > fov_down && fov_up <= 1.5708f && fov_down >= -1.5708f, "range view: bad vertical field of view [%g, %g] rad",
                  (double)fov_down, (double)fov_up);
    g = BevGeom{0.f, 1.f, 0.f, 1.f, 0.f, 0.f, H, W, 1, fov_down, fov_up - fov_down};
    return KDF_OK;
}

int kdf_range_index(const float *points, int B, int64_t N, int point_stride, float fov_up, float fov_down, int H, int W,
                    int32_t *cell, int32_t *count, void *stream) {
    BevGeom g;
    if (int e = range_geom(g, B, N, H, W, fov_up, fov_down, point_stride)) return e;
    KDF_CHECK_ARG(((points && cell) || (int64_t)B * N == 0) && (count || B == 0), "range_index: null pointer");
    return launch_index(points, B, N, point_stride, g, cell, nullptr, count, as_stream(stream));
}

size_t kdf_bev_workspace_bytes(int B, int64_t N, int H, int W) {
    const size_t bn = align_up(sizeof(int32_t) * (size_t)B * (size_t)N, 256);
    const size_t off = align_up(sizeof(int32_t) * (size_t)B * ((size_t)H * W + 1), 256);
    const size_t nchunk = (size_t)((N + BEV_CHUNK - 1) / BEV_CHUNK);
    const size_t hist = align_up(sizeof(int32_t) * (size_t)B * nchunk * (size_t)H * W, 256);    // per-chunk histograms
    return 2 * bn + off + hist + 256;
}

// index -> scan -> fill: cell ids, occupancy and the cell ordering (counting sort) of every frame
static int build_order(const float *points, int point_stride, int B, int64_t N, const BevGeom &g,
                       int32_t *count, int32_t *cell, int32_t *order, int32_t *offsets, int32_t *rank, cudaStream_t st,
                       float *sorted_points = nullptr, int32_t *cell_sorted = nullptr) {
    const int HW = g.H * g.W;
    const int64_t total = (int64_t)B * N;
    const int64_t nchunk = (N + BEV_CHUNK - 1) / BEV_CHUNK;
    if (point_stride == 4 && (reinterpret_cast<uintptr_t>(points) & 15) == 0 && HW * sizeof(int) <= 96 * 1024 &&
        N > 0 && nchunk <= 65535 && B <= 65535) {
        // the histograms live behind rank / (unused second [B,N] block) / offsets in the workspace
        const size_t bn = align_up(sizeof(int32_t) * (size_t)B * (size_t)N, 256);
        const size_t off = align_up(sizeof(int32_t) * (size_t)B * ((size_t)HW + 1), 256);
        int32_t *chist = reinterpret_cast<int32_t *>(reinterpret_cast<uint8_t *>(rank) + 2 * bn + off);
        // cell-sorted points: the per-chunk counts of points outside the grid sit in the (otherwise unused) second [B,N] block
        int32_t *cout = sorted_points ? reinterpret_cast<int32_t *>(reinterpret_cast<uint8_t *>(rank) + bn) : nullptr;
        KDF_CHECK_ARG(!sorted_points || (size_t)B * nchunk * sizeof(int32_t) <= bn, "bev_build_sorted: frames too short for the workspace layout");
        const size_t smem = sizeof(int) * (size_t)HW;
        if (smem > 48 * 1024) {
            KDF_CUDA(cudaFuncSetAttribute(bev_chunk_permute_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            KDF_CUDA(cudaFuncSetAttribute(bev_chunk_count_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            KDF_CUDA(cudaFuncSetAttribute(bev_chunk_count_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            KDF_CUDA(cudaFuncSetAttribute(bev_chunk_fill_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        }
        const dim3 grid((unsigned)nchunk, (unsigned)B);
        if (g.range_view)
            bev_chunk_count_kernel<true><<<grid, 256, smem, st>>>(reinterpret_cast<const float4 *>(points), N, g, cell, rank, chist, (int)nchunk, cout);
        else
            bev_chunk_count_kernel<false><<<grid, 256, smem, st>>>(reinterpret_cast<const float4 *>(points), N, g, cell, rank, chist, (int)nchunk, cout);
        KDF_LAUNCH_CHECK();
        bev_chunk_scan_kernel<<<B, 1024, 0, st>>>(chist, count, offsets, HW, (int)nchunk, cout);
        KDF_LAUNCH_CHECK();
        if (sorted_points)
            bev_chunk_permute_kernel<<<grid, 256, smem, st>>>(reinterpret_cast<const float4 *>(points), cell, rank, chist, cout, offsets,
                                                              reinterpret_cast<float4 *>(sorted_points), cell_sorted, order, N, HW, (int)nchunk);
        else
            bev_chunk_fill_kernel<<<grid, 256, smem, st>>>(cell, rank, chist, offsets, order, N, HW, (int)nchunk);
        KDF_LAUNCH_CHECK();
        return KDF_OK;
    }
    KDF_CHECK_ARG(!sorted_points, "bev_build_sorted: needs 16-byte aligned (x, y, z, i) points, H*W <= 24576 cells and N > 0");
    if (int e = launch_index(points, B, N, point_stride, g, cell, rank, count, st)) return e;
    bev_scan_kernel<<<B, 1024, 0, st>>>(count, offsets, HW);
    KDF_LAUNCH_CHECK();
    if (total > 0) {
        bev_fill_kernel<<<grid_for(total, 256, 4), 256, 0, st>>>(cell, rank, offsets, order, total, N, HW);
        KDF_LAUNCH_CHECK();
    }
    return KDF_OK;
}

int kdf_bev_build_order(const float *points, int point_stride, int B, int64_t N,
                        float x0, float xspan, float y0, float yspan, int H, int W,
                        int32_t *count, int32_t *cell, int32_t *order, int32_t *offsets,
                        void *workspace, size_t workspace_bytes, void *stream) {
    if (int e = check_geom(B, N, H, W, xspan, yspan)) return e;
    KDF_CHECK_ARG(point_stride >= 2, "bev: point_stride must be >= 2");
    if (B == 0) return KDF_OK;
    KDF_CHECK_ARG(((points && cell && order) || N == 0) && count && offsets && workspace, "bev_build_order: null pointer");
    KDF_CHECK_ARG(workspace_bytes >= kdf_bev_workspace_bytes(B, N, H, W), "bev_build_order: workspace too small");
    BevGeom g{x0, xspan, y0, yspan, (float)(W - 1), (float)(H - 1), H, W};
    return build_order(points, point_stride, B, N, g, count, cell, order, offsets, reinterpret_cast<int32_t *>(workspace),
                       as_stream(stream));
}

// The same, with the points themselves written in cell order (f2): sorted_points f32 [B,N,4], cell_sorted i32 [B,N]
// (global cell id b*H*W + cell of every sorted row, -1 for the rows of points outside, which close each frame);
// order (nullable) = the permutation, sorted row -> point id.
int kdf_bev_build_sorted(const float *points, int B, int64_t N, float x0, float xspan, float y0, float yspan, int H, int W,
                         int32_t *count, int32_t *cell, int32_t *offsets, float *sorted_points, int32_t *cell_sorted,
                         int32_t *order, void *workspace, size_t workspace_bytes, void *stream) {
    if (int e = check_geom(B, N, H, W, xspan, yspan)) return e;
    if (B == 0) return KDF_OK;
    KDF_CHECK_ARG(((points && cell && sorted_points && cell_sorted) || N == 0) && count && offsets && workspace, "bev_build_sorted: null pointer");
    KDF_CHECK_ARG(workspace_bytes >= kdf_bev_workspace_bytes(B, N, H, W), "bev_build_sorted: workspace too small");
    KDF_CHECK_ARG((int64_t)B * H * W < (1ll << 31), "bev_build_sorted: too many cells for 32-bit global cell ids");
    KDF_CHECK_ARG((reinterpret_cast<uintptr_t>(sorted_points) & 15) == 0, "bev_build_sorted: sorted_points must be 16-byte aligned");
    BevGeom g{x0, xspan, y0, yspan, (float)(W - 1), (float)(H - 1), H, W};
    return build_order(points, 4, B, N, g, count, cell, order, offsets, reinterpret_cast<int32_t *>(workspace), as_stream(stream),
                       sorted_points, cell_sorted);
}

static int project_fwd_impl(const float *points, int point_stride, const void *feats, int dtype, int B, int64_t N, int C,
                            const BevGeom &g, int reduce, void *grid, int32_t *count, int32_t *cell, int32_t *ties,
                            int32_t *order, int32_t *offsets, void *workspace, size_t workspace_bytes, void *stream) {
    const int H = g.H, W = g.W;
    KDF_CHECK_ARG(C > 0 && C % 4 == 0, "bev: C=%d must be a positive multiple of 4", C);
    KDF_CHECK_ARG(dtype == KDF_F32 || dtype == KDF_BF16, "bev: bad dtype %d", dtype);
    KDF_CHECK_ARG(reduce == KDF_REDUCE_MAX || reduce == KDF_REDUCE_MEAN, "bev: bad reduce %d", reduce);
    KDF_CHECK_ARG(((points && feats && cell) || (int64_t)B * N == 0) && ((grid && count && workspace) || B == 0),
                  "bev: null pointer");
    KDF_CHECK_ARG(workspace_bytes >= kdf_bev_workspace_bytes(B, N, H, W), "bev: workspace too small");
    KDF_CHECK_ARG((reinterpret_cast<uintptr_t>(feats) & 15) == 0 && (reinterpret_cast<uintptr_t>(grid) & 15) == 0,
                  "bev: feats/grid must be 16-byte aligned");
    cudaStream_t st = as_stream(stream);
    const int HW = H * W;
    const int64_t total = (int64_t)B * N;
    char *ws = reinterpret_cast<char *>(workspace);
    const size_t bn = align_up(sizeof(int32_t) * (size_t)total, 256);
    int32_t *rank = reinterpret_cast<int32_t *>(ws);
    if (!order) order = reinterpret_cast<int32_t *>(ws + bn);
    if (!offsets) offsets = reinterpret_cast<int32_t *>(ws + 2 * bn);
    if (B == 0) return KDF_OK;
    if (int e = build_order(points, point_stride, B, N, g, count, cell, order, offsets, rank, st)) return e;
    return launch_reduce(feats, dtype, order, offsets, B, N, C, HW, reduce, grid, ties, st);
}

int kdf_bev_project_fwd(const float *points, int point_stride, const void *feats, int dtype,
                        int B, int64_t N, int C,
                        float x0, float xspan, float y0, float yspan, int H, int W, int reduce,
                        void *grid, int32_t *count, int32_t *cell, int32_t *ties,
                        int32_t *order, int32_t *offsets,
                        void *workspace, size_t workspace_bytes, void *stream) {
    if (int e = check_geom(B, N, H, W, xspan, yspan)) return e;
    KDF_CHECK_ARG(point_stride >= 2, "bev: point_stride must be >= 2");
    BevGeom g{x0, xspan, y0, yspan, (float)(W - 1), (float)(H - 1), H, W};
    return project_fwd_impl(points, point_stride, feats, dtype, B, N, C, g, reduce, grid, count, cell, ties, order, offsets, workspace,
                            workspace_bytes, stream);
}

int kdf_range_project_fwd(const float *points, int point_stride, const void *feats, int dtype, int B, int64_t N, int C,
                          float fov_up, float fov_down, int H, int W, int reduce,
                          void *grid, int32_t *count, int32_t *cell, int32_t *ties, int32_t *order, int32_t *offsets,
                          void *workspace, size_t workspace_bytes, void *stream) {
    BevGeom g;
    if (int e = range_geom(g, B, N, H, W, fov_up, fov_down, point_stride)) return e;
    return project_fwd_impl(points, point_stride, feats, dtype, B, N, C, g, reduce, grid, count, cell, ties, order, offsets, workspace,
                            workspace_bytes, stream);
}

int kdf_bev_reduce(const void *feats, int dtype, const int32_t *order, const int32_t *offsets,
                   int B, int64_t N, int C, int H, int W, int reduce,
                   void *grid, int32_t *ties, void *stream) {
    KDF_CHECK_ARG(B >= 0 && N >= 0 && H > 0 && W > 0, "bev_reduce: bad sizes");
    KDF_CHECK_ARG(C > 0 && C % 4 == 0, "bev_reduce: C=%d must be a positive multiple of 4", C);
    KDF_CHECK_ARG(dtype == KDF_F32 || dtype == KDF_BF16, "bev_reduce: bad dtype %d", dtype);
    KDF_CHECK_ARG(reduce == KDF_REDUCE_MAX || reduce == KDF_REDUCE_MEAN, "bev_reduce: bad reduce %d", reduce);
    if (B == 0) return KDF_OK;
    KDF_CHECK_ARG((feats || N == 0) && order && offsets && grid, "bev_reduce: null pointer");
    return launch_reduce(feats, dtype, order, offsets, B, N, C, H * W, reduce, grid, ties, as_stream(stream));
}

int kdf_bev_project_bwd(const void *grad_grid, const void *feats, const void *grid,
                        const int32_t *ties, const int32_t *count, const int32_t *cell,
                        const int32_t *order, const int32_t *offsets,
                        int dtype, int B, int64_t N, int C, int H, int W, int reduce,
                        void *grad_feats, void *stream) {
    KDF_CHECK_ARG(B >= 0 && N >= 0 && H > 0 && W > 0, "bev_bwd: bad sizes");
    KDF_CHECK_ARG(C > 0 && C % 4 == 0, "bev_bwd: C=%d must be a positive multiple of 4", C);
    KDF_CHECK_ARG(dtype == KDF_F32 || dtype == KDF_BF16, "bev_bwd: bad dtype %d", dtype);
    const int64_t total = (int64_t)B * N;
    KDF_CHECK_ARG(reduce == KDF_REDUCE_MAX || reduce == KDF_REDUCE_MEAN, "bev_bwd: bad reduce %d", reduce);
    if (total == 0) return KDF_OK;
    KDF_CHECK_ARG(grad_grid && cell && grad_feats, "bev_bwd: null pointer");
    if (reduce == KDF_REDUCE_MAX) KDF_CHECK_ARG(feats && grid, "bev_bwd(max): feats/grid required");
    else KDF_CHECK_ARG(count, "bev_bwd(mean): count required");
    KDF_CHECK_ARG((order == nullptr) == (offsets == nullptr), "bev_bwd: order and offsets come together");
    cudaStream_t st = as_stream(stream);
    const int HW = H * W;
    const int lpr = order ? wide_lpr(dtype, C) : 0;
    if (reduce == KDF_REDUCE_MAX) KDF_CHECK_ARG(ties || lpr, "bev_bwd(max): tie counts are required for this C / without the cell ordering");
    if (lpr && reduce == KDF_REDUCE_MAX && ties == nullptr) {       // tie counts formed here (sweep A), shares written in sweep B
        const int64_t n_cells = (int64_t)B * HW;
        int64_t pb = (n_cells + 7) / 8;
        if (pb > (int64_t)sm_count() * 3) pb = (int64_t)sm_count() * 3;
#define KDF_BM(T, L) bev_bwd_max_kernel<T, L, 3><<<(int)pb, 256, 0, st>>>(reinterpret_cast<const T *>(grad_grid),          \
        reinterpret_cast<const T *>(feats), reinterpret_cast<const T *>(grid), order, offsets, cell,                     \
        reinterpret_cast<T *>(grad_feats), n_cells, N, HW, total)
        if (dtype == KDF_F32) { if (lpr == 8) KDF_BM(float, 8); else if (lpr == 16) KDF_BM(float, 16); else KDF_BM(float, 32); }
        else { if (lpr == 8) KDF_BM(__nv_bfloat16, 8); else if (lpr == 16) KDF_BM(__nv_bfloat16, 16); else KDF_BM(__nv_bfloat16, 32); }
#undef KDF_BM
        KDF_LAUNCH_CHECK();
        return KDF_OK;
    }
    if (lpr) {                       // cell-major: per-cell rows loaded once, feature rows streamed
        const int64_t n_cells = (int64_t)B * HW;
        const int blocks = grid_for(n_cells * 32, 256, 64);
#define KDF_BW(T, L, R)                                                                                      \
    bev_bwd_wide_kernel<T, L, R><<<blocks, 256, 0, st>>>(reinterpret_cast<const T *>(grad_grid),              \
        reinterpret_cast<const T *>(feats), reinterpret_cast<const T *>(grid), ties, order, offsets, cell,    \
        reinterpret_cast<T *>(grad_feats), n_cells, N, HW, total)
#define KDF_BW_L(T, R)                                                          \
    do {                                                                        \
        if (lpr == 8) KDF_BW(T, 8, R); else if (lpr == 16) KDF_BW(T, 16, R); else KDF_BW(T, 32, R); \
    } while (0)
        if (dtype == KDF_F32) { if (reduce == KDF_REDUCE_MAX) KDF_BW_L(float, KDF_REDUCE_MAX); else KDF_BW_L(float, KDF_REDUCE_MEAN); }
        else { if (reduce == KDF_REDUCE_MAX) KDF_BW_L(__nv_bfloat16, KDF_REDUCE_MAX); else KDF_BW_L(__nv_bfloat16, KDF_REDUCE_MEAN); }
#undef KDF_BW_L
#undef KDF_BW
        KDF_LAUNCH_CHECK();
        return KDF_OK;
    }
    const int blocks = grid_for((total + 3) / 4 * 32, 256, 64);
#define KDF_BWD_LAUNCH(T, R)                                                                         \
    bev_bwd_kernel<T, R><<<blocks, 256, 0, st>>>(reinterpret_cast<const T *>(grad_grid),             \
        reinterpret_cast<const T *>(feats), reinterpret_cast<const T *>(grid), ties, count, cell,     \
        reinterpret_cast<T *>(grad_feats), total, N, C, HW)
    if (dtype == KDF_F32) {
        if (reduce == KDF_REDUCE_MAX) KDF_BWD_LAUNCH(float, KDF_REDUCE_MAX);
        else                          KDF_BWD_LAUNCH(float, KDF_REDUCE_MEAN);
    } else {
        if (reduce == KDF_REDUCE_MAX) KDF_BWD_LAUNCH(__nv_bfloat16, KDF_REDUCE_MAX);
        else                          KDF_BWD_LAUNCH(__nv_bfloat16, KDF_REDUCE_MEAN);
    }
#undef KDF_BWD_LAUNCH
    KDF_LAUNCH_CHECK();
    return KDF_OK;
}

int kdf_bev_reduce_affine(const void *z_bf16, const float *scale, const float *shift,
                          const int32_t *order, const int32_t *offsets, int B, int64_t N, int C, int H, int W,
                          void *grid_bf16, void *grid_z_bf16, void *stream) {
    KDF_CHECK_ARG(B >= 0 && N >= 0 && H > 0 && W > 0, "bev_reduce_affine: bad sizes");
    KDF_CHECK_ARG(C == 64 || C == 128 || C == 256, "bev_reduce_affine: C=%d not supported (64, 128, 256)", C);
    if (B == 0) return KDF_OK;
    KDF_CHECK_ARG((z_bf16 || N == 0) && scale && shift && offsets && grid_bf16, "bev_reduce_affine: null pointer");
    const int64_t n_cells = (int64_t)B * H * W;
    int64_t blocks = (n_cells + 7) / 8;
    if (blocks > (int64_t)sm_count() * 8) blocks = (int64_t)sm_count() * 8;      // persistent warps, several cells each
    cudaStream_t st = as_stream(stream);
    const __nv_bfloat16 *zz = reinterpret_cast<const __nv_bfloat16 *>(z_bf16);
    __nv_bfloat16 *gg = reinterpret_cast<__nv_bfloat16 *>(grid_bf16), *gz = reinterpret_cast<__nv_bfloat16 *>(grid_z_bf16);
    // order == nullptr: the rows are in cell order already (kdf_bev_build_sorted), a cell is a contiguous row segment
#define KDF_RA(L)                                                                                                             \
    do {                                                                                                                      \
        if (order) bev_reduce_affine_kernel<L, true><<<(int)blocks, 256, 0, st>>>(zz, scale, shift, order, offsets, gg, gz, n_cells, N, H * W); \
        else bev_reduce_affine_kernel<L, false><<<(int)blocks, 256, 0, st>>>(zz, scale, shift, order, offsets, gg, gz, n_cells, N, H * W);      \
    } while (0)
    if (C == 64) KDF_RA(8); else if (C == 128) KDF_RA(16); else KDF_RA(32);
#undef KDF_RA
    KDF_LAUNCH_CHECK();
    return KDF_OK;
}

// Projection backward over cell-sorted rows: share bf16 [B*H*W, C] (the gradient every row at its cell's extreme receives;
// rows of empty cells are not written), bits u8 [B*N, C/8] (bit q of byte g: channel 8g+q of the row sits at the extreme;
// rows outside the grid are not written), sums f64 [2, C] as kdf_bev_bwd_affine.  No gradient rows are materialised:
// kdf_mlp_layer_bwd (mode 1) forms them from (cell_sorted, share, bits).
int kdf_bev_bwd_share(const void *grad_grid_bf16, const void *z_bf16, const void *grid_bf16, const void *grid_z_bf16,
                      const int32_t *offsets, int B, int64_t N, int C, int H, int W, void *share_bf16, void *bits_u8,
                      double *sums, void *stream) {
    KDF_CHECK_ARG(B >= 0 && N >= 0 && H > 0 && W > 0, "bev_bwd_share: bad sizes");
    KDF_CHECK_ARG(C == 64 || C == 128 || C == 256, "bev_bwd_share: C=%d not supported (64, 128, 256)", C);
    KDF_CHECK_ARG(sums, "bev_bwd_share: null pointer");
    cudaStream_t st = as_stream(stream);
    KDF_CUDA(cudaMemsetAsync(sums, 0, sizeof(double) * 2 * C, st));
    if ((int64_t)B * N == 0) return KDF_OK;
    KDF_CHECK_ARG(grad_grid_bf16 && z_bf16 && grid_bf16 && grid_z_bf16 && offsets && share_bf16 && bits_u8, "bev_bwd_share: null pointer");
    const int64_t n_cells = (int64_t)B * H * W;
    int64_t blocks = (n_cells + 7) / 8;
    if (blocks > (int64_t)sm_count() * 4) blocks = (int64_t)sm_count() * 4;
    typedef const __nv_bfloat16 *cb;
#define KDF_BS(L)                                                                                                             \
    bev_bwd_share_kernel<L, 3><<<(int)blocks, 256, 0, st>>>((cb)grad_grid_bf16, (cb)z_bf16, (cb)grid_bf16, (cb)grid_z_bf16, offsets, \
        reinterpret_cast<__nv_bfloat16 *>(share_bf16), reinterpret_cast<uint8_t *>(bits_u8), sums, n_cells, N, H * W)
    if (C == 64) KDF_BS(8); else if (C == 128) KDF_BS(16); else KDF_BS(32);
#undef KDF_BS
    KDF_LAUNCH_CHECK();
    return KDF_OK;
}

int kdf_bev_bwd_affine(const void *grad_grid_bf16, const void *z_bf16, const void *grid_bf16, const void *grid_z_bf16,
                       const int32_t *order, const int32_t *offsets, const int32_t *cell,
                       int B, int64_t N, int C, int H, int W, void *dy_bf16, double *sums, void *stream) {
    KDF_CHECK_ARG(B >= 0 && N >= 0 && H > 0 && W > 0, "bev_bwd_affine: bad sizes");
    KDF_CHECK_ARG(C == 64 || C == 128 || C == 256, "bev_bwd_affine: C=%d not supported (64, 128, 256)", C);
    KDF_CHECK_ARG(sums, "bev_bwd_affine: null pointer");
    cudaStream_t st = as_stream(stream);
    KDF_CUDA(cudaMemsetAsync(sums, 0, sizeof(double) * 2 * C, st));
    const int64_t total = (int64_t)B * N;
    if (total == 0) return KDF_OK;
    KDF_CHECK_ARG(grad_grid_bf16 && z_bf16 && grid_bf16 && grid_z_bf16 && order && offsets && dy_bf16,
                  "bev_bwd_affine: null pointer");
    const int64_t n_cells = (int64_t)B * H * W;
    int64_t blocks = (n_cells + 7) / 8;
    if (blocks > (int64_t)sm_count() * 3) blocks = (int64_t)sm_count() * 3;          // persistent: per-CTA sums -> few atomics
    typedef const __nv_bfloat16 *cb;
#define KDF_BA(L, MB)                                                                                           \
    bev_bwd_affine_kernel<L, MB><<<(int)blocks, 256, 0, st>>>((cb)grad_grid_bf16, (cb)z_bf16, (cb)grid_bf16, (cb)grid_z_bf16, \
        order, offsets, cell, reinterpret_cast<__nv_bfloat16 *>(dy_bf16), sums, n_cells, N, H * W, total)
    if (C == 64) KDF_BA(8, 3); else if (C == 128) KDF_BA(16, 3); else KDF_BA(32, 3);
#undef KDF_BA
    KDF_LAUNCH_CHECK();
    return KDF_OK;
}

// Sums over all points of (x, y, z, i) and of their 10 distinct pairwise products: the first MLP layer is
// linear in the point, so its BatchNorm statistics (and the weight gradient's z1-term) follow from these
// 14 numbers -- no pass over a [M,64] activation is needed.  out f64 [14] (zeroed by the call):
// [0..3] = sum x_k, then xx, xy, xz, xi, yy, yz, yi, zz, zi, ii.
__global__ void __launch_bounds__(256)
point_moments_kernel(const float4 *__restrict__ pts, int64_t M, double *__restrict__ out) {
    __shared__ float red[8][14];
    float acc[14];
#pragma unroll
    for (int i = 0; i < 14; ++i) acc[i] = 0.f;
    const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < M; i += nthreads) {
        const float4 p = ldg_stream_f4(pts + i);
        acc[0] += p.x; acc[1] += p.y; acc[2] += p.z; acc[3] += p.w;
        acc[4] = fmaf(p.x, p.x, acc[4]); acc[5] = fmaf(p.x, p.y, acc[5]); acc[6] = fmaf(p.x, p.z, acc[6]); acc[7] = fmaf(p.x, p.w, acc[7]);
        acc[8] = fmaf(p.y, p.y, acc[8]); acc[9] = fmaf(p.y, p.z, acc[9]); acc[10] = fmaf(p.y, p.w, acc[10]);
        acc[11] = fmaf(p.z, p.z, acc[11]); acc[12] = fmaf(p.z, p.w, acc[12]); acc[13] = fmaf(p.w, p.w, acc[13]);
    }
#pragma unroll
    for (int i = 0; i < 14; ++i) acc[i] = warp_sum(acc[i]);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < 14; ++i) red[warp][i] = acc[i];
    }
    __syncthreads();
    if (threadIdx.x < 14) {
        float v = 0.f;
        for (int w = 0; w < 8; ++w) v += red[w][threadIdx.x];
        atomicAdd(out + threadIdx.x, (double)v);
    }
}

int kdf_point_moments(const float *points, int64_t M, double *out14, void *stream) {
    KDF_CHECK_ARG(M >= 0 && out14, "point_moments: bad arguments");
    cudaStream_t st = as_stream(stream);
    KDF_CUDA(cudaMemsetAsync(out14, 0, sizeof(double) * 14, st));
    if (M == 0) return KDF_OK;
    KDF_CHECK_ARG(points && (reinterpret_cast<uintptr_t>(points) & 15) == 0, "point_moments: points must be 16-byte aligned [M,4]");
    int64_t blocks = (M + 256 * 16 - 1) / (256 * 16);
    if (blocks > (int64_t)sm_count() * 8) blocks = (int64_t)sm_count() * 8;
    point_moments_kernel<<<(int)blocks, 256, 0, st>>>(reinterpret_cast<const float4 *>(points), M, out14);
    KDF_LAUNCH_CHECK();
    return KDF_OK;
}

}  // extern "C"
