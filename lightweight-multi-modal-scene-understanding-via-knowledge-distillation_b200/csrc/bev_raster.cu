// BEV label rasterisation on the device.
//
// The reference builds the 64x64 training target of a frame with a Python loop over its ~100k labelled
// points (src/data_loading/pandaset_dataset.py:23-45): points inside the closed range are binned with the
// SAME fp32 formula the LiDAR encoder uses for its cell ids, and a cell keeps the FIRST non-zero label that
// lands in it, in point order.  "First in point order" is an integer reduction: per cell, the MINIMUM point
// index among the points that carry a non-zero label.  Two kernels:
//
//   raster_first_kernel : one coalesced read of (x, y) + label per point, cell id with one IEEE rounding per
//                         reference op (no contraction, true division), atomicMin of the point index into a
//                         shared-memory table per CTA slice, flushed with one global atomicMin per touched cell
//   raster_label_kernel : out[b, cell] = label of that first point (0 for cells no labelled point reached)
//
// Integer work, bit-exact with the reference for ANY label alphabet (not only {0,1}).
#include <limits.h>

#include "kdf_common.cuh"

namespace kdf {

struct RasterGeom {
    float x_min, x_max, y_min, y_max, xspan, yspan, sx, sy;   // sx = W-1, sy = H-1 as fp32
    int H, W;
};

// pandaset_dataset.py:33 (closed range on the RAW coordinates -- not on the normalised ones, which is what the
// encoder tests: the two differ for the float just above x_max) and :39-40
__device__ __forceinline__ int raster_cell_of(float x, float y, const RasterGeom &g) {
    const bool inside = (x >= g.x_min) && (x <= g.x_max) && (y >= g.y_min) && (y <= g.y_max);   // NaN -> false
    if (!inside) return -1;
    int col = (int)__fmul_rn(__fdiv_rn(__fsub_rn(x, g.x_min), g.xspan), g.sx);     // .astype(int): truncation
    int row = (int)__fmul_rn(__fdiv_rn(__fsub_rn(y, g.y_min), g.yspan), g.sy);
    col = min(max(col, 0), g.W - 1);
    row = min(max(row, 0), g.H - 1);
    return row * g.W + col;
}

__global__ void __launch_bounds__(512)
raster_first_kernel(const float *__restrict__ points, int stride, const int64_t *__restrict__ labels, int64_t N,
                    int64_t slice, RasterGeom g, int32_t *__restrict__ first) {
    extern __shared__ int table[];
    const int HW = g.H * g.W;
    const int b = blockIdx.y;
    for (int i = threadIdx.x; i < HW; i += 512) table[i] = INT_MAX;
    __syncthreads();
    const int64_t beg = (int64_t)blockIdx.x * slice, end = (beg + slice < N) ? beg + slice : N;
    const float *pb = points + (int64_t)b * N * stride;
    const int64_t *lb = labels + (int64_t)b * N;
    for (int64_t i = beg + threadIdx.x; i < end; i += 512) {
        if (__ldg(lb + i) == 0) continue;                                       // :43 only non-zero labels claim a cell
        float x, y;
        if (stride == 4) {
            const float4 p = ldg_stream_f4(reinterpret_cast<const float4 *>(pb) + i);
            x = p.x; y = p.y;
        } else {
            x = __ldg(pb + i * stride);
            y = __ldg(pb + i * stride + 1);
        }
        const int cell = raster_cell_of(x, y, g);
        if (cell >= 0) atomicMin(&table[cell], (int)i);
    }
    __syncthreads();
    int32_t *fb = first + (int64_t)b * HW;
    for (int i = threadIdx.x; i < HW; i += 512) {
        const int v = table[i];
        if (v != INT_MAX) atomicMin(fb + i, v);
    }
}

__global__ void __launch_bounds__(256)
raster_label_kernel(const int32_t *__restrict__ first, const int64_t *__restrict__ labels, int64_t N, int HW,
                    int64_t total, int64_t *__restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int f = first[i];
    out[i] = (f == INT_MAX) ? 0 : __ldg(labels + (i / HW) * N + f);
}

__global__ void __launch_bounds__(256)
raster_init_kernel(int32_t *__restrict__ first, int64_t total) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < total) first[i] = INT_MAX;
}

}  // namespace kdf

extern "C" int kdf_bev_rasterize(const float *points, int point_stride, const int64_t *labels, int B, int64_t N,
                                 float x_min, float x_max, float y_min, float y_max, float xspan, float yspan,
                                 int H, int W, int32_t *first_ws, int64_t *out, void *stream) {
    using namespace kdf;
    KDF_CHECK_ARG(B >= 0 && N >= 0 && H > 0 && W > 0, "bev_rasterize: bad sizes");
    KDF_CHECK_ARG(N < INT_MAX, "bev_rasterize: N=%lld does not fit the int32 point index", (long long)N);
    KDF_CHECK_ARG(point_stride >= 2, "bev_rasterize: point_stride=%d < 2", point_stride);
    KDF_CHECK_ARG((int64_t)H * W * 4 <= 200 * 1024, "bev_rasterize: grid %dx%d exceeds the shared-memory table", H, W);
    KDF_CHECK_ARG(first_ws && out && ((points && labels) || N == 0 || B == 0), "bev_rasterize: null pointer");
    if (B == 0) return KDF_OK;
    cudaStream_t st = as_stream(stream);
    const int HW = H * W;
    const int64_t total = (int64_t)B * HW;
    raster_init_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(first_ws, total);
    KDF_LAUNCH_CHECK();
    if (N > 0) {
        RasterGeom g{x_min, x_max, y_min, y_max, xspan, yspan, (float)(W - 1), (float)(H - 1), H, W};
        // about four CTAs per SM over the batch, slices of at least 4096 points
        int64_t slices = (4 * (int64_t)sm_count() + B - 1) / B;
        if (slices < 1) slices = 1;
        int64_t slice = (N + slices - 1) / slices;
        if (slice < 4096) slice = 4096;
        slices = (N + slice - 1) / slice;
        const size_t smem = (size_t)HW * 4;
        if (smem > 48 * 1024)
            KDF_CUDA(cudaFuncSetAttribute(raster_first_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        raster_first_kernel<<<dim3((unsigned)slices, (unsigned)B), 512, smem, st>>>(points, point_stride, labels, N, slice, g, first_ws);
        KDF_LAUNCH_CHECK();
    }
    raster_label_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(first_ws, labels, N, HW, total, out);
    KDF_LAUNCH_CHECK();
    return KDF_OK;
}
