// FPN-lite merge (sm_100a): out = base + bilinear(lo_a) [+ bilinear(lo_b)] over pixel-major (NHWC) rows.
//
// Replaces, in CameraFPNLite.forward (reference src/models/fusion_module.py:51-64), the per-stage
// F.interpolate(..., mode="bilinear", align_corners=False) to the largest map (:61-62) and the running sum
// (:63): two upsample kernels plus two strided adds per model in eager PyTorch, one streaming pass here.
// Source index / weights follow ATen's upsample_bilinear2d (align_corners=False):
//   src = scale*(dst+0.5)-0.5 clamped at 0, i0 = floor(src), i1 = min(i0+1, in-1), lambda = src-i0,
//   out = (1-ly)*((1-lx)*v00 + lx*v01) + ly*((1-lx)*v10 + lx*v11)          (fp32 arithmetic)
// The backward is the adjoint of that resize in gather form for the exact 2x case the model uses
// (32x32 -> 64x64): every low-resolution pixel collects its (up to) 4x4 high-resolution taps with weights
// {0.25, 0.75|1, 0.75|1, 0.25} per axis -- deterministic, no atomics.  The gradient w.r.t. `base` is the
// incoming gradient itself, and both low-resolution inputs receive the same adjoint (computed once).
#include "kdf_common.cuh"

namespace kdf {

template <typename T> struct Row16;        // 16 bytes of one pixel's channels <-> floats
template <> struct Row16<float> {
    static constexpr int N = 4;
    static __device__ __forceinline__ void unpack(const uint4 &u, float *v) {
        v[0] = __uint_as_float(u.x); v[1] = __uint_as_float(u.y); v[2] = __uint_as_float(u.z); v[3] = __uint_as_float(u.w);
    }
    static __device__ __forceinline__ uint4 pack(const float *v) {
        return make_uint4(__float_as_uint(v[0]), __float_as_uint(v[1]), __float_as_uint(v[2]), __float_as_uint(v[3]));
    }
};
template <> struct Row16<__nv_bfloat16> {
    static constexpr int N = 8;
    static __device__ __forceinline__ void unpack(const uint4 &u, float *v) {
        v[0] = bf16_lo(u.x); v[1] = bf16_hi(u.x); v[2] = bf16_lo(u.y); v[3] = bf16_hi(u.y);
        v[4] = bf16_lo(u.z); v[5] = bf16_hi(u.z); v[6] = bf16_lo(u.w); v[7] = bf16_hi(u.w);
    }
    static __device__ __forceinline__ uint4 pack(const float *v) {
        return make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
    }
};

struct Tap { int i0, i1; float l0, l1; };
__device__ __forceinline__ Tap bilinear_tap(int dst, float scale, int in_size) {
    float src = scale * ((float)dst + 0.5f) - 0.5f;
    if (src < 0.f) src = 0.f;
    Tap t;
    t.i0 = (int)src;
    if (t.i0 > in_size - 1) t.i0 = in_size - 1;
    t.i1 = t.i0 + (t.i0 < in_size - 1 ? 1 : 0);
    t.l1 = src - (float)t.i0;
    t.l0 = 1.f - t.l1;
    return t;
}

template <typename T>
__global__ void __launch_bounds__(256)
fpn_merge_fwd_kernel(const T *__restrict__ base, const T *__restrict__ lo_a, const T *__restrict__ lo_b,
                     int B, int H, int W, int h, int w, int C, float sy, float sx, T *__restrict__ out) {
    constexpr int V = Row16<T>::N;
    const int cg = C / V;                                   // 16-byte groups per pixel
    const int64_t total = (int64_t)B * H * W * cg;
    const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += nthreads) {
        const int g = (int)(i % cg);
        const int64_t pix = i / cg;
        const int x = (int)(pix % W);
        const int y = (int)((pix / W) % H);
        const int64_t b = pix / ((int64_t)W * H);
        const Tap ty = bilinear_tap(y, sy, h), tx = bilinear_tap(x, sx, w);
        float acc[V];
        Row16<T>::unpack(ldg_stream_u4(reinterpret_cast<const uint4 *>(base + pix * C) + g), acc);
        const int64_t o00 = ((b * h + ty.i0) * w + tx.i0) * C, o01 = ((b * h + ty.i0) * w + tx.i1) * C;
        const int64_t o10 = ((b * h + ty.i1) * w + tx.i0) * C, o11 = ((b * h + ty.i1) * w + tx.i1) * C;
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            const T *lo = s == 0 ? lo_a : lo_b;
            if (lo == nullptr) continue;
            float v00[V], v01[V], v10[V], v11[V];
            Row16<T>::unpack(__ldg(reinterpret_cast<const uint4 *>(lo + o00) + g), v00);
            Row16<T>::unpack(__ldg(reinterpret_cast<const uint4 *>(lo + o01) + g), v01);
            Row16<T>::unpack(__ldg(reinterpret_cast<const uint4 *>(lo + o10) + g), v10);
            Row16<T>::unpack(__ldg(reinterpret_cast<const uint4 *>(lo + o11) + g), v11);
#pragma unroll
            for (int q = 0; q < V; ++q)
                acc[q] += ty.l0 * (tx.l0 * v00[q] + tx.l1 * v01[q]) + ty.l1 * (tx.l0 * v10[q] + tx.l1 * v11[q]);
        }
        *(reinterpret_cast<uint4 *>(out + pix * C) + g) = Row16<T>::pack(acc);
    }
}

// adjoint of the exact 2x bilinear resize (H = 2h, W = 2w), gather form
template <typename T>
__global__ void __launch_bounds__(256)
fpn_up2_bwd_kernel(const T *__restrict__ gout, int B, int h, int w, int C, T *__restrict__ glo) {
    constexpr int V = Row16<T>::N;
    const int cg = C / V, H = 2 * h, W = 2 * w;
    const int64_t total = (int64_t)B * h * w * cg;
    const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += nthreads) {
        const int g = (int)(i % cg);
        const int64_t pix = i / cg;
        const int ix = (int)(pix % w);
        const int iy = (int)((pix / w) % h);
        const int64_t b = pix / ((int64_t)w * h);
        const float wy[4] = {0.25f, iy == 0 ? 1.f : 0.75f, iy == h - 1 ? 1.f : 0.75f, 0.25f};
        const float wx[4] = {0.25f, ix == 0 ? 1.f : 0.75f, ix == w - 1 ? 1.f : 0.75f, 0.25f};
        float acc[V];
#pragma unroll
        for (int q = 0; q < V; ++q) acc[q] = 0.f;
#pragma unroll
        for (int dy = 0; dy < 4; ++dy) {
            const int oy = 2 * iy - 1 + dy;
            if (oy < 0 || oy >= H) continue;
#pragma unroll
            for (int dx = 0; dx < 4; ++dx) {
                const int ox = 2 * ix - 1 + dx;
                if (ox < 0 || ox >= W) continue;
                float v[V];
                Row16<T>::unpack(__ldg(reinterpret_cast<const uint4 *>(gout + ((b * H + oy) * W + ox) * C) + g), v);
                const float wgt = wy[dy] * wx[dx];
#pragma unroll
                for (int q = 0; q < V; ++q) acc[q] = fmaf(wgt, v[q], acc[q]);
            }
        }
        *(reinterpret_cast<uint4 *>(glo + pix * C) + g) = Row16<T>::pack(acc);
    }
}

}  // namespace kdf

using namespace kdf;

extern "C" {

int kdf_fpn_merge_fwd(const void *base, const void *lo_a, const void *lo_b, int dtype,
                      int B, int H, int W, int h, int w, int C, void *out, void *stream) {
    KDF_CHECK_ARG(B >= 0 && H > 0 && W > 0 && h > 0 && w > 0, "fpn_merge_fwd: bad sizes");
    KDF_CHECK_ARG(dtype == KDF_F32 || dtype == KDF_BF16, "fpn_merge_fwd: bad dtype %d", dtype);
    const int V = dtype == KDF_F32 ? 4 : 8;
    KDF_CHECK_ARG(C > 0 && C % V == 0, "fpn_merge_fwd: C=%d must be a multiple of %d", C, V);
    if (B == 0) return KDF_OK;
    KDF_CHECK_ARG(base && lo_a && out, "fpn_merge_fwd: null pointer");
    KDF_CHECK_ARG(((reinterpret_cast<uintptr_t>(base) | reinterpret_cast<uintptr_t>(lo_a) | reinterpret_cast<uintptr_t>(lo_b) |
                    reinterpret_cast<uintptr_t>(out)) & 15) == 0, "fpn_merge_fwd: buffers must be 16-byte aligned");
    const int64_t total = (int64_t)B * H * W * (C / V);
    int64_t blocks = (total + 255) / 256;
    if (blocks > (int64_t)sm_count() * 16) blocks = (int64_t)sm_count() * 16;
    const float sy = (float)h / (float)H, sx = (float)w / (float)W;       // ATen: area_pixel_compute_scale, align_corners=False
    cudaStream_t st = as_stream(stream);
    if (dtype == KDF_F32)
        fpn_merge_fwd_kernel<float><<<(int)blocks, 256, 0, st>>>((const float *)base, (const float *)lo_a, (const float *)lo_b,
                                                                 B, H, W, h, w, C, sy, sx, (float *)out);
    else
        fpn_merge_fwd_kernel<__nv_bfloat16><<<(int)blocks, 256, 0, st>>>((const __nv_bfloat16 *)base, (const __nv_bfloat16 *)lo_a,
                                                                         (const __nv_bfloat16 *)lo_b, B, H, W, h, w, C, sy, sx,
                                                                         (__nv_bfloat16 *)out);
    KDF_LAUNCH_CHECK();
    return KDF_OK;
}

int kdf_fpn_up2_bwd(const void *grad_out, int dtype, int B, int h, int w, int C, void *grad_lo, void *stream) {
    KDF_CHECK_ARG(B >= 0 && h > 0 && w > 0, "fpn_up2_bwd: bad sizes");
    KDF_CHECK_ARG(dtype == KDF_F32 || dtype == KDF_BF16, "fpn_up2_bwd: bad dtype %d", dtype);
    const int V = dtype == KDF_F32 ? 4 : 8;
    KDF_CHECK_ARG(C > 0 && C % V == 0, "fpn_up2_bwd: C=%d must be a multiple of %d", C, V);
    if (B == 0) return KDF_OK;
    KDF_CHECK_ARG(grad_out && grad_lo, "fpn_up2_bwd: null pointer");
    KDF_CHECK_ARG(((reinterpret_cast<uintptr_t>(grad_out) | reinterpret_cast<uintptr_t>(grad_lo)) & 15) == 0,
                  "fpn_up2_bwd: buffers must be 16-byte aligned");
    const int64_t total = (int64_t)B * h * w * (C / V);
    int64_t blocks = (total + 255) / 256;
    if (blocks > (int64_t)sm_count() * 16) blocks = (int64_t)sm_count() * 16;
    cudaStream_t st = as_stream(stream);
    if (dtype == KDF_F32)
        fpn_up2_bwd_kernel<float><<<(int)blocks, 256, 0, st>>>((const float *)grad_out, B, h, w, C, (float *)grad_lo);
    else
        fpn_up2_bwd_kernel<__nv_bfloat16><<<(int)blocks, 256, 0, st>>>((const __nv_bfloat16 *)grad_out, B, h, w, C,
                                                                       (__nv_bfloat16 *)grad_lo);
    KDF_LAUNCH_CHECK();
    return KDF_OK;
}

}  // extern "C"
