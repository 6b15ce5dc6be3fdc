// Depthwise 3x3 convolution (padding 1, stride 1 or 2, no bias) over pixel-major (NHWC) maps, sm_100a.
//
// The camera branch's inverted-residual blocks (reference src/models/camera_encoder.py:9-51), the FPN-lite
// smoothing and the segmentation head (fusion_module.py:20-34) run a depthwise 3x3 between two 1x1
// convolutions.  It is a 9-tap stencil per channel -- 18 FLOP per output element against 2*s bytes in and
// out -- i.e. pure HBM traffic; the library kernels the eager path lands on take 3-4x the streaming time
// (2.0 ms of the 14.8 ms step).  Three kernels here:
//   forward : each thread owns 4 channels (its 9x4 taps stay in registers) and R=4 vertically adjacent output
//             pixels: 18 (stride 1) or 27 (stride 2) input chunks, all in flight at once, for 4 outputs instead
//             of 36; fp32 accumulation.  The same kernel with the taps flipped is the stride-1 data gradient.
//   dgrad2  : stride-2 data gradient in gather form, one 2x2 input patch per thread (no parity divergence).
//   wgrad   : per-thread [9][4] partial sums over strips of 8 output pixels with a sliding 3x3 input window,
//             block reduction, fp32 atomics.
#include <stdlib.h>

#include "kdf_common.cuh"

namespace kdf {

// A thread owns 4 channels (8-byte accesses for bf16, 16-byte for fp32) so that its 9x4 taps live in registers.
constexpr int DW_V = 4;
constexpr int DW_R = 4;          // forward: output rows per thread
constexpr int DW_L = 8;          // weight gradient: output pixels per strip (sliding window along x)

template <typename T> struct DwChunk;      // 4 consecutive channels of one pixel
template <> struct DwChunk<float> {
    typedef uint4 raw_t;
    static __device__ __forceinline__ raw_t ld(const float *p) { return __ldg(reinterpret_cast<const uint4 *>(p)); }
    static __device__ __forceinline__ raw_t zero() { return make_uint4(0u, 0u, 0u, 0u); }
    static __device__ __forceinline__ void unpack(const raw_t &u, float *v) {
        v[0] = __uint_as_float(u.x); v[1] = __uint_as_float(u.y); v[2] = __uint_as_float(u.z); v[3] = __uint_as_float(u.w);
    }
    static __device__ __forceinline__ void st(float *p, const float *v) { *reinterpret_cast<float4 *>(p) = make_float4(v[0], v[1], v[2], v[3]); }
    static __device__ __forceinline__ void round(float *) {}                      // storage rounding: none for fp32
};
template <> struct DwChunk<__nv_bfloat16> {
    typedef uint2 raw_t;
    static __device__ __forceinline__ raw_t ld(const __nv_bfloat16 *p) { return __ldg(reinterpret_cast<const uint2 *>(p)); }
    static __device__ __forceinline__ raw_t zero() { return make_uint2(0u, 0u); }
    static __device__ __forceinline__ void unpack(const raw_t &u, float *v) {
        v[0] = bf16_lo(u.x); v[1] = bf16_hi(u.x); v[2] = bf16_lo(u.y); v[3] = bf16_hi(u.y);
    }
    static __device__ __forceinline__ void st(__nv_bfloat16 *p, const float *v) {
        *reinterpret_cast<uint2 *>(p) = make_uint2(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]));
    }
    // the values as they are stored: two packed conversions + four shifts (a scalar cvt.rn.bf16.f32 per element compiles
    // to F2F on the conversion pipe, 8x the issue cost of an FMA -- it was as expensive as the convolution itself)
    static __device__ __forceinline__ void round(float *v) {
        const uint32_t a = pack_bf16(v[0], v[1]), b = pack_bf16(v[2], v[3]);
        v[0] = bf16_lo(a); v[1] = bf16_hi(a); v[2] = bf16_lo(b); v[3] = bf16_hi(b);
    }
};

// the 9 taps of this thread's 4 channels: wr[k][q] = w[(c0+q)*9 + (FLIP ? 8-k : k)]
template <bool FLIP>
__device__ __forceinline__ void dw_load_taps(const float *__restrict__ w, int c0, float (&wr)[9][DW_V]) {
#pragma unroll
    for (int q = 0; q < DW_V; ++q)
#pragma unroll
        for (int k = 0; k < 9; ++k) wr[k][q] = __ldg(w + (c0 + q) * 9 + (FLIP ? 8 - k : k));
}

// ----------------------------------------------------------------------------- forward (and stride-1 dgrad with FLIP)
template <typename T, int STRIDE, bool FLIP, int MINB = 2>
__global__ void __launch_bounds__(256, MINB)
dwconv3x3_fwd_kernel(const T *__restrict__ in, const float *__restrict__ w /* [C][9] */, T *__restrict__ out,
                     int B, int H, int W, int C, int OH, int OW, double *__restrict__ stats /* nullable [2][C] */,
                     const float *__restrict__ post_scale /* nullable [C] */, const float *__restrict__ post_shift, int post_act) {
    typedef DwChunk<T> K;
    __shared__ float sred[256 * 2 * DW_V];              // BatchNorm statistics of the stored outputs (only when asked for)
    constexpr int IN_ROWS = (DW_R - 1) * STRIDE + 3;
    // grid.x * blockDim.x covers one output row of (ox, channel group); grid.y strides over (frame, row block):
    // all index arithmetic is 32-bit and the thread keeps its channel group (taps in registers)
    const int cg = C / DW_V;
    const int oyb_n = (OH + DW_R - 1) / DW_R;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = idx < OW * cg;
    const int ox = live ? idx / cg : 0, g = live ? idx - ox * cg : 0;
    float wr[9][DW_V];
    dw_load_taps<FLIP>(w, g * DW_V, wr);
    float s_sum[DW_V], s_sq[DW_V];
#pragma unroll
    for (int q = 0; q < DW_V; ++q) { s_sum[q] = 0.f; s_sq[q] = 0.f; }
    // optional epilogue: the BatchNorm (running statistics) + activation that follows, applied to the value as it
    // would have been stored (same rounding as the two-kernel sequence)
    float psc[DW_V], psh[DW_V];
#pragma unroll
    for (int q = 0; q < DW_V; ++q) {
        psc[q] = post_scale ? __ldg(post_scale + g * DW_V + q) : 1.f;
        psh[q] = post_scale ? __ldg(post_shift + g * DW_V + q) : 0.f;
    }
    const int ix0 = ox * STRIDE - 1;
    for (int rb = blockIdx.y; live && rb < B * oyb_n; rb += gridDim.y) {
        const int b = rb / oyb_n, oyb = rb - b * oyb_n;
        const int oy0 = oyb * DW_R;
        const int iy0 = oy0 * STRIDE - 1;
        const T *ib = in + ((int64_t)b * H * W) * C + g * DW_V;
        float acc[DW_R][DW_V];
#pragma unroll
        for (int r = 0; r < DW_R; ++r)
#pragma unroll
            for (int q = 0; q < DW_V; ++q) acc[r][q] = 0.f;
        typename K::raw_t raw[IN_ROWS][3];                             // all input chunks of the item in flight at once
        // one 64-bit base per item, 32-bit offsets per tap (a frame's map is < 2^31 elements: checked by the host)
        const T *ibase = ib + ((int64_t)iy0 * W + ix0) * C;
        const int rs = W * C;
        const bool xok0 = ix0 >= 0, xok2 = ix0 + 2 < W;
#pragma unroll
        for (int j = 0; j < IN_ROWS; ++j) {
            const bool yok = (unsigned)(iy0 + j) < (unsigned)H;
            raw[j][0] = (yok && xok0) ? K::ld(ibase + j * rs) : K::zero();
            raw[j][1] = yok ? K::ld(ibase + j * rs + C) : K::zero();
            raw[j][2] = (yok && xok2) ? K::ld(ibase + j * rs + 2 * C) : K::zero();
        }
#pragma unroll
        for (int j = 0; j < IN_ROWS; ++j) {
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                float v[DW_V];
                K::unpack(raw[j][kx], v);
#pragma unroll
                for (int r = 0; r < DW_R; ++r) {
                    const int ky = j - r * STRIDE;                  // compile-time after unrolling
                    if (ky >= 0 && ky < 3) {
#pragma unroll
                        for (int q = 0; q < DW_V; ++q) acc[r][q] = fmaf(wr[ky * 3 + kx][q], v[q], acc[r][q]);
                    }
                }
            }
        }
#pragma unroll
        for (int r = 0; r < DW_R; ++r) {
            const int oy = oy0 + r;
            if (oy < OH) {
                if (post_scale) {
                    K::round(acc[r]);
#pragma unroll
                    for (int q = 0; q < DW_V; ++q) {
                        const float z = acc[r][q];
                        float y = fmaf(z, psc[q], psh[q]);
                        if (post_act == 1) y = fmaxf(y, 0.f);
                        else if (post_act == 2) y = fminf(fmaxf(y, 0.f), 6.f);
                        acc[r][q] = y;
                    }
                }
                K::st(out + (((int64_t)b * OH + oy) * OW + ox) * C + g * DW_V, acc[r]);
                if (stats) {
                    K::round(acc[r]);                                              // the values stored
#pragma unroll
                    for (int q = 0; q < DW_V; ++q) {
                        const float v = acc[r][q];
                        s_sum[q] += v;
                        s_sq[q] = fmaf(v, v, s_sq[q]);
                    }
                }
            }
        }
    }
    if (stats) {                                         // threads tid = g (mod cg) share a channel group
#pragma unroll
        for (int q = 0; q < DW_V; ++q) { sred[threadIdx.x * 2 * DW_V + q] = s_sum[q]; sred[threadIdx.x * 2 * DW_V + DW_V + q] = s_sq[q]; }
        __syncthreads();
        const int g0 = (blockIdx.x * blockDim.x) % cg;   // channel group of thread 0 of this CTA
        for (int i = threadIdx.x; i < cg * 2 * DW_V; i += blockDim.x) {
            const int gg = i / (2 * DW_V), e = i - gg * 2 * DW_V;
            // first thread of the CTA whose channel group is gg, then every cg-th
            int t0 = gg - g0;
            if (t0 < 0) t0 += cg;
            float v = 0.f;
            for (int t = t0; t < (int)blockDim.x; t += cg) v += sred[t * 2 * DW_V + e];
            const int which = e / DW_V, q = e - which * DW_V;
            if (v != 0.f) atomicAdd(stats + which * C + gg * DW_V + q, (double)v);
        }
    }
}

// ----------------------------------------------------------------------------- stride-2 data gradient (gather, 2x2 input patch per thread)
// out(oy,ox) reads in(2*oy+ky-1, 2*ox+kx-1).  For the input patch rows {2a, 2a+1} x cols {2c, 2c+1}:
//   even row 2a   <- (oy=a,   ky=1);          odd row 2a+1 <- (oy=a, ky=2) and (oy=a+1, ky=0); same along x.
// Four gradient chunks in, four out, 9 taps: no parity divergence inside a warp.
template <typename T>
__global__ void __launch_bounds__(256, 2)
dwconv3x3_dgrad2_kernel(const T *__restrict__ gout, const float *__restrict__ w, T *__restrict__ gin,
                        int B, int H, int W, int C, int OH, int OW) {
    typedef DwChunk<T> K;
    const int cg = C / DW_V;
    const int PH = (H + 1) / 2, PW = (W + 1) / 2;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= PW * cg) return;
    const int c = idx / cg, g = idx - c * cg;
    float wr[9][DW_V];
    dw_load_taps<false>(w, g * DW_V, wr);
    for (int rb = blockIdx.y; rb < B * PH; rb += gridDim.y) {
        const int b = rb / PH, a = rb - b * PH;
        const T *gb = gout + ((int64_t)b * OH * OW) * C + g * DW_V;
        float v[2][2][DW_V];
#pragma unroll
        for (int dy = 0; dy < 2; ++dy)
#pragma unroll
            for (int dx = 0; dx < 2; ++dx) {
                const int oy = a + dy, ox = c + dx;
                K::unpack((oy < OH && ox < OW) ? K::ld(gb + ((int64_t)oy * OW + ox) * C) : K::zero(), v[dy][dx]);
            }
        float o[2][2][DW_V];
#pragma unroll
        for (int q = 0; q < DW_V; ++q) {
            o[0][0][q] = wr[4][q] * v[0][0][q];
            o[0][1][q] = wr[5][q] * v[0][0][q] + wr[3][q] * v[0][1][q];
            o[1][0][q] = wr[7][q] * v[0][0][q] + wr[1][q] * v[1][0][q];
            o[1][1][q] = wr[8][q] * v[0][0][q] + wr[6][q] * v[0][1][q] + wr[2][q] * v[1][0][q] + wr[0][q] * v[1][1][q];
        }
        T *ob = gin + ((int64_t)b * H * W) * C + g * DW_V;
#pragma unroll
        for (int dy = 0; dy < 2; ++dy)
#pragma unroll
            for (int dx = 0; dx < 2; ++dx) {
                const int iy = 2 * a + dy, ix = 2 * c + dx;
                if (iy < H && ix < W) K::st(ob + ((int64_t)iy * W + ix) * C, o[dy][dx]);
            }
    }
}

// ----------------------------------------------------------------------------- weight gradient
// dw[c][ky][kx] = sum over (b, oy, ox) of gout(b,oy,ox,c) * in(b, s*oy+ky-1, s*ox+kx-1, c).
// A thread owns 4 channels and strips of DW_L output pixels along x with a sliding 3x3 input window
// (3*s new chunks per output instead of 9); 36 fp32 partial sums per thread, block reduction, fp32 atomics.
template <typename T, int STRIDE>
__global__ void __launch_bounds__(256, 2)
dwconv3x3_wgrad_kernel(const T *__restrict__ in, const T *__restrict__ gout, float *__restrict__ dw /* [C][9] */,
                       int B, int H, int W, int C, int OH, int OW) {
    typedef DwChunk<T> K;
    extern __shared__ float red[];                      // [blockDim/cg][cg][36]
    const int cg = C / DW_V;
    const int g = threadIdx.x % cg, lane_row = threadIdx.x / cg, rows_per_cta = blockDim.x / cg;
    float acc[9][DW_V];
#pragma unroll
    for (int k = 0; k < 9; ++k)
#pragma unroll
        for (int q = 0; q < DW_V; ++q) acc[k][q] = 0.f;
    const int nstrip = (OW + DW_L - 1) / DW_L;
    const int nitems = B * OH * nstrip;                                  // < 2^31 (checked by the host wrapper)
    for (int p = blockIdx.x * rows_per_cta + lane_row; p < nitems; p += gridDim.x * rows_per_cta) {
        const int row = p / nstrip, xs = p - row * nstrip;
        const int b = row / OH, oy = row - b * OH;
        const int ox0 = xs * DW_L;
        const T *ib = in + ((int64_t)b * H * W) * C + g * DW_V;
        const T *gb = gout + (((int64_t)b * OH + oy) * OW) * C + g * DW_V;
        const int iy0 = oy * STRIDE - 1;
        auto ld_col = [&](int ix, float (&col)[3][DW_V]) {
            const bool xok = ix >= 0 && ix < W;
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) {
                const int iy = iy0 + ky;
                K::unpack((xok && iy >= 0 && iy < H) ? K::ld(ib + ((int64_t)iy * W + ix) * C) : K::zero(), col[ky]);
            }
        };
        float win[3][3][DW_V];                           // [kx][ky][q]
        ld_col(ox0 * STRIDE - 1, win[0]);
        if (STRIDE == 1) ld_col(ox0 * STRIDE, win[1]);
#pragma unroll
        for (int l = 0; l < DW_L; ++l) {
            const int ox = ox0 + l;
            if (ox >= OW) break;
            if (STRIDE == 1) {
                ld_col(ox + 1, win[2]);
            } else {
                ld_col(2 * ox, win[1]);
                ld_col(2 * ox + 1, win[2]);
            }
            float gv[DW_V];
            K::unpack(K::ld(gb + (int64_t)ox * C), gv);
#pragma unroll
            for (int kx = 0; kx < 3; ++kx)
#pragma unroll
                for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                    for (int q = 0; q < DW_V; ++q) acc[ky * 3 + kx][q] = fmaf(gv[q], win[kx][ky][q], acc[ky * 3 + kx][q]);
            // slide: the last column of this window is the first of the next one
#pragma unroll
            for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                for (int q = 0; q < DW_V; ++q) {
                    if (STRIDE == 1) { win[0][ky][q] = win[1][ky][q]; win[1][ky][q] = win[2][ky][q]; }
                    else win[0][ky][q] = win[2][ky][q];
                }
        }
    }
#pragma unroll
    for (int k = 0; k < 9; ++k)
#pragma unroll
        for (int q = 0; q < DW_V; ++q) red[(lane_row * cg + g) * 36 + k * DW_V + q] = acc[k][q];
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

static int dw_check(const char *who, int dtype, int B, int H, int W, int C, int stride) {
    KDF_CHECK_ARG(B >= 0 && H > 0 && W > 0, "%s: bad sizes", who);
    KDF_CHECK_ARG(dtype == KDF_F32 || dtype == KDF_BF16, "%s: bad dtype %d", who, dtype);
    KDF_CHECK_ARG(stride == 1 || stride == 2, "%s: stride %d not supported (1, 2)", who, stride);
    KDF_CHECK_ARG(C > 0 && C % DW_V == 0 && C / DW_V <= 256, "%s: C=%d must be a multiple of %d, at most %d", who, C, DW_V, 256 * DW_V);
    return KDF_OK;
}

// threads per CTA: the largest multiple of the channel-group count <= 256, so that a thread keeps its channel group
static int dw_block(int cg) { return (256 / cg) * cg; }

}  // namespace kdf

using namespace kdf;

extern "C" {

static int dwconv_fwd_impl(const void *in, const float *weight, int dtype, int B, int H, int W, int C, int stride,
                           int flip, void *out, double *stats, const float *post_scale, const float *post_shift, int post_act,
                           void *stream) {
    if (int e = dw_check("dwconv3x3_fwd", dtype, B, H, W, C, stride)) return e;
    KDF_CHECK_ARG(!(flip && stride != 1), "dwconv3x3_fwd: flipped taps are the stride-1 data gradient only");
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
    const int row_blocks = B * ((OH + DW_R - 1) / DW_R);
    const int gx = (OW * cg + nt - 1) / nt;
    int gy = (sm_count() * 8 + gx - 1) / gx;                 // ~8 CTAs per SM in total; a thread then walks several row blocks
    if (gy > row_blocks) gy = row_blocks;
    const dim3 grid((unsigned)gx, (unsigned)(gy < 1 ? 1 : gy));
    cudaStream_t st = as_stream(stream);
    if (stats) KDF_CUDA(cudaMemsetAsync(stats, 0, sizeof(double) * 2 * C, st));
    // resident CTAs per SM the kernel is compiled for: 3 (85 registers, 128 B of spills) is 5 % faster than 2 (111 registers)
    // at stride 1 and 15 % slower at stride 2 (measured in the step)
    const int minb = stride == 1 ? 3 : 2;
#define KDF_DW(T, S, F)                                                                                                              \
    do {                                                                                                                             \
        if (minb == 3) dwconv3x3_fwd_kernel<T, S, F, 3><<<grid, nt, 0, st>>>((const T *)in, weight, (T *)out, B, H, W, C, OH, OW, stats, post_scale, post_shift, post_act); \
        else dwconv3x3_fwd_kernel<T, S, F, 2><<<grid, nt, 0, st>>>((const T *)in, weight, (T *)out, B, H, W, C, OH, OW, stats, post_scale, post_shift, post_act);           \
    } while (0)
    if (dtype == KDF_F32) { if (stride == 2) KDF_DW(float, 2, false); else if (flip) KDF_DW(float, 1, true); else KDF_DW(float, 1, false); }
    else { if (stride == 2) KDF_DW(__nv_bfloat16, 2, false); else if (flip) KDF_DW(__nv_bfloat16, 1, true); else KDF_DW(__nv_bfloat16, 1, false); }
#undef KDF_DW
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
    KDF_CHECK_ARG((int64_t)W * cg < (1ll << 30) && (int64_t)B * H < (1ll << 30), "dwconv3x3_bwd_data: map too large for 32-bit indexing");
    const int PH = (H + 1) / 2, PW = (W + 1) / 2;
    const int gx = (PW * cg + nt - 1) / nt;
    int gy = (sm_count() * 8 + gx - 1) / gx;
    if (gy > B * PH) gy = B * PH;
    const dim3 grid((unsigned)gx, (unsigned)(gy < 1 ? 1 : gy));
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
    const int cg = C / DW_V;
    const int nt = dw_block(cg), rows = nt / cg;
    const int64_t nitems = (int64_t)B * OH * ((OW + DW_L - 1) / DW_L);
    KDF_CHECK_ARG(nitems < (1ll << 30), "dwconv3x3_bwd_weight: map too large for 32-bit indexing");
    int64_t blocks = (nitems + rows - 1) / rows;
    if (blocks > (int64_t)sm_count() * 4) blocks = (int64_t)sm_count() * 4;
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
