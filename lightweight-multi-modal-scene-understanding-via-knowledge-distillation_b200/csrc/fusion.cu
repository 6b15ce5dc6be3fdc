// Camera-LiDAR feature fusion, fp32-exact path (sm_100a).
//
// Replaces the inline fusion of CompleteSegmentationModel.forward (reference
// src/models/fusion_module.py:242-256): the BatchNorm-apply + ReLU tail of both
// 1x1 projection blocks (fusion_module.py:8-17), the concat, the 256->128->2
// attention MLP, the 2-way softmax and the blend are ONE kernel per direction,
// where the reference runs 7 eager kernels plus a materialised [B,256,H,W]
// concat.  Pixel-major (NHWC) rows in, pixel-major rows out.
//
// This file is the fp32-arithmetic implementation used for the 1e-5 parity bar
// (CUDA-core FMA; every product and sum in fp32).  fusion_tc.cu holds the bf16
// tensor-core (tcgen05 + TMEM) implementation of the same contract, used for bf16 rows with C = 128.
#include "kdf_common.cuh"

namespace kdf {

// bf16 / C=128 tensor-core implementation (fusion_tc.cu)
bool fusion_tc_enabled();
int fusion_weighted_fwd_tc(const void *cam_pre, const void *lid_pre, int64_t M,
                           const float *csc, const float *csh, const float *lsc, const float *lsh,
                           const float *w1, const float *b1, const float *w2, const float *b2,
                           void *out, float *attn, cudaStream_t st);
int fusion_weighted_bwd_tc(const void *grad_out, const void *cam_pre, const void *lid_pre, int64_t M,
                           const float *csc, const float *csh, const float *lsc, const float *lsh,
                           const float *w1, const float *b1, const float *w2, const float *attn,
                           void *gcam, void *glid, float *gaff, float *gw1, float *gb1, float *gw2, float *gb2,
                           cudaStream_t st);

constexpr int FW_TM = 32;         // pixels per tile
constexpr int FW_THREADS = 256;   // 8 warps, warp w owns pixels 4w..4w+3

// --------------------------------------------------------------------------- affine+relu pair (minimal / concat head)
template <typename T>
__global__ void __launch_bounds__(256)
affine_relu_pair_fwd_kernel(const T *__restrict__ cam, const T *__restrict__ lid, int64_t M, int C,
                            const float *__restrict__ csc, const float *__restrict__ csh,
                            const float *__restrict__ lsc, const float *__restrict__ lsh,
                            int mode, T *__restrict__ out) {
    const int cg = C >> 2;                       // float4 groups per row; 256 % cg == 0
    const int g = threadIdx.x % cg;
    const int rows_per_pass = 256 / cg;
    const int c = g * 4;
    const float4 a0 = *reinterpret_cast<const float4 *>(csc + c), b0 = *reinterpret_cast<const float4 *>(csh + c);
    const float4 a1 = *reinterpret_cast<const float4 *>(lsc + c), b1 = *reinterpret_cast<const float4 *>(lsh + c);
    const int64_t ostride = mode ? 2 * (int64_t)C : C;
    for (int64_t m = (int64_t)blockIdx.x * rows_per_pass + threadIdx.x / cg; m < M;
         m += (int64_t)gridDim.x * rows_per_pass) {
        const float4 x = Vec4<T>::load_stream(cam + m * C + c);
        const float4 y = Vec4<T>::load_stream(lid + m * C + c);
        const float4 p = make_float4(fmaxf(x.x * a0.x + b0.x, 0.f), fmaxf(x.y * a0.y + b0.y, 0.f),
                                     fmaxf(x.z * a0.z + b0.z, 0.f), fmaxf(x.w * a0.w + b0.w, 0.f));
        const float4 q = make_float4(fmaxf(y.x * a1.x + b1.x, 0.f), fmaxf(y.y * a1.y + b1.y, 0.f),
                                     fmaxf(y.z * a1.z + b1.z, 0.f), fmaxf(y.w * a1.w + b1.w, 0.f));
        if (mode == 0) {
            Vec4<T>::store(out + m * ostride + c, make_float4(p.x + q.x, p.y + q.y, p.z + q.z, p.w + q.w));
        } else {
            Vec4<T>::store(out + m * ostride + c, p);
            Vec4<T>::store(out + m * ostride + C + c, q);
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(256)
affine_relu_pair_bwd_kernel(const T *__restrict__ gout, const T *__restrict__ cam, const T *__restrict__ lid,
                            int64_t M, int C,
                            const float *__restrict__ csc, const float *__restrict__ csh,
                            const float *__restrict__ lsc, const float *__restrict__ lsh, int mode,
                            T *__restrict__ gcam, T *__restrict__ glid, float *__restrict__ gaff) {
    extern __shared__ float sred[];              // [rows_per_pass][16] partials -> reduced per channel group
    const int cg = C >> 2;
    const int g = threadIdx.x % cg, r = threadIdx.x / cg;
    const int rows_per_pass = 256 / cg;
    const int c = g * 4;
    const float4 a0 = *reinterpret_cast<const float4 *>(csc + c), b0 = *reinterpret_cast<const float4 *>(csh + c);
    const float4 a1 = *reinterpret_cast<const float4 *>(lsc + c), b1 = *reinterpret_cast<const float4 *>(lsh + c);
    const int64_t gstride = mode ? 2 * (int64_t)C : C;
    float acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = 0.f;
    for (int64_t m = (int64_t)blockIdx.x * rows_per_pass + r; m < M; m += (int64_t)gridDim.x * rows_per_pass) {
        const float4 x = Vec4<T>::load_stream(cam + m * C + c);
        const float4 y = Vec4<T>::load_stream(lid + m * C + c);
        const float4 gc = Vec4<T>::load_stream(gout + m * gstride + c);
        const float4 gl = mode ? Vec4<T>::load_stream(gout + m * gstride + C + c) : gc;
        const float xs[4] = {x.x, x.y, x.z, x.w}, ys[4] = {y.x, y.y, y.z, y.w};
        const float gcs[4] = {gc.x, gc.y, gc.z, gc.w}, gls[4] = {gl.x, gl.y, gl.z, gl.w};
        const float A0[4] = {a0.x, a0.y, a0.z, a0.w}, B0[4] = {b0.x, b0.y, b0.z, b0.w};
        const float A1[4] = {a1.x, a1.y, a1.z, a1.w}, B1[4] = {b1.x, b1.y, b1.z, b1.w};
        float oc[4], ol[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float dc = (xs[q] * A0[q] + B0[q] > 0.f) ? gcs[q] : 0.f;
            const float dl = (ys[q] * A1[q] + B1[q] > 0.f) ? gls[q] : 0.f;
            oc[q] = dc * A0[q];
            ol[q] = dl * A1[q];
            acc[q] += dc * xs[q];        // d cam_scale
            acc[4 + q] += dc;            // d cam_shift
            acc[8 + q] += dl * ys[q];    // d lid_scale
            acc[12 + q] += dl;           // d lid_shift
        }
        Vec4<T>::store(gcam + m * C + c, make_float4(oc[0], oc[1], oc[2], oc[3]));
        Vec4<T>::store(glid + m * C + c, make_float4(ol[0], ol[1], ol[2], ol[3]));
    }
    // reduce over the rows_per_pass threads that share a channel group, then one atomic per value
#pragma unroll
    for (int i = 0; i < 16; ++i) sred[(r * cg + g) * 16 + i] = acc[i];
    __syncthreads();
    if (r == 0) {
        for (int rr = 1; rr < rows_per_pass; ++rr)
#pragma unroll
            for (int i = 0; i < 16; ++i) acc[i] += sred[(rr * cg + g) * 16 + i];
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
            for (int q = 0; q < 4; ++q) atomicAdd(gaff + (int64_t)k * C + c + q, acc[k * 4 + q]);
    }
}

// --------------------------------------------------------------------------- weighted fusion, forward
// smem: Y[FW_TM][2C+4] (post BN+ReLU, cam | lid), Ws[32][C+1] (k-chunk of W1, transposed)
template <typename T>
__global__ void __launch_bounds__(FW_THREADS)
fusion_weighted_fwd_kernel(const T *__restrict__ cam, const T *__restrict__ lid, int64_t M, int C,
                           const float *__restrict__ csc, const float *__restrict__ csh,
                           const float *__restrict__ lsc, const float *__restrict__ lsh,
                           const float *__restrict__ w1, const float *__restrict__ b1,
                           const float *__restrict__ w2, const float *__restrict__ b2,
                           T *__restrict__ out, float *__restrict__ attn) {
    extern __shared__ __align__(16) float smem[];
    const int K2 = 2 * C, YS = K2 + 4, WS = C + 1, R = C / 32;
    float *Y = smem;
    float *Ws = smem + FW_TM * YS;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int64_t n_tiles = (M + FW_TM - 1) / FW_TM;

    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t m0 = tile * FW_TM;
        __syncthreads();
        // ---- load + BN-apply + ReLU -> Y
        for (int idx = tid; idx < FW_TM * (K2 / 4); idx += FW_THREADS) {
            const int p = idx / (K2 / 4), k = (idx % (K2 / 4)) * 4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (m0 + p < M) {
                const bool is_cam = k < C;
                const int c = is_cam ? k : k - C;
                const float4 x = Vec4<T>::load_stream((is_cam ? cam : lid) + (m0 + p) * C + c);
                const float4 a = *reinterpret_cast<const float4 *>((is_cam ? csc : lsc) + c);
                const float4 b = *reinterpret_cast<const float4 *>((is_cam ? csh : lsh) + c);
                v = make_float4(fmaxf(x.x * a.x + b.x, 0.f), fmaxf(x.y * a.y + b.y, 0.f),
                                fmaxf(x.z * a.z + b.z, 0.f), fmaxf(x.w * a.w + b.w, 0.f));
            }
            *reinterpret_cast<float4 *>(Y + p * YS + k) = v;
        }
        // ---- hidden = relu(Y . W1^T + b1): acc[p][r] for pixel 4w+p, hidden unit lane+32r
        float acc[4][8];
#pragma unroll
        for (int p = 0; p < 4; ++p)
#pragma unroll
            for (int r = 0; r < 8; ++r) acc[p][r] = 0.f;
        for (int k0 = 0; k0 < K2; k0 += 32) {
            __syncthreads();
            for (int idx = tid; idx < C * 32; idx += FW_THREADS) {
                const int kk = idx & 31, j = idx >> 5;
                Ws[kk * WS + j] = __ldg(w1 + (int64_t)j * K2 + k0 + kk);
            }
            __syncthreads();
#pragma unroll 4
            for (int kk = 0; kk < 32; ++kk) {
                float a[4];
#pragma unroll
                for (int p = 0; p < 4; ++p) a[p] = Y[(4 * w + p) * YS + k0 + kk];
#pragma unroll
                for (int r = 0; r < 8; ++r)
                    if (r < R) {
                        const float wv = Ws[kk * WS + lane + 32 * r];
#pragma unroll
                        for (int p = 0; p < 4; ++p) acc[p][r] = fmaf(a[p], wv, acc[p][r]);
                    }
            }
        }
        // ---- 2 attention logits per pixel, softmax, blend
        float a0[4], a1[4];
#pragma unroll
        for (int p = 0; p < 4; ++p) { a0[p] = 0.f; a1[p] = 0.f; }
#pragma unroll
        for (int r = 0; r < 8; ++r)
            if (r < R) {
                const int j = lane + 32 * r;
                const float bj = __ldg(b1 + j), u0 = __ldg(w2 + j), u1 = __ldg(w2 + C + j);
#pragma unroll
                for (int p = 0; p < 4; ++p) {
                    const float h = fmaxf(acc[p][r] + bj, 0.f);
                    a0[p] = fmaf(h, u0, a0[p]);
                    a1[p] = fmaf(h, u1, a1[p]);
                }
            }
        const float bb0 = __ldg(b2), bb1 = __ldg(b2 + 1);
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            const float s0 = warp_sum(a0[p]) + bb0, s1 = warp_sum(a1[p]) + bb1;
            const float mx = fmaxf(s0, s1);
            const float e0 = expf(s0 - mx), e1 = expf(s1 - mx);
            const float inv = 1.f / (e0 + e1);
            const float w0 = e0 * inv, w1v = e1 * inv;
            const int64_t m = m0 + 4 * w + p;
            if (m < M) {
                if (lane == 0) { attn[2 * m] = w0; attn[2 * m + 1] = w1v; }
                const float *yr = Y + (4 * w + p) * YS;
                for (int c = lane * 4; c < C; c += 128) {
                    const float4 yc = *reinterpret_cast<const float4 *>(yr + c);
                    const float4 yl = *reinterpret_cast<const float4 *>(yr + C + c);
                    Vec4<T>::store(out + m * C + c,
                                   make_float4(yc.x * w0 + yl.x * w1v, yc.y * w0 + yl.y * w1v,
                                               yc.z * w0 + yl.z * w1v, yc.w * w0 + yl.w * w1v));
                }
            }
        }
    }
}

// --------------------------------------------------------------------------- weighted fusion, backward
// Persistent CTAs; per-CTA register accumulators for dW1 (C*2C/256 per thread),
// flushed with one atomicAdd per element at the end.
template <typename T, int C>
__global__ void __launch_bounds__(FW_THREADS, 1)
fusion_weighted_bwd_kernel(const T *__restrict__ gout, const T *__restrict__ cam, const T *__restrict__ lid,
                           int64_t M,
                           const float *__restrict__ csc, const float *__restrict__ csh,
                           const float *__restrict__ lsc, const float *__restrict__ lsh,
                           const float *__restrict__ w1, const float *__restrict__ b1,
                           const float *__restrict__ w2, const float *__restrict__ attn,
                           T *__restrict__ gcam, T *__restrict__ glid, float *__restrict__ gaff,
                           float *__restrict__ gw1, float *__restrict__ gb1, float *__restrict__ gw2,
                           float *__restrict__ gb2) {
    constexpr int K2 = 2 * C, YS = K2 + 4, HS = C + 1, R = C / 32, R2 = K2 / 32;
    constexpr int JB = 16, KB = 16;                  // dW1 ownership: j = jb + 16*qj, k = kb + 16*qk
    constexpr int NJ = C / JB, NK = K2 / KB;
    extern __shared__ __align__(16) float smem[];
    float *Y = smem;                                 // [TM][YS]   post BN+ReLU activations
    float *G = Y + FW_TM * YS;                       // [TM][C]    upstream gradient tile
    float *Hs = G + FW_TM * C;                       // [TM][HS]   hidden, then d hidden
    float *Ws = Hs + FW_TM * HS;                     // max(32*(C+1), 32*K2) weight staging
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int jb = tid / KB, kb = tid % KB;
    const int64_t n_tiles = (M + FW_TM - 1) / FW_TM;

    float dW1[NJ][NK];
#pragma unroll
    for (int a = 0; a < NJ; ++a)
#pragma unroll
        for (int b = 0; b < NK; ++b) dW1[a][b] = 0.f;
    float gsc[R2], gsh[R2];                          // d scale / d shift for k = lane + 32 r (cam | lid)
#pragma unroll
    for (int r = 0; r < R2; ++r) { gsc[r] = 0.f; gsh[r] = 0.f; }
    float gW2a[R], gW2b[R], gB1[R];
#pragma unroll
    for (int r = 0; r < R; ++r) { gW2a[r] = 0.f; gW2b[r] = 0.f; gB1[r] = 0.f; }
    float gB2a = 0.f, gB2b = 0.f;

    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t m0 = tile * FW_TM;
        __syncthreads();
        // ---- phase 0: Y and G tiles
        for (int idx = tid; idx < FW_TM * (K2 / 4); idx += FW_THREADS) {
            const int p = idx / (K2 / 4), k = (idx % (K2 / 4)) * 4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (m0 + p < M) {
                const bool is_cam = k < C;
                const int c = is_cam ? k : k - C;
                const float4 x = Vec4<T>::load_stream((is_cam ? cam : lid) + (m0 + p) * C + c);
                const float4 a = *reinterpret_cast<const float4 *>((is_cam ? csc : lsc) + c);
                const float4 b = *reinterpret_cast<const float4 *>((is_cam ? csh : lsh) + c);
                v = make_float4(fmaxf(x.x * a.x + b.x, 0.f), fmaxf(x.y * a.y + b.y, 0.f),
                                fmaxf(x.z * a.z + b.z, 0.f), fmaxf(x.w * a.w + b.w, 0.f));
            }
            *reinterpret_cast<float4 *>(Y + p * YS + k) = v;
        }
        for (int idx = tid; idx < FW_TM * (C / 4); idx += FW_THREADS) {
            const int p = idx / (C / 4), c = (idx % (C / 4)) * 4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (m0 + p < M) v = Vec4<T>::load_stream(gout + (m0 + p) * C + c);
            *reinterpret_cast<float4 *>(G + p * C + c) = v;
        }
        // ---- phase 1: recompute hidden (same arithmetic as the forward)
        {
            float acc[4][R];
#pragma unroll
            for (int p = 0; p < 4; ++p)
#pragma unroll
                for (int r = 0; r < R; ++r) acc[p][r] = 0.f;
            for (int k0 = 0; k0 < K2; k0 += 32) {
                __syncthreads();
                for (int idx = tid; idx < C * 32; idx += FW_THREADS) {
                    const int kk = idx & 31, j = idx >> 5;
                    Ws[kk * HS + j] = __ldg(w1 + (int64_t)j * K2 + k0 + kk);
                }
                __syncthreads();
#pragma unroll 4
                for (int kk = 0; kk < 32; ++kk) {
                    float a[4];
#pragma unroll
                    for (int p = 0; p < 4; ++p) a[p] = Y[(4 * w + p) * YS + k0 + kk];
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        const float wv = Ws[kk * HS + lane + 32 * r];
#pragma unroll
                        for (int p = 0; p < 4; ++p) acc[p][r] = fmaf(a[p], wv, acc[p][r]);
                    }
                }
            }
            // ---- phase 2: softmax / blend backward per pixel, d hidden -> Hs
#pragma unroll
            for (int p = 0; p < 4; ++p) {
                const int row = 4 * w + p;
                const int64_t m = m0 + row;
                const bool live = m < M;
                float d0 = 0.f, d1 = 0.f;                     // dL/dw0, dL/dw1
                for (int c = lane; c < C; c += 32) {
                    const float g = G[row * C + c];
                    d0 = fmaf(g, Y[row * YS + c], d0);
                    d1 = fmaf(g, Y[row * YS + C + c], d1);
                }
                d0 = warp_sum(d0); d1 = warp_sum(d1);
                const float w0 = live ? __ldg(attn + 2 * m) : 0.f, w1v = live ? __ldg(attn + 2 * m + 1) : 0.f;
                const float dot = w0 * d0 + w1v * d1;
                const float da0 = w0 * (d0 - dot), da1 = w1v * (d1 - dot);     // softmax backward
                if (lane == 0) { gB2a += da0; gB2b += da1; }
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const int j = lane + 32 * r;
                    const float hpre = acc[p][r] + __ldg(b1 + j);
                    const float h = fmaxf(hpre, 0.f);
                    gW2a[r] = fmaf(da0, h, gW2a[r]);
                    gW2b[r] = fmaf(da1, h, gW2b[r]);
                    const float dh = (hpre > 0.f) ? (da0 * __ldg(w2 + j) + da1 * __ldg(w2 + C + j)) : 0.f;
                    gB1[r] += dh;
                    Hs[row * HS + j] = dh;
                }
            }
        }
        // ---- phase 3: dY = dH . W1 (+ blend path), ReLU mask, BN-affine backward
        {
            float acc2[4][R2];
#pragma unroll
            for (int p = 0; p < 4; ++p)
#pragma unroll
                for (int r = 0; r < R2; ++r) acc2[p][r] = 0.f;
            for (int j0 = 0; j0 < C; j0 += 32) {
                __syncthreads();
                for (int idx = tid; idx < 32 * K2; idx += FW_THREADS)
                    Ws[idx] = __ldg(w1 + (int64_t)j0 * K2 + idx);           // rows j0..j0+31, contiguous
                __syncthreads();
#pragma unroll 4
                for (int jj = 0; jj < 32; ++jj) {
                    float d[4];
#pragma unroll
                    for (int p = 0; p < 4; ++p) d[p] = Hs[(4 * w + p) * HS + j0 + jj];
#pragma unroll
                    for (int r = 0; r < R2; ++r) {
                        const float wv = Ws[jj * K2 + lane + 32 * r];
#pragma unroll
                        for (int p = 0; p < 4; ++p) acc2[p][r] = fmaf(d[p], wv, acc2[p][r]);
                    }
                }
            }
#pragma unroll
            for (int p = 0; p < 4; ++p) {
                const int row = 4 * w + p;
                const int64_t m = m0 + row;
                if (m >= M) continue;
                const float w0 = __ldg(attn + 2 * m), w1v = __ldg(attn + 2 * m + 1);
#pragma unroll
                for (int r = 0; r < R2; ++r) {
                    const int k = lane + 32 * r;
                    const bool is_cam = k < C;
                    const int c = is_cam ? k : k - C;
                    const float y = Y[row * YS + k];
                    const float gy = (y > 0.f) ? acc2[p][r] + G[row * C + c] * (is_cam ? w0 : w1v) : 0.f;
                    const float sc = __ldg((is_cam ? csc : lsc) + c);
                    const float x = to_float<T>((is_cam ? cam : lid)[m * C + c]);
                    gsc[r] = fmaf(gy, x, gsc[r]);
                    gsh[r] += gy;
                    (is_cam ? gcam : glid)[m * C + c] = from_float<T>(gy * sc);
                }
            }
        }
        // ---- phase 4: dW1[j][k] += sum_p dH[p][j] * Y[p][k]
#pragma unroll 2
        for (int p = 0; p < FW_TM; ++p) {
            float dh[NJ], yv[NK];
#pragma unroll
            for (int a = 0; a < NJ; ++a) dh[a] = Hs[p * HS + jb + JB * a];
#pragma unroll
            for (int b = 0; b < NK; ++b) yv[b] = Y[p * YS + kb + KB * b];
#pragma unroll
            for (int a = 0; a < NJ; ++a)
#pragma unroll
                for (int b = 0; b < NK; ++b) dW1[a][b] = fmaf(dh[a], yv[b], dW1[a][b]);
        }
    }

    // ---- flush the per-CTA accumulators
#pragma unroll
    for (int a = 0; a < NJ; ++a)
#pragma unroll
        for (int b = 0; b < NK; ++b)
            atomicAdd(gw1 + (int64_t)(jb + JB * a) * K2 + kb + KB * b, dW1[a][b]);
    // per-warp values: reduce across the 8 warps through smem, then one atomic each
    __syncthreads();
    float *red = smem;                               // [8 warps][32 lanes][slots]
    constexpr int SLOTS = 2 * R2 + 3 * R;
#pragma unroll
    for (int r = 0; r < R2; ++r) { red[(w * 32 + lane) * SLOTS + r] = gsc[r]; red[(w * 32 + lane) * SLOTS + R2 + r] = gsh[r]; }
#pragma unroll
    for (int r = 0; r < R; ++r) {
        red[(w * 32 + lane) * SLOTS + 2 * R2 + r] = gW2a[r];
        red[(w * 32 + lane) * SLOTS + 2 * R2 + R + r] = gW2b[r];
        red[(w * 32 + lane) * SLOTS + 2 * R2 + 2 * R + r] = gB1[r];
    }
    __syncthreads();
    if (w == 0) {
        for (int s = 0; s < SLOTS; ++s) {
            float v = 0.f;
            for (int ww = 0; ww < 8; ++ww) v += red[(ww * 32 + lane) * SLOTS + s];
            if (s < R2) {                                           // d scale
                const int k = lane + 32 * s;
                atomicAdd(gaff + (k < C ? 0 : 2 * C) + (k < C ? k : k - C), v);
            } else if (s < 2 * R2) {                                // d shift
                const int k = lane + 32 * (s - R2);
                atomicAdd(gaff + (k < C ? C : 3 * C) + (k < C ? k : k - C), v);
            } else if (s < 2 * R2 + R) {
                atomicAdd(gw2 + lane + 32 * (s - 2 * R2), v);
            } else if (s < 2 * R2 + 2 * R) {
                atomicAdd(gw2 + C + lane + 32 * (s - 2 * R2 - R), v);
            } else {
                atomicAdd(gb1 + lane + 32 * (s - 2 * R2 - 2 * R), v);
            }
        }
    }
    // gB2 lives in lane 0 of each warp
    __syncthreads();
    if (lane == 0) { red[w * 2] = gB2a; red[w * 2 + 1] = gB2b; }
    __syncthreads();
    if (tid < 2) {
        float v = 0.f;
        for (int ww = 0; ww < 8; ++ww) v += red[ww * 2 + tid];
        atomicAdd(gb2 + tid, v);
    }
}

static size_t fw_fwd_smem(int C) { return sizeof(float) * ((size_t)FW_TM * (2 * C + 4) + 32 * (size_t)(C + 1)); }
static size_t fw_bwd_smem(int C) {
    const size_t ws = (size_t)32 * 2 * C > (size_t)32 * (C + 1) ? (size_t)32 * 2 * C : (size_t)32 * (C + 1);
    size_t fl = (size_t)FW_TM * (2 * C + 4) + (size_t)FW_TM * C + (size_t)FW_TM * (C + 1) + ws;
    const size_t red = (size_t)256 * (2 * (2 * C / 32) + 3 * (C / 32));
    if (red > fl) fl = red;
    return sizeof(float) * fl;
}

static int check_pair_args(const void *a, const void *b, int dtype, int64_t M, int C, const char *who) {
    KDF_CHECK_ARG(M >= 0, "%s: negative M", who);
    KDF_CHECK_ARG(dtype == KDF_F32 || dtype == KDF_BF16, "%s: bad dtype %d", who, dtype);
    KDF_CHECK_ARG(a && b, "%s: null pointer", who);
    KDF_CHECK_ARG(C >= 4 && C % 4 == 0 && 256 % (C / 4) == 0, "%s: C=%d must be 4*2^k <= 1024", who, C);
    return KDF_OK;
}

}  // namespace kdf

using namespace kdf;

extern "C" {

int kdf_fusion_affine_relu_pair_fwd(const void *cam_pre, const void *lid_pre, int dtype, int64_t M, int C,
                                    const float *cam_scale, const float *cam_shift,
                                    const float *lid_scale, const float *lid_shift,
                                    int mode, void *out, void *stream) {
    if (int e = check_pair_args(cam_pre, lid_pre, dtype, M, C, "affine_relu_pair_fwd")) return e;
    KDF_CHECK_ARG(cam_scale && cam_shift && lid_scale && lid_shift && out, "affine_relu_pair_fwd: null pointer");
    KDF_CHECK_ARG(mode == 0 || mode == 1, "affine_relu_pair_fwd: bad mode");
    if (M == 0) return KDF_OK;
    const int rows = 256 / (C / 4);
    int64_t blocks = (M + rows - 1) / rows;
    if (blocks > sm_count() * 16) blocks = sm_count() * 16;
    cudaStream_t st = as_stream(stream);
    if (dtype == KDF_F32)
        affine_relu_pair_fwd_kernel<float><<<(int)blocks, 256, 0, st>>>(
            (const float *)cam_pre, (const float *)lid_pre, M, C, cam_scale, cam_shift, lid_scale, lid_shift, mode, (float *)out);
    else
        affine_relu_pair_fwd_kernel<__nv_bfloat16><<<(int)blocks, 256, 0, st>>>(
            (const __nv_bfloat16 *)cam_pre, (const __nv_bfloat16 *)lid_pre, M, C, cam_scale, cam_shift, lid_scale, lid_shift,
            mode, (__nv_bfloat16 *)out);
    KDF_LAUNCH_CHECK();
    return KDF_OK;
}

int kdf_fusion_affine_relu_pair_bwd(const void *grad_out, const void *cam_pre, const void *lid_pre,
                                    int dtype, int64_t M, int C,
                                    const float *cam_scale, const float *cam_shift,
                                    const float *lid_scale, const float *lid_shift, int mode,
                                    void *grad_cam_pre, void *grad_lid_pre, float *grad_affine, void *stream) {
    if (int e = check_pair_args(cam_pre, lid_pre, dtype, M, C, "affine_relu_pair_bwd")) return e;
    KDF_CHECK_ARG(grad_out && cam_scale && cam_shift && lid_scale && lid_shift && grad_cam_pre && grad_lid_pre && grad_affine,
                  "affine_relu_pair_bwd: null pointer");
    KDF_CHECK_ARG(mode == 0 || mode == 1, "affine_relu_pair_bwd: bad mode");
    cudaStream_t st = as_stream(stream);
    KDF_CUDA(cudaMemsetAsync(grad_affine, 0, sizeof(float) * 4 * C, st));
    if (M == 0) return KDF_OK;
    const int rows = 256 / (C / 4);
    int64_t blocks = (M + rows - 1) / rows;
    if (blocks > sm_count() * 4) blocks = sm_count() * 4;
    const size_t smem = sizeof(float) * 256 * 16;
    if (dtype == KDF_F32)
        affine_relu_pair_bwd_kernel<float><<<(int)blocks, 256, smem, st>>>(
            (const float *)grad_out, (const float *)cam_pre, (const float *)lid_pre, M, C, cam_scale, cam_shift, lid_scale,
            lid_shift, mode, (float *)grad_cam_pre, (float *)grad_lid_pre, grad_affine);
    else
        affine_relu_pair_bwd_kernel<__nv_bfloat16><<<(int)blocks, 256, smem, st>>>(
            (const __nv_bfloat16 *)grad_out, (const __nv_bfloat16 *)cam_pre, (const __nv_bfloat16 *)lid_pre, M, C, cam_scale,
            cam_shift, lid_scale, lid_shift, mode, (__nv_bfloat16 *)grad_cam_pre, (__nv_bfloat16 *)grad_lid_pre, grad_affine);
    KDF_LAUNCH_CHECK();
    return KDF_OK;
}

int kdf_fusion_weighted_fwd(const void *cam_pre, const void *lid_pre, int dtype, int64_t M, int C,
                            const float *cam_scale, const float *cam_shift,
                            const float *lid_scale, const float *lid_shift,
                            const float *w1, const float *b1, const float *w2, const float *b2,
                            void *out, float *attn, void *stream) {
    KDF_CHECK_ARG(M >= 0, "fusion_weighted_fwd: negative M");
    KDF_CHECK_ARG(dtype == KDF_F32 || dtype == KDF_BF16, "fusion_weighted_fwd: bad dtype %d", dtype);
    KDF_CHECK_ARG(C >= 32 && C % 32 == 0 && C <= 256, "fusion_weighted_fwd: C=%d must be a multiple of 32 in [32,256]", C);
    KDF_CHECK_ARG(cam_pre && lid_pre && cam_scale && cam_shift && lid_scale && lid_shift && w1 && b1 && w2 && b2 && out && attn,
                  "fusion_weighted_fwd: null pointer");
    if (M == 0) return KDF_OK;
    cudaStream_t st = as_stream(stream);
    if (dtype == KDF_BF16 && C == 128 && fusion_tc_enabled())
        return fusion_weighted_fwd_tc(cam_pre, lid_pre, M, cam_scale, cam_shift, lid_scale, lid_shift, w1, b1, w2, b2, out, attn, st);
    const size_t smem = fw_fwd_smem(C);
    int64_t blocks = (M + FW_TM - 1) / FW_TM;
    if (blocks > sm_count() * 4) blocks = sm_count() * 4;
    if (dtype == KDF_F32) {
        KDF_CUDA(cudaFuncSetAttribute(fusion_weighted_fwd_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        fusion_weighted_fwd_kernel<float><<<(int)blocks, FW_THREADS, smem, st>>>(
            (const float *)cam_pre, (const float *)lid_pre, M, C, cam_scale, cam_shift, lid_scale, lid_shift, w1, b1, w2, b2,
            (float *)out, attn);
    } else {
        KDF_CUDA(cudaFuncSetAttribute(fusion_weighted_fwd_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        fusion_weighted_fwd_kernel<__nv_bfloat16><<<(int)blocks, FW_THREADS, smem, st>>>(
            (const __nv_bfloat16 *)cam_pre, (const __nv_bfloat16 *)lid_pre, M, C, cam_scale, cam_shift, lid_scale, lid_shift, w1,
            b1, w2, b2, (__nv_bfloat16 *)out, attn);
    }
    KDF_LAUNCH_CHECK();
    return KDF_OK;
}

int kdf_fusion_weighted_bwd(const void *grad_out, const void *cam_pre, const void *lid_pre,
                            int dtype, int64_t M, int C,
                            const float *cam_scale, const float *cam_shift,
                            const float *lid_scale, const float *lid_shift,
                            const float *w1, const float *b1, const float *w2, const float *b2,
                            const float *attn,
                            void *grad_cam_pre, void *grad_lid_pre, float *grad_affine,
                            float *grad_w1, float *grad_b1, float *grad_w2, float *grad_b2,
                            void *stream) {
    (void)b2;
    KDF_CHECK_ARG(M >= 0, "fusion_weighted_bwd: negative M");
    KDF_CHECK_ARG(dtype == KDF_F32 || dtype == KDF_BF16, "fusion_weighted_bwd: bad dtype %d", dtype);
    KDF_CHECK_ARG(C == 64 || C == 128, "fusion_weighted_bwd: C=%d not instantiated (64, 128)", C);
    KDF_CHECK_ARG(grad_out && cam_pre && lid_pre && cam_scale && cam_shift && lid_scale && lid_shift && w1 && b1 && w2 && attn &&
                  grad_cam_pre && grad_lid_pre && grad_affine && grad_w1 && grad_b1 && grad_w2 && grad_b2,
                  "fusion_weighted_bwd: null pointer");
    cudaStream_t st = as_stream(stream);
    KDF_CUDA(cudaMemsetAsync(grad_affine, 0, sizeof(float) * 4 * C, st));
    KDF_CUDA(cudaMemsetAsync(grad_w1, 0, sizeof(float) * 2 * C * C, st));
    KDF_CUDA(cudaMemsetAsync(grad_b1, 0, sizeof(float) * C, st));
    KDF_CUDA(cudaMemsetAsync(grad_w2, 0, sizeof(float) * 2 * C, st));
    KDF_CUDA(cudaMemsetAsync(grad_b2, 0, sizeof(float) * 2, st));
    if (M == 0) return KDF_OK;
    if (dtype == KDF_BF16 && C == 128 && fusion_tc_enabled())
        return fusion_weighted_bwd_tc(grad_out, cam_pre, lid_pre, M, cam_scale, cam_shift, lid_scale, lid_shift, w1, b1, w2, attn,
                                      grad_cam_pre, grad_lid_pre, grad_affine, grad_w1, grad_b1, grad_w2, grad_b2, st);
    const size_t smem = fw_bwd_smem(C);
    int64_t blocks = (M + FW_TM - 1) / FW_TM;
    if (blocks > sm_count()) blocks = sm_count();
#define KDF_FWB(T, CC)                                                                                              \
    do {                                                                                                            \
        KDF_CUDA(cudaFuncSetAttribute(fusion_weighted_bwd_kernel<T, CC>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                      (int)smem));                                                                  \
        fusion_weighted_bwd_kernel<T, CC><<<(int)blocks, FW_THREADS, smem, st>>>(                                   \
            (const T *)grad_out, (const T *)cam_pre, (const T *)lid_pre, M, cam_scale, cam_shift, lid_scale,         \
            lid_shift, w1, b1, w2, attn, (T *)grad_cam_pre, (T *)grad_lid_pre, grad_affine, grad_w1, grad_b1,        \
            grad_w2, grad_b2);                                                                                      \
    } while (0)
    if (dtype == KDF_F32) { if (C == 128) KDF_FWB(float, 128); else KDF_FWB(float, 64); }
    else                  { if (C == 128) KDF_FWB(__nv_bfloat16, 128); else KDF_FWB(__nv_bfloat16, 64); }
#undef KDF_FWB
    KDF_LAUNCH_CHECK();
    return KDF_OK;
}

}  // extern "C"
