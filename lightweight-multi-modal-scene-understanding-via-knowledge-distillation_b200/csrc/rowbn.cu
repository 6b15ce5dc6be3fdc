// BatchNorm (+ReLU / ReLU6, + residual) over pixel- or point-major rows [M, C]  (sm_100a).
//
// Every BatchNorm of the reference -- BatchNorm1d in the point MLP
// (src/models/lidar_encoder.py:25-35), BatchNorm2d in the camera encoder, FPN, fusion and
// heads (camera_encoder.py:19-41, fusion_module.py:11-32) -- sees its input here as rows
// of a channels-last tensor: M = B*N points or B*H*W pixels, C channels contiguous.
// Training mode needs the batch statistics before anything can be applied, so forward and
// backward are two streaming passes each, and nothing else:
//
//   forward : stats  (read x)            -> mean/invstd, scale/shift, running-stat update
//             apply  (read x, write y)   y = act(x*scale + shift) [+ residual]
//   backward: reduce (read g, x)         -> S0 = sum dy, S1 = sum dy*x  => d gamma, d beta,
//                                           per-channel (A, B) of the statistics chain
//             apply  (read g, x, write)  dx = dy*scale + B*x + A
//
// Both reductions are per-thread fp32 partial sums over a few hundred rows, reduced per CTA
// in shared memory and accumulated across CTAs with fp64 atomics (2C adds per CTA; in fp64 the
// accumulation order perturbs the result far below fp32 resolution), finalised by the last
// CTA to arrive.  All four
// kernels are HBM-bound streaming kernels; the per-channel finalisation lives in the last
// CTA of the reduction so no tiny host-driven launches are needed.
#include <cooperative_groups.h>
#include <stdlib.h>

#include "kdf_common.cuh"

namespace kdf {

constexpr int RB_THREADS = 256;
constexpr int RB_MAX_BLOCKS = 1184;          // 148 SMs x 8 CTAs
constexpr int RB_COPIES = 8;                 // accumulator sets of the column reductions

struct RowMap {           // thread -> (row lane r, column group g) for rows of C = G*VEC channels
    int G, rows, r, g;
    bool active;
};
template <int VEC>
__device__ __forceinline__ RowMap row_map(int C) {
    RowMap m;
    m.G = C / VEC;
    m.rows = RB_THREADS / m.G;               // >= 1 because C <= 256*VEC is checked on the host
    m.active = threadIdx.x < m.rows * m.G;
    m.r = threadIdx.x / m.G;
    m.g = threadIdx.x - m.r * m.G;
    return m;
}

// Rows are moved as raw 16-byte (8-byte for bf16 x4) registers and converted to fp32 only when
// consumed, so U loads in flight cost U*4 registers instead of U*VEC.
template <typename T, int VEC> struct VecIO;
template <> struct VecIO<float, 4> {
    using Raw = uint4;
    static __device__ __forceinline__ Raw load(const float *p) { return ldg_stream_u4(reinterpret_cast<const uint4 *>(p)); }
    static __device__ __forceinline__ void unpack(const Raw &u, float *v) {
        v[0] = __uint_as_float(u.x); v[1] = __uint_as_float(u.y); v[2] = __uint_as_float(u.z); v[3] = __uint_as_float(u.w);
    }
    static __device__ __forceinline__ void store(float *p, const float *v) {
        *reinterpret_cast<float4 *>(p) = make_float4(v[0], v[1], v[2], v[3]);
    }
};
template <> struct VecIO<__nv_bfloat16, 4> {
    using Raw = uint2;
    static __device__ __forceinline__ Raw load(const __nv_bfloat16 *p) { return ldg_stream_u2(reinterpret_cast<const uint2 *>(p)); }
    static __device__ __forceinline__ void unpack(const Raw &u, float *v) {
        v[0] = bf16_lo(u.x); v[1] = bf16_hi(u.x); v[2] = bf16_lo(u.y); v[3] = bf16_hi(u.y);
    }
    static __device__ __forceinline__ void store(__nv_bfloat16 *p, const float *v) {
        uint2 u;
        u.x = pack_bf16(v[0], v[1]); u.y = pack_bf16(v[2], v[3]);
        *reinterpret_cast<uint2 *>(p) = u;
    }
};
template <> struct VecIO<__nv_bfloat16, 8> {
    using Raw = uint4;
    static __device__ __forceinline__ Raw load(const __nv_bfloat16 *p) { return ldg_stream_u4(reinterpret_cast<const uint4 *>(p)); }
    static __device__ __forceinline__ void unpack(const Raw &u, float *v) {
        v[0] = bf16_lo(u.x); v[1] = bf16_hi(u.x); v[2] = bf16_lo(u.y); v[3] = bf16_hi(u.y);
        v[4] = bf16_lo(u.z); v[5] = bf16_hi(u.z); v[6] = bf16_lo(u.w); v[7] = bf16_hi(u.w);
    }
    static __device__ __forceinline__ void store(__nv_bfloat16 *p, const float *v) {
        uint4 u;
        u.x = pack_bf16(v[0], v[1]); u.y = pack_bf16(v[2], v[3]);
        u.z = pack_bf16(v[4], v[5]); u.w = pack_bf16(v[6], v[7]);
        *reinterpret_cast<uint4 *>(p) = u;
    }
};

// VEC per-channel coefficients of this thread's channel group from shared memory as 16-byte loads (the group starts at a
// multiple of VEC floats, so it is 16-byte aligned; left to the compiler these are VEC 4-byte loads per table)
template <int VEC>
__device__ __forceinline__ void ld_coef(const float *p, float (&v)[VEC]) {
#pragma unroll
    for (int q = 0; q < VEC; q += 4) {
        const float4 t = *reinterpret_cast<const float4 *>(p + q);
        v[q] = t.x; v[q + 1] = t.y; v[q + 2] = t.z; v[q + 3] = t.w;
    }
}

__device__ __forceinline__ float act_fwd(float y, int act) {
    if (act == 1) return fmaxf(y, 0.f);
    if (act == 2) return fminf(fmaxf(y, 0.f), 6.f);
    return y;
}
__device__ __forceinline__ bool act_open(float y, int act) {      // derivative is 1 (else 0)
    if (act == 1) return y > 0.f;
    if (act == 2) return y > 0.f && y < 6.f;
    return true;
}

struct RowBnWs {            // workspace header followed by acc[RB_COPIES][2][C] (fp64)
    unsigned int ticket;
    unsigned int pad[3];
};

// Two per-channel sums over the rows; `MODE` 0: (x, x^2)   1: (dy, dy*x) with dy = g * act'(x*scale+shift).
// The last CTA to arrive finalises from the fp64 accumulators.
struct ReduceArgs {
    const void *x, *g;
    int64_t M;
    int C, act;
    const float *scale, *shift;      // MODE 1
    // MODE 0 finalisation (training statistics)
    const float *gamma, *beta, *pre_bias;
    float eps, momentum;
    float *mean, *invstd, *out_scale, *out_shift, *running_mean, *running_var;
    // MODE 1 finalisation
    const float *in_mean, *in_invstd;
    int batch_stats;
    float *dgamma, *dbeta, *coefA, *coefB;
    RowBnWs *ws;
};

template <typename T, int VEC, int MODE>
__global__ void __launch_bounds__(RB_THREADS, (VEC == 8 ? 3 : 4))
rowbn_reduce_kernel(ReduceArgs a) {
    extern __shared__ float sm[];                     // [rows][G][2*VEC]
    __shared__ bool last;
    const RowMap m = row_map<VEC>(a.C);
    const T *x = reinterpret_cast<const T *>(a.x);
    const T *g = reinterpret_cast<const T *>(a.g);
    float s0[VEC], s1[VEC], sc[VEC], sh[VEC];
#pragma unroll
    for (int q = 0; q < VEC; ++q) { s0[q] = 0.f; s1[q] = 0.f; sc[q] = 1.f; sh[q] = 0.f; }
    if (m.active) {
        const int c = m.g * VEC;
        if (MODE == 1) {
#pragma unroll
            for (int q = 0; q < VEC; ++q) { sc[q] = a.scale[c + q]; sh[q] = a.shift[c + q]; }
        }
        const int64_t stride = (int64_t)gridDim.x * m.rows;
        constexpr int U = 4;
        using IO = VecIO<T, VEC>;
        for (int64_t row = (int64_t)blockIdx.x * m.rows + m.r; row < a.M; row += stride * U) {
            typename IO::Raw xr[U], gr[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int64_t rr = row + u * stride;
                if (rr < a.M) {
                    xr[u] = IO::load(x + rr * a.C + c);
                    if (MODE == 1) gr[u] = IO::load(g + rr * a.C + c);
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (row + u * stride < a.M) {
                    float xv[VEC], gv[VEC];
                    IO::unpack(xr[u], xv);
                    if (MODE == 1) IO::unpack(gr[u], gv);
#pragma unroll
                    for (int q = 0; q < VEC; ++q) {
                        if (MODE == 0) {
                            s0[q] += xv[q];
                            s1[q] = fmaf(xv[q], xv[q], s1[q]);
                        } else {
                            const float dy = act_open(fmaf(xv[q], sc[q], sh[q]), a.act) ? gv[q] : 0.f;
                            s0[q] += dy;
                            s1[q] = fmaf(dy, xv[q], s1[q]);
                        }
                    }
                }
            }
        }
        float *dst = sm + (m.r * m.G + m.g) * 2 * VEC;
#pragma unroll
        for (int q = 0; q < VEC; ++q) { dst[q] = s0[q]; dst[VEC + q] = s1[q]; }
    }
    __syncthreads();
    // RB_COPIES accumulator sets spread the same-address fp64 atomics (they serialise in L2); the last CTA adds them up
    double *acc0 = reinterpret_cast<double *>(a.ws + 1);
    double *acc = acc0 + (size_t)(blockIdx.x % RB_COPIES) * 2 * a.C;
    if (m.active && m.r == 0) {
        for (int rr = 1; rr < m.rows; ++rr) {
            const float *src = sm + (rr * m.G + m.g) * 2 * VEC;
#pragma unroll
            for (int q = 0; q < VEC; ++q) { s0[q] += src[q]; s1[q] += src[VEC + q]; }
        }
#pragma unroll
        for (int q = 0; q < VEC; ++q) {
            atomicAdd(acc + m.g * VEC + q, (double)s0[q]);
            atomicAdd(acc + a.C + m.g * VEC + q, (double)s1[q]);
        }
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) last = (atomicAdd(&a.ws->ticket, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!last) return;
    __threadfence();
    const double invM = 1.0 / (double)a.M;
    for (int c = threadIdx.x; c < a.C; c += RB_THREADS) {
        double t0 = 0.0, t1 = 0.0;
#pragma unroll
        for (int k = 0; k < RB_COPIES; ++k) {
            t0 += __ldcg(acc0 + (size_t)k * 2 * a.C + c);
            t1 += __ldcg(acc0 + (size_t)k * 2 * a.C + a.C + c);
        }
        if (MODE == 0) {
            const double mean = t0 * invM;
            double var = t1 * invM - mean * mean;                 // biased variance (normalisation)
            if (var < 0.0) var = 0.0;
            const float invstd = (float)(1.0 / sqrt(var + (double)a.eps));
            a.mean[c] = (float)mean;
            a.invstd[c] = invstd;
            const float scl = (a.gamma ? a.gamma[c] : 1.f) * invstd;
            a.out_scale[c] = scl;
            a.out_shift[c] = (a.beta ? a.beta[c] : 0.f) - (float)mean * scl;
            if (a.running_mean) {                                 // nn.BatchNorm: unbiased variance in the running stat
                const double unb = a.M > 1 ? var * ((double)a.M / (double)(a.M - 1)) : var;
                const float bias = a.pre_bias ? a.pre_bias[c] : 0.f;      // BN(x + bias): only the running mean sees it
                a.running_mean[c] = (1.f - a.momentum) * a.running_mean[c] + a.momentum * ((float)mean + bias);
                a.running_var[c] = (1.f - a.momentum) * a.running_var[c] + a.momentum * (float)unb;
            }
        } else {
            const float mean = a.in_mean[c], invstd = a.in_invstd[c], scl = a.scale[c];
            const float S0 = (float)t0, S1 = (float)t1;
            const float dgamma = invstd * (S1 - mean * S0);       // sum dy * xhat
            if (a.dgamma) a.dgamma[c] = dgamma;
            if (a.dbeta) a.dbeta[c] = S0;
            float A = 0.f, B = 0.f;
            if (a.batch_stats) {                                  // chain through batch mean / variance
                B = -(scl * invstd * dgamma) * (float)invM;
                A = -(scl * S0) * (float)invM - B * mean;
            }
            a.coefA[c] = A;
            a.coefB[c] = B;
        }
    }
}

// y = act(x*scale + shift) [+ residual]   (activation as a template parameter: no per-element selects of the other cases)
template <int ACT>
__device__ __forceinline__ float act_fwd_t(float y) {
    if (ACT == 1) return fmaxf(y, 0.f);
    if (ACT == 2) return fminf(fmaxf(y, 0.f), 6.f);
    return y;
}
template <typename T, int VEC, int ACT>
__global__ void __launch_bounds__(RB_THREADS, (VEC == 8 ? 3 : 4))
rowbn_apply_fwd_kernel(const T *__restrict__ x, const T *__restrict__ res, T *__restrict__ y, int64_t M, int C,
                       const float *__restrict__ scale, const float *__restrict__ shift) {
    const RowMap m = row_map<VEC>(C);
    if (!m.active) return;
    const int c = m.g * VEC;
    float sc[VEC], sh[VEC];
#pragma unroll
    for (int q = 0; q < VEC; ++q) { sc[q] = scale[c + q]; sh[q] = shift[c + q]; }
    const int64_t stride = (int64_t)gridDim.x * m.rows;
    constexpr int U = 4;
    using IO = VecIO<T, VEC>;
    for (int64_t row = (int64_t)blockIdx.x * m.rows + m.r; row < M; row += stride * U) {
        typename IO::Raw xr[U], rr_[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t rr = row + u * stride;
            if (rr < M) {
                xr[u] = IO::load(x + rr * C + c);
                if (res) rr_[u] = IO::load(res + rr * C + c);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t rr = row + u * stride;
            if (rr < M) {
                float xv[VEC], rv[VEC], o[VEC];
                IO::unpack(xr[u], xv);
                if (res) IO::unpack(rr_[u], rv);
#pragma unroll
                for (int q = 0; q < VEC; ++q) {
                    o[q] = act_fwd_t<ACT>(fmaf(xv[q], sc[q], sh[q]));
                    if (res) o[q] += rv[q];
                }
                IO::store(y + rr * C + c, o);
            }
        }
    }
}

// dx = dy*scale + B*x + A,  dy = g * act'(x*scale + shift)
template <typename T, int VEC>
__global__ void __launch_bounds__(RB_THREADS, (VEC == 8 ? 3 : 4))
rowbn_apply_bwd_kernel(const T *__restrict__ g, const T *__restrict__ x, T *__restrict__ dx, int64_t M, int C,
                       const float *__restrict__ scale, const float *__restrict__ shift,
                       const float *__restrict__ coefA, const float *__restrict__ coefB, int act) {
    // per-channel coefficients live in shared memory (4*C floats) to keep registers for loads in flight
    extern __shared__ float coef[];
    float *sc = coef, *sh = coef + C, *A = coef + 2 * C, *B = coef + 3 * C;
    for (int i = threadIdx.x; i < C; i += RB_THREADS) {
        sc[i] = scale[i]; sh[i] = shift[i]; A[i] = coefA[i]; B[i] = coefB[i];
    }
    __syncthreads();
    const RowMap m = row_map<VEC>(C);
    if (!m.active) return;
    const int c = m.g * VEC;
    sc += c; sh += c; A += c; B += c;
    const int64_t stride = (int64_t)gridDim.x * m.rows;
    constexpr int U = 4;
    using IO = VecIO<T, VEC>;
    for (int64_t row = (int64_t)blockIdx.x * m.rows + m.r; row < M; row += stride * U) {
        typename IO::Raw xr[U], gr[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t rr = row + u * stride;
            if (rr < M) {
                xr[u] = IO::load(x + rr * C + c);
                gr[u] = IO::load(g + rr * C + c);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t rr = row + u * stride;
            if (rr < M) {
                float xv[VEC], gv[VEC], o[VEC], vsc[VEC], vsh[VEC], vA[VEC], vB[VEC];
                IO::unpack(xr[u], xv);
                IO::unpack(gr[u], gv);
                ld_coef<VEC>(sc, vsc); ld_coef<VEC>(sh, vsh); ld_coef<VEC>(A, vA); ld_coef<VEC>(B, vB);
#pragma unroll
                for (int q = 0; q < VEC; ++q) {
                    const float dy = act_open(fmaf(xv[q], vsc[q], vsh[q]), act) ? gv[q] : 0.f;
                    o[q] = fmaf(dy, vsc[q], fmaf(vB[q], xv[q], vA[q]));
                }
                IO::store(dx + rr * C + c, o);
            }
        }
    }
}

// BatchNorm backward as ONE cooperative kernel for tensors that stay in L2 between the two passes: column reduction,
// grid-wide barrier, per-channel coefficients (every CTA, from the accumulator sets), apply.  Same arithmetic as
// rowbn_reduce_kernel<MODE 1> + rowbn_apply_bwd_kernel; what it saves is a launch, the last-CTA finalisation round
// trip and the second kernel's ramp -- ~10 us of the ~40 us such a pair costs on a 30 MB tensor.
// (the activation is a template parameter here: with a run-time `act` both loops carried the selects of all three cases)
template <int ACT>
__device__ __forceinline__ bool act_open_t(float y) {
    if (ACT == 1) return y > 0.f;
    if (ACT == 2) return y > 0.f && y < 6.f;
    return true;
}

template <typename T, int VEC, int ACT>
__global__ void __launch_bounds__(RB_THREADS, (VEC == 8 ? 3 : 4))
rowbn_bwd_coop_kernel(ReduceArgs a, T *__restrict__ dx) {
    extern __shared__ float sm[];                     // phase 1: [rows][G][2*VEC]; phase 2: scale, shift, A, B [4][C]
    const RowMap m = row_map<VEC>(a.C);
    const T *x = reinterpret_cast<const T *>(a.x);
    const T *g = reinterpret_cast<const T *>(a.g);
    const int c = m.g * VEC;
    const int64_t stride = (int64_t)gridDim.x * m.rows;
    constexpr int U = 4;
    using IO = VecIO<T, VEC>;
    {
        float s0[VEC], s1[VEC], sc[VEC], sh[VEC];
#pragma unroll
        for (int q = 0; q < VEC; ++q) { s0[q] = 0.f; s1[q] = 0.f; sc[q] = 1.f; sh[q] = 0.f; }
        if (m.active) {
#pragma unroll
            for (int q = 0; q < VEC; ++q) { sc[q] = a.scale[c + q]; sh[q] = a.shift[c + q]; }
            for (int64_t row = (int64_t)blockIdx.x * m.rows + m.r; row < a.M; row += stride * U) {
                typename IO::Raw xr[U], gr[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int64_t rr = row + u * stride;
                    if (rr < a.M) {                              // L2-allocating loads: phase 2 reads the same rows again
                        xr[u] = *reinterpret_cast<const typename IO::Raw *>(x + rr * a.C + c);
                        gr[u] = *reinterpret_cast<const typename IO::Raw *>(g + rr * a.C + c);
                    }
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    if (row + u * stride < a.M) {
                        float xv[VEC], gv[VEC];
                        IO::unpack(xr[u], xv);
                        IO::unpack(gr[u], gv);
#pragma unroll
                        for (int q = 0; q < VEC; ++q) {
                            const float dy = act_open_t<ACT>(fmaf(xv[q], sc[q], sh[q])) ? gv[q] : 0.f;
                            s0[q] += dy;
                            s1[q] = fmaf(dy, xv[q], s1[q]);
                        }
                    }
                }
            }
            float *dst = sm + (m.r * m.G + m.g) * 2 * VEC;
#pragma unroll
            for (int q = 0; q < VEC; ++q) { dst[q] = s0[q]; dst[VEC + q] = s1[q]; }
        }
        __syncthreads();
        double *acc0 = reinterpret_cast<double *>(a.ws + 1);
        double *acc = acc0 + (size_t)(blockIdx.x % RB_COPIES) * 2 * a.C;
        if (m.active && m.r == 0) {
            for (int rr = 1; rr < m.rows; ++rr) {
                const float *src = sm + (rr * m.G + m.g) * 2 * VEC;
#pragma unroll
                for (int q = 0; q < VEC; ++q) { s0[q] += src[q]; s1[q] += src[VEC + q]; }
            }
#pragma unroll
            for (int q = 0; q < VEC; ++q) {
                atomicAdd(acc + m.g * VEC + q, (double)s0[q]);
                atomicAdd(acc + a.C + m.g * VEC + q, (double)s1[q]);
            }
        }
    }
    __threadfence();
    cooperative_groups::this_grid().sync();
    // ---- per-channel coefficients (every CTA computes them; CTA 0 also publishes d gamma / d beta)
    float *csc = sm, *csh = sm + a.C, *cA = sm + 2 * a.C, *cB = sm + 3 * a.C;
    {
        const double *acc0 = reinterpret_cast<const double *>(a.ws + 1);
        const double invM = 1.0 / (double)a.M;
        for (int ch = threadIdx.x; ch < a.C; ch += RB_THREADS) {
            double t0 = 0.0, t1 = 0.0;
#pragma unroll
            for (int k = 0; k < RB_COPIES; ++k) {
                t0 += __ldcg(acc0 + (size_t)k * 2 * a.C + ch);
                t1 += __ldcg(acc0 + (size_t)k * 2 * a.C + a.C + ch);
            }
            const float mean = a.in_mean[ch], invstd = a.in_invstd[ch], scl = a.scale[ch];
            const float S0 = (float)t0, S1 = (float)t1;
            const float dgamma = invstd * (S1 - mean * S0);
            float A = 0.f, B = 0.f;
            if (a.batch_stats) {
                B = -(scl * invstd * dgamma) * (float)invM;
                A = -(scl * S0) * (float)invM - B * mean;
            }
            csc[ch] = scl; csh[ch] = a.shift[ch]; cA[ch] = A; cB[ch] = B;
            if (blockIdx.x == 0) {
                if (a.dgamma) a.dgamma[ch] = dgamma;
                if (a.dbeta) a.dbeta[ch] = S0;
            }
        }
    }
    __syncthreads();
    if (!m.active) return;
    for (int64_t row = (int64_t)blockIdx.x * m.rows + m.r; row < a.M; row += stride * U) {
        typename IO::Raw xr[U], gr[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t rr = row + u * stride;
            if (rr < a.M) {
                xr[u] = IO::load(x + rr * a.C + c);
                gr[u] = IO::load(g + rr * a.C + c);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t rr = row + u * stride;
            if (rr < a.M) {
                float xv[VEC], gv[VEC], o[VEC], vsc[VEC], vsh[VEC], vA[VEC], vB[VEC];
                IO::unpack(xr[u], xv);
                IO::unpack(gr[u], gv);
                ld_coef<VEC>(csc + c, vsc); ld_coef<VEC>(csh + c, vsh); ld_coef<VEC>(cA + c, vA); ld_coef<VEC>(cB + c, vB);
#pragma unroll
                for (int q = 0; q < VEC; ++q) {
                    const float dy = act_open_t<ACT>(fmaf(xv[q], vsc[q], vsh[q])) ? gv[q] : 0.f;
                    o[q] = fmaf(dy, vsc[q], fmaf(vB[q], xv[q], vA[q]));
                }
                IO::store(dx + rr * a.C + c, o);
            }
        }
    }
}

template <typename T, int VEC>
static int launch_bwd_coop(ReduceArgs a, void *dx, int blocks, size_t smem, cudaStream_t st) {
    T *dxp = reinterpret_cast<T *>(dx);
    void *args[] = {&a, &dxp};
    const void *fn = a.act == 1 ? reinterpret_cast<const void *>(&rowbn_bwd_coop_kernel<T, VEC, 1>)
                   : a.act == 2 ? reinterpret_cast<const void *>(&rowbn_bwd_coop_kernel<T, VEC, 2>)
                                : reinterpret_cast<const void *>(&rowbn_bwd_coop_kernel<T, VEC, 0>);
    KDF_CUDA(cudaLaunchCooperativeKernel(fn, dim3(blocks), dim3(RB_THREADS), args, smem, st));
    return KDF_OK;
}

static int rb_vec(int dtype, int C) { return (dtype == KDF_BF16 && C % 8 == 0) ? 8 : 4; }

static int rb_check(int dtype, int64_t M, int C, const char *who) {
    KDF_CHECK_ARG(M >= 0, "%s: negative M", who);
    KDF_CHECK_ARG(dtype == KDF_F32 || dtype == KDF_BF16, "%s: bad dtype %d", who, dtype);
    KDF_CHECK_ARG(C >= 4 && C % 4 == 0, "%s: C=%d must be a positive multiple of 4", who, C);
    KDF_CHECK_ARG(C / rb_vec(dtype, C) <= RB_THREADS, "%s: C=%d too wide", who, C);
    return KDF_OK;
}

static int rb_blocks(int64_t M, int C, int vec, int per_thread_rows, int max_blocks = 0) {
    const int rows = RB_THREADS / (C / vec);

    int64_t b = (M + (int64_t)rows * per_thread_rows - 1) / ((int64_t)rows * per_thread_rows);
    int cap = sm_count() * 8 < RB_MAX_BLOCKS ? sm_count() * 8 : RB_MAX_BLOCKS;
    if (max_blocks > 0 && max_blocks < cap) cap = max_blocks;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (int)b;
}

template <int MODE>
static int launch_reduce(const ReduceArgs &a, int dtype, cudaStream_t st) {
    const int vec = rb_vec(dtype, a.C);
    // every CTA ends with 2C fp64 atomics on the same 2C addresses: the grid is capped (measured: more, smaller CTAs
    // lose more to the serialised atomics than they gain in latency hiding)
    const int blocks = rb_blocks(a.M, a.C, vec, 16, 2 * sm_count());
    const int rows = RB_THREADS / (a.C / vec);
    const size_t smem = sizeof(float) * (size_t)rows * (a.C / vec) * 2 * vec;
    KDF_CUDA(cudaMemsetAsync(a.ws, 0, sizeof(RowBnWs) + sizeof(double) * 2 * RB_COPIES * (size_t)a.C, st));
    if (dtype == KDF_F32) rowbn_reduce_kernel<float, 4, MODE><<<blocks, RB_THREADS, smem, st>>>(a);
    else if (vec == 8)    rowbn_reduce_kernel<__nv_bfloat16, 8, MODE><<<blocks, RB_THREADS, smem, st>>>(a);
    else                  rowbn_reduce_kernel<__nv_bfloat16, 4, MODE><<<blocks, RB_THREADS, smem, st>>>(a);
    KDF_LAUNCH_CHECK();
    return KDF_OK;
}


// g[m,c] += bc[c]*x[m,c] + ac[c] over rows [M,C]: the second half of a BatchNorm backward whose first half (the
// per-row part and the two column sums) was produced by another kernel (the fused fusion backward).
template <typename T>
__global__ void __launch_bounds__(256)
rows_axpb_kernel(T *__restrict__ g, const T *__restrict__ x, const float *__restrict__ bc, const float *__restrict__ ac,
                 int64_t M, int C) {
    constexpr int V = 16 / sizeof(T);
    const int cg = C / V;
    const int64_t total = M * cg, nthreads = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += nthreads) {
        const int c = (int)(i % cg) * V;
        uint4 gv = *(reinterpret_cast<const uint4 *>(g) + i);
        const uint4 xv = ldg_stream_u4(reinterpret_cast<const uint4 *>(x) + i);
        float gf[V], xf[V];
        if (sizeof(T) == 2) {
            gf[0] = bf16_lo(gv.x); gf[1] = bf16_hi(gv.x); gf[2] = bf16_lo(gv.y); gf[3] = bf16_hi(gv.y);
            gf[4 % V] = bf16_lo(gv.z); gf[5 % V] = bf16_hi(gv.z); gf[6 % V] = bf16_lo(gv.w); gf[7 % V] = bf16_hi(gv.w);
            xf[0] = bf16_lo(xv.x); xf[1] = bf16_hi(xv.x); xf[2] = bf16_lo(xv.y); xf[3] = bf16_hi(xv.y);
            xf[4 % V] = bf16_lo(xv.z); xf[5 % V] = bf16_hi(xv.z); xf[6 % V] = bf16_lo(xv.w); xf[7 % V] = bf16_hi(xv.w);
        } else {
            gf[0] = __uint_as_float(gv.x); gf[1] = __uint_as_float(gv.y); gf[2] = __uint_as_float(gv.z); gf[3] = __uint_as_float(gv.w);
            xf[0] = __uint_as_float(xv.x); xf[1] = __uint_as_float(xv.y); xf[2] = __uint_as_float(xv.z); xf[3] = __uint_as_float(xv.w);
        }
#pragma unroll
        for (int q = 0; q < V; ++q) gf[q] = fmaf(__ldg(bc + c + q), xf[q], gf[q]) + __ldg(ac + c + q);
        if (sizeof(T) == 2) gv = make_uint4(pack_bf16(gf[0], gf[1]), pack_bf16(gf[2], gf[3]), pack_bf16(gf[4 % V], gf[5 % V]), pack_bf16(gf[6 % V], gf[7 % V]));
        else gv = make_uint4(__float_as_uint(gf[0]), __float_as_uint(gf[1]), __float_as_uint(gf[2]), __float_as_uint(gf[3]));
        *(reinterpret_cast<uint4 *>(g) + i) = gv;
    }
}

}  // namespace kdf

using namespace kdf;

extern "C" {

size_t kdf_rowbn_workspace_bytes(int C) { return sizeof(RowBnWs) + sizeof(double) * 2 * RB_COPIES * (size_t)C; }

int kdf_rowbn_stats(const void *x, int dtype, int64_t M, int C, const float *gamma, const float *beta,
                    const float *pre_bias, float eps, float momentum, float *running_mean, float *running_var,
                    float *mean, float *invstd, float *scale, float *shift, void *workspace, void *stream) {
    if (int e = rb_check(dtype, M, C, "rowbn_stats")) return e;
    KDF_CHECK_ARG(M > 0, "rowbn_stats: batch statistics need at least one row");
    KDF_CHECK_ARG(x && mean && invstd && scale && shift && workspace, "rowbn_stats: null pointer");
    KDF_CHECK_ARG((running_mean == nullptr) == (running_var == nullptr), "rowbn_stats: running stats come in pairs");
    ReduceArgs a{};
    a.x = x; a.M = M; a.C = C; a.gamma = gamma; a.beta = beta; a.pre_bias = pre_bias; a.eps = eps; a.momentum = momentum;
    a.mean = mean; a.invstd = invstd; a.out_scale = scale; a.out_shift = shift;
    a.running_mean = running_mean; a.running_var = running_var;
    a.ws = reinterpret_cast<RowBnWs *>(workspace);
    return launch_reduce<0>(a, dtype, as_stream(stream));
}

int kdf_rowbn_apply_fwd(const void *x, const void *residual, int dtype, int64_t M, int C,
                        const float *scale, const float *shift, int act, void *y, void *stream) {
    if (int e = rb_check(dtype, M, C, "rowbn_apply_fwd")) return e;
    KDF_CHECK_ARG(act >= 0 && act <= 2, "rowbn_apply_fwd: bad activation %d", act);
    if (M == 0) return KDF_OK;
    KDF_CHECK_ARG(x && y && scale && shift, "rowbn_apply_fwd: null pointer");
    const int vec = rb_vec(dtype, C);
    const int blocks = rb_blocks(M, C, vec, 8);
    cudaStream_t st = as_stream(stream);
#define KDF_RB_APPLY(T, V)                                                                                                      \
    do {                                                                                                                        \
        if (act == 1) rowbn_apply_fwd_kernel<T, V, 1><<<blocks, RB_THREADS, 0, st>>>((const T *)x, (const T *)residual, (T *)y, M, C, scale, shift);      \
        else if (act == 2) rowbn_apply_fwd_kernel<T, V, 2><<<blocks, RB_THREADS, 0, st>>>((const T *)x, (const T *)residual, (T *)y, M, C, scale, shift); \
        else rowbn_apply_fwd_kernel<T, V, 0><<<blocks, RB_THREADS, 0, st>>>((const T *)x, (const T *)residual, (T *)y, M, C, scale, shift);               \
    } while (0)
    if (dtype == KDF_F32) KDF_RB_APPLY(float, 4);
    else if (vec == 8) KDF_RB_APPLY(__nv_bfloat16, 8);
    else KDF_RB_APPLY(__nv_bfloat16, 4);
#undef KDF_RB_APPLY
    KDF_LAUNCH_CHECK();
    return KDF_OK;
}

int kdf_rowbn_bwd(const void *grad_out, const void *x, int dtype, int64_t M, int C,
                  const float *scale, const float *shift, const float *mean, const float *invstd,
                  int act, int batch_stats, void *grad_x, float *dgamma, float *dbeta,
                  void *workspace, void *stream) {
    if (int e = rb_check(dtype, M, C, "rowbn_bwd")) return e;
    KDF_CHECK_ARG(act >= 0 && act <= 2, "rowbn_bwd: bad activation %d", act);
    KDF_CHECK_ARG(M > 0, "rowbn_bwd: empty input");
    KDF_CHECK_ARG(grad_out && x && scale && shift && mean && invstd && grad_x && workspace, "rowbn_bwd: null pointer");
    cudaStream_t st = as_stream(stream);
    // coefficient vectors live at the tail of the workspace
    char *ws = reinterpret_cast<char *>(workspace);
    float *coefA = reinterpret_cast<float *>(ws + kdf_rowbn_workspace_bytes(C));
    float *coefB = coefA + C;
    ReduceArgs a{};
    a.x = x; a.g = grad_out; a.M = M; a.C = C; a.act = act; a.scale = scale; a.shift = shift;
    a.in_mean = mean; a.in_invstd = invstd; a.batch_stats = batch_stats;
    a.dgamma = dgamma; a.dbeta = dbeta; a.coefA = coefA; a.coefB = coefB;
    a.ws = reinterpret_cast<RowBnWs *>(workspace);
    const int vec = rb_vec(dtype, C);
    // one cooperative kernel (reduce, grid barrier, apply) up to 250 MB of input; measured in the step (B=32): never
    // 12.02 ms, up to 72 MB 12.00 ms, up to 250 MB 11.88 ms (all but the 2 x 201 MB expand BatchNorm of stage 2)
    constexpr long coop_mb = 250;
    const size_t in_bytes = 2 * (size_t)M * C * (dtype == KDF_F32 ? 4 : 2);
    if (coop_mb > 0 && in_bytes <= (size_t)coop_mb << 20) {
        const int rows = RB_THREADS / (C / vec);
        size_t smem = sizeof(float) * (size_t)rows * (C / vec) * 2 * vec;
        if (smem < sizeof(float) * 4 * (size_t)C) smem = sizeof(float) * 4 * (size_t)C;
        const int cblocks = rb_blocks(M, C, vec, 4, 3 * sm_count());                    // co-resident by construction (<= 3 CTAs/SM: the launch bounds)
        KDF_CUDA(cudaMemsetAsync(a.ws, 0, sizeof(RowBnWs) + sizeof(double) * 2 * RB_COPIES * (size_t)a.C, st));
        if (dtype == KDF_F32) return launch_bwd_coop<float, 4>(a, grad_x, cblocks, smem, st);
        if (vec == 8) return launch_bwd_coop<__nv_bfloat16, 8>(a, grad_x, cblocks, smem, st);
        return launch_bwd_coop<__nv_bfloat16, 4>(a, grad_x, cblocks, smem, st);
    }
    if (int e = launch_reduce<1>(a, dtype, st)) return e;
    const int blocks = rb_blocks(M, C, vec, 8);
    if (dtype == KDF_F32)
        rowbn_apply_bwd_kernel<float, 4><<<blocks, RB_THREADS, sizeof(float) * 4 * C, st>>>((const float *)grad_out, (const float *)x, (float *)grad_x, M, C, scale, shift, coefA, coefB, act);
    else if (vec == 8)
        rowbn_apply_bwd_kernel<__nv_bfloat16, 8><<<blocks, RB_THREADS, sizeof(float) * 4 * C, st>>>((const __nv_bfloat16 *)grad_out, (const __nv_bfloat16 *)x, (__nv_bfloat16 *)grad_x, M, C, scale, shift, coefA, coefB, act);
    else
        rowbn_apply_bwd_kernel<__nv_bfloat16, 4><<<blocks, RB_THREADS, sizeof(float) * 4 * C, st>>>((const __nv_bfloat16 *)grad_out, (const __nv_bfloat16 *)x, (__nv_bfloat16 *)grad_x, M, C, scale, shift, coefA, coefB, act);
    KDF_LAUNCH_CHECK();
    return KDF_OK;
}

size_t kdf_rowbn_bwd_workspace_bytes(int C) { return kdf_rowbn_workspace_bytes(C) + sizeof(float) * 2 * (size_t)C; }

int kdf_rows_axpb(void *g, const void *x, int dtype, int64_t M, int C, const float *bc, const float *ac, void *stream) {
    KDF_CHECK_ARG(M >= 0 && C > 0, "rows_axpb: bad sizes");
    KDF_CHECK_ARG(dtype == KDF_F32 || dtype == KDF_BF16, "rows_axpb: bad dtype %d", dtype);
    KDF_CHECK_ARG(C % (dtype == KDF_F32 ? 4 : 8) == 0, "rows_axpb: C=%d must fill 16-byte groups", C);
    if (M == 0) return KDF_OK;
    KDF_CHECK_ARG(g && x && bc && ac, "rows_axpb: null pointer");
    KDF_CHECK_ARG(((reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(x)) & 15) == 0, "rows_axpb: rows must be 16-byte aligned");
    const int64_t total = M * (C / (dtype == KDF_F32 ? 4 : 8));
    int64_t blocks = (total + 255) / 256;
    if (blocks > (int64_t)sm_count() * 16) blocks = (int64_t)sm_count() * 16;
    cudaStream_t st = as_stream(stream);
    if (dtype == KDF_F32) rows_axpb_kernel<float><<<(int)blocks, 256, 0, st>>>((float *)g, (const float *)x, bc, ac, M, C);
    else rows_axpb_kernel<__nv_bfloat16><<<(int)blocks, 256, 0, st>>>((__nv_bfloat16 *)g, (const __nv_bfloat16 *)x, bc, ac, M, C);
    KDF_LAUNCH_CHECK();
    return KDF_OK;
}

}  // extern "C"
