// Shared device/host helpers for the kdfusion_b200 kernels (sm_100a only).
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/kdfusion_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "kdfusion_b200 targets sm_100a (B200) only"
#endif

namespace kdf {

// ----------------------------------------------------------------------------- errors
void set_error(const char *fmt, ...);
int sm_count();                       // cached cudaDevAttrMultiProcessorCount of the current device

#define KDF_CHECK_ARG(cond, ...)                      \
    do {                                              \
        if (!(cond)) {                                \
            kdf::set_error(__VA_ARGS__);              \
            return KDF_ERR_ARG;                       \
        }                                             \
    } while (0)

#define KDF_CUDA(call)                                                                     \
    do {                                                                                   \
        cudaError_t _e = (call);                                                           \
        if (_e != cudaSuccess) {                                                           \
            kdf::set_error("%s failed at %s:%d: %s", #call, __FILE__, __LINE__,            \
                           cudaGetErrorString(_e));                                        \
            return KDF_ERR_CUDA;                                                           \
        }                                                                                  \
    } while (0)

#define KDF_LAUNCH_CHECK()                                                                 \
    do {                                                                                   \
        cudaError_t _e = cudaGetLastError();                                               \
        if (_e != cudaSuccess) {                                                           \
            kdf::set_error("kernel launch failed at %s:%d: %s", __FILE__, __LINE__,        \
                           cudaGetErrorString(_e));                                        \
            return KDF_ERR_CUDA;                                                           \
        }                                                                                  \
    } while (0)

static inline cudaStream_t as_stream(void *s) { return reinterpret_cast<cudaStream_t>(s); }
static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ----------------------------------------------------------------------------- warp / block reductions
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Sum over a block of NT threads (NT % 32 == 0, NT <= 1024); result valid in every thread.
template <int NT>
__device__ __forceinline__ float block_sum(float v, float *smem /* >= NT/32 floats */) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) smem[wid] = v;
    __syncthreads();
    float r = (lane < NT / 32) ? smem[lane] : 0.f;
    r = warp_sum(r);
    return r;
}

// ----------------------------------------------------------------------------- streaming loads / stores
__device__ __forceinline__ float4 ldg_stream_f4(const float4 *p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ uint2 ldg_stream_u2(const uint2 *p) {
    uint2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ uint4 ldg_stream_u4(const uint4 *p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}

__device__ __forceinline__ float bf16_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t *>(&v);
}

// 4 consecutive feature elements <-> float4, for both storage types.
template <typename T> struct Vec4;
template <> struct Vec4<float> {
    static __device__ __forceinline__ float4 load(const float *p) { return *reinterpret_cast<const float4 *>(p); }
    static __device__ __forceinline__ float4 load_stream(const float *p) {
        return ldg_stream_f4(reinterpret_cast<const float4 *>(p));
    }
    static __device__ __forceinline__ void store(float *p, float4 v) { *reinterpret_cast<float4 *>(p) = v; }
};
template <> struct Vec4<__nv_bfloat16> {
    static __device__ __forceinline__ float4 cvt(uint2 u) {
        return make_float4(bf16_lo(u.x), bf16_hi(u.x), bf16_lo(u.y), bf16_hi(u.y));
    }
    static __device__ __forceinline__ float4 load(const __nv_bfloat16 *p) {
        return cvt(*reinterpret_cast<const uint2 *>(p));
    }
    static __device__ __forceinline__ float4 load_stream(const __nv_bfloat16 *p) {
        return cvt(ldg_stream_u2(reinterpret_cast<const uint2 *>(p)));
    }
    static __device__ __forceinline__ void store(__nv_bfloat16 *p, float4 v) {
        uint2 u;
        u.x = pack_bf16(v.x, v.y);
        u.y = pack_bf16(v.z, v.w);
        *reinterpret_cast<uint2 *>(p) = u;
    }
};

template <typename T> __device__ __forceinline__ float to_float(T v);
template <> __device__ __forceinline__ float to_float<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_float<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_float(float v);
template <> __device__ __forceinline__ float from_float<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_float<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

}  // namespace kdf
