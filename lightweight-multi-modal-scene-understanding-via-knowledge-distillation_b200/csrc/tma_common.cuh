// TMA (cp.async.bulk.tensor) plumbing shared by the kernels that stage row tiles by TMA (sm_100a).
//
// The row tensors here are [M, 128] bf16 (pixel- or point-major rows).  A tensor map over such a tensor with boxes of
// 64 columns (one 128-byte swizzle span) x 128 rows and SWIZZLE_128B drops a tile into shared memory in exactly the
// panel layout tc_common.cuh describes (row r at r*128 bytes, 16-byte chunks XORed with r & 7), i.e. directly usable
// as a tcgen05 operand in either its K-major or its MN-major view.  Rows past M are zero-filled by the hardware.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "tc_common.cuh"

namespace kdf {
namespace tma {

__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(tc::smem_u32(bar)), "r"(bytes) : "memory");
}
// one box: c0 = first column (element index), c1 = first row
__device__ __forceinline__ void load_2d(void *smem_dst, const CUtensorMap *tmap, int c0, int c1, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(tc::smem_u32(smem_dst)), "l"(tmap), "r"(tc::smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}

// one box of a 4-D tensor (channels, x, y, frame); coordinates may be negative or run past the extents: the hardware
// zero-fills what lies outside, which is exactly a convolution's zero padding
__device__ __forceinline__ void load_4d(void *smem_dst, const CUtensorMap *tmap, int c0, int c1, int c2, int c3, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(tc::smem_u32(smem_dst)), "l"(tmap), "r"(tc::smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time dependency on libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static inline EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}
// [M, cols] bf16 rows (cols a multiple of 64), boxes of 64 columns x 128 rows, SWIZZLE_128B
static inline bool make_row_map(CUtensorMap *tm, const void *base, int64_t M, int cols) {
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc || M <= 0 || M >= (1ll << 31)) return false;
    const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)M};
    const cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
    const cuuint32_t box[2] = {64, 128};
    const cuuint32_t estr[2] = {1, 1};
    return enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// [B, H, W, C] bf16 map (pixel-major), boxes of box_c channels x box_w columns x box_h rows of one frame, no swizzle:
// a box lands densely as [row][column][channel]
static inline bool make_nhwc_map(CUtensorMap *tm, const void *base, int B, int H, int W, int C, int box_c, int box_w, int box_h) {
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc || B <= 0 || box_c > 256 || box_w > 256 || box_h > 256 || (C * 2) % 16 != 0) return false;
    const cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    const cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
    const cuuint32_t box[4] = {(cuuint32_t)box_c, (cuuint32_t)box_w, (cuuint32_t)box_h, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    return enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace tma
}  // namespace kdf
