// tcgen05 / TMEM / mbarrier plumbing for the fused point-MLP kernels (sm_100a only).
//
// Operand tiles live in shared memory in the canonical 128-byte-swizzled layout the 5th-gen tensor
// cores read: a tile of R rows x 64 bf16 (= 128 bytes per row) per "panel"; inside each group of 8
// rows the 16-byte chunk index is XORed with (row & 7) (Swizzle<3,4,3>), groups are 1024 bytes apart.
// The SAME bytes can be handed to the MMA either as a K-major operand (GEMM-K = the 64 columns of
// the panel, MN = rows) or as an MN-major operand (MN = columns, GEMM-K = rows); the fused backward
// uses both views of one tile (dX = dZ.W and dW = dZ^T.A).
#pragma once

#include <cuda_bf16.h>
#include <stdint.h>

namespace kdf {
namespace tc {

constexpr int PANEL_COLS = 64;                 // bf16 elements per 128-byte swizzle row
constexpr int ROW_BYTES = 128;
constexpr int GROUP_BYTES = 1024;              // 8 rows

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// First 1024-byte-aligned address of the dynamic shared memory, as pointer arithmetic ON the __shared__ array: the
// compiler keeps the address space and emits LDS / STS with 32-bit addresses.  (Rounding the pointer up through
// uintptr_t loses it: every shared-memory access of the kernel became a generic LD.E / ST.E with 64-bit address math.)
__device__ __forceinline__ uint8_t *align_smem_1024(uint8_t *raw) {
    const uint32_t base = static_cast<uint32_t>(__cvta_generic_to_shared(raw));
    return raw + ((1024u - (base & 1023u)) & 1023u);
}

// byte offset of 16-byte chunk `chunk` (0..7) of row `row` inside one panel
__device__ __forceinline__ uint32_t sw128_offset(int row, int chunk) {
    return (uint32_t)row * ROW_BYTES + (uint32_t)((chunk ^ (row & 7)) << 4);
}

// ----------------------------------------------------------------------------- descriptors
// 64-bit shared-memory matrix descriptor (PTX ISA "tcgen05 matrix descriptor"):
//   [0,14)  start address >> 4      [16,30) leading byte offset >> 4     [32,46) stride byte offset >> 4
//   [46,48) version = 1 (Blackwell) [49,52) base offset (0: tiles are 1024-byte aligned)
//   [61,64) layout: 2 = SWIZZLE_128B
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// K-major view: MN = rows (8-row groups 1024 B apart), K = the panel's 64 columns; LBO unused (1).
__device__ __forceinline__ uint64_t desc_kmajor(uint32_t saddr) { return make_desc(saddr, 16, GROUP_BYTES); }
// MN-major view: MN = columns (64 per panel, panels `panel_bytes` apart), K = rows (8-row groups 1024 B apart).
__device__ __forceinline__ uint64_t desc_mnmajor(uint32_t saddr, uint32_t panel_bytes) {
    return make_desc(saddr, panel_bytes, GROUP_BYTES);
}

// 32-bit instruction descriptor for kind::f16, bf16 x bf16 -> fp32:
//   [4,6) D format 1=F32   [7,10) A format 1=BF16   [10,13) B format 1=BF16
//   [15] A major (0 K, 1 MN)   [16] B major   [17,23) N>>3   [24,29) M>>4
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, int a_mn_major, int b_mn_major) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ----------------------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

// Bounded wait: a protocol bug traps (launch error) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    for (uint32_t spin = 0;; ++spin) {
        uint32_t done;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(addr), "r"(parity) : "memory");
        if (done) return;
        if (spin > (1u << 26)) __trap();
    }
}

// Pull `bytes` (multiple of 16) of global memory into L2 without occupying registers or shared memory (one thread).
__device__ __forceinline__ void prefetch_l2(const void *gptr, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gptr), "r"(bytes) : "memory");
}

// One 16-byte vector reduction into global memory (sm_90+; addr 16-byte aligned): a flush of an accumulator tile
// costs 4x fewer L2 atomic operations than scalar adds (they, not the bytes, bound such a flush).
__device__ __forceinline__ void red_add_v4(float *addr, float a, float b, float c, float d) {
    asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// ----------------------------------------------------------------------------- TMEM
__device__ __forceinline__ void tmem_alloc(uint32_t *dst_smem, uint32_t ncols) {      // one full warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {        // the same warp
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// generic-proxy shared-memory writes -> visible to the async proxy (tensor core operand reads)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// D[tmem] (+)= A[smem] . B[smem]; issued by ONE thread.
__device__ __forceinline__ void mma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, bool accumulate) {
    const uint32_t acc = accumulate ? 1u : 0u;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(acc) : "memory");
}
// mbarrier arrives when every MMA issued so far by this thread has completed (implies fence::before_thread_sync).
__device__ __forceinline__ void mma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// 32 lanes x 32 consecutive fp32 columns: thread `lane` of warp w reads TMEM lane 32*(w%4)+lane.
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
}
// 32 lanes x 16 consecutive fp32 columns
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
}
template <int N> __device__ __forceinline__ void tmem_ld(uint32_t taddr, uint32_t (&r)[N]);
template <> __device__ __forceinline__ void tmem_ld<32>(uint32_t taddr, uint32_t (&r)[32]) { tmem_ld32(taddr, r); }
template <> __device__ __forceinline__ void tmem_ld<16>(uint32_t taddr, uint32_t (&r)[16]) { tmem_ld16(taddr, r); }
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

}  // namespace tc
}  // namespace kdf
