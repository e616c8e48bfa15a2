// tcgen05 (5th-generation tensor core) helpers for the exact-integer A-products: kind::i8 MMA with int32
// accumulators in tensor memory, operands in shared memory in the no-swizzle canonical layout, operand tiles
// brought in by bulk asynchronous copies (cp.async.bulk, the 1-D form of TMA) that complete on an mbarrier.
// sm_100a only (B200 keeps the INT8 tensor path).
//
// Operand block (the unit that is stored in global memory and copied to shared memory): 128 x 128 bytes,
//   byte(i, k) at  (k / 16) * 2048 + i * 16 + (k % 16),   i, k in [0, 128)
// i.e. eight column slabs of 128 rows x 16 bytes.  The same block is a valid tcgen05 operand in two ways:
//   * K-major   (rows i = M, bytes k = K):  8 x 16 B core matrices, LBO (next 16 k) = 2048, SBO (next 8 i) = 128
//   * MN-major  (bytes k = M, rows i = K):  8 (K) x 16 (M) core matrices, SBO (next 16 k) = 2048, LBO (next 8 i) = 128
// so one stored copy of the sensing matrix serves A (.) and A' (.).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace twoace {
namespace umma {

constexpr int BLK = 128;                 // block edge (rows and bytes)
constexpr int BLK_BYTES = BLK * BLK;     // 16 KB
constexpr int SLAB = BLK * 16;           // 2048: one 16-byte column slab of a block

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// one lane of a converged warp (warp-uniform control flow around it keeps descriptors in uniform registers:
// tcgen05.mma / cp.async.bulk issued from a divergent `if (lane == 0)` region compile to an ELECT + R2UR waterfall
// loop that costs ~200 cycles per instruction)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.b32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ int uniform_warp_id() { return __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0); }

// ---- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "W_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra D_%=;\n\t"
      "bra W_%=;\n\t"
      "D_%=:\n\t}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

// ---- bulk asynchronous copy global -> shared, completion counted in bytes on `bar`
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// generic-proxy writes (st.shared) -> visible to the async proxy (tcgen05.mma operand reads)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- tensor memory
__device__ __forceinline__ void tmem_alloc512(uint32_t* slot_smem) {   // one full warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(slot_smem)) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_free512(uint32_t addr) {          // the same warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(addr) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// lane = 32 * (warp % 4) + laneid reads 8 consecutive 32-bit columns starting at `taddr`
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, int32_t (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, int32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- descriptors
// shared-memory matrix descriptor, no swizzle: start >> 4 | LBO >> 4 << 16 | SBO >> 4 << 32 | version 1 << 46
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) |
         ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) | ((uint64_t)1 << 46);
}
// instruction descriptor of kind::i8: D = S32, A = B = signed 8 bit, M = 128, N = n
__host__ __device__ constexpr uint32_t idesc_i8(int n, bool a_mn_major, bool b_mn_major) {
  return (2u << 4) | (1u << 7) | (1u << 10) | ((a_mn_major ? 1u : 0u) << 15) | ((b_mn_major ? 1u : 0u) << 16) |
         ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem]', one K = 32 step; issued by ONE thread
__device__ __forceinline__ void mma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, bool accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate ? 1u : 0u)
      : "memory");
}
// all MMAs issued so far by this thread complete -> one arrival on `bar`
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

}  // namespace umma
}  // namespace twoace
