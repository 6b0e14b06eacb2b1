// sm_100a building blocks used by the tensor-core GEMM (mlp_gemm_tc.cu): mbarrier, async-proxy fence,
// tcgen05 TMEM allocation / MMA issue / commit / load, UMMA shared-memory and instruction descriptors.
// Bit layouts follow the CUTLASS sm100 headers (cute/arch/mma_sm100_desc.hpp) that ship with the image;
// nothing here includes them.
#pragma once
#include <stdint.h>

namespace pn2 {
namespace sm100 {

__device__ __forceinline__ uint32_t smem_addr(const void *p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---- mbarrier -------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  const uint32_t a = smem_addr(bar);
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(a), "r"(parity)
        : "memory");
  } while (!done);
}

// generic-proxy shared-memory writes -> visible to the async proxy (tcgen05.mma reads operands through it)
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ---- tensor memory --------------------------------------------------------------------------------------
template <uint32_t COLS>
__device__ __forceinline__ void tmem_alloc(uint32_t *dst_in_smem) {  // one full warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(dst_in_smem)),
               "n"(COLS)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <uint32_t COLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {  // the same warp that allocated
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(COLS) : "memory");
}
__device__ __forceinline__ void tc_fence_before_sync() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after_sync() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}

// 32 consecutive fp32 accumulator columns of this thread's TMEM lane (warp w owns lanes 32*(w%4)..+31)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n\t"
      "tcgen05.wait::ld.sync.aligned;"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// ---- UMMA descriptors -----------------------------------------------------------------------------------
// K-major operand tile, 128-byte swizzle: rows of 128 B (32 tf32), 8-row atoms of 1024 B (SBO), 16-byte
// chunks XOR-swizzled with (row % 8).  Tile base must be 1024-byte aligned.
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);        // start address  [0,14)
  d |= static_cast<uint64_t>(0) << 16;                         // leading byte offset (unused for SW128 K-major)
  d |= static_cast<uint64_t>(1024 >> 4) << 32;                 // stride byte offset [32,46)
  d |= static_cast<uint64_t>(1) << 46;                         // descriptor version (sm_100)
  d |= static_cast<uint64_t>(2) << 61;                         // layout type: SWIZZLE_128B
  return d;
}
// byte offset of element (row r, 16-byte chunk c16 in 0..7) inside such a tile
__device__ __forceinline__ uint32_t sw128_offset(int r, int c16) {
  return static_cast<uint32_t>((r >> 3) * 1024 + (r & 7) * 128 + ((c16 ^ (r & 7)) << 4));
}

// kind::tf32, fp32 accumulate, M x N tile, both operands K-major
__host__ __device__ constexpr uint32_t idesc_tf32(int m, int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(m >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem]; issued by ONE thread
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                         bool accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate ? 1u : 0u)
      : "memory");
}
// same with the A operand read from tensor memory (lane = tile row, one 32-bit column per k element)
__device__ __forceinline__ void mma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                            bool accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate ? 1u : 0u)
      : "memory");
}
// 8 consecutive 32-bit columns of this thread's TMEM lane (warp w owns lanes 32*(w%4)..+31)
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const float (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr),
               "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
               "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])),
               "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7]))
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// arrive on an mbarrier once every MMA issued so far by this thread has completed
__device__ __forceinline__ void mma_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_addr(bar))
               : "memory");
}

// One lane of a CONVERGED warp (elect.sync).  The MMA warp runs its loop with all 32 lanes on warp-uniform values and
// issues tcgen05.mma / tcgen05.commit under this predicate: inside an `if (lane == 0)` region the compiler cannot use
// the uniform datapath and rebuilds every descriptor operand with an ELECT / R2UR.BROADCAST / BRA.U.ANY loop per MMA
// -- ~87 cycles of issue per 64-cycle MMA (tools/mma_floor2.cu, profiles/r2_mma_floor2.txt).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred != 0;
}

// fp32 -> (hi, lo): hi = nearest tf32 of v, lo = nearest tf32 of the (exact) remainder v - hi.  The tensor core
// TRUNCATES fp32 operands to tf32; an unrounded remainder would lose up to 2^-22 |v| in one direction on every
// element -- a bias that adds up linearly over K and put the 3xTF32 GEMM at ~1e-6 relative where fp32 FMA is at
// 3e-7 (tests/arbiter.py).  Rounded, the error per operand is 2^-24 |v| with either sign.
// Round-to-nearest (ties away from zero) to tf32 is done on the bit pattern, add half an ulp of the 10-bit mantissa and
// clear the 13 dropped bits -- what cvt.rna.tf32.f32 computes for every finite value, in 2 instructions: ptxas expands
// the cvt into ~7 (NaN / Inf handling), and at 2 cvt per operand element that was a quarter of the producers' k-block.
// Inf stays Inf (lo = NaN), a NaN keeps a NaN in lo: either way the product is NaN, as with the cvt form.
__device__ __forceinline__ float round_tf32(float v) { return __uint_as_float((__float_as_uint(v) + 0x1000u) & 0xffffe000u); }
__device__ __forceinline__ void split_tf32(float v, float &hi, float &lo) {
  hi = round_tf32(v);
  lo = round_tf32(v - hi);
}

}  // namespace sm100
}  // namespace pn2
