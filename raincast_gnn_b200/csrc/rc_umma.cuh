// tcgen05 / TMEM / mbarrier / bulk-copy primitives shared by the tensor-core kernels (sm_100a inline PTX).
// The encodings below were pinned on a B200 with tools/ubench/umma_probe.cu (K-major operands without swizzle for any
// LBO / SBO, K-major with the 128-byte swizzle, and the A operand read from TMEM all reproduce the CPU product exactly;
// MN-major TF32 operands do NOT - they return zeros - so every operand here is K-major and transposition, where a
// contraction needs it, is done by the threads that convert the operand anyway).
#pragma once
#include "rc_common.cuh"

namespace rc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

// K-major shared-memory operand without swizzle: 16-byte chunks (4 tf32 / 8 bf16 along K) of 8 consecutive rows form a
// 128-byte core matrix; LBO = bytes between K-adjacent core matrices, SBO = bytes between 8-row groups.
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) |
         (1ull << 46);
}

// The same descriptor as two 32-bit halves, so that stepping through an operand is ONE 32-bit add on the low word.
// (The thread that issues the MMAs is a single instruction stream: with ~16 instructions of descriptor arithmetic per
//  tcgen05.mma it issued one MMA per ~90 cycles - slower than the tensor core executes a 128 x 128 x 8 TF32 MMA.)
struct UmmaDesc {
  uint32_t lo, hi;
  __device__ __forceinline__ uint64_t at(uint32_t byte_off) const { return ((uint64_t)hi << 32) | (uint64_t)(lo + (byte_off >> 4)); }
};
__device__ __forceinline__ UmmaDesc umma_desc2(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  UmmaDesc d;
  d.lo = ((smem_addr & 0x3FFFFu) >> 4) | ((lbo_bytes >> 4) << 16);
  d.hi = (sbo_bytes >> 4) | (1u << 14);
  return d;
}
// one lane of a converged warp (the MMA issuer): unlike `if (lane == 0)` the compiler then knows the predicate is
// warp-uniform and emits the tensor-core instructions without a per-lane election loop around each of them
__device__ __forceinline__ bool elect_one() {
  uint32_t p;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(p));
  return p != 0;
}

// instruction descriptor: D fp32, A/B tf32 (fmt 2) or bf16 (fmt 1), both K-major, M = 128
__device__ __forceinline__ uint32_t umma_idesc(uint32_t fmt, int n) {
  return (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24);
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
               :: "r"(tmem_d), "l"(a), "l"(b), "r"(idesc), "r"(accumulate));
}
// A operand from TMEM (lane = row of A, 32-bit column = K index)
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t b, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n"
               :: "r"(tmem_d), "r"(tmem_a), "l"(b), "r"(idesc), "r"(accumulate));
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
               :: "r"(tmem_d), "l"(a), "l"(b), "r"(idesc), "r"(accumulate));
}
// all MMAs issued so far by this thread arrive on the mbarrier when they have completed
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar) : "memory");
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" :: "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
// a wait that lasts two seconds is a protocol bug: trap (the launch fails loudly) instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) __trap();
  }
}

// one contiguous block global -> shared through the TMA engine (UBLKCP); bytes % 16 == 0, both addresses 16-byte aligned
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :: "r"(dst_smem), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// whole warp; cols: power of two in [32, 512]
__device__ __forceinline__ void tmem_alloc(uint32_t slot_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(slot_smem), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(taddr), "r"(cols) : "memory");
}

// 32 lanes x 32 columns -> one column per register (lane = this warp's quarter of the 128 TMEM lanes); completes at tmem_ld_wait()
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
               "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
               "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                 "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                 "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                 "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
               : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 "
               "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                 "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// 32 lanes x 16 columns from registers; completes at tmem_st_wait()
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
               "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};\n"
               :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
                  "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ float to_tf32(float v) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return __uint_as_float(r);
}
// 3xTF32 split: v = hi + lo with hi = round-to-nearest TF32 (the tensor core truncates lo to TF32 itself);
// a*b ~= a_hi*b_hi + a_hi*b_lo + a_lo*b_hi, error ~2^-21 per product
__device__ __forceinline__ void split_tf32(const float4 v, float4& hi, float4& lo) {
  hi.x = to_tf32(v.x); hi.y = to_tf32(v.y); hi.z = to_tf32(v.z); hi.w = to_tf32(v.w);
  lo.x = v.x - hi.x; lo.y = v.y - hi.y; lo.z = v.z - hi.z; lo.w = v.w - hi.w;
}

// round-to-nearest split without cvt.rna's NaN / Inf handling (three instructions per element; finite inputs):
// hi = (bits + 2^12) & ~(2^13 - 1), |lo| <= 2^-12 |v| with either sign, so the tensor core's truncation of lo is unbiased
__device__ __forceinline__ void split_tf32_rn(const float4 v, float4& hi, float4& lo) {
  hi.x = __uint_as_float((__float_as_uint(v.x) + 0x1000u) & 0xffffe000u); hi.y = __uint_as_float((__float_as_uint(v.y) + 0x1000u) & 0xffffe000u);
  hi.z = __uint_as_float((__float_as_uint(v.z) + 0x1000u) & 0xffffe000u); hi.w = __uint_as_float((__float_as_uint(v.w) + 0x1000u) & 0xffffe000u);
  lo.x = v.x - hi.x; lo.y = v.y - hi.y; lo.z = v.z - hi.z; lo.w = v.w - hi.w;
}
// the same split with hi = v truncated to TF32 (two instructions per element instead of five; hi exact in 11 bits, lo the
// remaining 13 bits of which the tensor core keeps 11: error <= 2^-22 per product, well inside the 1e-5 budget)
__device__ __forceinline__ void split_tf32_trunc(const float4 v, float4& hi, float4& lo) {
  hi.x = __uint_as_float(__float_as_uint(v.x) & 0xffffe000u); hi.y = __uint_as_float(__float_as_uint(v.y) & 0xffffe000u);
  hi.z = __uint_as_float(__float_as_uint(v.z) & 0xffffe000u); hi.w = __uint_as_float(__float_as_uint(v.w) & 0xffffe000u);
  lo.x = v.x - hi.x; lo.y = v.y - hi.y; lo.z = v.z - hi.z; lo.w = v.w - hi.w;
}

}  // namespace rc
