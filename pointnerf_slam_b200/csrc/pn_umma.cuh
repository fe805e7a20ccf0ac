// tcgen05 (5th-gen tensor core) building blocks for sm_100a, hand-written PTX.
//
// Used for the 32-wide decoder layers: D[128 samples x N] (+)= A[128 x K] . B[N x K]^T with
// kind::tf32 operands in shared memory (canonical K-major, no-swizzle "interleave" layout) and
// the FP32 accumulator in tensor memory.  FP32-level accuracy comes from the 3xTF32 split
//     a.b  ~=  a_hi.b_hi + a_lo.b_hi + a_hi.b_lo ,   x_hi = tf32(x), x_lo = x - x_hi
// (the dropped a_lo.b_lo term is 2^-22 relative).
//
// Canonical K-major no-swizzle layout (cute::UMMA::make_umma_desc<Major::K>, LayoutType::INTERLEAVE):
// element (row r, col k) of a [rows x K] fp32 operand lives at byte
//     (r/8)*SBO + (k/4)*LBO + (r%8)*16 + (k%4)*4
// i.e. 8-row x 16-byte "core matrices" stored contiguously (128 B); LBO = byte distance between
// core matrices adjacent in K, SBO = between 8-row groups.  One MMA consumes K = 8 (two core
// matrices in K); advancing k by 8 advances the descriptor start address by 2*LBO.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pn {
namespace umma {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- operand split -----------------------------------------------------------------------
// Round to TF32 (10 mantissa bits), nearest with ties away from zero: bit-identical to cvt.rna.tf32.f32, but as one
// integer add and one mask on the ALU pipe (the cvt is a multi-cycle conversion-pipe instruction: 5 % of the forward
// kernel's stall samples in profiles/r2_*; a decoder pass converts 256 values per sample).
__device__ __forceinline__ float tf32_hi(float x) {
  return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);
}
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
  hi = tf32_hi(x);
  lo = x - hi;  // exact in fp32
}

// byte offset of element (r,k) in the canonical layout
__device__ __forceinline__ uint32_t kmajor_off(int r, int k, uint32_t lbo, uint32_t sbo) {
  return (uint32_t)(r >> 3) * sbo + (uint32_t)(k >> 2) * lbo + (uint32_t)(r & 7) * 16u + (uint32_t)(k & 3) * 4u;
}

// ---- descriptors -------------------------------------------------------------------------
// 64-bit shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start>>4 [0,14),
// LBO>>4 [16,30), SBO>>4 [32,46), version=1 [46,48), base_offset 0, layout_type 0 (no swizzle).
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}

// 32-bit instruction descriptor (cute::UMMA::InstrDescriptor) for kind::tf32, FP32 accumulate,
// A and B K-major: c_format=F32 [4,6), a_format=TF32 [7,10), b_format=TF32 [10,13),
// n_dim=N>>3 [17,23), m_dim=M>>4 [24,29).
__host__ __device__ constexpr uint32_t instr_desc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ---- mbarrier ----------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

// Bounded wait: a protocol bug traps (reported as a launch failure) instead of hanging the GPU.
// try_wait suspends the thread for up to the hinted time before it reports failure, so the loop
// spins rarely; the bound is a poll count (~seconds), far beyond any legitimate wait.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done = 0;
  for (uint32_t polls = 0;; ++polls) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity), "r"(20000u)
        : "memory");
    if (done) break;
    if (polls > (1u << 20)) __trap();
  }
}

// ---- fences ------------------------------------------------------------------------------
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ---- tensor memory -----------------------------------------------------------------------
// Warp-collective.  ncols: power of two >= 32.  The base address is written to *dst (shared).
__device__ __forceinline__ void tmem_alloc(uint32_t* dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// ---- MMA ---------------------------------------------------------------------------------
// D[tmem] (+)= A[smem] . B[smem]^T, issued by ONE thread.
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Make the mbarrier track completion of all MMAs issued so far by this thread.
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// 3xTF32 product of one [128 x K] A operand (hi/lo copies) with one [N x K] B operand (hi/lo):
// K/8 k-steps x 3 MMAs.  `first` = 0 overwrites D, otherwise accumulates.
__device__ __forceinline__ void mma_3xtf32(uint32_t d_tmem, uint32_t a_hi, uint32_t a_lo, uint32_t a_lbo, uint32_t a_sbo,
                                           uint32_t b_hi, uint32_t b_lo, uint32_t b_lbo, uint32_t b_sbo, int K, uint32_t idesc,
                                           uint32_t accumulate) {
  for (int ks = 0; ks < K / 8; ++ks) {
    const uint64_t ah = smem_desc(a_hi + 2 * ks * a_lbo, a_lbo, a_sbo), al = smem_desc(a_lo + 2 * ks * a_lbo, a_lbo, a_sbo);
    const uint64_t bh = smem_desc(b_hi + 2 * ks * b_lbo, b_lbo, b_sbo), bl = smem_desc(b_lo + 2 * ks * b_lbo, b_lbo, b_sbo);
    mma_tf32(d_tmem, al, bh, idesc, accumulate | (uint32_t)(ks > 0));  // small terms first
    mma_tf32(d_tmem, ah, bl, idesc, 1u);
    mma_tf32(d_tmem, ah, bh, idesc, 1u);
  }
}

// Same product with descriptors prepared once: a_hi/a_lo/b_hi/b_lo are descriptors of the
// operands' first K-step; a K-step of 8 advances the 14-bit start-address field by
// 2*LBO/16 (no carry into the LBO field for any shared-memory address).  K = 32.
__device__ __forceinline__ void mma_3xtf32_k32(uint32_t d_tmem, uint64_t a_hi, uint64_t a_lo, uint64_t b_hi, uint64_t b_lo,
                                               uint32_t a_step, uint32_t b_step, uint32_t idesc, uint32_t accumulate) {
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
    const uint64_t ah = a_hi + (uint64_t)(ks * a_step), al = a_lo + (uint64_t)(ks * a_step);
    const uint64_t bh = b_hi + (uint64_t)(ks * b_step), bl = b_lo + (uint64_t)(ks * b_step);
    mma_tf32(d_tmem, al, bh, idesc, ks == 0 ? accumulate : 1u);
    mma_tf32(d_tmem, ah, bl, idesc, 1u);
    mma_tf32(d_tmem, ah, bh, idesc, 1u);
  }
}

// ---- TMEM -> registers: 32 consecutive columns of this thread's lane ----------------------
// taddr: (lane << 16) | column; a warp may only touch lanes [32*(warp%4), +32).
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

}  // namespace umma
}  // namespace pn
