// Shared declarations of the grid-decoder kernels (pn_gridmlp*.cu).
#pragma once
#include <stdlib.h>

#include "pn_common.cuh"
#include "pn_umma.cuh"

namespace pn {

struct MlpDev {  // device copy of pn_grid_mlp pointers
  const float* B; const float* W[5]; const float* b[5]; const float* Wc[5]; const float* bc[5];
  const float* Wo; const float* bo;
};
inline MlpDev make_mlp(const pn_grid_mlp* w) {
  MlpDev m;
  m.B = w->B; m.Wo = w->Wo; m.bo = w->bo;
  for (int i = 0; i < 5; ++i) { m.W[i] = w->W[i]; m.b[i] = w->b[i]; m.Wc[i] = w->Wc[i]; m.bc[i] = w->bc[i]; }
  return m;
}

__device__ __forceinline__ void store_planar32(float* base, int64_t N, int64_t n, const float (&v)[32]) {
  float4* o = reinterpret_cast<float4*>(base);
#pragma unroll
  for (int q = 0; q < 8; ++q) o[(int64_t)q * N + n] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
}


struct FwdArgs {
  pn_points pts;
  MlpDev w;
  GridDev ga, gb;
  Bound6 nb, mb;
  int apply_mask, out_mode;
  float* raw;
  uint32_t* relu_bits; float* H; float* C; float* E;
};


struct BwdArgs {
  pn_points pts;
  MlpDev w;
  GridDev ga, gb;
  Bound6 nb, mb;
  int apply_mask, accumulate_pts;
  const float* g_raw;
  const uint32_t* relu_bits;
  float* g_grid; float* g_pts;
  float* GA; float* GH; float* GARG; float* P32; float* GO;
};


// tensor-core kernels (pn_gridmlp_tc_fwd.cu / pn_gridmlp_tc_bwd.cu): c_dim in {32, 64}, n_out in {1, 4}
int launch_fwd_tc(int c_dim, int n_out, const FwdArgs& a, cudaStream_t st);
int launch_bwd_tc(int c_dim, int n_out, bool grid_grad, bool need_dp, bool wstash, const BwdArgs& a, cudaStream_t st);

// ---- tcgen05 decoder kernels: shared layout constants and helpers ---------------------------
// CTA = 512 threads = two independent groups of 256; a group owns one tile of 128 consecutive
// samples at a time (two threads per sample row = TMEM lane, 16 columns each).  Per tile every
// decoder layer is a D[128 x N] (+)= A[128 x 32] . W[N x 32]^T tensor-core product: the activation
// operand A is written by the group (hi and lo copies, canonical K-major layout) into its 32 KB
// shared buffer, the weights sit pre-split in shared memory for the whole kernel, the accumulators
// live in the group's 256 tensor-memory columns:
//     cols   0..159  D2_l = Wc_l . c           (feature terms of all five blocks: ONE N=160 product)
//     cols 160..191  D1_0 = W0 . emb           (3 K-chunks of 32; D1_0 and D1_3 are ONE N=64 product)
//     cols 192..223  D1_3 = W3[:, :93] . emb + W3[:, 93:] . h2
//     cols 224..255  D1_x = W_l . h_{l-1}      (blocks 1, 2, 4)
// The epilogue of a block reads its accumulators back (tcgen05.ld), applies
// relu(D1 + b) + D2 + bc in registers, re-splits into hi/lo and stores the next A operand.
// Everything that does not depend on the product in flight -- the next chunk's sines, the second
// grid's gather, the stash stores, the next block's D2 + bc -- is done between issuing the MMAs
// and waiting for them; while one group waits the other group runs.
namespace tc {
constexpr uint32_t kLbo = 128;                       // core matrices adjacent in K
constexpr uint32_t kASbo = 8 * 128;                  // A buffer: K = 32 per 8-row group
constexpr uint32_t kABytes = 16 * kASbo;             // one 128 x 32 operand copy (16 KB)
__host__ __device__ constexpr uint32_t bsbo(int K) { return (uint32_t)(K / 4) * 128u; }
__host__ __device__ constexpr uint32_t bbytes(int K) { return 4u * bsbo(K); }  // 32 rows
// byte offsets of the pre-split weight operands
constexpr uint32_t O_WE = 0;                         // [W0; W3[:, :93]] as one [64 x 96] operand: hi, then lo
constexpr uint32_t O_WH = O_WE + 4 * bbytes(96);     // 4 x [32 x 32]: hi, lo per block
template <int CD> __host__ __device__ constexpr uint32_t o_wc() { return O_WH + 4 * 2 * bbytes(32); }   // [160 x CD]: hi, then lo
template <int CD> __host__ __device__ constexpr uint32_t o_a() { return o_wc<CD>() + 5u * 2u * bbytes(CD); }   // A buffers
// forward kernels: K-stride of the A operand's core matrices (padded where shared memory allows: no store bank conflicts)
template <int CD> __host__ __device__ constexpr uint32_t a_lbo() { return CD == 32 ? 144u : 128u; }
template <int CD> __host__ __device__ constexpr uint32_t o_small() { return o_a<CD>() + 2u * 2u * 16u * 8u * a_lbo<CD>(); }
constexpr int S_B = 0, S_BIAS = 288, S_BC = 448, S_WO = 608, S_BO = 736, S_TOTAL = 740;  // floats
// the two mbarriers and the TMEM base address follow the float area (kept in dynamic shared
// memory: with c_dim 64 the kernel uses all but ~100 bytes of the 227 KB an SM offers)
template <int CD> __host__ __device__ constexpr uint32_t smem_total() { return o_small<CD>() + S_TOTAL * 4u + 24u; }

// split a [32 x K] row-major weight block (row stride ld, starting column col0, `kvalid` real
// columns, zero beyond) into canonical hi / lo operands
__device__ __forceinline__ void stage_b(unsigned char* hi, unsigned char* lo, const float* __restrict__ src, int ld, int col0, int K,
                                        int kvalid) {
  // lane bits = (k & 3, n & 7): a warp's 32 stores fill one 128-byte core matrix (no bank conflicts)
  for (int i = threadIdx.x; i < 32 * K; i += blockDim.x) {
    const int rest = i >> 5, kc = rest % (K / 4), ng = rest / (K / 4);
    const int n = ng * 8 + ((i >> 2) & 7), k = kc * 4 + (i & 3);
    const float w = k < kvalid ? src[n * ld + col0 + k] : 0.f;
    float h, l;
    umma::split_tf32(w, h, l);
    const uint32_t off = umma::kmajor_off(n, k, kLbo, bsbo(K));
    *reinterpret_cast<float*>(hi + off) = h;
    *reinterpret_cast<float*>(lo + off) = l;
  }
}
}  // namespace tc

// 128-bit vector reduction into global memory (sm_90+): four float atomics in one instruction
__device__ __forceinline__ void red_add_v4(float* p, float x, float y, float z, float w) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(x), "f"(y), "f"(z), "f"(w) : "memory");
}

// TMEM -> registers, 16 columns
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}


}  // namespace pn
