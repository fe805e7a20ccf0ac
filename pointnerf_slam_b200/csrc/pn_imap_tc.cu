// Tensor-core GEMM for the iMAP* 256-wide MLP (decoder.MLP with c_dim 0; src/conv_onet/config.py:28-32,
// src/conv_onet/models/decoder.py:189-203): forward layers, input gradients and weight gradients.
//
//     C[M x N] (op)= sum_k A(m,k) * B(k,n),   A(m,k) = A[m*sam + k*sak],   B(k,n) = B[k*sbk + n*sbn]
//
// the same strided interface as the FFMA k_sgemm of pn_imap.cu, so the layer chain there is unchanged.  One CTA owns
// a 128 x N tile (N <= 256 = the tensor-memory columns of one accumulator) and walks K in chunks of 32:
//   * producers (all 256 threads): global -> registers -> 3xTF32 split (hi = tf32(x), lo = x - hi) -> canonical K-major
//     operand tiles in shared memory.  An operand whose K index is contiguous in memory (forward: activations and
//     weights) is read as 16-byte units that ARE the layout's 16-byte units; an operand whose row index is contiguous
//     (input gradients: W read along its output index; weight gradients: both operands, K = samples) is read as
//     4 x 4 blocks -- four coalesced 16-byte loads along the row index -- transposed in registers and stored as four
//     16-byte units, so neither case issues a scalar shared-memory store.  LBO = 144 and SBO = 1168 bytes (padding) keep
//     both store patterns free of bank conflicts;
//   * two operand stages: while the tensor core multiplies chunk c, the threads load and split chunk c+1; an mbarrier
//     per stage, armed by tcgen05.commit, says when a stage may be overwritten;
//   * one thread issues 12 MMAs per chunk (4 k-steps x {lo.hi, hi.lo, hi.hi}), M = 128, N = N rounded up to 16;
//   * epilogue: tensor memory -> registers (two threads per row, half of the columns each) -> bias + ReLU / ReLU mask
//     / plain store / split-K atomics -> global.
#include "pn_common.cuh"
#include "pn_umma.cuh"

namespace pn {
namespace {

constexpr int kGT = 256;                       // threads
constexpr uint32_t kGLbo = 144;                // 16-byte unit stride along K (128-byte core matrix + 16 bytes of padding)
constexpr uint32_t kGSbo = 8 * kGLbo + 16;     // 8-row group stride: 8 units along K (K chunk = 32) + 16 bytes
constexpr uint32_t kGA = 16 * kGSbo;           // one copy of a 128-row operand
constexpr uint32_t kGB = 32 * kGSbo;           // one copy of a 256-row operand
constexpr uint32_t kGStage = 2 * kGA + 2 * kGB;
constexpr uint32_t kGSmem = 2 * kGStage + 64;

enum { G_STORE = 0, G_BIAS_RELU = 1, G_MASK = 2, G_ATOMIC = 3 };

struct GemmArgs {
  const float* A; int64_t sam, sak;
  const float* B; int64_t sbk, sbn;
  float* C; int64_t ldc;
  int64_t M, K;
  int N;
  const float* bias; const float* aux;
  int a_vec, b_vec;      // 16-byte loads allowed (base and strides 16-byte aligned)
  int c_vec;             // 16-byte stores to C (and loads of aux / bias) allowed
  float* a_rowsum;       // optional (G_ATOMIC, row-contiguous A): a_rowsum[m] += sum_k A(m,k)  (bias gradients)
};

__device__ __forceinline__ void sts_split(unsigned char* hi, unsigned char* lo, uint32_t off, float4 v) {
  float4 h, l;
  umma::split_tf32(v.x, h.x, l.x); umma::split_tf32(v.y, h.y, l.y);
  umma::split_tf32(v.z, h.z, l.z); umma::split_tf32(v.w, h.w, l.w);
  *reinterpret_cast<float4*>(hi + off) = h;
  *reinterpret_cast<float4*>(lo + off) = l;
}

// four consecutive elements along the contiguous index starting at element (r, c) of a strided operand;
// `rs` = stride of the other index, elements beyond the limits read 0
__device__ __forceinline__ float4 ld_row4(const float* __restrict__ base, int64_t other, int64_t rs, int64_t c, int64_t climit,
                                          bool other_ok, bool vec) {
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (!other_ok) return v;
  const float* p = base + other * rs + c;
  if (vec && c + 3 < climit) return __ldg(reinterpret_cast<const float4*>(p));
  if (c < climit) v.x = __ldg(p);
  if (c + 1 < climit) v.y = __ldg(p + 1);
  if (c + 2 < climit) v.z = __ldg(p + 2);
  if (c + 3 < climit) v.w = __ldg(p + 3);
  return v;
}

// A_KC / B_KC: the operand's K index is the contiguous one (sak == 1 / sbk == 1)
template <int EP, bool A_KC, bool B_KC>
__global__ void __launch_bounds__(kGT, 1) k_tc_gemm(const GemmArgs a) {
  extern __shared__ __align__(128) unsigned char smraw[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smraw + 2 * kGStage);    // [2] stage free
  uint32_t& tmem_base_s = *reinterpret_cast<uint32_t*>(smraw + 2 * kGStage + 32);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t m0 = (int64_t)blockIdx.x * 128;
  const int N16 = (a.N + 15) & ~15;
  int64_t k_begin = 0, k_end = a.K;
  if (EP == G_ATOMIC) {
    const int64_t per = ((a.K + gridDim.z - 1) / gridDim.z + 31) / 32 * 32;
    k_begin = (int64_t)blockIdx.z * per;
    k_end = k_begin + per < a.K ? k_begin + per : a.K;
    if (k_begin >= k_end) return;
  }
  if (warp == 0) umma::tmem_alloc(&tmem_base_s, 256);
  if (tid == 0) { umma::mbar_init(&bars[0], 1); umma::mbar_init(&bars[1], 1); umma::fence_mbar_init(); }
  umma::tc_fence_before();
  __syncthreads();
  umma::tc_fence_after();
  const uint32_t tm = tmem_base_s;
  const uint32_t sbase = umma::smem_u32(smraw);
  const uint32_t idesc = umma::instr_desc_tf32(128, N16);
  constexpr uint32_t kStep = (2u * kGLbo) >> 4;
  uint32_t phase[2] = {0u, 0u};
  const int64_t nchunks = (k_end - k_begin + 31) / 32;
  // rows of the B tile that exist (rounded to the 4-row blocks the loaders use)
  const int nB = N16;
  float4 rsum = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 ra[4], rb[8];
  // global -> registers for chunk c (issued one chunk ahead: the loads fly under the barrier, the MMA issue and the next wait)
  auto load_chunk = [&](int64_t c) {
    const int64_t k0 = k_begin + c * 32;
    if (A_KC) {          // 128 rows x 8 units: unit (row = tid/8 + 32 j, kq = tid%8)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int row = (tid >> 3) + 32 * j, kq = tid & 7;
        ra[j] = ld_row4(a.A, m0 + row, a.sam, k0 + 4 * kq, k_end, m0 + row < a.M, a.a_vec);
      }
    } else {             // 4 x 4 blocks: (mq = lane, kq = warp): rows 4 mq .. +3, k = 4 kq + j
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int64_t k = k0 + 4 * warp + j;
        ra[j] = ld_row4(a.A, k, a.sak, m0 + 4 * lane, a.M, k < k_end, a.a_vec);
      }
    }
    if (B_KC) {          // nB rows x 8 units
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int row = (tid >> 3) + 32 * j, kq = tid & 7;
        rb[j] = (row < nB) ? ld_row4(a.B, row, a.sbn, k0 + 4 * kq, k_end, row < a.N, a.b_vec) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    } else {             // blocks (nq = lane + 32 h, kq = warp), h = 0, 1
#pragma unroll
      for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int64_t k = k0 + 4 * warp + j;
          const int n = 4 * (lane + 32 * h);
          rb[4 * h + j] = (n < nB) ? ld_row4(a.B, k, a.sbk, n, a.N, k < k_end, a.b_vec) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
  };
  load_chunk(0);
  for (int64_t c = 0; c < nchunks; ++c) {
    const int st = (int)(c & 1);
    unsigned char* ah = smraw + st * kGStage;
    unsigned char* al = ah + kGA;
    unsigned char* bh = al + kGA;
    unsigned char* bl = bh + kGB;
    if (c >= 2) { umma::mbar_wait(&bars[st], phase[st]); phase[st] ^= 1u; umma::tc_fence_after(); }
    // ---- split + store in the canonical layout
    if (A_KC) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int row = (tid >> 3) + 32 * j, kq = tid & 7;
        sts_split(ah, al, (uint32_t)(row >> 3) * kGSbo + (uint32_t)kq * kGLbo + (uint32_t)(row & 7) * 16u, ra[j]);
      }
    } else {
      const float x[4][4] = {{ra[0].x, ra[1].x, ra[2].x, ra[3].x}, {ra[0].y, ra[1].y, ra[2].y, ra[3].y},
                             {ra[0].z, ra[1].z, ra[2].z, ra[3].z}, {ra[0].w, ra[1].w, ra[2].w, ra[3].w}};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int row = 4 * lane + i;
        sts_split(ah, al, (uint32_t)(row >> 3) * kGSbo + (uint32_t)warp * kGLbo + (uint32_t)(row & 7) * 16u,
                  make_float4(x[i][0], x[i][1], x[i][2], x[i][3]));
      }
      rsum.x += (x[0][0] + x[0][1]) + (x[0][2] + x[0][3]); rsum.y += (x[1][0] + x[1][1]) + (x[1][2] + x[1][3]);
      rsum.z += (x[2][0] + x[2][1]) + (x[2][2] + x[2][3]); rsum.w += (x[3][0] + x[3][1]) + (x[3][2] + x[3][3]);
    }
    if (B_KC) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int row = (tid >> 3) + 32 * j, kq = tid & 7;
        if (row < nB) sts_split(bh, bl, (uint32_t)(row >> 3) * kGSbo + (uint32_t)kq * kGLbo + (uint32_t)(row & 7) * 16u, rb[j]);
      }
    } else {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int nq = lane + 32 * h;
        if (4 * nq < nB) {
          const float4* r = rb + 4 * h;
          const float x[4][4] = {{r[0].x, r[1].x, r[2].x, r[3].x}, {r[0].y, r[1].y, r[2].y, r[3].y},
                                 {r[0].z, r[1].z, r[2].z, r[3].z}, {r[0].w, r[1].w, r[2].w, r[3].w}};
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int row = 4 * nq + i;
            sts_split(bh, bl, (uint32_t)(row >> 3) * kGSbo + (uint32_t)warp * kGLbo + (uint32_t)(row & 7) * 16u,
                      make_float4(x[i][0], x[i][1], x[i][2], x[i][3]));
          }
        }
      }
    }
    if (c + 1 < nchunks) load_chunk(c + 1);
    umma::fence_proxy_async();
    umma::tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      umma::tc_fence_after();
      const uint32_t sa = sbase + (uint32_t)st * kGStage;
      const uint64_t dah = umma::smem_desc(sa, kGLbo, kGSbo), dal = umma::smem_desc(sa + kGA, kGLbo, kGSbo);
      const uint64_t dbh = umma::smem_desc(sa + 2 * kGA, kGLbo, kGSbo), dbl = umma::smem_desc(sa + 2 * kGA + kGB, kGLbo, kGSbo);
      umma::mma_3xtf32_k32(tm, dah, dal, dbh, dbl, kStep, kStep, idesc, c == 0 ? 0u : 1u);
      umma::mma_commit(&bars[st]);
    }
  }
  // drain: the last commit on each used stage
  {
    const int s1 = (int)((nchunks - 1) & 1);
    umma::mbar_wait(&bars[s1], phase[s1]); phase[s1] ^= 1u;
    if (nchunks >= 2) { const int s0 = s1 ^ 1; umma::mbar_wait(&bars[s0], phase[s0]); phase[s0] ^= 1u; }
  }
  umma::tc_fence_after();
  if (EP == G_ATOMIC && !A_KC && a.a_rowsum) {      // rows 4*lane.. of this tile, partial over this warp's k slots
    const int64_t m = m0 + 4 * lane;
    if (m < a.M) atomicAdd(a.a_rowsum + m, rsum.x);
    if (m + 1 < a.M) atomicAdd(a.a_rowsum + m + 1, rsum.y);
    if (m + 2 < a.M) atomicAdd(a.a_rowsum + m + 2, rsum.z);
    if (m + 3 < a.M) atomicAdd(a.a_rowsum + m + 3, rsum.w);
  }
  // ---- epilogue: warp w reads lanes 32*(w%4).., columns [half*N16/2 ...) in chunks of 32
  {
    const int quarter = warp & 3, half = warp >> 2;
    const int64_t gm = m0 + quarter * 32 + lane;
    const uint32_t tl = tm + ((uint32_t)(quarter * 32) << 16);
    const int ncol32 = (N16 + 31) / 32;                      // 32-column chunks in all
    for (int cc = half; cc < ncol32; cc += 2) {
      float v[32];
      umma::tmem_ld32(tl + 32u * cc, v);
      if (gm < a.M) {
        float* crow = a.C + gm * a.ldc + 32 * cc;
        const float* arow = (EP == G_MASK) ? a.aux + gm * a.ldc + 32 * cc : nullptr;
        if (EP != G_ATOMIC && a.c_vec && 32 * cc + 31 < a.N) {      // 128-bit stores (and mask loads)
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            float4 o = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
            if (EP == G_BIAS_RELU) {
              const float4 b = __ldg(reinterpret_cast<const float4*>(a.bias + 32 * cc) + q);
              o = make_float4(fmaxf(o.x + b.x, 0.f), fmaxf(o.y + b.y, 0.f), fmaxf(o.z + b.z, 0.f), fmaxf(o.w + b.w, 0.f));
            } else if (EP == G_MASK) {
              const float4 m = __ldg(reinterpret_cast<const float4*>(arow) + q);
              o = make_float4(m.x > 0.f ? o.x : 0.f, m.y > 0.f ? o.y : 0.f, m.z > 0.f ? o.z : 0.f, m.w > 0.f ? o.w : 0.f);
            }
            reinterpret_cast<float4*>(crow)[q] = o;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int gn = 32 * cc + j;
            if (gn < a.N) {
              if (EP == G_BIAS_RELU) crow[j] = fmaxf(v[j] + a.bias[gn], 0.f);
              else if (EP == G_MASK) crow[j] = arow[j] > 0.f ? v[j] : 0.f;
              else if (EP == G_ATOMIC) atomicAdd(crow + j, v[j]);
              else crow[j] = v[j];
            }
          }
        }
      }
    }
  }
  umma::tc_fence_before();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tm, 256);
}

// ---------------------------------------------------------------------------------------------
// Variant for the forward layers and the input gradients: B (the weights) is the same for every row tile, so it is
// split ONCE per launch into a scratch copy that already has the operand layout (k_presplit_b), and each CTA fetches a
// K chunk of it -- hi and lo copies, 2 x N16 x 64 bytes -- with two bulk asynchronous copies (cp.async.bulk, the TMA
// engine: no registers, no split, no shared-memory stores by the threads) that complete on the stage's mbarrier.
// K chunks of 16 and FOUR stages: the copy of chunk c+3 is issued as soon as the MMAs of chunk c-1 have released its
// stage, i.e. three chunks (~1.2 us of tensor work) ahead of its use, which covers the L2 round trip of the copy (with
// two stages of 32 the tensor pipe waited ~0.7 us per chunk for its weights).  The threads only produce the A operand,
// two chunks ahead in registers.
// ---------------------------------------------------------------------------------------------
constexpr int kTK = 16;                            // K per chunk
constexpr int kTStages = 4;
constexpr uint32_t kBL = 128, kBS = 512;           // unpadded canonical layout of a 16-wide chunk (the TMA engine writes it)
constexpr uint32_t kBC = 32 * kBS;                 // one copy of a 256-row B chunk (16 KB)
constexpr uint32_t kALbo = 160, kASbo = 4 * kALbo;        // A: 4 units along K, 32 bytes of padding each (conflict-free 128-bit stores)
constexpr uint32_t kAC = 16 * kASbo;               // one copy of the 128-row A chunk
constexpr uint32_t kTStage = 2 * kAC + 2 * kBC;
constexpr uint32_t kTSmem = kTStages * kTStage + 128 + 1024;

// scratch layout: [chunk][hi: kBC | lo: kBC]; element (n, k) of chunk c at (n/8)*kBS + ((k%16)/4)*kBL + (n%8)*16 + (k%4)*4
__global__ void __launch_bounds__(256) k_presplit_b(const float* __restrict__ B, int64_t sbk, int64_t sbn, int N, int N16, int64_t K,
                                                    unsigned char* __restrict__ out) {
  const int64_t c = blockIdx.x, k0 = c * kTK;
  unsigned char* hi = out + (size_t)c * 2 * kBC;
  unsigned char* lo = hi + kBC;
  for (int i = threadIdx.x; i < N16 * kTK; i += blockDim.x) {
    int n, kk;      // the contiguous index fastest for the source reads
    if (sbk == 1) { kk = i & (kTK - 1); n = i / kTK; } else { n = i % N16; kk = i / N16; }
    const int64_t k = k0 + kk;
    const float w = (n < N && k < K) ? B[k * sbk + (int64_t)n * sbn] : 0.f;
    float h, l;
    umma::split_tf32(w, h, l);
    const uint32_t off = (uint32_t)(n >> 3) * kBS + (uint32_t)(kk >> 2) * kBL + (uint32_t)(n & 7) * 16u + (uint32_t)(kk & 3) * 4u;
    *reinterpret_cast<float*>(hi + off) = h;
    *reinterpret_cast<float*>(lo + off) = l;
  }
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(umma::smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(umma::smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(umma::smem_u32(bar))
               : "memory");
}

template <int EP, bool A_KC>
__global__ void __launch_bounds__(kGT, 1) k_tc_gemm_bt(const GemmArgs a, const unsigned char* __restrict__ Bs) {
  extern __shared__ __align__(128) unsigned char smraw[];
  uint64_t* freeb = reinterpret_cast<uint64_t*>(smraw + kTStages * kTStage);   // [4] stage free (the MMAs that read it are done)
  uint64_t* fullb = freeb + kTStages;                                           // [4] B chunk has landed
  uint32_t& tmem_base_s = *reinterpret_cast<uint32_t*>(smraw + kTStages * kTStage + 96);
  float* bias_s = reinterpret_cast<float*>(smraw + kTStages * kTStage + 128);     // [256]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t m0 = (int64_t)blockIdx.x * 128;
  const int N16 = (a.N + 15) & ~15;
  const int64_t k_end = a.K;
  if (EP == G_BIAS_RELU) bias_s[tid] = tid < a.N ? a.bias[tid] : 0.f;
  if (warp == 0) umma::tmem_alloc(&tmem_base_s, 256);
  if (tid == 0) {
    for (int i = 0; i < kTStages; ++i) { umma::mbar_init(&freeb[i], 1); umma::mbar_init(&fullb[i], 1); }
    umma::fence_mbar_init();
  }
  umma::tc_fence_before();
  __syncthreads();
  umma::tc_fence_after();
  const uint32_t tm = tmem_base_s;
  const uint32_t sbase = umma::smem_u32(smraw);
  const uint32_t idesc = umma::instr_desc_tf32(128, N16);
  constexpr uint32_t kStepA = (2u * kALbo) >> 4, kStepB = (2u * kBL) >> 4;
  const int64_t nchunks = (k_end + kTK - 1) / kTK;
  const uint32_t b_bytes = (uint32_t)(N16 >> 3) * kBS;     // bytes of one copy actually used

  // thread 0: B chunk c -> its stage, asynchronously (the stage's previous reader, chunk c-4, must have finished)
  auto fetch_b = [&](int64_t c) {
    const int st = (int)(c % kTStages);
    if (c >= kTStages) umma::mbar_wait(&freeb[st], (uint32_t)(((c / kTStages) - 1) & 1));
    unsigned char* bh = smraw + st * kTStage + 2 * kAC;
    mbar_expect_tx(&fullb[st], 2u * b_bytes);
    const unsigned char* src = Bs + (size_t)c * 2 * kBC;
    bulk_g2s(bh, src, b_bytes, &fullb[st]);
    bulk_g2s(bh + kBC, src + kBC, b_bytes, &fullb[st]);
  };
  // A chunk c -> registers: 128 rows x 16 k = 512 16-byte units, two per thread
  float4 ra0[2], ra1[2];
  auto load_a = [&](int64_t c, float4 (&r)[2]) {
    const int64_t k0 = c * kTK;
    if (A_KC) {          // unit (row = tid/4 + 64 j, kq = tid%4)
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int row = (tid >> 2) + 64 * j, kq = tid & 3;
        r[j] = ld_row4(a.A, m0 + row, a.sam, k0 + 4 * kq, k_end, m0 + row < a.M, a.a_vec);
      }
    } else {             // 4 x 4 blocks (mq = lane, kq = warp % 4); warps 0-3 take k, warps 4-7 take k + ... two k rows each
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int64_t k = k0 + 4 * (warp & 3) + 2 * (warp >> 2) + j;
        r[j] = ld_row4(a.A, k, a.sak, m0 + 4 * lane, a.M, k < k_end, a.a_vec);
      }
    }
  };
  if (tid == 0)
    for (int64_t c = 0; c < kTStages - 1 && c < nchunks; ++c) fetch_b(c);
  load_a(0, ra0);
  if (nchunks > 1) load_a(1, ra1);
  // one chunk: r holds its A values; on the way out r is refilled with chunk c+2 (the loop body is instantiated twice so that
  // the two register sets are addressed statically)
  auto chunk = [&](int64_t c, float4 (&r)[2]) {
    const int st = (int)(c % kTStages);
    const uint32_t use = (uint32_t)((c / kTStages) & 1);          // parity of this stage's current use
    unsigned char* ah = smraw + st * kTStage;
    unsigned char* al = ah + kAC;
    // the A half of the stage: free once the MMAs of chunk c-4 are done
    if (c >= kTStages) { umma::mbar_wait(&freeb[st], use ^ 1u); umma::tc_fence_after(); }
    if (A_KC) {
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int row = (tid >> 2) + 64 * j, kq = tid & 3;
        sts_split(ah, al, (uint32_t)(row >> 3) * kASbo + (uint32_t)kq * kALbo + (uint32_t)(row & 7) * 16u, r[j]);
      }
    } else {
      // this thread holds rows 4*lane..+3 at two consecutive k (k%4 = 2*(warp>>2) + {0,1}) of unit kq = warp%4: 8-byte stores
      const int kq = warp & 3, khalf = warp >> 2;
      const float x[4][2] = {{r[0].x, r[1].x}, {r[0].y, r[1].y}, {r[0].z, r[1].z}, {r[0].w, r[1].w}};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int row = 4 * lane + i;
        const uint32_t off = (uint32_t)(row >> 3) * kASbo + (uint32_t)kq * kALbo + (uint32_t)(row & 7) * 16u + (uint32_t)khalf * 8u;
        float2 h, l;
        umma::split_tf32(x[i][0], h.x, l.x); umma::split_tf32(x[i][1], h.y, l.y);
        *reinterpret_cast<float2*>(ah + off) = h;
        *reinterpret_cast<float2*>(al + off) = l;
      }
    }
    if (c + 2 < nchunks) load_a(c + 2, r);     // in flight for two iterations
    umma::fence_proxy_async();
    umma::tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      umma::mbar_wait(&fullb[st], use);
      umma::tc_fence_after();
      const uint32_t sa = sbase + (uint32_t)st * kTStage;
      const uint64_t dah = umma::smem_desc(sa, kALbo, kASbo), dal = umma::smem_desc(sa + kAC, kALbo, kASbo);
      const uint64_t dbh = umma::smem_desc(sa + 2 * kAC, kBL, kBS), dbl = umma::smem_desc(sa + 2 * kAC + kBC, kBL, kBS);
#pragma unroll
      for (int ks = 0; ks < kTK / 8; ++ks) {
        const uint64_t ahk = dah + (uint64_t)(ks * kStepA), alk = dal + (uint64_t)(ks * kStepA);
        const uint64_t bhk = dbh + (uint64_t)(ks * kStepB), blk = dbl + (uint64_t)(ks * kStepB);
        umma::mma_tf32(tm, alk, bhk, idesc, (c == 0 && ks == 0) ? 0u : 1u);
        umma::mma_tf32(tm, ahk, blk, idesc, 1u);
        umma::mma_tf32(tm, ahk, bhk, idesc, 1u);
      }
      umma::mma_commit(&freeb[st]);
      // chunk c+3's weights: its stage was read by chunk c-1, whose MMAs finish while those of chunk c (just queued) run
      if (c + kTStages - 1 < nchunks) fetch_b(c + kTStages - 1);
    }
  };
  for (int64_t c = 0; c < nchunks; c += 2) {
    chunk(c, ra0);
    if (c + 1 < nchunks) chunk(c + 1, ra1);
  }
  {   // every MMA has completed when the last commit lands (commits complete in order)
    const int64_t last = nchunks - 1;
    umma::mbar_wait(&freeb[last % kTStages], (uint32_t)((last / kTStages) & 1));
  }
  umma::tc_fence_after();
  {
    const int quarter = warp & 3, half = warp >> 2;
    const int64_t gm = m0 + quarter * 32 + lane;
    const uint32_t tl = tm + ((uint32_t)(quarter * 32) << 16);
    const int ncol32 = (N16 + 31) / 32;
    for (int cc = half; cc < ncol32; cc += 2) {
      const bool vec = a.c_vec && 32 * cc + 31 < a.N;
      const float* arow = (EP == G_MASK) ? a.aux + gm * a.ldc + 32 * cc : nullptr;
      float4 mk[8];
      if (EP == G_MASK && vec && gm < a.M) {     // the row's mask values: eight loads in flight under the tensor-memory read
#pragma unroll
        for (int q = 0; q < 8; ++q) mk[q] = __ldg(reinterpret_cast<const float4*>(arow) + q);
      }
      float v[32];
      umma::tmem_ld32(tl + 32u * cc, v);
      if (gm < a.M) {
        float* crow = a.C + gm * a.ldc + 32 * cc;
        if (vec) {
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            float4 o = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
            if (EP == G_BIAS_RELU) {
              const float4 b = *reinterpret_cast<const float4*>(bias_s + 32 * cc + 4 * q);
              o = make_float4(fmaxf(o.x + b.x, 0.f), fmaxf(o.y + b.y, 0.f), fmaxf(o.z + b.z, 0.f), fmaxf(o.w + b.w, 0.f));
            } else if (EP == G_MASK) {
              const float4 m = mk[q];
              o = make_float4(m.x > 0.f ? o.x : 0.f, m.y > 0.f ? o.y : 0.f, m.z > 0.f ? o.z : 0.f, m.w > 0.f ? o.w : 0.f);
            }
            reinterpret_cast<float4*>(crow)[q] = o;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int gn = 32 * cc + j;
            if (gn < a.N) {
              if (EP == G_BIAS_RELU) crow[j] = fmaxf(v[j] + a.bias[gn], 0.f);
              else if (EP == G_MASK) crow[j] = arow[j] > 0.f ? v[j] : 0.f;
              else crow[j] = v[j];
            }
          }
        }
      }
    }
  }
  umma::tc_fence_before();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tm, 256);
}

template <int EP>
int launch_gemm_bt(const GemmArgs& a, bool akc, cudaStream_t st) {
  const int N16 = (a.N + 15) & ~15;
  const int64_t nchunks = (a.K + kTK - 1) / kTK;
  unsigned char* scratch = nullptr;
  if (cudaMallocAsync(reinterpret_cast<void**>(&scratch), (size_t)nchunks * 2 * kBC, st) != cudaSuccess) {
    cudaGetLastError();
    set_error("tc_gemm: cudaMallocAsync of the split-weight scratch failed");
    return 1;
  }
  k_presplit_b<<<(unsigned)nchunks, 256, 0, st>>>(a.B, a.sbk, a.sbn, a.N, N16, a.K, scratch);
  int rc = launch_status("k_presplit_b");
  if (!rc) {
    const dim3 grid((unsigned)((a.M + 127) / 128));
    if (akc) {
      auto kern = k_tc_gemm_bt<EP, true>;
      cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTSmem);
      kern<<<grid, kGT, kTSmem, st>>>(a, scratch);
    } else {
      auto kern = k_tc_gemm_bt<EP, false>;
      cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTSmem);
      kern<<<grid, kGT, kTSmem, st>>>(a, scratch);
    }
    rc = launch_status("k_tc_gemm_bt");
  }
  cudaFreeAsync(scratch, st);
  return rc;
}

template <int EP>
int launch_gemm(const GemmArgs& a, bool akc, bool bkc, int splitk, cudaStream_t st) {
  const dim3 grid((unsigned)((a.M + 127) / 128), 1, EP == G_ATOMIC ? splitk : 1);
#define PN_GEMM_LAUNCH(AK, BK)                                                                                   \
  {                                                                                                              \
    auto kern = k_tc_gemm<EP, AK, BK>;                                                                           \
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGSmem);                        \
    kern<<<grid, kGT, kGSmem, st>>>(a);                                                                          \
  }
  if (akc && bkc) PN_GEMM_LAUNCH(true, true)
  else if (akc) PN_GEMM_LAUNCH(true, false)
  else if (bkc) PN_GEMM_LAUNCH(false, true)
  else PN_GEMM_LAUNCH(false, false)
#undef PN_GEMM_LAUNCH
  return launch_status("k_tc_gemm");
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace

// 0 ok, 1 error, -1 not applicable (the caller then uses the FFMA GEMM): needs N <= 256, one contiguous index per
// operand, and enough rows to fill 128-row tiles
int tc_gemm(int ep, const float* A, int64_t sam, int64_t sak, const float* B, int64_t sbk, int64_t sbn, float* C, int64_t ldc,
            int64_t M, int N, int64_t K, const float* bias, const float* aux, int splitk, float* a_rowsum, cudaStream_t st) {
  if (N > 256 || N < 16 || M < 64 || K < 32) return -1;
  if (!(sak == 1 || sam == 1) || !(sbk == 1 || sbn == 1)) return -1;
  GemmArgs a;
  a.A = A; a.sam = sam; a.sak = sak; a.B = B; a.sbk = sbk; a.sbn = sbn; a.C = C; a.ldc = ldc; a.M = M; a.K = K; a.N = N;
  a.bias = bias; a.aux = aux; a.a_rowsum = a_rowsum;
  const bool akc = sak == 1, bkc = sbk == 1;
  if (a_rowsum && (ep != G_ATOMIC || akc)) return -1;
  if (ep == G_ATOMIC) {        // split K so that the (few) row tiles fill the machine; at least 8 chunks per CTA
    const int64_t mt = (M + 127) / 128;
    int64_t s = ((int64_t)sm_count() + mt - 1) / mt;
    const int64_t smax = K / 256 > 0 ? K / 256 : 1;
    splitk = (int)(s < smax ? s : smax);
  }
  a.a_vec = aligned16(A) && ((akc ? sam : sak) % 4 == 0);
  a.b_vec = aligned16(B) && ((bkc ? sbn : sbk) % 4 == 0);
  a.c_vec = aligned16(C) && (ldc % 4 == 0) && (!aux || aligned16(aux)) && (!bias || aligned16(bias));
  switch (ep) {   // B reused by every row tile (weights): split once, fetched by the TMA engine; split-K (B = activations): register path
    case G_STORE: return launch_gemm_bt<G_STORE>(a, akc, st);
    case G_BIAS_RELU: return launch_gemm_bt<G_BIAS_RELU>(a, akc, st);
    case G_MASK: return launch_gemm_bt<G_MASK>(a, akc, st);
    default: return launch_gemm<G_ATOMIC>(a, akc, bkc, splitk < 1 ? 1 : splitk, st);
  }
}

}  // namespace pn
