// Tensor-core GEMMs for the iMAP* 256-wide MLP (decoder.MLP with c_dim 0; src/conv_onet/config.py:28-32,
// src/conv_onet/models/decoder.py:189-203): forward layers, input gradients and weight gradients.
//
//     C[M x N] (op)= sum_k A(m,k) * B(k,n),   A(m,k) = A[m*sam + k*sak],   B(k,n) = B[k*sbk + n*sbn]
//
// the same strided interface as the FFMA k_sgemm of pn_imap.cu, so the layer chain there is unchanged.  Two kernels,
// both tcgen05 kind::tf32 with the 3xTF32 split (hi = tf32(x), lo = x - hi; a.b ~ a_lo.b_hi + a_hi.b_lo + a_hi.b_hi),
// operands in shared memory in the canonical K-major no-swizzle layout, a 256-row tile = TWO 128-row accumulators
// (512 tensor-memory columns), K walked in chunks of 16:
//
//   k_tc_gemm_bt  forward layers and input gradients (M = samples).  B (the weights) is the same for every row tile:
//     it is split ONCE per launch into a scratch copy that already has the operand layout (k_presplit_b) and each CTA
//     fetches a K chunk of it -- hi and lo copies -- with two bulk asynchronous copies (cp.async.bulk, the TMA engine:
//     no registers, no split, no shared-memory stores by the threads) that complete on the stage's mbarrier, two chunks
//     ahead of its use.  The threads produce only the A operand (two chunks ahead in registers).
//   k_tc_wgrad    weight gradients: M = output features (256), N = input features, K = SAMPLES, split over CTAs and
//     finished with atomics.  Both operands arrive with their row index contiguous in memory; a thread reads a 4 x 4
//     block -- four coalesced 16-byte loads along the row index -- transposes it in registers and stores four 16-byte
//     units, so no scalar shared-memory store is issued.  The bias gradient (column sums of the gradient operand) rides
//     along in the loaders.
//
// One elected thread issues the MMAs of a chunk (2 accumulators x 2 k-steps x 3) and commits them to the stage's "free"
// mbarrier; epilogue: tensor memory -> registers (two threads per row, half of the columns each) -> bias + ReLU / ReLU
// mask / plain store / atomics -> global.
#include "pn_common.cuh"
#include "pn_umma.cuh"

namespace pn {
namespace {

constexpr int kGT = 256;                           // threads
constexpr int kTK = 16;                            // K per chunk
enum { G_STORE = 0, G_BIAS_RELU = 1, G_MASK = 2, G_ATOMIC = 3 };

struct GemmArgs {
  const float* A; int64_t sam, sak;
  const float* B; int64_t sbk, sbn;
  float* C; int64_t ldc;
  int64_t M, K;
  int N;
  const float* bias; const float* aux;
  int a_vec, b_vec;      // 16-byte loads allowed (base and strides 16-byte aligned)
  int c_vec;             // 16-byte stores to C (and loads of aux) allowed
  float* a_rowsum;       // k_tc_wgrad, optional: a_rowsum[m] += sum_k A(m,k)  (bias gradients)
};

__device__ __forceinline__ void sts_split(unsigned char* hi, unsigned char* lo, uint32_t off, float4 v) {
  float4 h, l;
  umma::split_tf32(v.x, h.x, l.x); umma::split_tf32(v.y, h.y, l.y);
  umma::split_tf32(v.z, h.z, l.z); umma::split_tf32(v.w, h.w, l.w);
  *reinterpret_cast<float4*>(hi + off) = h;
  *reinterpret_cast<float4*>(lo + off) = l;
}

// four consecutive elements along the contiguous index starting at column c of row `other` (row stride rs);
// elements beyond the limits read 0
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

__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(umma::smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(umma::smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(umma::smem_u32(bar))
               : "memory");
}

// one chunk's MMAs: both 128-row halves of the A tile against the same B chunk (2 x 2 k-steps x 3)
__device__ __forceinline__ void issue_chunk(uint32_t tm, uint32_t a_hi, uint32_t a_lo, uint32_t a_half, uint32_t a_lbo, uint32_t a_sbo,
                                            uint32_t b_hi, uint32_t b_lo, uint32_t b_lbo, uint32_t b_sbo, uint32_t idesc, bool first,
                                            bool two_halves) {
  const uint64_t dbh = umma::smem_desc(b_hi, b_lbo, b_sbo), dbl = umma::smem_desc(b_lo, b_lbo, b_sbo);
  const uint32_t sa = (2u * a_lbo) >> 4, sb = (2u * b_lbo) >> 4;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    if (h == 1 && !two_halves) break;
    const uint64_t dah = umma::smem_desc(a_hi + h * a_half, a_lbo, a_sbo), dal = umma::smem_desc(a_lo + h * a_half, a_lbo, a_sbo);
#pragma unroll
    for (int ks = 0; ks < kTK / 8; ++ks) {
      umma::mma_tf32(tm + 256u * h, dal + (uint64_t)(ks * sa), dbh + (uint64_t)(ks * sb), idesc, (first && ks == 0) ? 0u : 1u);
      umma::mma_tf32(tm + 256u * h, dah + (uint64_t)(ks * sa), dbl + (uint64_t)(ks * sb), idesc, 1u);
      umma::mma_tf32(tm + 256u * h, dah + (uint64_t)(ks * sa), dbh + (uint64_t)(ks * sb), idesc, 1u);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// forward layers / input gradients: B by the TMA engine from a pre-split copy
// ---------------------------------------------------------------------------------------------
// ROWS = rows of A per CTA: 128 (one accumulator, four stages) or 256 (two accumulators, three stages).  Measured on the
// iMAP* iteration: 128 is faster for these kernels (0.97 vs 1.12 ms per forward call: the per-tile fixed cost is amortised
// better by 256, but 625 tiles leave a 4.2-wave tail on 148 SMs and the fourth stage hides more of the weight fetch)
constexpr int kBtRows = 128;
constexpr uint32_t kBL = 128, kBS = 512;           // unpadded canonical layout of a 16-wide chunk (the TMA engine writes it)
constexpr uint32_t kBC = 32 * kBS;                 // one copy of a 256-row B chunk (16 KB)
constexpr uint32_t kALbo = 160, kASbo = 4 * kALbo; // A: 4 units along K, 32 bytes of padding each (conflict-free 128-bit stores)
constexpr uint32_t kAH = 16 * kASbo;               // one 128-row half of one copy
template <int ROWS> struct BtCfg {
  static constexpr int stages = ROWS == 128 ? 4 : 3;
  static constexpr uint32_t ac = (ROWS / 128) * kAH;            // one copy of the A chunk
  static constexpr uint32_t stage = 2 * ac + 2 * kBC;
  static constexpr uint32_t smem = stages * stage + 128 + 1024;
};

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

// the shared epilogue: 256 rows (two accumulators), N16 columns
template <int EP>
__device__ __forceinline__ void epilogue(const GemmArgs& a, uint32_t tm, int64_t m0, int halves, int N16, const float* bias_s, int warp, int lane) {
  const int quarter = warp & 3, half = warp >> 2;
  const int ncol32 = (N16 + 31) / 32;
#pragma unroll 1
  for (int h = 0; h < halves; ++h) {
    if (m0 + 128 * h >= a.M) break;
    const int64_t gm = m0 + 128 * h + quarter * 32 + lane;
    const uint32_t tl = tm + 256u * h + ((uint32_t)(quarter * 32) << 16);
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
        if (EP != G_ATOMIC && vec) {
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
              if (EP == G_BIAS_RELU) crow[j] = fmaxf(v[j] + bias_s[gn], 0.f);
              else if (EP == G_MASK) crow[j] = arow[j] > 0.f ? v[j] : 0.f;
              else if (EP == G_ATOMIC) atomicAdd(crow + j, v[j]);
              else crow[j] = v[j];
            }
          }
        }
      }
    }
  }
}

template <int EP, bool A_KC, int ROWS>
__global__ void __launch_bounds__(kGT, 1) k_tc_gemm_bt(const GemmArgs a, const unsigned char* __restrict__ Bs) {
  constexpr int kTStages = BtCfg<ROWS>::stages;
  constexpr uint32_t kAC = BtCfg<ROWS>::ac, kTStage = BtCfg<ROWS>::stage;
  constexpr int kUnits = ROWS / 64;                  // 16-byte A units per thread and chunk
  extern __shared__ __align__(128) unsigned char smraw[];
  uint64_t* freeb = reinterpret_cast<uint64_t*>(smraw + kTStages * kTStage);   // [stages] the MMAs that read the stage are done
  uint64_t* fullb = freeb + kTStages;                                           // [stages] the B chunk has landed
  uint32_t& tmem_base_s = *reinterpret_cast<uint32_t*>(smraw + kTStages * kTStage + 96);
  float* bias_s = reinterpret_cast<float*>(smraw + kTStages * kTStage + 128);     // [256]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t m0 = (int64_t)blockIdx.x * ROWS;
  const int N16 = (a.N + 15) & ~15;
  const int64_t k_end = a.K;
  const bool two = ROWS == 256 && m0 + 128 < a.M;
  if (EP == G_BIAS_RELU) bias_s[tid] = tid < a.N ? a.bias[tid] : 0.f;
  if (warp == 0) umma::tmem_alloc(&tmem_base_s, ROWS == 256 ? 512 : 256);
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
  const int64_t nchunks = (k_end + kTK - 1) / kTK;
  const uint32_t b_bytes = (uint32_t)(N16 >> 3) * kBS;     // bytes of one copy actually used

  // thread 0: B chunk c -> its stage, asynchronously (the stage's previous reader, chunk c - stages, must have finished)
  auto fetch_b = [&](int64_t c) {
    const int st = (int)(c % kTStages);
    if (c >= kTStages) umma::mbar_wait(&freeb[st], (uint32_t)(((c / kTStages) - 1) & 1));
    unsigned char* bh = smraw + st * kTStage + 2 * kAC;
    mbar_expect_tx(&fullb[st], 2u * b_bytes);
    const unsigned char* src = Bs + (size_t)c * 2 * kBC;
    bulk_g2s(bh, src, b_bytes, &fullb[st]);
    bulk_g2s(bh + kBC, src + kBC, b_bytes, &fullb[st]);
  };
  // A chunk c -> registers: ROWS x 16 k = ROWS*4 16-byte units, kUnits per thread
  float4 ra0[4], ra1[4];
  auto load_a = [&](int64_t c, float4 (&r)[4]) {
    const int64_t k0 = c * kTK;
    if (A_KC) {          // unit (row = tid/4 + 64 j, kq = tid%4)
#pragma unroll
      for (int j = 0; j < kUnits; ++j) {
        const int row = (tid >> 2) + 64 * j, kq = tid & 3;
        r[j] = ld_row4(a.A, m0 + row, a.sam, k0 + 4 * kq, k_end, m0 + row < a.M, a.a_vec);
      }
    } else if (ROWS == 256) {   // 4 x 4 blocks: (row quad = lane + 32 (warp/4), kq = warp%4): rows 4 rq .. +3, k = 4 kq + j
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int64_t k = k0 + 4 * (warp & 3) + j;
        r[j] = ld_row4(a.A, k, a.sak, m0 + 4 * (lane + 32 * (warp >> 2)), a.M, k < k_end, a.a_vec);
      }
    } else {                    // 128 rows: 4 x 2 blocks (row quad = lane, kq = warp%4, k pair = warp/4)
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
  // one chunk: r holds its A values; on the way out r is refilled with chunk c+2 (the body is instantiated twice so that the
  // two register sets are addressed statically)
  auto chunk = [&](int64_t c, float4 (&r)[4]) {
    const int st = (int)(c % kTStages);
    const uint32_t use = (uint32_t)((c / kTStages) & 1);          // parity of this stage's current use
    unsigned char* ah = smraw + st * kTStage;
    unsigned char* al = ah + kAC;
    if (c >= kTStages) { umma::mbar_wait(&freeb[st], use ^ 1u); umma::tc_fence_after(); }
    if (A_KC) {
#pragma unroll
      for (int j = 0; j < kUnits; ++j) {
        const int row = (tid >> 2) + 64 * j, kq = tid & 3;
        sts_split(ah, al, (uint32_t)(row >> 3) * kASbo + (uint32_t)kq * kALbo + (uint32_t)(row & 7) * 16u, r[j]);
      }
    } else if (ROWS == 256) {
      const float x[4][4] = {{r[0].x, r[1].x, r[2].x, r[3].x}, {r[0].y, r[1].y, r[2].y, r[3].y},
                             {r[0].z, r[1].z, r[2].z, r[3].z}, {r[0].w, r[1].w, r[2].w, r[3].w}};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int row = 4 * (lane + 32 * (warp >> 2)) + i;
        sts_split(ah, al, (uint32_t)(row >> 3) * kASbo + (uint32_t)(warp & 3) * kALbo + (uint32_t)(row & 7) * 16u,
                  make_float4(x[i][0], x[i][1], x[i][2], x[i][3]));
      }
    } else {
      const float x[4][2] = {{r[0].x, r[1].x}, {r[0].y, r[1].y}, {r[0].z, r[1].z}, {r[0].w, r[1].w}};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int row = 4 * lane + i;
        const uint32_t off = (uint32_t)(row >> 3) * kASbo + (uint32_t)(warp & 3) * kALbo + (uint32_t)(row & 7) * 16u + (uint32_t)(warp >> 2) * 8u;
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
      issue_chunk(tm, sa, sa + kAC, kAH, kALbo, kASbo, sa + 2 * kAC, sa + 2 * kAC + kBC, kBL, kBS, idesc, c == 0, two);
      umma::mma_commit(&freeb[st]);
      // chunk c+2's weights: its stage was read by chunk c-1, whose MMAs finish while those of chunk c (just queued) run
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
  epilogue<EP>(a, tm, m0, ROWS / 128, N16, bias_s, warp, lane);
  umma::tc_fence_before();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tm, ROWS == 256 ? 512 : 256);
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
    const dim3 grid((unsigned)((a.M + kBtRows - 1) / kBtRows));
    constexpr uint32_t smem = BtCfg<kBtRows>::smem;
    if (akc) {
      auto kern = k_tc_gemm_bt<EP, true, kBtRows>;
      cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      kern<<<grid, kGT, smem, st>>>(a, scratch);
    } else {
      auto kern = k_tc_gemm_bt<EP, false, kBtRows>;
      cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      kern<<<grid, kGT, smem, st>>>(a, scratch);
    }
    rc = launch_status("k_tc_gemm_bt");
  }
  cudaFreeAsync(scratch, st);
  return rc;
}

// ---------------------------------------------------------------------------------------------
// weight gradients: both operands row-contiguous (sam == 1, sbn == 1), K = samples, split over CTAs
// ---------------------------------------------------------------------------------------------
constexpr uint32_t kWLbo = 144, kWSbo = 4 * kWLbo + 16;    // padding keeps the 4 x 4 block stores conflict-free
constexpr uint32_t kWC = 32 * kWSbo;                        // one copy of a 256-row operand chunk
constexpr uint32_t kWH = 16 * kWSbo;                        // its 128-row half
constexpr uint32_t kWStage = 4 * kWC;                       // A hi/lo, B hi/lo
constexpr int kWStages = 2;
constexpr uint32_t kWSmem = kWStages * kWStage + 64;

__global__ void __launch_bounds__(kGT, 1) k_tc_wgrad(const GemmArgs a) {
  extern __shared__ __align__(128) unsigned char smraw[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smraw + kWStages * kWStage);    // [2] stage free
  uint32_t& tmem_base_s = *reinterpret_cast<uint32_t*>(smraw + kWStages * kWStage + 32);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int N16 = (a.N + 15) & ~15;
  const int64_t per = ((a.K + gridDim.x - 1) / gridDim.x + kTK - 1) / kTK * kTK;
  const int64_t k_begin = (int64_t)blockIdx.x * per;
  const int64_t k_end = k_begin + per < a.K ? k_begin + per : a.K;
  if (k_begin >= k_end) return;
  const bool two = a.M > 128;
  if (warp == 0) umma::tmem_alloc(&tmem_base_s, 512);
  if (tid == 0) { umma::mbar_init(&bars[0], 1); umma::mbar_init(&bars[1], 1); umma::fence_mbar_init(); }
  umma::tc_fence_before();
  __syncthreads();
  umma::tc_fence_after();
  const uint32_t tm = tmem_base_s;
  const uint32_t sbase = umma::smem_u32(smraw);
  const uint32_t idesc = umma::instr_desc_tf32(128, N16);
  uint32_t phase[2] = {0u, 0u};
  const int64_t nchunks = (k_end - k_begin + kTK - 1) / kTK;
  // 4 x 4 blocks: row quad rq = lane + 32 (warp/4) (rows 4 rq .. +3 of the 256), k quad kq = warp%4 (k = 4 kq + j)
  const int rq = lane + 32 * (warp >> 2), kq = warp & 3;
  float4 ra[4], rb[4];
  float4 rsum = make_float4(0.f, 0.f, 0.f, 0.f);
  auto load_chunk = [&](int64_t c) {
    const int64_t k0 = k_begin + c * kTK + 4 * kq;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      ra[j] = ld_row4(a.A, k0 + j, a.sak, 4 * rq, a.M, k0 + j < k_end, a.a_vec);
      rb[j] = (4 * rq < N16) ? ld_row4(a.B, k0 + j, a.sbk, 4 * rq, a.N, k0 + j < k_end, a.b_vec) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  auto store_block = [&](unsigned char* hi, unsigned char* lo, const float4 (&r)[4], float4* sum) {
    const float x[4][4] = {{r[0].x, r[1].x, r[2].x, r[3].x}, {r[0].y, r[1].y, r[2].y, r[3].y},
                           {r[0].z, r[1].z, r[2].z, r[3].z}, {r[0].w, r[1].w, r[2].w, r[3].w}};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int row = 4 * rq + i;
      sts_split(hi, lo, (uint32_t)(row >> 3) * kWSbo + (uint32_t)kq * kWLbo + (uint32_t)(row & 7) * 16u,
                make_float4(x[i][0], x[i][1], x[i][2], x[i][3]));
    }
    if (sum) {
      sum->x += (x[0][0] + x[0][1]) + (x[0][2] + x[0][3]); sum->y += (x[1][0] + x[1][1]) + (x[1][2] + x[1][3]);
      sum->z += (x[2][0] + x[2][1]) + (x[2][2] + x[2][3]); sum->w += (x[3][0] + x[3][1]) + (x[3][2] + x[3][3]);
    }
  };
  load_chunk(0);
  for (int64_t c = 0; c < nchunks; ++c) {
    const int st = (int)(c & 1);
    unsigned char* base = smraw + st * kWStage;
    if (c >= 2) { umma::mbar_wait(&bars[st], phase[st]); phase[st] ^= 1u; umma::tc_fence_after(); }
    store_block(base, base + kWC, ra, &rsum);
    if (4 * rq < N16) store_block(base + 2 * kWC, base + 3 * kWC, rb, nullptr);
    if (c + 1 < nchunks) load_chunk(c + 1);     // in flight under the barrier, the MMA issue and the next wait
    umma::fence_proxy_async();
    umma::tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      umma::tc_fence_after();
      const uint32_t sa = sbase + (uint32_t)st * kWStage;
      issue_chunk(tm, sa, sa + kWC, kWH, kWLbo, kWSbo, sa + 2 * kWC, sa + 3 * kWC, kWLbo, kWSbo, idesc, c == 0, two);
      umma::mma_commit(&bars[st]);
    }
  }
  {
    const int s1 = (int)((nchunks - 1) & 1);
    umma::mbar_wait(&bars[s1], phase[s1]); phase[s1] ^= 1u;
  }
  umma::tc_fence_after();
  if (a.a_rowsum) {      // rows 4 rq .. of the tile, partial over this thread's k slots
    const int64_t m = 4 * rq;
    if (m < a.M) atomicAdd(a.a_rowsum + m, rsum.x);
    if (m + 1 < a.M) atomicAdd(a.a_rowsum + m + 1, rsum.y);
    if (m + 2 < a.M) atomicAdd(a.a_rowsum + m + 2, rsum.z);
    if (m + 3 < a.M) atomicAdd(a.a_rowsum + m + 3, rsum.w);
  }
  epilogue<G_ATOMIC>(a, tm, 0, 2, N16, nullptr, warp, lane);
  umma::tc_fence_before();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tm, 512);
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace

// 0 ok, 1 error, -1 not applicable (the caller then uses the FFMA GEMM).  ep 0..2 (store, bias+ReLU, ReLU mask): M = samples,
// N <= 256, one contiguous index per operand.  ep 3 (split-K atomics, weight gradients): M <= 256 output features, both
// operands row-contiguous, K = samples; a_rowsum (optional) receives the column sums of the gradient operand.
int tc_gemm(int ep, const float* A, int64_t sam, int64_t sak, const float* B, int64_t sbk, int64_t sbn, float* C, int64_t ldc,
            int64_t M, int N, int64_t K, const float* bias, const float* aux, int splitk, float* a_rowsum, cudaStream_t st) {
  (void)splitk;
  if (N > 256 || N < 16 || M < 64 || K < 32) return -1;
  GemmArgs a;
  a.A = A; a.sam = sam; a.sak = sak; a.B = B; a.sbk = sbk; a.sbn = sbn; a.C = C; a.ldc = ldc; a.M = M; a.K = K; a.N = N;
  a.bias = bias; a.aux = aux; a.a_rowsum = a_rowsum;
  a.c_vec = aligned16(C) && (ldc % 4 == 0) && (!aux || aligned16(aux));
  if (ep == G_ATOMIC) {
    if (sam != 1 || sbn != 1 || M > 256) return -1;
    a.a_vec = aligned16(A) && (sak % 4 == 0);
    a.b_vec = aligned16(B) && (sbk % 4 == 0);
    // split K so that the CTAs fill the machine, at least 16 chunks each
    int64_t s = sm_count();
    const int64_t smax = K / (16 * kTK) > 0 ? K / (16 * kTK) : 1;
    if (s > smax) s = smax;
    cudaFuncSetAttribute(k_tc_wgrad, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kWSmem);
    k_tc_wgrad<<<(unsigned)s, kGT, kWSmem, st>>>(a);
    return launch_status("k_tc_wgrad");
  }
  if (a_rowsum) return -1;
  if (!(sak == 1 || sam == 1) || !(sbk == 1 || sbn == 1)) return -1;
  const bool akc = sak == 1;
  a.a_vec = aligned16(A) && ((akc ? sam : sak) % 4 == 0);
  a.b_vec = 0;
  switch (ep) {
    case G_STORE: return launch_gemm_bt<G_STORE>(a, akc, st);
    case G_BIAS_RELU: return launch_gemm_bt<G_BIAS_RELU>(a, akc, st);
    default: return launch_gemm_bt<G_MASK>(a, akc, st);
  }
}

}  // namespace pn
