// Weight gradients of a grid decoder on the tensor cores.
//
// dW[j][f] = sum_n G[n][j] * X[n][f] is a GEMM whose reduction runs over SAMPLES: the sample
// axis is the MMA K dimension, features (M = 128) and outputs (N = 32 per block) are the rows of the two
// operands.  The operands are read from the planar-4 stash ([feature/4][sample] float4, coalesced
// 256-byte rows per 16-sample chunk), split into TF32 hi/lo in registers and stored TRANSPOSED
// into the canonical K-major no-swizzle layout
//     element (row r, sample k) at (r/8)*SBO + (k/4)*LBO + (r%8)*16 + (k%4)*4   (bytes)
// with LBO = 144 (a 128-byte core matrix + 16 bytes of padding, which keeps the scalar transposing
// stores of a half-warp -- 16 consecutive samples of one row -- on distinct banks) and SBO = 4*LBO.
// (The MN-major operand mode, which would take the stash rows unchanged, returns zeros for
// kind::tf32 on sm_100a -- measured with pn_tc_selftest_mn -- so it is not used.)
//
// Per 16-sample chunk (K = 16 = 2 MMA k-steps), M = 128 features, 3xTF32, one wide-N product per A operand:
//     A_H = [h0|h1|h2|h3]   x [GA_3|GA_1|GA_2|GA_4] (N = 128) -> dW3[:, 93:], dW1, dW2, dW4 (diagonal 32-row blocks)
//     A_E = [emb(96)|0(32)] x [GA_0|GA_3]           (N = 64)  -> dW0, dW3[:, :93]
//     A_C = [c(CD)|0]       x [GH_0..GH_4]          (N = 160) -> dWc_0..4
// 11 FP32 accumulators (352 tensor-memory columns) stay resident for the whole kernel; each CTA
// flushes them once with atomics.  Two operand stages alternate: while the tensor core works on
// chunk i, the 512 threads load / split / transpose chunk i+1 (an mbarrier per stage, armed by
// tcgen05.commit, says when a stage may be overwritten); chunks i+2.. are on their way into L2
// (prefetch.global.L2).  Bias gradients (column sums of GA_l /
// GH_l) are accumulated by the loading threads on the side.
#include "pn_common.cuh"
#include "pn_umma.cuh"

namespace pn {
namespace {

constexpr int kChunk = 16;                     // samples per stage
constexpr int kWgThreads = 512;
constexpr uint32_t kLboW = 144;                // padded core matrix
constexpr uint32_t kSboW = 4 * kLboW;          // 4 core matrices = 16 samples per 8-row group
constexpr uint32_t kACopy = 16 * kSboW;        // one copy of a 128-feature operand
constexpr uint32_t kBCopy = 4 * kSboW;         // one copy of a 32-output operand
constexpr uint32_t O_AH = 0, O_AE = O_AH + 2 * kACopy, O_AC = O_AE + 2 * kACopy, O_B = O_AC + 2 * kACopy;
// B operands: 10 hi copies back to back, then the 10 lo copies, in the order
// GA_0, GA_3, GA_1, GA_2, GA_4, GH_0..GH_4, so that the operands one A matrix multiplies are
// contiguous rows and a single wide-N MMA covers them: A_E x [GA_0|GA_3] (N=64),
// A_H x [GA_3|GA_1|GA_2|GA_4] (N=128), A_C x [GH_0..GH_4] (N=160).
constexpr uint32_t kBLo = 10 * kBCopy;
constexpr uint32_t kStage = O_B + 2 * kBLo;
constexpr uint32_t O_END = 2 * kStage;
constexpr uint32_t kSmem = O_END + 64;

struct WgTcArgs {
  const float* H; const float* C; const float* E; const float* GA; const float* GH;
  float* W[5]; float* b[5]; float* Wc[5]; float* bc[5];
  int64_t N;
  int cd;
};

// rows 4q..4q+3 (the four components of v), sample slot s -> transposed K-major stores
__device__ __forceinline__ void put_split(unsigned char* hi, uint32_t copy, int q, int s, float4 v) {
  const float x[4] = {v.x, v.y, v.z, v.w};
  const uint32_t base = (uint32_t)(q >> 1) * kSboW + (uint32_t)(s >> 2) * kLboW + (uint32_t)((q & 1) * 4) * 16u + (uint32_t)(s & 3) * 4u;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float h, l;
    umma::split_tf32(x[i], h, l);
    *reinterpret_cast<float*>(hi + base + i * 16) = h;
    *reinterpret_cast<float*>(hi + copy + base + i * 16) = l;
  }
}

// D (+)= A . B^T over K = 16 (two k-steps), 3xTF32, descriptors of the first k-step given
__device__ __forceinline__ void mma_3xtf32_k16(uint32_t d, uint64_t a_hi, uint64_t a_lo, uint64_t b_hi, uint64_t b_lo, uint32_t step,
                                               uint32_t idesc, uint32_t accumulate) {
#pragma unroll
  for (int ks = 0; ks < 2; ++ks) {
    const uint64_t ah = a_hi + (uint64_t)(ks * step), al = a_lo + (uint64_t)(ks * step);
    const uint64_t bh = b_hi + (uint64_t)(ks * step), bl = b_lo + (uint64_t)(ks * step);
    umma::mma_tf32(d, al, bh, idesc, ks == 0 ? accumulate : 1u);
    umma::mma_tf32(d, ah, bl, idesc, 1u);
    umma::mma_tf32(d, ah, bh, idesc, 1u);
  }
}

__global__ void __launch_bounds__(kWgThreads, 1) k_wgrad_tc(const WgTcArgs a) {
  extern __shared__ __align__(128) unsigned char smraw[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smraw + O_END);      // one per stage
  uint32_t& tmem_base_s = *reinterpret_cast<uint32_t*>(smraw + O_END + 16);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int slot = tid >> 4, s = tid & 15;       // 32 quad slots x 16 samples
  const int64_t N = a.N;
  const int cq = a.cd / 4;                       // feature quads of c
  if (warp == 0) umma::tmem_alloc(&tmem_base_s, 512);
  if (tid == 0) { umma::mbar_init(&bars[0], 1); umma::mbar_init(&bars[1], 1); umma::fence_mbar_init(); }
  // zero the padding rows once (A_E rows 96..127, A_C rows cd..127), both stages, hi and lo copies
  for (int st = 0; st < 2; ++st)
    for (int i = tid; i < 2 * (int)kACopy / 16; i += kWgThreads) {
      reinterpret_cast<float4*>(smraw + st * kStage + O_AE)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      reinterpret_cast<float4*>(smraw + st * kStage + O_AC)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  umma::fence_proxy_async();
  umma::tc_fence_before();
  __syncthreads();
  umma::tc_fence_after();
  const uint32_t tm = tmem_base_s;
  const uint32_t sW = umma::smem_u32(smraw);
  constexpr uint32_t kStep = (2u * kLboW) >> 4;  // one K-step = 8 samples = two core matrices
  uint32_t phase[2] = {0u, 0u};
  // bias partial sums: this thread loads G row (r*32 + slot) for r = 0..2 -> array (row / 8), quad (row % 8)
  float4 bsum[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) bsum[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
  const float4* H4 = reinterpret_cast<const float4*>(a.H);
  const float4* E4 = reinterpret_cast<const float4*>(a.E);
  const float4* C4 = reinterpret_cast<const float4*>(a.C);
  const float4* GA4 = reinterpret_cast<const float4*>(a.GA);
  const float4* GH4 = reinterpret_cast<const float4*>(a.GH);
  const int64_t nchunks = (N + kChunk - 1) / kChunk;
  // this thread's share of one chunk: up to six 16-byte loads (coalesced 256-byte rows per half-warp).
  // The rows of a chunk lie megabytes apart, so their DRAM latency is what a lock-step iteration would
  // wait for; an L2 prefetch three chunks ahead (no register, no scoreboard, not ordered by the proxy
  // fence below) turns those loads into L2 hits.
  auto prefetch_chunk = [&](int64_t c) {
    const int64_t n = c * kChunk + s;
    if (c >= nchunks || n >= N) return;
    auto pf = [](const float4* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); };
    pf(H4 + (int64_t)slot * N + n);
    if (slot < 24) pf(E4 + (int64_t)slot * N + n);
    if (slot < cq) pf(C4 + (int64_t)slot * N + n);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const int row = i * 32 + slot;
      if (row < 80) pf(row < 40 ? GA4 + (int64_t)row * N + n : GH4 + (int64_t)(row - 40) * N + n);
    }
  };
  constexpr int kAhead = 3;
  for (int d = 1; d < kAhead; ++d) prefetch_chunk(blockIdx.x + (int64_t)d * gridDim.x);
  int it = 0;
  for (int64_t c = blockIdx.x; c < nchunks; c += gridDim.x, ++it) {
    const int st = it & 1;
    unsigned char* base = smraw + st * kStage;
    const int64_t n = c * kChunk + s;
    const bool ok = n < N;
    prefetch_chunk(c + (int64_t)kAhead * gridDim.x);
    // global loads first (they do not touch shared memory), then wait for the stage to be free
    const float4 vh = ok ? H4[(int64_t)slot * N + n] : z4;                       // [h0|h1|h2|h3]: quad = slot
    const float4 ve = (ok && slot < 24) ? E4[(int64_t)slot * N + n] : z4;
    const float4 vc = (ok && slot < cq) ? C4[(int64_t)slot * N + n] : z4;
    float4 vg[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const int row = r * 32 + slot;             // 0..79: GA_0..4 (rows 0..39), GH_0..4 (40..79)
      vg[r] = z4;
      if (ok && row < 80) vg[r] = row < 40 ? GA4[(int64_t)row * N + n] : GH4[(int64_t)(row - 40) * N + n];
    }
    if (it >= 2) { umma::mbar_wait(&bars[st], phase[st]); phase[st] ^= 1u; umma::tc_fence_after(); }
    put_split(base + O_AH, kACopy, slot, s, vh);
    if (slot < 24) put_split(base + O_AE, kACopy, slot, s, ve);
    if (slot < cq) put_split(base + O_AC, kACopy, slot, s, vc);
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const int row = r * 32 + slot;
      if (row < 80) {
        const int arr = row >> 3;                                        // GA_l = l, GH_l = 5 + l
        const int pos = arr == 1 ? 2 : arr == 2 ? 3 : arr == 3 ? 1 : arr;  // position in shared memory
        put_split(base + O_B + (uint32_t)pos * kBCopy, kBLo, row & 7, s, vg[r]);
        bsum[r].x += vg[r].x; bsum[r].y += vg[r].y; bsum[r].z += vg[r].z; bsum[r].w += vg[r].w;
      }
    }
    umma::fence_proxy_async();
    umma::tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      umma::tc_fence_after();
      const uint32_t acc = it == 0 ? 0u : 1u;
      const uint32_t sb = sW + (uint32_t)st * kStage;
      auto desc = [&](uint32_t off) { return umma::smem_desc(sb + off, kLboW, kSboW); };
      auto group = [&](uint32_t dcol, uint32_t a_off, uint32_t b_pos, int n) {
        const uint32_t b_off = O_B + b_pos * kBCopy;
        mma_3xtf32_k16(tm + dcol, desc(a_off), desc(a_off + kACopy), desc(b_off), desc(b_off + kBLo), kStep,
                       umma::instr_desc_tf32(128, n), acc);
      };
      group(0u, O_AH, 1, 128);     // cols   0..127 : A_H x [GA_3|GA_1|GA_2|GA_4]
      group(128u, O_AE, 0, 64);    // cols 128..191 : A_E x [GA_0|GA_3]
      group(192u, O_AC, 5, 160);   // cols 192..351 : A_C x [GH_0..GH_4]
      umma::mma_commit(&bars[st]);
    }
  }
  // drain: the last commit on each used stage
  if (it >= 1) { const int st = (it - 1) & 1; umma::mbar_wait(&bars[st], phase[st]); phase[st] ^= 1u; }
  if (it >= 2) { const int st = it & 1; umma::mbar_wait(&bars[st], phase[st]); phase[st] ^= 1u; }
  umma::tc_fence_after();
  if (it > 0) {
    // ---- flush: thread = feature row (TMEM lane); warp w takes accumulators (w / 4) + 4 * i
    const int f = (warp & 3) * 32 + lane;
    const uint32_t tl = tm + ((uint32_t)((warp & 3) * 32) << 16);
    for (int ai = warp >> 2; ai < 11; ai += 4) {
      float v[32];
      umma::tmem_ld32(tl + 32u * ai, v);
      float* dst = nullptr;   // dW[j*ld + col]
      int ld = 0, col = -1;
      if (ai < 4) {           // A_H x GA_l with l = 3,1,2,4: useful rows are h_{l-1} = features 32*(l-1) ..
        const int l = ai == 0 ? 3 : ai == 3 ? 4 : ai;
        if ((f >> 5) == l - 1) { dst = a.W[l]; ld = l == 3 ? PN_EMBED + 32 : 32; col = (l == 3 ? PN_EMBED : 0) + (f & 31); }
      } else if (ai == 4) { if (f < PN_EMBED) { dst = a.W[0]; ld = PN_EMBED; col = f; } }
      else if (ai == 5) { if (f < PN_EMBED) { dst = a.W[3]; ld = PN_EMBED + 32; col = f; } }
      else { if (f < a.cd) { dst = a.Wc[ai - 6]; ld = a.cd; col = f; } }
      if (dst) {
#pragma unroll
        for (int j = 0; j < 32; ++j) atomicAdd(dst + (int64_t)j * ld + col, v[j]);
      }
    }
    // ---- biases: reduce each partial sum over the 16 sample lanes of the half-warp
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      float4 t = bsum[r];
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) {
        t.x += __shfl_xor_sync(kFull, t.x, o); t.y += __shfl_xor_sync(kFull, t.y, o);
        t.z += __shfl_xor_sync(kFull, t.z, o); t.w += __shfl_xor_sync(kFull, t.w, o);
      }
      const int row = r * 32 + slot;
      if (s == 0 && row < 80) {
        const int arr = row >> 3, q = row & 7;
        float* dst = arr < 5 ? a.b[arr] : a.bc[arr - 5];
        atomicAdd(dst + 4 * q, t.x); atomicAdd(dst + 4 * q + 1, t.y); atomicAdd(dst + 4 * q + 2, t.z); atomicAdd(dst + 4 * q + 3, t.w);
      }
    }
  }
  umma::tc_fence_before();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem_base_s, 512);
}

}  // namespace

// Called by pn_grid_mlp_wgrad.  Requires every W / b / Wc / bc sink (the common case); returns -1
// if the tensor-core path does not apply so that the caller uses the FFMA GEMM instead.
int launch_wgrad_tc(int64_t N, int c_dim, const float* H, const float* C, const float* E, const float* GA, const float* GH,
                    float* const* W, float* const* b, float* const* Wc, float* const* bc, cudaStream_t st) {
  for (int l = 0; l < 5; ++l)
    if (!W[l] || !b[l] || !Wc[l] || !bc[l]) return -1;
  WgTcArgs a;
  a.H = H; a.C = C; a.E = E; a.GA = GA; a.GH = GH; a.N = N; a.cd = c_dim;
  for (int l = 0; l < 5; ++l) { a.W[l] = W[l]; a.b[l] = b[l]; a.Wc[l] = Wc[l]; a.bc[l] = bc[l]; }
  cudaFuncSetAttribute(k_wgrad_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem);
  const int64_t nchunks = (N + kChunk - 1) / kChunk;
  const int grid = (int)(nchunks < (int64_t)sm_count() ? nchunks : (int64_t)sm_count());
  k_wgrad_tc<<<grid, kWgThreads, kSmem, st>>>(a);
  return launch_status("k_wgrad_tc");
}

}  // namespace pn
