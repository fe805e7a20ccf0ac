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
// Two kernels share this scheme: k_wgrad_tc32 (further down; c_dim 32, every parameter gradient, warp-specialised
// producers / issuer) is the one the mapping iteration uses; k_wgrad_tc (below; any c_dim, W / b / Wc / bc only,
// lock-step stages, followed by k_wgrad_out and k_wgrad_B of pn_gridmlp.cu) remains for c_dim 64.
//
// k_wgrad_tc, per 16-sample chunk (K = 16 = 2 MMA k-steps), M = 128 features, 3xTF32, one wide-N product per A operand:
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


// ---------------------------------------------------------------------------------------------
// c_dim = 32: ONE kernel for every parameter gradient of the decoder.
//
// Three 128-row feature operands and 12 accumulators (368 tensor-memory columns):
//     A_H  = [h0|h1|h2|h3]        x [GA_3|GA_1|GA_2|GA_4]  (N = 128) -> dW3[:, 93:], dW1, dW2, dW4
//     A_EC = [emb(96)|c(32)]      x [GA_0|GA_3]            (N = 64)  -> dW0, dW3[:, :93]      (rows 0..95)
//                                 x [GH_0..GH_4]           (N = 160) -> dWc_0..4              (rows 96..127)
//     A_G  = [garg(96)|h4(32)]    x [go(4)|p(3)|0]         (N = 16)  -> dB (rows 0..92), dWo (rows 96..127)
// GA_l is not read from memory: it is GH_l masked with the forward's ReLU bits, so the backward
// kernel stashes GH only.  Per 16-sample chunk a thread issues five 16-byte loads.
constexpr uint32_t P_AH = 0, P_AEC = P_AH + 2 * kACopy, P_AG = P_AEC + 2 * kACopy, P_B = P_AG + 2 * kACopy;
constexpr uint32_t P_B2 = P_B + 2 * kBLo;            // [16 x 16] operand: hi copy, lo copy (2 row groups each)
constexpr uint32_t kB2Copy = 2 * kSboW;
constexpr uint32_t kStage32 = P_B2 + 2 * kB2Copy;
constexpr uint32_t P_END = 2 * kStage32;
constexpr uint32_t kSmem32 = P_END + 64;   // 5 mbarriers + the TMEM base address

struct WgTc32Args {
  const float* H; const float* C; const float* E; const float* GH; const float* GARG; const float* GO; const float* P32;
  const uint32_t* bits;
  float* W[5]; float* b[5]; float* Wc[5]; float* bc[5]; float* Wo; float* bo; float* B;
  int64_t N;
  int nout;
};

// Warp-specialised: two producer groups of 256 threads (16 quad slots x 16 samples, two slots per thread) own one
// operand stage each and take alternate chunks; warp 16 is the MMA issuer.  A group signals "stage full" on an
// mbarrier (256 arrivals, each after its own proxy fence), the issuer multiplies and commits to the group's "stage
// empty" mbarrier.  While one group waits for its loads the other splits and stores, so load latency, conversion
// and the tensor pipe overlap without a block-wide barrier.
constexpr int kWg32Threads = kWgThreads + 32;

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(umma::smem_u32(bar)) : "memory");
}

__global__ void __launch_bounds__(kWg32Threads, 1) k_wgrad_tc32(const WgTc32Args a) {
  extern __shared__ __align__(128) unsigned char smraw[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smraw + P_END);          // [2]
  uint64_t* empty = full + 2;                                            // [2]
  uint64_t* done = full + 4;
  uint32_t& tmem_base_s = *reinterpret_cast<uint32_t*>(smraw + P_END + 48);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t N = a.N;
  pdl_launch_dependents();
  if (warp == 0) umma::tmem_alloc(&tmem_base_s, 512);
  if (tid == 0) {
    for (int i = 0; i < 2; ++i) { umma::mbar_init(&full[i], 256); umma::mbar_init(&empty[i], 1); }
    umma::mbar_init(done, 1);
    umma::fence_mbar_init();
  }
  // rows 7..15 of the [go|p] operand stay zero: clear both stages once
  for (int st = 0; st < 2; ++st)
    for (int i = tid; i < 2 * (int)kB2Copy / 16; i += kWg32Threads)
      reinterpret_cast<float4*>(smraw + st * kStage32 + P_B2)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  umma::fence_proxy_async();
  umma::tc_fence_before();
  __syncthreads();
  umma::tc_fence_after();
  pdl_wait();   // the stash written by the backward kernel is read from here on
  const uint32_t tm = tmem_base_s;
  const int64_t nchunks = (N + kChunk - 1) / kChunk;
  const int64_t mine = (nchunks - blockIdx.x + gridDim.x - 1) / gridDim.x;   // chunks of this CTA: blockIdx.x + k*gridDim.x
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 s_ga[2] = {z4, z4}, s_gh[2] = {z4, z4}, s_ga4 = z4, s_gh4 = z4, s_go = z4;   // bias partial sums
  const int grp = (tid >> 8) & 1, s = tid & 15, slot0 = (tid >> 4) & 15;

  if (warp == 16) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      const uint32_t sW = umma::smem_u32(smraw);
      constexpr uint32_t kStep = (2u * kLboW) >> 4;
      uint32_t ph[2] = {0u, 0u};
      for (int64_t k = 0; k < mine; ++k) {
        const int st = (int)(k & 1);
        umma::mbar_wait(&full[st], ph[st]); ph[st] ^= 1u;
        umma::tc_fence_after();
        const uint32_t acc = k == 0 ? 0u : 1u;
        const uint32_t sb = sW + (uint32_t)st * kStage32;
        auto desc = [&](uint32_t off) { return umma::smem_desc(sb + off, kLboW, kSboW); };
        auto group = [&](uint32_t dcol, uint32_t a_off, uint32_t b_off, uint32_t b_lo, int n) {
          mma_3xtf32_k16(tm + dcol, desc(a_off), desc(a_off + kACopy), desc(b_off), desc(b_off + b_lo), kStep,
                         umma::instr_desc_tf32(128, n), acc);
        };
        group(0u, P_AH, P_B + 1 * kBCopy, kBLo, 128);      // cols   0..127 : A_H  x [GA_3|GA_1|GA_2|GA_4]
        group(128u, P_AEC, P_B, kBLo, 64);                 // cols 128..191 : A_EC x [GA_0|GA_3]
        group(192u, P_AEC, P_B + 5 * kBCopy, kBLo, 160);   // cols 192..351 : A_EC x [GH_0..GH_4]
        group(352u, P_AG, P_B2, kB2Copy, 16);              // cols 352..367 : A_G  x [go|p]
        umma::mma_commit(&empty[st]);
      }
      umma::mma_commit(done);
    }
  } else {
    // ------------------------------------------------------------------ producers
    const float4* H4 = reinterpret_cast<const float4*>(a.H);
    const float4* E4 = reinterpret_cast<const float4*>(a.E);
    const float4* C4 = reinterpret_cast<const float4*>(a.C);
    const float4* GH4 = reinterpret_cast<const float4*>(a.GH);
    const float4* GR4 = reinterpret_cast<const float4*>(a.GARG);
    const float4* GO4 = reinterpret_cast<const float4*>(a.GO);
    unsigned char* base = smraw + grp * kStage32;
    auto pf = [](const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); };
    auto prefetch_chunk = [&](int64_t c) {
      const int64_t n = c * kChunk + s;
      if (c >= nchunks || n >= N) return;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int slot = slot0 + 16 * h;
        pf(H4 + (int64_t)slot * N + n);
        pf(slot < 8 ? H4 + (int64_t)(32 + slot) * N + n : E4 + (int64_t)(slot - 8) * N + n);
        pf(slot < 8 ? C4 + (int64_t)slot * N + n : GR4 + (int64_t)(slot - 8) * N + n);
        pf(GH4 + (int64_t)slot * N + n);
        if (slot < 8) pf(GH4 + (int64_t)(32 + slot) * N + n);
      }
    };
    auto masked = [](float4 v, uint32_t m) {
      return make_float4((m & 1u) ? v.x : 0.f, (m & 2u) ? v.y : 0.f, (m & 4u) ? v.z : 0.f, (m & 8u) ? v.w : 0.f);
    };
    auto acc4 = [](float4& t, const float4& v) { t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w; };
    // position of GA_l / GH_l among the B operands (GA_0, GA_3, GA_1, GA_2, GA_4, GH_0..GH_4)
    auto pos_ga = [](int l) { return l == 1 ? 2 : l == 2 ? 3 : l == 3 ? 1 : l; };
    constexpr int kAhead = 1;     // L2 prefetch distance in this group's chunks (measured: 1: 158 us, 2: 160 us, 4: 168 us)
    const int64_t stride = 2 * (int64_t)gridDim.x;
    const int64_t c0 = blockIdx.x + (int64_t)grp * gridDim.x;
    for (int d = 1; d < kAhead; ++d) prefetch_chunk(c0 + d * stride);
    uint32_t ph = 0u;
    int it = 0;
    for (int64_t c = c0; c < nchunks; c += stride, ++it) {
      const int64_t n = c * kChunk + s;
      const bool ok = n < N;
      prefetch_chunk(c + kAhead * stride);
      // global loads of both slots first (they do not touch shared memory), then wait for the stage to be free
      float4 v0[2], v1[2], v2[2], g0[2], g1 = z4;
      uint32_t m0[2], m1 = 0u;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int slot = slot0 + 16 * h;
        v0[h] = ok ? H4[(int64_t)slot * N + n] : z4;
        v1[h] = ok ? (slot < 8 ? H4[(int64_t)(32 + slot) * N + n] : E4[(int64_t)(slot - 8) * N + n]) : z4;
        v2[h] = ok ? (slot < 8 ? C4[(int64_t)slot * N + n] : GR4[(int64_t)(slot - 8) * N + n]) : z4;
        g0[h] = ok ? GH4[(int64_t)slot * N + n] : z4;
        m0[h] = ok ? (a.bits[(int64_t)(slot >> 3) * N + n] >> (4 * (slot & 7))) & 0xFu : 0u;
      }
      if (ok && slot0 < 8) { g1 = GH4[(int64_t)(32 + slot0) * N + n]; m1 = (a.bits[(int64_t)4 * N + n] >> (4 * slot0)) & 0xFu; }
      else if (ok && slot0 == 8) g1 = GO4[n];
      else if (ok && slot0 == 9) g1 = make_float4(a.P32[n], a.P32[N + n], a.P32[2 * N + n], 0.f);
      if (it >= 1) { umma::mbar_wait(&empty[grp], ph); ph ^= 1u; umma::tc_fence_after(); }
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int slot = slot0 + 16 * h, l3 = slot >> 3, q3 = slot & 7;
        put_split(base + P_AH, kACopy, slot, s, v0[h]);
        if (slot < 8) {
          put_split(base + P_AG, kACopy, 24 + slot, s, v1[h]);     // h4
          put_split(base + P_AEC, kACopy, 24 + slot, s, v2[h]);    // c
        } else {
          put_split(base + P_AEC, kACopy, slot - 8, s, v1[h]);     // emb
          put_split(base + P_AG, kACopy, slot - 8, s, v2[h]);      // garg
        }
        const float4 ga = masked(g0[h], m0[h]);
        put_split(base + P_B + (uint32_t)(5 + l3) * kBCopy, kBLo, q3, s, g0[h]);
        put_split(base + P_B + (uint32_t)pos_ga(l3) * kBCopy, kBLo, q3, s, ga);
        acc4(s_gh[h], g0[h]); acc4(s_ga[h], ga);
      }
      if (slot0 < 8) {
        const float4 ga = masked(g1, m1);
        put_split(base + P_B + 9u * kBCopy, kBLo, slot0, s, g1);
        put_split(base + P_B + 4u * kBCopy, kBLo, slot0, s, ga);
        acc4(s_gh4, g1); acc4(s_ga4, ga);
      } else if (slot0 == 8) {
        put_split(base + P_B2, kB2Copy, 0, s, g1);            // go -> rows 0..3
        acc4(s_go, g1);
      } else if (slot0 == 9) {
        put_split(base + P_B2, kB2Copy, 1, s, g1);            // p  -> rows 4..6
      }
      umma::fence_proxy_async();
      mbar_arrive(&full[grp]);
    }
  }
  // every MMA of this CTA has completed when `done` flips
  umma::mbar_wait(done, 0u);
  const int it = 1;
  umma::tc_fence_after();
  if (it > 0) {
    // ---- flush: thread = feature row (TMEM lane); warp w takes accumulators (w / 4) + 4 * i
    const int f = (warp & 3) * 32 + lane;
    const uint32_t tl = tm + ((uint32_t)((warp & 3) * 32) << 16);
    for (int ai = warp >> 2; ai < 12 && warp < 16; ai += 4) {
      if (ai == 11) {
        float v[16];
        uint32_t r[16];
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
              "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
            : "r"(tl + 352u)
            : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
        if (f < PN_EMBED) {
          if (a.B) { atomicAdd(a.B + f, v[4]); atomicAdd(a.B + PN_EMBED + f, v[5]); atomicAdd(a.B + 2 * PN_EMBED + f, v[6]); }
        } else if (f >= 96 && a.Wo) {
#pragma unroll
          for (int o = 0; o < 4; ++o)
            if (o < a.nout) atomicAdd(a.Wo + o * 32 + (f - 96), v[o]);
        }
        continue;
      }
      float v[32];
      umma::tmem_ld32(tl + 32u * ai, v);
      float* dst = nullptr;   // dW[j*ld + col]
      int ld = 0, col = -1;
      if (ai < 4) {           // A_H x GA_l with l = 3,1,2,4: useful rows are h_{l-1} = features 32*(l-1) ..
        const int l = ai == 0 ? 3 : ai == 3 ? 4 : ai;
        if ((f >> 5) == l - 1) { dst = a.W[l]; ld = l == 3 ? PN_EMBED + 32 : 32; col = (l == 3 ? PN_EMBED : 0) + (f & 31); }
      } else if (ai == 4) { if (f < PN_EMBED) { dst = a.W[0]; ld = PN_EMBED; col = f; } }
      else if (ai == 5) { if (f < PN_EMBED) { dst = a.W[3]; ld = PN_EMBED + 32; col = f; } }
      else { if (f >= 96) { dst = a.Wc[ai - 6]; ld = 32; col = f - 96; } }
      if (dst) {
#pragma unroll
        for (int j = 0; j < 32; ++j) atomicAdd(dst + (int64_t)j * ld + col, v[j]);
      }
    }
    // ---- biases: reduce each partial sum over the 16 sample lanes of the half-warp
    auto reduce16 = [&](float4 t) {
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) {
        t.x += __shfl_xor_sync(kFull, t.x, o); t.y += __shfl_xor_sync(kFull, t.y, o);
        t.z += __shfl_xor_sync(kFull, t.z, o); t.w += __shfl_xor_sync(kFull, t.w, o);
      }
      return t;
    };
    auto add4 = [&](float* dst, int q, const float4& t) {
      if (dst && s == 0) { atomicAdd(dst + 4 * q, t.x); atomicAdd(dst + 4 * q + 1, t.y); atomicAdd(dst + 4 * q + 2, t.z); atomicAdd(dst + 4 * q + 3, t.w); }
    };
    if (warp < 16) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int slot = slot0 + 16 * h;
        add4(a.b[slot >> 3], slot & 7, reduce16(s_ga[h]));
        add4(a.bc[slot >> 3], slot & 7, reduce16(s_gh[h]));
      }
      const float4 ta4 = reduce16(s_ga4), th4 = reduce16(s_gh4), tgo = reduce16(s_go);
      if (slot0 < 8) { add4(a.b[4], slot0, ta4); add4(a.bc[4], slot0, th4); }
      if (slot0 == 8 && s == 0 && a.bo) {
        const float t[4] = {tgo.x, tgo.y, tgo.z, tgo.w};
        for (int o = 0; o < a.nout; ++o) atomicAdd(a.bo + o, t[o]);
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


// c_dim 32: every parameter gradient (W, b, Wc, bc, Wo, bo, B; null sinks are skipped) in one kernel.
// GH is the only gradient stash it reads; the ReLU masks come from the forward's relu_bits.
int launch_wgrad_tc32(int64_t N, int n_out, const float* H, const float* C, const float* E, const float* GH, const float* GARG,
                      const float* GO, const float* P32, const uint32_t* relu_bits, float* const* W, float* const* b,
                      float* const* Wc, float* const* bc, float* Wo, float* bo, float* B, cudaStream_t st) {
  WgTc32Args a;
  a.H = H; a.C = C; a.E = E; a.GH = GH; a.GARG = GARG; a.GO = GO; a.P32 = P32; a.bits = relu_bits; a.N = N; a.nout = n_out;
  for (int l = 0; l < 5; ++l) { a.W[l] = W[l]; a.b[l] = b[l]; a.Wc[l] = Wc[l]; a.bc[l] = bc[l]; }
  a.Wo = Wo; a.bo = bo; a.B = B;
  cudaFuncSetAttribute(k_wgrad_tc32, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem32);
  const int64_t nchunks = (N + kChunk - 1) / kChunk;
  const int grid = (int)(nchunks < (int64_t)sm_count() ? nchunks : (int64_t)sm_count());
  launch_pdl(k_wgrad_tc32, dim3(grid), dim3(kWg32Threads), (size_t)kSmem32, st, a);
  return launch_status("k_wgrad_tc32");
}

}  // namespace pn
