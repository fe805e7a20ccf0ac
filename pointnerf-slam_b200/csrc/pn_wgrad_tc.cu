// Weight gradients of a grid decoder on the tensor cores.
//
// dW[j][f] = sum_n G[n][j] * X[n][f] is a GEMM whose reduction runs over SAMPLES: the sample
// axis is the MMA K dimension, features (M = 128) and outputs (N = 32) are the rows of the two
// operands.  The operands are read from the planar-4 stash ([feature/4][sample] float4, coalesced
// 512-byte rows per 32-sample chunk), split into TF32 hi/lo in registers and stored TRANSPOSED
// into the canonical K-major no-swizzle layout
//     element (row r, sample k) at (r/8)*SBO + (k/4)*LBO + (r%8)*16 + (k%4)*4   (bytes)
// with LBO = 144 (a 128-byte core matrix + 16 bytes of padding, which makes the scalar transposing
// stores of a warp -- 32 consecutive samples of one row -- bank-conflict free) and SBO = 8*LBO.
// (The MN-major operand mode, which would take the stash rows unchanged, returns zeros for
// kind::tf32 on sm_100a -- measured with pn_tc_selftest_mn -- so it is not used.)
//
// Per 32-sample chunk (K = 32 = 4 MMA k-steps), M = 128 features, N = 32 outputs, 3xTF32:
//     A_H = [h0|h1|h2|h3]   x GA_1..GA_4  -> dW1, dW2, dW3[:, 93:], dW4 (diagonal 32-row blocks)
//     A_E = [emb(96)|0(32)] x GA_0, GA_3  -> dW0, dW3[:, :93]
//     A_C = [c(CD)|0]       x GH_0..GH_4  -> dWc_0..4
// 11 FP32 accumulators (352 tensor-memory columns) stay resident for the whole kernel; each CTA
// flushes them once with atomics.  Bias gradients (column sums of GA_l / GH_l) are accumulated
// by the loading threads on the side.
#include "pn_common.cuh"
#include "pn_umma.cuh"

namespace pn {
namespace {

constexpr int kChunk = 32;                     // samples per step
constexpr uint32_t kLboW = 144;                // padded core matrix
constexpr uint32_t kSboW = 8 * kLboW;          // 8 core matrices = 32 samples per 8-row group
constexpr uint32_t kACopy = 16 * kSboW;        // one copy of a 128-feature operand (18 KB)
constexpr uint32_t kBCopy = 4 * kSboW;         // one copy of a 32-output operand (4.5 KB)
constexpr uint32_t O_AH = 0, O_AE = O_AH + 2 * kACopy, O_AC = O_AE + 2 * kACopy, O_B = O_AC + 2 * kACopy;
constexpr uint32_t O_END = O_B + 10 * 2 * kBCopy;   // GA_0..4, GH_0..4
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

__global__ void __launch_bounds__(256, 1) k_wgrad_tc(const WgTcArgs a) {
  extern __shared__ __align__(128) unsigned char smraw[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smraw + O_END);
  uint32_t& tmem_base_s = *reinterpret_cast<uint32_t*>(smraw + O_END + 16);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t N = a.N;
  const int cq = a.cd / 4;                       // feature quads of c
  if (warp == 0) umma::tmem_alloc(&tmem_base_s, 512);
  if (tid == 0) { umma::mbar_init(bar, 1); umma::fence_mbar_init(); }
  // zero the padding rows once (A_E quads 24..31, A_C quads cq..31), hi and lo copies
  for (int i = tid; i < 2 * (int)kACopy / 16; i += 256) {
    reinterpret_cast<float4*>(smraw + O_AE)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    reinterpret_cast<float4*>(smraw + O_AC)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  umma::fence_proxy_async();
  umma::tc_fence_before();
  __syncthreads();
  umma::tc_fence_after();
  const uint32_t tm = tmem_base_s;
  const uint32_t sW = umma::smem_u32(smraw);
  constexpr uint32_t idesc = umma::instr_desc_tf32(128, 32);
  constexpr uint32_t kStep = (2u * kLboW) >> 4;  // one K-step = 8 samples = two core matrices
  auto desc = [&](uint32_t off) { return umma::smem_desc(sW + off, kLboW, kSboW); };
  uint32_t phase = 0;
  float4 bsum[10];
#pragma unroll
  for (int i = 0; i < 10; ++i) bsum[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
  const int64_t nchunks = (N + kChunk - 1) / kChunk;
  bool first = true;
  const int s = lane;                            // sample slot of this thread
  for (int64_t c = blockIdx.x; c < nchunks; c += gridDim.x) {
    const int64_t n = c * kChunk + s;
    const bool ok = n < N;
    // ---- operands: thread (warp = quad group, lane = sample).  H: 32 quads, E: 24, C: cq, G: 8 each
    const float4* H4 = reinterpret_cast<const float4*>(a.H);
    const float4* E4 = reinterpret_cast<const float4*>(a.E);
    const float4* C4 = reinterpret_cast<const float4*>(a.C);
    const float4* GA4 = reinterpret_cast<const float4*>(a.GA);
    const float4* GH4 = reinterpret_cast<const float4*>(a.GH);
#pragma unroll
    for (int i = 0; i < 4; ++i) {                // h_0..h_3: block i, quad = warp
      const float4 v = ok ? H4[((int64_t)i * 8 + warp) * N + n] : z4;
      put_split(smraw + O_AH, kACopy, i * 8 + warp, s, v);
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) {                // embedding quads i*8 + warp
      const float4 v = ok ? E4[((int64_t)i * 8 + warp) * N + n] : z4;
      put_split(smraw + O_AE, kACopy, i * 8 + warp, s, v);
    }
    for (int q = warp; q < cq; q += 8) {
      const float4 v = ok ? C4[(int64_t)q * N + n] : z4;
      put_split(smraw + O_AC, kACopy, q, s, v);
    }
#pragma unroll
    for (int l = 0; l < 5; ++l) {                // GA_l, GH_l: quad = warp
      const float4 ga = ok ? GA4[((int64_t)l * 8 + warp) * N + n] : z4;
      const float4 gh = ok ? GH4[((int64_t)l * 8 + warp) * N + n] : z4;
      put_split(smraw + O_B + (uint32_t)l * 2u * kBCopy, kBCopy, warp, s, ga);
      put_split(smraw + O_B + (uint32_t)(5 + l) * 2u * kBCopy, kBCopy, warp, s, gh);
      bsum[l].x += ga.x; bsum[l].y += ga.y; bsum[l].z += ga.z; bsum[l].w += ga.w;
      bsum[5 + l].x += gh.x; bsum[5 + l].y += gh.y; bsum[5 + l].z += gh.z; bsum[5 + l].w += gh.w;
    }
    umma::fence_proxy_async();
    umma::tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      umma::tc_fence_after();
      const uint32_t acc = first ? 0u : 1u;
      auto group = [&](uint32_t dcol, uint32_t a_off, uint32_t b_idx) {
        const uint32_t b_off = O_B + b_idx * 2u * kBCopy;
        umma::mma_3xtf32_k32(tm + dcol, desc(a_off), desc(a_off + kACopy), desc(b_off), desc(b_off + kBCopy), kStep, kStep, idesc, acc);
      };
      for (uint32_t l = 1; l <= 4; ++l) group(32u * (l - 1), O_AH, l);       // cols   0..127 : A_H x GA_l
      group(128u, O_AE, 0);                                                   // cols 128..159 : A_E x GA_0
      group(160u, O_AE, 3);                                                   // cols 160..191 : A_E x GA_3
      for (uint32_t l = 0; l < 5; ++l) group(192u + 32u * l, O_AC, 5 + l);    // cols 192..351 : A_C x GH_l
      umma::mma_commit(bar);
    }
    first = false;
    umma::mbar_wait(bar, phase);
    phase ^= 1u;
    umma::tc_fence_after();
  }
  if (!first) {
    // ---- flush: thread = feature row (TMEM lane); warps 0-3 take accumulators 0..5, warps 4-7 take 6..10
    const int f = (warp & 3) * 32 + lane;
    const uint32_t tl = tm + ((uint32_t)((warp & 3) * 32) << 16);
    const int a0 = warp < 4 ? 0 : 6, a1 = warp < 4 ? 6 : 11;
    for (int ai = a0; ai < a1; ++ai) {
      float v[32];
      umma::tmem_ld32(tl + 32u * ai, v);
      float* dst = nullptr;   // dW[j*ld + col]
      int ld = 0, col = -1;
      if (ai < 4) {           // A_H x GA_{ai+1}: useful rows are block (ai) = features 32*ai .. 32*ai+31
        const int l = ai + 1;
        if ((f >> 5) == ai) { dst = a.W[l]; ld = l == 3 ? PN_EMBED + 32 : 32; col = (l == 3 ? PN_EMBED : 0) + (f & 31); }
      } else if (ai == 4) { if (f < PN_EMBED) { dst = a.W[0]; ld = PN_EMBED; col = f; } }
      else if (ai == 5) { if (f < PN_EMBED) { dst = a.W[3]; ld = PN_EMBED + 32; col = f; } }
      else { if (f < a.cd) { dst = a.Wc[ai - 6]; ld = a.cd; col = f; } }
      if (dst) {
#pragma unroll
        for (int j = 0; j < 32; ++j) atomicAdd(dst + (int64_t)j * ld + col, v[j]);
      }
    }
    // ---- biases: warp = output quad, reduce over the 32 sample slots
#pragma unroll
    for (int i = 0; i < 10; ++i) {
      const float x = warp_sum(bsum[i].x), y = warp_sum(bsum[i].y), zz = warp_sum(bsum[i].z), w = warp_sum(bsum[i].w);
      float* dst = i < 5 ? a.b[i] : a.bc[i - 5];
      if (lane == 0 && dst) { atomicAdd(dst + 4 * warp, x); atomicAdd(dst + 4 * warp + 1, y); atomicAdd(dst + 4 * warp + 2, zz); atomicAdd(dst + 4 * warp + 3, w); }
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
  k_wgrad_tc<<<grid, 256, kSmem, st>>>(a);
  return launch_status("k_wgrad_tc");
}

}  // namespace pn
