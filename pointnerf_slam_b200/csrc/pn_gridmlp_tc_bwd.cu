// Backward (input gradients) of a NICE grid decoder on the tensor cores.
#include "pn_gridmlp.cuh"

namespace pn {
namespace {

// ---------------------------------------------------------------------------
// backward (input gradients) on the tensor cores
// ---------------------------------------------------------------------------
// Same tiling as the forward: CTA = 3 groups x 256 threads, two threads per sample row.
// Every transposed mat-vec of the FFMA kernel becomes D[128 x N] (+)= G[128 x 32] . (W^T)[N x 32]^T
// with the gradient operand G (hi/lo) in the group's shared buffer and the transposed weights
// pre-split in shared memory.  Tensor-memory columns per group:
//     0..31   D_gc = feature gradient, accumulated over blocks
//    32..63   D_x  = ga_l . W_l                     (gradient gh_{l-1} at the previous block's output)
//    64..159  D_ge = ga_3 . W3[:, :93] + ga_0 . W0  (gradient at the Fourier embedding)
// One tensor-core round trip per block: the feature gradient sum_l gh_l . Wc_l is rewritten with
// gh_{l-1} = ga_l . W_l as  go . (Wo Wc_4) + sum_{l>=1} ga_l . (W_l Wc_{l-1}),  so the SAME operand ga_l
// feeds both products of a block (the weight products M_{l-1} = W_l Wc_{l-1} are formed once per
// CTA while staging; the go term is four FMAs per column in registers).
namespace tcb {
using tc::kLbo; using tc::kASbo; using tc::kABytes;
constexpr int kGroups = 3;                     // independent 128-sample tiles in flight per CTA (160 TMEM columns each)
constexpr uint32_t kBB = 4096;                 // one [32 x 32] operand copy
constexpr uint32_t kBE = 12288;                // one [96 x 32] operand copy
constexpr uint32_t O_WT = 0;                   // W1^T, W2^T, W3h^T, W4^T (hi, lo each)
constexpr uint32_t O_WCT = O_WT + 8 * kBB;     // M_m^T = (W_{m+1} Wc_m[:, :32])^T, m = 0..3
constexpr uint32_t O_W0T = O_WCT + 8 * kBB;    // W0^T  [96 x 32]
constexpr uint32_t O_W3ET = O_W0T + 2 * kBE;   // W3[:, :93]^T
constexpr uint32_t O_A = O_W3ET + 2 * kBE;     // 2 groups x (hi, lo)
constexpr uint32_t O_SMALL = O_A + kGroups * 2 * kABytes;
constexpr int S_B = 0, S_WO = 288, S_WOC = 416, S_TOTAL = 544;  // floats
constexpr uint32_t kSmem = O_SMALL + S_TOTAL * 4u + 40u;
constexpr int kTileLd = 36;                    // padded row of the feature-gradient tile (reuses the A buffer; 16-byte aligned rows)

// B operand = transpose of a row-major [32 x ld] weight block: element (n, j) = src[j*ld + col0 + n], n < nrows
__device__ __forceinline__ void stage_bt(unsigned char* hi, uint32_t copy_bytes, const float* __restrict__ src, int ld, int col0,
                                         int nrows, int nvalid) {
  unsigned char* lo = hi + copy_bytes;
  for (int i = threadIdx.x; i < nrows * 32; i += blockDim.x) {
    const int j = i / nrows, n = i - j * nrows;   // n fastest: coalesced over the source row
    const float w = n < nvalid ? src[j * ld + col0 + n] : 0.f;
    float h, l;
    umma::split_tf32(w, h, l);
    const uint32_t off = umma::kmajor_off(n, j, kLbo, 1024u);
    *reinterpret_cast<float*>(hi + off) = h;
    *reinterpret_cast<float*>(lo + off) = l;
  }
}

// B operand = transpose of the weight product M = Wl[:, colw .. colw+31] . Wc[:, :32]   ([32 x 32], FP32 sums)
__device__ __forceinline__ void stage_prod_t(unsigned char* hi, uint32_t copy_bytes, const float* __restrict__ Wl, int ldw, int colw,
                                             const float* __restrict__ Wc, int cd) {
  unsigned char* lo = hi + copy_bytes;
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) {
    const int k = i >> 5, n = i & 31;
    float m = 0.f;
    for (int t = 0; t < 32; ++t) m = fmaf(Wl[k * ldw + colw + t], Wc[t * cd + n], m);
    float h, l;
    umma::split_tf32(m, h, l);
    const uint32_t off = umma::kmajor_off(n, k, kLbo, 1024u);
    *reinterpret_cast<float*>(hi + off) = h;
    *reinterpret_cast<float*>(lo + off) = l;
  }
}
}  // namespace tcb

template <int CD, int NOUT, bool GRID_GRAD, bool NEED_DP, bool WS>
__global__ void __launch_bounds__(tcb::kGroups * 256, 1) k_grid_mlp_bwd_tc(const BwdArgs a) {
  extern __shared__ __align__(128) unsigned char smraw[];
  using namespace tcb;
  constexpr bool EMB = NEED_DP || WS;
  const int tid = threadIdx.x, lane = tid & 31;
  const int grp = tid >> 8, gw = (tid >> 5) & 7, quarter = gw & 3, half = gw >> 2;
  const int row = quarter * 32 + lane, col0 = 16 * half;
  float* sm = reinterpret_cast<float*>(smraw + O_SMALL);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + S_TOTAL);
  uint32_t& tmem_base_s = *reinterpret_cast<uint32_t*>(sm + S_TOTAL + 8);
  unsigned char* a_hi = smraw + O_A + (uint32_t)grp * 2u * kABytes;
  unsigned char* a_lo = a_hi + kABytes;
  float* gtile = reinterpret_cast<float*>(a_hi);   // [128][36] floats, valid between the last MMA and the next tile
  pdl_launch_dependents();   // prologue below is independent of the previous kernel
  if (tid < 32) umma::tmem_alloc(&tmem_base_s, 512);
  if (tid == 0) { for (int i = 0; i < kGroups; ++i) umma::mbar_init(&bars[i], 1); umma::fence_mbar_init(); }
  stage_bt(smraw + O_WT, kBB, a.w.W[1], 32, 0, 32, 32);
  stage_bt(smraw + O_WT + 2 * kBB, kBB, a.w.W[2], 32, 0, 32, 32);
  stage_bt(smraw + O_WT + 4 * kBB, kBB, a.w.W[3], PN_EMBED + 32, PN_EMBED, 32, 32);
  stage_bt(smraw + O_WT + 6 * kBB, kBB, a.w.W[4], 32, 0, 32, 32);
  constexpr bool GC = GRID_GRAD || NEED_DP;
  if (GC) {
    stage_prod_t(smraw + O_WCT, kBB, a.w.W[1], 32, 0, a.w.Wc[0], CD);
    stage_prod_t(smraw + O_WCT + 2 * kBB, kBB, a.w.W[2], 32, 0, a.w.Wc[1], CD);
    stage_prod_t(smraw + O_WCT + 4 * kBB, kBB, a.w.W[3], PN_EMBED + 32, PN_EMBED, a.w.Wc[2], CD);
    stage_prod_t(smraw + O_WCT + 6 * kBB, kBB, a.w.W[4], 32, 0, a.w.Wc[3], CD);
    for (int i = tid; i < 128; i += blockDim.x) {   // (Wo Wc_4)[o][j]
      const int o = i >> 5, j = i & 31;
      float m = 0.f;
      if (o < NOUT)
        for (int t = 0; t < 32; ++t) m = fmaf(a.w.Wo[o * 32 + t], a.w.Wc[4][t * CD + j], m);
      sm[S_WOC + i] = m;
    }
  }
  if (EMB) {
    stage_bt(smraw + O_W0T, kBE, a.w.W[0], PN_EMBED, 0, 96, PN_EMBED);
    stage_bt(smraw + O_W3ET, kBE, a.w.W[3], PN_EMBED + 32, 0, 96, PN_EMBED);
  }
  for (int i = tid; i < 288; i += blockDim.x) { const int d = i / 96, k = i % 96; sm[S_B + i] = k < PN_EMBED ? a.w.B[d * PN_EMBED + k] : 0.f; }
  for (int i = tid; i < 128; i += blockDim.x) sm[S_WO + i] = i < NOUT * 32 ? a.w.Wo[i] : 0.f;
  umma::fence_proxy_async();
  umma::tc_fence_before();
  __syncthreads();
  umma::tc_fence_after();
  pdl_wait();   // gradients, stash and masks written by earlier kernels are read from here on
  const uint32_t tm = tmem_base_s + (uint32_t)grp * 160u;
  const uint32_t tm_lane = tm + ((uint32_t)(quarter * 32) << 16);
  const uint32_t sA = umma::smem_u32(a_hi), sW = umma::smem_u32(smraw);
  constexpr uint32_t idesc32 = umma::instr_desc_tf32(128, 32), idesc96 = umma::instr_desc_tf32(128, 96);
  uint64_t* bar = &bars[grp];
  uint32_t phase = 0;
  const bool issuer = (tid & 255) == 0;
  const uint64_t dA_hi = umma::smem_desc(sA, kLbo, kASbo), dA_lo = umma::smem_desc(sA + kABytes, kLbo, kASbo);
  constexpr uint32_t kStep = (2u * kLbo) >> 4;
  auto mma = [&](uint32_t dcol, uint32_t boff, uint32_t copy_bytes, uint32_t idesc, uint32_t acc) {
    const uint64_t dB_hi = umma::smem_desc(sW + boff, kLbo, 1024u), dB_lo = umma::smem_desc(sW + boff + copy_bytes, kLbo, 1024u);
    umma::mma_3xtf32_k32(tm + dcol, dA_hi, dA_lo, dB_hi, dB_lo, kStep, kStep, idesc, acc);
  };
  auto group_bar = [&] { asm volatile("bar.sync %0, 256;" ::"r"(grp + 1) : "memory"); };
  auto publish_issue = [&](auto&& issue) {
    umma::fence_proxy_async();
    umma::tc_fence_before();
    group_bar();
    if (issuer) { umma::tc_fence_after(); issue(); umma::mma_commit(bar); }
  };
  auto wait_mma = [&] {   // one lane polls the mbarrier; the rest of the group sleeps on a named barrier
    if ((tid & 255) < 32) { if (lane == 0) umma::mbar_wait(bar, phase); __syncwarp(); }
    asm volatile("bar.sync %0, 256;" ::"r"(grp + 4) : "memory");
    phase ^= 1u;
    umma::tc_fence_after();
  };
  auto store_half_row = [&](const float (&v)[16]) {
    const uint32_t base = (uint32_t)(row >> 3) * kASbo + (uint32_t)(row & 7) * 16u + (uint32_t)(4 * half) * kLbo;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float4 h, l;
      umma::split_tf32(v[4 * q], h.x, l.x); umma::split_tf32(v[4 * q + 1], h.y, l.y);
      umma::split_tf32(v[4 * q + 2], h.z, l.z); umma::split_tf32(v[4 * q + 3], h.w, l.w);
      *reinterpret_cast<float4*>(a_hi + base + q * kLbo) = h;
      *reinterpret_cast<float4*>(a_lo + base + q * kLbo) = l;
    }
  };
  auto stash_half = [&](float* base, int64_t N, int64_t n, const float (&v)[16]) {
    float4* o = reinterpret_cast<float4*>(base);
#pragma unroll
    for (int q = 0; q < 4; ++q) o[(int64_t)(4 * half + q) * N + n] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
  };

  const int64_t N = a.pts.N, ntiles = (N + 127) / 128;
  // tile order as in the forward: the partial last wave is spread over the SMs instead of filling whole CTAs
  for (int64_t t = (int64_t)blockIdx.x + (int64_t)gridDim.x * grp; t < ntiles; t += (int64_t)gridDim.x * kGroups) {
    const int64_t n = t * 128 + row;
    const bool valid = n < N;
    Sample sp;
    sp.pf[0] = sp.pf[1] = sp.pf[2] = 0.f; sp.xn[0] = sp.xn[1] = sp.xn[2] = 0.f; sp.inside = true;
    if (valid) load_sample(a.pts, n, a.nb, a.mb, sp);
    const unsigned vm = __ballot_sync(kFull, valid);
    float go[4] = {0.f, 0.f, 0.f, 0.f};
    if (valid) {
      const float4 g = reinterpret_cast<const float4*>(a.g_raw)[n];
      if (NOUT == 4) { go[0] = g.x; go[1] = g.y; go[2] = g.z; }
      else go[0] = (a.apply_mask && !sp.inside) ? 0.f : g.w;
    }
    if (WS && valid && half == 0) {
      reinterpret_cast<float4*>(a.GO)[n] = make_float4(go[0], go[1], go[2], go[3]);
      a.P32[n] = sp.pf[0]; a.P32[N + n] = sp.pf[1]; a.P32[2 * N + n] = sp.pf[2];
    }
    float gh[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      float s = 0.f;
#pragma unroll
      for (int o = 0; o < NOUT; ++o) s = fmaf(sm[S_WO + o * 32 + col0 + j], go[o], s);
      gh[j] = s;
    }
    // the previous tile's scatter phase used the A buffer as a scratch tile: all of the group must be done with it
    group_bar();
    // ReLU masks of the five blocks: five independent loads in flight now, consumed block by block
    uint32_t b01 = 0u, b23 = 0u, b4 = 0u;
    if (valid) {
      const uint16_t* rb = reinterpret_cast<const uint16_t*>(a.relu_bits) + 2 * n + half;
      const uint32_t m0 = rb[0], m1 = rb[2 * N], m2 = rb[4 * N], m3 = rb[6 * N];
      b4 = rb[8 * N];
      b01 = m0 | (m1 << 16); b23 = m2 | (m3 << 16);
    }
#pragma unroll 1
    for (int l = 4; l >= 0; --l) {
      if (WS && valid) stash_half(a.GH + (int64_t)l * 32 * N, N, n, gh);
      const uint32_t bits = l == 4 ? b4 : (l >= 2 ? b23 >> (16 * (l - 2)) : b01 >> (16 * l));
#pragma unroll
      for (int j = 0; j < 16; ++j) gh[j] = ((bits >> j) & 1u) ? gh[j] : 0.f;    // ga_l
      if (l > 0 || EMB) {
        store_half_row(gh);
        publish_issue([&] {
          if (l == 4) mma(32u, O_WT + 6 * kBB, kBB, idesc32, 0u);
          else if (l == 3) { mma(32u, O_WT + 4 * kBB, kBB, idesc32, 0u); if (EMB) mma(64u, O_W3ET, kBE, idesc96, 0u); }
          else if (l == 2) mma(32u, O_WT + 2 * kBB, kBB, idesc32, 0u);
          else if (l == 1) mma(32u, O_WT, kBB, idesc32, 0u);
          else mma(64u, O_W0T, kBE, idesc96, 1u);
          if (GC && l > 0) mma(0u, O_WCT + (uint32_t)(l - 1) * 2u * kBB, kBB, idesc32, l < 4 ? 1u : 0u);   // D_gc (+)= ga_l . M_{l-1}
        });
      }
      // c_dim 32: the weight-gradient kernel rebuilds ga_l from gh_l and the ReLU bits, so only GH is stashed
      if (WS && CD == 64 && valid) stash_half(a.GA + (int64_t)l * 32 * N, N, n, gh);   // under the products
      if (l > 0 || EMB) {
        wait_mma();
        if (l > 0) tmem_ld16(tm_lane + 32u + col0, gh);
      }
    }
    // ---- Fourier embedding: gradient at the arguments, point gradient, dB scratch
    float gp[3] = {0.f, 0.f, 0.f};
    if (EMB) {
#pragma unroll 1
      for (int c = 0; c < 3; ++c) {
        float ge[16];
        tmem_ld16(tm_lane + 64u + 32u * c + col0, ge);
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          const int kk = 32 * c + col0 + k;
          const float bx = sm[S_B + kk], by = sm[S_B + 96 + kk], bz = sm[S_B + 192 + kk];
          const float garg = ge[k] * fourier_cos(fmaf(sp.pf[2], bz, fmaf(sp.pf[1], by, sp.pf[0] * bx)));
          ge[k] = garg;
          gp[0] = fmaf(bx, garg, gp[0]); gp[1] = fmaf(by, garg, gp[1]); gp[2] = fmaf(bz, garg, gp[2]);
        }
        if (WS && valid) {
          float4* o = reinterpret_cast<float4*>(a.GARG);
#pragma unroll
          for (int q = 0; q < 4; ++q)
            o[(int64_t)(8 * c + 4 * half + q) * N + n] = make_float4(ge[4 * q], ge[4 * q + 1], ge[4 * q + 2], ge[4 * q + 3]);
        }
      }
    }
    // ---- feature gradient: TMEM -> scratch tile [row][channel] -> warp-cooperative scatter
    if (GRID_GRAD || NEED_DP) {
      float gc[16];
      tmem_ld16(tm_lane + col0, gc);
#pragma unroll
      for (int j = 0; j < 16; ++j) {
#pragma unroll
        for (int o = 0; o < NOUT; ++o) gc[j] = fmaf(sm[S_WOC + o * 32 + col0 + j], go[o], gc[j]);   // gh_4 . Wc_4
      }
#pragma unroll
      for (int q = 0; q < 4; ++q)
        *reinterpret_cast<float4*>(gtile + row * kTileLd + col0 + 4 * q) = make_float4(gc[4 * q], gc[4 * q + 1], gc[4 * q + 2], gc[4 * q + 3]);
      umma::tc_fence_before();
      group_bar();
      // warp (quarter, half) scatters rows 32*quarter + 16*half + [0,16): 8 lanes per sample (lane&7 = channel
      // quad, 128-bit vector reductions), 4 samples per iteration
      const GridDev& g = a.ga;
      const float ux = unnormalise(sp.xn[0], g.W), uy = unnormalise(sp.xn[1], g.H), uz = unnormalise(sp.xn[2], g.D);
      const int q = lane & 7, sub = lane >> 3;
      float dux = 0.f, duy = 0.f, duz = 0.f;
#pragma unroll 1
      for (int it = 0; it < 4; ++it) {
        const int src = 16 * half + 4 * it + sub;
        const float sx = __shfl_sync(kFull, ux, src), sy = __shfl_sync(kFull, uy, src), sz = __shfl_sync(kFull, uz, src);
        float gx = 0.f, gy = 0.f, gz = 0.f;
        if ((vm >> src) & 1u) {
          const Cell c = make_cell(sx, sy, sz, g.W, g.H, g.D);
          const float4 gv = *reinterpret_cast<const float4*>(gtile + (quarter * 32 + src) * kTileLd + 4 * q);
          if (GRID_GRAD) {
            float* gg = a.g_grid + c.base + 4 * q;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              if ((c.ok >> k) & 1u) {
                const float w = corner_weight(c, k);
                red_add_v4(gg + corner_offset(k, g.W, g.H), w * gv.x, w * gv.y, w * gv.z, w * gv.w);
              }
            }
          }
          if (NEED_DP) {
            const float4* gd = reinterpret_cast<const float4*>(g.data + c.base) + q;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              if ((c.ok >> k) & 1u) {
                const float4 f = __ldg(gd + corner_offset(k, g.W, g.H) / 4);
                const float v = fmaf(f.w, gv.w, fmaf(f.z, gv.z, fmaf(f.y, gv.y, f.x * gv.x)));
                const float wx = c.wx[k & 1], wy = c.wy[(k >> 1) & 1], wz = c.wz[k >> 2];
                gx += ((k & 1) ? v : -v) * wy * wz;
                gy += (((k >> 1) & 1) ? v : -v) * wx * wz;
                gz += ((k >> 2) ? v : -v) * wx * wy;
              }
            }
            gx *= c.gm[0]; gy *= c.gm[1]; gz *= c.gm[2];
          }
        }
        if (NEED_DP) {
          // sum over the sample's 8 lanes, then hand the result to the lane that owns the row
#pragma unroll
          for (int o = 4; o > 0; o >>= 1) {
            gx += __shfl_xor_sync(kFull, gx, o); gy += __shfl_xor_sync(kFull, gy, o); gz += __shfl_xor_sync(kFull, gz, o);
          }
          const int from = 8 * (lane & 3);
          const float rx = __shfl_sync(kFull, gx, from), ry = __shfl_sync(kFull, gy, from), rz = __shfl_sync(kFull, gz, from);
          if (lane == 16 * half + 4 * it + (lane & 3)) { dux = rx; duy = ry; duz = rz; }
        }
      }
      if (NEED_DP) {
        // rows 16*half..16*half+15 of this quarter got their grid-path gradient in lanes 16*half + i of THIS warp;
        // combine with the embedding path: each (row, half) thread holds gp of its own 16 embedding columns
        const bool mine = (lane >> 4) == half;   // this lane's row was scattered by this warp
        if (mine && valid) {
          gp[0] += norm_grad(a.pts, a.nb, 0, dux); gp[1] += norm_grad(a.pts, a.nb, 1, duy); gp[2] += norm_grad(a.pts, a.nb, 2, duz);
        }
        if (valid) {
          // each (row, half) thread adds its partial sum; float atomics on 3 words per row, 2 adders per word
          float* o = a.g_pts + 3 * n;
          atomicAdd(o, gp[0]); atomicAdd(o + 1, gp[1]); atomicAdd(o + 2, gp[2]);
        }
      }
    } else if (NEED_DP && valid) {
      float* o = a.g_pts + 3 * n;
      atomicAdd(o, gp[0]); atomicAdd(o + 1, gp[1]); atomicAdd(o + 2, gp[2]);
    }
    umma::tc_fence_before();
  }
  umma::tc_fence_before();
  __syncthreads();
  if (tid < 32) umma::tmem_dealloc(tmem_base_s, 512);
}


template <int CD, int NOUT, bool GG, bool DP, bool WS>
int launch_t(const BwdArgs& a, cudaStream_t st) {
  auto kern = k_grid_mlp_bwd_tc<CD, NOUT, GG, DP, WS>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tcb::kSmem);
  const int64_t ntiles = (a.pts.N + 127) / 128;   // fewer tiles than SMs: one tile (group 0) per CTA
  const int grid = (int)((ntiles < (int64_t)sm_count()) ? ntiles : (int64_t)sm_count());
  launch_pdl(kern, dim3(grid), dim3(tcb::kGroups * 256), (size_t)tcb::kSmem, st, a);
  return launch_status("k_grid_mlp_bwd_tc");
}

template <int CD, int NOUT>
int launch_flags(const BwdArgs& a, bool gg, bool dp, bool ws, cudaStream_t st) {
  const int key = (gg ? 4 : 0) | (dp ? 2 : 0) | (ws ? 1 : 0);
  switch (key) {
    case 0: return launch_t<CD, NOUT, false, false, false>(a, st);
    case 1: return launch_t<CD, NOUT, false, false, true>(a, st);
    case 2: return launch_t<CD, NOUT, false, true, false>(a, st);
    case 3: return launch_t<CD, NOUT, false, true, true>(a, st);
    case 4: return launch_t<CD, NOUT, true, false, false>(a, st);
    case 5: return launch_t<CD, NOUT, true, false, true>(a, st);
    case 6: return launch_t<CD, NOUT, true, true, false>(a, st);
    default: return launch_t<CD, NOUT, true, true, true>(a, st);
  }
}

}  // namespace

int launch_bwd_tc(int c_dim, int n_out, bool gg, bool dp, bool ws, const BwdArgs& a, cudaStream_t st) {
  if (c_dim == 32) return n_out == 4 ? launch_flags<32, 4>(a, gg, dp, ws, st) : launch_flags<32, 1>(a, gg, dp, ws, st);
  return n_out == 4 ? launch_flags<64, 4>(a, gg, dp, ws, st) : launch_flags<64, 1>(a, gg, dp, ws, st);
}

}  // namespace pn
