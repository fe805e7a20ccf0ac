// Forward of a NICE grid decoder on the 5th-gen tensor cores (tcgen05, kind::tf32, 3xTF32 split).
// Layout constants and the design notes are in pn_gridmlp.cuh (namespace tc).
#include "pn_gridmlp.cuh"

namespace pn {
namespace {

// CTA = 512 threads = two groups of 256.  Inside a group TWO threads serve each of the 128
// sample rows: warp (quarter q = warp&3, half = warp>>2) owns rows 32q..32q+31 (the TMEM lanes
// a warp with that id may address) and columns [16*half, 16*half+16) of every 32-wide operand.
// BITS: record the ReLU masks (needed by any backward); inference launches compile the bookkeeping away.
template <int CD, int NOUT, bool BITS>
__global__ void __launch_bounds__(512, 1) k_grid_mlp_fwd_tc(const FwdArgs a) {
  extern __shared__ __align__(128) unsigned char smraw[];
  using namespace tc;
  // A operand: core matrices adjacent in K are kALbo bytes apart.  144 instead of 128 puts the eight 16-byte stores of a
  // quarter-warp in put_rows (one sample row, channel quads 0..7) into eight different bank groups; with 128 they were
  // 8-way bank conflicts = 21 % (c_dim 32) to 44 % (c_dim 64) of the kernel's shared-memory wavefronts, and the
  // shared-memory data pipe (operand reads of the tensor core + these stores) is what bounds the forward kernels.
  // c_dim 64 has no room for the padding (8 KB): it keeps 128.
  constexpr uint32_t kALbo = a_lbo<CD>(), kASbo = 8u * kALbo, kABytes = 16u * kASbo;
  const int tid = threadIdx.x, lane = tid & 31;
  const int grp = tid >> 8, gw = (tid >> 5) & 7, quarter = gw & 3, half = gw >> 2;
  const int row = quarter * 32 + lane, col0 = 16 * half;
  float* sm = reinterpret_cast<float*>(smraw + o_small<CD>());
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + S_TOTAL);
  uint32_t& tmem_base_s = *reinterpret_cast<uint32_t*>(sm + S_TOTAL + 4);
  unsigned char* a_hi = smraw + o_a<CD>() + (uint32_t)grp * 2u * kABytes;
  unsigned char* a_lo = a_hi + kABytes;
  // ---- one-time set-up: TMEM, barriers, weights (independent of the previous kernel: see pdl_wait below)
  pdl_launch_dependents();
  if (tid < 32) umma::tmem_alloc(&tmem_base_s, 512);
  if (tid == 0) { umma::mbar_init(&bars[0], 1); umma::mbar_init(&bars[1], 1); umma::fence_mbar_init(); }
  constexpr uint32_t O_WC = o_wc<CD>();
  stage_b(smraw + O_WE, smraw + O_WE + 2 * bbytes(96), a.w.W[0], PN_EMBED, 0, 96, PN_EMBED);
  stage_b(smraw + O_WE + bbytes(96), smraw + O_WE + 3 * bbytes(96), a.w.W[3], PN_EMBED + 32, 0, 96, PN_EMBED);
  stage_b(smraw + O_WH, smraw + O_WH + bbytes(32), a.w.W[1], 32, 0, 32, 32);
  stage_b(smraw + O_WH + 2 * bbytes(32), smraw + O_WH + 3 * bbytes(32), a.w.W[2], 32, 0, 32, 32);
  stage_b(smraw + O_WH + 4 * bbytes(32), smraw + O_WH + 5 * bbytes(32), a.w.W[3], PN_EMBED + 32, PN_EMBED, 32, 32);
  stage_b(smraw + O_WH + 6 * bbytes(32), smraw + O_WH + 7 * bbytes(32), a.w.W[4], 32, 0, 32, 32);
  for (int l = 0; l < 5; ++l)
    stage_b(smraw + O_WC + (uint32_t)l * bbytes(CD), smraw + O_WC + (uint32_t)(5 + l) * bbytes(CD), a.w.Wc[l], CD, 0, CD, CD);
  for (int i = tid; i < 288; i += 512) { const int d = i / 96, k = i % 96; sm[S_B + i] = k < PN_EMBED ? a.w.B[d * PN_EMBED + k] : 0.f; }
  for (int i = tid; i < 160; i += 512) { sm[S_BIAS + i] = a.w.b[i >> 5][i & 31]; sm[S_BC + i] = a.w.bc[i >> 5][i & 31]; }
  for (int i = tid; i < 128; i += 512) sm[S_WO + i] = i < NOUT * 32 ? a.w.Wo[i] : 0.f;
  if (tid < 4) sm[S_BO + tid] = tid < NOUT ? a.w.bo[tid] : 0.f;
  umma::fence_proxy_async();
  umma::tc_fence_before();
  __syncthreads();
  umma::tc_fence_after();
  pdl_wait();   // everything below may read what the previous kernel of the stream wrote (raw, stash, grids)
  const uint32_t tm = tmem_base_s + (uint32_t)grp * 256u;               // this group's columns
  const uint32_t tm_lane = tm + ((uint32_t)(quarter * 32) << 16);        // this warp's lanes
  const uint32_t sA = umma::smem_u32(a_hi), sAlo = sA + kABytes;
  const uint32_t sW = umma::smem_u32(smraw);
  constexpr uint32_t idesc32 = umma::instr_desc_tf32(128, 32), idesc64 = umma::instr_desc_tf32(128, 64),
                     idesc160 = umma::instr_desc_tf32(128, 160);
  uint64_t* bar = &bars[grp];
  uint32_t phase = 0;
  const bool issuer = (tid & 255) == 0;
  const uint64_t dA_hi = umma::smem_desc(sA, kALbo, kASbo), dA_lo = umma::smem_desc(sAlo, kALbo, kASbo);
  constexpr uint32_t kStep = (2u * kLbo) >> 4, kAStep = (2u * kALbo) >> 4;   // one K-step of 8 in descriptor address units
  // D[:, dcol .. dcol+N) (+)= A . B[:, 32*k32 .. 32*k32+31]^T for the [N x Kb] operand whose hi copy
  // starts at byte offset boff and whose lo copy follows lo_off bytes later
  auto mma = [&](uint32_t dcol, uint32_t boff, uint32_t lo_off, int Kb, int k32, uint32_t idesc, uint32_t acc) {
    const uint32_t bh = sW + boff + (uint32_t)k32 * 8u * kLbo;
    const uint64_t dB_hi = umma::smem_desc(bh, kLbo, bsbo(Kb)), dB_lo = umma::smem_desc(bh + lo_off, kLbo, bsbo(Kb));
    umma::mma_3xtf32_k32(tm + dcol, dA_hi, dA_lo, dB_hi, dB_lo, kAStep, kStep, idesc, acc);
  };
  // A written by every thread of the group -> one thread issues the MMAs and commits them to the barrier
  auto publish_issue = [&](auto&& issue) {
    umma::fence_proxy_async();
    umma::tc_fence_before();
    asm volatile("bar.sync %0, 256;" ::"r"(grp + 1) : "memory");
    if (issuer) { umma::tc_fence_after(); issue(); umma::mma_commit(bar); }
  };
  // one lane of every warp polls the mbarrier (try_wait suspends the lane between polls)
  auto wait_mma = [&] {
    if (lane == 0) umma::mbar_wait(bar, phase);   // one polling lane per warp
    __syncwarp();
    phase ^= 1u;
    umma::tc_fence_after();
  };
  // this thread's 16 columns of its row -> A operand (hi and lo)
  auto store_half_row = [&](const float (&v)[16]) {
    const uint32_t base = (uint32_t)(row >> 3) * kASbo + (uint32_t)(row & 7) * 16u + (uint32_t)(4 * half) * kALbo;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float4 h, l;
      umma::split_tf32(v[4 * q], h.x, l.x); umma::split_tf32(v[4 * q + 1], h.y, l.y);
      umma::split_tf32(v[4 * q + 2], h.z, l.z); umma::split_tf32(v[4 * q + 3], h.w, l.w);
      *reinterpret_cast<float4*>(a_hi + base + q * kALbo) = h;
      *reinterpret_cast<float4*>(a_lo + base + q * kALbo) = l;
    }
  };
  // trilinear features of rows 32*quarter + 16*half + [0,16), gathered with 128-bit loads into registers (gather16) and
  // stored as the A operand later (put_rows).  Two lane mappings, both free of shared-memory bank conflicts:
  //   c_dim 32: 8 lanes per sample (lane & 7 = channel quad: one 128-byte voxel row per corner), 4 samples per iteration;
  //             the quarter-warp's stores (one row, quads 0..7) are kALbo = 144 bytes apart -> eight different bank groups;
  //   c_dim 64: no room for the padding, so the SAMPLE sits in the low lane bits (s = lane & 7) and a lane loads quads
  //             lane >> 3 and (lane >> 3) + 4 of its sample, 8 samples per iteration: a quarter-warp stores eight
  //             consecutive rows of one core matrix = 128 contiguous bytes.  (Costs L1 tag lookups -- 8 lines per load
  //             instruction instead of 4 -- which is why c_dim 32 prefers the padding.)
  auto gather16 = [&](const GridDev& g, float ux, float uy, float uz, unsigned vm, float* __restrict__ Cst, int64_t N, int64_t nq,
                      float4 (&out)[4]) {
    auto corners = [&](const Cell& c, int q) {
      const float4* gp = reinterpret_cast<const float4*>(g.data + c.base) + q;
      float4 val[8];
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int k = 0; k < 8; ++k)
        val[k] = ((c.ok >> k) & 1u) ? __ldg(gp + corner_offset(k, g.W, g.H) / 4) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        if ((c.ok >> k) & 1u) {
          const float w = corner_weight(c, k);
          v.x = __fadd_rn(v.x, __fmul_rn(val[k].x, w)); v.y = __fadd_rn(v.y, __fmul_rn(val[k].y, w));
          v.z = __fadd_rn(v.z, __fmul_rn(val[k].z, w)); v.w = __fadd_rn(v.w, __fmul_rn(val[k].w, w));
        }
      }
      return v;
    };
    if constexpr (CD == 32) {
      const int q = lane & 7, sub = lane >> 3;
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        const int src = 16 * half + 4 * it + sub;   // lane (within this warp) that owns the row
        const float sx = __shfl_sync(kFull, ux, src), sy = __shfl_sync(kFull, uy, src), sz = __shfl_sync(kFull, uz, src);
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if ((vm >> src) & 1u) {
          v = corners(make_cell(sx, sy, sz, g.W, g.H, g.D), q);
          if (Cst) reinterpret_cast<float4*>(Cst)[(int64_t)q * N + nq + src] = v;
        }
        out[it] = v;
      }
    } else {
      const int s = lane & 7, qh = lane >> 3;
#pragma unroll
      for (int it = 0; it < 2; ++it) {
        const int src = 16 * half + 8 * it + s;
        const float sx = __shfl_sync(kFull, ux, src), sy = __shfl_sync(kFull, uy, src), sz = __shfl_sync(kFull, uz, src);
        float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0;
        if ((vm >> src) & 1u) {
          const Cell c = make_cell(sx, sy, sz, g.W, g.H, g.D);
          v0 = corners(c, qh);
          v1 = corners(c, qh + 4);
          if (Cst) {
            reinterpret_cast<float4*>(Cst)[(int64_t)qh * N + nq + src] = v0;
            reinterpret_cast<float4*>(Cst)[(int64_t)(qh + 4) * N + nq + src] = v1;
          }
        }
        out[2 * it] = v0;
        out[2 * it + 1] = v1;
      }
    }
  };
  auto put_rows = [&](const float4 (&v)[4]) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int r, q;
      if constexpr (CD == 32) { q = lane & 7; r = quarter * 32 + 16 * half + 4 * j + (lane >> 3); }
      else { q = (lane >> 3) + 4 * (j & 1); r = quarter * 32 + 16 * half + 8 * (j >> 1) + (lane & 7); }
      float4 h, l;
      umma::split_tf32(v[j].x, h.x, l.x); umma::split_tf32(v[j].y, h.y, l.y);
      umma::split_tf32(v[j].z, h.z, l.z); umma::split_tf32(v[j].w, h.w, l.w);
      const uint32_t off = (uint32_t)(r >> 3) * kASbo + (uint32_t)(r & 7) * 16u + (uint32_t)q * kALbo;
      *reinterpret_cast<float4*>(a_hi + off) = h;
      *reinterpret_cast<float4*>(a_lo + off) = l;
    }
  };

  const int64_t N = a.pts.N, ntiles = (N + 127) / 128;
  // Tile order: a wave fills group 0 of every CTA before group 1, so the partial last wave is spread over as many
  // SMs as possible (a group that runs alone on its SM finishes its tile much sooner than one that shares it).
  for (int64_t t = (int64_t)blockIdx.x + (int64_t)gridDim.x * grp; t < ntiles; t += (int64_t)gridDim.x * 2) {
    const int64_t n = t * 128 + row;
    const bool valid = n < N;
    Sample sp;
    sp.pf[0] = sp.pf[1] = sp.pf[2] = 0.f; sp.xn[0] = sp.xn[1] = sp.xn[2] = 0.f; sp.inside = true;
    if (valid) load_sample(a.pts, n, a.nb, a.mb, sp);
    const unsigned vm = __ballot_sync(kFull, valid);
    const int64_t nq = t * 128 + quarter * 32;
    // ---- feature terms of all five blocks
    {
      float4 f[4];
      gather16(a.ga, unnormalise(sp.xn[0], a.ga.W), unnormalise(sp.xn[1], a.ga.H), unnormalise(sp.xn[2], a.ga.D), vm, a.C, N, nq, f);
      put_rows(f);
      publish_issue([&] { mma(0u, O_WC, 5u * bbytes(CD), CD, 0, idesc160, 0u); });
      if (CD == 64) {   // second grid: gathered while the first product runs
        gather16(a.gb, unnormalise(sp.xn[0], a.gb.W), unnormalise(sp.xn[1], a.gb.H), unnormalise(sp.xn[2], a.gb.D), vm,
                 a.C ? a.C + (int64_t)32 * N : nullptr, N, nq, f);
        wait_mma();
        put_rows(f);
        publish_issue([&] { mma(0u, O_WC, 5u * bbytes(CD), CD, 1, idesc160, 1u); });
      }
    }
    // ---- Fourier embedding, three K-chunks of 32 (16 columns per thread); chunk c+1 is computed
    //      while the product of chunk c (or of the features) is in flight
#pragma unroll 1
    for (int c = 0; c < 3; ++c) {
      float e[16];
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        const int kk = 32 * c + col0 + k;
        e[k] = fourier_sin(fmaf(sp.pf[2], sm[S_B + 192 + kk], fmaf(sp.pf[1], sm[S_B + 96 + kk], sp.pf[0] * sm[S_B + kk])));
      }
      if (a.E && valid) {
        float4* o = reinterpret_cast<float4*>(a.E);
#pragma unroll
        for (int q = 0; q < 4; ++q)
          o[(int64_t)(8 * c + 4 * half + q) * N + n] = make_float4(e[4 * q], e[4 * q + 1], e[4 * q + 2], e[4 * q + 3]);
      }
      wait_mma();
      store_half_row(e);
      publish_issue([&] { mma(160u, O_WE, 2u * bbytes(96), 96, c, idesc64, c > 0 ? 1u : 0u); });   // D1_0 and D1_3
    }
    // ---- blocks 0..3: each thread finishes its 16 columns; s2 = D2_l + bc_l is read ahead of the wait
    float s2[16];
    {
      float d2[16];
      tmem_ld16(tm_lane + col0, d2);
#pragma unroll
      for (int j = 0; j < 16; ++j) s2[j] = d2[j] + sm[S_BC + col0 + j];
    }
    wait_mma();
#pragma unroll 1
    for (int l = 0; l < 4; ++l) {
      float d1[16], h[16];
      const uint32_t c1 = (l == 0) ? 160u : (l == 3 ? 192u : 224u);
      tmem_ld16(tm_lane + c1 + col0, d1);
      uint32_t bits = 0;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float pre = d1[j] + sm[S_BIAS + l * 32 + col0 + j];
        if (BITS) bits |= (pre > 0.f) ? (1u << j) : 0u;
        h[j] = fmaxf(pre, 0.f) + s2[j];
      }
      store_half_row(h);
      publish_issue([&] {
        if (l == 2) mma(192u, O_WH + 4 * bbytes(32), bbytes(32), 32, 0, idesc32, 1u);          // D1_3 += W3[:, 93:] . h2
        else mma(224u, O_WH + (uint32_t)(l == 3 ? 6 : 2 * l) * bbytes(32), bbytes(32), 32, 0, idesc32, 0u);  // W1, W2, W4
      });
      // stash + the next block's feature term, under the product
      if (valid) {
        if (BITS) reinterpret_cast<uint16_t*>(a.relu_bits)[((int64_t)l * N + n) * 2 + half] = (uint16_t)bits;
        if (a.H) {
          float4* o = reinterpret_cast<float4*>(a.H + (int64_t)l * 32 * N);
#pragma unroll
          for (int q = 0; q < 4; ++q) o[(int64_t)(4 * half + q) * N + n] = make_float4(h[4 * q], h[4 * q + 1], h[4 * q + 2], h[4 * q + 3]);
        }
      }
      if (l < 3) {
        float d2[16];
        tmem_ld16(tm_lane + 32u * (l + 1) + col0, d2);
#pragma unroll
        for (int j = 0; j < 16; ++j) s2[j] = d2[j] + sm[S_BC + (l + 1) * 32 + col0 + j];
      }
      wait_mma();
    }
    // ---- block 4 + output layer: half 0 finishes the whole row (32 columns)
    if (half == 0) {
      float d1[32], d2[32], h[32];
      umma::tmem_ld32(tm_lane + 224u, d1);
      umma::tmem_ld32(tm_lane + 128u, d2);
      uint32_t bits = 0;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float pre = d1[j] + sm[S_BIAS + 128 + j];
        if (BITS) bits |= (pre > 0.f) ? (1u << j) : 0u;
        h[j] = fmaxf(pre, 0.f) + (d2[j] + sm[S_BC + 128 + j]);
      }
      float out[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int o = 0; o < NOUT; ++o) {
        float s = sm[S_BO + o];
#pragma unroll
        for (int j = 0; j < 32; ++j) s = fmaf(sm[S_WO + o * 32 + j], h[j], s);
        out[o] = s;
      }
      if (valid) {
        if (BITS) a.relu_bits[(int64_t)4 * N + n] = bits;
        if (a.H) store_planar32(a.H + (int64_t)4 * 32 * N, N, n, h);
        store_raw<NOUT>(a.raw, n, out, a.out_mode, a.apply_mask && !sp.inside);
      }
    }
    // this tile's TMEM reads (tcgen05.wait::ld) are ordered before the next tile's MMAs by the
    // fence + group barrier inside the next publish_and_issue
    umma::tc_fence_before();
  }
  umma::tc_fence_before();
  __syncthreads();
  if (tid < 32) umma::tmem_dealloc(tmem_base_s, 512);
}


template <int CD, int NOUT, bool BITS>
int launch_b(const FwdArgs& a, cudaStream_t st) {
  auto kern = k_grid_mlp_fwd_tc<CD, NOUT, BITS>;
  const size_t sm = tc::smem_total<CD>();
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
  const int64_t ntiles = (a.pts.N + 127) / 128;   // fewer tiles than SMs: one tile (group 0) per CTA
  const int grid = (int)((ntiles < (int64_t)sm_count()) ? ntiles : (int64_t)sm_count());
  launch_pdl(kern, dim3(grid), dim3(512), sm, st, a);
  return launch_status("k_grid_mlp_fwd_tc");
}

template <int CD, int NOUT>
int launch_t(const FwdArgs& a, cudaStream_t st) {
  return a.relu_bits ? launch_b<CD, NOUT, true>(a, st) : launch_b<CD, NOUT, false>(a, st);
}

}  // namespace

int launch_fwd_tc(int c_dim, int n_out, const FwdArgs& a, cudaStream_t st) {
  if (c_dim == 32) return n_out == 4 ? launch_t<32, 4>(a, st) : launch_t<32, 1>(a, st);
  return n_out == 4 ? launch_t<64, 4>(a, st) : launch_t<64, 1>(a, st);
}

}  // namespace pn
