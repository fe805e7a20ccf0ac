// C ABI of the NICE grid decoders (decoder.MLP, hidden 32) and the coarse-level decoder (decoder.MLP_no_xyz).
//
// The grid decoders run on the tensor cores (pn_gridmlp_tc_fwd.cu, pn_gridmlp_tc_bwd.cu, pn_wgrad_tc.cu); this file
// holds their entry points (argument checks, dispatch on c_dim / n_out / requested gradients) and the generic tiled
// weight-gradient GEMM used when only some gradient sinks are requested.
//
// The coarse decoder (one 77 KB grid, 32-wide layers without embedding, its own Mapper process upstream) is a SIMT
// kernel: a CTA of 256 threads owns 256 consecutive samples and alternates two phases per tile:
//   * feature phase, warp-cooperative: lane = channel.  For each of the warp's 32 samples the 8 voxel corners are
//     8 fully coalesced 128-byte rows of the channels-last grid; the interpolated feature goes to a padded shared
//     tile [channel][sample].  The backward walks the same pattern to scatter feature gradients (run-length
//     aggregated over consecutive samples that share a voxel, then coalesced RED) and to form the coordinate
//     gradient (warp-shuffle reduction over channels).
//   * MLP phase, thread = sample: activations live in registers, weights are staged once per CTA in shared memory
//     ([out][in] rows, 16-byte aligned) and read as warp-uniform LDS.128 broadcasts.
#include "pn_gridmlp.cuh"

namespace pn {
namespace {

// acc[j] += sum_k W[j][k] x[k]   (W in shared memory, row stride LD)
template <int K, int LD>
__device__ __forceinline__ void matvec(const float* __restrict__ W, const float (&x)[K], float (&acc)[32]) {
#pragma unroll
  for (int kc = 0; kc < K / 4; ++kc) {
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const float4 w = ld4(W + j * LD + 4 * kc);
      acc[j] = fmaf(w.x, x[4 * kc], acc[j]);
      acc[j] = fmaf(w.y, x[4 * kc + 1], acc[j]);
      acc[j] = fmaf(w.z, x[4 * kc + 2], acc[j]);
      acc[j] = fmaf(w.w, x[4 * kc + 3], acc[j]);
    }
  }
}

// gx[k] += sum_j W[j][k] g[j]
template <int K, int LD>
__device__ __forceinline__ void matvec_t(const float* __restrict__ W, const float (&g)[32], float (&gx)[K]) {
#pragma unroll
  for (int j = 0; j < 32; ++j) {
#pragma unroll
    for (int kc = 0; kc < K / 4; ++kc) {
      const float4 w = ld4(W + j * LD + 4 * kc);
      gx[4 * kc] = fmaf(w.x, g[j], gx[4 * kc]);
      gx[4 * kc + 1] = fmaf(w.y, g[j], gx[4 * kc + 1]);
      gx[4 * kc + 2] = fmaf(w.z, g[j], gx[4 * kc + 2]);
      gx[4 * kc + 3] = fmaf(w.w, g[j], gx[4 * kc + 3]);
    }
  }
}

// ---------------------------------------------------------------------------
// feature phase (warp-cooperative)
// ---------------------------------------------------------------------------
// Interpolated features of the warp's 32 samples -> rows [0,32) of `crow`
// (crow = &tile[lane * kLdc + warp_base]).
__device__ __forceinline__ void warp_gather(const GridDev& g, float ux, float uy, float uz, unsigned valid_mask,
                                            int lane, float* crow) {
  const float* gd = g.data + lane;
#pragma unroll 2
  for (int s = 0; s < 32; ++s) {
    const float sx = __shfl_sync(kFull, ux, s), sy = __shfl_sync(kFull, uy, s), sz = __shfl_sync(kFull, uz, s);
    float v = 0.f;
    if ((valid_mask >> s) & 1u) {
      const Cell c = make_cell(sx, sy, sz, g.W, g.H, g.D);
      float val[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) val[k] = ((c.ok >> k) & 1u) ? __ldg(gd + c.base + corner_offset(k, g.W, g.H)) : 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k)
        if ((c.ok >> k) & 1u) v = __fadd_rn(v, __fmul_rn(val[k], corner_weight(c, k)));
    }
    crow[s] = v;
  }
}

// Backward of warp_gather for one grid.  grow = &tile[lane*kLdc + warp_base]
// holds dL/dc for channel `lane` of the warp's samples.  Scatters into ggrid
// (optional) and returns, in lane s, dL/d(unnormalised coordinate) * gm of
// sample s (optional, via dux/duy/duz).
template <bool GRID_GRAD, bool NEED_DP>
__device__ __forceinline__ void warp_scatter(const GridDev& g, float* __restrict__ ggrid, float ux, float uy, float uz,
                                             unsigned valid_mask, int lane, const float* grow, float& dux, float& duy,
                                             float& duz) {
  const float* gd = g.data + lane;
  int64_t run_base = -1;
  unsigned run_ok = 0;
  float run[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) run[k] = 0.f;
  dux = duy = duz = 0.f;
  for (int s = 0; s < 32; ++s) {
    const float sx = __shfl_sync(kFull, ux, s), sy = __shfl_sync(kFull, uy, s), sz = __shfl_sync(kFull, uz, s);
    if (!((valid_mask >> s) & 1u)) continue;  // warp-uniform
    const Cell c = make_cell(sx, sy, sz, g.W, g.H, g.D);
    const float gc = grow[s];
    if (GRID_GRAD) {
      if (c.base != run_base) {  // warp-uniform: flush the finished run
        if (run_base >= 0) {
#pragma unroll
          for (int k = 0; k < 8; ++k)
            if ((run_ok >> k) & 1u) atomicAdd(ggrid + run_base + corner_offset(k, g.W, g.H) + lane, run[k]);
        }
        run_base = c.base; run_ok = c.ok;
#pragma unroll
        for (int k = 0; k < 8; ++k) run[k] = 0.f;
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) run[k] = fmaf(corner_weight(c, k), gc, run[k]);
    }
    if (NEED_DP) {
      float gx = 0.f, gy = 0.f, gz = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        if ((c.ok >> k) & 1u) {
          const float v = __ldg(gd + c.base + corner_offset(k, g.W, g.H));
          const float wx = c.wx[k & 1], wy = c.wy[(k >> 1) & 1], wz = c.wz[k >> 2];
          gx += ((k & 1) ? v : -v) * wy * wz;
          gy += (((k >> 1) & 1) ? v : -v) * wx * wz;
          gz += ((k >> 2) ? v : -v) * wx * wy;
        }
      }
      gx = warp_sum(gx * gc); gy = warp_sum(gy * gc); gz = warp_sum(gz * gc);
      if (lane == s) { dux = gx * c.gm[0]; duy = gy * c.gm[1]; duz = gz * c.gm[2]; }
    }
  }
  if (GRID_GRAD && run_base >= 0) {
#pragma unroll
    for (int k = 0; k < 8; ++k)
      if ((run_ok >> k) & 1u) atomicAdd(ggrid + run_base + corner_offset(k, g.W, g.H) + lane, run[k]);
  }
}

// ---------------------------------------------------------------------------
// weight gradients: out[j][k] += sum_n G[n][j] X[n][k] as tiled FFMA GEMMs
// ---------------------------------------------------------------------------
struct WgMat {
  const float* G;  // planar-4 block (N x 32)
  const float* X;  // planar-4 block (N x 32) (chunk of a wider matrix)
  float* out;      // &dW[0][col0]
  float* bias;     // optional column sums of G
  int ld, ncols;   // row stride of out, valid columns in this chunk (<=32)
};
constexpr int kMaxMats = 40;
struct WgArgs {
  WgMat m[kMaxMats];
  int nmats;
  int64_t N;
};

__global__ void __launch_bounds__(kThreads, 2) k_wgrad_gemm(const WgArgs a) {
  extern __shared__ __align__(16) float smem[];
  float* Gt = smem;                  // [256][32] swizzled rows
  float* Xt = smem + kThreads * 32;  // [256][32]
  const WgMat m = a.m[blockIdx.x];
  if (m.out == nullptr && m.bias == nullptr) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int jg = lane >> 2, kg = lane & 3;
  const int64_t N = a.N;
  float acc[4][8];
  float bsum[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[i][k] = 0.f;
  const float4* G4 = reinterpret_cast<const float4*>(m.G);
  const float4* X4 = reinterpret_cast<const float4*>(m.X);
  float4* Gs = reinterpret_cast<float4*>(Gt);
  float4* Xs = reinterpret_cast<float4*>(Xt);
  for (int64_t base = (int64_t)blockIdx.y * kThreads; base < N; base += (int64_t)gridDim.y * kThreads) {
    const int64_t n = base + tid;
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      Gs[tid * 8 + (q ^ (tid & 7))] = n < N ? G4[(int64_t)q * N + n] : z4;
      Xs[tid * 8 + (q ^ (tid & 7))] = n < N ? X4[(int64_t)q * N + n] : z4;
    }
    __syncthreads();
#pragma unroll 4
    for (int r = 0; r < 32; ++r) {
      const int s = warp * 32 + r;
      const float4 g = Gs[s * 8 + (jg ^ (s & 7))];
      const float4 x0 = Xs[s * 8 + ((2 * kg) ^ (s & 7))], x1 = Xs[s * 8 + ((2 * kg + 1) ^ (s & 7))];
      const float gv[4] = {g.x, g.y, g.z, g.w};
      const float xv[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        bsum[i] += gv[i];
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[i][k] = fmaf(gv[i], xv[k], acc[i][k]);
      }
    }
    __syncthreads();
  }
  // cross-warp reduction through shared memory (reuses Gt: 8 warps x 1024 floats)
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int k = 0; k < 8; ++k) Gt[warp * 1024 + (4 * jg + i) * 32 + 8 * kg + k] = acc[i][k];
  if (kg == 0) {
#pragma unroll
    for (int i = 0; i < 4; ++i) Xt[warp * 32 + 4 * jg + i] = bsum[i];
  }
  __syncthreads();
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int idx = tid * 4 + e;
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += Gt[w * 1024 + idx];
    const int j = idx >> 5, k = idx & 31;
    if (m.out && k < m.ncols) atomicAdd(m.out + (int64_t)j * m.ld + k, s);
  }
  if (m.bias && tid < 32) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += Xt[w * 32 + tid];
    atomicAdd(m.bias + tid, s);
  }
}

// dWo[o][k] += sum_n GO[n][o] H4[n][k],  dbo[o] += sum_n GO[n][o]
// warp = feature quad of h_4 (8 quads), lane = sample slot: every load is a coalesced 512-byte row of the
// planar-4 stash; 16 + 4 partial sums per thread, reduced over the 32 sample lanes by shuffles at the end.
__global__ void __launch_bounds__(kThreads) k_wgrad_out(const float* __restrict__ GO, const float* __restrict__ H4, int64_t N,
                                                       int nout, float* __restrict__ dWo, float* __restrict__ dbo) {
  const int lane = threadIdx.x & 31, q = threadIdx.x >> 5;
  float acc[4][4];
  float bs[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int o = 0; o < 4; ++o)
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[o][i] = 0.f;
  const float4* H = reinterpret_cast<const float4*>(H4) + (int64_t)q * N;
  const float4* G = reinterpret_cast<const float4*>(GO);
#pragma unroll 4
  for (int64_t n = (int64_t)blockIdx.x * 32 + lane; n < N; n += (int64_t)gridDim.x * 32) {
    const float4 h = H[n], g = G[n];
    const float gv[4] = {g.x, g.y, g.z, g.w}, hv[4] = {h.x, h.y, h.z, h.w};
#pragma unroll
    for (int o = 0; o < 4; ++o) {
      bs[o] += gv[o];
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[o][i] = fmaf(gv[o], hv[i], acc[o][i]);
    }
  }
#pragma unroll
  for (int o = 0; o < 4; ++o) {
    if (o < nout) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float v = warp_sum(acc[o][i]);
        if (lane == 0 && dWo) atomicAdd(dWo + o * 32 + 4 * q + i, v);
      }
      if (q == 0) {
        const float v = warp_sum(bs[o]);
        if (lane == 0 && dbo) atomicAdd(dbo + o, v);
      }
    }
  }
}

// dB[d][k] += sum_n P32[d][n] GARG[n][k]
// warp = embedding quad (24 quads, 768 threads), lane = sample slot; coalesced rows, 12 partial sums per thread.
__global__ void __launch_bounds__(768) k_wgrad_B(const float* __restrict__ P32, const float* __restrict__ GARG, int64_t N,
                                                float* __restrict__ dB) {
  const int lane = threadIdx.x & 31, q = threadIdx.x >> 5;
  float acc[3][4];
#pragma unroll
  for (int d = 0; d < 3; ++d)
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[d][i] = 0.f;
  const float4* G = reinterpret_cast<const float4*>(GARG) + (int64_t)q * N;
#pragma unroll 4
  for (int64_t n = (int64_t)blockIdx.x * 32 + lane; n < N; n += (int64_t)gridDim.x * 32) {
    const float4 g = G[n];
    const float gv[4] = {g.x, g.y, g.z, g.w};
    const float p[3] = {P32[n], P32[N + n], P32[2 * N + n]};
#pragma unroll
    for (int d = 0; d < 3; ++d)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[d][i] = fmaf(p[d], gv[i], acc[d][i]);
  }
#pragma unroll
  for (int d = 0; d < 3; ++d)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float v = warp_sum(acc[d][i]);
      const int k = 4 * q + i;
      if (lane == 0 && k < PN_EMBED) atomicAdd(dB + d * PN_EMBED + k, v);
    }
}

// ---------------------------------------------------------------------------
// coarse level: decoder.MLP_no_xyz (decoder.py:206-274).  h = c; five 32-wide
// blocks, block 3 sees cat[c, h]; no embedding, no per-block feature term.
// ---------------------------------------------------------------------------
constexpr int CO_W0 = 0, CO_W1 = 1024, CO_W2 = 2048, CO_W3 = 3072 /*[32][64]*/, CO_W4 = 5120, CO_B = 6144 /*[5][32]*/,
              CO_WO = 6304, CO_BO = 6336, CO_TOTAL = 6340;

struct CoarseDev { const float* W[5]; const float* b[5]; const float* Wo; const float* bo; };
inline CoarseDev make_coarse(const pn_coarse_mlp* w) {
  CoarseDev m;
  for (int i = 0; i < 5; ++i) { m.W[i] = w->W[i]; m.b[i] = w->b[i]; }
  m.Wo = w->Wo; m.bo = w->bo;
  return m;
}

__device__ void stage_coarse(const CoarseDev& m, float* wsm) {
  const int t = threadIdx.x, nt = blockDim.x;
  for (int i = t; i < 1024; i += nt) {
    wsm[CO_W0 + i] = m.W[0][i]; wsm[CO_W1 + i] = m.W[1][i]; wsm[CO_W2 + i] = m.W[2][i]; wsm[CO_W4 + i] = m.W[4][i];
  }
  for (int i = t; i < 2048; i += nt) wsm[CO_W3 + i] = m.W[3][i];
  for (int i = t; i < 160; i += nt) wsm[CO_B + i] = m.b[i >> 5][i & 31];
  for (int i = t; i < 32; i += nt) wsm[CO_WO + i] = m.Wo[i];
  if (t == 0) wsm[CO_BO] = m.bo[0];
}

struct CoarseFwdArgs {
  pn_points pts; CoarseDev w; GridDev g; Bound6 nb, mb;
  int apply_mask, out_mode;
  float* raw; uint32_t* relu_bits; float* H; float* C;
};

__device__ __forceinline__ uint32_t relu_inplace(float (&a)[32]) {
  uint32_t bits = 0;
#pragma unroll
  for (int j = 0; j < 32; ++j) { bits |= (a[j] > 0.f) ? (1u << j) : 0u; a[j] = fmaxf(a[j], 0.f); }
  return bits;
}

__global__ void __launch_bounds__(kThreads, 2) k_coarse_fwd(const CoarseFwdArgs a) {
  extern __shared__ __align__(16) float smem[];
  float* wsm = smem;
  float* tile = smem + CO_TOTAL;
  stage_coarse(a.w, wsm);
  __syncthreads();
  const int tid = threadIdx.x, lane = tid & 31, wbase = tid & ~31;
  const int64_t N = a.pts.N, ntiles = (N + kThreads - 1) / kThreads;
  for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const int64_t n = t * kThreads + tid;
    const bool valid = n < N;
    Sample sp;
    sp.pf[0] = sp.pf[1] = sp.pf[2] = 0.f; sp.xn[0] = sp.xn[1] = sp.xn[2] = 0.f; sp.inside = true;
    if (valid) load_sample(a.pts, n, a.nb, a.mb, sp);
    const unsigned vm = __ballot_sync(kFull, valid);
    __syncwarp();
    warp_gather(a.g, unnormalise(sp.xn[0], a.g.W), unnormalise(sp.xn[1], a.g.H), unnormalise(sp.xn[2], a.g.D), vm, lane,
                tile + lane * kLdc + wbase);
    __syncwarp();
    float c[32], h[32], acc[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) c[j] = tile[j * kLdc + tid];
    if (a.C && valid) store_planar32(a.C, N, n, c);
#pragma unroll
    for (int j = 0; j < 32; ++j) h[j] = c[j];
#pragma unroll 1
    for (int l = 0; l < 5; ++l) {
#pragma unroll
      for (int j = 0; j < 32; ++j) acc[j] = wsm[CO_B + l * 32 + j];
      if (l == 3) {
        matvec<32, 64>(wsm + CO_W3, c, acc);
        matvec<32, 64>(wsm + CO_W3 + 32, h, acc);
      } else {
        const int off = (l == 4) ? CO_W4 : (l == 2) ? CO_W2 : (l == 1) ? CO_W1 : CO_W0;
        matvec<32, 32>(wsm + off, h, acc);
      }
      const uint32_t bits = relu_inplace(acc);
#pragma unroll
      for (int j = 0; j < 32; ++j) h[j] = acc[j];
      if (valid) {
        if (a.relu_bits) a.relu_bits[(int64_t)l * N + n] = bits;
        if (a.H) store_planar32(a.H + (int64_t)l * 32 * N, N, n, h);
      }
    }
    float out = wsm[CO_BO];
#pragma unroll
    for (int j = 0; j < 32; ++j) out = fmaf(wsm[CO_WO + j], h[j], out);
    if (valid) {
      const float o4[4] = {out, 0.f, 0.f, 0.f};
      store_raw<1>(a.raw, n, o4, a.out_mode, a.apply_mask && !sp.inside);
    }
  }
}

struct CoarseBwdArgs {
  pn_points pts; CoarseDev w; GridDev g; Bound6 nb, mb;
  int apply_mask;
  const float* g_raw; const uint32_t* relu_bits;
  float* g_grid; float* g_pts; float* GA; float* GO;
};

template <bool GRID_GRAD, bool NEED_DP, bool WS>
__global__ void __launch_bounds__(kThreads, 1) k_coarse_bwd(const CoarseBwdArgs a) {
  extern __shared__ __align__(16) float smem[];
  float* wsm = smem;
  float* tile = smem + CO_TOTAL;
  stage_coarse(a.w, wsm);
  __syncthreads();
  const int tid = threadIdx.x, lane = tid & 31, wbase = tid & ~31;
  const int64_t N = a.pts.N, ntiles = (N + kThreads - 1) / kThreads;
  for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const int64_t n = t * kThreads + tid;
    const bool valid = n < N;
    Sample sp;
    sp.pf[0] = sp.pf[1] = sp.pf[2] = 0.f; sp.xn[0] = sp.xn[1] = sp.xn[2] = 0.f; sp.inside = true;
    if (valid) load_sample(a.pts, n, a.nb, a.mb, sp);
    const unsigned vm = __ballot_sync(kFull, valid);
    float go = 0.f;
    if (valid) {
      go = reinterpret_cast<const float4*>(a.g_raw)[n].w;
      if (a.apply_mask && !sp.inside) go = 0.f;
    }
    if (WS && valid) reinterpret_cast<float4*>(a.GO)[n] = make_float4(go, 0.f, 0.f, 0.f);
    float gh[32], gx[32], gc[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) { gh[j] = wsm[CO_WO + j] * go; gc[j] = 0.f; }
#pragma unroll 1
    for (int l = 4; l >= 0; --l) {
      const uint32_t bits = valid ? a.relu_bits[(int64_t)l * N + n] : 0u;
#pragma unroll
      for (int j = 0; j < 32; ++j) gh[j] = ((bits >> j) & 1u) ? gh[j] : 0.f;
      if (WS && valid) store_planar32(a.GA + (int64_t)l * 32 * N, N, n, gh);
#pragma unroll
      for (int j = 0; j < 32; ++j) gx[j] = 0.f;
      if (l == 3) {
        matvec_t<32, 64>(wsm + CO_W3, gh, gc);
        matvec_t<32, 64>(wsm + CO_W3 + 32, gh, gx);
      } else {
        const int off = (l == 4) ? CO_W4 : (l == 2) ? CO_W2 : (l == 1) ? CO_W1 : CO_W0;
        matvec_t<32, 32>(wsm + off, gh, gx);
      }
      if (l == 0) {
#pragma unroll
        for (int j = 0; j < 32; ++j) gc[j] += gx[j];
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) gh[j] = gx[j];
      }
    }
    if (GRID_GRAD || NEED_DP) {
      __syncwarp();
#pragma unroll
      for (int j = 0; j < 32; ++j) tile[j * kLdc + tid] = gc[j];
      __syncwarp();
      float dux, duy, duz;
      warp_scatter<GRID_GRAD, NEED_DP>(a.g, a.g_grid, unnormalise(sp.xn[0], a.g.W), unnormalise(sp.xn[1], a.g.H),
                                       unnormalise(sp.xn[2], a.g.D), vm, lane, tile + lane * kLdc + wbase, dux, duy, duz);
      if (NEED_DP && valid) {
        float* o = a.g_pts + 3 * n;
        o[0] += norm_grad(a.pts, a.nb, 0, dux); o[1] += norm_grad(a.pts, a.nb, 1, duy); o[2] += norm_grad(a.pts, a.nb, 2, duz);
      }
    }
  }
}

template <bool GG, bool DP, bool WS>
int launch_coarse_bwd_t(const CoarseBwdArgs& a, cudaStream_t st) {
  auto kern = k_coarse_bwd<GG, DP, WS>;
  const size_t sm = (size_t)(CO_TOTAL + 32 * kLdc) * sizeof(float);
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
  const int64_t ntiles = (a.pts.N + kThreads - 1) / kThreads;
  const int grid = (int)((ntiles < (int64_t)sm_count()) ? ntiles : (int64_t)sm_count());
  kern<<<grid, kThreads, sm, st>>>(a);
  return launch_status("k_coarse_bwd");
}

bool check_mlp(const pn_grid_mlp* w, const char* fn) {
  if (!w) { set_error("%s: null decoder", fn); return false; }
  if (!((w->c_dim == 32 || w->c_dim == 64) && (w->n_out == 1 || w->n_out == 4))) {
    set_error("%s: unsupported decoder shape c_dim=%d n_out=%d (want 32|64, 1|4)", fn, w->c_dim, w->n_out);
    return false;
  }
  if (!w->B || !w->Wo || !w->bo) { set_error("%s: null parameter pointer", fn); return false; }
  for (int i = 0; i < 5; ++i)
    if (!w->W[i] || !w->b[i] || !w->Wc[i] || !w->bc[i]) { set_error("%s: null parameter pointer", fn); return false; }
  return true;
}

bool check_points(const pn_points* p, const char* fn) {
  if (!p || p->N < 0) { set_error("%s: bad point source", fn); return false; }
  const int modes = (p->pts64 ? 1 : 0) + (p->pts32 ? 1 : 0) + ((p->rays_o && p->rays_d && p->z) ? 1 : 0);
  if (modes != 1) { set_error("%s: exactly one of pts64 / pts32 / (rays_o,rays_d,z) must be set", fn); return false; }
  if (!p->pts64 && !p->pts32 && p->S <= 0) { set_error("%s: ray mode needs S > 0", fn); return false; }
  return true;
}

}  // namespace
}  // namespace pn

using namespace pn;

extern "C" int pn_grid_mlp_fwd(const pn_points* pts, const pn_grid_mlp* w, const pn_grid* gridA, const pn_grid* gridB,
                               const double* norm_bound, const double* mask_bound, int apply_mask, int out_mode,
                               float* raw, const pn_stash* stash, void* stream) {
  if (!check_points(pts, "pn_grid_mlp_fwd") || !check_mlp(w, "pn_grid_mlp_fwd")) return 1;
  if (!gridA || !gridA->data || (w->c_dim == 64 && (!gridB || !gridB->data)) || !raw || !norm_bound) {
    set_error("pn_grid_mlp_fwd: null grid/raw/bound");
    return 1;
  }
  if (apply_mask && !mask_bound) { set_error("pn_grid_mlp_fwd: apply_mask needs mask_bound"); return 1; }
  if (out_mode < PN_OUT_SET_ALL || out_mode > PN_OUT_SET_RGB || (out_mode == PN_OUT_SET_RGB && w->n_out != 4)) {
    set_error("pn_grid_mlp_fwd: bad out_mode %d for n_out %d", out_mode, w->n_out);
    return 1;
  }
  if (pts->N == 0) return 0;
  FwdArgs a;
  a.pts = *pts; a.w = make_mlp(w); a.ga = make_grid(gridA); a.gb = make_grid(gridB);
  a.nb = make_bound(norm_bound); a.mb = make_bound(mask_bound ? mask_bound : norm_bound);
  a.apply_mask = apply_mask; a.out_mode = out_mode; a.raw = raw;
  a.relu_bits = stash ? stash->relu_bits : nullptr; a.H = stash ? stash->H : nullptr;
  a.C = stash ? stash->C : nullptr; a.E = stash ? stash->E : nullptr;
  cudaStream_t st = (cudaStream_t)stream;
  return launch_fwd_tc(w->c_dim, w->n_out, a, st);
}

extern "C" int pn_grid_mlp_bwd(const pn_points* pts, const pn_grid_mlp* w, const pn_grid* gridA, const pn_grid* gridB,
                               const double* norm_bound, const double* mask_bound, int apply_mask, const float* g_raw,
                               const pn_stash* stash, float* g_gridA, float* g_pts, int accumulate_pts,
                               const pn_wscratch* ws, void* stream) {
  if (!check_points(pts, "pn_grid_mlp_bwd") || !check_mlp(w, "pn_grid_mlp_bwd")) return 1;
  if (!gridA || !gridA->data || !g_raw || !stash || !stash->relu_bits || !norm_bound) {
    set_error("pn_grid_mlp_bwd: null grid/g_raw/stash/bound");
    return 1;
  }
  if (ws && (!ws->GA || !ws->GH || !ws->GARG || !ws->P32 || !ws->GO)) {
    set_error("pn_grid_mlp_bwd: incomplete weight-gradient scratch");
    return 1;
  }
  if (pts->N == 0) return 0;
  BwdArgs a;
  a.pts = *pts; a.w = make_mlp(w); a.ga = make_grid(gridA); a.gb = make_grid(gridB);
  a.nb = make_bound(norm_bound); a.mb = make_bound(mask_bound ? mask_bound : norm_bound);
  a.apply_mask = apply_mask; a.accumulate_pts = accumulate_pts; a.g_raw = g_raw; a.relu_bits = stash->relu_bits;
  a.g_grid = g_gridA; a.g_pts = g_pts;
  a.GA = ws ? ws->GA : nullptr; a.GH = ws ? ws->GH : nullptr; a.GARG = ws ? ws->GARG : nullptr;
  a.P32 = ws ? ws->P32 : nullptr; a.GO = ws ? ws->GO : nullptr;
  cudaStream_t st = (cudaStream_t)stream;
  const bool gg = g_gridA != nullptr, dp = g_pts != nullptr, wsb = ws != nullptr;
  return launch_bwd_tc(w->c_dim, w->n_out, gg, dp, wsb, a, st);
}

extern "C" int pn_grid_mlp_wgrad(int64_t N, const pn_grid_mlp* w, const pn_stash* stash, const pn_wscratch* ws,
                                 const pn_grid_mlp_grad* g, void* stream) {
  if (!check_mlp(w, "pn_grid_mlp_wgrad")) return 1;
  if (!stash || !stash->H || !stash->C || !stash->E || !ws || !ws->GA || !ws->GH || !ws->GARG || !ws->P32 || !ws->GO || !g) {
    set_error("pn_grid_mlp_wgrad: stash/scratch/gradient sinks incomplete");
    return 1;
  }
  if (N == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (w->c_dim == 32) {   // one fused tensor-core kernel (GA = GH masked by the ReLU bits)
    if (!stash->relu_bits) {   // the tensor-core backward does not write GA for c_dim 32
      set_error("pn_grid_mlp_wgrad: stash->relu_bits is required");
      return 1;
    }
    return launch_wgrad_tc32(N, w->n_out, stash->H, stash->C, stash->E, ws->GH, ws->GARG, ws->GO, ws->P32, stash->relu_bits, g->W,
                             g->b, g->Wc, g->bc, g->Wo, g->bo, g->B, st);
  }
  const int64_t blk = 32 * N;  // floats per planar-4 (N x 32) block
  WgArgs a;
  a.N = N;
  int nm = 0;
  auto add = [&](const float* G, const float* X, float* out, float* bias, int ld, int ncols) {
    if (!out && !bias) return;
    a.m[nm++] = WgMat{G, X, out, bias, ld, ncols};
  };
  const int cd = w->c_dim;
  for (int l = 0; l < 5; ++l) {
    const float* GA = ws->GA + l * blk;
    const float* GH = ws->GH + l * blk;
    if (l == 0 || l == 3) {
      const int ld = l == 0 ? PN_EMBED : PN_EMBED + 32;
      for (int c = 0; c < 3; ++c)
        add(GA, stash->E + c * blk, g->W[l] ? g->W[l] + 32 * c : nullptr, c == 0 ? g->b[l] : nullptr, ld,
            c == 2 ? PN_EMBED - 64 : 32);
      if (l == 3) add(GA, stash->H + 2 * blk, g->W[3] ? g->W[3] + PN_EMBED : nullptr, nullptr, ld, 32);
    } else {
      add(GA, stash->H + (l - 1) * blk, g->W[l], g->b[l], 32, 32);
    }
    for (int c = 0; c < cd / 32; ++c)
      add(GH, stash->C + c * blk, g->Wc[l] ? g->Wc[l] + 32 * c : nullptr, c == 0 ? g->bc[l] : nullptr, cd, 32);
  }
  a.nmats = nm;
  if (nm > 0) {  // all W / b / Wc / bc sinks present -> tensor-core GEMMs
    const int rc = launch_wgrad_tc(N, cd, stash->H, stash->C, stash->E, ws->GA, ws->GH, g->W, g->b, g->Wc, g->bc, st);
    if (rc > 0) return 1;
    if (rc == 0) nm = 0;
  }
  if (nm > 0) {
    int64_t tiles = (N + kThreads - 1) / kThreads;
    int split = (int)((2 * sm_count() + nm - 1) / nm);
    if (split > tiles) split = (int)tiles;
    if (split < 1) split = 1;
    const int sm = 2 * kThreads * 32 * (int)sizeof(float);
    cudaFuncSetAttribute(k_wgrad_gemm, cudaFuncAttributeMaxDynamicSharedMemorySize, sm);
    k_wgrad_gemm<<<dim3(nm, split), kThreads, sm, st>>>(a);
    if (launch_status("k_wgrad_gemm")) return 1;
  }
  if (g->Wo || g->bo) {
    int grid = (int)((N + 31) / 32); if (grid > 2 * sm_count()) grid = 2 * sm_count();
    k_wgrad_out<<<grid, kThreads, 0, st>>>(ws->GO, stash->H + 4 * blk, N, w->n_out, g->Wo, g->bo);
    if (launch_status("k_wgrad_out")) return 1;
  }
  if (g->B) {
    int grid = (int)((N + 31) / 32); if (grid > 2 * sm_count()) grid = 2 * sm_count();
    k_wgrad_B<<<grid, 768, 0, st>>>(ws->P32, ws->GARG, N, g->B);
    if (launch_status("k_wgrad_B")) return 1;
  }
  return 0;
}

// ------------------------------------------------------------------ coarse level C ABI
static bool check_coarse(const pn_coarse_mlp* w, const char* fn) {
  if (!w || !w->Wo || !w->bo) { set_error("%s: null decoder", fn); return false; }
  for (int i = 0; i < 5; ++i)
    if (!w->W[i] || !w->b[i]) { set_error("%s: null parameter pointer", fn); return false; }
  return true;
}

extern "C" int pn_coarse_mlp_fwd(const pn_points* pts, const pn_coarse_mlp* w, const pn_grid* grid, const double* norm_bound,
                                 const double* mask_bound, int apply_mask, int out_mode, float* raw, const pn_stash* stash,
                                 void* stream) {
  if (!check_points(pts, "pn_coarse_mlp_fwd") || !check_coarse(w, "pn_coarse_mlp_fwd")) return 1;
  if (!grid || !grid->data || !raw || !norm_bound || (apply_mask && !mask_bound)) {
    set_error("pn_coarse_mlp_fwd: null grid/raw/bound");
    return 1;
  }
  if (out_mode < PN_OUT_SET_ALL || out_mode > PN_OUT_ADD_W) { set_error("pn_coarse_mlp_fwd: bad out_mode %d", out_mode); return 1; }
  if (pts->N == 0) return 0;
  CoarseFwdArgs a;
  a.pts = *pts; a.w = make_coarse(w); a.g = make_grid(grid);
  a.nb = make_bound(norm_bound); a.mb = make_bound(mask_bound ? mask_bound : norm_bound);
  a.apply_mask = apply_mask; a.out_mode = out_mode; a.raw = raw;
  a.relu_bits = stash ? stash->relu_bits : nullptr; a.H = stash ? stash->H : nullptr; a.C = stash ? stash->C : nullptr;
  const size_t sm = (size_t)(CO_TOTAL + 32 * kLdc) * sizeof(float);
  cudaFuncSetAttribute(k_coarse_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
  const int64_t ntiles = (pts->N + kThreads - 1) / kThreads;
  const int grid_n = (int)((ntiles < (int64_t)sm_count() * 2) ? ntiles : (int64_t)sm_count() * 2);
  k_coarse_fwd<<<grid_n, kThreads, sm, (cudaStream_t)stream>>>(a);
  return launch_status("k_coarse_fwd");
}

extern "C" int pn_coarse_mlp_bwd(const pn_points* pts, const pn_coarse_mlp* w, const pn_grid* grid, const double* norm_bound,
                                 const double* mask_bound, int apply_mask, const float* g_raw, const pn_stash* stash,
                                 float* g_grid, float* g_pts, int accumulate_pts, const pn_wscratch* ws, void* stream) {
  if (!check_points(pts, "pn_coarse_mlp_bwd") || !check_coarse(w, "pn_coarse_mlp_bwd")) return 1;
  if (!grid || !grid->data || !g_raw || !stash || !stash->relu_bits || !norm_bound) {
    set_error("pn_coarse_mlp_bwd: null grid/g_raw/stash/bound");
    return 1;
  }
  if (ws && (!ws->GA || !ws->GO)) { set_error("pn_coarse_mlp_bwd: incomplete weight-gradient scratch"); return 1; }
  if (g_pts && !accumulate_pts) { set_error("pn_coarse_mlp_bwd: g_pts is accumulate-only (zero it first)"); return 1; }
  if (pts->N == 0) return 0;
  CoarseBwdArgs a;
  a.pts = *pts; a.w = make_coarse(w); a.g = make_grid(grid);
  a.nb = make_bound(norm_bound); a.mb = make_bound(mask_bound ? mask_bound : norm_bound);
  a.apply_mask = apply_mask; a.g_raw = g_raw; a.relu_bits = stash->relu_bits;
  a.g_grid = g_grid; a.g_pts = g_pts; a.GA = ws ? ws->GA : nullptr; a.GO = ws ? ws->GO : nullptr;
  cudaStream_t st = (cudaStream_t)stream;
  const int key = (g_grid ? 4 : 0) | (g_pts ? 2 : 0) | (ws ? 1 : 0);
  switch (key) {
    case 0: return launch_coarse_bwd_t<false, false, false>(a, st);
    case 1: return launch_coarse_bwd_t<false, false, true>(a, st);
    case 2: return launch_coarse_bwd_t<false, true, false>(a, st);
    case 3: return launch_coarse_bwd_t<false, true, true>(a, st);
    case 4: return launch_coarse_bwd_t<true, false, false>(a, st);
    case 5: return launch_coarse_bwd_t<true, false, true>(a, st);
    case 6: return launch_coarse_bwd_t<true, true, false>(a, st);
    default: return launch_coarse_bwd_t<true, true, true>(a, st);
  }
}

extern "C" int pn_coarse_mlp_wgrad(int64_t N, const pn_stash* stash, const pn_wscratch* ws, const pn_coarse_mlp_grad* g,
                                   void* stream) {
  if (!stash || !stash->H || !stash->C || !ws || !ws->GA || !ws->GO || !g) {
    set_error("pn_coarse_mlp_wgrad: stash/scratch/gradient sinks incomplete");
    return 1;
  }
  if (N == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t blk = 32 * N;
  WgArgs a;
  a.N = N;
  int nm = 0;
  auto add = [&](const float* G, const float* X, float* out, float* bias, int ld) {
    if (!out && !bias) return;
    a.m[nm++] = WgMat{G, X, out, bias, ld, 32};
  };
  add(ws->GA, stash->C, g->W[0], g->b[0], 32);
  add(ws->GA + blk, stash->H, g->W[1], g->b[1], 32);
  add(ws->GA + 2 * blk, stash->H + blk, g->W[2], g->b[2], 32);
  add(ws->GA + 3 * blk, stash->C, g->W[3], g->b[3], 64);
  add(ws->GA + 3 * blk, stash->H + 2 * blk, g->W[3] ? g->W[3] + 32 : nullptr, nullptr, 64);
  add(ws->GA + 4 * blk, stash->H + 3 * blk, g->W[4], g->b[4], 32);
  a.nmats = nm;
  if (nm > 0) {
    const int64_t tiles = (N + kThreads - 1) / kThreads;
    int split = (2 * sm_count() + nm - 1) / nm;
    if (split > tiles) split = (int)tiles;
    if (split < 1) split = 1;
    const int sm = 2 * kThreads * 32 * (int)sizeof(float);
    cudaFuncSetAttribute(k_wgrad_gemm, cudaFuncAttributeMaxDynamicSharedMemorySize, sm);
    k_wgrad_gemm<<<dim3(nm, split), kThreads, sm, st>>>(a);
    if (launch_status("k_wgrad_gemm")) return 1;
  }
  if (g->Wo || g->bo) {
    int grid = (int)((N + 31) / 32); if (grid > 2 * sm_count()) grid = 2 * sm_count();
    k_wgrad_out<<<grid, kThreads, 0, st>>>(ws->GO, stash->H + 4 * blk, N, 1, g->Wo, g->bo);
    if (launch_status("k_wgrad_out")) return 1;
  }
  return 0;
}
