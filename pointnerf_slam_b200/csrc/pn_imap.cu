// iMAP* single-MLP decoder (decoder.MLP with c_dim = 0: Fourier-93 -> n_blocks x
// hidden relu layers -> 4 outputs; src/conv_onet/config.py:28-32,
// src/conv_onet/models/decoder.py:189-203).  At hidden 256 the layers are real
// GEMMs ([N x 256] . [256 x 256]), so the path is a chain of GEMM launches with
// fused epilogues (bias+relu, relu-mask, split-K atomics) over row-major
// activations kept in HBM (1 KB per sample and layer), plus small kernels for the
// embedding and the 4-wide output layer.  The GEMMs run on the tensor cores
// (tcgen05, 3xTF32; pn_imap_tc.cu); the FFMA tiles below serve the slivers
// (4- and 3-row weight gradients) and odd shapes.
#include "pn_common.cuh"

namespace pn {
namespace {

constexpr int TM = 64, TN = 64, TK = 16;  // CTA tile; 256 threads, 4x4 micro-tile

enum { EP_STORE = 0, EP_BIAS_RELU = 1, EP_MASK = 2, EP_ATOMIC = 3 };

// C[M x N] (op)= sum_k A(m,k) * B(k,n) with A(m,k) = A[m*sam + k*sak], B(k,n) = B[k*sbk + n*sbn].
// gridDim.z splits K (EP_ATOMIC only).  EP_MASK: C = acc where aux[m*ldc+n] > 0 else 0.
template <int EP>
__global__ void __launch_bounds__(256) k_sgemm(const float* __restrict__ A, int64_t sam, int64_t sak,
                                              const float* __restrict__ B, int64_t sbk, int64_t sbn,
                                              float* __restrict__ C, int64_t ldc, int64_t M, int N, int64_t K,
                                              const float* __restrict__ bias, const float* __restrict__ aux) {
  __shared__ float As[TK][TM + 4];
  __shared__ float Bs[TK][TN + 4];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int64_t m0 = (int64_t)blockIdx.y * TM;
  const int n0 = blockIdx.x * TN;
  int64_t k_begin = 0, k_end = K;
  if (EP == EP_ATOMIC) {
    const int64_t per = (K + gridDim.z - 1) / gridDim.z;
    k_begin = (int64_t)blockIdx.z * per;
    k_end = k_begin + per < K ? k_begin + per : K;
  }
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  // loader mapping: choose the thread->element order whose fastest index is contiguous in memory
  const bool a_k_contig = sak == 1;
  const bool b_n_contig = sbn == 1;
  for (int64_t k0 = k_begin; k0 < k_end; k0 += TK) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int idx = tid + e * 256;  // 0..1023 over TM x TK
      int m, k;
      if (a_k_contig) { k = idx & (TK - 1); m = idx >> 4; } else { m = idx & (TM - 1); k = idx >> 6; }
      const int64_t gm = m0 + m, gk = k0 + k;
      As[k][m] = (gm < M && gk < k_end) ? A[gm * sam + gk * sak] : 0.f;
      int n, kb;
      if (b_n_contig) { n = idx & (TN - 1); kb = idx >> 6; } else { kb = idx & (TK - 1); n = idx >> 4; }
      const int64_t gkb = k0 + kb;
      const int gn = n0 + n;
      Bs[kb][n] = (gn < N && gkb < k_end) ? B[gkb * sbk + (int64_t)gn * sbn] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < TK; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t gm = m0 + ty * 4 + i;
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gn = n0 + tx * 4 + j;
      if (gn >= N) continue;
      float v = acc[i][j];
      float* c = C + gm * ldc + gn;
      if (EP == EP_BIAS_RELU) *c = fmaxf(v + bias[gn], 0.f);
      else if (EP == EP_MASK) *c = aux[gm * ldc + gn] > 0.f ? v : 0.f;
      else if (EP == EP_ATOMIC) atomicAdd(c, v);
      else *c = v;
    }
  }
}

template <int EP>
int sgemm(const float* A, int64_t sam, int64_t sak, const float* B, int64_t sbk, int64_t sbn, float* C, int64_t ldc,
          int64_t M, int N, int64_t K, const float* bias, const float* aux, int splitk, cudaStream_t st) {
  dim3 grid((N + TN - 1) / TN, (unsigned)((M + TM - 1) / TM), EP == EP_ATOMIC ? splitk : 1);
  k_sgemm<EP><<<grid, 256, 0, st>>>(A, sam, sak, B, sbk, sbn, C, ldc, M, N, K, bias, aux);
  return launch_status("k_sgemm");
}

// Thin weight gradients (dWo: 4 rows, dB: 3 rows): out[r][c] += sum_n A(r,n) * X[n*ldx + c], A(r,n) = A[r*sar + n*san].
// Thread = column (coalesced reads of X's rows), R accumulators in registers, one atomicAdd per thread and row: X is read
// once at streaming speed instead of going through 64-row GEMM tiles that would be 94 % padding.
template <int R>
__global__ void __launch_bounds__(256) k_thin_wgrad(const float* __restrict__ A, int64_t sar, int64_t san, const float* __restrict__ X,
                                                    int64_t ldx, int64_t N, int C, float* __restrict__ out, int ldo) {
  const int c = threadIdx.x;
  const int64_t per = (N + gridDim.x - 1) / gridDim.x;
  const int64_t n0 = (int64_t)blockIdx.x * per, n1 = n0 + per < N ? n0 + per : N;
  float acc[R];
#pragma unroll
  for (int r = 0; r < R; ++r) acc[r] = 0.f;
  if (c < C) {
    for (int64_t n = n0; n < n1; ++n) {
      const float x = X[n * ldx + c];
#pragma unroll
      for (int r = 0; r < R; ++r) acc[r] = fmaf(__ldg(A + r * sar + n * san), x, acc[r]);
    }
#pragma unroll
    for (int r = 0; r < R; ++r) atomicAdd(out + r * ldo + c, acc[r]);
  }
}

// the layer GEMMs: tensor cores (pn_imap_tc.cu) whenever the shape fills 128-row tiles, the FFMA tiles for the slivers
// (dWo: 4 rows, dB: 3 rows)
template <int EP>
int gemm(const float* A, int64_t sam, int64_t sak, const float* B, int64_t sbk, int64_t sbn, float* C, int64_t ldc,
         int64_t M, int N, int64_t K, const float* bias, const float* aux, int splitk, cudaStream_t st) {
  const int rc = tc_gemm(EP, A, sam, sak, B, sbk, sbn, C, ldc, M, N, K, bias, aux, splitk, nullptr, st);
  if (rc >= 0) return rc;
  return sgemm<EP>(A, sam, sak, B, sbk, sbn, C, ldc, M, N, K, bias, aux, splitk, st);
}

// E[n][96] = sin(p . B[:,k]) (cols 93..95 = 0); optionally P32 [3][N]
__global__ void k_imap_embed(pn_points pts, Bound6 nb, Bound6 mb, const float* __restrict__ Bm, float* __restrict__ E,
                             float* __restrict__ P32) {
  __shared__ float Bs[3 * PN_EMBED];
  for (int i = threadIdx.x; i < 3 * PN_EMBED; i += blockDim.x) Bs[i] = Bm[i];
  __syncthreads();
  const int64_t n = (int64_t)blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5);
  if (n >= pts.N) return;
  const int lane = threadIdx.x & 31;
  Sample sp;
  load_sample(pts, n, nb, mb, sp);
  if (P32 && lane < 3) P32[(int64_t)lane * pts.N + n] = sp.pf[lane];
  for (int k = lane; k < 96; k += 32) {
    float v = 0.f;
    if (k < PN_EMBED) v = fourier_sin(fmaf(sp.pf[2], Bs[2 * PN_EMBED + k], fmaf(sp.pf[1], Bs[PN_EMBED + k], sp.pf[0] * Bs[k])));
    E[n * 96 + k] = v;
  }
}

// raw[n] = H[n] . Wo^T + bo (4 outputs), with the out-of-bound override on component 3
__global__ void k_imap_out(pn_points pts, Bound6 nb, Bound6 mb, int apply_mask, const float* __restrict__ Hl, int hidden,
                           const float* __restrict__ Wo, const float* __restrict__ bo, float* __restrict__ raw) {
  const int64_t n = (int64_t)blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5);
  if (n >= pts.N) return;
  const int lane = threadIdx.x & 31;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int k = lane; k < hidden; k += 32) {
    const float h = Hl[n * hidden + k];
#pragma unroll
    for (int o = 0; o < 4; ++o) acc[o] = fmaf(Wo[o * hidden + k], h, acc[o]);
  }
#pragma unroll
  for (int o = 0; o < 4; ++o) acc[o] = warp_sum(acc[o]);
  if (lane == 0) {
    float4 v = make_float4(acc[0] + bo[0], acc[1] + bo[1], acc[2] + bo[2], acc[3] + bo[3]);
    if (apply_mask) {
      Sample sp;
      load_sample(pts, n, nb, mb, sp);
      if (!sp.inside) v.w = 100.f;
    }
    reinterpret_cast<float4*>(raw)[n] = v;
  }
}

// GO[n] = g_raw[n] with component 3 zeroed outside the bound; GH[n][k] = (sum_o GO[n][o] Wo[o][k]) * [H[n][k] > 0]
__global__ void k_imap_out_bwd(pn_points pts, Bound6 nb, Bound6 mb, int apply_mask, const float* __restrict__ g_raw,
                               const float* __restrict__ Hl, int hidden, const float* __restrict__ Wo,
                               float* __restrict__ GO, float* __restrict__ GA) {
  const int64_t n = (int64_t)blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5);
  if (n >= pts.N) return;
  const int lane = threadIdx.x & 31;
  float4 g = reinterpret_cast<const float4*>(g_raw)[n];
  if (apply_mask) {
    Sample sp;
    load_sample(pts, n, nb, mb, sp);
    if (!sp.inside) g.w = 0.f;
  }
  if (lane == 0) reinterpret_cast<float4*>(GO)[n] = g;
  for (int k = lane; k < hidden; k += 32) {
    const float v = g.x * Wo[k] + g.y * Wo[hidden + k] + g.z * Wo[2 * hidden + k] + g.w * Wo[3 * hidden + k];
    GA[n * hidden + k] = Hl[n * hidden + k] > 0.f ? v : 0.f;
  }
}

// GE (N x 96, in place) *= cos(arg);  g_pts[n] += GE[n] . B^T
__global__ void k_imap_embed_bwd(pn_points pts, Bound6 nb, Bound6 mb, const float* __restrict__ Bm, float* __restrict__ GE,
                                 float* __restrict__ g_pts) {
  __shared__ float Bs[3 * PN_EMBED];
  for (int i = threadIdx.x; i < 3 * PN_EMBED; i += blockDim.x) Bs[i] = Bm[i];
  __syncthreads();
  const int64_t n = (int64_t)blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5);
  if (n >= pts.N) return;
  const int lane = threadIdx.x & 31;
  Sample sp;
  load_sample(pts, n, nb, mb, sp);
  float gp[3] = {0.f, 0.f, 0.f};
  for (int k = lane; k < 96; k += 32) {
    float ga = 0.f;
    if (k < PN_EMBED) {
      const float arg = fmaf(sp.pf[2], Bs[2 * PN_EMBED + k], fmaf(sp.pf[1], Bs[PN_EMBED + k], sp.pf[0] * Bs[k]));
      ga = GE[n * 96 + k] * fourier_cos(arg);
      gp[0] = fmaf(Bs[k], ga, gp[0]); gp[1] = fmaf(Bs[PN_EMBED + k], ga, gp[1]); gp[2] = fmaf(Bs[2 * PN_EMBED + k], ga, gp[2]);
    }
    GE[n * 96 + k] = ga;
  }
  if (g_pts) {
#pragma unroll
    for (int a = 0; a < 3; ++a) gp[a] = warp_sum(gp[a]);
    if (lane == 0) { g_pts[3 * n] += gp[0]; g_pts[3 * n + 1] += gp[1]; g_pts[3 * n + 2] += gp[2]; }
  }
}

// out[j] += sum_n X[n*ld + j]
__global__ void __launch_bounds__(256) k_colsum(const float* __restrict__ X, int64_t N, int ld, int ncols, float* __restrict__ out) {
  const int j = blockIdx.x * 32 + (threadIdx.x & 31);
  const int grp = threadIdx.x >> 5;
  __shared__ float red[8][32];
  float s = 0.f;
  if (j < ncols)
    for (int64_t n = (int64_t)blockIdx.y * 8 + grp; n < N; n += (int64_t)gridDim.y * 8) s += X[n * ld + j];
  red[grp][threadIdx.x & 31] = s;
  __syncthreads();
  if (threadIdx.x < 32 && j < ncols) {
    float t = 0.f;
    for (int g = 0; g < 8; ++g) t += red[g][threadIdx.x];
    atomicAdd(out + j, t);
  }
}

bool check_imap(const pn_imap_mlp* w, const char* fn) {
  if (!w || !w->B || !w->Wo || !w->bo || w->n_blocks < 1 || w->n_blocks > PN_IMAP_MAX_BLOCKS || w->hidden < 4 ||
      (w->hidden % 4) != 0) {
    set_error("%s: bad iMAP decoder description", fn);
    return false;
  }
  for (int i = 0; i < w->n_blocks; ++i)
    if (!w->W[i] || !w->b[i]) { set_error("%s: null parameter pointer", fn); return false; }
  return true;
}

}  // namespace
}  // namespace pn

using namespace pn;

extern "C" int pn_imap_mlp_fwd(const pn_points* pts, const pn_imap_mlp* w, const double* mask_bound, int apply_mask,
                               float* raw, float* E, float* H, float* P32, void* stream) {
  if (!check_imap(w, "pn_imap_mlp_fwd")) return 1;
  if (!pts || !raw || !E || !H || (apply_mask && !mask_bound)) { set_error("pn_imap_mlp_fwd: null argument"); return 1; }
  const int64_t N = pts->N;
  if (N == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const Bound6 mb = make_bound(mask_bound);
  const int hid = w->hidden;
  const unsigned wg = (unsigned)((N + 7) / 8);
  k_imap_embed<<<wg, 256, 0, st>>>(*pts, mb, mb, w->B, E, P32);
  if (launch_status("k_imap_embed")) return 1;
  for (int l = 0; l < w->n_blocks; ++l) {
    const float* X = l == 0 ? E : H + (int64_t)(l - 1) * N * hid;
    const int K = l == 0 ? PN_EMBED : hid;
    const int ldx = l == 0 ? 96 : hid;
    // Y = relu(X . W^T + b): B(k,n) = W[n*K + k]
    if (gemm<EP_BIAS_RELU>(X, ldx, 1, w->W[l], 1, K, H + (int64_t)l * N * hid, hid, N, hid, K, w->b[l], nullptr, 1, st)) return 1;
  }
  k_imap_out<<<wg, 256, 0, st>>>(*pts, mb, mb, apply_mask, H + (int64_t)(w->n_blocks - 1) * N * hid, hid, w->Wo, w->bo, raw);
  return launch_status("k_imap_out");
}

extern "C" int pn_imap_mlp_bwd(const pn_points* pts, const pn_imap_mlp* w, const double* mask_bound, int apply_mask,
                               const float* g_raw, const float* E, const float* H, const float* P32, float* GA, float* GB,
                               float* GO, float* g_pts, const pn_imap_mlp_grad* g, void* stream) {
  if (!check_imap(w, "pn_imap_mlp_bwd")) return 1;
  if (!pts || !g_raw || !E || !H || !GA || !GB || !GO || (apply_mask && !mask_bound)) {
    set_error("pn_imap_mlp_bwd: null argument");
    return 1;
  }
  if (g && g->B && !P32) { set_error("pn_imap_mlp_bwd: dB needs the P32 stash"); return 1; }
  const int64_t N = pts->N;
  if (N == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const Bound6 mb = make_bound(mask_bound);
  const int hid = w->hidden, nb = w->n_blocks;
  const unsigned wg = (unsigned)((N + 7) / 8);
  const int splitk = (int)((N + 4095) / 4096 < 1 ? 1 : ((N + 4095) / 4096 > 64 ? 64 : (N + 4095) / 4096));
  const float* Hlast = H + (int64_t)(nb - 1) * N * hid;
  float* cur = GA;   // gradient at the pre-activations of block l (N x hid)
  float* nxt = GB;
  k_imap_out_bwd<<<wg, 256, 0, st>>>(*pts, mb, mb, apply_mask, g_raw, Hlast, hid, w->Wo, GO, cur);
  if (launch_status("k_imap_out_bwd")) return 1;
  if (g && g->Wo) {  // dWo (4 x hid) = GO^T . H_last
    if (hid <= 256) {
      k_thin_wgrad<4><<<4 * sm_count(), 256, 0, st>>>(GO, 1, 4, Hlast, hid, N, hid, g->Wo, hid);
      if (launch_status("k_thin_wgrad")) return 1;
    } else if (sgemm<EP_ATOMIC>(GO, 1, 4, Hlast, hid, 1, g->Wo, hid, 4, hid, N, nullptr, nullptr, splitk, st)) return 1;
  }
  if (g && g->bo) { k_colsum<<<dim3(1, 64), 256, 0, st>>>(GO, N, 4, 4, g->bo); if (launch_status("k_colsum")) return 1; }
  for (int l = nb - 1; l >= 0; --l) {
    const float* X = l == 0 ? E : H + (int64_t)(l - 1) * N * hid;
    const int K = l == 0 ? PN_EMBED : hid;
    const int ldx = l == 0 ? 96 : hid;
    bool bias_done = false;
    if (g && g->W[l]) {  // dW_l (hid x K) = cur^T . X : A(m,k) = cur[k*hid + m], B(k,n) = X[k*ldx + n]; db_l rides along
      const int rc = tc_gemm(EP_ATOMIC, cur, 1, hid, X, ldx, 1, g->W[l], K, hid, K, N, nullptr, nullptr, splitk, g->b[l], st);
      if (rc > 0) return 1;
      if (rc == 0) bias_done = g->b[l] != nullptr;
      else if (sgemm<EP_ATOMIC>(cur, 1, hid, X, ldx, 1, g->W[l], K, hid, K, N, nullptr, nullptr, splitk, st)) return 1;
    }
    if (g && g->b[l] && !bias_done) { k_colsum<<<dim3((hid + 31) / 32, 32), 256, 0, st>>>(cur, N, hid, hid, g->b[l]); if (launch_status("k_colsum")) return 1; }
    if (l > 0) {  // next = (cur . W_l) masked by H_{l-1} > 0 : B(k,n) = W[k*K + n]
      if (gemm<EP_MASK>(cur, hid, 1, w->W[l], K, 1, nxt, hid, N, hid, hid, nullptr, X, 1, st)) return 1;
      float* t = cur; cur = nxt; nxt = t;
    } else if (g_pts || (g && g->B)) {
      // GE (N x 96) = cur . W_0 (hid x 93); then * cos(arg), dp, dB
      float* GE = nxt;  // reuse as N x 96 (hid >= 96 is not required: buffers are sized max(hid,96))
      if (gemm<EP_STORE>(cur, hid, 1, w->W[0], PN_EMBED, 1, GE, 96, N, PN_EMBED, hid, nullptr, nullptr, 1, st)) return 1;
      k_imap_embed_bwd<<<wg, 256, 0, st>>>(*pts, mb, mb, w->B, GE, g_pts);
      if (launch_status("k_imap_embed_bwd")) return 1;
      if (g && g->B) {  // dB (3 x 93) = P^T . GE : A(r,n) = P32[r*N + n]
        k_thin_wgrad<3><<<4 * sm_count(), 256, 0, st>>>(P32, N, 1, GE, 96, N, PN_EMBED, g->B, PN_EMBED);
        if (launch_status("k_thin_wgrad")) return 1;
      }
    }
  }
  return 0;
}
