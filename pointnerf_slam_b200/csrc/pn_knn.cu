// k-nearest neural-point feature aggregation (BASELINE.json config 4, "pointNeRF_slam config").
//
// BUILDER-DEFINED SEMANTICS, NOT REFERENCE PARITY: the reference tree has no 3-D neural-point aggregation at all
// (the nearest code is the 2-D cKDTree radius search of src/frame.py:362-366 and
// src/search_points.py:122,223,445).  The specification implemented here is SURVEY 8c's: Point-NeRF style, the K = 8
// nearest neural points within a radius, inverse-squared-distance weights, blended 32-channel feature.  The oracle is
// oracle/knn_oracle.py (indices pinned against scipy.spatial.cKDTree.query).
//
//   index   uniform cell lattice over the scene bound, cell edge h >= radius; points sorted by cell
//           (histogram -> one-CTA scan -> scatter); a query visits the 3 x 3 rows of <= 3 x-adjacent cells
//           around its own cell: nine contiguous candidate ranges of the sorted array;
//   query   thread per sample; d2 = ((dx*dx) + (dy*dy)) + (dz*dz) in float32 with explicitly rounded
//           operations (no FMA contraction), candidates with d2 <= r2; the K smallest 64-bit keys
//           (d2 bits << 32 | point index) -- a total order, so the result does not depend on the order in which
//           the candidates are met: the indices are bit-exact against the oracle and reproducible;
//   blend   w_k = 1 / (d2_k + eps), f = (sum_k w_k F[i_k]) / (sum_k w_k); 8 lanes per sample, one 128-bit load
//           per lane and neighbour of the [P][32] feature rows (the 128-byte row layout of the voxel grids);
//   VJP     g_F[i_k] += (w_k / W) g   (red.global.add.v4.f32 on the rows), g_p = sum_k (g.F_k - g.f) / W * (-2 w_k^2)(p - x_k)
//           (lane k of a sample's 8 lanes owns neighbour k; 8-lane shuffle reductions).
#include "pn_common.cuh"

namespace pn {
namespace {

constexpr int K = PN_KNN_K;
static_assert(K == 8, "eight lanes per sample own the eight neighbours");

struct KnnCells {
  float lo[3];
  float inv_h;
  int nx, ny, nz;
};

inline KnnCells make_cells(const pn_knn_index* ix) {
  KnnCells c;
  for (int a = 0; a < 3; ++a) c.lo[a] = ix->lo[a];
  c.inv_h = ix->inv_h; c.nx = ix->nx; c.ny = ix->ny; c.nz = ix->nz;
  return c;
}

// floor((x - lo) * inv_h) in float32, as the oracle forms it (two rounded operations)
__device__ __forceinline__ int cell_raw(float x, float lo, float inv_h) {
  const float t = floorf(__fmul_rn(__fsub_rn(x, lo), inv_h));
  return (int)fminf(fmaxf(t, -2.0f), 1.0e9f);   // keeps the int conversion defined; NaN -> -2 (no cell)
}

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// ---------------------------------------------------------------------------------------------
// index build
// ---------------------------------------------------------------------------------------------
__global__ void k_knn_count(const float* __restrict__ xyz, int P, KnnCells c, int32_t* __restrict__ counts, int32_t* __restrict__ cell_of,
                            int32_t* __restrict__ rank) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P) return;
  const int cx = clampi(cell_raw(xyz[3 * i], c.lo[0], c.inv_h), 0, c.nx - 1);
  const int cy = clampi(cell_raw(xyz[3 * i + 1], c.lo[1], c.inv_h), 0, c.ny - 1);
  const int cz = clampi(cell_raw(xyz[3 * i + 2], c.lo[2], c.inv_h), 0, c.nz - 1);
  const int cell = (cz * c.ny + cy) * c.nx + cx;
  cell_of[i] = cell;
  rank[i] = atomicAdd(&counts[cell], 1);
}

// exclusive scan of counts[0..n) -> start[0..n], one CTA of 1024 threads, a contiguous chunk per thread
__global__ void __launch_bounds__(1024, 1) k_knn_scan(const int32_t* __restrict__ counts, int n, int32_t* __restrict__ start) {
  __shared__ int32_t part[1024];
  const int t = threadIdx.x;
  const int chunk = (n + 1023) / 1024;
  const int b = t * chunk, e = min(b + chunk, n);
  int32_t s = 0;
  for (int i = b; i < e; ++i) s += counts[i];
  part[t] = s;
  __syncthreads();
  for (int o = 1; o < 1024; o <<= 1) {   // Hillis-Steele inclusive scan
    const int32_t v = t >= o ? part[t - o] : 0;
    __syncthreads();
    part[t] += v;
    __syncthreads();
  }
  int32_t run = t > 0 ? part[t - 1] : 0;
  for (int i = b; i < e; ++i) { start[i] = run; run += counts[i]; }
  if (t == 1023) start[n] = part[1023];
}

__global__ void k_knn_scatter(const float* __restrict__ xyz, int P, const int32_t* __restrict__ start, const int32_t* __restrict__ cell_of,
                              const int32_t* __restrict__ rank, float4* __restrict__ sorted) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P) return;
  sorted[start[cell_of[i]] + rank[i]] = make_float4(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], __int_as_float(i));
}

// ---------------------------------------------------------------------------------------------
// sample point in float32 (p.float() of the reference's float64 o + d*z, Renderer.py:177-179)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void load_pf(const pn_points& ps, int64_t n, float (&p)[3]) {
  if (ps.pts32) {
#pragma unroll
    for (int a = 0; a < 3; ++a) p[a] = ps.pts32[3 * n + a];
  } else if (ps.pts64) {
#pragma unroll
    for (int a = 0; a < 3; ++a) p[a] = (float)ps.pts64[3 * n + a];
  } else {
    const int64_t r = n / ps.S;
    const double z = ps.z[n];
#pragma unroll
    for (int a = 0; a < 3; ++a) p[a] = (float)__dadd_rn((double)ps.rays_o[3 * r + a], __dmul_rn((double)ps.rays_d[3 * r + a], z));
  }
}

// ---------------------------------------------------------------------------------------------
// query (+ blend): CTA = 128 samples
// ---------------------------------------------------------------------------------------------
constexpr uint64_t kNoKey = ~0ull;

template <bool BLEND>
__global__ void __launch_bounds__(128) k_knn_fwd(const pn_points ps, KnnCells c, const int32_t* __restrict__ start,
                                                 const float4* __restrict__ sorted, float r2, float eps, const float* __restrict__ feat,
                                                 int32_t* __restrict__ idx_out, float* __restrict__ d2_out, float* __restrict__ out) {
  const int64_t n = (int64_t)blockIdx.x * 128 + threadIdx.x;
  const int lane = threadIdx.x & 31;
  const bool valid = n < ps.N;
  uint64_t key[K];
#pragma unroll
  for (int k = 0; k < K; ++k) key[k] = kNoKey;
  float p[3] = {0.f, 0.f, 0.f};
  if (valid) {
    load_pf(ps, n, p);
    const int cx = cell_raw(p[0], c.lo[0], c.inv_h), cy = cell_raw(p[1], c.lo[1], c.inv_h), cz = cell_raw(p[2], c.lo[2], c.inv_h);
    const int x0 = max(cx - 1, 0), x1 = min(cx + 1, c.nx - 1);
    if (x0 <= x1) {
      for (int z = max(cz - 1, 0); z <= min(cz + 1, c.nz - 1); ++z) {
        for (int y = max(cy - 1, 0); y <= min(cy + 1, c.ny - 1); ++y) {
          const int row = (z * c.ny + y) * c.nx;
          const int b = __ldg(start + row + x0), e = __ldg(start + row + x1 + 1);
          for (int j = b; j < e; ++j) {
            const float4 q = __ldg(sorted + j);
            const float dx = __fsub_rn(p[0], q.x), dy = __fsub_rn(p[1], q.y), dz = __fsub_rn(p[2], q.z);
            const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
            if (d2 <= r2) {
              const uint64_t kk = ((uint64_t)__float_as_uint(d2) << 32) | (uint64_t)__float_as_uint(q.w);
              if (kk < key[K - 1]) {
                key[K - 1] = kk;
#pragma unroll
                for (int i = K - 1; i > 0; --i) {
                  const uint64_t a = key[i - 1], bb = key[i];
                  key[i - 1] = a < bb ? a : bb;
                  key[i] = a < bb ? bb : a;
                }
              }
            }
          }
        }
      }
    }
  }
  int32_t id[K];
  float d2[K];
#pragma unroll
  for (int k = 0; k < K; ++k) {
    const bool have = key[k] != kNoKey;
    id[k] = have ? (int32_t)(uint32_t)key[k] : -1;
    d2[k] = have ? __uint_as_float((uint32_t)(key[k] >> 32)) : 0.f;
  }
  if (valid) {
    int4* io = reinterpret_cast<int4*>(idx_out + n * K);
    io[0] = make_int4(id[0], id[1], id[2], id[3]);
    io[1] = make_int4(id[4], id[5], id[6], id[7]);
    if (d2_out) {
      float4* dd = reinterpret_cast<float4*>(d2_out + n * K);
      dd[0] = make_float4(d2[0], d2[1], d2[2], d2[3]);
      dd[1] = make_float4(d2[4], d2[5], d2[6], d2[7]);
    }
  }
  if (!BLEND) return;
  // ---- blend: 8 lanes per sample (lane & 7 = channel quad), 4 samples of the warp per iteration
  float w[K];
  float W = 0.f;
#pragma unroll
  for (int k = 0; k < K; ++k) {
    w[k] = id[k] >= 0 ? __fdiv_rn(1.0f, __fadd_rn(d2[k], eps)) : 0.f;
    W = __fadd_rn(W, w[k]);
  }
  const int q = lane & 7, sub = lane >> 3;
  const int64_t n_warp = n - lane;   // first sample of this warp
  const float4* feat4 = reinterpret_cast<const float4*>(feat);
#pragma unroll 1
  for (int it = 0; it < 8; ++it) {
    const int src = 4 * it + sub;
    const float Ws = __shfl_sync(kFull, W, src);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 f[K];
    float wk[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const int ik = __shfl_sync(kFull, id[k], src);
      wk[k] = __shfl_sync(kFull, w[k], src);
      f[k] = ik >= 0 ? __ldg(feat4 + (int64_t)ik * 8 + q) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int k = 0; k < K; ++k) {
      acc.x = fmaf(wk[k], f[k].x, acc.x); acc.y = fmaf(wk[k], f[k].y, acc.y);
      acc.z = fmaf(wk[k], f[k].z, acc.z); acc.w = fmaf(wk[k], f[k].w, acc.w);
    }
    const float inv = Ws > 0.f ? __fdiv_rn(1.0f, Ws) : 0.f;
    if (n_warp + src < ps.N)
      reinterpret_cast<float4*>(out)[(n_warp + src) * 8 + q] = make_float4(acc.x * inv, acc.y * inv, acc.z * inv, acc.w * inv);
  }
}

// ---------------------------------------------------------------------------------------------
// VJP of the blend
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_knn_bwd(const pn_points ps, const int32_t* __restrict__ idx, const float* __restrict__ d2in,
                                                 float eps, const float* __restrict__ feat, const float* __restrict__ xyz,
                                                 const float* __restrict__ g_out, float* __restrict__ g_feat, float* __restrict__ g_pts,
                                                 int accumulate_pts) {
  const int64_t n = (int64_t)blockIdx.x * 128 + threadIdx.x;
  const int lane = threadIdx.x & 31;
  const bool valid = n < ps.N;
  int32_t id[K];
  float w[K];
  float p[3] = {0.f, 0.f, 0.f};
  float W = 0.f;
  if (valid) {
    const int4 i0 = reinterpret_cast<const int4*>(idx + n * K)[0], i1 = reinterpret_cast<const int4*>(idx + n * K)[1];
    const float4 a0 = reinterpret_cast<const float4*>(d2in + n * K)[0], a1 = reinterpret_cast<const float4*>(d2in + n * K)[1];
    id[0] = i0.x; id[1] = i0.y; id[2] = i0.z; id[3] = i0.w; id[4] = i1.x; id[5] = i1.y; id[6] = i1.z; id[7] = i1.w;
    const float d2[K] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
    for (int k = 0; k < K; ++k) {
      w[k] = id[k] >= 0 ? __fdiv_rn(1.0f, __fadd_rn(d2[k], eps)) : 0.f;
      W = __fadd_rn(W, w[k]);
    }
    if (g_pts) load_pf(ps, n, p);
  } else {
#pragma unroll
    for (int k = 0; k < K; ++k) { id[k] = -1; w[k] = 0.f; }
  }
  const int q = lane & 7, sub = lane >> 3;
  const int64_t n_warp = n - lane;
  const float4* feat4 = reinterpret_cast<const float4*>(feat);
#pragma unroll 1
  for (int it = 0; it < 8; ++it) {
    const int src = 4 * it + sub;
    const int64_t ns = n_warp + src;
    const float Ws = __shfl_sync(kFull, W, src);
    const float inv = Ws > 0.f ? __fdiv_rn(1.0f, Ws) : 0.f;
    const float4 g = (ns < ps.N) ? __ldg(reinterpret_cast<const float4*>(g_out) + ns * 8 + q) : make_float4(0.f, 0.f, 0.f, 0.f);
    float gF[K];
    float gf = 0.f;       // g . f
    int my_id = -1;       // neighbour q of this sample, for the point gradient
    float my_w = 0.f;
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const int ik = __shfl_sync(kFull, id[k], src);
      const float wn = __shfl_sync(kFull, w[k], src) * inv;   // normalised weight
      float dot = 0.f;
      if (ik >= 0) {
        if (g_pts) {
          const float4 f = __ldg(feat4 + (int64_t)ik * 8 + q);
          dot = fmaf(g.x, f.x, fmaf(g.y, f.y, fmaf(g.z, f.z, g.w * f.w)));
        }
        if (g_feat) red_add_v4(g_feat + (int64_t)ik * 32 + 4 * q, make_float4(wn * g.x, wn * g.y, wn * g.z, wn * g.w));
      }
      if (g_pts) {
        dot += __shfl_xor_sync(kFull, dot, 1);
        dot += __shfl_xor_sync(kFull, dot, 2);
        dot += __shfl_xor_sync(kFull, dot, 4);
        gF[k] = dot;
        gf = fmaf(wn, dot, gf);
        if (k == q) { my_id = ik; my_w = wn; }
      }
    }
    if (g_pts) {
      const float px = __shfl_sync(kFull, p[0], src), py = __shfl_sync(kFull, p[1], src), pz = __shfl_sync(kFull, p[2], src);
      float mine = 0.f;
#pragma unroll
      for (int k = 0; k < K; ++k) mine = (k == q) ? gF[k] : mine;
      float cx = 0.f, cy = 0.f, cz = 0.f;
      if (my_id >= 0) {
        // d f / d w_k = (F_k - f) / W;  d w_k / d p = -2 w_k^2 (p - x_k);  my_w = w_k / W
        const float s = -2.0f * (mine - gf) * my_w * (my_w * Ws);
        cx = s * (px - __ldg(xyz + 3 * (int64_t)my_id));
        cy = s * (py - __ldg(xyz + 3 * (int64_t)my_id + 1));
        cz = s * (pz - __ldg(xyz + 3 * (int64_t)my_id + 2));
      }
#pragma unroll
      for (int o = 1; o < 8; o <<= 1) {
        cx += __shfl_xor_sync(kFull, cx, o); cy += __shfl_xor_sync(kFull, cy, o); cz += __shfl_xor_sync(kFull, cz, o);
      }
      if (q == 0 && ns < ps.N) {
        float* o = g_pts + 3 * ns;
        if (accumulate_pts) { o[0] += cx; o[1] += cy; o[2] += cz; }
        else { o[0] = cx; o[1] = cy; o[2] = cz; }
      }
    }
  }
}

bool points_ok(const pn_points* pts) {
  if (!pts || pts->N < 0) return false;
  if (pts->N == 0) return true;
  const int nsrc = (pts->pts32 != nullptr) + (pts->pts64 != nullptr) + (pts->rays_o != nullptr);
  if (nsrc != 1) return false;
  if (pts->rays_o && (!pts->rays_d || !pts->z || pts->S <= 0)) return false;
  return true;
}

bool index_ok(const pn_knn_index* ix) {
  return ix && ix->start && ix->sorted && ix->nx > 0 && ix->ny > 0 && ix->nz > 0 && ix->inv_h > 0.f && ix->P >= 0;
}

}  // namespace

extern "C" int pn_knn_build(const float* xyz, int P, const pn_knn_index* index, int32_t* scratch, void* stream) {
  if (!index_ok(index) || P != index->P || (P > 0 && (!xyz || !scratch))) { set_error("pn_knn_build: null pointer or bad index descriptor"); return 1; }
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t ncell = (int64_t)index->nx * index->ny * index->nz;
  if (ncell > (1ll << 30)) { set_error("pn_knn_build: %lld cells is more than the 2^30 the index supports", (long long)ncell); return 1; }
  int32_t* start = const_cast<int32_t*>(index->start);
  // scratch = cell_of [P] | rank [P] | counts [ncell]
  int32_t* cell_of = scratch;
  int32_t* rank = scratch + P;
  int32_t* counts = scratch + 2 * (int64_t)P;
  if (cudaMemsetAsync(counts, 0, sizeof(int32_t) * (size_t)ncell, st) != cudaSuccess) { set_error("pn_knn_build: memset failed"); return 1; }
  const KnnCells c = make_cells(index);
  if (P > 0) {
    k_knn_count<<<(unsigned)((P + 255) / 256), 256, 0, st>>>(xyz, P, c, counts, cell_of, rank);
    if (launch_status("k_knn_count")) return 1;
  }
  k_knn_scan<<<1, 1024, 0, st>>>(counts, (int)ncell, start);
  if (launch_status("k_knn_scan")) return 1;
  if (P > 0) {
    k_knn_scatter<<<(unsigned)((P + 255) / 256), 256, 0, st>>>(xyz, P, start, cell_of, rank, const_cast<float4*>(reinterpret_cast<const float4*>(index->sorted)));
    if (launch_status("k_knn_scatter")) return 1;
  }
  return 0;
}

extern "C" int pn_knn_query(const pn_points* pts, const pn_knn_index* index, float radius, int32_t* idx, float* d2, void* stream) {
  if (!points_ok(pts) || !index_ok(index) || !(radius > 0.f)) { set_error("pn_knn_query: bad points / index / radius"); return 1; }
  if (pts->N == 0) return 0;
  if (!idx) { set_error("pn_knn_query: null output"); return 1; }
  if (radius * index->inv_h > 1.0f) { set_error("pn_knn_query: radius %g exceeds the cell edge %g of the index", radius, 1.0 / index->inv_h); return 1; }
  k_knn_fwd<false><<<(unsigned)((pts->N + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
      *pts, make_cells(index), index->start, reinterpret_cast<const float4*>(index->sorted), radius * radius, 0.f, nullptr, idx, d2, nullptr);
  return launch_status("k_knn_fwd");
}

extern "C" int pn_knn_aggregate_fwd(const pn_points* pts, const pn_knn_index* index, float radius, float eps, const float* feat,
                                    int32_t* idx, float* d2, float* out, void* stream) {
  if (!points_ok(pts) || !index_ok(index) || !(radius > 0.f) || !(eps > 0.f)) { set_error("pn_knn_aggregate_fwd: bad points / index / radius / eps"); return 1; }
  if (pts->N == 0) return 0;
  if (!idx || !d2 || !out || !feat) { set_error("pn_knn_aggregate_fwd: null pointer"); return 1; }
  if (radius * index->inv_h > 1.0f) { set_error("pn_knn_aggregate_fwd: radius %g exceeds the cell edge %g of the index", radius, 1.0 / index->inv_h); return 1; }
  k_knn_fwd<true><<<(unsigned)((pts->N + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
      *pts, make_cells(index), index->start, reinterpret_cast<const float4*>(index->sorted), radius * radius, eps, feat, idx, d2, out);
  return launch_status("k_knn_fwd");
}

extern "C" int pn_knn_aggregate_bwd(const pn_points* pts, const int32_t* idx, const float* d2, float eps, const float* feat,
                                    const float* xyz, const float* g_out, float* g_feat, float* g_pts, int accumulate_pts,
                                    void* stream) {
  if (!points_ok(pts) || !(eps > 0.f)) { set_error("pn_knn_aggregate_bwd: bad points / eps"); return 1; }
  if (pts->N == 0 || (!g_feat && !g_pts)) return 0;
  if (!idx || !d2 || !g_out || (g_pts && (!feat || !xyz))) { set_error("pn_knn_aggregate_bwd: null pointer"); return 1; }
  k_knn_bwd<<<(unsigned)((pts->N + 127) / 128), 128, 0, (cudaStream_t)stream>>>(*pts, idx, d2, eps, feat, xyz, g_out, g_feat, g_pts,
                                                                                accumulate_pts);
  return launch_status("k_knn_bwd");
}

}  // namespace pn
