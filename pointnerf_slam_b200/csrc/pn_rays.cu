// Pose, ray generation, sample placement and compositing kernels.
//
// Each kernel replaces a chain of small ATen launches of the reference and
// keeps its dtype conventions (float32 rays, float64 bound / z-values / depth):
//   pn_camera_from_tensor_*   src/common.py:137-176
//   pn_sample_rays_fwd        src/common.py:74-134 (minus the randint)
//   pn_image_rays_fwd         src/common.py:248-266
//   pn_rays_bwd               autograd of the two above w.r.t. c2w
//   pn_ray_zvals              src/utils/Renderer.py:90-175
//   pn_importance_zvals       src/common.py:19-63 + Renderer.py:187-191
//   pn_regulation_points      src/utils/Renderer.py:280-298
//   pn_composite_*            src/common.py:204-245 (warp per ray, shuffle scans)
//   pn_points_to_rays_bwd     autograd of Renderer.py:177-178
#include "pn_common.cuh"

namespace pn {
namespace {

// ------------------------------------------------------------------ pose
__global__ void k_cam_fwd(const float* __restrict__ cam, int B, float* __restrict__ c2w) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const float* q = cam + 7 * b;
  const float qr = q[0], qi = q[1], qj = q[2], qk = q[3];
  const float n = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(qr, qr), __fmul_rn(qi, qi)), __fmul_rn(qj, qj)), __fmul_rn(qk, qk));
  const float s = __fdiv_rn(2.0f, n);
  float* o = c2w + 12 * b;
#define PN_SQ(a) __fmul_rn(a, a)
#define PN_M(a, b) __fmul_rn(a, b)
  o[0] = __fsub_rn(1.f, PN_M(s, __fadd_rn(PN_SQ(qj), PN_SQ(qk))));
  o[1] = PN_M(s, __fsub_rn(PN_M(qi, qj), PN_M(qk, qr)));
  o[2] = PN_M(s, __fadd_rn(PN_M(qi, qk), PN_M(qj, qr)));
  o[3] = q[4];
  o[4] = PN_M(s, __fadd_rn(PN_M(qi, qj), PN_M(qk, qr)));
  o[5] = __fsub_rn(1.f, PN_M(s, __fadd_rn(PN_SQ(qi), PN_SQ(qk))));
  o[6] = PN_M(s, __fsub_rn(PN_M(qj, qk), PN_M(qi, qr)));
  o[7] = q[5];
  o[8] = PN_M(s, __fsub_rn(PN_M(qi, qk), PN_M(qj, qr)));
  o[9] = PN_M(s, __fadd_rn(PN_M(qj, qk), PN_M(qi, qr)));
  o[10] = __fsub_rn(1.f, PN_M(s, __fadd_rn(PN_SQ(qi), PN_SQ(qj))));
  o[11] = q[6];
#undef PN_SQ
#undef PN_M
}

__global__ void k_cam_bwd(const float* __restrict__ cam, const float* __restrict__ g, int B, float* __restrict__ gc) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const float* q = cam + 7 * b;
  const float* G = g + 12 * b;  // G[a*4+c]
  const float qr = q[0], qi = q[1], qj = q[2], qk = q[3];
  const float n = qr * qr + qi * qi + qj * qj + qk * qk;
  const float s = 2.0f / n;
  const float g00 = G[0], g01 = G[1], g02 = G[2], g10 = G[4], g11 = G[5], g12 = G[6], g20 = G[8], g21 = G[9], g22 = G[10];
  // R = I + s*M(q);  GM = <G, M>
  const float gm = -g00 * (qj * qj + qk * qk) + g01 * (qi * qj - qk * qr) + g02 * (qi * qk + qj * qr) +
                   g10 * (qi * qj + qk * qr) - g11 * (qi * qi + qk * qk) + g12 * (qj * qk - qi * qr) +
                   g20 * (qi * qk - qj * qr) + g21 * (qj * qk + qi * qr) - g22 * (qi * qi + qj * qj);
  const float dr = -g01 * qk + g02 * qj + g10 * qk - g12 * qi - g20 * qj + g21 * qi;
  const float di = g01 * qj + g02 * qk + g10 * qj - 2.f * g11 * qi - g12 * qr + g20 * qk + g21 * qr - 2.f * g22 * qi;
  const float dj = -2.f * g00 * qj + g01 * qi + g02 * qr + g10 * qi + g12 * qk - g20 * qr + g21 * qk - 2.f * g22 * qj;
  const float dk = -2.f * g00 * qk - g01 * qr + g02 * qi + g10 * qr - 2.f * g11 * qk + g12 * qj + g20 * qi + g21 * qj;
  const float ds = -s * s;  // d s / d q_m = -s^2 q_m
  float* o = gc + 7 * b;
  o[0] = s * dr + gm * ds * qr;
  o[1] = s * di + gm * ds * qi;
  o[2] = s * dj + gm * ds * qj;
  o[3] = s * dk + gm * ds * qk;
  o[4] = G[3]; o[5] = G[7]; o[6] = G[11];
}

// ------------------------------------------------------------------ rays
__device__ __forceinline__ void pixel_dir(float i, float j, float fx, float fy, float cx, float cy, float (&d)[3]) {
  d[0] = __fdiv_rn(__fsub_rn(i, cx), fx);
  d[1] = -__fdiv_rn(__fsub_rn(j, cy), fy);
  d[2] = -1.0f;
}

__device__ __forceinline__ void rotate_dir(const float (&d)[3], const float* __restrict__ c2w, int ld, float* __restrict__ rd,
                                           float* __restrict__ ro) {
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const float* r = c2w + a * ld;
    rd[a] = __fadd_rn(__fadd_rn(__fmul_rn(d[0], r[0]), __fmul_rn(d[1], r[1])), __fmul_rn(d[2], r[2]));
    ro[a] = r[3];
  }
}

template <typename CT>
__global__ void k_sample_rays(const int64_t* __restrict__ idx, int n, int H0, int W0, int Wc, int W, float fx, float fy,
                              float cx, float cy, const float* __restrict__ c2w, int ld, const float* __restrict__ depth,
                              const CT* __restrict__ color, float* __restrict__ ro, float* __restrict__ rd,
                              float* __restrict__ dout, CT* __restrict__ cout) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const int64_t f = idx[t];
  const int r = (int)(f / Wc), c = (int)(f % Wc);
  const int64_t pix = (int64_t)(H0 + r) * W + (W0 + c);
  if (dout) dout[t] = depth[pix];
  if (cout) { cout[3 * t] = color[3 * pix]; cout[3 * t + 1] = color[3 * pix + 1]; cout[3 * t + 2] = color[3 * pix + 2]; }
  float d[3];
  pixel_dir((float)(W0 + c), (float)(H0 + r), fx, fy, cx, cy, d);
  rotate_dir(d, c2w, ld, rd + 3 * t, ro + 3 * t);
}

__global__ void k_image_rays(int H, int W, float fx, float fy, float cx, float cy, const float* __restrict__ c2w, int ld,
                             float* __restrict__ ro, float* __restrict__ rd) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (int64_t)H * W) return;
  float d[3];
  pixel_dir((float)(t % W), (float)(t / W), fx, fy, cx, cy, d);
  rotate_dir(d, c2w, ld, rd + 3 * t, ro + 3 * t);
}

__global__ void __launch_bounds__(256) k_rays_bwd(const int64_t* __restrict__ idx, int n, int H0, int W0, int Wc, float fx,
                                                 float fy, float cx, float cy, const float* __restrict__ go,
                                                 const float* __restrict__ gd, float* __restrict__ gc2w) {
  __shared__ float red[8][12];
  float acc[12];
#pragma unroll
  for (int k = 0; k < 12; ++k) acc[k] = 0.f;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) {
    const int64_t f = idx ? idx[t] : (int64_t)t;
    const int r = (int)(f / Wc), c = (int)(f % Wc);
    float d[3];
    pixel_dir((float)(W0 + c), (float)(H0 + r), fx, fy, cx, cy, d);
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const float g = gd ? gd[3 * t + a] : 0.f;
      acc[4 * a] = fmaf(g, d[0], acc[4 * a]);
      acc[4 * a + 1] = fmaf(g, d[1], acc[4 * a + 1]);
      acc[4 * a + 2] = fmaf(g, d[2], acc[4 * a + 2]);
      acc[4 * a + 3] += go ? go[3 * t + a] : 0.f;
    }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 12; ++k) {
    const float s = warp_sum(acc[k]);
    if (lane == 0) red[warp][k] = s;
  }
  __syncthreads();
  if (threadIdx.x < 12) {
    float s = 0.f;
    for (int w = 0; w < 8; ++w) s += red[w][threadIdx.x];
    atomicAdd(gc2w + threadIdx.x, s);
  }
}

// ---- the same two kernels over F keyframes at once (blockIdx.y = keyframe): one launch for the Mapper's
//      per-keyframe loop (src/Mapper.py:558-605) instead of one per keyframe
template <typename CT>
__global__ void k_sample_rays_multi(const int64_t* __restrict__ idx, int n, int H0, int W0, int Wc, int W, float fx, float fy,
                                    float cx, float cy, const float* __restrict__ c2w, const float* const* __restrict__ depth,
                                    const CT* const* __restrict__ color, float* __restrict__ ro, float* __restrict__ rd,
                                    float* __restrict__ dout, CT* __restrict__ cout) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x, f_ = blockIdx.y;
  if (t >= n) return;
  const int64_t o = (int64_t)f_ * n + t;
  const int64_t f = idx[o];
  const int r = (int)(f / Wc), c = (int)(f % Wc);
  const int64_t pix = (int64_t)(H0 + r) * W + (W0 + c);
  const float* dp = depth[f_];
  const CT* cp = color[f_];
  dout[o] = dp[pix];
  cout[3 * o] = cp[3 * pix]; cout[3 * o + 1] = cp[3 * pix + 1]; cout[3 * o + 2] = cp[3 * pix + 2];
  float d[3];
  pixel_dir((float)(W0 + c), (float)(H0 + r), fx, fy, cx, cy, d);
  rotate_dir(d, c2w + 12 * f_, 4, rd + 3 * o, ro + 3 * o);
}

__global__ void __launch_bounds__(256) k_rays_bwd_multi(const int64_t* __restrict__ idx, int n, int H0, int W0, int Wc, float fx,
                                                       float fy, float cx, float cy, const float* __restrict__ go,
                                                       const float* __restrict__ gd, float* __restrict__ gc2w) {
  __shared__ float red[8][12];
  const int f_ = blockIdx.y;
  float acc[12];
#pragma unroll
  for (int k = 0; k < 12; ++k) acc[k] = 0.f;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) {
    const int64_t o = (int64_t)f_ * n + t;
    const int64_t f = idx[o];
    const int r = (int)(f / Wc), c = (int)(f % Wc);
    float d[3];
    pixel_dir((float)(W0 + c), (float)(H0 + r), fx, fy, cx, cy, d);
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const float g = gd ? gd[3 * o + a] : 0.f;
      acc[4 * a] = fmaf(g, d[0], acc[4 * a]);
      acc[4 * a + 1] = fmaf(g, d[1], acc[4 * a + 1]);
      acc[4 * a + 2] = fmaf(g, d[2], acc[4 * a + 2]);
      acc[4 * a + 3] += go ? go[3 * o + a] : 0.f;
    }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 12; ++k) {
    const float s = warp_sum(acc[k]);
    if (lane == 0) red[warp][k] = s;
  }
  __syncthreads();
  if (threadIdx.x < 12) {
    float s = 0.f;
    for (int w = 0; w < 8; ++w) s += red[w][threadIdx.x];
    atomicAdd(gc2w + 12 * f_ + threadIdx.x, s);
  }
}

// ------------------------------------------------------------------ z-values
__device__ __forceinline__ void insertion_sort(double* z, int n) {
  for (int i = 1; i < n; ++i) {
    const double v = z[i];
    int j = i - 1;
    while (j >= 0 && z[j] > v) { z[j + 1] = z[j]; --j; }
    z[j + 1] = v;
  }
}

__global__ void __launch_bounds__(128) k_ray_zvals(const float* __restrict__ ro, const float* __restrict__ rd,
                                                  const float* __restrict__ gt, const float* __restrict__ dmax_p, int64_t R,
                                                  Bound6 b, int ns, int nsurf, int lindisp, const float* __restrict__ tv,
                                                  const double* __restrict__ ts, const float* __restrict__ t_rand,
                                                  double* __restrict__ zout) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  double z[PN_MAX_SAMPLES];
  // box exit distance (no-grad block, Renderer.py:98-105)
  double far_bb = 0.0;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const double o = (double)ro[3 * r + a], d = (double)rd[3 * r + a];
    const double t0 = __ddiv_rn(__dsub_rn(b.v[2 * a], o), d), t1 = __ddiv_rn(__dsub_rn(b.v[2 * a + 1], o), d);
    const double tm = fmax(t0, t1);
    far_bb = (a == 0) ? tm : fmin(far_bb, tm);
  }
  far_bb = __dadd_rn(far_bb, 0.01);
  const bool has_gt = gt != nullptr;
  float g = 0.f, dmax = 0.f;
  double far = far_bb;
  if (has_gt) {
    g = gt[r];
    dmax = *dmax_p;
    const double cap = (double)__fmul_rn(dmax, 1.2f);
    far = fmin(fmax(far_bb, 0.0), cap);
  }
  const float near32 = has_gt ? __fmul_rn(g, 0.01f) : 0.01f;
  for (int k = 0; k < ns; ++k) {
    const float t = tv[k];
    const float omt = __fsub_rn(1.0f, t);
    if (!lindisp) {
      z[k] = __dadd_rn((double)__fmul_rn(near32, omt), __dmul_rn(far, (double)t));
    } else {
      const float inv_near = has_gt ? __fdiv_rn(1.0f, near32) : 100.0f;
      const double a = (double)__fmul_rn(inv_near, omt);
      const double c = __dmul_rn(__ddiv_rn(1.0, far), (double)t);
      z[k] = __ddiv_rn(1.0, __dadd_rn(a, c));
    }
  }
  if (t_rand) {  // perturb > 0 (Renderer.py:164-171)
    double prev = z[0];
    double lower = z[0];
    for (int k = 0; k < ns; ++k) {
      const double cur = z[k];
      const double upper = (k + 1 < ns) ? __dmul_rn(0.5, __dadd_rn(z[k + 1], cur)) : cur;
      if (k > 0) lower = __dmul_rn(0.5, __dadd_rn(cur, prev));
      prev = cur;
      z[k] = __dadd_rn(lower, __dmul_rn(__dsub_rn(upper, lower), (double)t_rand[r * ns + k]));
    }
  }
  int S = ns;
  if (has_gt && nsurf > 0) {
    if (g > 0.f) {
      const double lo = (double)__fmul_rn(0.95f, g), hi = (double)__fmul_rn(1.05f, g);
      for (int k = 0; k < nsurf; ++k) z[ns + k] = __dadd_rn(__dmul_rn(lo, __dsub_rn(1.0, ts[k])), __dmul_rn(hi, ts[k]));
    } else {
      for (int k = 0; k < nsurf; ++k)
        z[ns + k] = __dadd_rn(__dmul_rn(0.001, __dsub_rn(1.0, ts[k])), __dmul_rn((double)dmax, ts[k]));
    }
    S = ns + nsurf;
    // torch.sort of the concatenation (Renderer.py:157).  Both runs are normally ascending already
    // (t_vals / t_surface increase, far >= near), so a merge gives the sorted row in S steps; a
    // descending run (far < near: the ray leaves the box at once) falls back to the general sort.
    bool ascending = true;
    for (int k = 1; k < ns; ++k) ascending = ascending && (z[k - 1] <= z[k]);
    for (int k = ns + 1; k < S; ++k) ascending = ascending && (z[k - 1] <= z[k]);
    if (ascending) {
      int i = 0, j = ns;
      double a = z[0], c = z[ns];
      for (int k = 0; k < S; ++k) {
        const bool take_a = (j >= S) || (i < ns && a <= c);
        zout[r * S + k] = take_a ? a : c;
        if (take_a) { ++i; a = i < ns ? z[i] : a; } else { ++j; c = j < S ? z[j] : c; }
      }
      return;
    }
    insertion_sort(z, S);
  }
  for (int k = 0; k < S; ++k) zout[r * S + k] = z[k];
}

// Warp-per-ray form of k_ray_zvals for the common case (no perturbation): lanes compute the samples of both runs,
// and -- when both runs ascend -- every sample finds its place in the merged row by counting the samples of the
// other run that precede it (ties: the uniform run first, as in the sequential merge).  Same arithmetic, same
// values; a descending run is sorted by lane 0.  8 rays per CTA.
__global__ void __launch_bounds__(256) k_ray_zvals_warp(const float* __restrict__ ro, const float* __restrict__ rd,
                                                       const float* __restrict__ gt, const float* __restrict__ dmax_p, int64_t R,
                                                       Bound6 b, int ns, int nsurf, int lindisp, const float* __restrict__ tv,
                                                       const double* __restrict__ ts, double* __restrict__ zout) {
  __shared__ double zs[8][PN_MAX_SAMPLES];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t r = (int64_t)blockIdx.x * 8 + w;
  if (r >= R) return;
  double* z = zs[w];
  double far_bb = 0.0;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const double o = (double)ro[3 * r + a], d = (double)rd[3 * r + a];
    const double t0 = __ddiv_rn(__dsub_rn(b.v[2 * a], o), d), t1 = __ddiv_rn(__dsub_rn(b.v[2 * a + 1], o), d);
    const double tm = fmax(t0, t1);
    far_bb = (a == 0) ? tm : fmin(far_bb, tm);
  }
  far_bb = __dadd_rn(far_bb, 0.01);
  const bool has_gt = gt != nullptr;
  float g = 0.f, dmax = 0.f;
  double far = far_bb;
  if (has_gt) {
    g = gt[r];
    dmax = *dmax_p;
    far = fmin(fmax(far_bb, 0.0), (double)__fmul_rn(dmax, 1.2f));
  }
  const float near32 = has_gt ? __fmul_rn(g, 0.01f) : 0.01f;
  for (int k = lane; k < ns; k += 32) {
    const float t = tv[k];
    const float omt = __fsub_rn(1.0f, t);
    if (!lindisp) {
      z[k] = __dadd_rn((double)__fmul_rn(near32, omt), __dmul_rn(far, (double)t));
    } else {
      const float inv_near = has_gt ? __fdiv_rn(1.0f, near32) : 100.0f;
      const double a = (double)__fmul_rn(inv_near, omt);
      const double c = __dmul_rn(__ddiv_rn(1.0, far), (double)t);
      z[k] = __ddiv_rn(1.0, __dadd_rn(a, c));
    }
  }
  const int S = ns + nsurf;
  if (nsurf > 0) {
    const double lo = g > 0.f ? (double)__fmul_rn(0.95f, g) : 0.001, hi = g > 0.f ? (double)__fmul_rn(1.05f, g) : (double)dmax;
    for (int k = lane; k < nsurf; k += 32) z[ns + k] = __dadd_rn(__dmul_rn(lo, __dsub_rn(1.0, ts[k])), __dmul_rn(hi, ts[k]));
  }
  __syncwarp();
  double* out = zout + r * S;
  if (nsurf == 0) {
    for (int k = lane; k < S; k += 32) out[k] = z[k];
    return;
  }
  bool asc = true;
  for (int k = lane; k < S; k += 32)
    if (k > 0 && k != ns) asc = asc && (z[k - 1] <= z[k]);
  if (__all_sync(kFull, asc)) {
    for (int k = lane; k < S; k += 32) {
      const double v = z[k];
      int pos;
      if (k < ns) {           // uniform sample: surface samples strictly below come first
        int c = 0;
        for (int j = ns; j < S; ++j) c += z[j] < v;
        pos = k + c;
      } else {                // surface sample: uniform samples below or equal come first
        int c = 0;
        for (int j = 0; j < ns; ++j) c += z[j] <= v;
        pos = (k - ns) + c;
      }
      out[pos] = v;
    }
  } else {
    if (lane == 0) insertion_sort(z, S);
    __syncwarp();
    for (int k = lane; k < S; k += 32) out[k] = z[k];
  }
}

__global__ void __launch_bounds__(128) k_importance(const double* __restrict__ zin, const float* __restrict__ w, int64_t R, int S,
                                                   int ni, const float* __restrict__ u_lin, const float* __restrict__ u_rand,
                                                   double* __restrict__ zout) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  double z[PN_MAX_SAMPLES];
  float cdf[PN_MAX_SAMPLES];
  const double* zi = zin + r * S;
  const float* wi = w + r * S;
  const int nb = S - 1;  // bins = mid points, cdf has nb entries
  float sum = 0.f;
  for (int k = 1; k < S - 1; ++k) sum = __fadd_rn(sum, __fadd_rn(wi[k], 1e-5f));
  float run = 0.f;
  cdf[0] = 0.f;
  for (int k = 1; k < S - 1; ++k) {
    run = __fadd_rn(run, __fdiv_rn(__fadd_rn(wi[k], 1e-5f), sum));
    cdf[k] = run;
  }
  for (int k = 0; k < S; ++k) z[k] = zi[k];
  for (int m = 0; m < ni; ++m) {
    const float u = u_rand ? u_rand[r * ni + m] : u_lin[m];
    int ind = 0;  // searchsorted(right=True): first index with cdf[ind] > u
    while (ind < nb && !(cdf[ind] > u)) ++ind;
    const int below = max(ind - 1, 0), above = min(ind, nb - 1);
    float denom = __fsub_rn(cdf[above], cdf[below]);
    if (denom < 1e-5f) denom = 1.0f;
    const float t = __fdiv_rn(__fsub_rn(u, cdf[below]), denom);
    const double b0 = __dmul_rn(0.5, __dadd_rn(zi[below + 1], zi[below]));
    const double b1 = __dmul_rn(0.5, __dadd_rn(zi[above + 1], zi[above]));
    z[S + m] = __dadd_rn(b0, __dmul_rn((double)t, __dsub_rn(b1, b0)));
  }
  insertion_sort(z, S + ni);
  for (int k = 0; k < S + ni; ++k) zout[r * (S + ni) + k] = z[k];
}

// standalone sample_pdf (common.py:19-63): bins (R,nb) f64, weights (R,nb-1) f32 -> out (R,n) f64
__global__ void __launch_bounds__(128) k_sample_pdf(const double* __restrict__ bins, const float* __restrict__ w, int64_t R, int nb,
                                                   int n, const float* __restrict__ u_lin, const float* __restrict__ u_rand,
                                                   double* __restrict__ out) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  float cdf[PN_MAX_SAMPLES];
  const double* bi = bins + r * nb;
  const float* wi = w + r * (nb - 1);
  float sum = 0.f;
  for (int k = 0; k < nb - 1; ++k) sum = __fadd_rn(sum, __fadd_rn(wi[k], 1e-5f));
  float run = 0.f;
  cdf[0] = 0.f;
  for (int k = 0; k < nb - 1; ++k) {
    run = __fadd_rn(run, __fdiv_rn(__fadd_rn(wi[k], 1e-5f), sum));
    cdf[k + 1] = run;
  }
  for (int m = 0; m < n; ++m) {
    const float u = u_rand ? u_rand[r * n + m] : u_lin[m];
    int ind = 0;
    while (ind < nb && !(cdf[ind] > u)) ++ind;
    const int below = max(ind - 1, 0), above = min(ind, nb - 1);
    float denom = __fsub_rn(cdf[above], cdf[below]);
    if (denom < 1e-5f) denom = 1.0f;
    const float t = __fdiv_rn(__fsub_rn(u, cdf[below]), denom);
    out[r * n + m] = __dadd_rn(bi[below], __dmul_rn((double)t, __dsub_rn(bi[above], bi[below])));
  }
}

__global__ void k_regulation_points(const float* __restrict__ ro, const float* __restrict__ rd, const float* __restrict__ gt,
                                    const float* __restrict__ tv, const float* __restrict__ t_rand, int64_t R, int ns,
                                    float* __restrict__ pts, double* __restrict__ zout) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= R * ns) return;
  const int64_t r = i / ns;
  const int k = (int)(i - r * ns);
  const float far = __fmul_rn(gt[r], 0.85f);
  auto zk = [&](int q) { return __fadd_rn(__fmul_rn(0.0f, __fsub_rn(1.0f, tv[q])), __fmul_rn(far, tv[q])); };
  const float zc = zk(k);
  const float lower = k > 0 ? __fmul_rn(0.5f, __fadd_rn(zc, zk(k - 1))) : zc;
  const float upper = k + 1 < ns ? __fmul_rn(0.5f, __fadd_rn(zk(k + 1), zc)) : zc;
  const float z = __fadd_rn(lower, __fmul_rn(__fsub_rn(upper, lower), t_rand[i]));
  if (zout) zout[i] = (double)z;
#pragma unroll
  for (int a = 0; a < 3; ++a) pts[3 * i + a] = __fadd_rn(ro[3 * r + a], __fmul_rn(rd[3 * r + a], z));
}

// ------------------------------------------------------------------ compositing
constexpr int kMaxChunks = PN_MAX_SAMPLES / 32;

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}

// alpha of one sample (common.py:222-237)
__device__ __forceinline__ float sample_alpha(float sigma, float dist, int occupancy) {
  if (occupancy) return 1.0f / (1.0f + expf(-10.0f * sigma));
  return 1.0f - expf(-fmaxf(sigma, 0.f) * dist);
}

__global__ void __launch_bounds__(256) k_composite_fwd(const float* __restrict__ raw, const double* __restrict__ z,
                                                      const float* __restrict__ rd, int64_t R, int S, int occupancy,
                                                      double* __restrict__ depth, double* __restrict__ var,
                                                      float* __restrict__ rgb, float* __restrict__ weights) {
  const int lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= R) return;
  const float4* rw = reinterpret_cast<const float4*>(raw) + r * S;
  const double* zr = z + r * S;
  const float dn = sqrtf(rd[3 * r] * rd[3 * r] + rd[3 * r + 1] * rd[3 * r + 1] + rd[3 * r + 2] * rd[3 * r + 2]);
  float carry = 1.0f;
  double dsum = 0.0;
  float c0 = 0.f, c1 = 0.f, c2 = 0.f;
  float wk[kMaxChunks];
  double zk[kMaxChunks];
#pragma unroll
  for (int ch = 0; ch < kMaxChunks; ++ch) {
    wk[ch] = 0.f; zk[ch] = 0.0;
    if (ch * 32 < S) {
      const int k = ch * 32 + lane;
      const bool ok = k < S;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      double zz = 0.0;
      float alpha = 0.f;
      if (ok) {
        v = rw[k]; zz = zr[k];
        const float dist = (k + 1 < S ? (float)(zr[k + 1] - zz) : 1e10f) * dn;
        alpha = sample_alpha(v.w, dist, occupancy);
      }
      const float om = ok ? (1.0f - alpha) + 1e-10f : 1.0f;
      float incl = om;  // inclusive product scan
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const float t = __shfl_up_sync(kFull, incl, o);
        if (lane >= o) incl *= t;
      }
      float excl = __shfl_up_sync(kFull, incl, 1);
      if (lane == 0) excl = 1.0f;
      const float T = carry * excl;
      carry *= __shfl_sync(kFull, incl, 31);
      const float w = alpha * T;
      if (ok) {
        if (weights) weights[r * S + k] = w;
        wk[ch] = w; zk[ch] = zz;
        dsum += (double)w * zz;
        c0 = fmaf(w, v.x, c0); c1 = fmaf(w, v.y, c1); c2 = fmaf(w, v.z, c2);
      }
    }
  }
  dsum = warp_sum_d(dsum);
  c0 = warp_sum(c0); c1 = warp_sum(c1); c2 = warp_sum(c2);
  double vs = 0.0;
#pragma unroll
  for (int ch = 0; ch < kMaxChunks; ++ch) {
    const double t = zk[ch] - dsum;
    vs += ((double)wk[ch] * t) * t;
  }
  vs = warp_sum_d(vs);
  if (lane == 0) {
    depth[r] = dsum; var[r] = vs;
    rgb[3 * r] = c0; rgb[3 * r + 1] = c1; rgb[3 * r + 2] = c2;
  }
}

__global__ void __launch_bounds__(256) k_composite_bwd(const float* __restrict__ raw, const double* __restrict__ z,
                                                      const float* __restrict__ rd, int64_t R, int S, int occupancy,
                                                      const double* __restrict__ g_depth, const double* __restrict__ g_var,
                                                      const float* __restrict__ g_rgb, float* __restrict__ g_raw,
                                                      float* __restrict__ g_rays_d) {
  const int lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= R) return;
  const float4* rw = reinterpret_cast<const float4*>(raw) + r * S;
  const double* zr = z + r * S;
  const float dx = rd[3 * r], dy = rd[3 * r + 1], dz = rd[3 * r + 2];
  const float dn = sqrtf(dx * dx + dy * dy + dz * dz);
  float4 v[kMaxChunks];
  double zz[kMaxChunks];
  float alpha[kMaxChunks], om[kMaxChunks], T[kMaxChunks], w[kMaxChunks], dist[kMaxChunks];
  float carry = 1.0f;
  double dsum = 0.0;
#pragma unroll
  for (int ch = 0; ch < kMaxChunks; ++ch) {
    v[ch] = make_float4(0.f, 0.f, 0.f, 0.f); zz[ch] = 0.0; alpha[ch] = 0.f; om[ch] = 1.f; T[ch] = 0.f; w[ch] = 0.f; dist[ch] = 0.f;
    if (ch * 32 < S) {
      const int k = ch * 32 + lane;
      const bool ok = k < S;
      if (ok) {
        v[ch] = rw[k]; zz[ch] = zr[k];
        dist[ch] = (k + 1 < S ? (float)(zr[k + 1] - zz[ch]) : 1e10f);
        alpha[ch] = sample_alpha(v[ch].w, dist[ch] * dn, occupancy);
        om[ch] = (1.0f - alpha[ch]) + 1e-10f;
      }
      float incl = om[ch];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const float t = __shfl_up_sync(kFull, incl, o);
        if (lane >= o) incl *= t;
      }
      float excl = __shfl_up_sync(kFull, incl, 1);
      if (lane == 0) excl = 1.0f;
      T[ch] = carry * excl;
      carry *= __shfl_sync(kFull, incl, 31);
      w[ch] = ok ? alpha[ch] * T[ch] : 0.f;
      dsum += (double)w[ch] * zz[ch];
    }
  }
  dsum = warp_sum_d(dsum);
  const double gv = g_var ? g_var[r] : 0.0;
  double gd_eff = g_depth ? g_depth[r] : 0.0;
  if (g_var) {  // d var / d depth = -2 sum w (z - depth)
    double s = 0.0;
#pragma unroll
    for (int ch = 0; ch < kMaxChunks; ++ch) s += (double)w[ch] * (zz[ch] - dsum);
    s = warp_sum_d(s);
    gd_eff += gv * (-2.0 * s);
  }
  const float gr0 = g_rgb ? g_rgb[3 * r] : 0.f, gr1 = g_rgb ? g_rgb[3 * r + 1] : 0.f, gr2 = g_rgb ? g_rgb[3 * r + 2] : 0.f;
  // suffix sums of G_m w_m, last chunk first
  float suffix_carry = 0.f;
  float gnorm = 0.f;
#pragma unroll
  for (int ch = kMaxChunks - 1; ch >= 0; --ch) {
    if (ch * 32 < S) {
      const int k = ch * 32 + lane;
      const bool ok = k < S;
      const double t = zz[ch] - dsum;
      const float G = (float)(gd_eff * zz[ch] + gv * t * t) + gr0 * v[ch].x + gr1 * v[ch].y + gr2 * v[ch].z;
      const float x = ok ? G * w[ch] : 0.f;
      float incl = x;  // inclusive suffix sum within the chunk
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const float u = __shfl_down_sync(kFull, incl, o);
        if (lane + o < 32) incl += u;
      }
      float after = __shfl_down_sync(kFull, incl, 1);  // sum over m > k inside the chunk
      if (lane == 31) after = 0.f;
      after += suffix_carry;
      suffix_carry += __shfl_sync(kFull, incl, 0);
      if (ok) {
        const float dalpha = G * T[ch] - after / om[ch];
        float dsig;
        if (occupancy) {
          dsig = dalpha * 10.0f * alpha[ch] * (1.0f - alpha[ch]);
        } else {
          const float e = 1.0f - alpha[ch];  // exp(-relu(sigma) dist)
          dsig = v[ch].w > 0.f ? dalpha * dist[ch] * dn * e : 0.f;
          gnorm += dalpha * fmaxf(v[ch].w, 0.f) * dist[ch] * e;
        }
        reinterpret_cast<float4*>(g_raw)[r * S + k] = make_float4(w[ch] * gr0, w[ch] * gr1, w[ch] * gr2, dsig);
      }
    }
  }
  if (g_rays_d && !occupancy) {
    gnorm = warp_sum(gnorm);
    if (lane == 0 && dn > 0.f) {
      g_rays_d[3 * r] += gnorm * dx / dn; g_rays_d[3 * r + 1] += gnorm * dy / dn; g_rays_d[3 * r + 2] += gnorm * dz / dn;
    }
  }
}

__global__ void __launch_bounds__(256) k_points_to_rays(const float* __restrict__ gp, const double* __restrict__ z, int64_t R, int S,
                                                       float* __restrict__ go, float* __restrict__ gd) {
  const int lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= R) return;
  double so[3] = {0, 0, 0}, sd[3] = {0, 0, 0};
  for (int k = lane; k < S; k += 32) {
    const double zz = z[r * S + k];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const double g = (double)gp[3 * (r * S + k) + a];
      so[a] += g; sd[a] += g * zz;
    }
  }
#pragma unroll
  for (int a = 0; a < 3; ++a) { so[a] = warp_sum_d(so[a]); sd[a] = warp_sum_d(sd[a]); }
  if (lane == 0) {
#pragma unroll
    for (int a = 0; a < 3; ++a) { go[3 * r + a] = (float)so[a]; gd[3 * r + a] = (float)sd[a]; }
  }
}

// ------------------------------------------------------------------ utilities
__global__ void k_grid_transpose(const float* __restrict__ src, float* __restrict__ dst, int64_t V, int to_cl) {
  __shared__ float t[32][33];
  const int64_t v0 = (int64_t)blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  if (to_cl) {  // src [32][V] -> dst [V][32]
    for (int c = ty; c < 32; c += 8) { const int64_t v = v0 + tx; t[c][tx] = v < V ? src[(int64_t)c * V + v] : 0.f; }
    __syncthreads();
    for (int i = ty; i < 32; i += 8) { const int64_t v = v0 + i; if (v < V) dst[v * 32 + tx] = t[tx][i]; }
  } else {      // src [V][32] -> dst [32][V]
    for (int i = ty; i < 32; i += 8) { const int64_t v = v0 + i; t[i][tx] = v < V ? src[v * 32 + tx] : 0.f; }
    __syncthreads();
    for (int c = ty; c < 32; c += 8) { const int64_t v = v0 + tx; if (v < V) dst[(int64_t)c * V + v] = t[tx][c]; }
  }
}

__global__ void __launch_bounds__(1024) k_max_f32(const float* __restrict__ x, int64_t n, float* __restrict__ out) {
  __shared__ float red[32];
  float m = -INFINITY;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) m = fmaxf(m, x[i]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(kFull, m, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x < 32) {
    m = red[threadIdx.x];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(kFull, m, o));
    if (threadIdx.x == 0) out[0] = m;
  }
}

// ---------------------------------------------------------------------------
// Mapper loss head (src/Mapper.py:628-646) with its gradient, one launch.
// One CTA, fixed summation order (thread-strided partial sums, shuffle + shared tree): deterministic.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) k_mapping_loss(const double* __restrict__ depth, const float* __restrict__ color,
                                                       const float* __restrict__ gt_depth, const float* __restrict__ gt_color,
                                                       int64_t R, int use_color, float w_color, int depth_sup,
                                                       double* __restrict__ loss, double* __restrict__ g_depth,
                                                       float* __restrict__ g_color) {
  __shared__ double red_d[32], red_c[32];
  double sd = 0.0, sc = 0.0;
  if (!depth_sup) w_color = 1.0f;                  // Mapper.py:633-637: the colour term alone, unweighted
  for (int64_t r = threadIdx.x; r < R; r += blockDim.x) {
    const float g = gt_depth[r];
    const double diff = (double)g - depth[r];      // float32 - float64 promotes to float64 (torch)
    const bool m = depth_sup && g > 0.f;
    if (m) sd += fabs(diff);
    // d|gt - depth| / d depth = -sign(gt - depth), sign(0) = 0
    g_depth[r] = m ? (diff > 0.0 ? -1.0 : (diff < 0.0 ? 1.0 : 0.0)) : 0.0;
    if (use_color) {
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        const float dc = __fsub_rn(gt_color[3 * r + a], color[3 * r + a]);
        sc += (double)fabsf(dc);
        g_color[3 * r + a] = dc > 0.f ? -w_color : (dc < 0.f ? w_color : 0.f);
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { sd += __shfl_xor_sync(kFull, sd, o); sc += __shfl_xor_sync(kFull, sc, o); }
  if ((threadIdx.x & 31) == 0) { red_d[threadIdx.x >> 5] = sd; red_c[threadIdx.x >> 5] = sc; }
  __syncthreads();
  if (threadIdx.x < 32) {
    sd = threadIdx.x < (blockDim.x >> 5) ? red_d[threadIdx.x] : 0.0;
    sc = threadIdx.x < (blockDim.x >> 5) ? red_c[threadIdx.x] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { sd += __shfl_xor_sync(kFull, sd, o); sc += __shfl_xor_sync(kFull, sc, o); }
    // the reference sums the colour residuals in float32 and scales in float32 before adding to the float64 depth term
    if (threadIdx.x == 0) *loss = use_color ? sd + (double)__fmul_rn(w_color, (float)sc) : sd;
  }
}

// ---------------------------------------------------------------------------
// Tracker loss head (src/Tracker.py:306-330) with its gradient, one launch, one CTA.
//   tmp  = |gt_depth - depth| / sqrt(var + 1e-10)                      (float64)
//   mask = gt_depth > 0  [ & tmp < 10 * median(tmp)  when handle_dynamic ]
//   loss = sum_mask tmp + w_color * sum_mask |gt_color - color|
// torch.median of n values is the element of rank (n-1)/2: tmp is sorted in shared memory (bitonic, padded with
// +inf to a power of two), so R <= kTrackMaxRays.  Fixed summation order: deterministic.
// ---------------------------------------------------------------------------
constexpr int kTrackMaxRays = 8192;   // up to here tmp is sorted in shared memory; beyond, the median comes from a radix select

__device__ __forceinline__ double track_tmp(const double* depth, const double* var, const float* gt_depth, int r) {
  return __ddiv_rn(fabs((double)gt_depth[r] - depth[r]), __dsqrt_rn(__dadd_rn(var[r], 1e-10)));
}

__global__ void __launch_bounds__(1024) k_tracking_loss(const double* __restrict__ depth, const double* __restrict__ var,
                                                        const float* __restrict__ color, const float* __restrict__ gt_depth,
                                                        const float* __restrict__ gt_color, int R, int handle_dynamic,
                                                        int use_color, float w_color, int depth_sup, double* __restrict__ loss,
                                                        double* __restrict__ g_depth, float* __restrict__ g_color) {
  extern __shared__ double srt[];
  __shared__ double red_d[32], red_c[32];
  __shared__ double thr_s;
  __shared__ int hist[256];
  __shared__ int nan_s;
  __shared__ unsigned long long prefix_s;
  __shared__ int rank_s;
  const int tid = threadIdx.x;
  double thr = 0.0;
  if (!depth_sup) w_color = 1.0f;     // Tracker.py:313-318: the masked colour term alone, unweighted
  if (handle_dynamic) {
    if (tid == 0) nan_s = 0;
    __syncthreads();
    const double kNaN = __longlong_as_double(0x7ff8000000000000LL);
    if (R <= kTrackMaxRays) {
      int P2 = 1;
      while (P2 < R) P2 <<= 1;
      for (int r = tid; r < P2; r += blockDim.x) {
        const double t = r < R ? track_tmp(depth, var, gt_depth, r) : __longlong_as_double(0x7ff0000000000000LL);
        if (t != t) nan_s = 1;          // torch.median propagates NaN: the mask becomes all-false
        srt[r] = t;
      }
      __syncthreads();
      for (int k = 2; k <= P2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
          for (int i = tid; i < P2; i += blockDim.x) {
            const int l = i ^ j;
            if (l > i) {
              const double a = srt[i], b = srt[l];
              const bool up = (i & k) == 0;
              if ((a > b) == up) { srt[i] = b; srt[l] = a; }
            }
          }
          __syncthreads();
        }
      }
      if (tid == 0) thr_s = nan_s ? kNaN : __dmul_rn(10.0, srt[(R - 1) / 2]);
    } else {
      // radix select of the element of rank (R-1)/2: tmp >= 0, so the unsigned order of the bit patterns is the
      // numeric order; eight 8-bit passes, most significant byte first, tmp recomputed on the fly
      if (tid == 0) { prefix_s = 0ull; rank_s = (R - 1) / 2; }
      __syncthreads();
      for (int shift = 56; shift >= 0; shift -= 8) {
        for (int i = tid; i < 256; i += blockDim.x) hist[i] = 0;
        __syncthreads();
        const unsigned long long prefix = prefix_s;
        const unsigned long long hi_mask = shift == 56 ? 0ull : (~0ull << (shift + 8));
        for (int r = tid; r < R; r += blockDim.x) {
          const double t = track_tmp(depth, var, gt_depth, r);
          if (t != t) { nan_s = 1; continue; }
          const unsigned long long b = (unsigned long long)__double_as_longlong(t);
          if ((b & hi_mask) == prefix) atomicAdd(&hist[(int)((b >> shift) & 0xffull)], 1);
        }
        __syncthreads();
        if (tid == 0) {
          int k = rank_s, bkt = 0;
          while (bkt < 255 && k >= hist[bkt]) { k -= hist[bkt]; ++bkt; }
          rank_s = k;
          prefix_s = prefix | ((unsigned long long)bkt << shift);
        }
        __syncthreads();
      }
      if (tid == 0) thr_s = nan_s ? kNaN : __dmul_rn(10.0, __longlong_as_double((long long)prefix_s));
    }
    __syncthreads();
    thr = thr_s;
  }
  double sd = 0.0, sc = 0.0;
  for (int r = tid; r < R; r += blockDim.x) {
    const float g = gt_depth[r];
    const double diff = (double)g - depth[r];
    const double inv = __ddiv_rn(1.0, __dsqrt_rn(__dadd_rn(var[r], 1e-10)));
    const double t = __ddiv_rn(fabs(diff), __dsqrt_rn(__dadd_rn(var[r], 1e-10)));
    const bool m = (g > 0.f) && (!handle_dynamic || t < thr);
    if (m && depth_sup) sd += t;
    // d(|gt - depth| / s)/d depth = -sign(gt - depth) / s
    g_depth[r] = (m && depth_sup) ? (diff > 0.0 ? -inv : (diff < 0.0 ? inv : 0.0)) : 0.0;
    if (use_color) {
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        const float dc = __fsub_rn(gt_color[3 * r + a], color[3 * r + a]);
        if (m) sc += (double)fabsf(dc);
        g_color[3 * r + a] = m ? (dc > 0.f ? -w_color : (dc < 0.f ? w_color : 0.f)) : 0.f;
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { sd += __shfl_xor_sync(kFull, sd, o); sc += __shfl_xor_sync(kFull, sc, o); }
  if ((tid & 31) == 0) { red_d[tid >> 5] = sd; red_c[tid >> 5] = sc; }
  __syncthreads();
  if (tid < 32) {
    sd = tid < (blockDim.x >> 5) ? red_d[tid] : 0.0;
    sc = tid < (blockDim.x >> 5) ? red_c[tid] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { sd += __shfl_xor_sync(kFull, sd, o); sc += __shfl_xor_sync(kFull, sc, o); }
    if (tid == 0) *loss = use_color ? sd + (double)__fmul_rn(w_color, (float)sc) : sd;
  }
}

}  // namespace
}  // namespace pn

using namespace pn;
#define PN_ST ((cudaStream_t)stream)

extern "C" int pn_camera_from_tensor_fwd(const float* cam, int batch, float* c2w, void* stream) {
  if (!cam || !c2w || batch < 0) { set_error("pn_camera_from_tensor_fwd: bad arguments"); return 1; }
  if (batch == 0) return 0;
  k_cam_fwd<<<(batch + 63) / 64, 64, 0, PN_ST>>>(cam, batch, c2w);
  return launch_status("k_cam_fwd");
}

extern "C" int pn_camera_from_tensor_bwd(const float* cam, const float* g_c2w, int batch, float* g_cam, void* stream) {
  if (!cam || !g_c2w || !g_cam || batch < 0) { set_error("pn_camera_from_tensor_bwd: bad arguments"); return 1; }
  if (batch == 0) return 0;
  k_cam_bwd<<<(batch + 63) / 64, 64, 0, PN_ST>>>(cam, g_c2w, batch, g_cam);
  return launch_status("k_cam_bwd");
}

extern "C" int pn_sample_rays_fwd(const int64_t* idx, int n, int H0, int W0, int Wc, int W, float fx, float fy, float cx,
                                  float cy, const float* c2w, int c2w_ld, const float* depth_img, const void* color_img,
                                  int color_is_f64, float* rays_o, float* rays_d, float* depth_out, void* color_out,
                                  void* stream) {
  if (!idx || !c2w || !rays_o || !rays_d || n < 0 || Wc <= 0 || (depth_out && !depth_img) || (color_out && !color_img)) {
    set_error("pn_sample_rays_fwd: bad arguments");
    return 1;
  }
  if (n == 0) return 0;
  const int g = (n + 127) / 128;
  if (color_is_f64)
    k_sample_rays<double><<<g, 128, 0, PN_ST>>>(idx, n, H0, W0, Wc, W, fx, fy, cx, cy, c2w, c2w_ld, depth_img,
                                                (const double*)color_img, rays_o, rays_d, depth_out, (double*)color_out);
  else
    k_sample_rays<float><<<g, 128, 0, PN_ST>>>(idx, n, H0, W0, Wc, W, fx, fy, cx, cy, c2w, c2w_ld, depth_img,
                                               (const float*)color_img, rays_o, rays_d, depth_out, (float*)color_out);
  return launch_status("k_sample_rays");
}

extern "C" int pn_sample_rays_multi_fwd(const int64_t* idx, int F, int n, int H0, int W0, int Wc, int W, float fx, float fy,
                                       float cx, float cy, const float* c2w, const float* const* depth_ptrs,
                                       const void* const* color_ptrs, int color_is_f64, float* rays_o, float* rays_d,
                                       float* depth_out, void* color_out, void* stream) {
  if (!idx || !c2w || !depth_ptrs || !color_ptrs || !rays_o || !rays_d || !depth_out || !color_out || F <= 0 || F > 65535 || n < 0 ||
      Wc <= 0) {
    set_error("pn_sample_rays_multi_fwd: bad arguments");
    return 1;
  }
  if (n == 0) return 0;
  const dim3 g((n + 127) / 128, F);
  if (color_is_f64)
    k_sample_rays_multi<double><<<g, 128, 0, PN_ST>>>(idx, n, H0, W0, Wc, W, fx, fy, cx, cy, c2w, depth_ptrs,
                                                      (const double* const*)color_ptrs, rays_o, rays_d, depth_out, (double*)color_out);
  else
    k_sample_rays_multi<float><<<g, 128, 0, PN_ST>>>(idx, n, H0, W0, Wc, W, fx, fy, cx, cy, c2w, depth_ptrs,
                                                     (const float* const*)color_ptrs, rays_o, rays_d, depth_out, (float*)color_out);
  return launch_status("k_sample_rays_multi");
}

extern "C" int pn_rays_multi_bwd(const int64_t* idx, int F, int n, int H0, int W0, int Wc, float fx, float fy, float cx, float cy,
                                 const float* g_rays_o, const float* g_rays_d, float* g_c2w, void* stream) {
  if (!idx || !g_c2w || F <= 0 || F > 65535 || n < 0 || Wc <= 0) { set_error("pn_rays_multi_bwd: bad arguments"); return 1; }
  if (n == 0) return 0;
  int gx = (n + 255) / 256;
  if (gx > 8) gx = 8;
  k_rays_bwd_multi<<<dim3(gx, F), 256, 0, PN_ST>>>(idx, n, H0, W0, Wc, fx, fy, cx, cy, g_rays_o, g_rays_d, g_c2w);
  return launch_status("k_rays_bwd_multi");
}

extern "C" int pn_image_rays_fwd(int H, int W, float fx, float fy, float cx, float cy, const float* c2w, int c2w_ld,
                                 float* rays_o, float* rays_d, void* stream) {
  if (!c2w || !rays_o || !rays_d || H <= 0 || W <= 0) { set_error("pn_image_rays_fwd: bad arguments"); return 1; }
  const int64_t n = (int64_t)H * W;
  k_image_rays<<<(unsigned)((n + 255) / 256), 256, 0, PN_ST>>>(H, W, fx, fy, cx, cy, c2w, c2w_ld, rays_o, rays_d);
  return launch_status("k_image_rays");
}

extern "C" int pn_rays_bwd(const int64_t* idx, int n, int H0, int W0, int Wc, float fx, float fy, float cx, float cy,
                           const float* g_rays_o, const float* g_rays_d, float* g_c2w, void* stream) {
  if (!g_c2w || n < 0 || Wc <= 0) { set_error("pn_rays_bwd: bad arguments"); return 1; }
  if (n == 0) return 0;
  int g = (n + 255) / 256;
  if (g > 2 * sm_count()) g = 2 * sm_count();
  k_rays_bwd<<<g, 256, 0, PN_ST>>>(idx, n, H0, W0, Wc, fx, fy, cx, cy, g_rays_o, g_rays_d, g_c2w);
  return launch_status("k_rays_bwd");
}

extern "C" int pn_ray_zvals(const float* rays_o, const float* rays_d, const float* gt_depth, const float* depth_max,
                            int64_t R, const double* bound, int n_samples, int n_surface, int lindisp,
                            const float* t_vals, const double* t_surface, const float* t_rand, double* z_out,
                            void* stream) {
  if (!rays_o || !rays_d || !bound || !t_vals || !z_out || n_samples <= 0 || n_surface < 0 ||
      n_samples + n_surface > PN_MAX_SAMPLES) {
    set_error("pn_ray_zvals: bad arguments (n_samples+n_surface must be <= %d)", PN_MAX_SAMPLES);
    return 1;
  }
  if (gt_depth && (!depth_max || (n_surface > 0 && !t_surface))) { set_error("pn_ray_zvals: depth given without depth_max/t_surface"); return 1; }
  if (R == 0) return 0;
  if (!t_rand) {   // common case: warp per ray
    k_ray_zvals_warp<<<(unsigned)((R + 7) / 8), 256, 0, PN_ST>>>(rays_o, rays_d, gt_depth, depth_max, R, make_bound(bound), n_samples,
                                                               gt_depth ? n_surface : 0, lindisp, t_vals, t_surface, z_out);
    return launch_status("k_ray_zvals_warp");
  }
  k_ray_zvals<<<(unsigned)((R + 127) / 128), 128, 0, PN_ST>>>(rays_o, rays_d, gt_depth, depth_max, R, make_bound(bound),
                                                             n_samples, gt_depth ? n_surface : 0, lindisp, t_vals,
                                                             t_surface, t_rand, z_out);
  return launch_status("k_ray_zvals");
}

extern "C" int pn_importance_zvals(const double* z, const float* weights, int64_t R, int S, int n_imp, const float* u_lin,
                                   const float* u_rand, double* z_out, void* stream) {
  if (!z || !weights || !z_out || S < 3 || n_imp <= 0 || S + n_imp > PN_MAX_SAMPLES || (!u_lin && !u_rand)) {
    set_error("pn_importance_zvals: bad arguments");
    return 1;
  }
  if (R == 0) return 0;
  k_importance<<<(unsigned)((R + 127) / 128), 128, 0, PN_ST>>>(z, weights, R, S, n_imp, u_lin, u_rand, z_out);
  return launch_status("k_importance");
}

extern "C" int pn_sample_pdf(const double* bins, const float* weights, int64_t R, int nb, int n, const float* u_lin,
                             const float* u_rand, double* out, void* stream) {
  if (!bins || !weights || !out || nb < 2 || nb > PN_MAX_SAMPLES || n <= 0 || (!u_lin && !u_rand)) {
    set_error("pn_sample_pdf: bad arguments");
    return 1;
  }
  if (R == 0) return 0;
  k_sample_pdf<<<(unsigned)((R + 127) / 128), 128, 0, PN_ST>>>(bins, weights, R, nb, n, u_lin, u_rand, out);
  return launch_status("k_sample_pdf");
}

extern "C" int pn_regulation_points(const float* rays_o, const float* rays_d, const float* gt_depth, const float* t_vals,
                                    const float* t_rand, int64_t R, int n_samples, float* pts_out, double* z_out,
                                    void* stream) {
  if (!rays_o || !rays_d || !gt_depth || !t_vals || !t_rand || !pts_out || n_samples <= 0) {
    set_error("pn_regulation_points: bad arguments");
    return 1;
  }
  if (R == 0) return 0;
  const int64_t n = R * n_samples;
  k_regulation_points<<<(unsigned)((n + 255) / 256), 256, 0, PN_ST>>>(rays_o, rays_d, gt_depth, t_vals, t_rand, R, n_samples, pts_out, z_out);
  return launch_status("k_regulation_points");
}

extern "C" int pn_composite_fwd(const float* raw, const double* z, const float* rays_d, int64_t R, int S, int occupancy,
                                double* depth, double* var, float* rgb, float* weights, void* stream) {
  if (!raw || !z || !rays_d || !depth || !var || !rgb || S <= 0 || S > PN_MAX_SAMPLES) {
    set_error("pn_composite_fwd: bad arguments (S must be in 1..%d)", PN_MAX_SAMPLES);
    return 1;
  }
  if (R == 0) return 0;
  k_composite_fwd<<<(unsigned)((R + 7) / 8), 256, 0, PN_ST>>>(raw, z, rays_d, R, S, occupancy, depth, var, rgb, weights);
  return launch_status("k_composite_fwd");
}

extern "C" int pn_composite_bwd(const float* raw, const double* z, const float* rays_d, int64_t R, int S, int occupancy,
                                const double* g_depth, const double* g_var, const float* g_rgb, float* g_raw,
                                float* g_rays_d, void* stream) {
  if (!raw || !z || !rays_d || !g_raw || S <= 0 || S > PN_MAX_SAMPLES) {
    set_error("pn_composite_bwd: bad arguments (S must be in 1..%d)", PN_MAX_SAMPLES);
    return 1;
  }
  if (R == 0) return 0;
  k_composite_bwd<<<(unsigned)((R + 7) / 8), 256, 0, PN_ST>>>(raw, z, rays_d, R, S, occupancy, g_depth, g_var, g_rgb, g_raw, g_rays_d);
  return launch_status("k_composite_bwd");
}

extern "C" int pn_points_to_rays_bwd(const float* g_pts, const double* z, int64_t R, int S, float* g_rays_o,
                                     float* g_rays_d, void* stream) {
  if (!g_pts || !z || !g_rays_o || !g_rays_d || S <= 0) { set_error("pn_points_to_rays_bwd: bad arguments"); return 1; }
  if (R == 0) return 0;
  k_points_to_rays<<<(unsigned)((R + 7) / 8), 256, 0, PN_ST>>>(g_pts, z, R, S, g_rays_o, g_rays_d);
  return launch_status("k_points_to_rays");
}

extern "C" int pn_grid_transpose(const float* src, float* dst, int D, int H, int W, int to_channels_last, void* stream) {
  if (!src || !dst || D <= 0 || H <= 0 || W <= 0) { set_error("pn_grid_transpose: bad arguments"); return 1; }
  const int64_t V = (int64_t)D * H * W;
  k_grid_transpose<<<(unsigned)((V + 31) / 32), 256, 0, PN_ST>>>(src, dst, V, to_channels_last);
  return launch_status("k_grid_transpose");
}

extern "C" int pn_tracking_loss(const double* depth, const double* var, const float* color, const float* gt_depth,
                                const float* gt_color, int64_t R, int handle_dynamic, int use_color, float w_color,
                                int depth_supervision, double* loss, double* g_depth, float* g_color, void* stream) {
  if (!depth || !var || !gt_depth || !loss || !g_depth || R < 1 || R > 0x7fffffff || (use_color && (!color || !gt_color || !g_color))) {
    set_error("pn_tracking_loss: bad arguments");
    return 1;
  }
  size_t sm = 0;
  if (handle_dynamic && R <= pn::kTrackMaxRays) { size_t p2 = 1; while ((int64_t)p2 < R) p2 <<= 1; sm = p2 * sizeof(double); }
  cudaFuncSetAttribute(pn::k_tracking_loss, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(pn::kTrackMaxRays * sizeof(double)));
  pn::k_tracking_loss<<<1, 1024, sm, PN_ST>>>(depth, var, color, gt_depth, gt_color, (int)R, handle_dynamic, use_color, w_color,
                                             depth_supervision, loss, g_depth, g_color);
  return launch_status("k_tracking_loss");
}

extern "C" int pn_mapping_loss(const double* depth, const float* color, const float* gt_depth, const float* gt_color, int64_t R,
                               int use_color, float w_color, int depth_supervision, double* loss, double* g_depth, float* g_color,
                               void* stream) {
  if (!depth || !gt_depth || !loss || !g_depth || R < 0 || (use_color && (!color || !gt_color || !g_color))) {
    set_error("pn_mapping_loss: bad arguments");
    return 1;
  }
  k_mapping_loss<<<1, 1024, 0, PN_ST>>>(depth, color, gt_depth, gt_color, R, use_color, w_color, depth_supervision, loss, g_depth, g_color);
  return launch_status("k_mapping_loss");
}

extern "C" int pn_max_f32(const float* x, int64_t n, float* out, void* stream) {
  if (!x || !out || n < 1) { set_error("pn_max_f32: bad arguments"); return 1; }
  k_max_f32<<<1, 1024, 0, PN_ST>>>(x, n, out);
  return launch_status("k_max_f32");
}
