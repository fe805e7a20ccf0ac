// Kernels for the callers either side of the rendering path: the Mapper's frustum
// feature selection, its masked Adam step, the ray pre-filter, pixel selection by depth, keyframe
// overlap selection, the Mesher's point masks, the Visualizer's residual panels -- and the sparse
// gradient exchange of the data-parallel mapper (touched 128-byte voxel rows instead of dense grids).
//
// Integer / boolean outputs (masks, indices, counts) are bit-exact against oracle/mapper_oracle.py:
// every floating-point step that feeds a comparison uses explicitly rounded intrinsics in the
// reference's dtype and evaluation order (no FMA contraction).
#include "pn_common.cuh"

namespace pn {
namespace {

// ---------------------------------------------------------------------------------------------
// projection of a world point into a camera (src/Mapper.py:152-163, 303-315; Mesher.py:127-145)
// ---------------------------------------------------------------------------------------------
struct Cam16 { float m[16]; };   // row-major world-to-camera matrix (host np.linalg.inv, as the reference)

struct Proj {
  float u, v;      // pixel coordinates, float32
  double z;        // camera z + eps (float64): visible points have z < 0
  float cz;        // camera z, float32
};

// cam = w2c @ [p;1] in float32, ((m0*x + m1*y) + m2*z) + m3*1, x negated; uv = K @ cam in float64 (all nine
// products, as numpy's float64 matmul forms them); uv / (z + eps) -> float32
__device__ __forceinline__ Proj project(const Cam16& w, float x, float y, float z, double fx, double fy, double cx, double cy,
                                        double eps) {
  float c[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    float acc = __fmul_rn(w.m[4 * i], x);
    acc = __fadd_rn(acc, __fmul_rn(w.m[4 * i + 1], y));
    acc = __fadd_rn(acc, __fmul_rn(w.m[4 * i + 2], z));
    acc = __fadd_rn(acc, __fmul_rn(w.m[4 * i + 3], 1.0f));
    c[i] = acc;
  }
  c[0] = __fmul_rn(c[0], -1.0f);
  const double c0 = (double)c[0], c1 = (double)c[1], c2 = (double)c[2];
  const double u = __dadd_rn(__dadd_rn(__dmul_rn(fx, c0), __dmul_rn(0.0, c1)), __dmul_rn(cx, c2));
  const double v = __dadd_rn(__dadd_rn(__dmul_rn(0.0, c0), __dmul_rn(fy, c1)), __dmul_rn(cy, c2));
  const double zz = __dadd_rn(__dadd_rn(__dmul_rn(0.0, c0), __dmul_rn(0.0, c1)), __dmul_rn(1.0, c2));
  Proj p;
  p.z = __dadd_rn(zz, eps);
  p.u = (float)__ddiv_rn(u, p.z);
  p.v = (float)__ddiv_rn(v, p.z);
  p.cz = c[2];
  return p;
}

// torch's float32 variant (Mesher.py:136-145): cam = w2c @ [p;1] float32, K.float() @ cam float32, z = w + 1e-8f
__device__ __forceinline__ Proj project_f32(const Cam16& w, float x, float y, float z, float fx, float fy, float cx, float cy) {
  float c[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    float acc = __fmul_rn(w.m[4 * i], x);
    acc = __fadd_rn(acc, __fmul_rn(w.m[4 * i + 1], y));
    acc = __fadd_rn(acc, __fmul_rn(w.m[4 * i + 2], z));
    acc = __fadd_rn(acc, __fmul_rn(w.m[4 * i + 3], 1.0f));
    c[i] = acc;
  }
  c[0] = __fmul_rn(c[0], -1.0f);
  const float u = __fadd_rn(__fadd_rn(__fmul_rn(fx, c[0]), __fmul_rn(0.0f, c[1])), __fmul_rn(cx, c[2]));
  const float v = __fadd_rn(__fadd_rn(__fmul_rn(0.0f, c[0]), __fmul_rn(fy, c[1])), __fmul_rn(cy, c[2]));
  const float zz = __fadd_rn(__fadd_rn(__fmul_rn(0.0f, c[0]), __fmul_rn(0.0f, c[1])), __fmul_rn(1.0f, c[2]));
  Proj p;
  const float zf = __fadd_rn(zz, 1e-8f);
  p.z = (double)zf;
  p.u = __fdiv_rn(u, zf);
  p.v = __fdiv_rn(v, zf);
  p.cz = c[2];
  return p;
}

// ---------------------------------------------------------------------------------------------
// cv2.remap(float32 image, INTER_LINEAR, BORDER_CONSTANT 0): 5-bit fixed-point coordinates
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int cv_round_scaled(float x) {   // cvRound(x * 32): half to even; NaN / overflow -> INT_MIN
  const float s = __fmul_rn(x, 32.0f);
  if (!(s >= -2147483648.0f && s < 2147483648.0f)) return INT_MIN;
  return __float2int_rn(s);
}

__device__ __forceinline__ float cv_remap_bilinear(const float* __restrict__ img, int H, int W, float x, float y) {
  const int sx = cv_round_scaled(x), sy = cv_round_scaled(y);
  int ix = sx >> 5, iy = sy >> 5;
  ix = max(-32768, min(32767, ix));
  iy = max(-32768, min(32767, iy));
  const float ax = __fmul_rn((float)(sx & 31), 0.03125f), ay = __fmul_rn((float)(sy & 31), 0.03125f);
  const float wx0 = __fsub_rn(1.0f, ax), wy0 = __fsub_rn(1.0f, ay);
  const float w0 = __fmul_rn(wy0, wx0), w1 = __fmul_rn(wy0, ax), w2 = __fmul_rn(ay, wx0), w3 = __fmul_rn(ay, ax);
  auto tap = [&](int yy, int xx) -> float {
    return (xx >= 0 && xx < W && yy >= 0 && yy < H) ? __ldg(img + (int64_t)yy * W + xx) : 0.0f;
  };
  float out = __fmul_rn(tap(iy, ix), w0);
  out = __fadd_rn(out, __fmul_rn(tap(iy, ix + 1), w1));
  out = __fadd_rn(out, __fmul_rn(tap(iy + 1, ix), w2));
  out = __fadd_rn(out, __fmul_rn(tap(iy + 1, ix + 1), w3));
  return out;
}

// ---------------------------------------------------------------------------------------------
// frustum feature selection (src/Mapper.py:129-200).  Pass 1: remapped depth of every voxel centre and
// its maximum; pass 2: the mask.  Voxel (x,y,z) of the (Z,Y,X) grid sits at (xs[x], ys[y], zs[z]).
// ---------------------------------------------------------------------------------------------
struct FrustumArgs {
  const float* xs; const float* ys; const float* zs;
  int nx, ny, nz;
  Cam16 w2c;
  float cam_o[3];
  double fx, fy, cx, cy;
  int H, W;
  const float* depth;
  float* vox_depth;        // [nz*ny*nx]
  unsigned int* max_bits;  // float bits of the maximum (depths are >= 0)
  uint8_t* mask;           // [nz][ny][nx]
};

__global__ void __launch_bounds__(256) k_frustum_depth(const FrustumArgs a) {
  const int64_t V = (int64_t)a.nx * a.ny * a.nz;
  float best = 0.0f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < V; i += (int64_t)gridDim.x * blockDim.x) {
    const int x = (int)(i % a.nx), y = (int)((i / a.nx) % a.ny), z = (int)(i / ((int64_t)a.nx * a.ny));
    const Proj p = project(a.w2c, a.xs[x], a.ys[y], a.zs[z], a.fx, a.fy, a.cx, a.cy, 1e-5);
    const float d = cv_remap_bilinear(a.depth, a.H, a.W, p.u, p.v);
    a.vox_depth[i] = d;
    best = fmaxf(best, d);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) best = fmaxf(best, __shfl_xor_sync(kFull, best, o));
  if ((threadIdx.x & 31) == 0 && best > 0.0f) atomicMax(a.max_bits, __float_as_uint(best));
}

__global__ void __launch_bounds__(256) k_frustum_mask(const FrustumArgs a) {
  const int64_t V = (int64_t)a.nx * a.ny * a.nz;
  const float dmax = __uint_as_float(*a.max_bits);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < V; i += (int64_t)gridDim.x * blockDim.x) {
    const int x = (int)(i % a.nx), y = (int)((i / a.nx) % a.ny), z = (int)(i / ((int64_t)a.nx * a.ny));
    const float px = a.xs[x], py = a.ys[y], pz = a.zs[z];
    const Proj p = project(a.w2c, px, py, pz, a.fx, a.fy, a.cx, a.cy, 1e-5);
    bool m = (p.u < (float)a.W) && (p.u > 0.0f) && (p.v < (float)a.H) && (p.v > 0.0f);
    float d = a.vox_depth[i];
    if (d == 0.0f) d = dmax;
    const double nz = -p.z;
    m = m && (0.0 <= nz) && (nz <= (double)__fadd_rn(d, 0.5f));
    const float dx = __fsub_rn(px, a.cam_o[0]), dy = __fsub_rn(py, a.cam_o[1]), dz = __fsub_rn(pz, a.cam_o[2]);
    const float dist = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
    m = m || ((double)dist < 0.25);
    a.mask[i] = m ? 1 : 0;
  }
}

// ---------------------------------------------------------------------------------------------
// masked multi-tensor Adam (torch.optim.Adam, single-tensor arithmetic; src/Mapper.py:482-505, 657-674)
// ---------------------------------------------------------------------------------------------
struct AdamTensor {      // one row of the device descriptor table (9 x 8 bytes)
  float* p; const float* g; float* m; float* v;
  const int64_t* idx;    // ascending indices of the selected voxels (the frustum mask, compacted once), or NULL = every element
  int64_t n;             // elements of the tensor
  int64_t nwork;         // elements to update: n, or (selected voxels) x (channels)
  int32_t row;           // > 0: channels-last grid, element = voxel * row + channel; < 0: NCDHW grid, element = channel * (-row) + voxel
  int32_t group;
  int64_t block0;        // first block of this tensor in the launch
};

struct AdamGroups {
  const double* lr;      // device, per group (changed by the host between replays of a captured step)
  int32_t* step;         // device, per TENSOR: steps taken so far
  float2* hyper;         // device, per TENSOR: (lr / (1 - beta1^step), sqrt(1 - beta2^step)) of the step being taken
  double beta1, beta2, eps;
  int ntensors;
};

constexpr int kAdamPerBlock = 256 * 4;

// torch.optim.Adam keeps one step count per parameter and advances it only when the parameter has a gradient
__global__ void k_adam_prepare(const AdamTensor* __restrict__ tab, const AdamGroups g) {
  for (int i = threadIdx.x; i < g.ntensors; i += blockDim.x) {
    if (tab[i].g == nullptr) continue;
    const int step = g.step[i] + 1;
    g.step[i] = step;
    const double bc1 = 1.0 - pow(g.beta1, (double)step), bc2 = 1.0 - pow(g.beta2, (double)step);
    g.hyper[i] = make_float2((float)(g.lr[tab[i].group] / bc1), (float)sqrt(bc2));
  }
}

__global__ void __launch_bounds__(256) k_adam(const AdamTensor* __restrict__ tab, const AdamGroups g) {
  __shared__ AdamTensor t;
  __shared__ float2 hyper_s;
  if (threadIdx.x == 0) {
    int lo = 0, hi = g.ntensors - 1;           // last tensor whose block0 <= blockIdx.x
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (tab[mid].block0 <= (int64_t)blockIdx.x) lo = mid; else hi = mid - 1;
    }
    t = tab[lo];
    hyper_s = g.hyper[lo];
  }
  __syncthreads();
  if (t.g == nullptr) return;
  const float b2 = (float)g.beta2, eps = (float)g.eps, step_size = hyper_s.x, bc2s = hyper_s.y;
  const float omb1 = (float)(1.0 - g.beta1), omb2 = (float)(1.0 - g.beta2);
  const int64_t base = ((int64_t)blockIdx.x - t.block0) * kAdamPerBlock;
  const int64_t chans = t.row > 0 ? (int64_t)t.row : (t.row < 0 ? t.n / (int64_t)(-t.row) : 1);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int64_t e = base + (int64_t)k * 256 + threadIdx.x;
    if (e >= t.nwork) break;
    int64_t i = e;
    if (t.idx) {
      const int64_t vox = t.idx[e / chans], ch = e % chans;
      i = t.row > 0 ? vox * chans + ch : ch * (int64_t)(-t.row) + vox;
    }
    const float gr = t.g[i];
    float m = t.m[i], v = t.v[i];
    m = __fadd_rn(m, __fmul_rn(omb1, __fsub_rn(gr, m)));                       // exp_avg.lerp_(grad, 1 - beta1)
    v = __fadd_rn(__fmul_rn(v, b2), __fmul_rn(__fmul_rn(omb2, gr), gr));       // mul_(beta2).addcmul_(grad, grad, 1 - beta2)
    const float denom = __fadd_rn(__fdiv_rn(__fsqrt_rn(v), bc2s), eps);
    t.p[i] = __fadd_rn(t.p[i], __fmul_rn(-step_size, __fdiv_rn(m, denom)));    // addcdiv_(exp_avg, denom, -step_size)
    t.m[i] = m; t.v[i] = v;
  }
}

// ---------------------------------------------------------------------------------------------
// stable compaction: flags (uint8, n) -> ascending indices of the set flags + their number
// ---------------------------------------------------------------------------------------------
constexpr int kCompBlock = 1024;

__global__ void __launch_bounds__(kCompBlock) k_flag_counts(const uint8_t* __restrict__ flags, int64_t n, int32_t* __restrict__ counts) {
  const int64_t i = (int64_t)blockIdx.x * kCompBlock + threadIdx.x;
  const int f = (i < n && flags[i]) ? 1 : 0;
  const int c = __syncthreads_count(f);
  if (threadIdx.x == 0) counts[blockIdx.x] = c;
}

// exclusive scan of the block counts in place (one CTA); total -> *count
__global__ void __launch_bounds__(1024) k_scan_counts(int32_t* __restrict__ counts, int nblocks, int64_t* __restrict__ count) {
  __shared__ int32_t warp_sums[32];
  __shared__ int32_t carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int base = 0; base < nblocks; base += 1024) {
    const int i = base + threadIdx.x;
    const int32_t x = i < nblocks ? counts[i] : 0;
    int32_t s = x;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int32_t y = __shfl_up_sync(kFull, s, o); if (lane >= o) s += y; }
    if (lane == 31) warp_sums[warp] = s;
    __syncthreads();
    if (warp == 0) {
      int32_t w = warp_sums[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const int32_t y = __shfl_up_sync(kFull, w, o); if (lane >= o) w += y; }
      warp_sums[lane] = w;
    }
    __syncthreads();
    const int32_t carry = carry_s;
    const int32_t incl = s + (warp > 0 ? warp_sums[warp - 1] : 0);
    if (i < nblocks) counts[i] = carry + incl - x;
    __syncthreads();
    if (threadIdx.x == 1023) carry_s = carry + incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) *count = carry_s;
}

__global__ void __launch_bounds__(kCompBlock) k_flag_scatter(const uint8_t* __restrict__ flags, int64_t n, const int32_t* __restrict__ offsets,
                                                             int64_t* __restrict__ idx_out, int64_t cap) {
  __shared__ int32_t warp_base[32];
  const int64_t i = (int64_t)blockIdx.x * kCompBlock + threadIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool f = i < n && flags[i];
  const unsigned b = __ballot_sync(kFull, f);
  if (lane == 0) warp_base[warp] = __popc(b);
  __syncthreads();
  if (warp == 0) {
    int32_t w = warp_base[lane], s = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int32_t y = __shfl_up_sync(kFull, s, o); if (lane >= o) s += y; }
    warp_base[lane] = s - w;
  }
  __syncthreads();
  if (f) {
    const int64_t pos = (int64_t)offsets[blockIdx.x] + warp_base[warp] + __popc(b & ((1u << lane) - 1u));
    if (pos < cap) idx_out[pos] = i;
  }
}

// ray pre-filter (src/Mapper.py:607-621): keep = min_axis(max_pair((bound - o) / d)) >= gt_depth, float64
__global__ void __launch_bounds__(256) k_ray_prefilter(const float* __restrict__ ro, const float* __restrict__ rd,
                                                       const float* __restrict__ gd, int64_t R, const Bound6 b, uint8_t* __restrict__ keep) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  double t = 0.0;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const double o = (double)ro[3 * r + a], d = (double)rd[3 * r + a];
    const double t0 = __ddiv_rn(__dsub_rn(b.v[2 * a], o), d), t1 = __ddiv_rn(__dsub_rn(b.v[2 * a + 1], o), d);
    // torch.max / torch.min propagate NaN (0/0 when a ray starts on a face and runs along it)
    double m = (t0 != t0 || t1 != t1) ? __longlong_as_double(0x7ff8000000000000LL) : fmax(t0, t1);
    if (a == 0) t = m;
    else t = (t != t || m != m) ? __longlong_as_double(0x7ff8000000000000LL) : fmin(t, m);
  }
  keep[r] = (t >= (double)gd[r]) ? 1 : 0;
}

// pixels of a crop with depth > thresh (src/Tracker.py:206-226): flag per crop-flattened pixel
__global__ void __launch_bounds__(256) k_depth_flags(const float* __restrict__ depth, int W, int H0, int W0, int Hc, int Wc, float thresh,
                                                     uint8_t* __restrict__ flags) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)Hc * Wc) return;
  const int y = (int)(i / Wc), x = (int)(i % Wc);
  flags[i] = depth[(int64_t)(H0 + y) * W + (W0 + x)] > thresh ? 1 : 0;
}

// ---------------------------------------------------------------------------------------------
// keyframe overlap selection (src/Mapper.py:267-333)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_overlap_points(const float* __restrict__ ro, const float* __restrict__ rd, const float* __restrict__ gd,
                                                        const float* __restrict__ t_vals, int64_t R, int S, float* __restrict__ verts) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= R * S) return;
  const int64_t r = i / S;
  const float t = t_vals[i % S], d = gd[r];
  const float near = __fmul_rn(d, 0.8f), far = __fadd_rn(d, 0.5f);
  const float z = __fadd_rn(__fmul_rn(near, __fsub_rn(1.0f, t)), __fmul_rn(far, t));
#pragma unroll
  for (int a = 0; a < 3; ++a) verts[3 * i + a] = __fadd_rn(ro[3 * r + a], __fmul_rn(rd[3 * r + a], z));
}

__global__ void __launch_bounds__(256) k_overlap_count(const float* __restrict__ verts, int64_t n, const float* __restrict__ w2c, double fx,
                                                       double fy, double cx, double cy, int H, int W, int edge, int32_t* __restrict__ counts) {
  __shared__ Cam16 cam;
  if (threadIdx.x < 16) cam.m[threadIdx.x] = w2c[16 * blockIdx.y + threadIdx.x];
  __syncthreads();
  int c = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const Proj p = project(cam, verts[3 * i], verts[3 * i + 1], verts[3 * i + 2], fx, fy, cx, cy, 1e-5);
    const bool m = (p.u < (float)(W - edge)) && (p.u > (float)edge) && (p.v < (float)(H - edge)) && (p.v > (float)edge) && (p.z < 0.0);
    c += m ? 1 : 0;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(kFull, c, o);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(counts + blockIdx.y, c);
}

// ---------------------------------------------------------------------------------------------
// Mesher.point_masks (src/utils/Mesher.py:127-196), keyframe branch.  Pass 1 (depth test only): per keyframe
// the maximum of the bilinear depth samples over the chunk; pass 2: seen / forecast per point.
// ---------------------------------------------------------------------------------------------
// F.grid_sample(bilinear, padding zeros, align_corners=True) of a (H,W) image at pixel (u,v) after the reference's
// normalise / unnormalise round trip (Mesher.py:153-154; ATen grid_sampler_unnormalize)
__device__ __forceinline__ float grid_sample_depth(const float* __restrict__ img, int H, int W, float u, float v) {
  const float gx = __fsub_rn(__fmul_rn(__fdiv_rn(u, (float)(W - 1)), 2.0f), 1.0f);
  const float gy = __fsub_rn(__fmul_rn(__fdiv_rn(v, (float)(H - 1)), 2.0f), 1.0f);
  const float ix = __fmul_rn(__fdiv_rn(__fadd_rn(gx, 1.0f), 2.0f), (float)(W - 1));
  const float iy = __fmul_rn(__fdiv_rn(__fadd_rn(gy, 1.0f), 2.0f), (float)(H - 1));
  const float x0f = floorf(ix), y0f = floorf(iy);
  const float x1f = x0f + 1.0f, y1f = y0f + 1.0f;
  // ATen's vectorised CPU kernel (GridSamplerKernel.cpp, ApplyGridSample bilinear): w = x - x_w, e = 1 - w, n = y - y_n, s = 1 - n
  const float tx = __fsub_rn(ix, x0f), ty = __fsub_rn(iy, y0f);
  const float ex = __fsub_rn(1.0f, tx), sy = __fsub_rn(1.0f, ty);
  const float nw = __fmul_rn(ex, sy), ne = __fmul_rn(tx, sy), sw = __fmul_rn(ex, ty), se = __fmul_rn(tx, ty);
  auto tap = [&](float yf, float xf) -> float {
    if (!(xf >= 0.0f && xf <= (float)(W - 1) && yf >= 0.0f && yf <= (float)(H - 1))) return 0.0f;
    return __ldg(img + (int64_t)(int)yf * W + (int)xf);
  };
  float out = __fmul_rn(tap(y0f, x0f), nw);
  out = __fadd_rn(out, __fmul_rn(tap(y0f, x1f), ne));
  out = __fadd_rn(out, __fmul_rn(tap(y1f, x0f), sw));
  out = __fadd_rn(out, __fmul_rn(tap(y1f, x1f), se));
  return out;
}

struct PMaskArgs {
  const float* pts; int64_t n;
  const float* w2c;            // (K,16) device
  const float* const* depth;   // K device pointers to (H,W) images (device array)
  const float* kf_max;         // (K) max(depth)*1.1 per keyframe (depth_test == 0), float32
  unsigned int* samp_max;      // (K) float bits of max(depth_sample) (depth_test == 1)
  int K, H, W, depth_test;
  float fx, fy, cx, cy;
  uint8_t* seen; uint8_t* forecast;
};

__global__ void __launch_bounds__(256) k_pmask_sample_max(const PMaskArgs a) {
  __shared__ Cam16 cam;
  if (threadIdx.x < 16) cam.m[threadIdx.x] = a.w2c[16 * blockIdx.y + threadIdx.x];
  __syncthreads();
  const float* img = a.depth[blockIdx.y];
  float best = -INFINITY;
  bool nan = false;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += (int64_t)gridDim.x * blockDim.x) {
    const Proj p = project_f32(cam, a.pts[3 * i], a.pts[3 * i + 1], a.pts[3 * i + 2], a.fx, a.fy, a.cx, a.cy);
    const float d = grid_sample_depth(img, a.H, a.W, p.u, p.v);
    nan = nan || (d != d);
    best = fmaxf(best, d);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) best = fmaxf(best, __shfl_xor_sync(kFull, best, o));
  // depth samples are >= 0 (zero padding, non-negative images): unsigned order of the bits = float order
  if ((threadIdx.x & 31) == 0 && best >= 0.0f) atomicMax(a.samp_max + blockIdx.y, __float_as_uint(best));
  (void)nan;
}

__global__ void __launch_bounds__(256) k_pmask(const PMaskArgs a) {
  extern __shared__ float cams[];   // K x 16
  for (int i = threadIdx.x; i < 16 * a.K; i += blockDim.x) cams[i] = a.w2c[i];
  __syncthreads();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.n) return;
  const float x = a.pts[3 * i], y = a.pts[3 * i + 1], z = a.pts[3 * i + 2];
  bool seen = false, fore = false;
  for (int k = 0; k < a.K; ++k) {
    Cam16 cam;
#pragma unroll
    for (int j = 0; j < 16; ++j) cam.m[j] = cams[16 * k + j];
    const Proj p = project_f32(cam, x, y, z, a.fx, a.fy, a.cx, a.cy);
    const bool front = p.z < 0.0;
    bool cs = (p.u < (float)a.W) && (p.u > 0.0f) && (p.v < (float)a.H) && (p.v > 0.0f) && front;
    bool cf = (p.u < (float)(a.W + 1000)) && (p.u > -1000.0f) && (p.v < (float)(a.H + 1000)) && (p.v > -1000.0f) && front;
    const float pd = -p.cz;
    if (a.depth_test) {
      const float mx = __uint_as_float(a.samp_max[k]);
      cf = cf && (pd < mx);
      if (cs) {
        const float ds = grid_sample_depth(a.depth[k], a.H, a.W, p.u, p.v);
        cs = (pd < __fadd_rn(ds, 2.4f)) && (__fsub_rn(ds, 2.4f) < pd);
      }
    } else {
      const float mx = a.kf_max[k];
      cf = cf && (pd < mx);
      cs = cs && (pd < mx);
    }
    seen = seen || cs;
    fore = fore || cf;
  }
  a.seen[i] = seen ? 1 : 0;
  a.forecast[i] = (fore && !seen) ? 1 : 0;
}

// ---------------------------------------------------------------------------------------------
// Visualizer residual panels (src/utils/Visualizer.py:60-89)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_vis_residuals(const float* __restrict__ gt_depth, const float* __restrict__ gt_color,
                                                       const double* __restrict__ depth, const float* __restrict__ color, int64_t n,
                                                       double* __restrict__ depth_res, float* __restrict__ gt_color_clip,
                                                       float* __restrict__ color_clip, float* __restrict__ color_res) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float gd = gt_depth[i];
  const bool hole = gd == 0.0f;
  depth_res[i] = hole ? 0.0 : fabs(__dsub_rn((double)gd, depth[i]));
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float g = gt_color[3 * i + c], r = color[3 * i + c];
    const float res = hole ? 0.0f : fabsf(__fsub_rn(g, r));
    gt_color_clip[3 * i + c] = fminf(fmaxf(g, 0.0f), 1.0f);
    color_clip[3 * i + c] = fminf(fmaxf(r, 0.0f), 1.0f);
    color_res[3 * i + c] = fminf(fmaxf(res, 0.0f), 1.0f);
  }
}

// ---------------------------------------------------------------------------------------------
// sparse gradient exchange: rows = 128-byte voxel rows (32 floats) of a dense channels-last gradient
// ---------------------------------------------------------------------------------------------
// Pass 1: per 32-row word a bitmap of rows that hold any non-zero value; per 1024-row block the number of such rows.
// 8 lanes per row (one 128-bit load each), a warp covers 4 rows per step and 32 rows (one bitmap word) in 8 steps.
constexpr int kRowsPerBlock = 1024;   // 32 warps x 32 rows

__global__ void __launch_bounds__(1024) k_rows_bitmap(const float* __restrict__ dense, int64_t V, uint32_t* __restrict__ bitmap,
                                                      int32_t* __restrict__ block_counts) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t row0 = (int64_t)blockIdx.x * kRowsPerBlock + warp * 32;
  uint32_t word = 0;
#pragma unroll
  for (int s = 0; s < 8; ++s) {
    const int64_t r = row0 + 4 * s + (lane >> 3);
    bool nz = false;
    if (r < V) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(dense + r * 32) + (lane & 7));
      nz = (v.x != 0.f) || (v.y != 0.f) || (v.z != 0.f) || (v.w != 0.f);
    }
    const unsigned b = __ballot_sync(kFull, nz);
#pragma unroll
    for (int q = 0; q < 4; ++q)
      if ((b >> (8 * q)) & 0xffu) word |= 1u << (4 * s + q);
  }
  __shared__ int32_t wc[32];
  if (lane == 0 && row0 < V) bitmap[row0 >> 5] = word;
  if (lane == 0) wc[warp] = __popc(word);
  __syncthreads();
  if (warp == 0) {
    int32_t s = wc[lane];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(kFull, s, o);
    if (lane == 0) block_counts[blockIdx.x] = s;
  }
}

// Pass 3 (after k_scan_counts over the block counts): word prefixes and the packed rows
__global__ void __launch_bounds__(1024) k_rows_pack(const float* __restrict__ dense, int64_t V, const uint32_t* __restrict__ bitmap,
                                                    const int32_t* __restrict__ block_offsets, uint32_t* __restrict__ prefix,
                                                    float* __restrict__ rows, int64_t cap, int32_t* __restrict__ overflow) {
  __shared__ int32_t wbase[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t row0 = (int64_t)blockIdx.x * kRowsPerBlock + warp * 32;
  const uint32_t word = row0 < V ? bitmap[row0 >> 5] : 0u;
  if (lane == 0) wbase[warp] = __popc(word);
  __syncthreads();
  if (warp == 0) {
    int32_t w = wbase[lane], s = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int32_t y = __shfl_up_sync(kFull, s, o); if (lane >= o) s += y; }
    wbase[lane] = s - w;
  }
  __syncthreads();
  const int64_t base = (int64_t)block_offsets[blockIdx.x] + wbase[warp];
  if (lane == 0 && row0 < V) prefix[row0 >> 5] = (uint32_t)base;
#pragma unroll
  for (int s = 0; s < 8; ++s) {
    const int rr = 4 * s + (lane >> 3);
    if ((word >> rr) & 1u) {
      const int64_t pos = base + __popc(word & ((1u << rr) - 1u));
      if (pos < cap) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(dense + (row0 + rr) * 32) + (lane & 7));
        reinterpret_cast<float4*>(rows + pos * 32)[lane & 7] = v;
      } else if ((lane & 7) == 0) {
        atomicAdd(overflow, 1);
      }
    }
  }
}

// dense[row] = sum over sources (in source order) of that source's packed row, for every row some source holds.
// Sources are the ranks' (bitmap, prefix, rows) triples: slices of an all-gathered buffer or peer pointers.
struct SparseSrc { const uint32_t* bitmap; const uint32_t* prefix; const float* rows; };
constexpr int kMaxSrc = 16;
struct SparseApplyArgs { SparseSrc src[kMaxSrc]; int nsrc; float* dense; int64_t V; int64_t cap; };

// NS = compile-time bound on the number of sources.  Every load of a step (bitmap words, prefixes, then the up-to-NS
// packed rows of one voxel row) is issued before the first add, so that with peer pointers the NVLink round trips
// (~1 us each) of a step overlap instead of queueing behind one another; absent rows contribute an exact +0.
template <int NS>
__global__ void __launch_bounds__(256) k_rows_apply(const SparseApplyArgs a) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwords = (a.V + 31) >> 5;
  for (int64_t w = warp; w < nwords; w += ((int64_t)gridDim.x * blockDim.x) >> 5) {
    uint32_t words[NS], pref[NS];
    uint32_t any = 0;
#pragma unroll
    for (int s = 0; s < NS; ++s) {
      words[s] = s < a.nsrc ? a.src[s].bitmap[w] : 0u;
      any |= words[s];
    }
    if (!any) continue;
#pragma unroll
    for (int s = 0; s < NS; ++s) pref[s] = words[s] ? a.src[s].prefix[w] : 0u;
#pragma unroll 2
    for (int st = 0; st < 8; ++st) {
      const int rr = 4 * st + (lane >> 3);
      if (!((any >> rr) & 1u)) continue;
      float4 v[NS];
#pragma unroll
      for (int s = 0; s < NS; ++s) {
        v[s] = make_float4(0.f, 0.f, 0.f, 0.f);
        if ((words[s] >> rr) & 1u) {
          const int64_t pos = (int64_t)pref[s] + __popc(words[s] & ((1u << rr) - 1u));
          if (pos < a.cap) v[s] = reinterpret_cast<const float4*>(a.src[s].rows + pos * 32)[lane & 7];
        }
      }
      float4 acc = v[0];
#pragma unroll
      for (int s = 1; s < NS; ++s) {
        acc.x = __fadd_rn(acc.x, v[s].x); acc.y = __fadd_rn(acc.y, v[s].y); acc.z = __fadd_rn(acc.z, v[s].z); acc.w = __fadd_rn(acc.w, v[s].w);
      }
      reinterpret_cast<float4*>(a.dense + (w * 32 + rr) * 32)[lane & 7] = acc;
    }
  }
}

// out[i] = sum_s src_s[i] in source order (the dense tail of the exchange: decoder and pose gradients)
struct DenseSumArgs { const float* src[kMaxSrc]; int nsrc; float* out; int64_t n; };
__global__ void __launch_bounds__(256) k_dense_sum(const DenseSumArgs a) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += (int64_t)gridDim.x * blockDim.x) {
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < kMaxSrc; ++k)
      if (k < a.nsrc) s = __fadd_rn(s, a.src[k][i]);
    a.out[i] = s;
  }
}

inline int blocks_for(int64_t n, int per_block, int cap_blocks) {
  int64_t b = (n + per_block - 1) / per_block;
  if (b < 1) b = 1;
  if (b > cap_blocks) b = cap_blocks;
  return (int)b;
}

}  // namespace
}  // namespace pn

using namespace pn;

extern "C" int pn_frustum_mask(const float* xs, int nx, const float* ys, int ny, const float* zs, int nz, const float* w2c_host,
                               const float* cam_o_host, double fx, double fy, double cx, double cy, int H, int W,
                               const float* depth_img, float* scratch_depth, float* scratch_max, uint8_t* mask, void* stream) {
  if (!xs || !ys || !zs || !w2c_host || !cam_o_host || !depth_img || !scratch_depth || !scratch_max || !mask || nx <= 0 || ny <= 0 ||
      nz <= 0) {
    set_error("pn_frustum_mask: null pointer or empty grid");
    return 1;
  }
  FrustumArgs a;
  a.xs = xs; a.ys = ys; a.zs = zs; a.nx = nx; a.ny = ny; a.nz = nz;
  for (int i = 0; i < 16; ++i) a.w2c.m[i] = w2c_host[i];
  for (int i = 0; i < 3; ++i) a.cam_o[i] = cam_o_host[i];
  a.fx = fx; a.fy = fy; a.cx = cx; a.cy = cy; a.H = H; a.W = W; a.depth = depth_img;
  a.vox_depth = scratch_depth; a.max_bits = reinterpret_cast<unsigned int*>(scratch_max); a.mask = mask;
  cudaStream_t st = (cudaStream_t)stream;
  if (cudaMemsetAsync(scratch_max, 0, 4, st) != cudaSuccess) { set_error("pn_frustum_mask: memset failed"); return 1; }
  const int grid = blocks_for((int64_t)nx * ny * nz, 256, 4 * sm_count());
  k_frustum_depth<<<grid, 256, 0, st>>>(a);
  if (launch_status("k_frustum_depth")) return 1;
  k_frustum_mask<<<grid, 256, 0, st>>>(a);
  return launch_status("k_frustum_mask");
}

extern "C" int pn_adam_step(const void* table, int ntensors, int64_t nblocks, const double* lr, int32_t* step, float* hyper,
                            double beta1, double beta2, double eps, void* stream) {
  if (!table || !lr || !step || !hyper || ntensors <= 0 || nblocks <= 0 || nblocks > 0x7fffffff) {
    set_error("pn_adam_step: bad arguments (ntensors %d, nblocks %lld)", ntensors, (long long)nblocks);
    return 1;
  }
  AdamGroups g;
  g.lr = lr; g.step = step; g.hyper = reinterpret_cast<float2*>(hyper); g.beta1 = beta1; g.beta2 = beta2; g.eps = eps; g.ntensors = ntensors;
  cudaStream_t st = (cudaStream_t)stream;
  k_adam_prepare<<<1, 128, 0, st>>>(reinterpret_cast<const AdamTensor*>(table), g);
  if (launch_status("k_adam_prepare")) return 1;
  k_adam<<<(unsigned)nblocks, 256, 0, st>>>(reinterpret_cast<const AdamTensor*>(table), g);
  return launch_status("k_adam");
}

// shared by the three compaction entry points: flags -> idx_out[0..count), *count.  scratch: int32[(n+1023)/1024 + 1]
static int compact_flags(const uint8_t* flags, int64_t n, int32_t* scratch, int64_t* idx_out, int64_t cap, int64_t* count, cudaStream_t st) {
  const int64_t nb = (n + kCompBlock - 1) / kCompBlock;
  if (nb > 0x7fffffff) { set_error("compaction: too many elements"); return 1; }
  if (n == 0) return cudaMemsetAsync(count, 0, 8, st) == cudaSuccess ? 0 : 1;
  k_flag_counts<<<(unsigned)nb, kCompBlock, 0, st>>>(flags, n, scratch);
  if (launch_status("k_flag_counts")) return 1;
  k_scan_counts<<<1, 1024, 0, st>>>(scratch, (int)nb, count);
  if (launch_status("k_scan_counts")) return 1;
  k_flag_scatter<<<(unsigned)nb, kCompBlock, 0, st>>>(flags, n, scratch, idx_out, cap);
  return launch_status("k_flag_scatter");
}

extern "C" int pn_compact_flags(const uint8_t* flags, int64_t n, int32_t* scratch, int64_t* idx_out, int64_t cap, int64_t* count,
                                void* stream) {
  if (!count || n < 0 || (n > 0 && (!flags || !scratch || !idx_out))) { set_error("pn_compact_flags: null pointer"); return 1; }
  return compact_flags(flags, n, scratch, idx_out, cap, count, (cudaStream_t)stream);
}

extern "C" int pn_ray_prefilter(const float* rays_o, const float* rays_d, const float* gt_depth, int64_t R, const double* bound,
                                uint8_t* keep, void* stream) {
  if (R == 0) return 0;
  if (!rays_o || !rays_d || !gt_depth || !bound || !keep || R < 0) { set_error("pn_ray_prefilter: null pointer"); return 1; }
  k_ray_prefilter<<<(unsigned)((R + 255) / 256), 256, 0, (cudaStream_t)stream>>>(rays_o, rays_d, gt_depth, R, make_bound(bound), keep);
  return launch_status("k_ray_prefilter");
}

extern "C" int pn_depth_pixel_flags(const float* depth_img, int W, int H0, int H1, int W0, int W1, float thresh, uint8_t* flags,
                                    void* stream) {
  if (!depth_img || !flags || H1 <= H0 || W1 <= W0) { set_error("pn_depth_pixel_flags: null pointer or empty crop"); return 1; }
  const int64_t n = (int64_t)(H1 - H0) * (W1 - W0);
  k_depth_flags<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(depth_img, W, H0, W0, H1 - H0, W1 - W0, thresh, flags);
  return launch_status("k_depth_flags");
}

extern "C" int pn_overlap_points(const float* rays_o, const float* rays_d, const float* gt_depth, const float* t_vals, int64_t R,
                                 int n_samples, float* verts, void* stream) {
  if (!rays_o || !rays_d || !gt_depth || !t_vals || !verts || n_samples <= 0) { set_error("pn_overlap_points: null pointer"); return 1; }
  if (R == 0) return 0;
  const int64_t n = R * n_samples;
  k_overlap_points<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(rays_o, rays_d, gt_depth, t_vals, R, n_samples, verts);
  return launch_status("k_overlap_points");
}

extern "C" int pn_keyframe_overlap(const float* verts, int64_t n, const float* w2c, int K, double fx, double fy, double cx, double cy,
                                   int H, int W, int edge, int32_t* counts, void* stream) {
  if (!verts || !w2c || !counts || K <= 0 || K > 65535) { set_error("pn_keyframe_overlap: null pointer or bad keyframe count"); return 1; }
  cudaStream_t st = (cudaStream_t)stream;
  if (cudaMemsetAsync(counts, 0, 4 * (size_t)K, st) != cudaSuccess) { set_error("pn_keyframe_overlap: memset failed"); return 1; }
  if (n == 0) return 0;
  const int gx = blocks_for(n, 256, 64);
  k_overlap_count<<<dim3(gx, K), 256, 0, st>>>(verts, n, w2c, fx, fy, cx, cy, H, W, edge, counts);
  return launch_status("k_overlap_count");
}

extern "C" int pn_point_masks(const float* pts, int64_t n, const float* w2c, const float* const* depth_ptrs, const float* kf_max, int K,
                              int H, int W, float fx, float fy, float cx, float cy, int depth_test, float* scratch_max, uint8_t* seen,
                              uint8_t* forecast, void* stream) {
  if (!pts || !w2c || !seen || !forecast || K <= 0 || K > 512 || (depth_test && (!depth_ptrs || !scratch_max)) || (!depth_test && !kf_max)) {
    set_error("pn_point_masks: null pointer or bad keyframe count");
    return 1;
  }
  if (n == 0) return 0;
  PMaskArgs a;
  a.pts = pts; a.n = n; a.w2c = w2c; a.depth = depth_ptrs; a.kf_max = kf_max; a.samp_max = reinterpret_cast<unsigned int*>(scratch_max);
  a.K = K; a.H = H; a.W = W; a.depth_test = depth_test; a.fx = fx; a.fy = fy; a.cx = cx; a.cy = cy; a.seen = seen; a.forecast = forecast;
  cudaStream_t st = (cudaStream_t)stream;
  if (depth_test) {
    if (cudaMemsetAsync(scratch_max, 0, 4 * (size_t)K, st) != cudaSuccess) { set_error("pn_point_masks: memset failed"); return 1; }
    k_pmask_sample_max<<<dim3(blocks_for(n, 256, 256), K), 256, 0, st>>>(a);
    if (launch_status("k_pmask_sample_max")) return 1;
  }
  k_pmask<<<(unsigned)((n + 255) / 256), 256, 64 * (size_t)K, st>>>(a);
  return launch_status("k_pmask");
}

extern "C" int pn_vis_residuals(const float* gt_depth, const float* gt_color, const double* depth, const float* color, int64_t n,
                                double* depth_res, float* gt_color_clip, float* color_clip, float* color_res, void* stream) {
  if (!gt_depth || !gt_color || !depth || !color || !depth_res || !gt_color_clip || !color_clip || !color_res) {
    set_error("pn_vis_residuals: null pointer");
    return 1;
  }
  if (n == 0) return 0;
  k_vis_residuals<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(gt_depth, gt_color, depth, color, n, depth_res,
                                                                                 gt_color_clip, color_clip, color_res);
  return launch_status("k_vis_residuals");
}

extern "C" int pn_sparse_rows_pack(const float* dense, int64_t V, uint32_t* bitmap, uint32_t* prefix, float* rows, int64_t cap,
                                   int32_t* scratch, int64_t* count, int32_t* overflow, void* stream) {
  if (!dense || !bitmap || !prefix || !rows || !scratch || !count || !overflow || V <= 0) { set_error("pn_sparse_rows_pack: null pointer"); return 1; }
  const int64_t nb = (V + kRowsPerBlock - 1) / kRowsPerBlock;
  cudaStream_t st = (cudaStream_t)stream;
  k_rows_bitmap<<<(unsigned)nb, 1024, 0, st>>>(dense, V, bitmap, scratch);
  if (launch_status("k_rows_bitmap")) return 1;
  k_scan_counts<<<1, 1024, 0, st>>>(scratch, (int)nb, count);
  if (launch_status("k_scan_counts")) return 1;
  k_rows_pack<<<(unsigned)nb, 1024, 0, st>>>(dense, V, bitmap, scratch, prefix, rows, cap, overflow);
  return launch_status("k_rows_pack");
}

extern "C" int pn_sparse_rows_apply(float* dense, int64_t V, int nsrc, const uint32_t* const* bitmaps, const uint32_t* const* prefixes,
                                    const float* const* rows, int64_t cap, void* stream) {
  if (!dense || !bitmaps || !prefixes || !rows || nsrc <= 0 || nsrc > kMaxSrc || V <= 0) {
    set_error("pn_sparse_rows_apply: bad arguments (nsrc %d, at most %d sources)", nsrc, kMaxSrc);
    return 1;
  }
  SparseApplyArgs a;
  for (int s = 0; s < kMaxSrc; ++s) a.src[s] = SparseSrc{nullptr, nullptr, nullptr};
  for (int s = 0; s < nsrc; ++s) a.src[s] = SparseSrc{bitmaps[s], prefixes[s], rows[s]};
  a.nsrc = nsrc; a.dense = dense; a.V = V; a.cap = cap;
  const int64_t nwords = (V + 31) / 32;
  const int grid = blocks_for(nwords, 8, 8 * sm_count());
  cudaStream_t st = (cudaStream_t)stream;
  if (nsrc <= 2) k_rows_apply<2><<<grid, 256, 0, st>>>(a);
  else if (nsrc <= 4) k_rows_apply<4><<<grid, 256, 0, st>>>(a);
  else if (nsrc <= 8) k_rows_apply<8><<<grid, 256, 0, st>>>(a);
  else k_rows_apply<kMaxSrc><<<grid, 256, 0, st>>>(a);
  return launch_status("k_rows_apply");
}

extern "C" int pn_dense_sum(float* out, int64_t n, int nsrc, const float* const* srcs, void* stream) {
  if (!out || !srcs || nsrc <= 0 || nsrc > kMaxSrc) { set_error("pn_dense_sum: bad arguments"); return 1; }
  if (n == 0) return 0;
  DenseSumArgs a;
  for (int s = 0; s < kMaxSrc; ++s) a.src[s] = s < nsrc ? srcs[s] : nullptr;
  a.nsrc = nsrc; a.out = out; a.n = n;
  k_dense_sum<<<blocks_for(n, 256, 4 * sm_count()), 256, 0, (cudaStream_t)stream>>>(a);
  return launch_status("k_dense_sum");
}
