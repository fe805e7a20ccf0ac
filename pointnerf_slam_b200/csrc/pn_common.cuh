// Shared device helpers for the sm_100a ray-rendering kernels.
//
// Geometry (sample points, bound test, coordinate normalisation) follows the
// reference's mixed float32/float64 arithmetic op by op with explicitly
// rounded intrinsics (no FMA contraction), so that the in-bound mask and the
// voxel a sample falls into are bit-identical to the reference's torch path:
//   points      src/utils/Renderer.py:177-179   (float64: o + d*z)
//   bound mask  src/utils/Renderer.py:43-46     (strict, in the points' dtype)
//   normalise   src/common.py:269-284           (in the points' dtype, then .float())
//   trilinear   F.grid_sample(align_corners=True, padding_mode='border'),
//               src/conv_onet/models/decoder.py:168-175
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "pnslam.h"

namespace pn {

constexpr int kThreads = 256;          // threads per CTA of the per-sample kernels
constexpr int kLdc = kThreads + 1;     // padded row of the per-CTA feature tile
constexpr unsigned kFull = 0xffffffffu;

void set_error(const char* fmt, ...);
int launch_status(const char* what);   // cudaGetLastError -> 0 / 1 (+message)
int sm_count();
// tensor-core weight gradients (pn_wgrad_tc.cu): 0 ok, 1 error, -1 not applicable
int launch_wgrad_tc(int64_t N, int c_dim, const float* H, const float* C, const float* E, const float* GA, const float* GH,
                    float* const* W, float* const* b, float* const* Wc, float* const* bc, cudaStream_t st);
// c_dim 32: all parameter gradients in one kernel; GA is rebuilt from GH and the forward's ReLU bits
int launch_wgrad_tc32(int64_t N, int n_out, const float* H, const float* C, const float* E, const float* GH, const float* GARG,
                      const float* GO, const float* P32, const uint32_t* relu_bits, float* const* W, float* const* b,
                      float* const* Wc, float* const* bc, float* Wo, float* bo, float* B, cudaStream_t st);

// tensor-core GEMM of the iMAP* MLP (pn_imap_tc.cu): 0 ok, 1 error, -1 not applicable.  ep: 0 store, 1 bias+relu,
// 2 relu mask (aux), 3 split-K atomics (a_rowsum: optional, += sum_k A(m,k): the bias gradient rides in the weight-gradient GEMM)
int tc_gemm(int ep, const float* A, int64_t sam, int64_t sak, const float* B, int64_t sbk, int64_t sbn, float* C, int64_t ldc,
            int64_t M, int N, int64_t K, const float* bias, const float* aux, int splitk, float* a_rowsum, cudaStream_t st);

// ---- programmatic dependent launch ------------------------------------------------------------
// The persistent decoder kernels start with a prologue that does not depend on the previous kernel
// of the stream (weights -> shared memory, TMEM allocation).  Launched with the programmatic-
// serialization attribute, a kernel is scheduled as soon as every CTA of its predecessor has called
// pdl_launch_dependents() (they do so at once) and an SM frees up, runs its prologue while the
// predecessor drains, and calls pdl_wait() -- which returns when the predecessor has completed and
// its writes are visible -- before touching anything the predecessor may have produced.  After a
// kernel that never triggers (any other kernel) the attribute changes nothing.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, args...);
}

// One decoder result -> raw[n] according to out_mode (pnslam.h).  `force`: the point lies outside the mask bound.
template <int NOUT>
__device__ __forceinline__ void store_raw(float* __restrict__ raw, int64_t n, const float (&out)[4], int out_mode, bool force) {
  float* r = raw + 4 * n;
  if (NOUT == 4) {
    if (out_mode == PN_OUT_SET_RGB) {
      *reinterpret_cast<float2*>(r) = make_float2(out[0], out[1]);
      r[2] = out[2];
    } else {
      *reinterpret_cast<float4*>(r) = make_float4(out[0], out[1], out[2], force ? 100.f : out[3]);
    }
  } else if (out_mode == PN_OUT_SET_ALL) {
    *reinterpret_cast<float4*>(r) = make_float4(0.f, 0.f, 0.f, force ? 100.f : out[0]);
  } else {
    const float w = (out_mode == PN_OUT_ADD_W) ? r[3] + out[0] : out[0];
    r[3] = force ? 100.f : w;
  }
}

struct Bound6 {  // [lo_x hi_x lo_y hi_y lo_z hi_z]
  double v[6];
};
inline Bound6 make_bound(const double* b) {
  Bound6 r;
  for (int i = 0; i < 6; ++i) r.v[i] = b ? b[i] : 0.0;
  return r;
}

struct GridDev {
  const float* data;
  int D, H, W;
};
inline GridDev make_grid(const pn_grid* g) {
  GridDev r{nullptr, 1, 1, 1};
  if (g) { r.data = g->data; r.D = g->D; r.H = g->H; r.W = g->W; }
  return r;
}

// ---------------------------------------------------------------------------
// sample point, bound mask, normalised coordinate
// ---------------------------------------------------------------------------
struct Sample {
  float pf[3];   // float32 point fed to the Fourier embedding (p.float())
  float xn[3];   // normalised coordinate in [-1,1] (float32, after .float())
  bool inside;   // strictly inside mask bound
};

__device__ __forceinline__ void load_sample(const pn_points& ps, int64_t n, const Bound6& nb, const Bound6& mb,
                                            Sample& s) {
  if (ps.pts32) {  // float32 points: everything in float32 with the bound cast to float32
    bool in = true;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const float p = ps.pts32[3 * n + a];
      s.pf[a] = p;
      in = in && (p < (float)mb.v[2 * a + 1]) && (p > (float)mb.v[2 * a]);
      const float ext = (float)__dsub_rn(nb.v[2 * a + 1], nb.v[2 * a]);
      const float t = __fdiv_rn(__fsub_rn(p, (float)nb.v[2 * a]), ext);
      s.xn[a] = __fsub_rn(__fmul_rn(t, 2.0f), 1.0f);
    }
    s.inside = in;
    return;
  }
  double p[3];
  if (ps.pts64) {
#pragma unroll
    for (int a = 0; a < 3; ++a) p[a] = ps.pts64[3 * n + a];
  } else {
    const int64_t r = n / ps.S;
    const double z = ps.z[n];
#pragma unroll
    for (int a = 0; a < 3; ++a)
      p[a] = __dadd_rn((double)ps.rays_o[3 * r + a], __dmul_rn((double)ps.rays_d[3 * r + a], z));
  }
  bool in = true;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    s.pf[a] = (float)p[a];
    in = in && (p[a] < mb.v[2 * a + 1]) && (p[a] > mb.v[2 * a]);
    const double t = __ddiv_rn(__dsub_rn(p[a], nb.v[2 * a]), __dsub_rn(nb.v[2 * a + 1], nb.v[2 * a]));
    s.xn[a] = (float)__dsub_rn(__dmul_rn(t, 2.0), 1.0);
  }
  s.inside = in;
}

// d(xn)/d(p) factor of the normalisation, applied to a float32 gradient.
__device__ __forceinline__ float norm_grad(const pn_points& ps, const Bound6& nb, int a, float g) {
  if (ps.pts32) return __fdiv_rn(__fmul_rn(g, 2.0f), (float)__dsub_rn(nb.v[2 * a + 1], nb.v[2 * a]));
  return (float)__ddiv_rn(__dmul_rn((double)g, 2.0), __dsub_rn(nb.v[2 * a + 1], nb.v[2 * a]));
}

// grid_sampler_unnormalize(align_corners=True): ((x+1)/2)*(size-1), float32.  The division by two is
// written as a multiplication by 0.5: exact either way, so the result is bit-identical, without the
// IEEE division sequence.
__device__ __forceinline__ float unnormalise(float xn, int size) {
  return __fmul_rn(__fmul_rn(__fadd_rn(xn, 1.0f), 0.5f), (float)(size - 1));
}

// One trilinear cell in ATen's corner order (tnw,tne,tsw,tse,bnw,bne,bsw,bse):
// corner c has x-offset c&1, y-offset (c>>1)&1, z-offset c>>2.
struct Cell {
  int64_t base;        // ((z0*H + y0)*W + x0) * 32
  float wx[2], wy[2], wz[2];  // [0] = weight of the low corner ((x0+1)-ix), [1] = ix-x0
  unsigned ok;         // bit c set if corner c is inside the grid
  float gm[3];         // (size-1)/2 if the coordinate was not clipped else 0
};

__device__ __forceinline__ Cell make_cell(float ux, float uy, float uz, int W, int H, int D) {
  Cell c;
  const float mx = (float)(W - 1), my = (float)(H - 1), mz = (float)(D - 1);
  const float ix = fminf(mx, fmaxf(ux, 0.0f));
  const float iy = fminf(my, fmaxf(uy, 0.0f));
  const float iz = fminf(mz, fmaxf(uz, 0.0f));
  c.gm[0] = (ux <= 0.0f || ux >= mx) ? 0.0f : 0.5f * mx;
  c.gm[1] = (uy <= 0.0f || uy >= my) ? 0.0f : 0.5f * my;
  c.gm[2] = (uz <= 0.0f || uz >= mz) ? 0.0f : 0.5f * mz;
  const float fx0 = floorf(ix), fy0 = floorf(iy), fz0 = floorf(iz);
  const int x0 = (int)fx0, y0 = (int)fy0, z0 = (int)fz0;
  c.wx[1] = __fsub_rn(ix, fx0); c.wx[0] = __fsub_rn(fx0 + 1.0f, ix);
  c.wy[1] = __fsub_rn(iy, fy0); c.wy[0] = __fsub_rn(fy0 + 1.0f, iy);
  c.wz[1] = __fsub_rn(iz, fz0); c.wz[0] = __fsub_rn(fz0 + 1.0f, iz);
  const bool x1 = (x0 + 1) < W, y1 = (y0 + 1) < H, z1 = (z0 + 1) < D;
  unsigned ok = 1u;
  ok |= x1 ? 2u : 0u;
  ok |= y1 ? 4u : 0u;
  ok |= (x1 && y1) ? 8u : 0u;
  if (z1) ok |= ok << 4;
  c.ok = ok;
  c.base = (((int64_t)z0 * H + y0) * W + x0) * 32;
  return c;
}

__device__ __forceinline__ float corner_weight(const Cell& c, int k) {
  return __fmul_rn(__fmul_rn(c.wx[k & 1], c.wy[(k >> 1) & 1]), c.wz[k >> 2]);
}

__device__ __forceinline__ int64_t corner_offset(int k, int W, int H) {
  return ((int64_t)(k & 1) + (int64_t)((k >> 1) & 1) * W + (int64_t)(k >> 2) * W * H) * 32;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}

// sin / cos of a Fourier argument (branch-free; |x| up to ~1e5 rad, far beyond any p.B of a scene): two-constant Cody-Waite
// reduction to [-pi, pi] followed by the SFU approximation.  Absolute error <= ~5e-7, an
// order of magnitude below the float32 rounding noise of the argument p.B itself (>= 1e-5
// at |x| ~ 100), at a fifth of the instructions of libdevice sinf.
__device__ __forceinline__ float reduce_2pi(float x) {
  const float n = rintf(x * 0.15915494309189535f);
  float r = fmaf(n, -6.2831854820251465f, x);
  return fmaf(n, 1.7484555e-7f, r);
}
__device__ __forceinline__ float fourier_sin(float x) { return __sinf(reduce_2pi(x)); }
__device__ __forceinline__ float fourier_cos(float x) { return __cosf(reduce_2pi(x)); }

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }

// vector reduction into global memory: red.global.add.v4.f32 (sm_90+)
__device__ __forceinline__ void red_add_v4(float* addr, float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}

}  // namespace pn
