/*
 * pnslam.h -- C ABI of the B200-native differentiable ray-rendering path.
 *
 * The reference (thua919/pointNeRF-SLAM, a NICE-SLAM fork) is pure Python: it
 * has no FFI for this path, only Python call signatures.  This
 * header is therefore the boundary a maintainer binds with ctypes; each entry
 * point names the reference lines whose ATen op chain it replaces.  All
 * pointers are DEVICE pointers unless stated; `stream` is a cudaStream_t.
 * Every function returns 0 on success, non-zero on error; pn_last_error()
 * returns a thread-local message.  The library never caches caller pointers
 * across calls; its only process-wide state is the launch counter
 * (pn_launch_count) and the SM reservation (pn_reserve_sms).
 *
 * Layout conventions
 *   - feature grids are channels-last: [Z][Y][X][32] float32 (the memory of a
 *     torch (1,32,Z,Y,X) tensor in torch.channels_last_3d format);
 *   - "planar-4" stash buffers hold an (N x K) float matrix as
 *     [K/4][N] float4, i.e. element (n,k) at ((k/4)*N + n)*4 + k%4, so that
 *     consecutive samples are consecutive 16-byte words;
 *   - decoder parameters are read in place from the torch tensors of the
 *     nn.Module (row-major [out][in]).
 */
#ifndef PNSLAM_H_
#define PNSLAM_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PN_HIDDEN 32 /* hidden width of the NICE decoders, decoder.py:295 */
#define PN_EMBED 93  /* Fourier mapping size, decoder.py:129 */
#define PN_CDIM 32   /* feature channels per grid, nice_slam.yaml:114 */
#define PN_MAX_SAMPLES 128

typedef struct pn_grid {
  const float* data; /* [D][H][W][32] */
  int D, H, W;       /* Z, Y, X voxel counts */
} pn_grid;

/* Parameters of one decoder.MLP (decoder.py:91-203) with hidden 32, 5 blocks,
 * skip after block 2.  in-dims of W: 93,32,32,125,32.  c_dim 32 or 64. */
typedef struct pn_grid_mlp {
  const float* B;     /* embedder._B (3,93) */
  const float* W[5];  /* pts_linears.i.weight */
  const float* b[5];  /* pts_linears.i.bias */
  const float* Wc[5]; /* fc_c.i.weight (32,c_dim) */
  const float* bc[5]; /* fc_c.i.bias */
  const float* Wo;    /* output_linear.weight (n_out,32) */
  const float* bo;    /* output_linear.bias */
  int c_dim;
  int n_out; /* 1 (occupancy) or 4 (colour) */
} pn_grid_mlp;

/* Gradient sinks with the shapes of pn_grid_mlp (accumulated with +=; the
 * caller zeroes them).  Any pointer may be NULL (that gradient is skipped). */
typedef struct pn_grid_mlp_grad {
  float* B;
  float* W[5];
  float* b[5];
  float* Wc[5];
  float* bc[5];
  float* Wo;
  float* bo;
} pn_grid_mlp_grad;

/* Parameters of decoder.MLP_no_xyz (decoder.py:206-274): in-dims 32,32,32,64,32. */
typedef struct pn_coarse_mlp {
  const float* W[5];
  const float* b[5];
  const float* Wo;
  const float* bo;
} pn_coarse_mlp;

typedef struct pn_coarse_mlp_grad {
  float* W[5];
  float* b[5];
  float* Wo;
  float* bo;
} pn_coarse_mlp_grad;

/* Where the sample points of an evaluation come from.  Exactly one of
 * pts64 / pts32 / (rays_o, rays_d, z) is set.  In ray mode sample n is
 * (ray n / S, k = n % S) and p = o + d * z[ray][k] evaluated in float64
 * (Renderer.py:177-179). */
typedef struct pn_points {
  const double* pts64; /* (N,3) */
  const float* pts32;  /* (N,3) */
  const float* rays_o; /* (R,3) */
  const float* rays_d; /* (R,3) */
  const double* z;     /* (R,S) */
  int S;
  int64_t N; /* number of sample points */
} pn_points;

/* Stash written by a forward evaluation for its backward. */
typedef struct pn_stash {
  uint32_t* relu_bits; /* [5][N]: bit j of word i = (pre-activation j of block i > 0) */
  float* H;            /* [5] planar-4 (N x 32): block outputs h_0..h_4; NULL unless weight grads wanted */
  float* C;            /* planar-4 (N x c_dim): gathered features; NULL unless weight grads wanted */
  float* E;            /* planar-4 (N x 96): Fourier embedding (cols 93..95 zero); NULL unless weight grads wanted */
} pn_stash;

/* Scratch written by the input-gradient kernel for the weight-gradient kernel. */
typedef struct pn_wscratch {
  float* GA;   /* [5] planar-4 (N x 32): gradient at block pre-activations.  For c_dim 32 the backward leaves it
                * unwritten: pn_grid_mlp_wgrad rebuilds it from GH and stash->relu_bits */
  float* GH;   /* [5] planar-4 (N x 32): gradient at block outputs */
  float* GARG; /* planar-4 (N x 96): gradient at the Fourier arguments */
  float* P32;  /* [3][N]: float32 points */
  float* GO;   /* (N,4): gradient at the decoder outputs actually consumed (masked) */
} pn_wscratch;

enum { PN_OUT_SET_ALL = 0, /* raw[n] = (0,0,0,occ) or the 4 colour-decoder outputs */
       PN_OUT_SET_W = 1,   /* raw[n].w = occ, rgb kept */
       PN_OUT_ADD_W = 2,   /* raw[n].w += occ */
       PN_OUT_SET_RGB = 3 }; /* raw[n].xyz = the colour-decoder outputs (n_out 4), w untouched.  SET_W / ADD_W touch only
                              * the 4th float and SET_RGB only the first three, so the colour pass may run concurrently
                              * (another stream) with the occupancy passes that fill w of the same buffer */

const char* pn_last_error(void);
int pn_version(void);
/* number of kernels this process has launched through the library (for bench accounting) */
long long pn_launch_count(void);
/* The decoder kernels are persistent, one CTA per SM with nearly all of its shared memory, so a kernel
 * of another stream (an NCCL all-reduce of finished gradients) cannot start beside them.  Reserving n
 * SMs makes every later launch size its grid for (SM count - n); returns the previous value.
 * No reference counterpart (the reference is single-GPU); used by the data-parallel mapper. */
int pn_reserve_sms(int n);

/* -------- pose and ray generation (src/common.py:74-176, 248-266) -------- */

/* get_camera_from_tensor, common.py:163-176: cam (B,7) [qw qx qy qz tx ty tz] -> c2w (B,3,4). */
int pn_camera_from_tensor_fwd(const float* cam, int batch, float* c2w, void* stream);
/* its vector-Jacobian product: g_c2w (B,3,4) -> g_cam (B,7) (overwritten). */
int pn_camera_from_tensor_bwd(const float* cam, const float* g_c2w, int batch, float* g_cam, void* stream);

/* get_samples without the randint, common.py:92-134: for flat crop indices idx (n,)
 * gather depth/colour and build rays.  c2w: 12+ floats, row stride `c2w_ld` (4 for
 * (3,4)/(4,4) row-major).  color_img float32 or float64 (H,W,3) selected by color_is_f64;
 * color_out has the same dtype. */
int pn_sample_rays_fwd(const int64_t* idx, int n, int H0, int W0, int Wc, int W,
                       float fx, float fy, float cx, float cy, const float* c2w, int c2w_ld,
                       const float* depth_img, const void* color_img, int color_is_f64,
                       float* rays_o, float* rays_d, float* depth_out, void* color_out, void* stream);
/* get_rays, common.py:248-266: all H*W rays, row-major. */
int pn_image_rays_fwd(int H, int W, float fx, float fy, float cx, float cy, const float* c2w, int c2w_ld,
                      float* rays_o, float* rays_d, void* stream);
/* VJP of both: g_c2w (3,4) += sum_n [ g_rays_d[n] (x) dir_n | g_rays_o[n] ].  Pixels given
 * either by idx (+crop) or, if idx == NULL, as the full H0..,W0.. Wc-wide lattice of n pixels.
 * Block reduction by warp shuffles, one atomicAdd per block and entry.  g_c2w must be zeroed. */
int pn_rays_bwd(const int64_t* idx, int n, int H0, int W0, int Wc, float fx, float fy, float cx, float cy,
                const float* g_rays_o, const float* g_rays_d, float* g_c2w, void* stream);

/* The Mapper's per-keyframe loop (src/Mapper.py:558-605: get_samples for every keyframe of the window, then cat) as
 * ONE launch: idx (F,n) flat crop indices per keyframe, c2w (F,3,4) device, depth_ptrs / color_ptrs: DEVICE arrays of F
 * device pointers to the keyframes' (H,W) depth and (H,W,3) colour images; outputs are the concatenated (F*n, ...) batch. */
int pn_sample_rays_multi_fwd(const int64_t* idx, int F, int n, int H0, int W0, int Wc, int W, float fx, float fy,
                             float cx, float cy, const float* c2w, const float* const* depth_ptrs,
                             const void* const* color_ptrs, int color_is_f64, float* rays_o, float* rays_d,
                             float* depth_out, void* color_out, void* stream);
/* VJP: g_c2w (F,3,4) += per keyframe, as pn_rays_bwd; must be zeroed. */
int pn_rays_multi_bwd(const int64_t* idx, int F, int n, int H0, int W0, int Wc, float fx, float fy, float cx, float cy,
                      const float* g_rays_o, const float* g_rays_d, float* g_c2w, void* stream);

/* -------- sample placement (src/utils/Renderer.py:82-175) -------- */

/* Depth-guided + stratified z-values, sorted.  gt_depth may be NULL (then n_surface
 * is ignored, near = 0.01, far = box exit).  depth_max: device pointer to max(gt_depth)
 * over the WHOLE batch (all shards).  bound: host array [lo_x hi_x lo_y hi_y lo_z hi_z]
 * float64.  t_vals: device linspace(0,1,n_samples) float32; t_surface: device
 * linspace(0,1,n_surface) as float64 (both made by the host with torch.linspace so that
 * they carry the reference's exact values).  t_rand: optional (R,n_samples) jitter for
 * perturb>0.  z_out: (R, S). */
int pn_ray_zvals(const float* rays_o, const float* rays_d, const float* gt_depth, const float* depth_max,
                 int64_t R, const double* bound, int n_samples, int n_surface, int lindisp,
                 const float* t_vals, const double* t_surface, const float* t_rand, double* z_out,
                 void* stream);
/* sample_pdf (common.py:19-63) + merge-sort (Renderer.py:187-191): z (R,S), weights (R,S)
 * -> z_out (R,S+n_imp).  u_lin: device linspace(0,1,n_imp) (deterministic mode) or
 * u_rand: (R,n_imp) uniforms; exactly one is used (u_rand wins). */
int pn_importance_zvals(const double* z, const float* weights, int64_t R, int S, int n_imp,
                        const float* u_lin, const float* u_rand, double* z_out, void* stream);
/* sample_pdf alone (common.py:19-63): bins (R,nb) f64, weights (R,nb-1) f32 -> out (R,n) f64. */
int pn_sample_pdf(const double* bins, const float* weights, int64_t R, int nb, int n, const float* u_lin,
                  const float* u_rand, double* out, void* stream);
/* regulation sample points (Renderer.py:280-298), all float32: z jittered in [0, 0.85*depth]
 * with t_rand (R,n_samples), pts_out (R*n_samples,3) = o + d*z; z_out (optional) gets z as
 * float64 (R,n_samples) for pn_points_to_rays_bwd. */
int pn_regulation_points(const float* rays_o, const float* rays_d, const float* gt_depth, const float* t_vals,
                         const float* t_rand, int64_t R, int n_samples, float* pts_out, double* z_out,
                         void* stream);

/* -------- decoders (src/conv_onet/models/decoder.py) -------- */

/* One decoder.MLP over N points: trilinear gather (decoder.py:168-175), Fourier
 * embedding (26-30), 5 blocks + output (189-203).  gridA is the decoder's own grid,
 * gridB the middle grid concatenated (detached) when w->c_dim == 64 (182-187).
 * norm_bound / mask_bound: host float64[6].  If apply_mask, points outside
 * mask_bound (strict) get raw.w = 100 (Renderer.py:43-57).  raw: (N,4) float32. */
int pn_grid_mlp_fwd(const pn_points* pts, const pn_grid_mlp* w, const pn_grid* gridA, const pn_grid* gridB,
                    const double* norm_bound, const double* mask_bound, int apply_mask, int out_mode,
                    float* raw, const pn_stash* stash, void* stream);
/* Input-side backward of the same decoder.  g_raw (N,4): this decoder consumes
 * components 0..2 (n_out 4) or component 3 (n_out 1; zeroed outside mask_bound if
 * apply_mask).  Produces (each optional): g_gridA [D][H][W][32] += (warp-aggregated
 * vector atomics), g_pts (N,3) float32 (=/+= per accumulate_pts), and the scratch the
 * weight-gradient kernel needs. */
int pn_grid_mlp_bwd(const pn_points* pts, const pn_grid_mlp* w, const pn_grid* gridA, const pn_grid* gridB,
                    const double* norm_bound, const double* mask_bound, int apply_mask,
                    const float* g_raw, const pn_stash* stash, float* g_gridA, float* g_pts,
                    int accumulate_pts, const pn_wscratch* ws, void* stream);
/* Weight gradients: sum over samples of (gradient x activation) outer products over the stash / scratch written
 * by pn_grid_mlp_fwd / _bwd of the SAME engine; results atomically added into `g` (NULL sinks are skipped).
 * c_dim 32: one warp-specialised tcgen05 kernel for W, b, Wc, bc, Wo, bo and B (needs stash->relu_bits);
 * c_dim 64: tcgen05 GEMM kernel + two small kernels (tiled FFMA GEMMs only when some W / b / Wc / bc sinks are NULL). */
int pn_grid_mlp_wgrad(int64_t N, const pn_grid_mlp* w, const pn_stash* stash, const pn_wscratch* ws,
                      const pn_grid_mlp_grad* g, void* stream);

/* decoder.MLP_no_xyz (coarse level).  Same conventions; norm_bound is the enlarged bound. */
int pn_coarse_mlp_fwd(const pn_points* pts, const pn_coarse_mlp* w, const pn_grid* grid,
                      const double* norm_bound, const double* mask_bound, int apply_mask, int out_mode,
                      float* raw, const pn_stash* stash, void* stream);
int pn_coarse_mlp_bwd(const pn_points* pts, const pn_coarse_mlp* w, const pn_grid* grid,
                      const double* norm_bound, const double* mask_bound, int apply_mask,
                      const float* g_raw, const pn_stash* stash, float* g_grid, float* g_pts,
                      int accumulate_pts, const pn_wscratch* ws, void* stream);
int pn_coarse_mlp_wgrad(int64_t N, const pn_stash* stash, const pn_wscratch* ws,
                        const pn_coarse_mlp_grad* g, void* stream);

/* iMAP* single MLP = decoder.MLP with c_dim 0 (conv_onet/config.py:28-32): Fourier-93 ->
 * n_blocks x [hidden] relu layers (no skips) -> 4 outputs.  Activations are row-major:
 * E (N,96) embedding, H [n_blocks](N,hidden) block outputs (also the relu masks), P32 [3][N]. */
#define PN_IMAP_MAX_BLOCKS 8
typedef struct pn_imap_mlp {
  const float* B;                      /* embedder._B (3,93) */
  const float* W[PN_IMAP_MAX_BLOCKS];  /* pts_linears.i.weight: (hidden,93), then (hidden,hidden) */
  const float* b[PN_IMAP_MAX_BLOCKS];
  const float* Wo;                     /* (4,hidden) */
  const float* bo;
  int hidden, n_blocks;
} pn_imap_mlp;
typedef struct pn_imap_mlp_grad {      /* accumulated (+=), any pointer may be NULL */
  float* B;
  float* W[PN_IMAP_MAX_BLOCKS];
  float* b[PN_IMAP_MAX_BLOCKS];
  float* Wo;
  float* bo;
} pn_imap_mlp_grad;
/* raw (N,4) = MLP(p); component 3 forced to 100 outside mask_bound if apply_mask.
 * E, H are always written (they are the backward's stash); P32 optional. */
int pn_imap_mlp_fwd(const pn_points* pts, const pn_imap_mlp* w, const double* mask_bound, int apply_mask,
                    float* raw, float* E, float* H, float* P32, void* stream);
/* Backward: g_raw (N,4) -> parameter gradients g (optional), g_pts (N,3) += (optional).
 * GA, GB: scratch (N, max(hidden,96)) each; GO: scratch (N,4). */
int pn_imap_mlp_bwd(const pn_points* pts, const pn_imap_mlp* w, const double* mask_bound, int apply_mask,
                    const float* g_raw, const float* E, const float* H, const float* P32, float* GA, float* GB,
                    float* GO, float* g_pts, const pn_imap_mlp_grad* g, void* stream);

/* -------- compositing (src/common.py:204-245) -------- */

/* raw (R,S,4), z (R,S) f64, rays_d (R,3) -> depth (R) f64, var (R) f64, rgb (R,3) f32,
 * weights (R,S) f32 (optional).  One warp per ray, transmittance by warp scan. */
int pn_composite_fwd(const float* raw, const double* z, const float* rays_d, int64_t R, int S, int occupancy,
                     double* depth, double* var, float* rgb, float* weights, void* stream);
/* VJP: g_depth, g_var (f64, optional), g_rgb (f32, optional) -> g_raw (R,S,4).  In density
 * mode alpha depends on |rays_d|; that term is added (+=) into g_rays_d (R,3) if non-NULL. */
int pn_composite_bwd(const float* raw, const double* z, const float* rays_d, int64_t R, int S, int occupancy,
                     const double* g_depth, const double* g_var, const float* g_rgb, float* g_raw,
                     float* g_rays_d, void* stream);
/* VJP of p = o + d*z: g_pts (R*S,3) -> g_rays_o (R,3), g_rays_d (R,3), float64 accumulation. */
int pn_points_to_rays_bwd(const float* g_pts, const double* z, int64_t R, int S,
                          float* g_rays_o, float* g_rays_d, void* stream);

/* -------- loss head of the mapping iteration (src/Mapper.py:628-646) -------- */
/* loss = sum_{gt_depth>0} |gt_depth - depth|  (+ w_color * sum |gt_color - color| when use_color), float64 scalar;
 * also its gradient: g_depth (R) float64, g_color (R,3) float32 (already scaled by w_color).  One CTA with a fixed
 * summation order (deterministic); the colour sum is rounded to float32 before scaling, as in the reference.
 * depth_supervision == 0 selects the fork's colour-only branch (Mapper.py:633-637): loss = sum |gt_color - color|
 * (unweighted), no depth gradient. */
int pn_mapping_loss(const double* depth, const float* color, const float* gt_depth, const float* gt_color, int64_t R,
                    int use_color, float w_color, int depth_supervision, double* loss, double* g_depth, float* g_color,
                    void* stream);

/* Tracker loss head (src/Tracker.py:306-330): tmp = |gt_depth - depth| / sqrt(var + 1e-10) (float64);
 * mask = gt_depth > 0, and with handle_dynamic also tmp < 10 * median(tmp) (torch.median: rank (R-1)/2; a NaN in tmp
 * makes the median NaN and the mask empty, as in torch);
 * loss = sum_mask tmp (+ w_color * sum_mask |gt_color - color| when use_color).  Also its gradient w.r.t. depth
 * (R, float64) and colour (R,3, float32, scaled by w_color); var is treated as detached, as in the reference.
 * depth_supervision == 0: the fork's colour-only branch (Tracker.py:313-318): loss = sum_mask |gt_color - color|.
 * One CTA, deterministic.  R <= 8192: tmp is sorted in shared memory; larger R (the fork tracker can pass every
 * pixel with depth): the median comes from an 8-pass radix select over the bit patterns. */
int pn_tracking_loss(const double* depth, const double* var, const float* color, const float* gt_depth,
                     const float* gt_color, int64_t R, int handle_dynamic, int use_color, float w_color,
                     int depth_supervision, double* loss, double* g_depth, float* g_color, void* stream);

/* tcgen05 self-test: Y (128,32) = X (128,K) . W (32,K)^T through the 3xTF32 tensor-core path
 * (operands in shared memory, accumulator in tensor memory); K multiple of 8, <= 128. */
int pn_tc_selftest(const float* X, const float* W, float* Y, int K, void* stream);
/* Hardware probe: the same product with MN-major (sample-contiguous) operands Xt (K,128), Wt (K,32); swap selects
 * which of LBO / SBO is the stride between core matrices along K.  On sm_100a kind::tf32 returns zeros for both
 * conventions with the no-swizzle layout, which is why the weight-gradient kernels transpose in shared memory. */
int pn_tc_selftest_mn(const float* Xt, const float* Wt, float* Y, int K, int swap, void* stream);

/* ======== callers either side of the rendering path ======== */

/* -------- frustum feature selection (src/Mapper.py:129-200, Mapper.get_mask_from_c2w) --------
 * xs / ys / zs: device float32 axes of the voxel lattice (the host's torch.linspace(bound[a][0], bound[a][1], n),
 * Mapper.py:145-147); w2c_host: HOST float32[16], row-major np.linalg.inv(c2w) (Mapper.py:154); cam_o_host: HOST
 * float32[3] = c2w[:3,3].  depth_img: device (H,W) float32.  The depth of a voxel is cv2.remap(INTER_LINEAR) of the
 * image at its projection (5-bit fixed-point coordinates, zero border); voxels whose remapped depth is 0 take the
 * maximum over all voxels (Mapper.py:180-181).  scratch_depth: nx*ny*nz floats; scratch_max: 1 float.
 * mask: uint8 [nz][ny][nx] (the memory order of the channels-last grid), 1 = optimise this voxel. */
int pn_frustum_mask(const float* xs, int nx, const float* ys, int ny, const float* zs, int nz, const float* w2c_host,
                    const float* cam_o_host, double fx, double fy, double cx, double cy, int H, int W,
                    const float* depth_img, float* scratch_depth, float* scratch_max, uint8_t* mask, void* stream);

/* -------- optimiser step (src/Mapper.py:482-505, 529-536, 657-674; configs/nice_slam.yaml:71-95) --------
 * torch.optim.Adam (no weight decay, no amsgrad) over a table of tensors in ONE launch (+ a one-block kernel that
 * advances the step counts and forms the bias corrections).  `table`: device array of `ntensors` rows
 * {float* p; const float* g; float* m; float* v; const int64_t* idx; int64_t n; int64_t nwork; int32_t row;
 * int32_t group; int64_t block0} (72 bytes each; g == NULL: tensor skipped, as Adam skips parameters without
 * gradient).  idx (optional): ascending indices of the selected voxels -- the frustum mask compacted once per
 * mapping call; only those voxels are visited and only they carry Adam state, exactly like the reference, which
 * optimises val[mask] and writes it back (Mapper.py:413-431, 511-518, 665-674).  nwork = elements to update
 * (n, or selected voxels x channels); row > 0: channels-last grid (element = voxel * row + channel, row = 32);
 * row < 0: NCDHW grid (element = channel * (-row) + voxel, -row = Z*Y*X).  block0: first block of the tensor,
 * blocks of 1024 work elements; nblocks = total.  lr: device double[groups] (the per-stage table, changed by the
 * host between iterations); step: device int32[ntensors], steps taken so far per tensor (advanced for the tensors
 * that have a gradient, as torch.optim.Adam does); hyper: device scratch, 2 floats per tensor.
 * Graph-capturable: no host scalar changes between iterations. */
int pn_adam_step(const void* table, int ntensors, int64_t nblocks, const double* lr, int32_t* step, float* hyper, double beta1,
                 double beta2, double eps, void* stream);

/* -------- ray pre-filter and pixel selection (src/Mapper.py:607-621, src/Tracker.py:206-226, 288-300) -------- */
/* keep[r] = min_axis(max_pair((bound - o)/d)) >= gt_depth[r], float64 like the reference (bound: HOST double[6]). */
int pn_ray_prefilter(const float* rays_o, const float* rays_d, const float* gt_depth, int64_t R, const double* bound,
                     uint8_t* keep, void* stream);
/* flags[y*Wc + x] = depth_img[H0+y][W0+x] > thresh over the crop (np.where(depth > 0.01) of Tracker.select_uv). */
int pn_depth_pixel_flags(const float* depth_img, int W, int H0, int H1, int W0, int W1, float thresh, uint8_t* flags,
                         void* stream);
/* Stable compaction: ascending indices i with flags[i] != 0 -> idx_out[0..min(count,cap)), *count = their number
 * (device int64).  scratch: device int32[(n + 1023)/1024]. */
int pn_compact_flags(const uint8_t* flags, int64_t n, int32_t* scratch, int64_t* idx_out, int64_t cap, int64_t* count,
                     void* stream);

/* -------- keyframe overlap selection (src/Mapper.py:267-333) -------- */
/* verts (R*n_samples,3) float32: points between 0.8*depth and depth+0.5 along each ray (Mapper.py:291-300);
 * t_vals: device linspace(0,1,n_samples) float32. */
int pn_overlap_points(const float* rays_o, const float* rays_d, const float* gt_depth, const float* t_vals, int64_t R,
                      int n_samples, float* verts, void* stream);
/* counts[k] = number of verts that project inside keyframe k's image (margin `edge`) in front of the camera
 * (Mapper.py:302-320).  w2c: device (K,16) float32, the host's np.linalg.inv of every keyframe pose. */
int pn_keyframe_overlap(const float* verts, int64_t n, const float* w2c, int K, double fx, double fy, double cx, double cy,
                        int H, int W, int edge, int32_t* counts, void* stream);

/* -------- dense-render consumers (src/utils/Mesher.py:53-212, src/utils/Visualizer.py:60-89) -------- */
/* Mesher.point_masks over ONE chunk of points (keyframe branch): seen / forecast uint8 (n) (unseen = neither).
 * w2c: device (K,16); depth_ptrs: DEVICE array of K device pointers to (H,W) depth images (depth_test != 0: bilinear
 * grid_sample test against each keyframe's depth, forecast limited by max(depth_sample) over the chunk);
 * kf_max: device float32 (K) = max(depth)*1.1 per keyframe (depth_test == 0).  scratch_max: K floats. */
int pn_point_masks(const float* pts, int64_t n, const float* w2c, const float* const* depth_ptrs, const float* kf_max, int K,
                   int H, int W, float fx, float fy, float cx, float cy, int depth_test, float* scratch_max, uint8_t* seen,
                   uint8_t* forecast, void* stream);
/* Visualizer panels: depth_res = |gt_depth - depth| (0 where gt_depth == 0), colour residual likewise, colours clipped
 * to [0,1].  n pixels; depth float64 as render_img returns it. */
int pn_vis_residuals(const float* gt_depth, const float* gt_color, const double* depth, const float* color, int64_t n,
                     double* depth_res, float* gt_color_clip, float* color_clip, float* color_res, void* stream);

/* -------- sparse gradient exchange of the data-parallel mapper (no reference counterpart: single GPU) --------
 * A mapping iteration touches a few per cent of the voxel rows (32 floats = 128 bytes each) of a gradient grid.
 * pn_sparse_rows_pack turns a dense [V][32] gradient into bitmap (V/32 words, bit r of word w = row 32w+r holds a
 * non-zero), prefix (per word: packed position of its first set row) and the packed rows in ascending row order
 * (at most `cap`; rows beyond it are counted in *overflow, which the caller must find 0).  *count = rows held.
 * scratch: int32[(V + 1023)/1024]. */
int pn_sparse_rows_pack(const float* dense, int64_t V, uint32_t* bitmap, uint32_t* prefix, float* rows, int64_t cap,
                        int32_t* scratch, int64_t* count, int32_t* overflow, void* stream);
/* dense[row] = sum over the nsrc sources, in source order, of that source's packed row, for every row some source
 * holds (other rows are left as they are).  bitmaps / prefixes / rows: HOST arrays of nsrc device pointers (slices
 * of an all-gathered buffer, or peer pointers into the other GPUs' symmetric buffers: the reads then travel over
 * NVLink inside this kernel).  Same source order on every rank => bit-identical sums on every rank. */
int pn_sparse_rows_apply(float* dense, int64_t V, int nsrc, const uint32_t* const* bitmaps, const uint32_t* const* prefixes,
                         const float* const* rows, int64_t cap, void* stream);
/* out[i] = src_0[i] + src_1[i] + ... in source order (decoder / pose gradients riding in the same exchange). */
int pn_dense_sum(float* out, int64_t n, int nsrc, const float* const* srcs, void* stream);

/* -------- k-nearest neural-point feature aggregation (BASELINE.json config 4) --------
 * BUILDER-DEFINED SEMANTICS, NOT REFERENCE PARITY: the reference has no 3-D neural-point aggregation;
 * its nearest code is the 2-D cKDTree radius search of src/frame.py:362-366 / src/search_points.py:122,223,445.
 * Specification: the K = 8 nearest neural points within `radius` of a sample point, weights
 * w_k = 1 / (d_k^2 + eps), blended feature f = sum_k w_k F[i_k] / sum_k w_k (zero when no point lies within the radius).
 * Distances are float32, d2 = ((dx*dx) + (dy*dy)) + (dz*dz) with individually rounded operations; neighbours are ordered
 * by (d2, point index), so the index lists are deterministic and bit-exact against oracle/knn_oracle.py. */
#define PN_KNN_K 8
typedef struct pn_knn_index {
  const int32_t* start; /* [nx*ny*nz + 1]: first slot of every cell in `sorted` (cell = (cz*ny + cy)*nx + cx) */
  const float* sorted;  /* [P] float4 {x, y, z, bit pattern of the point's index}: the points grouped by cell */
  float lo[3];          /* lower corner of the cell lattice */
  float inv_h;          /* 1 / cell edge; cell coordinate = clamp(floor((x - lo) * inv_h), 0, n-1), float32 */
  int nx, ny, nz;
  int P;
} pn_knn_index;
/* Fill index->start and index->sorted from xyz (P,3) float32 (every point must lie inside the lattice: the caller's
 * check).  scratch: int32[2*P + nx*ny*nz]. */
int pn_knn_build(const float* xyz, int P, const pn_knn_index* index, int32_t* scratch, void* stream);
/* idx (N,8) int32: original indices of the nearest points in ascending (d2, index) order, -1 where fewer than 8 lie
 * within the radius; d2 (N,8) float32 (optional; 0 in the empty slots).  radius <= the index's cell edge. */
int pn_knn_query(const pn_points* pts, const pn_knn_index* index, float radius, int32_t* idx, float* d2, void* stream);
/* query + blend in one launch: also out (N,32) float32.  feat: [P][32] float32 rows. */
int pn_knn_aggregate_fwd(const pn_points* pts, const pn_knn_index* index, float radius, float eps, const float* feat,
                         int32_t* idx, float* d2, float* out, void* stream);
/* VJP of the blend for g_out (N,32): g_feat [P][32] += (vector atomics; optional), g_pts (N,3) =/+= (optional;
 * the neighbour SET is piecewise constant in p, so only the weights carry a point gradient). */
int pn_knn_aggregate_bwd(const pn_points* pts, const int32_t* idx, const float* d2, float eps, const float* feat,
                         const float* xyz, const float* g_out, float* g_feat, float* g_pts, int accumulate_pts,
                         void* stream);

/* -------- utilities -------- */
/* (1,32,Z,Y,X) contiguous <-> channels-last [Z][Y][X][32]; `to_channels_last` = 1 or 0. */
int pn_grid_transpose(const float* src, float* dst, int D, int H, int W, int to_channels_last, void* stream);
/* out[0] = max(x[0..n)) ; n >= 1. */
int pn_max_f32(const float* x, int64_t n, float* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PNSLAM_H_ */
