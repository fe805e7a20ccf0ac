"""CPU oracle for the differentiable ray-rendering path.  TEST INFRASTRUCTURE ONLY.

This module restates, with plain torch ops on the CPU, the algorithm of the
reference's hot path (thua919/pointNeRF-SLAM, a NICE-SLAM fork).  It exists to
*check* the CUDA product path; nothing under ``pointnerf-slam_b200/`` may import
it.  Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` use it.

Parity status: PINNED.  ``oracle/pin_against_reference.py`` imports the
reference's own modules from ``/root/reference`` (possible only in the build
container), runs both on identical seeded inputs and requires bit-equal
outputs/gradients on the CPU; it also writes the golden vectors under
``tests/golden/`` that ``tests/test_oracle_golden.py`` replays everywhere.

The restatement is functional: decoders are evaluated from a flat
``state_dict``-style mapping that uses the reference's parameter names
(``middle_decoder.pts_linears.0.weight`` ...), so reference checkpoints load
unchanged.  Each function cites the reference lines it follows
(paths relative to the reference root).

Floating-point conventions that matter for parity and are reproduced here:
  * ``bound`` is float64 and its upper edge carries a float32 rounding
    (src/NICE_SLAM.py:208-213);
  * z-values, sample points, coordinate normalisation, depth and variance are
    float64 whenever a depth image is given (src/utils/Renderer.py:98-179);
  * points are cast to float32 only right before the trilinear gather and the
    Fourier embedding (src/conv_onet/models/decoder.py:171,189).
"""
from __future__ import annotations

import math
from typing import Dict, Mapping, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor
GRID_KEYS = ("grid_coarse", "grid_middle", "grid_fine", "grid_color")
STAGES = ("coarse", "middle", "fine", "color")
EMBED = 93  # Fourier mapping size, src/conv_onet/models/decoder.py:129


# --------------------------------------------------------------------------
# scene set-up (orchestrator duties; src/NICE_SLAM.py:200-315)
# --------------------------------------------------------------------------
def scene_bound(bound_cfg, scale: float = 1.0, bound_divisible: float = 0.32) -> Tensor:
    """float64 (3,2) bound whose upper edge is enlarged to a multiple of
    ``bound_divisible``.  The int32-tensor * python-float product is float32,
    which is why the upper edge carries float32 rounding
    (src/NICE_SLAM.py:208-213)."""
    b = torch.from_numpy(np.array(bound_cfg, dtype=np.float64) * scale)
    steps = ((b[:, 1] - b[:, 0]) / bound_divisible).int() + 1
    b[:, 1] = steps * bound_divisible + b[:, 0]
    return b


def grid_shapes(bound: Tensor, grid_len: Mapping[str, float], coarse_enlarge: float = 2,
                coarse: bool = True) -> Dict[str, Tuple[int, int, int]]:
    """(Z, Y, X) voxel counts per level; x and z are swapped relative to the
    bound order and the division truncates (src/NICE_SLAM.py:276-311)."""
    extent = bound[:, 1] - bound[:, 0]
    out = {}
    for level in ("coarse", "middle", "fine", "color"):
        if level == "coarse":
            if not coarse:
                continue
            n = list(map(int, (extent * coarse_enlarge / grid_len[level]).tolist()))
        else:
            n = list(map(int, (extent / grid_len[level]).tolist()))
        out["grid_" + level] = (n[2], n[1], n[0])
    return out


def init_grids(bound: Tensor, grid_len: Mapping[str, float], c_dim: int = 32,
               coarse_enlarge: float = 2, coarse: bool = True,
               generator: Optional[torch.Generator] = None) -> Dict[str, Tensor]:
    """Feature grids (1, C, Z, Y, X) ~ N(0, 0.01^2), fine ~ N(0, 1e-4^2)
    (src/NICE_SLAM.py:283-313)."""
    grids = {}
    for key, (z, y, x) in grid_shapes(bound, grid_len, coarse_enlarge, coarse).items():
        std = 0.0001 if key == "grid_fine" else 0.01
        grids[key] = torch.zeros(1, c_dim, z, y, x).normal_(mean=0, std=std, generator=generator)
    return grids


def decoder_bounds(bound: Tensor, coarse_enlarge: float = 2) -> Dict[str, Tensor]:
    """Per-decoder bound; the coarse level sees the enlarged box
    (src/NICE_SLAM.py:216-221)."""
    return {"coarse": bound * coarse_enlarge, "middle": bound, "fine": bound, "color": bound}


def _xavier(out_dim: int, in_dim: int, gain: float, g) -> Tensor:
    a = gain * math.sqrt(6.0 / (in_dim + out_dim))
    return (torch.rand(out_dim, in_dim, generator=g) * 2 - 1) * a


def init_nice_state(c_dim: int = 32, hidden: int = 32, coarse: bool = True, seed: int = 0) -> Dict[str, Tensor]:
    """Random NICE decoder parameters with the reference's names, shapes and
    init family (xavier-uniform weights, zero pts/output biases, B = 25*randn;
    src/conv_onet/models/decoder.py:17-22,75-79,124-159,235-245,298-310).
    The fc_c layers are ``nn.Linear`` (kaiming-uniform family); a uniform of
    the same scale is used here.  Not the reference's RNG stream -- parity runs
    share one state dict between both sides instead."""
    g = torch.Generator().manual_seed(seed)
    relu_gain = math.sqrt(2.0)
    sd: Dict[str, Tensor] = {}
    if coarse:
        pre = "coarse_decoder."
        ins = [hidden, hidden, hidden, hidden + c_dim, hidden]
        for i, k in enumerate(ins):
            sd[f"{pre}pts_linears.{i}.weight"] = _xavier(hidden, k, relu_gain, g)
            sd[f"{pre}pts_linears.{i}.bias"] = torch.zeros(hidden)
        sd[pre + "output_linear.weight"] = _xavier(1, hidden, 1.0, g)
        sd[pre + "output_linear.bias"] = torch.zeros(1)
    for name, cd, n_out in (("middle", c_dim, 1), ("fine", 2 * c_dim, 1), ("color", c_dim, 4)):
        pre = f"{name}_decoder."
        for i in range(5):
            lim = 1.0 / math.sqrt(cd)
            sd[f"{pre}fc_c.{i}.weight"] = (torch.rand(hidden, cd, generator=g) * 2 - 1) * lim
            sd[f"{pre}fc_c.{i}.bias"] = (torch.rand(hidden, generator=g) * 2 - 1) * lim
        sd[pre + "embedder._B"] = torch.randn(3, EMBED, generator=g) * 25
        ins = [EMBED, hidden, hidden, hidden + EMBED, hidden]
        for i, k in enumerate(ins):
            sd[f"{pre}pts_linears.{i}.weight"] = _xavier(hidden, k, relu_gain, g)
            # the reference zero-initialises these; small random values make
            # the bias path observable in parity tests
            sd[f"{pre}pts_linears.{i}.bias"] = (torch.rand(hidden, generator=g) - 0.5) * 0.1
        sd[pre + "output_linear.weight"] = _xavier(n_out, hidden, 1.0, g)
        sd[pre + "output_linear.bias"] = (torch.rand(n_out, generator=g) - 0.5) * 0.1
    return sd


def init_imap_state(hidden: int = 256, n_blocks: int = 4, seed: int = 0) -> Dict[str, Tensor]:
    """Random iMAP* single-MLP parameters (src/conv_onet/config.py:28-32)."""
    g = torch.Generator().manual_seed(seed)
    sd = {"embedder._B": torch.randn(3, EMBED, generator=g) * 25}
    ins = [EMBED] + [hidden] * (n_blocks - 1)
    for i, k in enumerate(ins):
        sd[f"pts_linears.{i}.weight"] = _xavier(hidden, k, math.sqrt(2.0), g)
        sd[f"pts_linears.{i}.bias"] = (torch.rand(hidden, generator=g) - 0.5) * 0.1
    sd["output_linear.weight"] = _xavier(4, hidden, 1.0, g)
    sd["output_linear.bias"] = (torch.rand(4, generator=g) - 0.5) * 0.1
    return sd


# --------------------------------------------------------------------------
# camera / ray generation (src/common.py:74-176, 248-266)
# --------------------------------------------------------------------------
def quaternion_to_rotation(quad: Tensor) -> Tensor:
    """(B,4) [w,x,y,z] -> (B,3,3); src/common.py:137-160."""
    qr, qi, qj, qk = quad[:, 0], quad[:, 1], quad[:, 2], quad[:, 3]
    two_s = 2.0 / (quad * quad).sum(-1)
    rows = [
        [1 - two_s * (qj ** 2 + qk ** 2), two_s * (qi * qj - qk * qr), two_s * (qi * qk + qj * qr)],
        [two_s * (qi * qj + qk * qr), 1 - two_s * (qi ** 2 + qk ** 2), two_s * (qj * qk - qi * qr)],
        [two_s * (qi * qk - qj * qr), two_s * (qj * qk + qi * qr), 1 - two_s * (qi ** 2 + qj ** 2)],
    ]
    return torch.stack([torch.stack(r, -1) for r in rows], -2)


def camera_from_tensor(cam: Tensor) -> Tensor:
    """[qw,qx,qy,qz,tx,ty,tz] -> (3,4) (or batched); src/common.py:163-176."""
    single = cam.dim() == 1
    x = cam.unsqueeze(0) if single else cam
    rt = torch.cat([quaternion_to_rotation(x[:, :4]), x[:, 4:, None]], 2)
    return rt[0] if single else rt


def rays_from_pixels(i: Tensor, j: Tensor, c2w: Tensor, fx, fy, cx, cy) -> Tuple[Tensor, Tensor]:
    """Un-normalised world rays through pixel centres; src/common.py:74-89."""
    dirs = torch.stack([(i - cx) / fx, -(j - cy) / fy, -torch.ones_like(i)], -1)
    dirs = dirs.reshape(-1, 1, 3)
    rays_d = torch.sum(dirs * c2w[:3, :3], -1)
    rays_o = c2w[:3, -1].expand(rays_d.shape)
    return rays_o, rays_d


def pixel_lattice(H0: int, H1: int, W0: int, W1: int) -> Tuple[Tensor, Tensor]:
    """Flattened (i=column, j=row) coordinates of the crop, row-major over the
    crop; src/common.py:117-120 with 97-98."""
    i, j = torch.meshgrid(torch.linspace(W0, W1 - 1, W1 - W0), torch.linspace(H0, H1 - 1, H1 - H0),
                          indexing="ij")
    return i.t().reshape(-1), j.t().reshape(-1)


def get_samples(H0, H1, W0, W1, n, fx, fy, cx, cy, c2w, depth, color,
                indices: Optional[Tensor] = None):
    """n random rays of the crop with their depth/colour.  ``indices`` (flat,
    row-major over the crop) may be supplied; otherwise drawn with
    ``torch.randint`` exactly like src/common.py:99.  Returns
    (rays_o, rays_d, depth, color, indices); src/common.py:92-134."""
    i, j = pixel_lattice(H0, H1, W0, W1)
    i, j = i.to(depth.device), j.to(depth.device)
    if indices is None:
        indices = torch.randint(i.shape[0], (n,), device=depth.device)
    indices = indices.clamp(0, i.shape[0])
    d = depth[H0:H1, W0:W1].reshape(-1)[indices]
    c = color[H0:H1, W0:W1].reshape(-1, 3)[indices]
    rays_o, rays_d = rays_from_pixels(i[indices], j[indices], c2w, fx, fy, cx, cy)
    return rays_o, rays_d, d, c, indices


def get_rays(H, W, fx, fy, cx, cy, c2w) -> Tuple[Tensor, Tensor]:
    """All rays of an image, (H,W,3) each; src/common.py:248-266."""
    i, j = torch.meshgrid(torch.linspace(0, W - 1, W), torch.linspace(0, H - 1, H), indexing="ij")
    i, j = i.t().to(c2w.device), j.t().to(c2w.device)
    dirs = torch.stack([(i - cx) / fx, -(j - cy) / fy, -torch.ones_like(i)], -1).reshape(H, W, 1, 3)
    rays_d = torch.sum(dirs * c2w[:3, :3], -1)
    return c2w[:3, -1].expand(rays_d.shape), rays_d


# --------------------------------------------------------------------------
# decoders (src/conv_onet/models/decoder.py)
# --------------------------------------------------------------------------
def normalise_points(p: Tensor, bound: Tensor) -> Tensor:
    """Map to [-1,1] in p's own dtype; src/common.py:269-284."""
    p = p.reshape(-1, 3)
    cols = [((p[:, a] - bound[a, 0]) / (bound[a, 1] - bound[a, 0])) * 2 - 1.0 for a in range(3)]
    return torch.stack(cols, -1).to(p.dtype)


def trilinear_feature(p: Tensor, grid: Tensor, bound: Tensor) -> Tensor:
    """(N,3) points -> (N,C) features, border padding, align_corners;
    src/conv_onet/models/decoder.py:168-175."""
    vgrid = normalise_points(p, bound).unsqueeze(0)[:, :, None, None].float()
    c = F.grid_sample(grid, vgrid, padding_mode="border", align_corners=True, mode="bilinear")
    return c.squeeze(-1).squeeze(-1).transpose(1, 2).squeeze(0)


# The reference forms the Fourier argument with a float32 matmul, `p @ B`, i.e. a 3-term dot product whose evaluation order
# (fused or not, which product first) is whatever the BLAS of the box does.  With |p.B| of hundreds of radians one float32
# ulp of the argument is 1.5e-5..3e-5 rad, so two BLAS builds already disagree by ~1e-4 in a few per cent of the outputs
# (seen between GPU boxes with different host CPUs).  Comparisons that run the oracle on the box's CPU therefore pin the order
# to ONE of the admissible ones -- fma(p_z, B_z, fma(p_y, B_y, p_x * B_x)), the order the CUDA kernels use -- by setting
# EMBED_ORDER = "fma" (see fixed_embed_order()); the default None keeps the reference's own expression, which is what
# oracle/pin_against_reference.py pins bit-equal against /root/reference and what the goldens were minted with.
EMBED_ORDER: Optional[str] = None


class fixed_embed_order:
    """``with fixed_embed_order():`` evaluates the Fourier argument in the fixed fused order (box-independent results)."""

    def __enter__(self):
        global EMBED_ORDER
        self.prev, EMBED_ORDER = EMBED_ORDER, "fma"
        return self

    def __exit__(self, *exc):
        global EMBED_ORDER
        EMBED_ORDER = self.prev
        return False


def _fma32(a: Tensor, b: Tensor, c: Tensor) -> Tensor:
    """float32 fused multiply-add: the product of two float32 is exact in float64, so one float64 addition followed by the
    rounding to float32 reproduces fmaf (up to double rounding in ~2^-29 of the cases)."""
    return (a.double() * b.double() + c.double()).float()


def fourier_embed(p32: Tensor, B: Tensor) -> Tensor:
    """sin(p @ B); src/conv_onet/models/decoder.py:26-30."""
    if EMBED_ORDER == "fma":
        t = p32[:, 0:1] * B[0:1, :]
        t = _fma32(p32[:, 1:2], B[1:2, :], t)
        t = _fma32(p32[:, 2:3], B[2:3, :], t)
        return torch.sin(t)
    return torch.sin(p32 @ B)


def grid_mlp(sd: Mapping[str, Tensor], prefix: str, p: Tensor, feat: Optional[Tensor],
             skips: Sequence[int] = (2,), squeeze_out: bool = True) -> Tensor:
    """Fourier-embedded MLP with a feature term added after every ReLU
    (``feat`` None for iMAP*); src/conv_onet/models/decoder.py:189-203."""
    emb = fourier_embed(p.reshape(-1, 3).float(), sd[prefix + "embedder._B"])
    h = emb
    i = 0
    while f"{prefix}pts_linears.{i}.weight" in sd:
        h = F.relu(F.linear(h, sd[f"{prefix}pts_linears.{i}.weight"], sd[f"{prefix}pts_linears.{i}.bias"]))
        if feat is not None:
            h = h + F.linear(feat, sd[f"{prefix}fc_c.{i}.weight"], sd[f"{prefix}fc_c.{i}.bias"])
        if i in skips:
            h = torch.cat([emb, h], -1)
        i += 1
    out = F.linear(h, sd[prefix + "output_linear.weight"], sd[prefix + "output_linear.bias"])
    return out.squeeze(-1) if (squeeze_out and out.shape[-1] == 1) else out


def coarse_mlp(sd: Mapping[str, Tensor], prefix: str, feat: Tensor) -> Tensor:
    """Feature-only MLP (no xyz input); src/conv_onet/models/decoder.py:262-274."""
    h = feat
    for i in range(5):
        h = F.relu(F.linear(h, sd[f"{prefix}pts_linears.{i}.weight"], sd[f"{prefix}pts_linears.{i}.bias"]))
        if i == 2:
            h = torch.cat([feat, h], -1)
    return F.linear(h, sd[prefix + "output_linear.weight"], sd[prefix + "output_linear.bias"]).squeeze(-1)


def middle_occ(sd, p, grids, bounds):
    return grid_mlp(sd, "middle_decoder.", p, trilinear_feature(p, grids["grid_middle"], bounds["middle"]))


def fine_occ(sd, p, grids, bounds):
    c = trilinear_feature(p, grids["grid_fine"], bounds["fine"])
    with torch.no_grad():  # middle feature enters the fine decoder detached, decoder.py:184-186
        cm = trilinear_feature(p, grids["grid_middle"], bounds["fine"])
    return grid_mlp(sd, "fine_decoder.", p, torch.cat([c, cm], 1))


def color_raw(sd, p, grids, bounds):
    return grid_mlp(sd, "color_decoder.", p, trilinear_feature(p, grids["grid_color"], bounds["color"]))


def nice_forward(sd: Mapping[str, Tensor], p: Tensor, grids: Mapping[str, Tensor],
                 bounds: Mapping[str, Tensor], stage: str) -> Tensor:
    """(N,3) -> (N,4) [r,g,b,occupancy logit] per stage;
    src/conv_onet/models/decoder.py:312-342."""
    if stage == "coarse":
        occ = coarse_mlp(sd, "coarse_decoder.", trilinear_feature(p, grids["grid_coarse"], bounds["coarse"]))
        raw = torch.zeros(occ.shape[0], 4, device=occ.device)
        raw[..., -1] = occ
    elif stage == "middle":
        occ = middle_occ(sd, p, grids, bounds)
        raw = torch.zeros(occ.shape[0], 4, device=occ.device)
        raw[..., -1] = occ
    elif stage == "fine":
        f = fine_occ(sd, p, grids, bounds)
        raw = torch.zeros(f.shape[0], 4, device=f.device)
        raw[..., -1] = f + middle_occ(sd, p, grids, bounds)
    elif stage == "color":
        f = fine_occ(sd, p, grids, bounds)
        raw = color_raw(sd, p, grids, bounds)
        raw[..., -1] = f + middle_occ(sd, p, grids, bounds)
    else:
        raise ValueError(stage)
    return raw


def imap_forward(sd: Mapping[str, Tensor], p: Tensor) -> Tensor:
    """iMAP* single MLP, hidden 256, 4 blocks, no skips, 4 outputs."""
    return grid_mlp(sd, "", p, None, skips=(), squeeze_out=False)


# --------------------------------------------------------------------------
# renderer (src/utils/Renderer.py)
# --------------------------------------------------------------------------
class Scene:
    """Everything ``render_batch_ray`` needs besides the rays."""

    def __init__(self, sd, grids, bound, *, nice=True, occupancy=True, n_samples=32, n_surface=16,
                 n_importance=0, lindisp=False, perturb=0.0, coarse_enlarge=2,
                 points_batch_size=500000, ray_batch_size=100000):
        self.sd, self.grids, self.bound = sd, grids, bound
        self.bounds = decoder_bounds(bound, coarse_enlarge)
        self.nice, self.occupancy = nice, occupancy
        self.n_samples, self.n_surface, self.n_importance = n_samples, n_surface, n_importance
        self.lindisp, self.perturb = lindisp, perturb
        self.points_batch_size, self.ray_batch_size = points_batch_size, ray_batch_size


def eval_points(scene: Scene, p: Tensor, stage: str = "color") -> Tensor:
    """Occupancy/colour of points; logit forced to 100 outside the (strict)
    bound; src/utils/Renderer.py:23-61."""
    b = scene.bound
    outs = []
    for pi in torch.split(p, scene.points_batch_size):
        inside = ((pi[:, 0] < b[0][1]) & (pi[:, 0] > b[0][0]) & (pi[:, 1] < b[1][1]) & (pi[:, 1] > b[1][0])
                  & (pi[:, 2] < b[2][1]) & (pi[:, 2] > b[2][0]))
        if scene.nice:
            ret = nice_forward(scene.sd, pi, scene.grids, scene.bounds, stage)
        else:
            ret = imap_forward(scene.sd, pi)
        ret[~inside, 3] = 100
        outs.append(ret)
    return torch.cat(outs, 0)


def composite(raw: Tensor, z_vals: Tensor, rays_d: Tensor, occupancy: bool):
    """raw (N,S,4) -> depth, variance, rgb, weights; src/common.py:204-245.
    (The reference overwrites raw[...,3] in place with the sigmoid; a copy is
    used here, the values are the same.)"""
    dists = (z_vals[..., 1:] - z_vals[..., :-1]).float()
    dists = torch.cat([dists, torch.full_like(dists[..., :1], 1e10)], -1)
    dists = dists * torch.norm(rays_d[..., None, :], dim=-1)
    rgb = raw[..., :-1]
    if occupancy:
        alpha = torch.sigmoid(10 * raw[..., -1])
    else:
        alpha = 1.0 - torch.exp(-F.relu(raw[..., -1]) * dists)
    ones = torch.ones((alpha.shape[0], 1), device=alpha.device)
    trans = torch.cumprod(torch.cat([ones, (1.0 - alpha + 1e-10).float()], -1).float(), -1)[:, :-1]
    weights = alpha.float() * trans
    rgb_map = torch.sum(weights[..., None] * rgb, -2)
    depth = torch.sum(weights * z_vals, -1)
    dev = z_vals - depth.unsqueeze(-1)
    var = torch.sum(weights * dev * dev, dim=1)
    return depth, var, rgb_map, weights


def sample_pdf(bins: Tensor, weights: Tensor, n: int, det: bool = True, u: Optional[Tensor] = None) -> Tensor:
    """Inverse-CDF importance samples; src/common.py:19-63."""
    weights = weights + 1e-5
    pdf = weights / torch.sum(weights, -1, keepdim=True)
    cdf = torch.cumsum(pdf, -1)
    cdf = torch.cat([torch.zeros_like(cdf[..., :1]), cdf], -1)
    if u is None:
        if det:
            u = torch.linspace(0.0, 1.0, steps=n).expand(list(cdf.shape[:-1]) + [n])
        else:
            u = torch.rand(list(cdf.shape[:-1]) + [n])
    u = u.to(cdf.device).contiguous()
    inds = torch.searchsorted(cdf, u, right=True)
    below = torch.clamp(inds - 1, min=0)
    above = torch.clamp(inds, max=cdf.shape[-1] - 1)
    pair = torch.stack([below, above], -1)
    shape = [pair.shape[0], pair.shape[1], cdf.shape[-1]]
    cdf_g = torch.gather(cdf.unsqueeze(1).expand(shape), 2, pair)
    bins_g = torch.gather(bins.unsqueeze(1).expand(shape), 2, pair)
    denom = cdf_g[..., 1] - cdf_g[..., 0]
    denom = torch.where(denom < 1e-5, torch.ones_like(denom), denom)
    t = (u - cdf_g[..., 0]) / denom
    return bins_g[..., 0] + t * (bins_g[..., 1] - bins_g[..., 0])


def ray_z_values(scene: Scene, rays_o: Tensor, rays_d: Tensor, gt_depth: Optional[Tensor],
                 t_rand: Optional[Tensor] = None) -> Tensor:
    """Stratified + depth-guided sample depths, sorted; src/utils/Renderer.py:82-175."""
    n_samples, n_surface = scene.n_samples, scene.n_surface
    if gt_depth is None:
        n_surface = 0
        near = 0.01
    else:
        gt_depth = gt_depth.reshape(-1, 1)
        near = gt_depth.repeat(1, n_samples) * 0.01
    with torch.no_grad():
        o = rays_o.clone().detach().unsqueeze(-1)
        d = rays_d.clone().detach().unsqueeze(-1)
        t = (scene.bound.unsqueeze(0).to(o.device) - o) / d
        far_bb = torch.min(torch.max(t, dim=2)[0], dim=1)[0].unsqueeze(-1)
        far_bb += 0.01
    far = torch.clamp(far_bb, 0, (gt_depth * 1.2).max()) if gt_depth is not None else far_bb
    z_surface = None
    if n_surface > 0:
        has = gt_depth > 0
        ts = torch.linspace(0.0, 1.0, steps=n_surface).double().to(rays_o.device)
        g = gt_depth[has].unsqueeze(-1).repeat(1, n_surface)
        near_surface = 0.95 * g * (1.0 - ts) + 1.05 * g * ts
        z_surface = torch.zeros(gt_depth.shape[0], n_surface).to(rays_o.device).double()
        has = has.squeeze(-1)
        z_surface[has, :] = near_surface
        z_surface[~has, :] = 0.001 * (1.0 - ts) + torch.max(gt_depth) * ts
    tv = torch.linspace(0.0, 1.0, steps=n_samples).to(rays_o.device)
    if not scene.lindisp:
        z = near * (1.0 - tv) + far * tv
    else:
        z = 1.0 / (1.0 / near * (1.0 - tv) + 1.0 / far * tv)
    if scene.perturb > 0.0:
        mids = 0.5 * (z[..., 1:] + z[..., :-1])
        upper = torch.cat([mids, z[..., -1:]], -1)
        lower = torch.cat([z[..., :1], mids], -1)
        if t_rand is None:
            t_rand = torch.rand(z.shape)
        z = lower + (upper - lower) * t_rand.to(z.device)
    if n_surface > 0:
        z, _ = torch.sort(torch.cat([z, z_surface.double()], -1), -1)
    return z


def render_batch_ray(scene: Scene, rays_d: Tensor, rays_o: Tensor, stage: str,
                     gt_depth: Optional[Tensor] = None, t_rand: Optional[Tensor] = None,
                     return_extras: bool = False):
    """depth (N,), variance (N,), colour (N,3); src/utils/Renderer.py:63-203."""
    n = rays_o.shape[0]
    z = ray_z_values(scene, rays_o, rays_d, gt_depth, t_rand)
    pts = rays_o[..., None, :] + rays_d[..., None, :] * z[..., :, None]
    raw = eval_points(scene, pts.reshape(-1, 3), stage).reshape(n, z.shape[1], -1)
    depth, var, rgb, w = composite(raw, z, rays_d, scene.occupancy)
    if scene.n_importance > 0:
        mid = 0.5 * (z[..., 1:] + z[..., :-1])
        zs = sample_pdf(mid, w[..., 1:-1], scene.n_importance, det=(scene.perturb == 0.0)).detach()
        z, _ = torch.sort(torch.cat([z, zs], -1), -1)
        pts = rays_o[..., None, :] + rays_d[..., None, :] * z[..., :, None]
        raw = eval_points(scene, pts.reshape(-1, 3), stage).reshape(n, z.shape[1], -1)
        depth, var, rgb, w = composite(raw, z, rays_d, scene.occupancy)
    if return_extras:
        return depth, var, rgb, {"z_vals": z, "raw": raw, "weights": w}
    return depth, var, rgb


def render_img(scene: Scene, c2w: Tensor, H, W, fx, fy, cx, cy, stage: str, gt_depth: Tensor):
    """Full-frame, no-grad, chunked; src/utils/Renderer.py:205-260."""
    with torch.no_grad():
        ro, rd = get_rays(H, W, fx, fy, cx, cy, c2w)
        ro, rd = ro.reshape(-1, 3), rd.reshape(-1, 3)
        gd = gt_depth.reshape(-1)
        ds, us, cs = [], [], []
        for s in range(0, rd.shape[0], scene.ray_batch_size):
            e = s + scene.ray_batch_size
            d, u, c = render_batch_ray(scene, rd[s:e], ro[s:e], stage, gd[s:e])
            ds.append(d.double()); us.append(u.double()); cs.append(c)
        return (torch.cat(ds).reshape(H, W), torch.cat(us).reshape(H, W), torch.cat(cs).reshape(H, W, 3))


def regulation(scene: Scene, rays_d: Tensor, rays_o: Tensor, gt_depth: Tensor, stage: str = "color",
               t_rand: Optional[Tensor] = None) -> Tensor:
    """Densities of jittered samples in [0, 0.85*depth]; src/utils/Renderer.py:263-301."""
    g = gt_depth.reshape(-1, 1).repeat(1, scene.n_samples)
    tv = torch.linspace(0.0, 1.0, steps=scene.n_samples).to(rays_o.device)
    z = 0.0 * (1.0 - tv) + (g * 0.85) * tv
    mids = 0.5 * (z[..., 1:] + z[..., :-1])
    upper = torch.cat([mids, z[..., -1:]], -1)
    lower = torch.cat([z[..., :1], mids], -1)
    if t_rand is None:
        t_rand = torch.rand(z.shape)
    z = lower + (upper - lower) * t_rand.to(z.device)
    pts = rays_o[..., None, :] + rays_d[..., None, :] * z[..., :, None]
    return eval_points(scene, pts.reshape(-1, 3), stage)[:, -1]


# --------------------------------------------------------------------------
# caller losses (define the gradients fed to the backward)
# --------------------------------------------------------------------------
def tracking_loss(depth, var, color, gt_depth, gt_color, w_color=0.5, handle_dynamic=True, depth_supervision=True):
    """src/Tracker.py:306-330 (depth_supervision=False: the fork's colour-only branch, :313-318)."""
    var = var.detach()
    if handle_dynamic:
        tmp = torch.abs(gt_depth - depth) / torch.sqrt(var + 1e-10)
        mask = (tmp < 10 * tmp.median()) & (gt_depth > 0)
    else:
        mask = gt_depth > 0
    if not depth_supervision:
        return torch.abs(gt_color - color)[mask].sum()
    loss = (torch.abs(gt_depth - depth) / torch.sqrt(var + 1e-10))[mask].sum()
    return loss + w_color * torch.abs(gt_color - color)[mask].sum()


def mapping_loss(depth, color, gt_depth, gt_color, stage, w_color=0.2, nice=True, depth_supervision=True):
    """src/Mapper.py:628-646 (depth_supervision=False: the fork's colour-only branch, :633-637)."""
    if not depth_supervision:
        return torch.abs(gt_color - color).sum()
    m = gt_depth > 0
    loss = torch.abs(gt_depth[m] - depth[m]).sum()
    if (not nice) or stage == "color":
        loss = loss + w_color * torch.abs(gt_color - color).sum()
    return loss
