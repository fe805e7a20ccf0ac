"""Drop-in ``Renderer`` (reference: src/utils/Renderer.py) on the CUDA library.

Same constructor, attributes and method signatures as the reference class so
that ``Tracker.py`` / ``Mapper.py`` / ``Visualizer.py`` / ``Mesher.py`` can
switch over unchanged:

    render_batch_ray(c, decoders, rays_d, rays_o, device, stage, gt_depth=None)
    eval_points(p, decoders, c=None, stage='color', device='cuda:0')
    render_img(c, decoders, c2w, device, stage, gt_depth=None)
    regulation(c, decoders, rays_d, rays_o, gt_depth, device, stage='color')

Outputs keep the reference's dtypes (depth / uncertainty float64, colour
float32).  ``decoders`` may be this package's modules or the reference's own
``NICE`` / ``MLP`` instances (parameters are read in place through the same
attribute names).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Dict, Optional, Tuple

import torch

from . import _lib as L
from . import engine as E


@dataclass
class _RayCfg:
    n_samples: int
    n_surface: int
    n_importance: int
    lindisp: bool
    perturb: float
    occupancy: bool
    bound: object
    freeze_map: bool
    consts: Dict[str, torch.Tensor]


def _ray_constants(device, n_samples: int, n_surface: int, n_importance: int) -> Dict[str, torch.Tensor]:
    """linspace tables made on the host exactly as the reference makes them
    (Renderer.py:137-138,157; common.py:33), then uploaded once."""
    out = {"t_vals": torch.linspace(0.0, 1.0, steps=n_samples).to(device)}
    if n_surface > 0:
        out["t_surface"] = torch.linspace(0.0, 1.0, steps=n_surface).double().to(device)
    if n_importance > 0:
        out["u_lin"] = torch.linspace(0.0, 1.0, steps=n_importance).to(device)
    return out


def composite(raw: torch.Tensor, z: torch.Tensor, rays_d: torch.Tensor, occupancy: bool, want_weights: bool):
    """Launch pn_composite_fwd; returns depth, var, rgb, weights."""
    R, S = z.shape
    dev = z.device
    depth = torch.empty(R, dtype=torch.float64, device=dev)
    var = torch.empty(R, dtype=torch.float64, device=dev)
    rgb = torch.empty((R, 3), dtype=torch.float32, device=dev)
    w = torch.empty((R, S), dtype=torch.float32, device=dev) if want_weights else None
    with L.device_guard(dev):
        L.check(L.lib().pn_composite_fwd(C.c_void_p(raw.data_ptr()), C.c_void_p(z.data_ptr()), C.c_void_p(rays_d.data_ptr()),
                                         C.c_int64(R), S, int(occupancy), C.c_void_p(depth.data_ptr()),
                                         C.c_void_p(var.data_ptr()), C.c_void_p(rgb.data_ptr()), C.c_void_p(L.ptr(w)),
                                         C.c_void_p(L.stream_ptr(dev))), "pn_composite_fwd")
    return depth, var, rgb, w


def batch_depth_max(gt: torch.Tensor) -> torch.Tensor:
    """max over the batch of gt_depth as a 1-element device tensor."""
    out = torch.empty(1, dtype=torch.float32, device=gt.device)
    with L.device_guard(gt.device):
        L.check(L.lib().pn_max_f32(C.c_void_p(gt.data_ptr()), C.c_int64(gt.numel()), C.c_void_p(out.data_ptr()),
                                   C.c_void_p(L.stream_ptr(gt.device))), "pn_max_f32")
    return out


class RenderRaysFn(torch.autograd.Function):
    """rays -> (depth, variance, colour); backward into grids, decoder
    parameters and rays.  Replaces Renderer.py:82-203 + autograd."""

    N_LEAD = 7

    @staticmethod
    def forward(ctx, plan: E.Plan, cfg: _RayCfg, rays_o, rays_d, gt_depth, depth_max, t_rand, *tensors):
        lib = L.lib()
        dev = rays_o.device
        st = C.c_void_p(L.stream_ptr(dev))
        ctx.set_materialize_grads(False)
        ro = rays_o.detach().float().contiguous()
        rd = rays_d.detach().float().contiguous()
        R = ro.shape[0]
        gt = gt_depth.detach().reshape(-1).float().contiguous() if gt_depth is not None else None
        if gt is not None and depth_max is None and R > 0:
            depth_max = batch_depth_max(gt)
        n_surface = cfg.n_surface if gt is not None else 0
        S = cfg.n_samples + n_surface
        z = torch.empty((R, S), dtype=torch.float64, device=dev)
        k = cfg.consts
        tr = t_rand.detach().float().contiguous() if t_rand is not None else None
        with L.device_guard(dev):
            L.check(lib.pn_ray_zvals(C.c_void_p(ro.data_ptr()), C.c_void_p(rd.data_ptr()), C.c_void_p(L.ptr(gt)),
                                     C.c_void_p(L.ptr(depth_max)), C.c_int64(R), E.host_bound(cfg.bound), cfg.n_samples,
                                     n_surface, int(bool(cfg.lindisp)), C.c_void_p(k["t_vals"].data_ptr()),
                                     C.c_void_p(L.ptr(k.get("t_surface"))), C.c_void_p(L.ptr(tr)),
                                     C.c_void_p(z.data_ptr()), st), "pn_ray_zvals")
        grids = {key: E.grid_channels_last(tensors[i]) for i, key in enumerate(plan.grid_keys)}
        needs = ctx.needs_input_grad
        need_grid, want_w = E._grad_flags(plan, needs, RenderRaysFn.N_LEAD, cfg.freeze_map)
        save = any(needs)
        two_pass = cfg.n_importance > 0
        pts = E.Points(n=R * S, rays_o=ro, rays_d=rd, z=z)
        raw, stashes = E.plan_forward(plan, grids, pts, dev, save and not two_pass, want_w)
        depth, var, rgb, w = composite(raw, z, rd, cfg.occupancy, two_pass)
        if two_pass:
            S2 = S + cfg.n_importance
            z2 = torch.empty((R, S2), dtype=torch.float64, device=dev)
            with L.device_guard(dev):
                L.check(lib.pn_importance_zvals(C.c_void_p(z.data_ptr()), C.c_void_p(w.data_ptr()), C.c_int64(R), S,
                                                cfg.n_importance, C.c_void_p(L.ptr(k.get("u_lin"))), None,
                                                C.c_void_p(z2.data_ptr()), st), "pn_importance_zvals")
            z, S = z2, S2
            pts = E.Points(n=R * S, rays_o=ro, rays_d=rd, z=z)
            raw, stashes = E.plan_forward(plan, grids, pts, dev, save, want_w)
            depth, var, rgb, _ = composite(raw, z, rd, cfg.occupancy, False)
        if save:
            ctx.plan, ctx.cfg, ctx.grids, ctx.pts, ctx.stashes, ctx.raw = plan, cfg, grids, pts, stashes, raw
            ctx.need_grid, ctx.want_w = need_grid, want_w
        return depth, var, rgb

    @staticmethod
    def backward(ctx, g_depth, g_var, g_rgb):
        lib = L.lib()
        plan, cfg, pts = ctx.plan, ctx.cfg, ctx.pts
        needs = ctx.needs_input_grad
        z, rd = pts.z, pts.rays_d
        R, S = z.shape
        dev = z.device
        st = C.c_void_p(L.stream_ptr(dev))
        need_rays = bool(needs[2] or needs[3])
        gd = g_depth.double().contiguous() if g_depth is not None else None
        gv = g_var.double().contiguous() if g_var is not None else None
        gc = g_rgb.float().contiguous() if g_rgb is not None else None
        g_raw = torch.empty((R * S, 4), dtype=torch.float32, device=dev)
        g_rd_extra = E.zeros((R, 3), dev) if (need_rays and not cfg.occupancy) else None
        with L.device_guard(dev):
            L.check(lib.pn_composite_bwd(C.c_void_p(ctx.raw.data_ptr()), C.c_void_p(z.data_ptr()), C.c_void_p(rd.data_ptr()),
                                         C.c_int64(R), S, int(cfg.occupancy), C.c_void_p(L.ptr(gd)), C.c_void_p(L.ptr(gv)),
                                         C.c_void_p(L.ptr(gc)), C.c_void_p(g_raw.data_ptr()), C.c_void_p(L.ptr(g_rd_extra)),
                                         st), "pn_composite_bwd")
        g_grids, g_pts, g_params = E.plan_backward(plan, ctx.grids, pts, dev, g_raw, ctx.stashes, ctx.need_grid,
                                                   need_rays, ctx.want_w)
        g_o = g_d = None
        if need_rays:
            g_o = torch.empty((R, 3), dtype=torch.float32, device=dev)
            g_d = torch.empty((R, 3), dtype=torch.float32, device=dev)
            with L.device_guard(dev):
                L.check(lib.pn_points_to_rays_bwd(C.c_void_p(g_pts.data_ptr()), C.c_void_p(z.data_ptr()), C.c_int64(R), S,
                                                  C.c_void_p(g_o.data_ptr()), C.c_void_p(g_d.data_ptr()), st),
                        "pn_points_to_rays_bwd")
            if g_rd_extra is not None:
                g_d = g_d + g_rd_extra
        grads = E._assemble_grads(plan, needs, RenderRaysFn.N_LEAD, g_grids, g_params)
        return (None, None, g_o if needs[2] else None, g_d if needs[3] else None, None, None, None, *grads)


class Renderer(object):
    def __init__(self, cfg, args, slam, points_batch_size=500000, ray_batch_size=100000):
        # same fields as Renderer.py:6-21
        self.ray_batch_size = ray_batch_size
        self.points_batch_size = points_batch_size
        self.lindisp = cfg['rendering']['lindisp']
        self.perturb = cfg['rendering']['perturb']
        self.N_samples = cfg['rendering']['N_samples']
        self.N_surface = cfg['rendering']['N_surface']
        self.N_importance = cfg['rendering']['N_importance']
        self.scale = cfg['scale']
        self.occupancy = cfg['occupancy']
        # the fork hard-codes False (Renderer.py:18); upstream reads slam.nice
        self.nice = bool(getattr(slam, 'nice', False))
        self.bound = slam.bound
        self.H, self.W, self.fx, self.fy, self.cx, self.cy = slam.H, slam.W, slam.fx, slam.fy, slam.cx, slam.cy
        # extensions (not in the reference)
        self.freeze_map = False     # True: never form grid / decoder gradients (tracking)
        self.depth_max_override = None  # 1-element device tensor: batch max of gt_depth when rays are sharded
        self._consts: Dict[Tuple, Dict[str, torch.Tensor]] = {}

    # ------------------------------------------------------------------
    def _plan(self, decoders, stage: str, masked: bool = True) -> E.Plan:
        if self.nice:
            passes = E.stage_passes(decoders, stage, self.bound)
        else:
            passes = E.single_pass(decoders, "imap", self.bound)
        return E.Plan(passes, self.bound if masked else None)

    def _constants(self, device) -> Dict[str, torch.Tensor]:
        key = (str(device), self.N_samples, self.N_surface, self.N_importance)
        if key not in self._consts:
            self._consts[key] = _ray_constants(device, self.N_samples, self.N_surface, self.N_importance)
        return self._consts[key]

    # ------------------------------------------------------------------
    def eval_points(self, p, decoders, c=None, stage='color', device='cuda:0'):
        """Occupancy / colour of points, (N,3) -> (N,4); logit 100 outside the
        bound (Renderer.py:23-61).  One fused pass per decoder.  The NICE decoders keep activations on chip,
        so no chunking is needed; the iMAP* MLP stages (N x 256) activations per layer, so without gradients
        it walks the batch in chunks of 500k points through one scratch buffer (imap.forward), like the
        reference's points_batch_size loop (Renderer.py:38-41)."""
        plan = self._plan(decoders, stage)
        if p.dim() != 2 or p.shape[1] != 3:
            raise RuntimeError(f"eval_points expects (N,3) points, got {tuple(p.shape)}")
        if p.dtype not in (torch.float32, torch.float64):
            p = p.float()
        return E.eval_plan(plan, p, c if c is not None else {}, self.freeze_map)

    def sample_z(self, rays_d, rays_o, gt_depth=None, t_rand=None):
        """The sorted sample depths (N, S) float64 that render_batch_ray places
        along each ray (Renderer.py:90-175); exposed for parity tests."""
        dev = rays_o.device
        ro = rays_o.detach().float().contiguous()
        rd = rays_d.detach().float().contiguous()
        R = ro.shape[0]
        gt = gt_depth.detach().reshape(-1).float().contiguous() if gt_depth is not None else None
        dmax = None
        if gt is not None:
            dmax = self.depth_max_override if self.depth_max_override is not None else batch_depth_max(gt)
        n_surface = self.N_surface if gt is not None else 0
        z = torch.empty((R, self.N_samples + n_surface), dtype=torch.float64, device=dev)
        k = self._constants(dev)
        tr = t_rand.detach().float().contiguous().to(dev) if t_rand is not None else None
        with L.device_guard(dev):
            L.check(L.lib().pn_ray_zvals(C.c_void_p(ro.data_ptr()), C.c_void_p(rd.data_ptr()), C.c_void_p(L.ptr(gt)),
                                         C.c_void_p(L.ptr(dmax)), C.c_int64(R), E.host_bound(self.bound), self.N_samples,
                                         n_surface, int(bool(self.lindisp)), C.c_void_p(k["t_vals"].data_ptr()),
                                         C.c_void_p(L.ptr(k.get("t_surface"))), C.c_void_p(L.ptr(tr)),
                                         C.c_void_p(z.data_ptr()), C.c_void_p(L.stream_ptr(dev))), "pn_ray_zvals")
        return z

    def render_batch_ray(self, c, decoders, rays_d, rays_o, device, stage, gt_depth=None):
        """depth (N,) f64, uncertainty (N,) f64, colour (N,3) f32 (Renderer.py:63-203)."""
        plan = self._plan(decoders, stage)
        dev = rays_o.device
        if rays_o.shape[0] == 0:  # empty batch: nothing to launch
            return (torch.zeros(0, dtype=torch.float64, device=dev), torch.zeros(0, dtype=torch.float64, device=dev),
                    torch.zeros((0, 3), dtype=torch.float32, device=dev))
        cfg = _RayCfg(self.N_samples, self.N_surface, self.N_importance, bool(self.lindisp), float(self.perturb),
                      bool(self.occupancy), self.bound, bool(self.freeze_map), self._constants(dev))
        t_rand = None
        if self.perturb > 0. and self.N_importance > 0:
            # Renderer.py:188 draws random u for the importance samples when perturb > 0; this build resamples with the
            # deterministic table only (every shipped config has perturb 0.0)
            raise NotImplementedError("N_importance > 0 with perturb > 0 (random importance samples) is not built")
        if self.perturb > 0.:
            # the reference draws on the CPU generator and uploads (Renderer.py:170)
            t_rand = torch.rand((rays_o.shape[0], self.N_samples)).to(dev)
        grids = [c[k] for k in plan.grid_keys]
        dmax = self.depth_max_override if gt_depth is not None else None
        return RenderRaysFn.apply(plan, cfg, rays_o, rays_d, gt_depth, dmax, t_rand, *grids, *plan.flat_params())

    def render_img(self, c, decoders, c2w, device, stage, gt_depth=None):
        """Full-frame render, (H,W) f64, (H,W) f64, (H,W,3) f32 (Renderer.py:205-260).
        Chunked by ray_batch_size like the reference because the far clamp uses
        the per-chunk maximum depth."""
        from .common import get_rays
        with torch.no_grad():
            H, W = self.H, self.W
            rays_o, rays_d = get_rays(H, W, self.fx, self.fy, self.cx, self.cy, c2w, device)
            rays_o = rays_o.reshape(-1, 3)
            rays_d = rays_d.reshape(-1, 3)
            gt_depth = gt_depth.reshape(-1)
            ds, us, cs = [], [], []
            for i in range(0, rays_d.shape[0], self.ray_batch_size):
                d, u, col = self.render_batch_ray(c, decoders, rays_d[i:i + self.ray_batch_size],
                                                  rays_o[i:i + self.ray_batch_size], device, stage,
                                                  gt_depth=gt_depth[i:i + self.ray_batch_size])
                ds.append(d.double()); us.append(u.double()); cs.append(col)
            return (torch.cat(ds, 0).reshape(H, W), torch.cat(us, 0).reshape(H, W), torch.cat(cs, 0).reshape(H, W, 3))

    def regulation(self, c, decoders, rays_d, rays_o, gt_depth, device, stage='color', t_rand=None):
        """Densities of jittered samples between the camera and 0.85*depth
        (iMAP* only, Renderer.py:263-301).  Differentiable w.r.t. the decoders and -- through
        ``pts = o + d*z`` (Renderer.py:296-297) -- the rays, so that under bundle adjustment the
        regulation loss reaches the camera tensors as it does in the reference (Mapper.py:650-655)."""
        dev = rays_o.device
        R = rays_o.shape[0]
        if t_rand is None:
            t_rand = torch.rand((R, self.N_samples))  # CPU generator, as the reference
        t_rand = t_rand.to(dev).float().contiguous()
        gt = gt_depth.detach().reshape(-1).float().contiguous()
        pts = _RegulationPointsFn.apply(rays_o, rays_d, gt, t_rand, self._constants(dev)["t_vals"], self.N_samples)
        raw = self.eval_points(pts, decoders, c, stage, device)
        return raw[:, -1]


class _RegulationPointsFn(torch.autograd.Function):
    """(rays_o, rays_d) -> the float32 regulation sample points (R*S,3); backward sums the point gradients
    back onto the rays (z depends on gt_depth and the jitter only, which carry no gradient)."""

    @staticmethod
    def forward(ctx, rays_o, rays_d, gt, t_rand, t_vals, n_samples):
        dev = rays_o.device
        ro = rays_o.detach().float().contiguous()
        rd = rays_d.detach().float().contiguous()
        R = ro.shape[0]
        need = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]
        pts = torch.empty((R * n_samples, 3), dtype=torch.float32, device=dev)
        z = torch.empty((R, n_samples), dtype=torch.float64, device=dev) if need else None
        with L.device_guard(dev):
            L.check(L.lib().pn_regulation_points(C.c_void_p(ro.data_ptr()), C.c_void_p(rd.data_ptr()), C.c_void_p(gt.data_ptr()),
                                                 C.c_void_p(t_vals.data_ptr()), C.c_void_p(t_rand.data_ptr()),
                                                 C.c_int64(R), n_samples, C.c_void_p(pts.data_ptr()), C.c_void_p(L.ptr(z)),
                                                 C.c_void_p(L.stream_ptr(dev))), "pn_regulation_points")
        ctx.z = z
        ctx.set_materialize_grads(False)
        return pts

    @staticmethod
    def backward(ctx, g_pts):
        if g_pts is None or ctx.z is None:
            return None, None, None, None, None, None
        z = ctx.z
        R, S = z.shape
        dev = z.device
        g = g_pts.float().contiguous()
        g_o = torch.empty((R, 3), dtype=torch.float32, device=dev)
        g_d = torch.empty((R, 3), dtype=torch.float32, device=dev)
        with L.device_guard(dev):
            L.check(L.lib().pn_points_to_rays_bwd(C.c_void_p(g.data_ptr()), C.c_void_p(z.data_ptr()), C.c_int64(R), S,
                                                  C.c_void_p(g_o.data_ptr()), C.c_void_p(g_d.data_ptr()),
                                                  C.c_void_p(L.stream_ptr(dev))), "pn_points_to_rays_bwd")
        return (g_o if ctx.needs_input_grad[0] else None, g_d if ctx.needs_input_grad[1] else None, None, None, None, None)
