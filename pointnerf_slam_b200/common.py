"""Drop-ins for the hot-path helpers of the reference's ``src/common.py``.

Same names, argument order and return conventions:

    get_samples(H0, H1, W0, W1, n, H, W, fx, fy, cx, cy, c2w, depth, color, device)
    get_rays(H, W, fx, fy, cx, cy, c2w, device)
    get_rays_from_uv(i, j, c2w, H, W, fx, fy, cx, cy, device)
    get_camera_from_tensor(inputs)          get_tensor_from_camera(RT, Tquad=False)
    raw2outputs_nerf_color(raw, z_vals, rays_d, occupancy=False, device='cuda:0')
    sample_pdf(bins, weights, N_samples, det=False, device='cuda:0')
    normalize_3d_coordinate(p, bound)

Each launches one kernel of ``libpnslam.so`` (plus one for its backward); the
pixel indices are still drawn by ``torch.randint`` on the target device so the
Philox stream -- and therefore the sampled pixels -- are bit-identical to the
reference's (src/common.py:99).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np
import torch

from . import _lib as L
from . import engine as E

f32 = C.c_float


def _as_c2w(c2w, device) -> torch.Tensor:
    if isinstance(c2w, np.ndarray):
        c2w = torch.from_numpy(c2w)
    return c2w.to(device=device, dtype=torch.float32)


class _CameraFn(torch.autograd.Function):
    """(7,) -> (3,4) or (B,7) -> (B,3,4).  The single-pose form is handled here (not by unsqueeze / select around
    the Function) so that its backward is one kernel, without select_backward's zero fill and copy."""

    @staticmethod
    def forward(ctx, cam):
        single = cam.dim() == 1
        camc = cam.detach().float().contiguous().reshape(-1, 7)
        B = camc.shape[0]
        out = torch.empty((B, 3, 4), dtype=torch.float32, device=cam.device)
        with L.device_guard(cam.device):
            L.check(L.lib().pn_camera_from_tensor_fwd(C.c_void_p(camc.data_ptr()), B, C.c_void_p(out.data_ptr()),
                                                      C.c_void_p(L.stream_ptr(cam.device))), "pn_camera_from_tensor_fwd")
        ctx.save_for_backward(camc)
        ctx.single, ctx.in_dtype = single, cam.dtype
        return out[0] if single else out

    @staticmethod
    def backward(ctx, g):
        (camc,) = ctx.saved_tensors
        g = g.float().contiguous()
        out = torch.empty_like(camc)
        with L.device_guard(camc.device):
            L.check(L.lib().pn_camera_from_tensor_bwd(C.c_void_p(camc.data_ptr()), C.c_void_p(g.data_ptr()), camc.shape[0],
                                                      C.c_void_p(out.data_ptr()), C.c_void_p(L.stream_ptr(camc.device))),
                    "pn_camera_from_tensor_bwd")
        out = out[0] if ctx.single else out
        return out if ctx.in_dtype == torch.float32 else out.to(ctx.in_dtype)


def get_camera_from_tensor(inputs: torch.Tensor) -> torch.Tensor:
    """[qw,qx,qy,qz,tx,ty,tz] (7,) or (B,7) -> (3,4) or (B,3,4); differentiable
    (src/common.py:163-176)."""
    return _CameraFn.apply(inputs)


def quad2rotation(quad: torch.Tensor) -> torch.Tensor:
    """(B,4) quaternion -> (B,3,3) (src/common.py:137-160)."""
    cam = torch.cat([quad, torch.zeros((quad.shape[0], 3), dtype=quad.dtype, device=quad.device)], 1)
    return _CameraFn.apply(cam)[:, :, :3]


def get_tensor_from_camera(RT, Tquad: bool = False) -> torch.Tensor:
    """(3,4)/(4,4) matrix -> 7-vector [quat, T] (or [T, quat]); host-side and
    non-differentiable like the reference (src/common.py:179-201, which uses
    ``mathutils``; the same [w,x,y,z] Shepperd conversion is done in numpy here)."""
    dev = None
    if isinstance(RT, torch.Tensor):
        if RT.is_cuda:
            dev = RT.device
        RT = RT.detach().cpu().numpy()
    R, T = np.asarray(RT[:3, :3], dtype=np.float64), np.asarray(RT[:3, 3], dtype=np.float64)
    tr = np.trace(R)
    if tr > 0:
        s = np.sqrt(tr + 1.0) * 2
        q = [0.25 * s, (R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s]
    elif R[0, 0] > R[1, 1] and R[0, 0] > R[2, 2]:
        s = np.sqrt(1.0 + R[0, 0] - R[1, 1] - R[2, 2]) * 2
        q = [(R[2, 1] - R[1, 2]) / s, 0.25 * s, (R[0, 1] + R[1, 0]) / s, (R[0, 2] + R[2, 0]) / s]
    elif R[1, 1] > R[2, 2]:
        s = np.sqrt(1.0 + R[1, 1] - R[0, 0] - R[2, 2]) * 2
        q = [(R[0, 2] - R[2, 0]) / s, (R[0, 1] + R[1, 0]) / s, 0.25 * s, (R[1, 2] + R[2, 1]) / s]
    else:
        s = np.sqrt(1.0 + R[2, 2] - R[0, 0] - R[1, 1]) * 2
        q = [(R[1, 0] - R[0, 1]) / s, (R[0, 2] + R[2, 0]) / s, (R[1, 2] + R[2, 1]) / s, 0.25 * s]
    q = np.asarray(q)
    if q[0] < 0:
        q = -q
    vec = np.concatenate([T, q], 0) if Tquad else np.concatenate([q, T], 0)
    out = torch.from_numpy(vec).float()
    return out.to(dev) if dev is not None else out


class _SampleRaysFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, c2w, idx, geom, depth, color):
        H0, W0, Wc, W, fx, fy, cx, cy = geom
        dev = c2w.device
        n = idx.shape[0]
        c2wc = c2w.detach().float().contiguous()
        ro = torch.empty((n, 3), dtype=torch.float32, device=dev)
        rd = torch.empty((n, 3), dtype=torch.float32, device=dev)
        d_out = torch.empty(n, dtype=torch.float32, device=dev) if depth is not None else None
        c_out = torch.empty((n, 3), dtype=color.dtype, device=dev) if color is not None else None
        with L.device_guard(dev):
            L.check(L.lib().pn_sample_rays_fwd(C.c_void_p(idx.data_ptr()), n, H0, W0, Wc, W, f32(fx), f32(fy), f32(cx), f32(cy),
                                               C.c_void_p(c2wc.data_ptr()), c2wc.shape[-1], C.c_void_p(L.ptr(depth)),
                                               C.c_void_p(L.ptr(color)), int(color is not None and color.dtype == torch.float64),
                                               C.c_void_p(ro.data_ptr()), C.c_void_p(rd.data_ptr()), C.c_void_p(L.ptr(d_out)),
                                               C.c_void_p(L.ptr(c_out)), C.c_void_p(L.stream_ptr(dev))), "pn_sample_rays_fwd")
        ctx.geom, ctx.idx, ctx.shape = geom, idx, tuple(c2w.shape)
        ctx.mark_non_differentiable(*[t for t in (d_out, c_out) if t is not None])
        ctx.set_materialize_grads(False)   # no zero tensors for the (non-differentiable) depth / colour outputs
        return ro, rd, d_out, c_out

    @staticmethod
    def backward(ctx, g_ro, g_rd, _gd, _gc):
        H0, W0, Wc, W, fx, fy, cx, cy = ctx.geom
        idx = ctx.idx
        dev = idx.device
        g = E.zeros((3, 4), dev)
        g_ro = g_ro.float().contiguous() if g_ro is not None else None
        g_rd = g_rd.float().contiguous() if g_rd is not None else None
        with L.device_guard(dev):
            L.check(L.lib().pn_rays_bwd(C.c_void_p(idx.data_ptr()), idx.shape[0], H0, W0, Wc, f32(fx), f32(fy), f32(cx), f32(cy),
                                        C.c_void_p(L.ptr(g_ro)), C.c_void_p(L.ptr(g_rd)), C.c_void_p(g.data_ptr()),
                                        C.c_void_p(L.stream_ptr(dev))), "pn_rays_bwd")
        if ctx.shape[0] == 4:
            g = torch.cat([g, torch.zeros((1, 4), dtype=torch.float32, device=dev)], 0)
        return g, None, None, None, None


def get_samples(H0, H1, W0, W1, n, H, W, fx, fy, cx, cy, c2w, depth, color, device, indices: Optional[torch.Tensor] = None):
    """n random rays of the crop H0..H1 x W0..W1 with their depth / colour
    (src/common.py:92-134).  ``indices`` (flat, row-major over the crop) may be
    passed to replay a draw; by default they come from ``torch.randint`` on
    ``device`` exactly as in the reference."""
    Hc, Wc = H1 - H0, W1 - W0
    if indices is None:
        indices = torch.randint(Hc * Wc, (n,), device=device)
    indices = indices.to(device=device, dtype=torch.int64).contiguous()
    c2w = _as_c2w(c2w, device)
    depth = depth.to(device)
    color = color.to(device)
    if depth.dtype != torch.float32:
        depth = depth.float()
    if color.dtype not in (torch.float32, torch.float64):
        color = color.float()
    geom = (int(H0), int(W0), int(Wc), int(W), float(fx), float(fy), float(cx), float(cy))
    ro, rd, d, c = _SampleRaysFn.apply(c2w, indices, geom, depth.contiguous(), color.contiguous())
    return ro, rd, d, c


class _SampleRaysMultiFn(torch.autograd.Function):
    """(F,3,4) poses + (F,n) pixel indices + F keyframe images -> the concatenated ray batch, one launch each way."""

    @staticmethod
    def forward(ctx, c2w, idx, geom, depth_ptrs, color_ptrs, color_dtype):
        H0, W0, Wc, W, fx, fy, cx, cy = geom
        dev = c2w.device
        F, n = idx.shape
        c2wc = c2w.detach().float().contiguous()
        ro = torch.empty((F * n, 3), dtype=torch.float32, device=dev)
        rd = torch.empty((F * n, 3), dtype=torch.float32, device=dev)
        d_out = torch.empty(F * n, dtype=torch.float32, device=dev)
        c_out = torch.empty((F * n, 3), dtype=color_dtype, device=dev)
        with L.device_guard(dev):
            L.check(L.lib().pn_sample_rays_multi_fwd(C.c_void_p(idx.data_ptr()), F, n, H0, W0, Wc, W, f32(fx), f32(fy), f32(cx), f32(cy),
                                                     C.c_void_p(c2wc.data_ptr()), C.c_void_p(depth_ptrs.data_ptr()),
                                                     C.c_void_p(color_ptrs.data_ptr()), int(color_dtype == torch.float64),
                                                     C.c_void_p(ro.data_ptr()), C.c_void_p(rd.data_ptr()), C.c_void_p(d_out.data_ptr()),
                                                     C.c_void_p(c_out.data_ptr()), C.c_void_p(L.stream_ptr(dev))), "pn_sample_rays_multi_fwd")
        ctx.geom, ctx.idx = geom, idx
        ctx.mark_non_differentiable(d_out, c_out)
        ctx.set_materialize_grads(False)
        return ro, rd, d_out, c_out

    @staticmethod
    def backward(ctx, g_ro, g_rd, _gd, _gc):
        H0, W0, Wc, W, fx, fy, cx, cy = ctx.geom
        idx = ctx.idx
        dev = idx.device
        F, n = idx.shape
        g = E.zeros((F, 3, 4), dev)
        g_ro = g_ro.float().contiguous() if g_ro is not None else None
        g_rd = g_rd.float().contiguous() if g_rd is not None else None
        with L.device_guard(dev):
            L.check(L.lib().pn_rays_multi_bwd(C.c_void_p(idx.data_ptr()), F, n, H0, W0, Wc, f32(fx), f32(fy), f32(cx), f32(cy),
                                              C.c_void_p(L.ptr(g_ro)), C.c_void_p(L.ptr(g_rd)), C.c_void_p(g.data_ptr()),
                                              C.c_void_p(L.stream_ptr(dev))), "pn_rays_multi_bwd")
        return g, None, None, None, None, None


class KeyframeBatch:
    """The keyframes of a mapping window prepared for batched sampling: device arrays of the images' addresses (the
    images themselves stay where the caller keeps them and may be overwritten in place between calls)."""

    def __init__(self, frames, device):
        self.depth = [d if (d.is_cuda and d.dtype == torch.float32 and d.is_contiguous()) else d.to(device).float().contiguous() for d, _ in frames]
        self.color = [c if (c.is_cuda and c.is_contiguous() and c.dtype in (torch.float32, torch.float64)) else c.to(device).float().contiguous()
                      for _, c in frames]
        if len({c.dtype for c in self.color}) != 1:
            raise RuntimeError("the keyframes' colour images must share one dtype")
        self.color_dtype = self.color[0].dtype
        self.depth_ptrs = torch.tensor([d.data_ptr() for d in self.depth], dtype=torch.int64, device=device)
        self.color_ptrs = torch.tensor([c.data_ptr() for c in self.color], dtype=torch.int64, device=device)
        self.F = len(frames)


def get_samples_multi(H0, H1, W0, W1, n, H, W, fx, fy, cx, cy, c2w, batch: "KeyframeBatch", device, indices: Optional[torch.Tensor] = None):
    """``get_samples`` (src/common.py:125-134) for every keyframe of a window at once and concatenated, i.e. the loop of
    src/Mapper.py:558-605: c2w (F,3,4), indices (F,n) (default: ONE ``torch.randint`` draw of F*n values; callers that
    must reproduce the reference's generator stream draw per keyframe and pass the indices, as mapping.MappingIteration does).  Returns rays_o, rays_d (F*n,3), depth (F*n,), colour (F*n,3); differentiable w.r.t. c2w."""
    Hc, Wc = H1 - H0, W1 - W0
    if indices is None:
        indices = torch.randint(Hc * Wc, (batch.F, n), device=device)
    indices = indices.to(device=device, dtype=torch.int64).contiguous()
    geom = (int(H0), int(W0), int(Wc), int(W), float(fx), float(fy), float(cx), float(cy))
    return _SampleRaysMultiFn.apply(c2w, indices, geom, batch.depth_ptrs, batch.color_ptrs, batch.color_dtype)


class _ImageRaysFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, c2w, geom):
        H, W, fx, fy, cx, cy = geom
        dev = c2w.device
        c2wc = c2w.detach().float().contiguous()
        ro = torch.empty((H, W, 3), dtype=torch.float32, device=dev)
        rd = torch.empty((H, W, 3), dtype=torch.float32, device=dev)
        with L.device_guard(dev):
            L.check(L.lib().pn_image_rays_fwd(H, W, f32(fx), f32(fy), f32(cx), f32(cy), C.c_void_p(c2wc.data_ptr()),
                                              c2wc.shape[-1], C.c_void_p(ro.data_ptr()), C.c_void_p(rd.data_ptr()),
                                              C.c_void_p(L.stream_ptr(dev))), "pn_image_rays_fwd")
        ctx.geom, ctx.shape = geom, tuple(c2w.shape)
        ctx.set_materialize_grads(False)
        return ro, rd

    @staticmethod
    def backward(ctx, g_ro, g_rd):
        H, W, fx, fy, cx, cy = ctx.geom
        dev = (g_ro if g_ro is not None else g_rd).device
        g = E.zeros((3, 4), dev)
        g_ro = g_ro.float().contiguous() if g_ro is not None else None
        g_rd = g_rd.float().contiguous() if g_rd is not None else None
        with L.device_guard(dev):
            L.check(L.lib().pn_rays_bwd(None, H * W, 0, 0, W, f32(fx), f32(fy), f32(cx), f32(cy), C.c_void_p(L.ptr(g_ro)),
                                        C.c_void_p(L.ptr(g_rd)), C.c_void_p(g.data_ptr()), C.c_void_p(L.stream_ptr(dev))),
                    "pn_rays_bwd")
        if ctx.shape[0] == 4:
            g = torch.cat([g, torch.zeros((1, 4), dtype=torch.float32, device=dev)], 0)
        return g, None


def get_rays(H, W, fx, fy, cx, cy, c2w, device):
    """All rays of an image, (H,W,3) origins and directions (src/common.py:248-266)."""
    c2w = _as_c2w(c2w, device)
    return _ImageRaysFn.apply(c2w, (int(H), int(W), float(fx), float(fy), float(cx), float(cy)))


def get_rays_from_uv(i, j, c2w, H, W, fx, fy, cx, cy, device):
    """Rays through given pixel coordinates (src/common.py:74-89).  Integer-valued
    i, j (what ``get_sample_uv`` produces) are required."""
    i = i.reshape(-1).to(device)
    j = j.reshape(-1).to(device)
    idx = (j.long() * int(W) + i.long()).contiguous()
    c2w = _as_c2w(c2w, device)
    geom = (0, 0, int(W), int(W), float(fx), float(fy), float(cx), float(cy))
    ro, rd, _, _ = _SampleRaysFn.apply(c2w, idx, geom, None, None)
    return ro, rd


class _CompositeFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, raw, z, rays_d, occupancy):
        from .renderer import composite
        rawc = raw.detach().float().contiguous()
        zc = z.detach().double().contiguous()
        rdc = rays_d.detach().float().contiguous()
        depth, var, rgb, w = composite(rawc, zc, rdc, occupancy, True)
        ctx.save_for_backward(rawc, zc, rdc)
        ctx.occupancy = occupancy
        ctx.set_materialize_grads(False)
        ctx.mark_non_differentiable(w)
        return depth, var, rgb, w

    @staticmethod
    def backward(ctx, g_depth, g_var, g_rgb, _gw):
        rawc, zc, rdc = ctx.saved_tensors
        R, S = zc.shape
        dev = zc.device
        g_raw = torch.empty_like(rawc)
        gd = g_depth.double().contiguous() if g_depth is not None else None
        gv = g_var.double().contiguous() if g_var is not None else None
        gc = g_rgb.float().contiguous() if g_rgb is not None else None
        g_rd = torch.zeros_like(rdc) if (ctx.needs_input_grad[2] and not ctx.occupancy) else None
        with L.device_guard(dev):
            L.check(L.lib().pn_composite_bwd(C.c_void_p(rawc.data_ptr()), C.c_void_p(zc.data_ptr()), C.c_void_p(rdc.data_ptr()),
                                             C.c_int64(R), S, int(ctx.occupancy), C.c_void_p(L.ptr(gd)), C.c_void_p(L.ptr(gv)),
                                             C.c_void_p(L.ptr(gc)), C.c_void_p(g_raw.data_ptr()), C.c_void_p(L.ptr(g_rd)),
                                             C.c_void_p(L.stream_ptr(dev))), "pn_composite_bwd")
        return g_raw, None, g_rd, None


def raw2outputs_nerf_color(raw, z_vals, rays_d, occupancy=False, device='cuda:0'):
    """raw (N,S,4), z (N,S), rays_d (N,3) -> depth, variance, rgb, weights
    (src/common.py:204-245).  One warp per ray; transmittance by warp scan."""
    R, S = z_vals.shape
    return _CompositeFn.apply(raw.reshape(R * S, 4), z_vals, rays_d, bool(occupancy))


def sample_pdf(bins, weights, N_samples, det=False, device='cuda:0'):
    """Inverse-CDF samples (src/common.py:19-63): bins (R,nb), weights (R,nb-1)
    -> (R,N_samples) in the dtype of ``bins``."""
    R, nb = bins.shape
    dev = bins.device
    b = bins.detach().double().contiguous()
    w = weights.detach().float().contiguous()
    out = torch.empty((R, N_samples), dtype=torch.float64, device=dev)
    if det:
        u_lin, u_rand = torch.linspace(0., 1., steps=N_samples).to(dev), None
    else:
        u_lin, u_rand = None, torch.rand((R, N_samples)).to(dev).contiguous()
    with L.device_guard(dev):
        L.check(L.lib().pn_sample_pdf(C.c_void_p(b.data_ptr()), C.c_void_p(w.data_ptr()), C.c_int64(R), nb, N_samples,
                                      C.c_void_p(L.ptr(u_lin)), C.c_void_p(L.ptr(u_rand)), C.c_void_p(out.data_ptr()),
                                      C.c_void_p(L.stream_ptr(dev))), "pn_sample_pdf")
    return out.to(bins.dtype)


def normalize_3d_coordinate(p, bound):
    """Map points into [-1,1] of the bound, in p's dtype (src/common.py:269-284).
    (Inside the fused decoders this step is part of the kernel; this standalone
    form exists for API parity and uses elementwise torch ops.)"""
    p = p.reshape(-1, 3)
    cols = [((p[:, a] - bound[a, 0]) / (bound[a, 1] - bound[a, 0])) * 2 - 1.0 for a in range(3)]
    return torch.stack(cols, -1).to(p.dtype)
