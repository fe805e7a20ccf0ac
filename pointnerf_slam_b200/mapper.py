"""The Mapper's / Tracker's / Mesher's work either side of the rendering path, on the GPU.

Reference (file:line)                                     here
  Mapper.get_mask_from_c2w       src/Mapper.py:129-200     get_mask_from_c2w, frustum_voxel_mask
  optimiser set-up + step        src/Mapper.py:482-505,    StageOptimizer (per-stage learning-rate table, masked
                                 529-536, 657-674          Adam over grids / decoders / cameras in ONE launch)
  ray pre-filter                 src/Mapper.py:607-621,    prefilter_rays
                                 src/Tracker.py:288-300
  Tracker.select_uv              src/Tracker.py:206-226    select_depth_pixels
  keyframe_selection_overlap     src/Mapper.py:267-333     keyframe_selection_overlap
  Mesher.point_masks             src/utils/Mesher.py:53-212  point_masks
  Visualizer residual panels     src/utils/Visualizer.py:60-89  vis_residuals

Host work that stays on the host is the reference's own host work: ``np.linalg.inv`` of 4x4 poses, ``torch.linspace``
of the voxel axes, the sort / random permutation of a few keyframe ids.  Everything per voxel, per ray, per point or
per parameter runs in ``libpnslam.so``; there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import struct
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import _lib as L
from .engine import grid_mlp_tensors, coarse_mlp_tensors, imap_mlp_tensors

GRID_GROUP = {"grid_coarse": 1, "grid_middle": 2, "grid_fine": 3, "grid_color": 4}
LR_KEYS = ("decoders_lr", "coarse_lr", "middle_lr", "fine_lr", "color_lr")


def _cuda(t: torch.Tensor, name: str, dtype=None) -> torch.Tensor:
    if not t.is_cuda:
        raise RuntimeError(f"{name} must live on a CUDA device (this framework has no CPU path)")
    if dtype is not None and t.dtype != dtype:
        raise RuntimeError(f"{name} must be {dtype}, got {t.dtype}")
    return t


def _w2c_host(c2w) -> np.ndarray:
    """np.linalg.inv of a float32 (4,4) / (3,4) pose on the host, exactly the reference's call (Mapper.py:153-154)."""
    m = c2w.detach().cpu().numpy() if isinstance(c2w, torch.Tensor) else np.asarray(c2w)
    m = m.astype(np.float32)
    if m.shape[0] == 3:
        m = np.concatenate([m, np.array([[0, 0, 0, 1]], dtype=np.float32)], 0)
    return np.linalg.inv(m)


# ----------------------------------------------------------------------------------------------
# frustum feature selection
# ----------------------------------------------------------------------------------------------
def frustum_voxel_mask(c2w, key: str, val_shape: Sequence[int], depth: torch.Tensor, bound: torch.Tensor, H, W, fx, fy, cx, cy) -> torch.Tensor:
    """uint8 (Z,Y,X) device mask of the voxels Mapper.get_mask_from_c2w selects (1 = optimise).  `val_shape` = (Z,Y,X) =
    ``val.shape[2:]`` as the reference passes it; `depth` the current frame's (H,W) float32 depth image on the device."""
    depth = _cuda(depth, "depth", torch.float32).contiguous()
    dev = depth.device
    nz, ny, nx = int(val_shape[0]), int(val_shape[1]), int(val_shape[2])
    mask = torch.empty((nz, ny, nx), dtype=torch.uint8, device=dev)
    if key == "grid_coarse":                              # Mapper.py:150-152
        return mask.fill_(1)
    axes = [torch.linspace(bound[a][0], bound[a][1], n).to(dev) for a, n in ((0, nx), (1, ny), (2, nz))]   # Mapper.py:145-147
    c2w_np = (c2w.detach().cpu().numpy() if isinstance(c2w, torch.Tensor) else np.asarray(c2w)).astype(np.float32)
    w2c = np.ascontiguousarray(_w2c_host(c2w_np), dtype=np.float32)
    cam_o = np.ascontiguousarray(c2w_np[:3, 3], dtype=np.float32)
    scratch = torch.empty(nx * ny * nz + 1, dtype=torch.float32, device=dev)
    with L.device_guard(dev):
        L.check(L.lib().pn_frustum_mask(C.c_void_p(axes[0].data_ptr()), nx, C.c_void_p(axes[1].data_ptr()), ny,
                                        C.c_void_p(axes[2].data_ptr()), nz, w2c.ctypes.data_as(C.c_void_p),
                                        cam_o.ctypes.data_as(C.c_void_p), float(fx), float(fy), float(cx), float(cy), int(H), int(W),
                                        C.c_void_p(depth.data_ptr()), C.c_void_p(scratch.data_ptr()),
                                        C.c_void_p(scratch[-1:].data_ptr()), C.c_void_p(mask.data_ptr()),
                                        C.c_void_p(L.stream_ptr(dev))), "pn_frustum_mask")
    return mask


def get_mask_from_c2w(c2w, key: str, val_shape: Sequence[int], depth: torch.Tensor, bound: torch.Tensor, H, W, fx, fy, cx, cy) -> torch.Tensor:
    """Mapper.get_mask_from_c2w: boolean mask of shape (X,Y,Z) (the reference's return value, which its caller
    permutes to (Z,Y,X), Mapper.py:424-425).  A device tensor instead of a numpy array."""
    return frustum_voxel_mask(c2w, key, val_shape, depth, bound, H, W, fx, fy, cx, cy).bool().permute(2, 1, 0)


# ----------------------------------------------------------------------------------------------
# optimiser
# ----------------------------------------------------------------------------------------------
class StageOptimizer:
    """The Mapper's Adam over feature grids, decoders and camera tensors with the per-stage learning rates.

    Mirrors src/Mapper.py:482-505 (six parameter groups: decoders, coarse, middle, fine, colour grid, cameras),
    :529-536 (``set_stage`` writes ``cfg['mapping']['stage'][stage][*_lr] * lr_factor`` into the groups; the camera
    group gets ``BA_cam_lr`` in stage ``color`` only) and :657-674 (``step``).  With frustum feature selection the
    reference optimises ``val[mask]`` and copies it back into the grid before and after every iteration
    (:413-431, 511-518, 665-674); here the masked voxels of the grid are updated in place, which is the same
    arithmetic on the same elements (Adam state exists for masked voxels only, as there).

    One kernel launch per step for every tensor (``pn_adam_step``); learning rates and step counts live on the
    device, so a captured mapping iteration (graphs.GraphedStep) can include the step.
    """
    NGROUPS = 6

    def __init__(self, grids: Dict[str, torch.Tensor], decoder_params: Sequence[torch.Tensor], cameras: Sequence[torch.Tensor] = (),
                 masks: Optional[Dict[str, torch.Tensor]] = None, stage_lr: Optional[dict] = None, lr_factor: float = 1.0,
                 BA_cam_lr: float = 0.001, betas=(0.9, 0.999), eps: float = 1e-8):
        self.stage_lr = stage_lr
        self.lr_factor, self.BA_cam_lr = float(lr_factor), float(BA_cam_lr)
        self.betas, self.eps = (float(betas[0]), float(betas[1])), float(eps)
        self.entries = []   # (param, selected voxel indices or None, row, group)
        dev = None
        for key, g in grids.items():
            _cuda(g, key, torch.float32)
            dev = g.device
            zyx = g.shape[2] * g.shape[3] * g.shape[4]
            if g.is_contiguous(memory_format=torch.channels_last_3d):
                row = g.shape[1]                      # memory [Z][Y][X][C]: one mask byte per C consecutive elements
            elif g.is_contiguous():
                row = -zyx                            # memory [C][Z][Y][X]: mask index = element index mod Z*Y*X
            else:
                raise RuntimeError(f"{key}: grid must be dense in NCDHW or channels-last memory")
            m = None if masks is None else masks.get(key)
            if m is not None:
                m = _cuda(m, key + " mask").reshape(-1)
                m = m.to(torch.uint8) if m.dtype != torch.uint8 else m
                if m.numel() != zyx:
                    raise RuntimeError(f"{key}: the mask must have Z*Y*X = {zyx} elements, got {m.numel()}")
                m = _compact(m.contiguous())          # ascending voxel indices: the step visits the selected voxels only
            self.entries.append((g, m, row, GRID_GROUP[key]))
        for p in decoder_params:
            _cuda(p, "decoder parameter", torch.float32)
            dev = p.device
            self.entries.append((p, None, 1, 0))
        for cam in cameras:
            _cuda(cam, "camera tensor", torch.float32)
            dev = cam.device
            self.entries.append((cam, None, 1, 5))
        if dev is None:
            raise ValueError("StageOptimizer: nothing to optimise")
        self.device = dev
        for p, *_ in self.entries:
            if not p.is_contiguous() and not p.is_contiguous(memory_format=torch.channels_last_3d):
                raise RuntimeError("StageOptimizer: parameters must be dense")
        self.exp_avg = [torch.zeros_like(p) for p, *_ in self.entries]          # preserve_format: same memory order as p
        self.exp_avg_sq = [torch.zeros_like(p) for p, *_ in self.entries]
        self.lr = torch.zeros(self.NGROUPS, dtype=torch.float64, device=dev)
        self.step_count = torch.zeros(len(self.entries), dtype=torch.int32, device=dev)   # per tensor, as torch.optim.Adam
        self._hyper = torch.zeros(2 * len(self.entries), dtype=torch.float32, device=dev)
        self._table = None
        self._table_key = None
        self._nblocks = 0
        self._lr_host = [0.0] * self.NGROUPS
        self._pinned_lr = torch.zeros(self.NGROUPS, dtype=torch.float64).pin_memory()

    # -- learning rates -------------------------------------------------------------------------
    def set_lrs(self, lrs: Sequence[float]) -> None:
        lrs = [float(x) for x in lrs]
        if lrs != self._lr_host:
            self._lr_host = lrs
            self._pinned_lr.copy_(torch.tensor(lrs, dtype=torch.float64))
            self.lr.copy_(self._pinned_lr, non_blocking=True)

    def set_stage(self, stage: str, BA: bool = True) -> None:
        """Mapper.py:529-536."""
        tab = self.stage_lr[stage]
        lrs = [tab[k] * self.lr_factor for k in LR_KEYS]
        lrs.append(self.BA_cam_lr if (BA and stage == "color") else 0.0)
        self.set_lrs(lrs)

    # -- step -----------------------------------------------------------------------------------
    def _build_table(self):
        grads = [p.grad for p, *_ in self.entries]
        key = tuple((p.data_ptr(), 0 if g is None else g.data_ptr()) for (p, *_), g in zip(self.entries, grads))
        if key == self._table_key:
            return
        rows, block0 = [], 0
        for (p, m, row, group), g, ea, es in zip(self.entries, grads, self.exp_avg, self.exp_avg_sq):
            gp = 0
            if g is not None:
                if g.dtype != torch.float32 or g.stride() != p.stride():
                    raise RuntimeError("StageOptimizer: a gradient must be float32 with its parameter's memory layout")
                gp = g.data_ptr()
            n = p.numel()
            chans = row if row > 0 else (n // -row if row < 0 else 1)
            nwork = n if m is None else m.numel() * chans
            rows.append(struct.pack("<5Q q q i i q", p.data_ptr(), gp, ea.data_ptr(), es.data_ptr(), 0 if m is None else m.data_ptr(),
                                    n, nwork, row if m is not None else 1, group, block0))
            block0 += max(1, (nwork + 1023) // 1024)
        raw = np.frombuffer(b"".join(rows), dtype=np.uint8).copy()
        self._table = torch.from_numpy(raw).to(self.device)
        self._table_key, self._nblocks = key, block0

    def step(self) -> None:
        self._build_table()
        with L.device_guard(self.device), L.timed("adam_step", self.device):
            L.check(L.lib().pn_adam_step(C.c_void_p(self._table.data_ptr()), len(self.entries), C.c_int64(self._nblocks),
                                         C.c_void_p(self.lr.data_ptr()), C.c_void_p(self.step_count.data_ptr()),
                                         C.c_void_p(self._hyper.data_ptr()), self.betas[0], self.betas[1], self.eps,
                                         C.c_void_p(L.stream_ptr(self.device))), "pn_adam_step")

    def zero_grad(self, set_to_none: bool = True) -> None:
        for p, *_ in self.entries:
            if set_to_none:
                p.grad = None
            elif p.grad is not None:
                p.grad.zero_()


def decoder_parameters(decoders, fix_fine: bool = True, fix_color: bool = False, nice: bool = True) -> List[torch.Tensor]:
    """The decoder parameter list of Mapper.py:447-458: fine decoder unless fix_fine, colour decoder unless
    fix_color (NICE); every parameter for iMAP*."""
    if not nice:
        return list(decoders.parameters())
    out: List[torch.Tensor] = []
    if not fix_fine:
        out += list(decoders.fine_decoder.parameters())
    if not fix_color:
        out += list(decoders.color_decoder.parameters())
    return out


# ----------------------------------------------------------------------------------------------
# ray pre-filter, pixel selection
# ----------------------------------------------------------------------------------------------
def _compact(flags: torch.Tensor) -> torch.Tensor:
    """Ascending int64 indices of the set flags (stable).  One host synchronisation to learn their number, as the
    reference's boolean indexing / np.where has."""
    n = flags.numel()
    dev = flags.device
    idx = torch.empty(max(n, 1), dtype=torch.int64, device=dev)
    count = torch.zeros(1, dtype=torch.int64, device=dev)
    scratch = torch.empty((n + 1023) // 1024 + 1, dtype=torch.int32, device=dev)
    with L.device_guard(dev):
        L.check(L.lib().pn_compact_flags(C.c_void_p(flags.data_ptr()), C.c_int64(n), C.c_void_p(scratch.data_ptr()),
                                         C.c_void_p(idx.data_ptr()), C.c_int64(n), C.c_void_p(count.data_ptr()),
                                         C.c_void_p(L.stream_ptr(dev))), "pn_compact_flags")
    return idx[:int(count.item())]


def prefilter_mask(rays_o, rays_d, gt_depth, bound) -> torch.Tensor:
    """inside_mask of Mapper.py:610-616 (uint8, device)."""
    ro = _cuda(rays_o, "rays_o", torch.float32).detach().contiguous()
    rd = _cuda(rays_d, "rays_d", torch.float32).detach().contiguous()
    gd = _cuda(gt_depth, "gt_depth", torch.float32).detach().contiguous()
    keep = torch.empty(ro.shape[0], dtype=torch.uint8, device=ro.device)
    with L.device_guard(ro.device):
        L.check(L.lib().pn_ray_prefilter(C.c_void_p(ro.data_ptr()), C.c_void_p(rd.data_ptr()), C.c_void_p(gd.data_ptr()),
                                         C.c_int64(ro.shape[0]), L.f64x6(bound), C.c_void_p(keep.data_ptr()),
                                         C.c_void_p(L.stream_ptr(ro.device))), "pn_ray_prefilter")
    return keep


def prefilter_rays(rays_o, rays_d, gt_depth, gt_color, bound):
    """Mapper.py:607-621 / Tracker.py:288-300: drop the rays whose measured depth lies beyond the scene bound.
    Differentiable w.r.t. the rays (an index_select)."""
    idx = _compact(prefilter_mask(rays_o, rays_d, gt_depth, bound))
    return rays_o.index_select(0, idx), rays_d.index_select(0, idx), gt_depth.index_select(0, idx), gt_color.index_select(0, idx)


def select_depth_pixels(depth: torch.Tensor, H0: int, H1: int, W0: int, W1: int, thresh: float = 0.01) -> torch.Tensor:
    """Flat indices (into the H0:H1, W0:W1 crop) of the pixels with depth > thresh, ascending: the fork tracker's
    select_uv (Tracker.py:206-226: ``np.where(depth_np > 0.01)`` after a device->host copy).  Feed the result to
    ``common.get_samples(..., indices=...)``."""
    depth = _cuda(depth, "depth", torch.float32).contiguous()
    n = (H1 - H0) * (W1 - W0)
    flags = torch.empty(n, dtype=torch.uint8, device=depth.device)
    with L.device_guard(depth.device):
        L.check(L.lib().pn_depth_pixel_flags(C.c_void_p(depth.data_ptr()), int(depth.shape[1]), int(H0), int(H1), int(W0), int(W1),
                                             C.c_float(thresh), C.c_void_p(flags.data_ptr()), C.c_void_p(L.stream_ptr(depth.device))),
                "pn_depth_pixel_flags")
    return _compact(flags)


# ----------------------------------------------------------------------------------------------
# keyframe overlap selection
# ----------------------------------------------------------------------------------------------
def keyframe_overlap_fractions(rays_o, rays_d, gt_depth, keyframe_c2w: Sequence, H, W, fx, fy, cx, cy, N_samples: int = 16,
                               edge: int = 20) -> np.ndarray:
    """percent_inside of every keyframe (Mapper.py:289-322) as float64 numpy."""
    ro = _cuda(rays_o, "rays_o", torch.float32).detach().contiguous()
    rd = _cuda(rays_d, "rays_d", torch.float32).detach().contiguous()
    gd = _cuda(gt_depth, "gt_depth", torch.float32).detach().contiguous()
    dev, R, K = ro.device, ro.shape[0], len(keyframe_c2w)
    if K == 0:
        return np.zeros(0)
    t_vals = torch.linspace(0.0, 1.0, steps=N_samples).to(dev)
    verts = torch.empty((R * N_samples, 3), dtype=torch.float32, device=dev)
    w2c = torch.from_numpy(np.stack([_w2c_host(c) for c in keyframe_c2w]).astype(np.float32)).to(dev)
    counts = torch.empty(K, dtype=torch.int32, device=dev)
    st = C.c_void_p(L.stream_ptr(dev))
    with L.device_guard(dev):
        L.check(L.lib().pn_overlap_points(C.c_void_p(ro.data_ptr()), C.c_void_p(rd.data_ptr()), C.c_void_p(gd.data_ptr()),
                                          C.c_void_p(t_vals.data_ptr()), C.c_int64(R), int(N_samples), C.c_void_p(verts.data_ptr()), st),
                "pn_overlap_points")
        L.check(L.lib().pn_keyframe_overlap(C.c_void_p(verts.data_ptr()), C.c_int64(R * N_samples), C.c_void_p(w2c.data_ptr()), K,
                                            float(fx), float(fy), float(cx), float(cy), int(H), int(W), int(edge),
                                            C.c_void_p(counts.data_ptr()), st), "pn_keyframe_overlap")
    return counts.cpu().numpy().astype(np.int64) / (R * N_samples)


def keyframe_selection_overlap(gt_color, gt_depth, c2w, keyframe_dict, k, N_samples=16, pixels=100, *, H, W, fx, fy, cx, cy, device,
                               rng=None, indices=None) -> List[int]:
    """Mapper.keyframe_selection_overlap (src/Mapper.py:267-333).  `rng`: numpy RandomState for the final permutation
    (default: numpy's global state, as the reference); `indices`: optional fixed pixel indices for get_samples."""
    from .common import get_samples
    ro, rd, gd, _ = get_samples(0, H, 0, W, pixels, H, W, fx, fy, cx, cy, c2w, gt_depth, gt_color, device, indices=indices)
    fr = keyframe_overlap_fractions(ro, rd, gd, [kf["est_c2w"] for kf in keyframe_dict], H, W, fx, fy, cx, cy, N_samples)
    order = sorted(range(len(fr)), key=lambda i: fr[i], reverse=True)
    keep = [i for i in order if fr[i] > 0.0]
    perm = (rng or np.random).permutation(np.array(keep, dtype=np.int64)) if keep else []
    return list(perm[:k])


# ----------------------------------------------------------------------------------------------
# dense-render consumers
# ----------------------------------------------------------------------------------------------
def point_masks(input_points, keyframe_dict, H, W, fx, fy, cx, cy, device, depth_test: bool = False, points_batch_size: int = 500000):
    """Mesher.point_masks, keyframe branch (src/utils/Mesher.py:53-212): (seen, forecast, unseen) numpy bool arrays."""
    pts_all = input_points if isinstance(input_points, torch.Tensor) else torch.from_numpy(np.asarray(input_points))
    pts_all = pts_all.detach().to(device).float().contiguous()
    K = len(keyframe_dict)
    dev = pts_all.device
    w2c = torch.from_numpy(np.stack([_w2c_host(kf["est_c2w"]) for kf in keyframe_dict]).astype(np.float32)).to(dev)
    depths = [_cuda(kf["depth"].to(dev), "keyframe depth").float().contiguous() for kf in keyframe_dict]
    depth_ptrs = torch.tensor([d.data_ptr() for d in depths], dtype=torch.int64, device=dev)
    kf_max = torch.stack([d.max() * 1.1 for d in depths]).float().contiguous()
    scratch = torch.empty(K, dtype=torch.float32, device=dev)
    seen_l, fore_l = [], []
    for pts in torch.split(pts_all, points_batch_size, dim=0):        # max(depth_sample) is per chunk (Mesher.py:159)
        pts = pts.contiguous()
        n = pts.shape[0]
        seen = torch.empty(n, dtype=torch.uint8, device=dev)
        fore = torch.empty(n, dtype=torch.uint8, device=dev)
        with L.device_guard(dev):
            L.check(L.lib().pn_point_masks(C.c_void_p(pts.data_ptr()), C.c_int64(n), C.c_void_p(w2c.data_ptr()),
                                           C.c_void_p(depth_ptrs.data_ptr()), C.c_void_p(kf_max.data_ptr()), K, int(H), int(W),
                                           C.c_float(fx), C.c_float(fy), C.c_float(cx), C.c_float(cy), int(bool(depth_test)),
                                           C.c_void_p(scratch.data_ptr()), C.c_void_p(seen.data_ptr()), C.c_void_p(fore.data_ptr()),
                                           C.c_void_p(L.stream_ptr(dev))), "pn_point_masks")
        seen_l.append(seen); fore_l.append(fore)
    seen = torch.cat(seen_l).bool()
    fore = torch.cat(fore_l).bool()
    unseen = ~(seen | fore)
    return seen.cpu().numpy(), fore.cpu().numpy(), unseen.cpu().numpy()


def vis_residuals(gt_depth, gt_color, depth, color):
    """The arrays Visualizer.vis plots (src/utils/Visualizer.py:60-89): depth residual (float64), clipped input /
    rendered / residual colours, and max(gt_depth), computed on the device in one launch + one reduction."""
    gd = _cuda(gt_depth, "gt_depth", torch.float32).contiguous()
    gc = _cuda(gt_color, "gt_color").float().contiguous()
    d = _cuda(depth, "depth").double().contiguous()
    c = _cuda(color, "color").float().contiguous()
    n, dev = gd.numel(), gd.device
    depth_res = torch.empty_like(d)
    gcc, cc, cres = torch.empty_like(gc), torch.empty_like(c), torch.empty_like(c)
    with L.device_guard(dev):
        L.check(L.lib().pn_vis_residuals(C.c_void_p(gd.data_ptr()), C.c_void_p(gc.data_ptr()), C.c_void_p(d.data_ptr()),
                                         C.c_void_p(c.data_ptr()), C.c_int64(n), C.c_void_p(depth_res.data_ptr()),
                                         C.c_void_p(gcc.data_ptr()), C.c_void_p(cc.data_ptr()), C.c_void_p(cres.data_ptr()),
                                         C.c_void_p(L.stream_ptr(dev))), "pn_vis_residuals")
    return depth_res, gcc, cc, cres, gd.max()
