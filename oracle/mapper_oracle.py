"""CPU restatement of the Mapper / Tracker / Mesher code either side of the rendering path
(SURVEY.md 8f rows 1-4).  TEST INFRASTRUCTURE ONLY: nothing under ``pointnerf_slam_b200/`` may
import this module; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU legs do.

Each function follows the cited reference lines op for op (numpy / torch on the CPU, the
reference's dtypes at every step).  Where the reference calls into a third-party routine whose
arithmetic is not spelled out in ``/root/reference`` the restatement fixes one evaluation order
and says so:

  * ``np.linalg.inv`` of a 4x4 float32 pose (LAPACK sgesv): called as is (numpy travels);
  * ``w2c @ homo_vertices`` on float32 (numpy matmul): restated as the sequential float32
    sum ((w0*x + w1*y) + w2*z) + w3*1 without fused multiply-adds;
  * ``cv2.remap(..., INTER_LINEAR)`` on a float32 image (OpenCV 4.x, remapBilinear with
    INTER_BITS = 5 fixed-point coordinates, BORDER_CONSTANT 0): restated in ``cv_remap_bilinear``.

``oracle/pin_mapper_against_reference.py`` runs the reference's own functions
(``Mapper.get_mask_from_c2w``, ``Mapper.keyframe_selection_overlap``, ``Mesher.point_masks``,
``torch.optim.Adam``, ``cv2.remap``) in the build container and reports how the restatement
compares; it also mints ``tests/golden/mapper.npz``.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

F32 = np.float32
INTER_BITS = 5
INTER_TAB_SIZE = 1 << INTER_BITS


# ----------------------------------------------------------------------------------------------
# cv2.remap(src float32 (H,W), map_x, map_y float32 (n,), INTER_LINEAR, BORDER_CONSTANT=0)
# OpenCV 4.x modules/imgproc/src/imgwarp.cpp: RemapInvoker (float maps -> sx = cvRound(x*32),
# integer part sx>>5 saturated to short, 5-bit fraction) + remapBilinear<Cast<float,float>,
# RemapNoVec, float>: D = S00*w0 + S01*w1 + S10*w2 + S11*w3 in float32, left to right, taps outside
# the image read 0; weights w = (1-fy)(1-fx), (1-fy)fx, fy(1-fx), fy fx with f = frac/32 (float32).
# ----------------------------------------------------------------------------------------------
def _cv_round_f32(x: np.ndarray) -> np.ndarray:
    """cvRound of a float32 array: round half to even, out-of-range / NaN -> INT_MIN (cvtps2dq)."""
    x = np.asarray(x, dtype=F32)
    with np.errstate(invalid="ignore", over="ignore"):
        r = np.rint(x.astype(np.float64))
    bad = ~np.isfinite(r) | (r >= 2147483648.0) | (r < -2147483648.0)
    out = np.where(bad, -2147483648.0, r).astype(np.int64)
    return out


def cv_remap_bilinear(src: np.ndarray, map_x: np.ndarray, map_y: np.ndarray) -> np.ndarray:
    src = np.ascontiguousarray(src, dtype=F32)
    Hs, Ws = src.shape
    sx = _cv_round_f32(np.asarray(map_x, F32) * F32(INTER_TAB_SIZE))
    sy = _cv_round_f32(np.asarray(map_y, F32) * F32(INTER_TAB_SIZE))
    fx = (sx & (INTER_TAB_SIZE - 1)).astype(np.int64)
    fy = (sy & (INTER_TAB_SIZE - 1)).astype(np.int64)
    ix = np.clip(sx >> INTER_BITS, -32768, 32767)
    iy = np.clip(sy >> INTER_BITS, -32768, 32767)
    scale = F32(1.0) / F32(INTER_TAB_SIZE)
    ax = fx.astype(F32) * scale
    ay = fy.astype(F32) * scale
    wx0, wx1 = (F32(1.0) - ax).astype(F32), ax
    wy0, wy1 = (F32(1.0) - ay).astype(F32), ay
    w = [(wy0 * wx0).astype(F32), (wy0 * wx1).astype(F32), (wy1 * wx0).astype(F32), (wy1 * wx1).astype(F32)]

    def tap(yy, xx):
        ok = (xx >= 0) & (xx < Ws) & (yy >= 0) & (yy < Hs)
        v = src[np.clip(yy, 0, Hs - 1), np.clip(xx, 0, Ws - 1)]
        return np.where(ok, v, F32(0.0)).astype(F32)

    s = [tap(iy, ix), tap(iy, ix + 1), tap(iy + 1, ix), tap(iy + 1, ix + 1)]
    out = (s[0] * w[0]).astype(F32)
    for k in (1, 2, 3):
        out = (out + (s[k] * w[k]).astype(F32)).astype(F32)
    return out


# ----------------------------------------------------------------------------------------------
# shared projection: world points (float32) -> pixel coordinates, as Mapper.py:152-163 / 303-315
# ----------------------------------------------------------------------------------------------
def project_points(points: np.ndarray, w2c: np.ndarray, fx, fy, cx, cy, eps: float = 1e-5):
    """points (N,3) float32, w2c (4,4) float32 -> uv (N,2) float32, z (N,) float64 (= cam_z + eps, camera looks
    along -z so visible points have z < 0).  cam = w2c @ [p;1] in float32 (sequential, no FMA), x negated,
    uv = K @ cam in float64 (K float64, all nine products as numpy forms them), uv/z -> float32."""
    p = np.asarray(points, F32)
    m = np.asarray(w2c, F32)
    cam = []
    for i in range(3):
        acc = (m[i, 0] * p[:, 0]).astype(F32)
        acc = (acc + (m[i, 1] * p[:, 1]).astype(F32)).astype(F32)
        acc = (acc + (m[i, 2] * p[:, 2]).astype(F32)).astype(F32)
        acc = (acc + (m[i, 3] * F32(1.0))).astype(F32)
        cam.append(acc)
    cam[0] = (cam[0] * F32(-1.0)).astype(F32)
    c0, c1, c2 = [c.astype(np.float64) for c in cam]
    with np.errstate(all="ignore"):
        u = (fx * c0 + 0.0 * c1) + cx * c2
        v = (0.0 * c0 + fy * c1) + cy * c2
        zz = (0.0 * c0 + 0.0 * c1) + 1.0 * c2
        z = zz + eps
        uv = np.stack([u / z, v / z], axis=1).astype(F32)
    return uv, z, cam


# ----------------------------------------------------------------------------------------------
# f2: frustum feature selection, Mapper.get_mask_from_c2w (src/Mapper.py:129-200)
# ----------------------------------------------------------------------------------------------
def grid_axes(bound: torch.Tensor, val_shape: Sequence[int]):
    """The three float32 linspace axes of Mapper.py:145-147 (val_shape = (Z, Y, X) of the grid tensor)."""
    X = torch.linspace(bound[0][0], bound[0][1], val_shape[2])
    Y = torch.linspace(bound[1][0], bound[1][1], val_shape[1])
    Z = torch.linspace(bound[2][0], bound[2][1], val_shape[0])
    return X.numpy(), Y.numpy(), Z.numpy()


def frustum_mask(c2w: np.ndarray, key: str, val_shape: Sequence[int], depth_np: np.ndarray, bound: torch.Tensor,
                 H, W, fx, fy, cx, cy, remap=None) -> np.ndarray:
    """Boolean (X, Y, Z) mask exactly as get_mask_from_c2w returns it (the caller permutes it to (Z,Y,X),
    Mapper.py:425).  `remap` defaults to the restated cv2.remap."""
    if key == "grid_coarse":
        return np.ones(tuple(val_shape[::-1]), dtype=bool)
    remap = remap or cv_remap_bilinear
    X, Y, Z = grid_axes(bound, val_shape)
    gx, gy, gz = np.meshgrid(X, Y, Z, indexing="ij")
    points = np.stack([gx, gy, gz], axis=-1).reshape(-1, 3).astype(F32)
    c2w = np.asarray(c2w, F32)
    w2c = np.linalg.inv(c2w)
    uv, z, _ = project_points(points, w2c, fx, fy, cx, cy, 1e-5)
    depths = remap(np.asarray(depth_np, F32), uv[:, 0], uv[:, 1]).astype(F32)
    edge = 0
    with np.errstate(invalid="ignore"):
        mask = (uv[:, 0] < W - edge) & (uv[:, 0] > edge) & (uv[:, 1] < H - edge) & (uv[:, 1] > edge)
        zero = depths == 0
        depths = depths.copy()
        depths[zero] = np.max(depths)
        mask = mask & (0 <= -z) & (-z <= (depths + F32(0.5)).astype(F32))
    ray_o = c2w[:3, 3].astype(F32)
    d = (points - ray_o[None]).astype(F32)
    dd = (d * d).astype(F32)
    dist = ((dd[:, 0] + dd[:, 1]).astype(F32) + dd[:, 2]).astype(F32)
    mask = mask | (dist < 0.5 * 0.5)
    return mask.reshape(val_shape[2], val_shape[1], val_shape[0])


# ----------------------------------------------------------------------------------------------
# f1: optimiser step -- torch.optim.Adam (single-tensor path) on the masked elements with the
# per-stage learning-rate table (Mapper.py:482-505, 529-536, 657-674; configs/nice_slam.yaml:71-95)
# ----------------------------------------------------------------------------------------------
STAGE_LR = {  # configs/nice_slam.yaml:71-95
    "coarse": {"decoders_lr": 0.0, "coarse_lr": 0.001, "middle_lr": 0.0, "fine_lr": 0.0, "color_lr": 0.0},
    "middle": {"decoders_lr": 0.0, "coarse_lr": 0.0, "middle_lr": 0.1, "fine_lr": 0.0, "color_lr": 0.0},
    "fine": {"decoders_lr": 0.0, "coarse_lr": 0.0, "middle_lr": 0.005, "fine_lr": 0.005, "color_lr": 0.0},
    "color": {"decoders_lr": 0.005, "coarse_lr": 0.0, "middle_lr": 0.005, "fine_lr": 0.005, "color_lr": 0.005},
}


def stage_of_iter(joint_iter: int, num_joint_iters: int, middle_iter_ratio=0.4, fine_iter_ratio=0.6, coarse_mapper=False) -> str:
    """Mapper.py:520-527."""
    if coarse_mapper:
        return "coarse"
    if joint_iter <= int(num_joint_iters * middle_iter_ratio):
        return "middle"
    if joint_iter <= int(num_joint_iters * fine_iter_ratio):
        return "fine"
    return "color"


def adam_step(p: torch.Tensor, g: torch.Tensor, m: torch.Tensor, v: torch.Tensor, step: int, lr: float,
              beta1=0.9, beta2=0.999, eps=1e-8, mask: Optional[torch.Tensor] = None) -> None:
    """torch.optim.adam._single_tensor_adam (no weight decay, no amsgrad), in place, on the elements selected
    by the boolean `mask` (the reference optimises ``val[mask]`` and writes it back, Mapper.py:427-431, 511-518,
    665-674: identical to updating those elements in place)."""
    sel = slice(None) if mask is None else mask
    gg = g[sel]
    mm = m[sel].lerp(gg, 1 - beta1)
    vv = v[sel].mul(beta2).addcmul(gg, gg, value=1 - beta2)
    bc1 = 1 - beta1 ** step
    bc2 = 1 - beta2 ** step
    step_size = lr / bc1
    denom = (vv.sqrt() / math.sqrt(bc2)).add(eps)
    p[sel] = p[sel].addcdiv(mm, denom, value=-step_size)
    m[sel] = mm
    v[sel] = vv


# ----------------------------------------------------------------------------------------------
# f3: ray pre-filter (Mapper.py:607-621, Tracker.py:288-300), pixel selection by depth
# (Tracker.py:206-226 / Mapper.py:216-236) and keyframe overlap selection (Mapper.py:267-333)
# ----------------------------------------------------------------------------------------------
def ray_prefilter_mask(rays_o: torch.Tensor, rays_d: torch.Tensor, gt_depth: torch.Tensor, bound: torch.Tensor) -> torch.Tensor:
    det_o = rays_o.clone().detach().unsqueeze(-1)
    det_d = rays_d.clone().detach().unsqueeze(-1)
    t = (bound.unsqueeze(0) - det_o) / det_d
    t, _ = torch.min(torch.max(t, dim=2)[0], dim=1)
    return t >= gt_depth


def select_depth_pixels(depth_crop: torch.Tensor, thresh: float = 0.01) -> torch.Tensor:
    """Flat indices of the crop's pixels with depth > thresh, ascending (np.where on the flattened crop)."""
    d = depth_crop.reshape(-1).cpu().numpy()
    return torch.from_numpy(np.where(d > thresh)[0].astype(np.int64))


def overlap_points(rays_o: torch.Tensor, rays_d: torch.Tensor, gt_depth: torch.Tensor, n_samples: int = 16) -> np.ndarray:
    """Mapper.py:291-300: n_samples float32 points per ray between 0.8*depth and depth+0.5."""
    gd = gt_depth.reshape(-1, 1).repeat(1, n_samples)
    t_vals = torch.linspace(0.0, 1.0, steps=n_samples)
    near, far = gd * 0.8, gd + 0.5
    z_vals = near * (1.0 - t_vals) + far * t_vals
    pts = rays_o[..., None, :] + rays_d[..., None, :] * z_vals[..., :, None]
    return pts.reshape(-1, 3).cpu().numpy()


def keyframe_overlap_fractions(vertices: np.ndarray, keyframe_c2w: Sequence[np.ndarray], H, W, fx, fy, cx, cy) -> np.ndarray:
    """percent_inside per keyframe, Mapper.py:302-322 (edge 20, in front of the camera)."""
    out = []
    for c2w in keyframe_c2w:
        w2c = np.linalg.inv(np.asarray(c2w, F32))
        uv, z, _ = project_points(vertices, w2c, fx, fy, cx, cy, 1e-5)
        edge = 20
        with np.errstate(invalid="ignore"):
            mask = (uv[:, 0] < W - edge) & (uv[:, 0] > edge) & (uv[:, 1] < H - edge) & (uv[:, 1] > edge)
            mask = mask & (z < 0)
        out.append(mask.sum() / uv.shape[0])
    return np.asarray(out, dtype=np.float64)


def select_overlapping_keyframes(fractions: np.ndarray, k: int, rng: np.random.RandomState) -> List[int]:
    """Mapper.py:324-333: sort by fraction (stable, descending), keep > 0, random permutation, first k."""
    order = sorted(range(len(fractions)), key=lambda i: fractions[i], reverse=True)
    keep = [i for i in order if fractions[i] > 0.0]
    return list(rng.permutation(np.array(keep, dtype=np.int64))[:k]) if keep else []


# ----------------------------------------------------------------------------------------------
# f4: consumers of the dense render -- Visualizer residual panels (src/utils/Visualizer.py:60-89) and
# Mesher.point_masks (src/utils/Mesher.py:53-212)
# ----------------------------------------------------------------------------------------------
def vis_residuals(gt_depth: np.ndarray, gt_color: np.ndarray, depth: np.ndarray, color: np.ndarray):
    depth_residual = np.abs(gt_depth - depth)
    depth_residual[gt_depth == 0.0] = 0.0
    color_residual = np.abs(gt_color - color)
    color_residual[gt_depth == 0.0] = 0.0
    return (depth_residual, np.clip(gt_color, 0, 1), np.clip(color, 0, 1), np.clip(color_residual, 0, 1),
            np.max(gt_depth))


def point_masks(points: torch.Tensor, keyframe_c2w: Sequence[torch.Tensor], keyframe_depth: Sequence[torch.Tensor],
                H, W, fx, fy, cx, cy, depth_test: bool, points_batch_size: int = 500000):
    """Mesher.point_masks, `get_mask_use_all_frames=False` branch (src/utils/Mesher.py:127-196), torch on the CPU."""
    import torch.nn.functional as Fn
    seen_l, fore_l, unseen_l = [], [], []
    for pnts in torch.split(points.clone().detach(), points_batch_size, dim=0):
        pts = pnts.float()
        seen = torch.zeros(pts.shape[0]).bool()
        fore = torch.zeros(pts.shape[0]).bool()
        for c2w_t, depth in zip(keyframe_c2w, keyframe_depth):
            w2c = torch.from_numpy(np.linalg.inv(c2w_t.cpu().numpy())).float()
            ones = torch.ones_like(pts[:, 0]).reshape(-1, 1)
            homo = torch.cat([pts, ones], dim=1).reshape(-1, 4, 1).float()
            cam_cord = (w2c @ homo)[:, :3]
            K = torch.from_numpy(np.array([[fx, .0, cx], [.0, fy, cy], [.0, .0, 1.0]]).reshape(3, 3))
            cam_cord[:, 0] *= -1
            uv = K.float() @ cam_cord.float()
            z = uv[:, -1:] + 1e-8
            uv = (uv[:, :2] / z).float()
            edge = 0
            cur_seen = (uv[:, 0] < W - edge) & (uv[:, 0] > edge) & (uv[:, 1] < H - edge) & (uv[:, 1] > edge)
            cur_seen = cur_seen & (z[:, :, 0] < 0)
            edge = -1000
            cur_fore = (uv[:, 0] < W - edge) & (uv[:, 0] > edge) & (uv[:, 1] < H - edge) & (uv[:, 1] > edge)
            cur_fore = cur_fore & (z[:, :, 0] < 0)
            if depth_test:
                gt_depth = depth.reshape(1, 1, H, W)
                vgrid = uv.reshape(1, 1, -1, 2)
                vgrid[..., 0] = (vgrid[..., 0] / (W - 1) * 2.0 - 1.0)
                vgrid[..., 1] = (vgrid[..., 1] / (H - 1) * 2.0 - 1.0)
                depth_sample = Fn.grid_sample(gt_depth, vgrid, padding_mode="zeros", align_corners=True).reshape(-1)
                max_depth = torch.max(depth_sample)
                cur_fore = cur_fore.reshape(-1)
                pdf = -cam_cord[cur_fore, 2].reshape(-1)
                cur_fore[cur_fore.clone()] &= pdf < max_depth
                cur_seen = cur_seen.reshape(-1)
                pds = -cam_cord[cur_seen, 2].reshape(-1)
                cur_seen[cur_seen.clone()] &= (pds < depth_sample[cur_seen] + 2.4) & (depth_sample[cur_seen] - 2.4 < pds)
            else:
                max_depth = torch.max(depth) * 1.1
                cur_fore = cur_fore.reshape(-1)
                pdf = -cam_cord[cur_fore, 2].reshape(-1)
                cur_fore[cur_fore.clone()] &= pdf < max_depth
                cur_seen = cur_seen.reshape(-1)
                pds = -cam_cord[cur_seen, 2].reshape(-1)
                cur_seen[cur_seen.clone()] &= pds < max_depth
            seen |= cur_seen
            fore |= cur_fore
        fore &= ~seen
        unseen = ~(seen | fore)
        seen_l.append(seen.numpy()); fore_l.append(fore.numpy()); unseen_l.append(unseen.numpy())
    return np.concatenate(seen_l), np.concatenate(fore_l), np.concatenate(unseen_l)
