"""CPU restatement of the k-nearest neural-point feature aggregation (BASELINE.json config 4).

TEST INFRASTRUCTURE ONLY (tests/, __graft_entry__.smoke(), bench.py's cpu_baseline leg); the product never imports it.

**Builder-defined semantics -- parity unpinned by the reference.**  The reference tree holds no 3-D neural-point
aggregation: `search_points.py` and `frame.py:362-366` are 2-D keypoint searches with
`scipy.spatial.cKDTree(...).query_ball_point`.  Following SURVEY 8c the specification is Point-NeRF style:

  * neighbours: the K = 8 points nearest to the sample within `radius`, ordered by (d2, point index);
    d2 = ((dx*dx) + (dy*dy)) + (dz*dz) in float32, every operation rounded on its own (numpy float32 arithmetic),
    accepted when d2 <= float32(radius*radius);
  * weights w_k = 1 / (d2_k + eps); feature f = sum_k (w_k / sum_j w_j) F[i_k]; zero when no point is in range.

What pins it: the index lists are checked against `scipy.spatial.cKDTree.query(k=8, distance_upper_bound=radius)`
(float64 distances; `pin_against_ckdtree`) -- the only admissible differences are neighbours whose float32 distances
tie or straddle the radius within float32 rounding -- and values / gradients come from torch autograd over the torch
restatement below.
"""
from __future__ import annotations

import numpy as np
import torch

K = 8
F32 = np.float32


def cell_coords(x: np.ndarray, lo, inv_h, dims) -> np.ndarray:
    """Cell of every point: clamp(floor((x - lo) * inv_h), 0, n-1) per axis in float32 (two rounded operations)."""
    x = np.asarray(x, dtype=F32)
    t = np.floor((x - np.asarray(lo, dtype=F32)[None, :]) * F32(inv_h))
    t = np.clip(t, 0, np.asarray(dims, dtype=F32)[None, :] - 1)
    return t.astype(np.int64)


def build_cells(xyz: np.ndarray, lo, inv_h, dims):
    """start (ncell+1) and the point order grouped by cell (ascending index inside a cell; the CUDA build's order inside
    a cell is arbitrary, so compare cells as sets).  dims = (nx, ny, nz); cell id = (cz*ny + cy)*nx + cx."""
    c = cell_coords(xyz, lo, inv_h, dims)
    nx, ny, nz = dims
    cid = (c[:, 2] * ny + c[:, 1]) * nx + c[:, 0]
    order = np.argsort(cid, kind="stable")
    counts = np.bincount(cid, minlength=nx * ny * nz)
    start = np.zeros(nx * ny * nz + 1, dtype=np.int64)
    np.cumsum(counts, out=start[1:])
    return start, order


def d2_f32(p: np.ndarray, q: np.ndarray) -> np.ndarray:
    """float32 squared distance between p (3,) / (M,3) and q (M,3) in the kernel's operation order."""
    p = np.asarray(p, dtype=F32)
    q = np.asarray(q, dtype=F32)
    dx = p[..., 0] - q[..., 0]
    dy = p[..., 1] - q[..., 1]
    dz = p[..., 2] - q[..., 2]
    return ((dx * dx) + (dy * dy)) + (dz * dz)


def _select(d2: np.ndarray, ids: np.ndarray, r2: np.float32):
    keep = d2 <= r2
    d2, ids = d2[keep], ids[keep]
    key = (d2.view(np.uint32).astype(np.uint64) << np.uint64(32)) | ids.astype(np.uint64)
    o = np.argsort(key, kind="stable")[:K]
    return ids[o], d2[o]


def knn_query(p, xyz, radius, tree=None):
    """idx (N,8) int32 (-1 padded), d2 (N,8) float32 (0 padded).  Candidates come from a float64 ball of radius
    r*(1 + 1e-4) (a superset of the float32 test by a margin of 500 float32 ulps) and are re-evaluated in float32."""
    from scipy.spatial import cKDTree
    p = np.ascontiguousarray(np.asarray(p, dtype=F32))
    xyz = np.ascontiguousarray(np.asarray(xyz, dtype=F32))
    r2 = F32(radius) * F32(radius)
    if tree is None:
        tree = cKDTree(xyz.astype(np.float64))
    cand = tree.query_ball_point(p.astype(np.float64), float(radius) * (1.0 + 1e-4))
    idx = np.full((p.shape[0], K), -1, dtype=np.int32)
    d2o = np.zeros((p.shape[0], K), dtype=F32)
    for n, c in enumerate(cand):
        if not c:
            continue
        ids = np.asarray(c, dtype=np.int64)
        i, d = _select(d2_f32(p[n], xyz[ids]), ids, r2)
        idx[n, :len(i)] = i
        d2o[n, :len(i)] = d
    return idx, d2o


def knn_query_bruteforce(p, xyz, radius):
    """The same selection over ALL points (small cases): no spatial structure at all."""
    p = np.asarray(p, dtype=F32)
    xyz = np.asarray(xyz, dtype=F32)
    r2 = F32(radius) * F32(radius)
    idx = np.full((p.shape[0], K), -1, dtype=np.int32)
    d2o = np.zeros((p.shape[0], K), dtype=F32)
    ids = np.arange(xyz.shape[0], dtype=np.int64)
    for n in range(p.shape[0]):
        i, d = _select(d2_f32(p[n], xyz), ids, r2)
        idx[n, :len(i)] = i
        d2o[n, :len(i)] = d
    return idx, d2o


def pin_against_ckdtree(p, xyz, radius, idx):
    """Compare float32-ordered index lists with scipy.spatial.cKDTree.query (float64).  Returns (n_rows_different,
    worst relative distance gap among the differing entries): a difference is admissible only between neighbours whose
    distances agree to float32 rounding."""
    from scipy.spatial import cKDTree
    p64 = np.asarray(p, dtype=F32).astype(np.float64)
    x64 = np.asarray(xyz, dtype=F32).astype(np.float64)
    tree = cKDTree(x64)
    dist, ref = tree.query(p64, k=K, distance_upper_bound=float(radius))
    ref = np.where(np.isfinite(dist), ref, -1).astype(np.int64)
    bad = np.nonzero((ref != idx).any(axis=1))[0]
    worst = 0.0
    for n in bad:
        for k in range(K):
            a, b = int(ref[n, k]), int(idx[n, k])
            if a == b:
                continue
            da = np.linalg.norm(p64[n] - x64[a]) if a >= 0 else float(radius)
            db = np.linalg.norm(p64[n] - x64[b]) if b >= 0 else float(radius)
            worst = max(worst, abs(da - db) / max(da, db, 1e-30))
    return len(bad), worst


def aggregate(p: torch.Tensor, xyz: torch.Tensor, feat: torch.Tensor, idx: torch.Tensor, eps: float) -> torch.Tensor:
    """f (N,C) from the index lists; differentiable w.r.t. p (through the weights; the neighbour set is piecewise
    constant) and feat.  p (N,3), xyz (P,3), feat (P,C) float32; idx (N,8) integer with -1 padding."""
    valid = idx >= 0
    j = idx.clamp(min=0).long()
    d = p.float()[:, None, :] - xyz[j]
    d2 = (d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1]) + d[..., 2] * d[..., 2]
    w = torch.where(valid, 1.0 / (d2 + eps), torch.zeros_like(d2))
    W = w.sum(1, keepdim=True)
    wn = torch.where(W > 0, w / torch.where(W > 0, W, torch.ones_like(W)), torch.zeros_like(w))
    return (wn[..., None] * feat[j]).sum(1)
