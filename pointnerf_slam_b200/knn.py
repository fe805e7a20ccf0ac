"""k-nearest neural-point feature aggregation (BASELINE.json config 4, the "pointNeRF_slam config").

**Builder-defined semantics, not reference parity**: the reference tree has no 3-D neural-point aggregation; its nearest
code is the 2-D ``cKDTree`` radius search of ``src/frame.py:362-366`` and ``src/search_points.py:122,223,445``
(SURVEY.md 0.3, 8c).  The specification is SURVEY 8c's (Point-NeRF style): the K = 8 nearest neural points within a
radius of every sample, inverse-squared-distance weights, blended 32-channel feature -- the per-sample feature a decoder
would consume in place of the trilinear grid gather (the CPU restatement used by the tests lives outside this package).

Everything runs in ``libpnslam.so`` (``csrc/pn_knn.cu``): there is no CPU path, CPU tensors raise.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import torch

from . import _lib as L

K = 8


class PnKnnIndex(C.Structure):
    _fields_ = [("start", L.P), ("sorted", L.P), ("lo", C.c_float * 3), ("inv_h", C.c_float),
                ("nx", C.c_int), ("ny", C.c_int), ("nz", C.c_int), ("P", C.c_int)]


def _need_cuda(t: torch.Tensor, what: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"{what}: CUDA tensor required (the k-NN path has no CPU implementation)")


def _points_struct(p: Optional[torch.Tensor], rays_o, rays_d, z) -> Tuple[L.PnPoints, int]:
    s = L.PnPoints()
    if p is not None:
        if p.dim() != 2 or p.shape[1] != 3 or not p.is_contiguous():
            raise ValueError("points must be a contiguous (N,3) tensor")
        if p.dtype == torch.float64:
            s.pts64 = p.data_ptr()
        elif p.dtype == torch.float32:
            s.pts32 = p.data_ptr()
        else:
            raise TypeError("points must be float32 or float64")
        s.S, s.N = 1, p.shape[0]
    else:
        if z.dtype != torch.float64 or rays_o.dtype != torch.float32 or rays_d.dtype != torch.float32:
            raise TypeError("ray mode: rays float32 (R,3), z float64 (R,S)")
        s.rays_o, s.rays_d, s.z = rays_o.data_ptr(), rays_d.data_ptr(), z.data_ptr()
        s.S, s.N = z.shape[1], z.shape[0] * z.shape[1]
    return s, int(s.N)


class NeuralPointIndex:
    """Uniform cell lattice over ``bound`` with the points grouped by cell.  ``radius`` is the search radius; the cell
    edge is radius * (1 + 1e-4), so that two points within the radius always sit in the same or adjacent cells even
    after the float32 rounding of the cell coordinate."""

    def __init__(self, xyz: torch.Tensor, bound, radius: float):
        _need_cuda(xyz, "NeuralPointIndex")
        if xyz.dtype != torch.float32 or xyz.dim() != 2 or xyz.shape[1] != 3:
            raise TypeError("xyz must be (P,3) float32")
        self.xyz = xyz.contiguous()
        self.device = xyz.device
        self.radius = float(radius)
        b = torch.as_tensor(bound, dtype=torch.float64).reshape(3, 2)
        lo = b[:, 0].to(torch.float32)
        hi = b[:, 1].to(torch.float32)
        P = self.xyz.shape[0]
        if P > 0:
            mn, mx = self.xyz.min(0).values.cpu(), self.xyz.max(0).values.cpu()
            if bool((mn < lo).any()) or bool((mx > hi).any()):
                raise ValueError("NeuralPointIndex: every neural point must lie inside the bound")
        h = self.radius * (1.0 + 1e-4)
        self.inv_h = float(torch.tensor(1.0 / h, dtype=torch.float32))
        self.lo = [float(v) for v in lo]
        self.dims = [max(1, int(float(hi[a] - lo[a]) / h) + 1) for a in range(3)]
        ncell = self.dims[0] * self.dims[1] * self.dims[2]
        self.start = torch.zeros(ncell + 1, dtype=torch.int32, device=self.device)
        self.sorted = torch.zeros((max(P, 1), 4), dtype=torch.float32, device=self.device)
        self.P = P
        self.rebuild()

    def struct(self) -> PnKnnIndex:
        s = PnKnnIndex()
        s.start, s.sorted = self.start.data_ptr(), self.sorted.data_ptr()
        for a in range(3):
            s.lo[a] = self.lo[a]
        s.inv_h = self.inv_h
        s.nx, s.ny, s.nz = self.dims
        s.P = self.P
        return s

    def rebuild(self) -> None:
        """(Re)group the points by cell: three launches (histogram, one-CTA scan, scatter)."""
        ncell = self.dims[0] * self.dims[1] * self.dims[2]
        scratch = torch.empty(2 * self.P + ncell, dtype=torch.int32, device=self.device)
        s = self.struct()
        with L.device_guard(self.device):
            L.check(L.lib().pn_knn_build(C.c_void_p(self.xyz.data_ptr()), self.P, C.byref(s), C.c_void_p(scratch.data_ptr()),
                                         C.c_void_p(L.stream_ptr(self.device))), "pn_knn_build")

    def query(self, p: Optional[torch.Tensor] = None, rays_o=None, rays_d=None, z=None, want_d2: bool = True):
        """idx (N,8) int32 in ascending (d2, index) order, -1 padded; d2 (N,8) float32."""
        pts, N = _points_struct(p, rays_o, rays_d, z)
        _need_cuda(p if p is not None else z, "NeuralPointIndex.query")
        idx = torch.empty((N, K), dtype=torch.int32, device=self.device)
        d2 = torch.empty((N, K), dtype=torch.float32, device=self.device) if want_d2 else None
        s = self.struct()
        with L.device_guard(self.device):
            L.check(L.lib().pn_knn_query(C.byref(pts), C.byref(s), C.c_float(self.radius), C.c_void_p(idx.data_ptr()),
                                         C.c_void_p(L.ptr(d2)), C.c_void_p(L.stream_ptr(self.device))), "pn_knn_query")
        return idx, d2


class _KnnAggregateFn(torch.autograd.Function):
    """f (N,32) = blend of the 8 nearest neural-point features; VJP into the feature rows and the sample points (or, in
    ray mode, the rays through ``pn_points_to_rays_bwd``)."""

    @staticmethod
    def forward(ctx, field: "NeuralPointField", p, rays_o, rays_d, z, feat):
        index = field.index
        dev = index.device
        pts, N = _points_struct(None if p is None else p.detach(), rays_o, rays_d, z)
        idx = torch.empty((N, K), dtype=torch.int32, device=dev)
        d2 = torch.empty((N, K), dtype=torch.float32, device=dev)
        out = torch.empty((N, 32), dtype=torch.float32, device=dev)
        s = index.struct()
        with L.device_guard(dev), L.timed("knn_fwd", dev):
            L.check(L.lib().pn_knn_aggregate_fwd(C.byref(pts), C.byref(s), C.c_float(index.radius), C.c_float(field.eps),
                                                 C.c_void_p(feat.data_ptr()), C.c_void_p(idx.data_ptr()), C.c_void_p(d2.data_ptr()),
                                                 C.c_void_p(out.data_ptr()), C.c_void_p(L.stream_ptr(dev))), "pn_knn_aggregate_fwd")
        ctx.field = field
        ctx.save_for_backward(*(t for t in (p, rays_o, rays_d, z) if t is not None), feat, idx, d2)
        ctx.ray_mode = p is None
        field.last_idx, field.last_d2 = idx, d2
        return out

    @staticmethod
    def backward(ctx, g_out):
        field = ctx.field
        index = field.index
        dev = index.device
        saved = ctx.saved_tensors
        if ctx.ray_mode:
            rays_o, rays_d, z, feat, idx, d2 = saved
            p = None
        else:
            p, feat, idx, d2 = saved
            rays_o = rays_d = z = None
        pts, N = _points_struct(None if p is None else p.detach(), rays_o, rays_d, z)
        need_f = ctx.needs_input_grad[5]
        need_p = ctx.needs_input_grad[1] or ctx.needs_input_grad[2] or ctx.needs_input_grad[3]
        g_out = g_out.contiguous()
        g_feat = field.grad_sink(feat) if need_f else None
        g_pts = torch.empty((N, 3), dtype=torch.float32, device=dev) if need_p else None
        with L.device_guard(dev), L.timed("knn_bwd", dev):
            L.check(L.lib().pn_knn_aggregate_bwd(C.byref(pts), C.c_void_p(idx.data_ptr()), C.c_void_p(d2.data_ptr()), C.c_float(field.eps),
                                                 C.c_void_p(feat.data_ptr()), C.c_void_p(index.xyz.data_ptr()), C.c_void_p(g_out.data_ptr()),
                                                 C.c_void_p(L.ptr(g_feat)), C.c_void_p(L.ptr(g_pts)), 0,
                                                 C.c_void_p(L.stream_ptr(dev))), "pn_knn_aggregate_bwd")
        g_p = g_o = g_d = None
        if need_p:
            if ctx.ray_mode:
                R, S = z.shape
                g_o = torch.empty((R, 3), dtype=torch.float32, device=dev)
                g_d = torch.empty((R, 3), dtype=torch.float32, device=dev)
                with L.device_guard(dev):
                    L.check(L.lib().pn_points_to_rays_bwd(C.c_void_p(g_pts.data_ptr()), C.c_void_p(z.data_ptr()), C.c_int64(R), S,
                                                          C.c_void_p(g_o.data_ptr()), C.c_void_p(g_d.data_ptr()),
                                                          C.c_void_p(L.stream_ptr(dev))), "pn_points_to_rays_bwd")
            else:
                g_p = g_pts.to(p.dtype)
        return None, g_p, g_o, g_d, None, g_feat


class NeuralPointField:
    """A neural point cloud: positions ``xyz`` (P,3) float32 and features ``feat`` (P,32) float32 rows (128 bytes per
    point, the row layout of the voxel grids, so the sparse row exchange of ``dist`` applies to its gradient as well)."""

    def __init__(self, xyz: torch.Tensor, feat: torch.Tensor, bound, radius: float = 0.16, eps: float = 1e-6):
        _need_cuda(feat, "NeuralPointField")
        if feat.dtype != torch.float32 or feat.dim() != 2 or feat.shape[1] != 32 or feat.shape[0] != xyz.shape[0] or not feat.is_contiguous():
            raise TypeError("feat must be a contiguous (P,32) float32 tensor")
        self.index = NeuralPointIndex(xyz, bound, radius)
        self.feat = feat
        self.eps = float(eps)
        self.last_idx = self.last_d2 = None

    def grad_sink(self, feat: torch.Tensor) -> torch.Tensor:
        """Zeroed (P,32) gradient buffer the backward adds into (one memset per iteration)."""
        return torch.zeros_like(feat)

    def aggregate(self, p: torch.Tensor) -> torch.Tensor:
        """(N,3) float32 / float64 sample points -> (N,32) blended features (differentiable in p and ``feat``)."""
        _need_cuda(p, "NeuralPointField.aggregate")
        return _KnnAggregateFn.apply(self, p.contiguous(), None, None, None, self.feat)

    def aggregate_rays(self, rays_o: torch.Tensor, rays_d: torch.Tensor, z: torch.Tensor) -> torch.Tensor:
        """Samples p = o + d*z (float64, as Renderer.py:177-179) of R rays x S depths -> (R*S,32)."""
        _need_cuda(z, "NeuralPointField.aggregate_rays")
        return _KnnAggregateFn.apply(self, None, rays_o.contiguous(), rays_d.contiguous(), z.contiguous(), self.feat)
