"""Host glue for the iMAP* single-MLP decoder kernels (pn_imap_mlp_fwd / _bwd)."""
from __future__ import annotations

import ctypes as C
from typing import List, Optional

import torch

from . import _lib as L


class PnImapMlpGrad(C.Structure):
    _fields_ = [("B", L.P), ("W", L.P * 8), ("b", L.P * 8), ("Wo", L.P), ("bo", L.P)]


def _struct(p) -> L.PnImapMlp:
    t = p.params  # [B, W0.., b0.., Wo, bo]
    nb = (len(t) - 3) // 2
    if nb > 8:
        raise RuntimeError("iMAP decoder: at most 8 blocks are supported")
    m = L.PnImapMlp()
    m.B = t[0].data_ptr()
    for i in range(nb):
        m.W[i] = t[1 + i].data_ptr()
        m.b[i] = t[1 + nb + i].data_ptr()
    m.Wo, m.bo = t[1 + 2 * nb].data_ptr(), t[2 + 2 * nb].data_ptr()
    m.hidden, m.n_blocks = t[1].shape[0], nb
    if t[1].shape[1] != 93 or t[1 + 2 * nb].shape[0] != 4 or any(s in getattr(p.dec, "skips", []) for s in range(nb)):
        raise RuntimeError("iMAP decoder must be Fourier-93 -> hidden x n_blocks (no skips) -> 4")
    return m


class ImapStash:
    def __init__(self, n: int, hidden: int, n_blocks: int, device):
        f32 = dict(dtype=torch.float32, device=device)
        self.E = torch.empty(n * 96, **f32)
        self.H = torch.empty(n_blocks * n * hidden, **f32)
        self.P32 = torch.empty(3 * n, **f32)


NO_GRAD_CHUNK = 500000   # points per launch when no stash is kept (the reference's points_batch_size, Renderer.py:6-8)


def _sub_points(pts, b: int, e: int) -> L.PnPoints:
    """pn_points of the sample range [b, e) (ray mode: b and e are multiples of S)."""
    s = L.PnPoints()
    s.N = e - b
    if pts.pts is not None:
        base = pts.pts.data_ptr() + 3 * b * pts.pts.element_size()
        if pts.pts.dtype == torch.float64:
            s.pts64 = base
        else:
            s.pts32 = base
        s.S = 1
    else:
        S = pts.z.shape[1]
        r0 = b // S
        s.rays_o, s.rays_d, s.z = pts.rays_o.data_ptr() + 12 * r0, pts.rays_d.data_ptr() + 12 * r0, pts.z.data_ptr() + 8 * b
        s.S = S
    return s


def forward(p, pts, raw: torch.Tensor, mask_bound, device, save: bool, want_w: bool) -> Optional[ImapStash]:
    """With `save` the whole batch is one launch and its activations (E 96 + H n_blocks x hidden + P32 3 floats per
    sample) are the backward's stash.  Without it (torch.no_grad: render_img, meshing, the first pass of a two-pass
    render) the batch is walked in chunks of NO_GRAD_CHUNK points through ONE chunk-sized activation buffer, so a
    100k-ray render_img chunk or a 256^3 mesh query needs ~2 GB of scratch instead of 20-75 GB."""
    m = _struct(p)
    n = pts.n
    mb = L.f64x6(mask_bound) if mask_bound is not None else None
    S = 1 if pts.pts is not None else pts.z.shape[1]
    step = n if save else max(S, NO_GRAD_CHUNK // S * S)
    st = ImapStash(min(n, step), m.hidden, m.n_blocks, device)
    for b in range(0, n, step):
        e = min(n, b + step)
        ps = pts.struct() if (b == 0 and e == n) else _sub_points(pts, b, e)
        with L.timed("imap_mlp_fwd", device):
            L.check(L.lib().pn_imap_mlp_fwd(C.byref(ps), C.byref(m), mb, 1 if mb is not None else 0,
                                            C.c_void_p(raw.data_ptr() + 16 * b), C.c_void_p(st.E.data_ptr()),
                                            C.c_void_p(st.H.data_ptr()), C.c_void_p(st.P32.data_ptr()),
                                            C.c_void_p(L.stream_ptr(device))), "pn_imap_mlp_fwd")
    return st if save else None


def backward(p, pts, g_raw: torch.Tensor, stash: ImapStash, g_pts: Optional[torch.Tensor], mask_bound, device,
             want_w: bool) -> Optional[List[torch.Tensor]]:
    m = _struct(p)
    n = pts.n
    f32 = dict(dtype=torch.float32, device=device)
    width = max(m.hidden, 96)
    GA, GB, GO = torch.empty(n * width, **f32), torch.empty(n * width, **f32), torch.empty(n * 4, **f32)
    gp = g = None
    if want_w:
        from .engine import zeros_like_flat
        gp = zeros_like_flat(p.params)
        nb = m.n_blocks
        g = PnImapMlpGrad()
        g.B = gp[0].data_ptr()
        for i in range(nb):
            g.W[i] = gp[1 + i].data_ptr()
            g.b[i] = gp[1 + nb + i].data_ptr()
        g.Wo, g.bo = gp[1 + 2 * nb].data_ptr(), gp[2 + 2 * nb].data_ptr()
    ps = pts.struct()
    mb = L.f64x6(mask_bound) if mask_bound is not None else None
    with L.timed("imap_mlp_bwd", device):
        L.check(L.lib().pn_imap_mlp_bwd(C.byref(ps), C.byref(m), mb, 1 if mb is not None else 0, C.c_void_p(g_raw.data_ptr()),
                                        C.c_void_p(stash.E.data_ptr()), C.c_void_p(stash.H.data_ptr()),
                                        C.c_void_p(stash.P32.data_ptr()), C.c_void_p(GA.data_ptr()), C.c_void_p(GB.data_ptr()),
                                        C.c_void_p(GO.data_ptr()), C.c_void_p(L.ptr(g_pts)), C.byref(g) if g is not None else None,
                                        C.c_void_p(L.stream_ptr(device))), "pn_imap_mlp_bwd")
    return gp
