"""Host-side orchestration of the CUDA kernels: decoder passes, stashes and the
two autograd boundaries (points -> raw, rays -> depth/variance/colour).

Nothing here computes on the CPU; every numerical step is a call into
``libpnslam.so``.  torch supplies device memory, streams and the autograd graph
edges only.
"""
from __future__ import annotations

import ctypes as C
import os as _os
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import _lib as L

GRID_KEYS = ("grid_coarse", "grid_middle", "grid_fine", "grid_color")


# --------------------------------------------------------------------------
# decoder parameter access (duck-typed: works for this package's modules and
# for the reference's decoder.MLP / MLP_no_xyz / NICE instances alike)
# --------------------------------------------------------------------------
def grid_mlp_tensors(dec) -> List[torch.Tensor]:
    """[B, W0..4, b0..4, Wc0..4, bc0..4, Wo, bo] (decoder.py:124-159 names)."""
    ts = [dec.embedder._B]
    ts += [l.weight for l in dec.pts_linears]
    ts += [l.bias for l in dec.pts_linears]
    ts += [l.weight for l in dec.fc_c]
    ts += [l.bias for l in dec.fc_c]
    ts += [dec.output_linear.weight, dec.output_linear.bias]
    return ts


def coarse_mlp_tensors(dec) -> List[torch.Tensor]:
    """[W0..4, b0..4, Wo, bo] (decoder.py:235-245 names)."""
    return ([l.weight for l in dec.pts_linears] + [l.bias for l in dec.pts_linears]
            + [dec.output_linear.weight, dec.output_linear.bias])


def imap_mlp_tensors(dec) -> List[torch.Tensor]:
    """[B, W0.., b0.., Wo, bo] of the iMAP* single MLP (c_dim == 0)."""
    return ([dec.embedder._B] + [l.weight for l in dec.pts_linears] + [l.bias for l in dec.pts_linears]
            + [dec.output_linear.weight, dec.output_linear.bias])


def _req(t: torch.Tensor, name: str, dtype=torch.float32) -> torch.Tensor:
    if not t.is_cuda:
        raise RuntimeError(f"{name} must live on a CUDA device (this framework has no CPU path)")
    if t.dtype != dtype:
        raise RuntimeError(f"{name} must be {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise RuntimeError(f"{name} must be contiguous")
    return t


def host_bound(bound) -> C.Array:
    return L.f64x6(bound)


# --------------------------------------------------------------------------
# passes
# --------------------------------------------------------------------------
@dataclass
class Pass:
    kind: str                 # 'grid' | 'coarse' | 'imap'
    dec: object               # decoder sub-module
    grid_a: Optional[str]     # own grid key
    grid_b: Optional[str]     # detached middle grid (fine decoder)
    out_mode: int
    norm_bound: object        # (3,2) bound used for coordinate normalisation
    params: List[torch.Tensor] = field(default_factory=list)

    def __post_init__(self):
        if self.kind == "grid":
            self.params = grid_mlp_tensors(self.dec)
        elif self.kind == "coarse":
            self.params = coarse_mlp_tensors(self.dec)
        else:
            self.params = imap_mlp_tensors(self.dec)

    @property
    def c_dim(self) -> int:
        return self.dec.fc_c[0].weight.shape[1] if self.kind == "grid" else 0

    @property
    def n_out(self) -> int:
        return self.dec.output_linear.weight.shape[0]


def _dec_bound(dec, default):
    b = getattr(dec, "bound", None)
    return default if b is None else b


def stage_passes(decoders, stage: str, default_bound) -> List[Pass]:
    """Kernel passes that realise NICE.forward for a stage (decoder.py:312-342).
    In stage colour the colour decoder writes only raw[..., :3] and the occupancy passes only
    raw[..., 3] (set by fine, then ``+= middle``: ``raw[..., -1] = fine + middle``), so the colour
    pass is independent of the other two and may run on another stream (plan_forward)."""
    if stage == "coarse":
        d = decoders.coarse_decoder
        return [Pass("coarse", d, "grid_coarse", None, L.OUT_SET_ALL, _dec_bound(d, default_bound))]
    mid = decoders.middle_decoder
    p_mid = lambda mode: Pass("grid", mid, "grid_middle", None, mode, _dec_bound(mid, default_bound))
    if stage == "middle":
        return [p_mid(L.OUT_SET_ALL)]
    fine = decoders.fine_decoder
    p_fine = lambda mode: Pass("grid", fine, "grid_fine", "grid_middle", mode, _dec_bound(fine, default_bound))
    if stage == "fine":
        return [p_fine(L.OUT_SET_ALL), p_mid(L.OUT_ADD_W)]
    if stage == "color":
        col = decoders.color_decoder
        return [Pass("grid", col, "grid_color", None, L.OUT_SET_RGB, _dec_bound(col, default_bound)),
                p_fine(L.OUT_SET_W), p_mid(L.OUT_ADD_W)]
    raise ValueError(f"unknown stage {stage!r}")


def single_pass(dec, kind: str, default_bound) -> List[Pass]:
    """One sub-decoder called on its own (MLP.forward / MLP_no_xyz.forward)."""
    if kind == "imap":
        return [Pass("imap", dec, None, None, L.OUT_SET_ALL, default_bound)]
    name = dec.name
    grid_b = "grid_middle" if (kind == "grid" and getattr(dec, "concat_feature", False)) else None
    return [Pass(kind, dec, "grid_" + name, grid_b, L.OUT_SET_ALL, _dec_bound(dec, default_bound))]


@dataclass
class Plan:
    passes: List[Pass]
    mask_bound: Optional[object]      # None -> no out-of-bound override (decoder called directly)
    grid_keys: List[str] = field(default_factory=list)

    def __post_init__(self):
        keys = []
        for p in self.passes:
            for k in (p.grid_a, p.grid_b):
                if k is not None and k not in keys:
                    keys.append(k)
        self.grid_keys = keys

    def flat_params(self) -> List[torch.Tensor]:
        out: List[torch.Tensor] = []
        for p in self.passes:
            out += p.params
        return out


# --------------------------------------------------------------------------
# grids
# --------------------------------------------------------------------------
def grid_channels_last(g: torch.Tensor) -> torch.Tensor:
    """(1,32,Z,Y,X) tensor whose memory is [Z][Y][X][32].  A grid that already is
    in torch.channels_last_3d format is used in place; a contiguous NCDHW grid
    (the reference's layout) is transposed by the library."""
    if g.dim() != 5 or g.shape[0] != 1 or g.shape[1] != 32:
        raise RuntimeError(f"feature grid must be (1,32,Z,Y,X), got {tuple(g.shape)}")
    if not g.is_cuda:
        raise RuntimeError("feature grids must live on a CUDA device (this framework has no CPU path)")
    if g.dtype != torch.float32:
        raise RuntimeError("feature grids must be float32")
    if g.is_contiguous(memory_format=torch.channels_last_3d):
        return g
    src = g.detach().contiguous()
    out = torch.empty_like(src, memory_format=torch.channels_last_3d)
    with L.device_guard(g.device):
        L.check(L.lib().pn_grid_transpose(C.c_void_p(src.data_ptr()), C.c_void_p(out.data_ptr()), g.shape[2], g.shape[3],
                                          g.shape[4], 1, C.c_void_p(L.stream_ptr(g.device))), "pn_grid_transpose")
    return out


class GradArena:
    """One flat float32 buffer that the backward carves its grid and parameter gradient
    sinks from (in a fixed order every iteration).  One memset zeroes them all and -- on
    several GPUs -- ONE all-reduce over ``used()`` sums every gradient of the iteration
    (launch latency, not link count, sizes the exchange).  Install with
    ``engine.GRAD_ARENA = arena`` and call ``arena.reset()`` before each backward."""

    def __init__(self, numel: int, device):
        self.buf = torch.zeros(int(numel), dtype=torch.float32, device=device)
        self.offset = 0

    def reset(self) -> None:
        if self.offset:
            self.buf[:self.offset].zero_()
        self.offset = 0

    def take(self, numel: int) -> Optional[torch.Tensor]:
        n = (int(numel) + 31) // 32 * 32          # keep every sink 128-byte aligned
        if self.offset + n > self.buf.numel():
            return None                            # arena too small: caller falls back to a fresh buffer
        v = self.buf[self.offset:self.offset + int(numel)]
        self.offset += n
        return v

    def used(self) -> torch.Tensor:
        return self.buf[:self.offset]

    def owns(self, t: Optional[torch.Tensor]) -> bool:
        if t is None:
            return False
        lo = self.buf.data_ptr()
        return lo <= t.data_ptr() < lo + self.buf.numel() * 4


GRAD_ARENA: Optional[GradArena] = None


def _zeros(numel: int, device) -> torch.Tensor:
    if GRAD_ARENA is not None and GRAD_ARENA.buf.device == torch.device(device):
        v = GRAD_ARENA.take(numel)
        if v is not None:
            return v                               # zeroed by GradArena.reset()
    return torch.zeros(int(numel), dtype=torch.float32, device=device)


def zeros(shape, device) -> torch.Tensor:
    """Zeroed float32 gradient sink of the given shape (from the arena when one is installed)."""
    n = 1
    for d in shape:
        n *= int(d)
    return _zeros(n, device).view(*shape)


def new_grid_grad(g: torch.Tensor) -> torch.Tensor:
    """Zeroed gradient buffer shaped like g with channels-last memory."""
    n = g.shape[2] * g.shape[3] * g.shape[4] * 32
    return _zeros(n, g.device).view(1, g.shape[2], g.shape[3], g.shape[4], 32).permute(0, 4, 1, 2, 3)


def _pn_grid(g: Optional[torch.Tensor]) -> Optional[L.PnGrid]:
    if g is None:
        return None
    return L.PnGrid(g.data_ptr(), g.shape[2], g.shape[3], g.shape[4])


# --------------------------------------------------------------------------
# point sources
# --------------------------------------------------------------------------
@dataclass
class Points:
    n: int
    pts: Optional[torch.Tensor] = None        # (N,3) float32 / float64
    rays_o: Optional[torch.Tensor] = None     # (R,3) float32
    rays_d: Optional[torch.Tensor] = None
    z: Optional[torch.Tensor] = None          # (R,S) float64

    def struct(self) -> L.PnPoints:
        s = L.PnPoints()
        s.N = self.n
        if self.pts is not None:
            if self.pts.dtype == torch.float64:
                s.pts64 = self.pts.data_ptr()
            else:
                s.pts32 = self.pts.data_ptr()
            s.S = 1
        else:
            s.rays_o, s.rays_d, s.z = self.rays_o.data_ptr(), self.rays_d.data_ptr(), self.z.data_ptr()
            s.S = self.z.shape[1]
        return s


# --------------------------------------------------------------------------
# forward / backward over a plan
# --------------------------------------------------------------------------
def _mlp_struct(p: Pass) -> L.PnGridMlp:
    t = p.params
    m = L.PnGridMlp()
    m.B = t[0].data_ptr()
    for i in range(5):
        m.W[i] = t[1 + i].data_ptr(); m.b[i] = t[6 + i].data_ptr()
        m.Wc[i] = t[11 + i].data_ptr(); m.bc[i] = t[16 + i].data_ptr()
    m.Wo, m.bo = t[21].data_ptr(), t[22].data_ptr()
    m.c_dim, m.n_out = p.c_dim, p.n_out
    return m


def _coarse_struct(p: Pass) -> L.PnCoarseMlp:
    t = p.params
    m = L.PnCoarseMlp()
    for i in range(5):
        m.W[i] = t[i].data_ptr(); m.b[i] = t[5 + i].data_ptr()
    m.Wo, m.bo = t[10].data_ptr(), t[11].data_ptr()
    return m


def _byref(x):
    return C.byref(x) if x is not None else None


class Stash:
    """Device buffers a forward pass leaves for its backward."""

    def __init__(self, p: Pass, n: int, device, weights: bool):
        self.relu_bits = torch.empty(5 * n, dtype=torch.int32, device=device)
        self.H = self.C = self.E = None
        if weights:
            self.H = torch.empty(5 * 32 * n, dtype=torch.float32, device=device)
            self.C = torch.empty(max(p.c_dim, 32) * n, dtype=torch.float32, device=device)
            if p.kind == "grid":
                self.E = torch.empty(96 * n, dtype=torch.float32, device=device)

    def struct(self) -> L.PnStash:
        return L.PnStash(self.relu_bits.data_ptr(), L.ptr(self.H), L.ptr(self.C), L.ptr(self.E))


def plan_forward(plan: Plan, grids_cl: Dict[str, torch.Tensor], pts: Points, device, save: bool,
                 want_w: Sequence[bool]) -> Tuple[torch.Tensor, List[Optional[Stash]]]:
    """Run every pass of the plan; returns raw (N,4) and the per-pass stashes."""
    lib = L.lib()
    n = pts.n
    raw = torch.empty((n, 4), dtype=torch.float32, device=device)
    if n == 0:
        return raw, [None] * len(plan.passes)
    ps = pts.struct()
    mb = host_bound(plan.mask_bound) if plan.mask_bound is not None else None
    apply_mask = 1 if mb is not None else 0
    # everything that outlives this call is allocated here, on the caller's stream (a buffer allocated while a side
    # stream is current would return to that stream's pool while kernels of the caller's stream may still use it)
    stashes: List[Optional[Stash]] = [Stash(p, n, device, bool(want_w[i])) if (save and p.kind != "imap") else None
                                      for i, p in enumerate(plan.passes)]

    def run_pass(i, p):
            st = C.c_void_p(L.stream_ptr(device))      # the stream that is current for THIS pass
            stash = stashes[i]
            sst = stash.struct() if stash is not None else None
            nb = host_bound(p.norm_bound)
            if p.kind == "grid":
                m = _mlp_struct(p)
                ga, gb = _pn_grid(grids_cl[p.grid_a]), _pn_grid(grids_cl.get(p.grid_b) if p.grid_b else None)
                with L.timed(f"grid_mlp_fwd:{p.dec.name}", device):
                    L.check(lib.pn_grid_mlp_fwd(C.byref(ps), C.byref(m), C.byref(ga), _byref(gb), nb, mb, apply_mask,
                                                p.out_mode, C.c_void_p(raw.data_ptr()), _byref(sst), st), "pn_grid_mlp_fwd")
            elif p.kind == "coarse":
                m = _coarse_struct(p)
                ga = _pn_grid(grids_cl[p.grid_a])
                L.check(lib.pn_coarse_mlp_fwd(C.byref(ps), C.byref(m), C.byref(ga), nb, mb, apply_mask, p.out_mode,
                                              C.c_void_p(raw.data_ptr()), _byref(sst), st), "pn_coarse_mlp_fwd")
            else:
                from . import imap
                stashes[i] = imap.forward(p, pts, raw, plan.mask_bound, device, save, bool(want_w[i]))

    # The colour pass (SET_RGB) and the occupancy passes write disjoint components of raw, so with PARALLEL_FORWARD the
    # colour pass goes to a side stream (the two chains could fill each other's tails; see the flag for what was measured).
    side = _side_stream(device) if (PARALLEL_FORWARD and len(plan.passes) > 1 and
                                    any(p.out_mode == L.OUT_SET_RGB for p in plan.passes)) else None
    with L.device_guard(device):
        main = torch.cuda.current_stream(device)
        if side is not None:
            side.wait_stream(main)
        # SM_SPLIT: the two chains get disjoint shares of the SMs (persistent kernels, one CTA per SM), so they really run
        # side by side instead of one filling the other's tail
        cost = [_FWD_COST.get(p.dec.name if p.kind == "grid" else p.kind, 1.0) for p in plan.passes]
        on_side = [side is not None and p.out_mode == L.OUT_SET_RGB for p in plan.passes]
        share = _sm_shares(sum(c for c, s_ in zip(cost, on_side) if s_), sum(c for c, s_ in zip(cost, on_side) if not s_)) \
            if (side is not None and SM_SPLIT) else None
        for i, p in enumerate(plan.passes):
            if on_side[i]:
                with torch.cuda.stream(side), _sm_budget(share[0] if share else None):
                    run_pass(i, p)
            else:
                with _sm_budget(share[1] if share else None):
                    run_pass(i, p)
        if side is not None:
            main.wait_stream(side)
    return raw, stashes


def plan_backward(plan: Plan, grids_cl: Dict[str, torch.Tensor], pts: Points, device, g_raw: torch.Tensor,
                  stashes: List[Optional[Stash]], need_grid: Dict[str, bool], need_pts: bool,
                  want_w: Sequence[bool]):
    """Backward of plan_forward.  Returns (grid grads by key, g_pts or None,
    per-pass parameter-gradient lists or None)."""
    lib = L.lib()
    n = pts.n
    g_grids: Dict[str, torch.Tensor] = {k: new_grid_grad(grids_cl[k]) for k in plan.grid_keys if need_grid.get(k)}
    g_pts = zeros((n, 3), device) if need_pts else None
    g_params: List[Optional[List[torch.Tensor]]] = []
    if n == 0:
        for i, p in enumerate(plan.passes):
            g_params.append([torch.zeros_like(t) for t in p.params] if want_w[i] else None)
        return g_grids, g_pts, g_params
    ps = pts.struct()
    mb = host_bound(plan.mask_bound) if plan.mask_bound is not None else None
    apply_mask = 1 if mb is not None else 0
    # gradient sinks outlive this call: allocated here, on the caller's stream (see plan_forward)
    sinks = [zeros_like_flat(p.params) if (want_w[i] and p.kind != "imap") else None for i, p in enumerate(plan.passes)]

    g_params.extend([None] * len(plan.passes))
    # scratch between a pass's input-gradient kernel and its weight-gradient kernel: allocated here, on the caller's stream
    scratch: List[Optional[dict]] = [None] * len(plan.passes)
    for i, p in enumerate(plan.passes):
        if want_w[i] and p.kind != "imap":
            f32 = dict(dtype=torch.float32, device=device)
            ws = dict(GH=torch.empty(5 * 32 * n, **f32), GARG=torch.empty(96 * n, **f32),
                      P32=torch.empty(3 * n, **f32), GO=torch.empty(4 * n, **f32))
            # GA is neither written nor read for c_dim 32 (pnslam.h, pn_wscratch): alias it
            tc32 = p.kind == "grid" and p.c_dim == 32
            ws["GA"] = ws["GH"] if tc32 else torch.empty(5 * 32 * n, **f32)
            scratch[i] = ws

    def run_input_grad(i, p):
            st = C.c_void_p(L.stream_ptr(device))      # the stream that is current for THIS pass
            nb = host_bound(p.norm_bound)
            gg = g_grids.get(p.grid_a) if p.grid_a else None
            if p.kind == "imap":
                from . import imap
                g_params[i] = imap.backward(p, pts, g_raw, stashes[i], g_pts, plan.mask_bound, device, bool(want_w[i]))
                return
            ws = scratch[i]
            wst = L.PnWscratch(*[ws[k].data_ptr() for k in ("GA", "GH", "GARG", "P32", "GO")]) if ws is not None else None
            sst = stashes[i].struct()
            if p.kind == "grid":
                m = _mlp_struct(p)
                ga, gb = _pn_grid(grids_cl[p.grid_a]), _pn_grid(grids_cl.get(p.grid_b) if p.grid_b else None)
                with L.timed(f"grid_mlp_bwd:{p.dec.name}", device):
                    L.check(lib.pn_grid_mlp_bwd(C.byref(ps), C.byref(m), C.byref(ga), _byref(gb), nb, mb, apply_mask,
                                                C.c_void_p(g_raw.data_ptr()), C.byref(sst), C.c_void_p(L.ptr(gg)),
                                                C.c_void_p(L.ptr(g_pts)), 1, _byref(wst), st), "pn_grid_mlp_bwd")
                if GRAD_READY_HOOK is not None and gg is not None:
                    GRAD_READY_HOOK(p.grid_a, gg)          # this grid's gradient is final (one pass writes each grid)
            else:  # coarse
                m = _coarse_struct(p)
                ga = _pn_grid(grids_cl[p.grid_a])
                L.check(lib.pn_coarse_mlp_bwd(C.byref(ps), C.byref(m), C.byref(ga), nb, mb, apply_mask,
                                              C.c_void_p(g_raw.data_ptr()), C.byref(sst), C.c_void_p(L.ptr(gg)),
                                              C.c_void_p(L.ptr(g_pts)), 1, _byref(wst), st), "pn_coarse_mlp_bwd")

    def run_weight_grad(i, p):
            if not want_w[i] or p.kind == "imap":
                return
            st = C.c_void_p(L.stream_ptr(device))
            ws = scratch[i]
            wst = L.PnWscratch(*[ws[k].data_ptr() for k in ("GA", "GH", "GARG", "P32", "GO")])
            sst = stashes[i].struct()
            gp = sinks[i]
            if p.kind == "grid":
                m = _mlp_struct(p)
                g = L.PnGridMlpGrad()
                g.B = gp[0].data_ptr()
                for k in range(5):
                    g.W[k] = gp[1 + k].data_ptr(); g.b[k] = gp[6 + k].data_ptr()
                    g.Wc[k] = gp[11 + k].data_ptr(); g.bc[k] = gp[16 + k].data_ptr()
                g.Wo, g.bo = gp[21].data_ptr(), gp[22].data_ptr()
                with L.timed(f"grid_mlp_wgrad:{p.dec.name}", device):
                    L.check(lib.pn_grid_mlp_wgrad(C.c_int64(n), C.byref(m), C.byref(sst), C.byref(wst), C.byref(g), st),
                            "pn_grid_mlp_wgrad")
                if GRAD_READY_HOOK is not None:
                    GRAD_READY_HOOK(("params", p.dec.name), gp)
            else:  # coarse
                g = L.PnCoarseMlpGrad()
                for k in range(5):
                    g.W[k] = gp[k].data_ptr(); g.b[k] = gp[5 + k].data_ptr()
                g.Wo, g.bo = gp[10].data_ptr(), gp[11].data_ptr()
                L.check(lib.pn_coarse_mlp_wgrad(C.c_int64(n), C.byref(sst), C.byref(wst), C.byref(g), st),
                        "pn_coarse_mlp_wgrad")
            g_params[i] = gp

    def run_pass(i, p):
            run_input_grad(i, p)
            run_weight_grad(i, p)

    # The passes of a stage are independent given g_raw (each writes its own grid and parameter gradients; the point
    # gradient is accumulated with atomics), so the first pass (+ its weight gradients) runs on the current stream and
    # the others on a side stream: as the persistent CTAs of one kernel drain, the CTAs of the other chain take the SMs.
    # (Not while a gradient-ready hook launches collectives from inside the backward: measured on 2 and 8 GPUs the
    # two chains plus NCCL's kernels are slower than the serial order.)
    side = _side_stream(device) if (PARALLEL_BACKWARD and GRAD_READY_HOOK is None and len(plan.passes) > 1 and
                                    all(p.kind != "imap" for p in plan.passes)) else None
    with L.device_guard(device):
        main = torch.cuda.current_stream(device)
        if side is not None:
            side.wait_stream(main)
        cost = [_BWD_COST.get(p.dec.name if p.kind == "grid" else p.kind, 1.0) + (_WGRAD_COST if want_w[i] else 0.0)
                for i, p in enumerate(plan.passes)]
        if GRIDS_READY_HOOK is not None and side is not None and SM_SPLIT and any(want_w):
            # Data-parallel mapper: every INPUT-gradient kernel first (pass 0 beside the chain of the others, SM shares by
            # cost), then the hook -- all grid gradients are final, it starts their exchange on another stream -- and only
            # then the weight-gradient kernels, which run while the exchange is in flight (on the SMs the hook leaves them).
            share = _sm_shares(sum(_BWD_COST.get(p.dec.name if p.kind == "grid" else p.kind, 1.0) for p in plan.passes[1:]),
                               _BWD_COST.get(plan.passes[0].dec.name if plan.passes[0].kind == "grid" else plan.passes[0].kind, 1.0))
            for i, p in enumerate(plan.passes):
                if i >= 1:
                    with torch.cuda.stream(side), _sm_budget(share[0]):
                        run_input_grad(i, p)
                else:
                    with _sm_budget(share[1]):
                        run_input_grad(i, p)
            main.wait_stream(side)
            leave = int(GRIDS_READY_HOOK(g_grids) or 0)
            full = sum(_sm_shares(1.0, 1.0))
            with _sm_budget(max(1, full - leave) if leave > 0 else None):
                for i, p in enumerate(plan.passes):
                    run_weight_grad(i, p)
            return g_grids, g_pts, g_params
        n_chains = BWD_STREAMS if BWD_STREAMS > 0 else (2 if any(want_w) else 3)
        three = side is not None and SM_SPLIT and n_chains >= 3 and len(plan.passes) == 3
        if three:      # one chain per pass, SMs in proportion to the passes' costs
            side2 = _side_stream(device, 1)
            side2.wait_stream(main)
            total = sum(_sm_shares(1.0, 1.0))
            a = max(1, int(round(total * cost[1] / sum(cost))))
            b = max(1, int(round(total * cost[2] / sum(cost))))
            budgets, streams = [total - a - b, a, b], [None, side, side2]
            for i, p in enumerate(plan.passes):
                if streams[i] is None:
                    with _sm_budget(budgets[i]):
                        run_pass(i, p)
                else:
                    with torch.cuda.stream(streams[i]), _sm_budget(budgets[i]):
                        run_pass(i, p)
            main.wait_stream(side)
            main.wait_stream(side2)
            return g_grids, g_pts, g_params
        share = _sm_shares(sum(cost[1:]), cost[0]) if (side is not None and SM_SPLIT) else None
        if share is not None and SM_SPLIT_BIAS:
            share = (share[0] + SM_SPLIT_BIAS, share[1] - SM_SPLIT_BIAS)
        for i, p in enumerate(plan.passes):
            if side is not None and i >= 1:
                with torch.cuda.stream(side), _sm_budget(share[0] if share else None):
                    run_pass(i, p)
            else:
                with _sm_budget(share[1] if share else None):
                    run_pass(i, p)
        if side is not None:
            main.wait_stream(side)
    return g_grids, g_pts, g_params


PARALLEL_BACKWARD = _os.environ.get("PN_PARALLEL_BACKWARD", "1") != "0"   # two-stream backward (see plan_backward)
# colour pass beside the occupancy passes: possible (disjoint components of raw) but measured without gain on a B200
# (mapping 1.25 ms, tracking 0.31 ms either way; +0.1 ms of host time in eager mode), so off by default
PARALLEL_FORWARD = _os.environ.get("PN_PARALLEL_FORWARD", "0") != "0"
# With two chains in flight, give each a fixed share of the SMs in proportion to its measured cost (B200, 240,000 samples:
# forward colour 0.157 / fine 0.184 / middle 0.139 ms; backward colour 0.19 (+ weight gradients 0.15) / fine 0.16 / middle 0.20)
#
# Measured on a B200 (graph replay, L2 flushed): mapping iteration 1.22-1.23 ms with two full-grid chains (the second chain's
# CTAs only fill the first's tail) -> 1.16-1.17 ms with the SMs split 76 / 72 between [colour backward + weight gradients] and
# [fine + middle backward]; +-4 SMs off that balance costs 1-3 %.  Tracking (no weight gradients, 375 tiles per kernel = less
# than one wave): one chain per pass on a third of the SMs each, 0.309 -> 0.295 ms.  The forward chains gain nothing from a split
# (colour beside fine -> middle: 1.19 ms against 1.17-1.19 with the backward split alone), so PARALLEL_FORWARD stays off.
SM_SPLIT = _os.environ.get("PN_SM_SPLIT", "1") != "0"
SM_SPLIT_BIAS = int(_os.environ.get("PN_SM_SPLIT_BIAS", "0"))     # SMs moved from the main chain to the side chain (tuning)
BWD_STREAMS = int(_os.environ.get("PN_BWD_STREAMS", "0"))         # 0: three chains when no pass computes weight gradients, else two
_FWD_COST = {"color": 0.157, "fine": 0.184, "middle": 0.139, "coarse": 0.05}
_BWD_COST = {"color": 0.19, "fine": 0.16, "middle": 0.20, "coarse": 0.05}
_WGRAD_COST = float(_os.environ.get("PN_WGRAD_COST", "0.19"))   # 0.15 ms alone; 0.19 balances the chains when they run side by side
_SIDE_STREAMS: Dict[Tuple[int, int], "torch.cuda.Stream"] = {}


def _sm_shares(cost_side: float, cost_main: float) -> Tuple[int, int]:
    """(SMs for the side chain, SMs for the main chain) out of what pn_reserve_sms currently leaves."""
    lib = L.lib()
    reserved = lib.pn_reserve_sms(0)
    lib.pn_reserve_sms(reserved)
    total = max(2, torch.cuda.get_device_properties(torch.cuda.current_device()).multi_processor_count - reserved)
    a = min(total - 1, max(1, int(round(total * cost_side / max(cost_side + cost_main, 1e-9)))))
    return a, total - a


class _sm_budget:
    """``with _sm_budget(n):`` the persistent kernels launched inside size their grids for n SMs (None: unchanged)."""

    def __init__(self, n: Optional[int]):
        self.n = n

    def __enter__(self):
        if self.n is not None:
            lib = L.lib()
            full = torch.cuda.get_device_properties(torch.cuda.current_device()).multi_processor_count
            self.prev = lib.pn_reserve_sms(max(0, full - self.n))
        return self

    def __exit__(self, *exc):
        if self.n is not None:
            L.lib().pn_reserve_sms(self.prev)
        return False


def _side_stream(device, which: int = 0) -> "torch.cuda.Stream":
    idx = torch.device(device).index
    idx = torch.cuda.current_device() if idx is None else idx
    key = (idx, which)
    if key not in _SIDE_STREAMS:
        _SIDE_STREAMS[key] = torch.cuda.Stream(device=idx)
    return _SIDE_STREAMS[key]


# Optional callback ``hook(key, grad)`` fired from inside the backward as soon as a gradient
# buffer is complete: key = grid name with the (1,32,Z,Y,X) channels-last gradient, or
# ("params", decoder name) with the list of that decoder's parameter gradients (views of one
# flat buffer).  dist.OverlappedGradReducer uses it to start the NCCL all-reduce of a finished
# gradient while the remaining decoder kernels still run.
GRAD_READY_HOOK = None
# Optional callback ``hook(g_grids) -> SMs to leave free`` fired from inside the backward once EVERY grid gradient is final and
# before the weight-gradient kernels are launched (they are deferred to after the hook): mapping.MappingIteration starts the
# sparse row exchange of the grid gradients there, so that it overlaps the weight-gradient kernels.
GRIDS_READY_HOOK = None


class FlatGrads(list):
    """Parameter-gradient views plus the exact flat slice they were carved from (``.flat``).  The slice, not
    ``view._base``, is what a collective must reduce: with a GradArena installed ``_base`` is the WHOLE arena."""
    flat: torch.Tensor


def zeros_like_flat(tensors: Sequence[torch.Tensor]) -> "FlatGrads":
    """Zeroed gradient sinks for a list of parameters, carved out of ONE buffer (one
    memset launch instead of one per parameter; every view stays 16-byte aligned)."""
    sizes = [(t.numel() + 3) // 4 * 4 for t in tensors]
    flat = _zeros(sum(sizes), tensors[0].device)
    out, off = FlatGrads(), 0
    for t, sz in zip(tensors, sizes):
        out.append(flat[off:off + t.numel()].view(t.shape))
        off += sz
    out.flat = flat
    return out


def _grad_flags(plan: Plan, needs: Sequence[bool], n_lead: int, freeze_map: bool):
    """Split autograd's needs_input_grad (after the first n_lead inputs) into
    per-grid and per-pass flags."""
    ng = len(plan.grid_keys)
    need_grid = {k: (bool(needs[n_lead + i]) and not freeze_map) for i, k in enumerate(plan.grid_keys)}
    want_w = []
    off = n_lead + ng
    for p in plan.passes:
        cnt = len(p.params)
        want_w.append((not freeze_map) and any(needs[off:off + cnt]))
        off += cnt
    return need_grid, want_w


def _assemble_grads(plan: Plan, needs: Sequence[bool], n_lead: int, g_grids, g_params):
    out: List[Optional[torch.Tensor]] = []
    for i, k in enumerate(plan.grid_keys):
        out.append(g_grids.get(k) if needs[n_lead + i] else None)
    off = n_lead + len(plan.grid_keys)
    for p, gp in zip(plan.passes, g_params):
        for j in range(len(p.params)):
            out.append(gp[j] if (gp is not None and needs[off + j]) else None)
        off += len(p.params)
    return out


class EvalPointsFn(torch.autograd.Function):
    """points (N,3) -> raw (N,4) through the plan's decoders."""

    @staticmethod
    def forward(ctx, plan: Plan, freeze_map: bool, p: torch.Tensor, *tensors):
        device = p.device
        grids = {k: grid_channels_last(tensors[i]) for i, k in enumerate(plan.grid_keys)}
        needs = ctx.needs_input_grad
        need_grid, want_w = _grad_flags(plan, needs, 3, freeze_map)
        save = any(needs)
        pts = Points(n=p.shape[0], pts=p.detach().contiguous())
        raw, stashes = plan_forward(plan, grids, pts, device, save, want_w)
        if save:
            ctx.plan, ctx.grids, ctx.pts, ctx.stashes = plan, grids, pts, stashes
            ctx.need_grid, ctx.want_w, ctx.p_dtype = need_grid, want_w, p.dtype
        return raw

    @staticmethod
    def backward(ctx, g_raw):
        needs = ctx.needs_input_grad
        plan = ctx.plan
        g_raw = g_raw.contiguous().float()
        g_grids, g_pts, g_params = plan_backward(plan, ctx.grids, ctx.pts, g_raw.device, g_raw, ctx.stashes,
                                                 ctx.need_grid, bool(needs[2]), ctx.want_w)
        gp = g_pts.to(ctx.p_dtype) if g_pts is not None else None
        return (None, None, gp, *_assemble_grads(plan, needs, 3, g_grids, g_params))


def eval_plan(plan: Plan, p: torch.Tensor, c: Optional[dict], freeze_map: bool = False) -> torch.Tensor:
    grids = [c[k] for k in plan.grid_keys]
    return EvalPointsFn.apply(plan, freeze_map, p, *grids, *plan.flat_params())
