"""Loss head of the mapping iteration (src/Mapper.py:628-646) as one kernel.

The reference computes ``|gt_depth[m] - depth[m]|.sum() + w_color * |gt_color - color|.sum()`` with ~10 elementwise
and reduction launches (and as many again in the backward); here the value and its gradient come from ONE launch,
and the backward is two scalings by the upstream gradient.  Optional: the callers' torch expression keeps working.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib as L


class _MappingLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, depth, color, gt_depth, gt_color, w_color, use_color, depth_supervision=True):
        dev = depth.device
        d = depth.detach().double().contiguous()
        gd = gt_depth.detach().float().contiguous()
        R = d.shape[0]
        loss = torch.empty((), dtype=torch.float64, device=dev)
        g_depth = torch.empty(R, dtype=torch.float64, device=dev)
        c = gc = g_color = None
        if use_color:
            c = color.detach().float().contiguous()
            gc = gt_color.detach().float().contiguous()
            g_color = torch.empty((R, 3), dtype=torch.float32, device=dev)
        with L.device_guard(dev):
            L.check(L.lib().pn_mapping_loss(C.c_void_p(d.data_ptr()), C.c_void_p(L.ptr(c)), C.c_void_p(gd.data_ptr()),
                                            C.c_void_p(L.ptr(gc)), C.c_int64(R), int(use_color), C.c_float(w_color),
                                            int(bool(depth_supervision)), C.c_void_p(loss.data_ptr()), C.c_void_p(g_depth.data_ptr()),
                                            C.c_void_p(L.ptr(g_color)), C.c_void_p(L.stream_ptr(dev))), "pn_mapping_loss")
        ctx.save_for_backward(g_depth, g_color)
        ctx.color_dtype = color.dtype if use_color else None
        ctx.depth_dtype = depth.dtype
        return loss

    @staticmethod
    def backward(ctx, go):
        g_depth, g_color = ctx.saved_tensors
        # a 0-dim multiplier does not promote the dtype of a dimensioned tensor: one launch each
        gd = (g_depth * go).to(ctx.depth_dtype) if ctx.needs_input_grad[0] else None
        gc = (g_color * go).to(ctx.color_dtype) if (g_color is not None and ctx.needs_input_grad[1]) else None
        return gd, gc, None, None, None, None, None


def mapping_loss_and_grads(depth: torch.Tensor, color: Optional[torch.Tensor], gt_depth: torch.Tensor,
                           gt_color: Optional[torch.Tensor], stage: str = "color", w_color: float = 0.2, nice: bool = True,
                           depth_supervision: bool = True):
    """The same launch as ``mapping_loss`` without an autograd node: (loss, d loss / d depth (R,) in depth's dtype,
    d loss / d colour (R,3) or None).  A caller that owns the whole iteration (``mapping.MappingIteration``) seeds the
    renderer's backward with these directly -- ``torch.autograd.backward([depth, color], [g_depth, g_color])`` -- which
    saves the three launches autograd spends on the loss node (a ones() fill and two scalings by it)."""
    if not depth.is_cuda:
        raise RuntimeError("pointnerf_slam_b200.losses.mapping_loss_and_grads needs CUDA tensors (there is no CPU path)")
    use_color = ((not nice) or stage == "color") and color is not None and gt_color is not None
    if not depth_supervision and not use_color:
        raise ValueError("mapping_loss: without depth supervision the loss is the colour term (stage 'color' or iMAP*)")
    dev = depth.device
    d = depth.detach().double().contiguous()
    gd = gt_depth.detach().float().contiguous()
    R = d.shape[0]
    loss = torch.empty((), dtype=torch.float64, device=dev)
    g_depth = torch.empty(R, dtype=torch.float64, device=dev)
    c = gc = g_color = None
    if use_color:
        c = color.detach().float().contiguous()
        gc = gt_color.detach().float().contiguous()
        g_color = torch.empty((R, 3), dtype=torch.float32, device=dev)
    with L.device_guard(dev):
        L.check(L.lib().pn_mapping_loss(C.c_void_p(d.data_ptr()), C.c_void_p(L.ptr(c)), C.c_void_p(gd.data_ptr()),
                                        C.c_void_p(L.ptr(gc)), C.c_int64(R), int(use_color), C.c_float(w_color),
                                        int(bool(depth_supervision)), C.c_void_p(loss.data_ptr()), C.c_void_p(g_depth.data_ptr()),
                                        C.c_void_p(L.ptr(g_color)), C.c_void_p(L.stream_ptr(dev))), "pn_mapping_loss")
    if g_depth.dtype != depth.dtype:
        g_depth = g_depth.to(depth.dtype)
    if g_color is not None and g_color.dtype != color.dtype:
        g_color = g_color.to(color.dtype)
    return loss, g_depth, g_color


def mapping_loss(depth: torch.Tensor, color: Optional[torch.Tensor], gt_depth: torch.Tensor,
                 gt_color: Optional[torch.Tensor], stage: str = "color", w_color: float = 0.2, nice: bool = True,
                 depth_supervision: bool = True) -> torch.Tensor:
    """Mapper.py:628-646: masked L1 depth loss, plus ``w_color`` times the L1 colour loss in stage
    ``color`` (always for iMAP*).  ``depth_supervision=False`` is the fork's colour-only branch (Mapper.py:633-637).
    Returns a float64 scalar; differentiable w.r.t. depth and colour."""
    if not depth.is_cuda:
        raise RuntimeError("pointnerf_slam_b200.losses.mapping_loss needs CUDA tensors (there is no CPU path)")
    use_color = ((not nice) or stage == "color") and color is not None and gt_color is not None
    if not depth_supervision and not use_color:
        raise ValueError("mapping_loss: without depth supervision the loss is the colour term (stage 'color' or iMAP*)")
    return _MappingLossFn.apply(depth, color if use_color else None, gt_depth, gt_color if use_color else None,
                                float(w_color), bool(use_color), bool(depth_supervision))


class _TrackingLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, depth, var, color, gt_depth, gt_color, w_color, use_color, handle_dynamic, depth_supervision=True):
        dev = depth.device
        d = depth.detach().double().contiguous()
        v = var.detach().double().contiguous()
        gd = gt_depth.detach().float().contiguous()
        R = d.shape[0]
        loss = torch.empty((), dtype=torch.float64, device=dev)
        g_depth = torch.empty(R, dtype=torch.float64, device=dev)
        c = gc = g_color = None
        if use_color:
            c = color.detach().float().contiguous()
            gc = gt_color.detach().float().contiguous()
            g_color = torch.empty((R, 3), dtype=torch.float32, device=dev)
        with L.device_guard(dev):
            L.check(L.lib().pn_tracking_loss(C.c_void_p(d.data_ptr()), C.c_void_p(v.data_ptr()), C.c_void_p(L.ptr(c)),
                                             C.c_void_p(gd.data_ptr()), C.c_void_p(L.ptr(gc)), C.c_int64(R), int(handle_dynamic),
                                             int(use_color), C.c_float(w_color), int(bool(depth_supervision)), C.c_void_p(loss.data_ptr()),
                                             C.c_void_p(g_depth.data_ptr()), C.c_void_p(L.ptr(g_color)),
                                             C.c_void_p(L.stream_ptr(dev))), "pn_tracking_loss")
        ctx.save_for_backward(g_depth, g_color)
        ctx.color_dtype = color.dtype if use_color else None
        ctx.depth_dtype = depth.dtype
        return loss

    @staticmethod
    def backward(ctx, go):
        g_depth, g_color = ctx.saved_tensors
        gd = (g_depth * go).to(ctx.depth_dtype) if ctx.needs_input_grad[0] else None
        gc = (g_color * go).to(ctx.color_dtype) if (g_color is not None and ctx.needs_input_grad[2]) else None
        return gd, None, gc, None, None, None, None, None, None



def tracking_loss_and_grads(depth: torch.Tensor, uncertainty: torch.Tensor, color: Optional[torch.Tensor], gt_depth: torch.Tensor,
                            gt_color: Optional[torch.Tensor], w_color: float = 0.5, use_color: bool = True,
                            handle_dynamic: bool = True, depth_supervision: bool = True):
    """The same launch as ``tracking_loss`` without an autograd node: (loss, d loss / d depth, d loss / d colour or None),
    for a caller that seeds the renderer's backward itself (``tracking.TrackingIteration``)."""
    if not depth.is_cuda:
        raise RuntimeError("pointnerf_slam_b200.losses.tracking_loss_and_grads needs CUDA tensors (there is no CPU path)")
    if depth.shape[0] == 0:
        raise ValueError("tracking_loss needs at least one ray")
    use_color = bool(use_color) and color is not None and gt_color is not None
    if not depth_supervision and not use_color:
        raise ValueError("tracking_loss: without depth supervision the loss is the colour term (use_color_in_tracking)")
    dev = depth.device
    d = depth.detach().double().contiguous()
    v = uncertainty.detach().double().contiguous()
    gd = gt_depth.detach().float().contiguous()
    R = d.shape[0]
    loss = torch.empty((), dtype=torch.float64, device=dev)
    g_depth = torch.empty(R, dtype=torch.float64, device=dev)
    c = gc = g_color = None
    if use_color:
        c = color.detach().float().contiguous()
        gc = gt_color.detach().float().contiguous()
        g_color = torch.empty((R, 3), dtype=torch.float32, device=dev)
    with L.device_guard(dev):
        L.check(L.lib().pn_tracking_loss(C.c_void_p(d.data_ptr()), C.c_void_p(v.data_ptr()), C.c_void_p(L.ptr(c)),
                                         C.c_void_p(gd.data_ptr()), C.c_void_p(L.ptr(gc)), C.c_int64(R), int(bool(handle_dynamic)),
                                         int(use_color), C.c_float(w_color), int(bool(depth_supervision)), C.c_void_p(loss.data_ptr()),
                                         C.c_void_p(g_depth.data_ptr()), C.c_void_p(L.ptr(g_color)),
                                         C.c_void_p(L.stream_ptr(dev))), "pn_tracking_loss")
    if g_depth.dtype != depth.dtype:
        g_depth = g_depth.to(depth.dtype)
    if g_color is not None and g_color.dtype != color.dtype:
        g_color = g_color.to(color.dtype)
    return loss, g_depth, g_color


def tracking_loss(depth: torch.Tensor, uncertainty: torch.Tensor, color: Optional[torch.Tensor], gt_depth: torch.Tensor,
                  gt_color: Optional[torch.Tensor], w_color: float = 0.5, use_color: bool = True,
                  handle_dynamic: bool = True, depth_supervision: bool = True) -> torch.Tensor:
    """Tracker.py:306-330: uncertainty-weighted masked L1 depth loss (the uncertainty is detached; with
    ``handle_dynamic`` rays whose weighted residual exceeds ten times the median are dropped) plus ``w_color`` times the
    masked L1 colour loss.  Value and gradient come from one launch (the reference spends ~28 elementwise, reduction
    and sort launches on it, forward and backward).  ``depth_supervision=False`` is the fork's colour-only branch
    (Tracker.py:313-318: the same mask, the masked colour residuals alone).  Any number of rays (the fork's tracker
    passes every pixel with depth, Tracker.py:206-226).  Returns a float64 scalar."""
    if not depth.is_cuda:
        raise RuntimeError("pointnerf_slam_b200.losses.tracking_loss needs CUDA tensors (there is no CPU path)")
    if depth.shape[0] == 0:
        raise ValueError("tracking_loss needs at least one ray")
    use_color = bool(use_color) and color is not None and gt_color is not None
    if not depth_supervision and not use_color:
        raise ValueError("tracking_loss: without depth supervision the loss is the colour term (use_color_in_tracking)")
    return _TrackingLossFn.apply(depth, uncertainty, color if use_color else None, gt_depth, gt_color if use_color else None,
                                 float(w_color), use_color, bool(handle_dynamic), bool(depth_supervision))
