"""One iteration of the Tracker's pose optimisation (src/Tracker.py:253-335, ``optimize_cam_in_batch``) through this
package's public API.

    optimizer.zero_grad()                                                          :269
    c2w = get_camera_from_tensor(camera_tensor)                                    :270
    get_samples(Hedge, H-Hedge, Wedge, W-Wedge, batch_size, ...)                   :271-286
    render_batch_ray(c, decoders, rays_d, rays_o, device, stage='color', gt_depth) :301-303
    uncertainty-weighted, median-masked L1 depth loss (+ w_color * colour loss)    :306-330
    loss.backward(); optimizer.step()                                              :331-333

``TrackingIteration`` is the callable a Tracker puts inside its ``for cam_iter`` loop -- eagerly or captured once with
``graphs.GraphedStep`` (nothing in it synchronises with the host; the reference reads ``loss.item()`` every iteration).
The map is frozen (the tracker's private copies, Tracker.py:349-352): only the camera 7-vector receives a gradient.
"""
from __future__ import annotations

from typing import Optional

import torch

from .common import get_camera_from_tensor, get_samples
from .losses import tracking_loss, tracking_loss_and_grads


class TrackingIteration:
    def __init__(self, renderer, decoders, grids, depth: torch.Tensor, color: torch.Tensor, cam: torch.Tensor, H, W, fx, fy, cx, cy,
                 n_pixels: int, ignore_edge_h: int = 0, ignore_edge_w: int = 0, w_color: float = 0.5, use_color: bool = True,
                 handle_dynamic: bool = True, depth_supervision: bool = True, generator: Optional[torch.Generator] = None,
                 optimizer=None, stage: str = "color"):
        """depth (H,W) float32, colour (H,W,3): the current frame on the device; cam: the camera 7-vector
        ([qw qx qy qz tx ty tz], requires_grad).  optimizer (optional): anything with ``step()`` -- e.g.
        ``mapper.StageOptimizer({}, [], [cam])`` with the camera learning rate in group 5, or a torch optimiser."""
        self.renderer, self.decoders, self.grids = renderer, decoders, grids
        self.depth, self.color, self.cam = depth, color, cam
        self.geom = (int(H), int(W), float(fx), float(fy), float(cx), float(cy))
        self.n, self.eh, self.ew = int(n_pixels), int(ignore_edge_h), int(ignore_edge_w)
        self.w_color, self.use_color, self.handle_dynamic = float(w_color), bool(use_color), bool(handle_dynamic)
        self.depth_supervision, self.gen, self.optimizer, self.stage = bool(depth_supervision), generator, optimizer, stage
        self.device = cam.device
        self.last_indices = None

    def _render(self, indices=None):
        H, W, fx, fy, cx, cy = self.geom
        H0, H1, W0, W1 = self.eh, H - self.eh, self.ew, W - self.ew
        if indices is None:
            indices = torch.randint((H1 - H0) * (W1 - W0), (self.n,), device=self.device, generator=self.gen)
        self.last_indices = indices
        c2w = get_camera_from_tensor(self.cam)
        ro, rd, gd, gc = get_samples(H0, H1, W0, W1, self.n, H, W, fx, fy, cx, cy, c2w, self.depth, self.color, self.device,
                                     indices=indices)
        frozen = self.renderer.freeze_map
        self.renderer.freeze_map = True
        try:
            d, v, c = self.renderer.render_batch_ray(self.grids, self.decoders, rd, ro, self.device, self.stage, gt_depth=gd)
        finally:
            self.renderer.freeze_map = frozen
        return d, v, c, gd, gc

    def forward_loss(self, indices=None) -> torch.Tensor:
        """The iteration's loss as a differentiable scalar (``loss.backward()`` works as in the reference)."""
        d, v, c, gd, gc = self._render(indices)
        return tracking_loss(d, v, c, gd, gc, self.w_color, self.use_color, self.handle_dynamic, self.depth_supervision)

    def __call__(self, indices=None) -> torch.Tensor:
        # the gradient lives in ONE buffer for the life of the object (zeroed, then accumulated into by autograd): a fixed
        # address, which the optimiser's device-side tensor table and a captured CUDA graph both rely on
        if self.cam.grad is None:
            self.cam.grad = torch.zeros_like(self.cam)
        else:
            self.cam.grad.zero_()
        d, v, c, gd, gc = self._render(indices)
        loss, g_depth, g_color = tracking_loss_and_grads(d, v, c, gd, gc, self.w_color, self.use_color, self.handle_dynamic,
                                                         self.depth_supervision)
        outs, grads = [], []
        if self.depth_supervision:
            outs.append(d)
            grads.append(g_depth)
        if g_color is not None:
            outs.append(c)
            grads.append(g_color)
        torch.autograd.backward(outs, grads)       # the loss kernel's gradients seed the renderer's backward directly
        if self.optimizer is not None:
            self.optimizer.step()
        return loss
