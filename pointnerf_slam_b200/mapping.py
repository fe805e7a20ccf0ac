"""One iteration of the Mapper's hot loop (src/Mapper.py:551-662) through this package's public API.

    for frame in optimize_frame:  c2w = get_camera_from_tensor(camera_tensor)      :558-590
                                  get_samples(0, H, 0, W, pixs_per_image, ...)       :592-600
    cat rays                                                                          :602-605
    render_batch_ray(c, decoders, rays_d, rays_o, device, stage, gt_depth)            :623-624
    loss = masked L1 depth (+ w_color * L1 colour in stage 'color')                   :628-646
    loss.backward()                                                                   :657
    [data-parallel: gradient exchange]                                                (no reference counterpart)
    optimizer.step()                                                                  :658

``MappingIteration`` is the callable a Mapper puts inside its ``for joint_iter`` loop -- and what ``bench.py`` times and
``tests/`` check: the same object, eagerly or captured once into a CUDA graph (``graphs.GraphedStep``).  Nothing in it
synchronises with the host.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import torch

from . import dist as D
from . import engine as E
from .common import get_camera_from_tensor, get_samples
from .losses import mapping_loss

TRAINED_GRIDS = {"coarse": ("grid_coarse",), "middle": ("grid_middle",), "fine": ("grid_middle", "grid_fine"),
                 "color": ("grid_middle", "grid_fine", "grid_color")}


class MappingIteration:
    def __init__(self, renderer, decoders, grids: Dict[str, torch.Tensor], frames: Sequence, cams: Sequence[torch.Tensor],
                 H, W, fx, fy, cx, cy, pix_per_frame: int, stage: str = "color", w_color: float = 0.2,
                 generator: Optional[torch.Generator] = None, arena: Optional[E.GradArena] = None,
                 exchange: str = "none", optimizer=None, trained_decoders: Sequence[str] = ("color",), shared_cameras: bool = True):
        """frames: [(depth (H,W) f32, colour (H,W,3))] device tensors, one per keyframe; cams: camera 7-vectors (those
        with requires_grad are bundle-adjusted).  exchange: 'none' | 'sparse' | 'sparse_p2p' | 'dense' | 'overlap' |
        'arena' -- how the gradients of a sharded batch are summed over the ranks (dist.py).  shared_cameras: True when
        every rank holds the same keyframes and the RAYS are sharded (pose gradients are summed too); False when the
        KEYFRAMES are sharded (each camera's gradient stays on the rank that owns the keyframe)."""
        self.renderer, self.decoders, self.grids, self.frames, self.cams = renderer, decoders, grids, list(frames), list(cams)
        self.geom = (int(H), int(W), float(fx), float(fy), float(cx), float(cy))
        self.n, self.stage, self.w_color, self.gen = int(pix_per_frame), stage, float(w_color), generator
        self.device = self.cams[0].device
        self.arena, self.exchange, self.optimizer = arena, exchange, optimizer
        self.grid_keys = [k for k in TRAINED_GRIDS[stage] if grids[k].requires_grad]
        self.dec_modules = {name: getattr(decoders, name + "_decoder") for name in trained_decoders
                            if any(p.requires_grad for p in getattr(decoders, name + "_decoder").parameters())}
        self.dec_params = [p for m in self.dec_modules.values() for p in m.parameters() if p.requires_grad]
        self.ba_cams = [c for c in self.cams if c.requires_grad]
        self.shared_cams = [c for c in self.ba_cams] if shared_cameras else []
        self.world = D.world_size()
        self._sparse = None
        self._reducer = None
        self._tail_items = self.dec_params + self.shared_cams
        self._tail_off, tail = [], 0
        for p in self._tail_items:                       # 16-byte aligned slots of the exchange's dense tail
            self._tail_off.append(tail)
            tail += (p.numel() + 3) // 4 * 4
        if self.world > 1 or exchange.startswith("sparse"):
            if exchange in ("sparse", "sparse_p2p"):
                shapes = {k: grids[k].shape[2:] for k in self.grid_keys}
                self._sparse = D.SparseGradExchange(shapes, tail, self.device, mode="p2p" if exchange == "sparse_p2p" else "allgather")
            elif exchange in ("overlap", "arena"):
                self._reducer = D.OverlappedGradReducer(arena if exchange == "arena" else None)
        self.last_indices: List[torch.Tensor] = []

    # ------------------------------------------------------------------------------------------
    def trained(self) -> List[torch.Tensor]:
        return [self.grids[k] for k in self.grid_keys] + self.dec_params + self.ba_cams

    def sample(self, indices: Optional[Sequence[torch.Tensor]] = None):
        H, W, fx, fy, cx, cy = self.geom
        ro, rd, gd, gc = [], [], [], []
        self.last_indices = []
        for k, (depth, color) in enumerate(self.frames):
            c2w = get_camera_from_tensor(self.cams[k])
            idx = indices[k] if indices is not None else torch.randint(H * W, (self.n,), device=self.device, generator=self.gen)
            self.last_indices.append(idx)
            o, d, dd, cc = get_samples(0, H, 0, W, self.n, H, W, fx, fy, cx, cy, c2w, depth, color, self.device, indices=idx)
            ro.append(o); rd.append(d); gd.append(dd); gc.append(cc)
        return torch.cat(ro), torch.cat(rd), torch.cat(gd), torch.cat(gc)

    def forward_loss(self, indices=None):
        ro, rd, gd, gc = self.sample(indices)
        self.renderer.depth_max_override = D.share_depth_max(gd) if self.world > 1 else None
        try:
            depth, var, color = self.renderer.render_batch_ray(self.grids, self.decoders, rd, ro, self.device, self.stage, gt_depth=gd)
        finally:
            self.renderer.depth_max_override = None
        return mapping_loss(depth, color, gd, gc, self.stage, self.w_color)

    def _exchange_sparse(self):
        sp = self._sparse
        tail = sp.tail_view()
        views = [tail[o:o + p.numel()].view(p.shape) for o, p in zip(self._tail_off, self._tail_items)]
        grads = [p.grad for p in self._tail_items]
        torch._foreach_copy_(views, grads)
        grid_grads = {}
        for k in self.grid_keys:
            g = self.grids[k].grad
            if not g.is_contiguous(memory_format=torch.channels_last_3d):
                raise RuntimeError(f"{k}: sparse exchange needs the channels-last gradient the backward produces")
            grid_grads[k] = g
        summed = torch.empty_like(tail)
        sp.exchange(grid_grads, summed)
        outs = [summed[o:o + p.numel()].view(p.shape) for o, p in zip(self._tail_off, self._tail_items)]
        torch._foreach_copy_(grads, outs)

    def __call__(self, indices=None) -> torch.Tensor:
        if self.arena is not None:
            E.GRAD_ARENA = self.arena
            self.arena.reset()
        try:
            loss = self.forward_loss(indices)
            if self._reducer is not None and self.world > 1:
                with self._reducer:
                    loss.backward()
                self._reducer.finish({k: self.grids[k] for k in self.grid_keys}, decoders=dict(self.dec_modules),
                                     others=[c.grad for c in self.shared_cams])
            else:
                loss.backward()
                if self._sparse is not None:
                    self._exchange_sparse()
                elif self.world > 1 and self.exchange == "dense":
                    D.allreduce_gradients([self.grids[k].grad for k in self.grid_keys] + [p.grad for p in self._tail_items])
        finally:
            if self.arena is not None:
                E.GRAD_ARENA = None
        if self.optimizer is not None:
            self.optimizer.step()
        return loss

    def zero_grad(self) -> None:
        for t in self.trained():
            t.grad = None
