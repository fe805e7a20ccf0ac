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

import os
from typing import Dict, List, Optional, Sequence

import torch

from . import dist as D
from . import engine as E
from .common import KeyframeBatch, get_camera_from_tensor, get_samples_multi
from .losses import mapping_loss, mapping_loss_and_grads

TRAINED_GRIDS = {"coarse": ("grid_coarse",), "middle": ("grid_middle",), "fine": ("grid_middle", "grid_fine"),
                 "color": ("grid_middle", "grid_fine", "grid_color")}


class MappingIteration:
    def __init__(self, renderer, decoders, grids: Dict[str, torch.Tensor], frames: Sequence, cams: Sequence[torch.Tensor],
                 H, W, fx, fy, cx, cy, pix_per_frame: int, stage: str = "color", w_color: float = 0.2,
                 generator: Optional[torch.Generator] = None, arena: Optional[E.GradArena] = None,
                 exchange: str = "none", optimizer=None, trained_decoders: Sequence[str] = ("color",), shared_cameras: bool = True,
                 world: Optional[int] = None):
        """frames: [(depth (H,W) f32, colour (H,W,3))] device tensors, one per keyframe; cams: camera 7-vectors (those
        with requires_grad are bundle-adjusted).  exchange: 'none' | 'sparse' | 'sparse_p2p' | 'dense' | 'overlap' |
        'arena' -- how the gradients of a sharded batch are summed over the ranks (dist.py).  shared_cameras: True when
        every rank holds the same keyframes and the RAYS are sharded (pose gradients are summed too); False when the
        KEYFRAMES are sharded (each camera's gradient stays on the rank that owns the keyframe).  world: number of ranks
        the batch is sharded over (default: the process group's size; 1 = no collective at all)."""
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
        self.world = D.world_size() if world is None else int(world)
        self._sparse = None
        self._reducer = None
        self._tail_items = self.dec_params + self.shared_cams
        self._tail_off, tail = [], 0
        for p in self._tail_items:                       # 16-byte aligned slots of the exchange's dense tail
            self._tail_off.append(tail)
            tail += (p.numel() + 3) // 4 * 4
        self.sparse_cap_frac = 0.15      # rows the exchange can carry per rank, as a fraction of all voxel rows
        self._use_sparse = exchange in ("sparse", "sparse_p2p") and self.world > 1
        if self.world > 1 and exchange in ("overlap", "arena"):
            self._reducer = D.OverlappedGradReducer(arena if exchange == "arena" else None)
        self.last_indices = None
        self._batch: Optional[KeyframeBatch] = None
        self._arena_join = None
        self._comm_pending = None
        # sparse (all-gather) exchange: start it as soon as the grid gradients are final, beside the weight-gradient kernels
        self.overlap_exchange = os.environ.get("PN_OVERLAP_EXCHANGE", "0") != "0"
        self.overlap_reserve_sms = int(os.environ.get("PN_OVERLAP_RESERVE_SMS", "16"))

    # ------------------------------------------------------------------------------------------
    def trained(self) -> List[torch.Tensor]:
        return [self.grids[k] for k in self.grid_keys] + self.dec_params + self.ba_cams

    def sample(self, indices=None):
        """The per-keyframe loop of Mapper.py:558-605 as one batched launch: camera tensors -> poses, pixel indices (one
        ``torch.randint`` call per keyframe, in keyframe order, so the generator stream is the reference's), rays, depths and
        colours of all keyframes concatenated.  `indices`: (F,n) tensor or a list of F index tensors to replay a draw."""
        H, W, fx, fy, cx, cy = self.geom
        F = len(self.frames)
        if self._batch is None:
            self._batch = KeyframeBatch(self.frames, self.device)
        if indices is None:
            idx = torch.empty((F, self.n), dtype=torch.int64, device=self.device)
            for k in range(F):
                torch.randint(H * W, (self.n,), device=self.device, generator=self.gen, out=idx[k])
        else:
            idx = indices if isinstance(indices, torch.Tensor) else torch.stack([i.to(self.device) for i in indices])
        self.last_indices = idx
        c2w = get_camera_from_tensor(torch.stack(self.cams))
        return get_samples_multi(0, H, 0, W, self.n, H, W, fx, fy, cx, cy, c2w, self._batch, self.device, indices=idx)

    def _render(self, indices=None):
        ro, rd, gd, gc = self.sample(indices)
        self.renderer.depth_max_override = D.share_depth_max(gd) if self.world > 1 else None
        try:
            depth, var, color = self.renderer.render_batch_ray(self.grids, self.decoders, rd, ro, self.device, self.stage, gt_depth=gd)
        finally:
            self.renderer.depth_max_override = None
        return depth, color, gd, gc

    def forward_loss(self, indices=None):
        """The iteration's loss as a differentiable scalar (``loss.backward()`` works as in the reference)."""
        depth, color, gd, gc = self._render(indices)
        return mapping_loss(depth, color, gd, gc, self.stage, self.w_color)

    def _backward(self, indices=None) -> torch.Tensor:
        """Forward, loss and backward.  The loss kernel hands back d loss / d depth and d loss / d colour, and the renderer's
        backward is seeded with them directly (no autograd node for the loss: three launches fewer per iteration)."""
        depth, color, gd, gc = self._render(indices)
        loss, g_depth, g_color = mapping_loss_and_grads(depth, color, gd, gc, self.stage, self.w_color)
        outs, grads = [depth], [g_depth]
        if g_color is not None:
            outs.append(color)
            grads.append(g_color)
        if self._arena_join is not None:
            torch.cuda.current_stream(self.device).wait_stream(self._arena_join)
            self._arena_join = None
        torch.autograd.backward(outs, grads)
        return loss

    @staticmethod
    def _contiguous_block(tensors, numels):
        """One flat float32 view over `tensors` if they lie back to back in memory (slots of numels[i] floats) inside the
        same storage -- the gradient arena hands out its sinks that way -- else None."""
        base = tensors[0].data_ptr()
        off = 0
        for t, n in zip(tensors, numels):
            if t.dtype != torch.float32 or t.data_ptr() != base + 4 * off or t.untyped_storage().data_ptr() != tensors[0].untyped_storage().data_ptr():
                return None
            off += n
        return torch.as_strided(tensors[0], (off,), (1,))

    def _sparse_setup(self, grid_grads):
        """(exchange object, key -> gradient block): with the gradient arena the grid gradients are ONE contiguous
        [V_total][32] block ("grids"), otherwise one block per grid."""
        for k, g in zip(self.grid_keys, grid_grads):
            if not g.is_contiguous(memory_format=torch.channels_last_3d):
                raise RuntimeError(f"{k}: sparse exchange needs the channels-last gradient the backward produces")
        order = sorted(range(len(grid_grads)), key=lambda i: grid_grads[i].data_ptr())
        flat_grids = self._contiguous_block([grid_grads[i] for i in order], [grid_grads[i].numel() for i in order])
        if self._sparse is None:
            if flat_grids is not None:
                shapes = {"grids": (flat_grids.numel() // 32, 1, 1)}
            else:
                shapes = {k: self.grids[k].shape[2:] for k in self.grid_keys}
            tail = sum((p.numel() + 3) // 4 * 4 for p in self._tail_items)
            self._sparse = D.SparseGradExchange(shapes, tail, self.device, cap_frac=self.sparse_cap_frac,
                                                mode="p2p" if self.exchange == "sparse_p2p" else "allgather", world=self.world)
            self._sparse_flat = flat_grids is not None
        if self._sparse_flat != (flat_grids is not None):
            raise RuntimeError("sparse exchange: the gradient buffers changed layout between iterations")
        return self._sparse, ({"grids": flat_grids} if flat_grids is not None else dict(zip(self.grid_keys, grid_grads)))

    def _fill_tail(self, sp):
        """Copy the decoder / shared-camera gradients into the exchange's dense tail; returns (flat_params or None,
        pgrads, cam_grads)."""
        pgrads = [p.grad for p in self.dec_params]
        flat_params = self._contiguous_block(pgrads, [(p.numel() + 3) // 4 * 4 for p in self.dec_params]) if pgrads else None
        tail = sp.tail_view()
        cam_grads = [c.grad for c in self.shared_cams]
        n_par = sum((p.numel() + 3) // 4 * 4 for p in self.dec_params)
        if flat_params is not None:
            tail[:n_par].copy_(flat_params)
        elif pgrads:
            torch._foreach_copy_([tail[o:o + p.numel()].view(p.shape) for o, p in zip(self._tail_off, self.dec_params)], pgrads)
        cam_views = [tail[o:o + c.numel()] for o, c in zip(self._tail_off[len(self.dec_params):], self.shared_cams)]
        if cam_grads:
            torch._foreach_copy_(cam_views, cam_grads)
        return flat_params, pgrads, cam_grads

    def _scatter_tail(self, summed, pgrads, cam_grads):
        outs = [summed[o:o + p.numel()].view(p.shape) for o, p in zip(self._tail_off, self._tail_items)]
        torch._foreach_copy_(pgrads + cam_grads, outs)

    def _exchange_sparse(self):
        """Sum the gradients over the ranks through dist.SparseGradExchange.  With the gradient arena the three grid
        gradients are one contiguous [V_total][32] block and the decoder gradients one flat slice, so the whole exchange
        is: pack (3 launches), tail copies, ONE all-gather (or none: P2P), apply, tail sums."""
        sp, blocks = self._sparse_setup([self.grids[k].grad for k in self.grid_keys])
        flat_params, pgrads, cam_grads = self._fill_tail(sp)
        if flat_params is not None and not cam_grads:
            sp.exchange(blocks, flat_params)
            return
        summed = torch.empty_like(sp.tail_view())
        sp.exchange(blocks, summed)
        self._scatter_tail(summed, pgrads, cam_grads)

    # -- overlapped variant: the grid rows travel while the weight-gradient kernels run -------------------------------
    def _on_grids_ready(self, g_grids) -> int:
        """engine.GRIDS_READY_HOOK: every grid gradient is final, the weight-gradient kernels have not been launched yet.
        Start the row exchange on the communication stream; returns the SMs the remaining kernels should leave free."""
        sp, blocks = self._sparse_setup([g_grids[k] for k in self.grid_keys])
        main, comm = torch.cuda.current_stream(self.device), E._side_stream(self.device, 3)
        comm.wait_stream(main)
        with torch.cuda.stream(comm):
            sp.exchange_grids(blocks)
        self._comm_pending = comm
        return self.overlap_reserve_sms

    def _finish_overlapped(self):
        """After the backward: exchange the dense tail (decoder + shared-camera gradients), then join the row exchange."""
        sp = self._sparse
        flat_params, pgrads, cam_grads = self._fill_tail(sp)
        if flat_params is not None and not cam_grads:
            sp.exchange_tail(flat_params)
        else:
            summed = torch.empty_like(sp.tail_view())
            sp.exchange_tail(summed)
            self._scatter_tail(summed, pgrads, cam_grads)
        torch.cuda.current_stream(self.device).wait_stream(self._comm_pending)
        self._comm_pending = None

    def __call__(self, indices=None) -> torch.Tensor:
        if self.arena is not None:
            E.GRAD_ARENA = self.arena
            if self.arena.offset:
                # the 46 MiB memset of the gradient sinks runs beside the forward (which takes nothing from the arena) on a
                # side stream and is joined right before the backward
                main, side = torch.cuda.current_stream(self.device), E._side_stream(self.device, 2)
                side.wait_stream(main)
                with torch.cuda.stream(side):
                    self.arena.reset()
                self._arena_join = side
            else:
                self.arena.reset()
        try:
            if self._reducer is not None and self.world > 1:
                with self._reducer:
                    loss = self._backward(indices)
                self._reducer.finish({k: self.grids[k] for k in self.grid_keys}, decoders=dict(self.dec_modules),
                                     others=[c.grad for c in self.shared_cams])
            elif self._use_sparse and self.exchange == "sparse" and self.overlap_exchange and E.SM_SPLIT and E.PARALLEL_BACKWARD:
                E.GRIDS_READY_HOOK = self._on_grids_ready
                try:
                    loss = self._backward(indices)
                finally:
                    E.GRIDS_READY_HOOK = None
                if self._comm_pending is not None:
                    self._finish_overlapped()
                else:                       # the backward did not take the deferred path (no weight gradients asked for)
                    self._exchange_sparse()
            else:
                loss = self._backward(indices)
                if self._use_sparse:
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
