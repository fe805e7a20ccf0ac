"""Data-parallel mapping over the GPUs of one box.

One process per GPU.  Rays shard across ranks with no data-path collective;
grids, decoders and poses are replicated.  Two exchanges per iteration:

  * ``share_depth_max``   all-reduce(MAX) of one float: ``render_batch_ray``
    couples all rays of a batch through ``max(gt_depth)``
    (src/utils/Renderer.py:112,149), so the batch-global value must be shared
    before the z-values are placed;
  * ``allreduce_gradients`` all-reduce(SUM) of grid, decoder and pose
    gradients.  The reference's losses are plain sums (src/Mapper.py:641-646),
    so summed shard gradients equal the single-GPU gradient -- no rescaling.

The collectives run on ``torch.distributed`` (NCCL over NVLink on the GPU box,
gloo in the CPU tests).
"""
from __future__ import annotations

import os
from typing import Iterable, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def init_from_env(backend: Optional[str] = None) -> Tuple[int, int, int]:
    """(rank, world_size, local_rank) from torchrun's environment; initialises
    the default process group when world_size > 1."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {}
        if backend == "nccl":
            torch.cuda.set_device(local)
            kw["device_id"] = torch.device("cuda", local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kw)
    return rank, world, local


def world_size() -> int:
    return dist.get_world_size() if dist.is_initialized() else 1


def shard_bounds(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [begin, end) slice of n units for this rank (sizes differ by <= 1)."""
    base, rem = divmod(n, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def share_depth_max(gt_depth: torch.Tensor) -> torch.Tensor:
    """1-element tensor holding max(gt_depth) over every rank's shard."""
    m = gt_depth.detach().reshape(-1).float().max().reshape(1) if gt_depth.numel() else \
        torch.full((1,), float("-inf"), device=gt_depth.device)
    if world_size() > 1:
        dist.all_reduce(m, op=dist.ReduceOp.MAX)
    return m


def allreduce_gradients(tensors: Iterable[Optional[torch.Tensor]], small_bytes: int = 1 << 20) -> None:
    """In-place SUM all-reduce of gradient tensors.  Large tensors (the feature
    grids) go individually; small ones (decoder parameters, poses) are packed
    into one bucket so that launch latency, not link count, sizes the calls."""
    if world_size() == 1:
        return
    big: List[torch.Tensor] = []
    small: List[torch.Tensor] = []
    for t in tensors:
        if t is None:
            continue
        (big if t.numel() * t.element_size() >= small_bytes else small).append(t)
    works = []
    copies = []
    for t in big:
        # channels-last grid gradients are dense in memory: reduce the storage order as is
        flat = t.permute(0, 2, 3, 4, 1) if (t.dim() == 5 and not t.is_contiguous()) else t
        if not flat.is_contiguous():
            tmp = flat.contiguous()
            copies.append((flat, tmp))
            flat = tmp
        works.append(dist.all_reduce(flat, op=dist.ReduceOp.SUM, async_op=True))
    if small:
        dtype = small[0].dtype
        bucket = torch.cat([t.reshape(-1).to(dtype) for t in small])        # one launch
        dist.all_reduce(bucket, op=dist.ReduceOp.SUM)
        views, off = [], 0
        for t in small:
            views.append(bucket[off:off + t.numel()].view(t.shape))
            off += t.numel()
        torch._foreach_copy_(small, views)                                   # one (fused) launch back
    for w in works:
        w.wait()
    for dst, tmp in copies:
        dst.copy_(tmp)


class SymmetricGradArena:
    """An ``engine.GradArena`` whose buffer lives in symmetric memory (every rank maps every peer's copy and the
    NVSwitch multicast address).  ``all_reduce_used`` sums the used part across ranks with the switch's in-fabric
    reduction: each rank owns 1/N of the range, pulls it with ``multimem.ld_reduce`` (the NVSwitch adds the N
    replicas) and pushes the sum back with ``multimem.st`` (broadcast) -- PyTorch's ``symm_mem.multimem_all_reduce_``
    kernel over NVLink peer memory, no NCCL ring.  Needs an NVSwitch box with multicast support."""

    def __new__(cls, numel: int, device, group=None):
        from .engine import GradArena
        import torch.distributed._symmetric_memory as symm

        class _Arena(GradArena):
            def __init__(self, numel, device, group):
                group = group or dist.group.WORLD
                self.group_name = group.group_name
                n = (int(numel) + 4095) // 4096 * 4096
                self.buf = symm.empty(n, dtype=torch.float32, device=device)
                self.handle = symm.rendezvous(self.buf, self.group_name)
                self.buf.zero_()
                self.offset = 0

            def all_reduce_used(self):
                n = min((self.offset + 4095) // 4096 * 4096, self.buf.numel())
                torch.ops.symm_mem.multimem_all_reduce_(self.buf[:n], "sum", self.group_name)

        return _Arena(numel, device, group)


def mapping_gradients(params: Sequence[torch.Tensor]) -> List[Optional[torch.Tensor]]:
    return [p.grad for p in params]


class OverlappedGradReducer:
    """Start the SUM all-reduce of each gradient the moment the backward has produced it.

    Installed as ``engine.GRAD_READY_HOOK`` for the duration of one ``loss.backward()``:
    the colour grid's gradient is reduced over NVLink while the fine and middle decoder
    kernels still run, and so on; only the last gradient's reduction is exposed.
    ``finish()`` waits for the collectives and makes sure every leaf's ``.grad`` holds the
    reduced values (autograd may have kept the buffer itself or a copy of it).

        red = OverlappedGradReducer()
        with red:                      # installs / removes the hook
            loss.backward()
        red.finish({"grid_color": grids["grid_color"], ...}, decoders={"color": model.color_decoder},
                   others=[cam.grad for cam in cams])
    """

    def __init__(self, arena=None, reserve_sms: int = 0):
        """With ``arena`` (an engine.GradArena that the backward allocates from) the hook only
        records the buffers and ``finish`` issues ONE all-reduce over the used part of the arena.
        ``reserve_sms`` > 0: from the first started all-reduce to the end of the backward the
        persistent decoder kernels leave that many SMs free (``pn_reserve_sms``), so that NCCL's
        CTAs run beside them instead of in the gaps between them."""
        self.items = []    # (key, tensor or list, flat buffer reduced, work)
        self.arena = arena
        self.reserve_sms = int(reserve_sms)
        self._reserved = False

    def __enter__(self):
        from . import engine
        self._prev = engine.GRAD_READY_HOOK
        engine.GRAD_READY_HOOK = self._ready
        return self

    def __exit__(self, *exc):
        from . import engine
        engine.GRAD_READY_HOOK = self._prev
        if self._reserved:
            from . import _lib as L
            L.lib().pn_reserve_sms(0)
            self._reserved = False
        return False

    def _reserve(self):
        if self.reserve_sms > 0 and not self._reserved:
            from . import _lib as L
            L.lib().pn_reserve_sms(self.reserve_sms)
            self._reserved = True

    def _ready(self, key, grad):
        if world_size() == 1:
            return
        if self.arena is not None:
            first = grad[0] if isinstance(grad, (list, tuple)) else grad
            in_arena = self.arena.owns(first)
            self.items.append((key, list(grad) if isinstance(grad, (list, tuple)) else grad, "arena" if in_arena else None, None))
            return
        if isinstance(grad, (list, tuple)):           # parameter gradients: views of one flat buffer
            # engine.zeros_like_flat hands over the exact slice it carved (never ``_base``: with a gradient arena
            # installed that is the whole arena, and reducing it would sum every other sink a second time)
            flat = getattr(grad, "flat", None)
            if flat is not None and not flat.is_contiguous():
                flat = None
            if flat is not None:
                work = dist.all_reduce(flat, op=dist.ReduceOp.SUM, async_op=True)
                self.items.append((key, list(grad), flat, work))
            else:
                self.items.append((key, list(grad), None, None))
            return
        view = grad.permute(0, 2, 3, 4, 1) if (grad.dim() == 5 and not grad.is_contiguous()) else grad
        if view.is_contiguous():
            work = dist.all_reduce(view, op=dist.ReduceOp.SUM, async_op=True)
            self._reserve()
            self.items.append((key, grad, view, work))
        else:
            self.items.append((key, grad, None, None))

    def finish(self, grids=None, decoders=None, others=()):
        """Wait, then reconcile ``.grad`` of the leaves.  ``grids``: key -> leaf tensor;
        ``decoders``: decoder name -> module whose parameters were trained; ``others``: extra
        gradient tensors (poses ...) reduced here in one bucket."""
        late = []
        if self.arena is not None and world_size() > 1 and self.arena.offset:
            if hasattr(self.arena, "all_reduce_used"):        # symmetric-memory arena: NVLS multimem all-reduce
                self.arena.all_reduce_used()
            else:
                dist.all_reduce(self.arena.used(), op=dist.ReduceOp.SUM)
        for key, grad, buf, work in self.items:
            if work is not None:
                work.wait()
            elif buf == "arena":
                work = True                            # reduced by the arena all-reduce above
            if isinstance(grad, list):
                dec = (decoders or {}).get(key[1])
                leaves = None
                if dec is not None:
                    from .engine import grid_mlp_tensors, coarse_mlp_tensors
                    leaves = grid_mlp_tensors(dec) if hasattr(dec, "fc_c") else coarse_mlp_tensors(dec)
                if work is None:
                    late += [p.grad for p in (leaves or []) if p.grad is not None]
                elif leaves is not None:
                    dst = [p.grad for p, g in zip(leaves, grad) if p.grad is not None and p.grad.data_ptr() != g.data_ptr()]
                    src = [g for p, g in zip(leaves, grad) if p.grad is not None and p.grad.data_ptr() != g.data_ptr()]
                    if dst:
                        torch._foreach_copy_(dst, src)
            else:
                leaf = (grids or {}).get(key)
                if work is None:
                    if leaf is not None and leaf.grad is not None:
                        late.append(leaf.grad)
                elif leaf is not None and leaf.grad is not None and leaf.grad.data_ptr() != grad.data_ptr():
                    leaf.grad.copy_(grad)
        self.items = []
        allreduce_gradients(list(late) + [t for t in others if t is not None])
