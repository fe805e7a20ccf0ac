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


class SparseGradExchange:
    """Gradient exchange of the data-parallel mapper that moves only what a mapping iteration touches.

    A 5,000-ray batch leaves gradient in a few per cent of the voxel rows (one row = the 32 channels of a voxel = 128
    bytes) of the middle / fine / colour grids, yet a dense all-reduce moves all 46 MiB of them.  Here every rank

      1. packs its touched rows (``pn_sparse_rows_pack``: a bitmap of V/32 words, a per-word prefix, the rows in
         ascending order) and its dense "tail" (decoder + pose gradients, ~60k floats) into ONE send buffer,
      2. exchanges the buffers -- ``mode="allgather"``: one NCCL all-gather of fixed-size buffers;
         ``mode="p2p"``: no collective at all, the buffers live in symmetric memory and step 3 reads the peers'
         copies straight over NVLink (two device-side barriers fence the reads),
      3. rebuilds the summed gradient in place (``pn_sparse_rows_apply`` / ``pn_dense_sum``): every touched row is the
         sum of the ranks' rows IN RANK ORDER, so all ranks hold bit-identical gradients (what a replicated Adam step
         needs) and the result does not depend on timing.

    Row capacity per grid is fixed (``cap_frac`` of its rows) so that buffer sizes never depend on the data and the
    whole exchange can sit inside a captured CUDA graph; a rank that touches more rows than that raises the device
    counter ``overflow`` (checked by ``check_overflow()`` outside the graph; the caller then falls back to
    ``allreduce_gradients``).  No reference counterpart: the reference is single-GPU.
    """

    def __init__(self, grid_shapes, tail_numel: int, device, cap_frac=0.25, mode: str = "allgather", group=None, world=None):
        from . import _lib as L
        self.L = L
        self.device = torch.device(device)
        self.world = world_size() if world is None else int(world)
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.mode = mode
        self.keys = list(grid_shapes)
        self.V = {k: int(grid_shapes[k][0]) * int(grid_shapes[k][1]) * int(grid_shapes[k][2]) for k in self.keys}
        frac = cap_frac if isinstance(cap_frac, dict) else {k: cap_frac for k in self.keys}
        self.cap = {k: max(32, (int(self.V[k] * frac[k]) + 31) // 32 * 32) for k in self.keys}
        self.tail_numel = int(tail_numel)
        off, self.off = 0, {}
        for k in self.keys:                       # float32 units; every section starts on a 128-byte boundary
            nw = (self.V[k] + 31) // 32
            nwp = (nw + 31) // 32 * 32
            self.off[k] = (off, off + nwp, off + 2 * nwp)        # bitmap, prefix, rows
            off += 2 * nwp + self.cap[k] * 32
        self.off_tail = off
        off += (self.tail_numel + 31) // 32 * 32
        self.length = off
        if mode == "p2p" and self.world > 1:
            import torch.distributed._symmetric_memory as symm
            group = group or dist.group.WORLD
            self.group_name = group.group_name
            self.send = symm.empty(self.length, dtype=torch.float32, device=self.device)
            self.handle = symm.rendezvous(self.send, self.group_name)
            self.send.zero_()
            base = [int(self.handle.buffer_ptrs[r]) for r in range(self.world)]
            self.recv = None
        else:
            self.mode = "allgather"
            self.send = torch.zeros(self.length, dtype=torch.float32, device=self.device)
            self.recv = torch.zeros(self.world * self.length, dtype=torch.float32, device=self.device)
            base = [self.recv.data_ptr() + 4 * r * self.length for r in range(self.world)]
        self.base = base
        nscr = max((v + 1023) // 1024 for v in self.V.values()) + 1
        self.scratch = torch.zeros(nscr, dtype=torch.int32, device=self.device)
        self.count = torch.zeros(len(self.keys), dtype=torch.int64, device=self.device)
        self.overflow = torch.zeros(1, dtype=torch.int32, device=self.device)
        import ctypes as C
        PtrArr = C.c_void_p * self.world
        self._ptrs = {}
        for k in self.keys:
            ob, op, orow = self.off[k]
            self._ptrs[k] = (PtrArr(*[b + 4 * ob for b in base]), PtrArr(*[b + 4 * op for b in base]), PtrArr(*[b + 4 * orow for b in base]))
        self._tail_ptrs = PtrArr(*[b + 4 * self.off_tail for b in base])

    def bytes_per_rank(self) -> int:
        return 4 * self.length

    def tail_view(self) -> torch.Tensor:
        """The send buffer's dense tail: write the decoder / pose gradients here before ``exchange``."""
        return self.send[self.off_tail:self.off_tail + self.tail_numel]

    def exchange(self, grid_grads, tail_out=None) -> None:
        """grid_grads: key -> dense channels-last gradient (any view of a [V][32] float32 block), summed in place
        across ranks.  tail_out: flat float32 tensor of ``tail_numel`` elements that receives the summed tail."""
        import ctypes as C
        L, lib = self.L, self.L.lib()
        st = C.c_void_p(L.stream_ptr(self.device))
        with L.device_guard(self.device):
            for i, k in enumerate(self.keys):
                g = grid_grads[k]
                ob, op, orow = self.off[k]
                sp = self.send.data_ptr()
                L.check(lib.pn_sparse_rows_pack(C.c_void_p(g.data_ptr()), C.c_int64(self.V[k]), C.c_void_p(sp + 4 * ob),
                                                C.c_void_p(sp + 4 * op), C.c_void_p(sp + 4 * orow), C.c_int64(self.cap[k]),
                                                C.c_void_p(self.scratch.data_ptr()), C.c_void_p(self.count[i:].data_ptr()),
                                                C.c_void_p(self.overflow.data_ptr()), st), "pn_sparse_rows_pack")
            if self.world > 1:
                if self.mode == "p2p":
                    self.handle.barrier(channel=0)                # every rank's buffer is complete
                else:
                    dist.all_gather_into_tensor(self.recv, self.send)
            else:
                self.recv.copy_(self.send)
            for k in self.keys:
                bm, pf, rows = self._ptrs[k]
                L.check(lib.pn_sparse_rows_apply(C.c_void_p(grid_grads[k].data_ptr()), C.c_int64(self.V[k]), self.world, bm, pf, rows,
                                                 C.c_int64(self.cap[k]), st), "pn_sparse_rows_apply")
            if tail_out is not None and self.tail_numel:
                L.check(lib.pn_dense_sum(C.c_void_p(tail_out.data_ptr()), C.c_int64(self.tail_numel), self.world, self._tail_ptrs, st),
                        "pn_dense_sum")
            if self.world > 1 and self.mode == "p2p":
                self.handle.barrier(channel=1)                    # nobody overwrites a buffer a peer still reads

    # -- the same exchange in two parts (allgather mode), for a caller that overlaps the first with other work ----------
    def exchange_grids(self, grid_grads) -> None:
        """Pack, all-gather and apply the touched rows of the grid gradients only (the tail section of the buffer travels
        along unused).  Everything is enqueued on the CURRENT stream."""
        import ctypes as C
        if self.mode != "allgather":
            raise RuntimeError("exchange_grids / exchange_tail: allgather mode only")
        L, lib = self.L, self.L.lib()
        st = C.c_void_p(L.stream_ptr(self.device))
        with L.device_guard(self.device):
            sp = self.send.data_ptr()
            for i, k in enumerate(self.keys):
                ob, op, orow = self.off[k]
                L.check(lib.pn_sparse_rows_pack(C.c_void_p(grid_grads[k].data_ptr()), C.c_int64(self.V[k]), C.c_void_p(sp + 4 * ob),
                                                C.c_void_p(sp + 4 * op), C.c_void_p(sp + 4 * orow), C.c_int64(self.cap[k]),
                                                C.c_void_p(self.scratch.data_ptr()), C.c_void_p(self.count[i:].data_ptr()),
                                                C.c_void_p(self.overflow.data_ptr()), st), "pn_sparse_rows_pack")
            if self.world > 1:
                dist.all_gather_into_tensor(self.recv, self.send)
            else:
                self.recv.copy_(self.send)
            for k in self.keys:
                bm, pf, rows = self._ptrs[k]
                L.check(lib.pn_sparse_rows_apply(C.c_void_p(grid_grads[k].data_ptr()), C.c_int64(self.V[k]), self.world, bm, pf, rows,
                                                 C.c_int64(self.cap[k]), st), "pn_sparse_rows_apply")

    def exchange_tail(self, tail_out: torch.Tensor) -> None:
        """Sum the dense tail (written to ``tail_view()`` beforehand) over the ranks in rank order into ``tail_out``: one small
        all-gather of the tails + pn_dense_sum."""
        import ctypes as C
        if self.mode != "allgather":
            raise RuntimeError("exchange_grids / exchange_tail: allgather mode only")
        if not self.tail_numel:
            return
        L, lib = self.L, self.L.lib()
        n = (self.tail_numel + 31) // 32 * 32
        if getattr(self, "_tail_recv", None) is None:
            self._tail_recv = torch.zeros(self.world * n, dtype=torch.float32, device=self.device)
            PtrArr = C.c_void_p * self.world
            self._tail_recv_ptrs = PtrArr(*[self._tail_recv.data_ptr() + 4 * r * n for r in range(self.world)])
        mine = self.send[self.off_tail:self.off_tail + n]
        if self.world > 1:
            dist.all_gather_into_tensor(self._tail_recv, mine)
        else:
            self._tail_recv.copy_(mine)
        with L.device_guard(self.device):
            L.check(lib.pn_dense_sum(C.c_void_p(tail_out.data_ptr()), C.c_int64(self.tail_numel), self.world, self._tail_recv_ptrs,
                                     C.c_void_p(L.stream_ptr(self.device))), "pn_dense_sum")

    def check_overflow(self) -> None:
        n = int(self.overflow.item())
        if n:
            self.overflow.zero_()
            raise RuntimeError(f"SparseGradExchange: {n} touched voxel rows did not fit the send buffer "
                               f"(capacities {self.cap}); raise cap_frac or fall back to allreduce_gradients")


def gather_shards(x: torch.Tensor, total: int, world: Optional[int] = None) -> torch.Tensor:
    """Concatenation over the ranks of their contiguous shards of a length-`total` leading dimension (shard r =
    ``shard_bounds(total, r, world)``).  One all-gather of equal-size (padded) buffers; every rank gets the whole."""
    world = world_size() if world is None else world
    if world == 1:
        return x
    n_max = (total + world - 1) // world
    buf = x.new_zeros((n_max,) + tuple(x.shape[1:]))
    buf[:x.shape[0]].copy_(x)
    out = x.new_empty((world * n_max,) + tuple(x.shape[1:]))
    dist.all_gather_into_tensor(out, buf)
    if total == world * n_max:
        return out
    parts = []
    for r in range(world):
        b, e = shard_bounds(total, r, world)
        parts.append(out[r * n_max:r * n_max + (e - b)])
    return torch.cat(parts, 0)


def render_img_sharded(renderer, c, decoders, c2w, device, stage, gt_depth, rank: Optional[int] = None, world: Optional[int] = None):
    """``Renderer.render_img`` (src/utils/Renderer.py:205-260) with the rays of every chunk sharded over the ranks and the
    outputs all-gathered: every rank returns the full (H,W) depth, uncertainty and (H,W,3) colour.

    The reference renders in chunks of ``ray_batch_size`` rays and clamps `far` with the maximum depth of the CHUNK
    (Renderer.py:112 inside the loop of :237); each rank therefore takes its slice of every chunk and is told the
    chunk's maximum (a local reduction: every rank holds the whole depth image), so the result equals the one-GPU
    render ray for ray."""
    from .common import get_rays
    from .renderer import batch_depth_max
    world = world_size() if world is None else world
    rank = (dist.get_rank() if dist.is_initialized() else 0) if rank is None else rank
    if world == 1:
        return renderer.render_img(c, decoders, c2w, device, stage, gt_depth=gt_depth)
    with torch.no_grad():
        Hh, Ww = renderer.H, renderer.W
        rays_o, rays_d = get_rays(Hh, Ww, renderer.fx, renderer.fy, renderer.cx, renderer.cy, c2w, device)
        rays_o, rays_d, gd = rays_o.reshape(-1, 3), rays_d.reshape(-1, 3), gt_depth.reshape(-1)
        n, bs = rays_d.shape[0], renderer.ray_batch_size
        ds, us, cs = [], [], []
        for i in range(0, n, bs):
            m = min(bs, n - i)
            b, e = shard_bounds(m, rank, world)
            renderer.depth_max_override = batch_depth_max(gd[i:i + m].float().contiguous())
            try:
                d, u, col = renderer.render_batch_ray(c, decoders, rays_d[i + b:i + e], rays_o[i + b:i + e], device, stage,
                                                      gt_depth=gd[i + b:i + e])
            finally:
                renderer.depth_max_override = None
            ds.append(d.double()); us.append(u.double()); cs.append(col)
        # one all-gather: [depth | uncertainty | colour] as float64 columns of this rank's rays (chunk-major)
        mine = torch.cat([torch.cat(ds)[:, None], torch.cat(us)[:, None], torch.cat(cs).double()], 1)
        per_rank = [sum(shard_bounds(min(bs, n - i), r, world)[1] - shard_bounds(min(bs, n - i), r, world)[0] for i in range(0, n, bs))
                    for r in range(world)]
        n_max = max(per_rank)
        buf = mine.new_zeros((n_max, 5))
        buf[:mine.shape[0]].copy_(mine)
        out = mine.new_empty((world * n_max, 5))
        dist.all_gather_into_tensor(out, buf)
        # reassemble in ray order: chunk by chunk, rank by rank
        pieces, off = [], [0] * world
        for i in range(0, n, bs):
            m = min(bs, n - i)
            for r in range(world):
                b, e = shard_bounds(m, r, world)
                pieces.append(out[r * n_max + off[r]:r * n_max + off[r] + (e - b)])
                off[r] += e - b
        full = torch.cat(pieces, 0)
        return full[:, 0].reshape(Hh, Ww), full[:, 1].reshape(Hh, Ww), full[:, 2:5].float().reshape(Hh, Ww, 3)
