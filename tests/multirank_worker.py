"""Worker of tests/test_gpu_multirank.py (launched by torchrun, one rank per GPU): one mapping iteration sharded over the
ranks under a given gradient-exchange mode; rank 0 also runs the un-sharded iteration on its own GPU and writes the
comparison as JSON."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench as B  # noqa: E402


def main(out_path, exchange, sharding):
    from pointnerf_slam_b200 import dist as D
    from pointnerf_slam_b200 import engine as E
    from pointnerf_slam_b200.mapping import MappingIteration
    import torch.distributed as dist
    rank, world, local = D.init_from_env()
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    overlap = exchange == "sparse_overlap"    # the row exchange starts inside the backward, beside the weight-gradient kernels
    E.PARALLEL_BACKWARD = overlap
    pix = 400
    shared = sharding == "rays"
    w = B.build_mapping(dev, 0 if shared else rank, world, pix, "sparse" if overlap else exchange, with_optimizer=False,
                        shared_cameras=shared)
    w.iteration.overlap_exchange = overlap
    g = torch.Generator().manual_seed(100 + rank)
    idx = [torch.randint(B.H * B.W, (pix,), generator=g).to(dev) for _ in range(B.N_KEYFRAMES)]
    for rep in range(2):                      # twice: buffers of the exchange are reused across iterations
        w.iteration.zero_grad()
        loss = w.iteration(idx)
    torch.cuda.synchronize()
    if w.iteration._sparse is not None:
        w.iteration._sparse.check_overflow()
    mine = {k: w.grids[k].grad.detach().clone() for k in w.iteration.grid_keys}
    params = [p.grad.detach().clone() for p in w.iteration.dec_params]
    cams = [c.grad.detach().clone() for c in w.iteration.ba_cams]
    # bit-identical replicas: every rank must hold exactly the same summed gradient
    same = True
    for t in list(mine.values()) + params + (cams if shared else []):
        ref = t.clone()
        dist.broadcast(ref, 0)
        same = same and bool(torch.equal(ref, t))
    flag = torch.tensor([1 if same else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    # everything rank 0 needs for the un-sharded iteration
    all_idx = [torch.empty_like(torch.stack(idx)) for _ in range(world)]
    dist.all_gather(all_idx, torch.stack(idx))
    res = None
    if rank == 0:
        E.GRAD_ARENA = None
        E.PARALLEL_BACKWARD = False
        ws = [B.build_mapping(dev, 0 if shared else r, 1, pix, "none", with_optimizer=False, arena=False) for r in range(world)]
        w0 = ws[0]
        if shared:    # one batch: frame k gets the pixels of every rank
            frames, cams0 = w0.frames, w0.cams
            full_idx = [torch.cat([all_idx[r][k] for r in range(world)]) for k in range(B.N_KEYFRAMES)]
            n_pix = pix * world
        else:         # world x 5 keyframes
            frames = [f for x in ws for f in x.frames]
            cams0 = [c for x in ws for c in x.cams]
            full_idx = [all_idx[r][k] for r in range(world) for k in range(B.N_KEYFRAMES)]
            n_pix = pix
        it = MappingIteration(w0.renderer, w0.model, w0.grids, frames, cams0, B.H, B.W, B.FX, B.FY, B.CX, B.CY, n_pix, "color", B.W_COLOR,
                              world=1)
        it(full_idx)
        torch.cuda.synchronize()
        rel = lambda a, b: float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))
        res = {"exchange": exchange, "sharding": sharding, "replicas_bit_identical": bool(flag.item()), "loss": float(loss)}
        res["grids"] = {k: rel(mine[k], w0.grids[k].grad) for k in mine}
        res["params"] = max(rel(a, p.grad) for a, p in zip(params, it.dec_params))
        ref_cams = [c.grad for c in it.ba_cams]
        res["cams"] = max(rel(a, b) for a, b in zip(cams, ref_cams[:len(cams)]))
        res["touched"] = {k: int((mine[k].abs().sum(1) > 0).sum()) for k in mine}
        json.dump(res, open(out_path, "w"))
    dist.barrier()
    torch.cuda.synchronize()
    sys.stdout.flush()
    os._exit(0)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3])
