"""Multi-rank host logic on CPU (gloo, world_size 2): ray sharding, the shared
batch depth maximum and the SUM all-reduce of gradients reproduce the
un-sharded oracle gradients."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from pointnerf_slam_b200 import dist as D
    from oracle import nice_oracle as O
    from tests import helpers as T
    r, w, _ = D.init_from_env("gloo")
    assert (r, w) == (rank, world)
    torch.set_num_threads(1)
    g = T.load_nice()
    n = g["rays_o"].shape[0]
    b, e = D.shard_bounds(n, rank, world)
    grids = {k: g[k].clone().requires_grad_(True) for k in O.GRID_KEYS}
    sd = {k: v.clone().requires_grad_(k.startswith("color_decoder.")) for k, v in T.state_dict(g).items()}
    gd = g["gt_depth"][b:e]
    dmax = D.share_depth_max(gd)
    assert float(dmax) == float(g["gt_depth"].max())
    # the oracle takes the batch maximum from the tensor it is given: append a far, zero-weight
    # sentinel so that every shard sees the global maximum (what depth_max_override does on GPU)
    scene = T.oracle_scene(g, grids, sd)
    ro, rd, gc = g["rays_o"][b:e], g["rays_d"][b:e], g["gt_color"][b:e]
    ro2 = torch.cat([ro, ro[:1]]); rd2 = torch.cat([rd, rd[:1]]); gd2 = torch.cat([gd, dmax])
    d, v, c = O.render_batch_ray(scene, rd2, ro2, "color", gd2)
    loss = O.mapping_loss(d[:-1], c[:-1], gd, gc, "color")
    loss.backward()
    trained = [grids[k] for k in ("grid_middle", "grid_fine", "grid_color")] + [p for p in sd.values() if p.requires_grad]
    D.allreduce_gradients([t.grad for t in trained])
    if rank == 0:
        torch.save({"grads": [t.grad for t in trained]}, out)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharded_mapping_matches_full(tmp_path):
    from oracle import nice_oracle as O
    from tests import helpers as T
    out = str(tmp_path / "g.pt")
    port = 29000 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    got = torch.load(out)["grads"]
    g = T.load_nice()
    keys = ["color/map/grad_grid_middle", "color/map/grad_grid_fine", "color/map/grad_grid_color"]
    for a, k in zip(got[:3], keys):
        assert T.rel_max(a, g[k]) < 1e-4, k
    names = [k for k in T.state_dict(g) if k.startswith("color_decoder.")]
    for a, nme in zip(got[3:], names):
        assert T.rel_max(a, g["color/map/gradsd/" + nme]) < 1e-4, nme


def test_shard_bounds_cover_everything():
    from pointnerf_slam_b200 import dist as D
    for n in (0, 1, 7, 96, 5000):
        for world in (1, 2, 3, 8):
            spans = [D.shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1


def _reducer_worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from pointnerf_slam_b200 import dist as D
    from pointnerf_slam_b200 import engine as E
    import pointnerf_slam_b200 as P
    D.init_from_env("gloo")
    torch.manual_seed(rank)
    # a grid leaf whose .grad autograd COPIED, and a decoder whose .grad are the hook's own views
    leaf = torch.zeros(1, 32, 3, 4, 5).contiguous(memory_format=torch.channels_last_3d)
    g = torch.randn(1, 3, 4, 5, 32).permute(0, 4, 1, 2, 3)       # channels-last gradient buffer
    leaf.grad = g.clone()
    dec = P.NICE(coarse=False).color_decoder
    params = E.grid_mlp_tensors(dec)
    gp = E.zeros_like_flat(params)
    for t in gp:
        t.copy_(torch.randn(t.shape))
    for p_, t in zip(params, gp):
        p_.grad = t
    pose = torch.randn(7)
    expect_grid, expect_p0, expect_pose = g.clone(), gp[1].clone(), pose.clone()
    red = D.OverlappedGradReducer()
    with red:
        assert E.GRAD_READY_HOOK is not None
        E.GRAD_READY_HOOK("grid_color", g)
        E.GRAD_READY_HOOK(("params", "color"), gp)
    assert E.GRAD_READY_HOOK is None
    red.finish({"grid_color": leaf}, decoders={"color": dec}, others=[pose])
    for t in (expect_grid, expect_p0, expect_pose):
        dist.all_reduce(t)
    ok = torch.allclose(leaf.grad, expect_grid) and torch.allclose(params[1].grad, expect_p0) and torch.allclose(pose, expect_pose)
    # arena mode: sinks carved from one buffer, ONE all-reduce in finish()
    arena = E.GradArena(1 << 16, "cpu")
    E.GRAD_ARENA = arena
    arena.reset()
    g2 = E.new_grid_grad(leaf)
    g2.copy_(torch.randn(g2.shape))
    gp2 = E.zeros_like_flat(params)
    for t in gp2:
        t.copy_(torch.randn(t.shape))
    assert arena.owns(g2) and arena.owns(gp2[3])
    leaf.grad = g2                       # autograd kept the buffer itself
    for p_, t in zip(params, gp2):
        p_.grad = t.clone()              # autograd copied
    e_grid, e_p3 = g2.clone(), gp2[3].clone()
    red2 = D.OverlappedGradReducer(arena)
    with red2:
        E.GRAD_READY_HOOK("grid_color", g2)
        E.GRAD_READY_HOOK(("params", "color"), gp2)
    red2.finish({"grid_color": leaf}, decoders={"color": dec})
    dist.all_reduce(e_grid); dist.all_reduce(e_p3)
    ok = ok and torch.allclose(leaf.grad, e_grid) and torch.allclose(params[3].grad, e_p3)
    # bench.py's multi-GPU default: arena installed, per-gradient overlapped reductions (reducer WITHOUT the arena).
    # Every sink is a view of a view of arena.buf; each must be summed exactly once.
    arena.reset()
    g3 = E.new_grid_grad(leaf); g3.copy_(torch.randn(g3.shape))
    other = E.zeros((17, 3), "cpu"); other.copy_(torch.randn(17, 3))       # a sink no hook ever announces (g_pts)
    gp3 = E.zeros_like_flat(params)
    for t in gp3:
        t.copy_(torch.randn(t.shape))
    assert gp3[0]._base is arena.buf and gp3.flat.numel() < arena.buf.numel()
    leaf.grad = g3
    for p_, t in zip(params, gp3):
        p_.grad = t
    e_grid3, e_all = g3.clone(), [t.clone() for t in gp3]
    keep_other = other.clone()
    red3 = D.OverlappedGradReducer(None)
    with red3:
        E.GRAD_READY_HOOK("grid_color", g3)
        E.GRAD_READY_HOOK(("params", "color"), gp3)
    red3.finish({"grid_color": leaf}, decoders={"color": dec})
    E.GRAD_ARENA = None
    dist.all_reduce(e_grid3)
    for t in e_all:
        dist.all_reduce(t)
    ok = ok and torch.allclose(leaf.grad, e_grid3) and all(torch.allclose(p_.grad, t) for p_, t in zip(params, e_all))
    ok = ok and torch.equal(other, keep_other)                               # untouched by any collective
    if rank == 0:
        torch.save({"ok": bool(ok)}, out)
    dist.barrier()
    dist.destroy_process_group()


def test_overlapped_reducer_two_ranks(tmp_path):
    out = str(tmp_path / "r.pt")
    port = 31000 + os.getpid() % 2000
    mp.spawn(_reducer_worker, args=(2, port, out), nprocs=2, join=True)
    assert torch.load(out)["ok"]
