"""Host-side logic that needs no GPU: pass planning, the gradient arena, sharding, and the loud failures of the
CUDA-only entry points on a CPU box."""
import types

import pytest
import torch

import pointnerf_slam_b200 as P
from pointnerf_slam_b200 import _lib as L
from pointnerf_slam_b200 import dist as D
from pointnerf_slam_b200 import engine as E


def _decoders():
    cfg = {"model": {"c_dim": 32, "coarse_bound_enlarge": 2, "pos_embedding_method": "fourier"}, "grid_len":
           {"coarse": 2, "middle": 0.32, "fine": 0.16, "color": 0.16, "bound_divisible": 0.32}, "coarse": True,
           "data": {"dim": 3}}
    return P.get_model(cfg, nice=True)


def test_stage_passes_follow_nice_forward():
    """decoder.py:312-342: which sub-decoders run per stage, on which grids, and how they share raw (N,4)."""
    dec = _decoders()
    bound = torch.tensor([[-1.0, 1.0]] * 3)
    modes = {st: [(p.kind, p.grid_a, p.grid_b, p.out_mode) for p in E.stage_passes(dec, st, bound)]
             for st in ("coarse", "middle", "fine", "color")}
    assert modes["coarse"] == [("coarse", "grid_coarse", None, L.OUT_SET_ALL)]
    assert modes["middle"] == [("grid", "grid_middle", None, L.OUT_SET_ALL)]
    assert modes["fine"] == [("grid", "grid_fine", "grid_middle", L.OUT_SET_ALL), ("grid", "grid_middle", None, L.OUT_ADD_W)]
    # colour: rgb from the colour decoder only, w = fine then += middle; the colour pass touches no w
    assert modes["color"] == [("grid", "grid_color", None, L.OUT_SET_RGB), ("grid", "grid_fine", "grid_middle", L.OUT_SET_W),
                              ("grid", "grid_middle", None, L.OUT_ADD_W)]
    with pytest.raises(ValueError):
        E.stage_passes(dec, "nope", bound)


def test_grad_arena_carves_aligned_zeroed_views():
    arena = E.GradArena(1000, "cpu")
    a = arena.take(10)
    b = arena.take(33)
    assert a.numel() == 10 and b.numel() == 33
    assert (b.data_ptr() - a.data_ptr()) == 32 * 4, "every sink starts on a 128-byte boundary"
    assert arena.owns(a) and arena.owns(b) and not arena.owns(torch.zeros(4))
    a.fill_(3.0); b.fill_(4.0)
    assert float(arena.used().sum()) == 30.0 + 132.0
    arena.reset()
    assert arena.offset == 0 and float(arena.buf.abs().sum()) == 0.0
    assert arena.take(2000) is None, "an exhausted arena makes the caller fall back to a fresh buffer"
    E.GRAD_ARENA = arena
    try:
        z = E.zeros((3, 4), "cpu")
        assert arena.owns(z) and z.shape == (3, 4)
        flat = E.zeros_like_flat([torch.empty(5), torch.empty(2, 3)])
        assert [tuple(t.shape) for t in flat] == [(5,), (2, 3)] and all(arena.owns(t) for t in flat)
        assert (flat[1].data_ptr() - flat[0].data_ptr()) % 16 == 0
    finally:
        E.GRAD_ARENA = None
    assert not arena.owns(E.zeros((3, 4), "cpu"))


def test_shard_bounds_partition_the_rays():
    for n in (0, 1, 7, 5000):
        for world in (1, 2, 3, 8):
            spans = [D.shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_cuda_only_entry_points_fail_loudly_on_cpu():
    """No CPU fallback anywhere on the product path."""
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(RuntimeError):
        P.graphs.GraphedStep(lambda: torch.zeros(()))
    with pytest.raises(RuntimeError):
        P.losses.mapping_loss(torch.zeros(4, dtype=torch.float64), torch.zeros(4, 3), torch.ones(4), torch.zeros(4, 3))
    with pytest.raises(RuntimeError):
        P.losses.tracking_loss(torch.zeros(4, dtype=torch.float64), torch.ones(4, dtype=torch.float64), torch.zeros(4, 3),
                               torch.ones(4), torch.zeros(4, 3))
    dec = _decoders()
    slam = types.SimpleNamespace(bound=torch.tensor([[-1.0, 1.0]] * 3), H=8, W=8, fx=8.0, fy=8.0, cx=3.5, cy=3.5, nice=True)
    cfg = {"rendering": {"lindisp": False, "perturb": 0.0, "N_samples": 4, "N_surface": 2, "N_importance": 0}, "scale": 1,
           "occupancy": True, "coarse": True, "data": {"dim": 3}, "model": {"c_dim": 32, "coarse_bound_enlarge": 2,
                                                                             "pos_embedding_method": "fourier"},
           "grid_len": {"coarse": 2, "middle": 0.32, "fine": 0.16, "color": 0.16, "bound_divisible": 0.32}}
    r = P.Renderer(cfg, None, slam)
    with pytest.raises(Exception):
        r.eval_points(torch.zeros(4, 3, dtype=torch.float64), dec, {}, "middle", "cpu")


def test_header_prototypes_cover_every_entry_point():
    """_lib parses include/pnslam.h: every declared function gets ctypes argtypes, so an argument-count mismatch raises."""
    lib = L.lib()
    for name in ("pn_grid_mlp_fwd", "pn_grid_mlp_bwd", "pn_grid_mlp_wgrad", "pn_mapping_loss", "pn_reserve_sms",
                 "pn_ray_zvals", "pn_composite_fwd", "pn_tc_selftest_mn"):
        fn = getattr(lib, name)
        assert fn.argtypes is not None, name
    with pytest.raises(Exception):
        lib.pn_reserve_sms()          # wrong argument count


def test_reserve_sms_round_trip():
    """pn_reserve_sms returns the previous reservation and clamps negatives (no device needed)."""
    lib = L.lib()
    prev = lib.pn_reserve_sms(12)
    try:
        assert lib.pn_reserve_sms(-5) == 12      # negative -> 0
        assert lib.pn_reserve_sms(0) == 0
    finally:
        lib.pn_reserve_sms(prev)
