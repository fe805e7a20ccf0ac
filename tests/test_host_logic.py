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


def test_product_scene_setup_matches_oracle_bit_for_bit():
    """config.load_bound / grid_shapes / grid_init (src/NICE_SLAM.py:200-315): the PRODUCT's functions, not only the
    oracle's, against the reference values (float32 rounding of the enlarged bound, truncated z extent, x<->z swap)."""
    from oracle import nice_oracle as O
    import bench as B
    bound = P.load_bound(B.CFG)
    ref = O.scene_bound(B.CFG["mapping"]["bound"], 1.0, 0.32)
    assert bound.dtype == torch.float64 and torch.equal(bound, ref)
    assert bound[:, 1].tolist() == [8.94000015258789, 5.7600000381469725, 3.5399999618530273]
    shapes = P.config.grid_shapes(B.CFG, bound)
    assert shapes == {"grid_coarse": (7, 8, 11), "grid_middle": (21, 28, 37), "grid_fine": (43, 56, 74), "grid_color": (43, 56, 74)}
    assert shapes == O.grid_shapes(ref, B.CFG["grid_len"])
    grids = P.grid_init(B.CFG, bound, "cpu", generator=torch.Generator().manual_seed(1))
    og = O.init_grids(ref, B.CFG["grid_len"], 32, 2, True, torch.Generator().manual_seed(1))
    for k, g in grids.items():
        assert g.shape == (1, 32) + shapes[k] and g.is_contiguous(memory_format=torch.channels_last_3d)
        assert torch.equal(g, og[k]), k                     # same values in the reference's logical (1,C,Z,Y,X) order
    assert float(grids["grid_fine"].std()) < 2e-4 < float(grids["grid_middle"].std())
    # a bound that is already a multiple of 0.32 still grows by one cell (NICE_SLAM.py:211: int(...)+1)
    cfg2 = dict(B.CFG, mapping={"bound": [[0.0, 0.64], [0.0, 0.32], [-0.32, 0.32]]})
    assert torch.equal(P.load_bound(cfg2), O.scene_bound(cfg2["mapping"]["bound"], 1.0, 0.32))


def test_get_tensor_from_camera_round_trip():
    """common.get_tensor_from_camera (src/common.py:179-201; host-side Shepperd conversion instead of mathutils): every
    branch of the conversion, [w,x,y,z] order with w >= 0, Tquad order, and the inverse through the oracle's
    camera_from_tensor."""
    import numpy as np
    from scipy.spatial.transform import Rotation
    from oracle import nice_oracle as O
    rots = [Rotation.from_euler("xyz", a).as_matrix() for a in ([0.1, -0.2, 0.3], [3.0, 0.1, 0.0], [0.1, 3.0, 0.2], [0.2, 0.1, 3.1],
                                                                [2.0, 2.0, 2.0], [0, 0, 0])]
    for i, R in enumerate(rots):
        RT = np.eye(4)
        RT[:3, :3], RT[:3, 3] = R, [0.5 * i, -1.0, 2.0]
        t = P.get_tensor_from_camera(torch.from_numpy(RT).float())
        assert t.shape == (7,) and t.dtype == torch.float32 and float(t[0]) >= 0.0
        q = Rotation.from_matrix(R).as_quat()               # scipy: [x,y,z,w]
        qs = np.array([q[3], q[0], q[1], q[2]])
        qs = -qs if qs[0] < 0 else qs
        assert np.allclose(t[:4].numpy(), qs, atol=1e-6) or np.allclose(t[:4].numpy(), -qs, atol=1e-6)
        assert np.allclose(t[4:].numpy(), RT[:3, 3], atol=1e-7)
        back = O.camera_from_tensor(t.double())
        assert np.allclose(back.numpy(), RT[:3, :4], atol=1e-6)
        tq = P.get_tensor_from_camera(RT[:3], Tquad=True)
        assert torch.allclose(tq[:3], t[4:]) and torch.allclose(tq[3:], t[:4])


def test_mapper_side_entry_points_fail_loudly_on_cpu():
    import pointnerf_slam_b200.mapper as PM
    with pytest.raises(RuntimeError, match="CUDA"):
        PM.frustum_voxel_mask(torch.eye(4), "grid_fine", (4, 4, 4), torch.zeros(8, 8), torch.tensor([[0.0, 1.0]] * 3), 8, 8, 1, 1, 1, 1)
    with pytest.raises(RuntimeError, match="CUDA"):
        PM.StageOptimizer({"grid_middle": torch.zeros(1, 32, 2, 2, 2)}, [], [])
    with pytest.raises(RuntimeError, match="CUDA"):
        PM.prefilter_mask(torch.zeros(4, 3), torch.ones(4, 3), torch.ones(4), torch.tensor([[0.0, 1.0]] * 3))
    with pytest.raises(RuntimeError, match="CUDA"):
        PM.select_depth_pixels(torch.zeros(8, 8), 0, 8, 0, 8)


def test_round2_entry_points_fail_loudly_on_cpu():
    """k-NN aggregation, the gradient-returning loss heads and the iteration callables: CUDA tensors or an exception."""
    import pointnerf_slam_b200.knn as PK
    from pointnerf_slam_b200.tracking import TrackingIteration
    xyz, feat = torch.rand(10, 3), torch.zeros(10, 32)
    bound = [[0.0, 1.0]] * 3
    with pytest.raises(RuntimeError, match="CUDA tensor required"):
        PK.NeuralPointIndex(xyz, bound, 0.2)
    with pytest.raises(RuntimeError, match="CUDA tensor required"):
        PK.NeuralPointField(xyz, feat, bound, 0.2)
    with pytest.raises(RuntimeError, match="CUDA"):
        P.losses.mapping_loss_and_grads(torch.zeros(4, dtype=torch.float64), torch.zeros(4, 3), torch.ones(4), torch.zeros(4, 3))
    with pytest.raises(RuntimeError, match="CUDA"):
        P.losses.tracking_loss_and_grads(torch.zeros(4, dtype=torch.float64), torch.ones(4, dtype=torch.float64), torch.zeros(4, 3),
                                         torch.ones(4), torch.zeros(4, 3))
    if not torch.cuda.is_available():
        cam = torch.tensor([1.0, 0, 0, 0, 0, 0, 0], requires_grad=True)
        it = TrackingIteration(None, None, {}, torch.ones(8, 8), torch.zeros(8, 8, 3), cam, 8, 8, 8.0, 8.0, 3.5, 3.5, 4)
        with pytest.raises(Exception):
            it()


def test_knn_struct_matches_the_header():
    """ctypes mirror of pn_knn_index: field order and sizes as declared in include/pnslam.h."""
    import ctypes as C
    import pointnerf_slam_b200.knn as PK
    names = [f[0] for f in PK.PnKnnIndex._fields_]
    assert names == ["start", "sorted", "lo", "inv_h", "nx", "ny", "nz", "P"]
    assert C.sizeof(PK.PnKnnIndex) == 8 + 8 + 12 + 4 + 16
    hdr = open(L.HEADER_PATH).read()
    body = hdr[hdr.index("typedef struct pn_knn_index {"):hdr.index("} pn_knn_index;")]
    order = [body.index(k) for k in ("start;", "sorted;", "lo[3];", "inv_h;", "nx, ny, nz;", "int P;")]
    assert order == sorted(order)
    for fn in ("pn_knn_build", "pn_knn_query", "pn_knn_aggregate_fwd", "pn_knn_aggregate_bwd"):
        assert hasattr(L.lib(), fn) and getattr(L.lib(), fn).argtypes is not None
