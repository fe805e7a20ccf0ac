"""CUDA kernels of the Mapper / Tracker / Mesher work around the rendering path (pointnerf_slam_b200.mapper) against
the oracle and the reference goldens: bit-exact masks, indices and counts; Adam within 1e-6."""
import numpy as np
import pytest
import torch

from oracle import mapper_oracle as M
from tests import helpers as T
from tests.test_mapper_oracle import CX, CY, FX, FY, H, SHAPES, W, golden, room0_bound, synthetic_depth

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_frustum_mask_bit_exact():
    import pointnerf_slam_b200.mapper as PM
    g = golden()
    bound = room0_bound()
    for kf in (0, 30):
        depth = synthetic_depth(int(g[f"frustum/{kf}/depth_seed"]))
        c2w = torch.from_numpy(g[f"frustum/{kf}/c2w"])
        for key, vs in SHAPES.items():
            m = PM.get_mask_from_c2w(c2w.to(DEV), key, vs, depth.to(DEV), bound, H, W, FX, FY, CX, CY)
            assert m.shape == (vs[2], vs[1], vs[0]) and m.dtype == torch.bool
            ref = np.unpackbits(g[f"frustum/{kf}/{key}"])[:m.numel()].astype(bool).reshape(tuple(m.shape))
            assert np.array_equal(m.cpu().numpy(), ref), (kf, key, int((m.cpu().numpy() != ref).sum()))
    assert PM.frustum_voxel_mask(c2w, "grid_coarse", (7, 8, 11), depth.to(DEV), bound, H, W, FX, FY, CX, CY).all()


def test_frustum_mask_degenerate_inputs():
    """Camera far outside the scene / all-zero depth image: same mask as the oracle (huge and negative projections)."""
    import pointnerf_slam_b200.mapper as PM
    bound = room0_bound()
    c2w = torch.eye(4)
    c2w[:3, 3] = torch.tensor([40.0, -30.0, 25.0])
    for depth in (torch.zeros(H, W), synthetic_depth(3)):
        ref = M.frustum_mask(c2w.numpy(), "grid_middle", SHAPES["grid_middle"], depth.numpy(), bound, H, W, FX, FY, CX, CY)
        m = PM.get_mask_from_c2w(c2w, "grid_middle", SHAPES["grid_middle"], depth.to(DEV), bound, H, W, FX, FY, CX, CY)
        assert np.array_equal(m.cpu().numpy(), ref)


def _adam_case(channels_last):
    g = golden()
    val0 = torch.from_numpy(g["adam/val0"])
    p = val0.to(DEV)
    p = p.contiguous(memory_format=torch.channels_last_3d) if channels_last else p.contiguous()
    return g, val0, p


@pytest.mark.parametrize("channels_last", [True, False])
def test_masked_adam_matches_torch_golden(channels_last):
    import pointnerf_slam_b200.mapper as PM
    g, val0, p = _adam_case(channels_last)
    mask = torch.from_numpy(g["adam/mask"]).to(DEV)
    grids = {"grid_middle": p}
    table = {s: dict(M.STAGE_LR[s]) for s in M.STAGE_LR}
    opt = PM.StageOptimizer(grids, [], [], masks={"grid_middle": mask}, stage_lr=table, lr_factor=5.0)
    for it in range(6):
        opt.set_stage("middle" if it < 3 else "color")
        gr = torch.from_numpy(g["adam/grads"][it]).to(DEV)
        p.grad = gr.contiguous(memory_format=torch.channels_last_3d) if channels_last else gr.contiguous()
        opt.step()
    ref = torch.from_numpy(g["adam/val6"])
    assert T.rel_max(p, ref) < 1e-6
    m5 = mask[None, None].expand_as(p).cpu()
    assert torch.equal(p.cpu()[~m5], val0[~m5])
    assert opt.step_count.tolist() == [6]


def test_adam_groups_follow_torch_optim():
    """Decoder + camera + two grids through the stage schedule against torch.optim.Adam with the reference's six groups
    (Mapper.py:482-505, 529-536); a parameter without gradient in a stage keeps its step count (colour grid in stage fine)."""
    import pointnerf_slam_b200.mapper as PM
    gen = torch.Generator().manual_seed(0)
    shapes = {"grid_middle": (1, 32, 3, 4, 5), "grid_fine": (1, 32, 5, 6, 7), "grid_color": (1, 32, 5, 6, 7)}
    ours = {k: (torch.randn(s, generator=gen) * 0.01).to(DEV).contiguous(memory_format=torch.channels_last_3d) for k, s in shapes.items()}
    ref = {k: v.detach().clone().cpu().requires_grad_(True) for k, v in ours.items()}
    dec_o = [torch.randn(32, 93, generator=gen).to(DEV), torch.randn(7, generator=gen).to(DEV)]
    dec_r = [t.detach().clone().cpu().requires_grad_(True) for t in dec_o]
    cam_o = [torch.randn(7, generator=gen).to(DEV)]
    cam_r = [t.detach().clone().cpu().requires_grad_(True) for t in cam_o]
    tab = M.STAGE_LR
    opt_r = torch.optim.Adam([{"params": dec_r, "lr": 0}, {"params": [], "lr": 0}, {"params": [ref["grid_middle"]], "lr": 0},
                              {"params": [ref["grid_fine"]], "lr": 0}, {"params": [ref["grid_color"]], "lr": 0}, {"params": cam_r, "lr": 0}])
    opt_o = PM.StageOptimizer(ours, dec_o, cam_o, stage_lr=tab, lr_factor=1.0, BA_cam_lr=0.001)
    n = 10
    for it in range(n):
        stage = M.stage_of_iter(it, n)
        for gi, k in enumerate(PM.LR_KEYS):
            opt_r.param_groups[gi]["lr"] = tab[stage][k]
        opt_r.param_groups[5]["lr"] = 0.001 if stage == "color" else 0.0
        opt_o.set_stage(stage)
        used = {"middle": ["grid_middle"], "fine": ["grid_middle", "grid_fine"], "color": list(shapes)}[stage]
        for k in shapes:
            gr = torch.randn(shapes[k], generator=gen) if k in used else None
            ref[k].grad = gr
            ours[k].grad = None if gr is None else gr.to(DEV).contiguous(memory_format=torch.channels_last_3d)
        for a, b in zip(dec_o + cam_o, dec_r + cam_r):
            gr = torch.randn(b.shape, generator=gen) if stage == "color" else None
            b.grad = gr
            a.grad = None if gr is None else gr.to(DEV)
        opt_r.step()
        opt_o.step()
    for k in shapes:
        assert T.rel_max(ours[k], ref[k]) < 1e-6, k
    for a, b in zip(dec_o + cam_o, dec_r + cam_r):
        assert T.rel_max(a, b) < 1e-6
    steps = opt_o.step_count.tolist()
    assert steps[0] == n and steps[2] < steps[1] < steps[0]      # middle every iteration, fine later, colour last


def test_prefilter_and_depth_pixel_selection_bit_exact():
    import pointnerf_slam_b200.mapper as PM
    g = golden()
    bound = room0_bound()
    ro, rd, gd = [torch.from_numpy(g[f"prefilter/{k}"]).to(DEV) for k in ("rays_o", "rays_d", "gt_depth")]
    keep = PM.prefilter_mask(ro, rd, gd, bound)
    assert np.array_equal(keep.cpu().numpy().astype(bool), g["prefilter/mask"])
    gc = torch.rand(ro.shape[0], 3, device=DEV)
    o2, d2, gd2, gc2 = PM.prefilter_rays(ro, rd, gd, gc, bound)
    m = torch.from_numpy(g["prefilter/mask"]).to(DEV)
    assert torch.equal(o2, ro[m]) and torch.equal(d2, rd[m]) and torch.equal(gd2, gd[m]) and torch.equal(gc2, gc[m])
    # empty / full selections
    assert PM.prefilter_rays(ro[:0], rd[:0], gd[:0], gc[:0], bound)[0].shape[0] == 0
    # the fork tracker's pixel selection (np.where(depth > 0.01)) on the shipped-pseudo-depth-like sparse map, full frame and crop
    depth = torch.zeros(H, W)
    idx = torch.randperm(H * W, generator=torch.Generator().manual_seed(9))[:7800]
    depth.view(-1)[idx] = 1.0 + torch.rand(7800, generator=torch.Generator().manual_seed(10))
    depth += 1e-5 * torch.rand(H, W, generator=torch.Generator().manual_seed(11))
    for (h0, h1, w0, w1) in ((0, H, 0, W), (100, H - 100, 100, W - 100)):
        ref = M.select_depth_pixels(depth[h0:h1, w0:w1])
        out = PM.select_depth_pixels(depth.to(DEV), h0, h1, w0, w1)
        assert out.dtype == torch.int64 and torch.equal(out.cpu(), ref)
    assert PM.select_depth_pixels(torch.zeros(H, W, device=DEV), 0, H, 0, W).numel() == 0
    assert PM.select_depth_pixels(torch.ones(H, W, device=DEV), 0, H, 0, W).numel() == H * W


def test_keyframe_overlap_bit_exact():
    import pointnerf_slam_b200.mapper as PM
    g = golden()
    poses = np.load(T.GOLDEN + "/room0_poses.npz")["c2w"]
    ro, rd, gd = [torch.from_numpy(g[f"overlap/{k}"]).to(DEV) for k in ("rays_o", "rays_d", "gt_depth")]
    fr = PM.keyframe_overlap_fractions(ro, rd, gd, [poses[i] for i in g["overlap/kf_ids"]], H, W, FX, FY, CX, CY, 16)
    assert np.array_equal(fr, g["overlap/fractions"])
    sel = M.select_overlapping_keyframes(fr, 4, np.random.RandomState(5))
    assert list(map(int, sel)) == list(map(int, g["overlap/selected"]))


def test_point_masks_bit_exact():
    import pointnerf_slam_b200.mapper as PM
    g = golden()
    poses = np.load(T.GOLDEN + "/room0_poses.npz")["c2w"]
    pts = torch.from_numpy(g["pmask/points"])
    kfs = [int(i) for i in g["pmask/kf_ids"]]
    kd = [{"est_c2w": torch.from_numpy(poses[i]), "depth": synthetic_depth(50 + i)} for i in kfs]
    for dt in (0, 1):
        for bs in (500000, 7000):          # one chunk / three chunks (max(depth_sample) is per chunk)
            seen, fore, unseen = PM.point_masks(pts, kd, H, W, FX, FY, CX, CY, DEV, depth_test=bool(dt), points_batch_size=bs)
            rs, rf, ru = M.point_masks(pts, [k["est_c2w"] for k in kd], [k["depth"] for k in kd], H, W, FX, FY, CX, CY, bool(dt), bs)
            assert np.array_equal(seen, rs) and np.array_equal(fore, rf) and np.array_equal(unseen, ru), (dt, bs)
        if True:
            s1, f1, _ = PM.point_masks(pts, kd, H, W, FX, FY, CX, CY, DEV, depth_test=bool(dt))
            assert np.array_equal(np.packbits(s1), g[f"pmask/{dt}/seen"]) and np.array_equal(np.packbits(f1), g[f"pmask/{dt}/forecast"])


def test_vis_residuals():
    import pointnerf_slam_b200.mapper as PM
    gen = torch.Generator().manual_seed(2)
    gd = synthetic_depth(4)[:64, :96].contiguous()
    gc = torch.rand(64, 96, 3, generator=gen) * 1.2 - 0.1
    d = (gd + 0.1 * torch.randn(64, 96, generator=gen)).double()
    c = torch.rand(64, 96, 3, generator=gen) * 1.4 - 0.2
    ref = M.vis_residuals(gd.numpy(), gc.numpy(), d.numpy(), c.numpy())
    out = PM.vis_residuals(gd.to(DEV), gc.to(DEV), d.to(DEV), c.to(DEV))
    for a, b in zip(out[:4], ref[:4]):
        assert np.array_equal(a.cpu().numpy(), b)
    assert float(out[4]) == float(ref[4])


def test_sparse_rows_roundtrip_and_ordered_sum():
    """pack -> apply over three 'ranks' reproduces the dense sum in rank order, bit for bit; untouched rows stay untouched."""
    import ctypes as C
    from pointnerf_slam_b200 import _lib as L
    lib = L.lib()
    V = 43 * 56 * 74 + 5          # not a multiple of 32 or 1024
    gen = torch.Generator().manual_seed(1)
    dense = []
    for r in range(3):
        d = torch.zeros(V, 32)
        rows = torch.randperm(V, generator=gen)[:9000 + 1000 * r]
        d[rows] = torch.randn(rows.numel(), 32, generator=gen)
        d[rows[:50], 1:] = 0.0         # rows with a single non-zero
        dense.append(d.to(DEV))
    expect = (dense[0] + dense[1]) + dense[2]
    nw, cap = (V + 31) // 32, 12000
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    packs = []
    for d in dense:
        bm = torch.zeros(nw, dtype=torch.int32, device=DEV); pf = torch.zeros(nw, dtype=torch.int32, device=DEV)
        rows = torch.zeros(cap, 32, device=DEV); scratch = torch.zeros((V + 1023) // 1024 + 1, dtype=torch.int32, device=DEV)
        count = torch.zeros(1, dtype=torch.int64, device=DEV); ovf = torch.zeros(1, dtype=torch.int32, device=DEV)
        L.check(lib.pn_sparse_rows_pack(C.c_void_p(d.data_ptr()), C.c_int64(V), C.c_void_p(bm.data_ptr()), C.c_void_p(pf.data_ptr()),
                                        C.c_void_p(rows.data_ptr()), C.c_int64(cap), C.c_void_p(scratch.data_ptr()),
                                        C.c_void_p(count.data_ptr()), C.c_void_p(ovf.data_ptr()), st), "pack")
        nz = (d != 0).any(1)
        assert int(count) == int(nz.sum()) and int(ovf) == 0
        assert torch.equal(rows[:int(count)], d[nz])                      # ascending row order
        packs.append((bm, pf, rows))
    out = dense[1].clone()                                                  # "rank 1": its own dense gradient
    arr = lambda ts: (C.c_void_p * 3)(*[t.data_ptr() for t in ts])
    L.check(lib.pn_sparse_rows_apply(C.c_void_p(out.data_ptr()), C.c_int64(V), 3, arr([p[0] for p in packs]), arr([p[1] for p in packs]),
                                     arr([p[2] for p in packs]), C.c_int64(cap), st), "apply")
    assert torch.equal(out, expect)
    # overflow is reported, not silent
    ovf = torch.zeros(1, dtype=torch.int32, device=DEV)
    L.check(lib.pn_sparse_rows_pack(C.c_void_p(dense[2].data_ptr()), C.c_int64(V), C.c_void_p(packs[2][0].data_ptr()),
                                    C.c_void_p(packs[2][1].data_ptr()), C.c_void_p(packs[2][2].data_ptr()), C.c_int64(100),
                                    C.c_void_p(scratch.data_ptr()), C.c_void_p(count.data_ptr()), C.c_void_p(ovf.data_ptr()), st), "pack")
    assert int(ovf) == int(count) - 100
    # dense tail
    srcs = [torch.randn(1001, generator=gen).to(DEV) for _ in range(3)]
    o = torch.empty(1001, device=DEV)
    L.check(lib.pn_dense_sum(C.c_void_p(o.data_ptr()), C.c_int64(1001), 3, arr(srcs), st), "sum")
    assert torch.equal(o, (srcs[0] + srcs[1]) + srcs[2])


def test_batched_keyframe_sampling_equals_the_per_keyframe_loop():
    """get_samples_multi (one launch for the Mapper's per-keyframe loop, Mapper.py:558-605) == get_samples per keyframe + cat,
    bit for bit, forward and pose gradient; float64 colour images as the reference's dataset path delivers them."""
    import pointnerf_slam_b200 as P
    gen = torch.Generator().manual_seed(4)
    F, n = 3, 257
    cams = [torch.randn(7, generator=gen).to(DEV).requires_grad_(True) for _ in range(F)]
    cams2 = [c.detach().clone().requires_grad_(True) for c in cams]
    for dtype in (torch.float32, torch.float64):
        frames = [(synthetic_depth(60 + k).to(DEV), torch.rand(H, W, 3, generator=gen).to(dtype).to(DEV)) for k in range(F)]
        idx = torch.randint((H - 40) * (W - 60), (F, n), generator=gen).to(DEV)
        c2w = P.get_camera_from_tensor(torch.stack(cams))
        out = P.get_samples_multi(20, H - 20, 30, W - 30, n, H, W, FX, FY, CX, CY, c2w, P.KeyframeBatch(frames, DEV), DEV, indices=idx)
        ref = [P.get_samples(20, H - 20, 30, W - 30, n, H, W, FX, FY, CX, CY, P.get_camera_from_tensor(cams2[k]), frames[k][0], frames[k][1],
                             DEV, indices=idx[k]) for k in range(F)]
        for j in range(4):
            assert torch.equal(out[j], torch.cat([r[j] for r in ref])), j
        assert out[3].dtype == dtype
        wgt = torch.randn(F * n, 3, generator=gen).to(DEV)
        for c in cams + cams2:
            c.grad = None
        ((out[0] * wgt).sum() + (out[1] * wgt.flip(0)).sum()).backward()
        (sum((r[0] * wgt[k * n:(k + 1) * n]).sum() for k, r in enumerate(ref)) + sum((r[1] * wgt.flip(0)[k * n:(k + 1) * n]).sum() for k, r in enumerate(ref))).backward()
        for a, b in zip(cams, cams2):
            assert T.rel_max(a.grad, b.grad) < 1e-5
