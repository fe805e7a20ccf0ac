"""oracle/mapper_oracle.py replays the goldens minted from the reference's own Mapper / Mesher functions
(oracle/pin_mapper_against_reference.py): frustum masks, cv2.remap, ray pre-filter, keyframe overlap, masked Adam,
point masks.  CPU only."""
import numpy as np
import torch

from oracle import mapper_oracle as M
from oracle import nice_oracle as O
from tests import helpers as T

H, W, FX, FY, CX, CY = 680, 1200, 600.0, 600.0, 599.5, 339.5
SHAPES = {"grid_middle": (21, 28, 37), "grid_fine": (43, 56, 74)}


def golden():
    return np.load(T.GOLDEN + "/mapper.npz")


def synthetic_depth(seed):
    g = torch.Generator().manual_seed(seed)
    depth = 1.0 + 2.0 * torch.rand(H, W, generator=g)
    depth[torch.rand(H, W, generator=g) < 0.02] = 0.0
    return depth


def room0_bound():
    return O.scene_bound([[-2.9, 8.9], [-3.2, 5.5], [-3.5, 3.3]], 1.0, 0.32)


def test_cv_remap_restatement_matches_opencv_golden():
    g = golden()
    depth0 = synthetic_depth(7).numpy()
    out = M.cv_remap_bilinear(depth0, g["remap/x"], g["remap/y"])
    assert np.array_equal(out.view(np.uint32), g["remap/out"].view(np.uint32))


def test_frustum_mask_matches_reference_golden():
    g = golden()
    bound = room0_bound()
    for kf in (0, 30):
        depth = synthetic_depth(int(g[f"frustum/{kf}/depth_seed"])).numpy()
        for key, vs in SHAPES.items():
            m = M.frustum_mask(g[f"frustum/{kf}/c2w"], key, vs, depth, bound, H, W, FX, FY, CX, CY)
            ref = np.unpackbits(g[f"frustum/{kf}/{key}"])[:m.size].astype(bool).reshape(m.shape)
            assert np.array_equal(m, ref), (kf, key)
            assert 0 < m.sum() < m.size
    assert M.frustum_mask(g["frustum/0/c2w"], "grid_coarse", (7, 8, 11), depth, bound, H, W, FX, FY, CX, CY).all()


def test_ray_prefilter_and_overlap_goldens():
    g = golden()
    bound = room0_bound()
    keep = M.ray_prefilter_mask(torch.from_numpy(g["prefilter/rays_o"]), torch.from_numpy(g["prefilter/rays_d"]),
                                torch.from_numpy(g["prefilter/gt_depth"]), bound)
    assert np.array_equal(keep.numpy(), g["prefilter/mask"]) and 0 < keep.sum() < keep.numel()
    verts = M.overlap_points(torch.from_numpy(g["overlap/rays_o"]), torch.from_numpy(g["overlap/rays_d"]),
                             torch.from_numpy(g["overlap/gt_depth"]), 16)
    assert np.array_equal(verts, g["overlap/vertices"])
    poses = np.load(T.GOLDEN + "/room0_poses.npz")["c2w"]
    fr = M.keyframe_overlap_fractions(verts, [poses[i] for i in g["overlap/kf_ids"]], H, W, FX, FY, CX, CY)
    assert np.array_equal(fr, g["overlap/fractions"])
    sel = M.select_overlapping_keyframes(fr, 4, np.random.RandomState(5))
    assert list(map(int, sel)) == list(map(int, g["overlap/selected"]))


def test_masked_adam_matches_torch_optim_golden():
    g = golden()
    p = torch.from_numpy(g["adam/val0"]).clone()
    mask = torch.from_numpy(g["adam/mask"])[None, None].repeat(1, 32, 1, 1, 1)
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    for it in range(6):
        lr = M.STAGE_LR["middle" if it < 3 else "color"]["middle_lr"] * 5
        M.adam_step(p, torch.from_numpy(g["adam/grads"][it]), m, v, it + 1, lr, mask=mask)
    torch.testing.assert_close(p, torch.from_numpy(g["adam/val6"]), rtol=1e-6, atol=1e-9)
    assert torch.equal(p[~mask], torch.from_numpy(g["adam/val0"])[~mask])        # unmasked voxels never move


def test_stage_schedule():
    # Mapper.py:520-527 with n = 60, ratios 0.4 / 0.6: 25 middle, 12 fine, 23 colour iterations
    st = [M.stage_of_iter(i, 60) for i in range(60)]
    assert (st.count("middle"), st.count("fine"), st.count("color")) == (25, 12, 23)
    assert M.stage_of_iter(0, 60, coarse_mapper=True) == "coarse"


def test_point_masks_golden():
    g = golden()
    poses = np.load(T.GOLDEN + "/room0_poses.npz")["c2w"]
    pts = torch.from_numpy(g["pmask/points"])
    kfs = [int(i) for i in g["pmask/kf_ids"]]
    c2w = [torch.from_numpy(poses[i]) for i in kfs]
    depth = [synthetic_depth(50 + i) for i in kfs]
    for dt in (0, 1):
        seen, fore, unseen = M.point_masks(pts, c2w, depth, H, W, FX, FY, CX, CY, bool(dt))
        assert np.array_equal(np.packbits(seen), g[f"pmask/{dt}/seen"]) and np.array_equal(np.packbits(fore), g[f"pmask/{dt}/forecast"])
        assert not (seen & fore).any() and (seen | fore | unseen).all()
