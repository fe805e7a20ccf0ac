#!/usr/bin/env python
"""Pin ``oracle/mapper_oracle.py`` against the reference's own Mapper / Mesher code and mint
``tests/golden/mapper.npz``.  Runs ONLY in the build container (needs ``/root/reference``).

``src.Mapper`` and ``src.utils.Mesher`` import packages that are absent here (colorama, matplotlib, open3d,
skimage, trimesh, termcolor, g2o ...).  None of them is used by the functions pinned below, so they are
stubbed in ``sys.modules`` and the reference's *unmodified* functions are called unbound with a
``SimpleNamespace`` standing in for ``self``:

  * ``Mapper.get_mask_from_c2w``           (src/Mapper.py:129-200; numpy + cv2.remap)
  * ``Mapper.keyframe_selection_overlap``  (src/Mapper.py:267-333)
  * ``Mesher.point_masks``                 (src/utils/Mesher.py:53-212; torch)
  * the ray pre-filter block               (src/Mapper.py:607-621; commented upstream code, restated verbatim here)
  * ``torch.optim.Adam`` on ``val[mask]``  (src/Mapper.py:427-431, 482-505, 657-674)
  * ``cv2.remap``                          (the installed OpenCV)

usage:  python oracle/pin_mapper_against_reference.py [--out tests/golden]
"""
import argparse
import importlib
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.dont_write_bytecode = True

from oracle import mapper_oracle as M  # noqa: E402

H, W, FX, FY, CX, CY = 680, 1200, 600.0, 600.0, 599.5, 339.5


class _Stub(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Stub(self.__name__ + "." + name)

    def __call__(self, *a, **k):
        return self


def import_reference():
    os.chdir(REF)
    sys.path.insert(0, REF)
    for name in ("colorama", "matplotlib", "matplotlib.pyplot", "open3d", "skimage", "skimage.measure", "trimesh",
                 "termcolor", "g2o", "ordered_set", "packaging", "mathutils"):
        if name not in sys.modules:
            try:
                importlib.import_module(name)
            except Exception:
                sys.modules[name] = _Stub(name)
    if not hasattr(np, "bool"):
        np.bool = bool          # Mapper.py:151 (numpy < 1.24 alias; only the coarse branch touches it)
    from src.Mapper import Mapper
    from src.utils.Mesher import Mesher
    return Mapper, Mesher


def room0_bound():
    from oracle import nice_oracle as O
    return O.scene_bound([[-2.9, 8.9], [-3.2, 5.5], [-3.5, 3.3]], 1.0, 0.32)


def synthetic_depth(seed):
    g = torch.Generator().manual_seed(seed)
    depth = 1.0 + 2.0 * torch.rand(H, W, generator=g)
    depth[torch.rand(H, W, generator=g) < 0.02] = 0.0
    return depth


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "tests", "golden"))
    args = ap.parse_args()
    Mapper, Mesher = import_reference()
    import cv2
    out = {}
    poses = np.load(os.path.join(ROOT, "tests", "golden", "room0_poses.npz"))["c2w"].astype(np.float32)
    bound = room0_bound()

    # ---- cv2.remap restatement ----
    rng = np.random.RandomState(0)
    depth0 = synthetic_depth(7).numpy()
    mx = rng.uniform(-40, W + 40, 200000).astype(np.float32)
    my = rng.uniform(-40, H + 40, 200000).astype(np.float32)
    mx[:8] = [0.0, W - 1.0, W - 0.5, -0.5, 1e9, -1e9, np.inf, np.nan]
    my[:8] = [0.0, H - 1.0, H - 0.5, -0.5, 10.0, 10.0, 10.0, 10.0]
    def cv_remap(src, x, y, chunk=30000):     # dst.rows must stay below SHRT_MAX (why Mapper.py:165 chunks by 3e4)
        return np.concatenate([cv2.remap(src, x[i:i + chunk], y[i:i + chunk], interpolation=cv2.INTER_LINEAR)[:, 0]
                               for i in range(0, x.shape[0], chunk)])

    ref = cv_remap(depth0, mx, my)
    mine = M.cv_remap_bilinear(depth0, mx, my)
    nbad = int((ref.view(np.uint32) != mine.view(np.uint32)).sum())
    print(f"cv2.remap vs restatement: {nbad} of {ref.size} samples differ in bits")
    assert nbad == 0
    out["remap/x"], out["remap/y"], out["remap/out"] = mx[:4096], my[:4096], ref[:4096]

    # ---- frustum feature selection ----
    self_ = types.SimpleNamespace(H=H, W=W, fx=FX, fy=FY, cx=CX, cy=CY, bound=bound)
    shapes = {"grid_middle": (21, 28, 37), "grid_fine": (43, 56, 74)}
    for kf, seed in ((0, 11), (30, 12)):
        depth_np = synthetic_depth(seed).numpy()
        c2w = torch.from_numpy(np.concatenate([poses[kf][:3], [[0, 0, 0, 1]]], 0).astype(np.float32)) if poses[kf].shape[0] == 3 \
            else torch.from_numpy(poses[kf])
        for key, vs in shapes.items():
            ref = Mapper.get_mask_from_c2w(self_, c2w, key, vs, depth_np)
            mine = M.frustum_mask(c2w.numpy(), key, vs, depth_np, bound, H, W, FX, FY, CX, CY)
            mine_cv = M.frustum_mask(c2w.numpy(), key, vs, depth_np, bound, H, W, FX, FY, CX, CY,
                                     remap=cv_remap)
            d = int((ref != mine).sum())
            print(f"get_mask_from_c2w kf {kf} {key}: {int(ref.sum())} of {ref.size} selected; restatement differs in {d} voxels "
                  f"(with cv2.remap inside: {int((ref != mine_cv).sum())})")
            assert d == 0, "frustum mask restatement differs from the reference"
            out[f"frustum/{kf}/{key}"] = np.packbits(ref.reshape(-1))
        out[f"frustum/{kf}/c2w"] = c2w.numpy()
        out[f"frustum/{kf}/depth_seed"] = np.array(seed)

    # ---- keyframe overlap selection ----
    from oracle import nice_oracle as O
    torch.manual_seed(3)
    depth_t = synthetic_depth(21)
    color_t = torch.rand(H, W, 3, generator=torch.Generator().manual_seed(22))
    cur = torch.from_numpy(poses[20])
    idx = torch.randint(H * W, (100,), generator=torch.Generator().manual_seed(23))
    ro, rd, gd, gc, _ = O.get_samples(0, H, 0, W, 100, FX, FY, CX, CY, cur[:3], depth_t, color_t, idx)
    verts = M.overlap_points(ro, rd, gd, 16)
    kf_ids = list(range(0, 40, 2))
    fr = M.keyframe_overlap_fractions(verts, [poses[i] for i in kf_ids], H, W, FX, FY, CX, CY)
    # reference: same loop body, its own get_samples replaced by the fixed rays above
    import src.Mapper as RM
    keep = RM.get_samples
    RM.get_samples = lambda *a, **k: (ro, rd, gd, gc)
    self2 = types.SimpleNamespace(H=H, W=W, fx=FX, fy=FY, cx=CX, cy=CY, device="cpu", weak_depth=False)
    kd = [{"est_c2w": torch.from_numpy(poses[i])} for i in kf_ids]
    np.random.seed(5)
    import io, contextlib
    with contextlib.redirect_stdout(io.StringIO()):
        ref_sel = Mapper.keyframe_selection_overlap(self2, color_t, depth_t, cur, kd, 4, N_samples=16, pixels=100)
    RM.get_samples = keep
    mine_sel = M.select_overlapping_keyframes(fr, 4, np.random.RandomState(5))
    print("keyframe_selection_overlap:", list(map(int, ref_sel)), "restatement:", list(map(int, mine_sel)))
    assert list(map(int, ref_sel)) == list(map(int, mine_sel))
    out["overlap/rays_o"], out["overlap/rays_d"], out["overlap/gt_depth"] = ro.numpy(), rd.numpy(), gd.numpy()
    out["overlap/kf_ids"], out["overlap/fractions"], out["overlap/selected"] = np.array(kf_ids), fr, np.array(ref_sel)
    out["overlap/vertices"] = verts

    # ---- ray pre-filter (Mapper.py:607-621 verbatim) ----
    ro5 = torch.cat([ro, ro[:20] + torch.tensor([0.0, 0.0, 30.0])]); rd5 = torch.cat([rd, rd[:20]]); gd5 = torch.cat([gd * 3.0, gd[:20]])
    det_rays_o = ro5.clone().detach().unsqueeze(-1)
    det_rays_d = rd5.clone().detach().unsqueeze(-1)
    t = (bound.unsqueeze(0) - det_rays_o) / det_rays_d
    t, _ = torch.min(torch.max(t, dim=2)[0], dim=1)
    inside_ref = t >= gd5
    assert torch.equal(inside_ref, M.ray_prefilter_mask(ro5, rd5, gd5, bound))
    print(f"ray pre-filter: {int(inside_ref.sum())} of {inside_ref.numel()} rays kept")
    out["prefilter/rays_o"], out["prefilter/rays_d"], out["prefilter/gt_depth"] = ro5.numpy(), rd5.numpy(), gd5.numpy()
    out["prefilter/mask"] = inside_ref.numpy()

    # ---- masked Adam = torch.optim.Adam on val[mask] with the stage learning rates ----
    g = torch.Generator().manual_seed(31)
    val = torch.randn(1, 32, 5, 6, 7, generator=g) * 0.01
    mask3 = torch.rand(5, 6, 7, generator=g) < 0.4
    mask = mask3[None, None].repeat(1, 32, 1, 1, 1)
    val_ref = val.clone()
    val_grad = torch.nn.Parameter(val_ref[mask].clone())
    opt = torch.optim.Adam([{"params": [val_grad], "lr": 0}])
    mine_p, mine_m, mine_v = val.clone(), torch.zeros_like(val), torch.zeros_like(val)
    grads = []
    for it in range(6):
        lr = M.STAGE_LR["middle" if it < 3 else "color"]["middle_lr"] * 5
        opt.param_groups[0]["lr"] = lr
        gr = torch.randn(val.shape, generator=g)
        gr[:, :, it % 5] = 0.0          # masked voxels without gradient still move (Adam momentum)
        grads.append(gr)
        val_grad.grad = gr[mask].clone()
        opt.step()
        val_ref[mask] = val_grad.detach()
        M.adam_step(mine_p, gr, mine_m, mine_v, it + 1, lr, mask=mask)
    err = (val_ref - mine_p).abs().max().item()
    print(f"masked Adam vs torch.optim.Adam on val[mask]: max abs diff {err:.3e}")
    assert err == 0.0
    out["adam/val0"], out["adam/mask"], out["adam/grads"], out["adam/val6"] = val.numpy(), mask3.numpy(), torch.stack(grads).numpy(), val_ref.numpy()

    # ---- Mesher.point_masks ----
    gp = torch.Generator().manual_seed(41)
    lo, hi = bound[:, 0].float(), bound[:, 1].float()
    pts = lo + (hi - lo) * torch.rand(20000, 3, generator=gp)
    kfs = [0, 10, 20, 30]
    kdict = [{"est_c2w": torch.from_numpy(poses[i]), "depth": synthetic_depth(50 + i)} for i in kfs]
    for dt in (False, True):
        selfm = types.SimpleNamespace(H=H, W=W, fx=FX, fy=FY, cx=CX, cy=CY, points_batch_size=500000, depth_test=dt)
        seen, fore, unseen = Mesher.point_masks(selfm, pts, kdict, None, 0, "cpu")
        s2, f2, u2 = M.point_masks(pts, [k["est_c2w"] for k in kdict], [k["depth"] for k in kdict], H, W, FX, FY, CX, CY, dt)
        assert (seen == s2).all() and (fore == f2).all() and (unseen == u2).all()
        print(f"point_masks depth_test={dt}: seen {int(seen.sum())}, forecast {int(fore.sum())}, unseen {int(unseen.sum())} (bit-equal)")
        out[f"pmask/{int(dt)}/seen"], out[f"pmask/{int(dt)}/forecast"] = np.packbits(seen), np.packbits(fore)
    out["pmask/points"], out["pmask/kf_ids"] = pts.numpy(), np.array(kfs)

    os.makedirs(args.out, exist_ok=True)
    path = os.path.join(args.out, "mapper.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
