#!/usr/bin/env python
"""Pin ``oracle/nice_oracle.py`` against the reference itself and mint goldens.

Runs ONLY in the build container (needs ``/root/reference``).  It imports the
reference's own ``src.common``, ``src.conv_onet.models.decoder``,
``src.utils.Renderer`` and ``src.config`` unmodified, feeds them and the oracle
the same seeded inputs on the CPU, and requires *bit-equal* results (same ATen
ops in the same order).  The reference outputs are then written to
``tests/golden/*.npz`` so that the pin can be replayed on any box
(``tests/test_oracle_golden.py``) where ``/root/reference`` does not exist.

CPU shims the reference needs (it is written for CUDA):
  * ``NICE.forward`` builds ``'cuda:-1'`` for stages other than ``color``
    (decoder.py:316-331): those stages are assembled here from the reference's
    live sub-decoders exactly as decoder.py:317-335 does;
  * ``quad2rotation`` calls ``.to(quad.get_device())`` (common.py:150):
    ``Tensor.get_device`` is patched to return ``'cpu'`` for that call.

usage:  python oracle/pin_against_reference.py [--out tests/golden]
"""
import argparse
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.dont_write_bytecode = True

from oracle import nice_oracle as O  # noqa: E402


def import_reference():
    os.chdir(REF)  # inherit_from paths in the yaml are cwd-relative (src/config.py:23)
    sys.path.insert(0, REF)
    from src import config as rconfig
    from src import common as rcommon
    from src.utils.Renderer import Renderer
    return rconfig, rcommon, Renderer


def bit_equal(a, b, what):
    a, b = a.detach(), b.detach()
    assert a.shape == b.shape and a.dtype == b.dtype, (what, a.shape, b.shape, a.dtype, b.dtype)
    if not torch.equal(a, b):
        err = (a.double() - b.double()).abs().max().item()
        raise AssertionError(f"{what}: oracle differs from reference, max abs err {err:g}")
    print(f"  pinned bit-equal: {what} {tuple(a.shape)} {a.dtype}")


def ref_nice_stage(model, p, c, stage):
    """decoder.py:317-342 with the ``.to('cuda:-1')`` removed."""
    if stage == "color":
        return model(p, c_grid=c, stage="color")
    if stage == "coarse":
        occ = model.coarse_decoder(p, c).squeeze(0)
        raw = torch.zeros(occ.shape[0], 4)
        raw[..., -1] = occ
        return raw
    mid = model.middle_decoder(p, c).squeeze(0)
    raw = torch.zeros(mid.shape[0], 4)
    raw[..., -1] = mid if stage == "middle" else model.fine_decoder(p, c) + mid
    return raw


class StageShim(torch.nn.Module):
    """Lets the unmodified reference Renderer drive stages that crash on CPU."""

    def __init__(self, model):
        super().__init__()
        self.model = model

    def forward(self, p, c_grid=None, stage="color", **kw):
        return ref_nice_stage(self.model, p, c_grid, stage)


def synthetic_frame(H, W, seed):
    g = torch.Generator().manual_seed(seed)
    depth = 1.0 + 2.0 * torch.rand(H, W, generator=g)
    depth[torch.rand(H, W, generator=g) < 0.02] = 0.0
    color = torch.rand(H, W, 3, generator=g)
    return depth, color


def to_np(d):
    return {k: v.detach().cpu().numpy() for k, v in d.items()}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(os.path.dirname(HERE), "tests", "golden"))
    args = ap.parse_args()
    out_dir = os.path.abspath(args.out)
    rconfig, rcommon, Renderer = import_reference()
    torch.set_num_threads(1)  # goldens independent of the sgemm thread split
    cfg = rconfig.load_config("configs/Replica/room0.yaml", "configs/nice_slam.yaml")

    # ---- bound / grid shapes: SURVEY 8(a-0) probe values -------------------
    b0 = O.scene_bound(cfg["mapping"]["bound"], cfg["scale"], cfg["grid_len"]["bound_divisible"])
    assert b0[:, 1].tolist() == [8.94000015258789, 5.7600000381469725, 3.5399999618530273], b0
    sh = O.grid_shapes(b0, cfg["grid_len"])
    assert sh == {"grid_coarse": (7, 8, 11), "grid_middle": (21, 28, 37), "grid_fine": (43, 56, 74),
                  "grid_color": (43, 56, 74)}, sh
    print("  pinned: room0 bound + grid shapes")

    # ---- small scene shared by both sides ----------------------------------
    bound = O.scene_bound([[-1.1, 1.4], [-0.9, 1.3], [-1.2, 0.7]], 1.0, 0.32)
    gen = torch.Generator().manual_seed(7)
    grids = O.init_grids(bound, cfg["grid_len"], 32, 2, True, gen)
    for k in grids:  # larger features so that every decoder path is observable
        grids[k] = grids[k] * 30.0
    torch.manual_seed(0)
    model = rconfig.get_model(cfg, nice=True)
    with torch.no_grad():
        for n_, p_ in model.named_parameters():
            if n_.endswith("bias"):
                p_.add_(0.05 * torch.randn(p_.shape))
    bounds = O.decoder_bounds(bound)
    model.bound = bound
    model.middle_decoder.bound = bound
    model.fine_decoder.bound = bound
    model.color_decoder.bound = bound
    model.coarse_decoder.bound = bound * 2
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}

    H, W, fx, fy, cx, cy = 68, 120, 60.0, 60.0, 59.5, 33.5
    slam = types.SimpleNamespace(bound=bound, H=H, W=W, fx=fx, fy=fy, cx=cx, cy=cy)
    renderer = Renderer(cfg, None, slam)
    renderer.nice = True
    shim = StageShim(model)
    scene = O.Scene(sd, grids, bound, nice=True, occupancy=True, n_samples=32, n_surface=16)

    # ---- case A: eval_points, all stages, edge cases ----------------------
    g = torch.Generator().manual_seed(11)
    lo, hi = bound[:, 0], bound[:, 1]
    pts = lo + (hi - lo) * torch.rand(1500, 3, generator=g).double()
    outside = lo - 0.3 + (hi - lo + 0.6) * torch.rand(300, 3, generator=g).double()
    face = pts[:50].clone(); face[:, 0] = hi[0]                      # exactly on a face => outside
    last = pts[50:100].clone(); last[:, 2] = hi[2] - 1e-4 * torch.rand(50, generator=g).double()
    pA = torch.cat([pts, outside, face, last], 0)
    gold = {"bound": bound, "points": pA}
    for k, v in grids.items():
        gold[k] = v
    for k, v in sd.items():
        gold["sd/" + k] = v
    for stage in O.STAGES:
        with torch.no_grad():
            r = renderer.eval_points(pA.clone(), shim, grids, stage, "cpu")
            o = O.eval_points(scene, pA.clone(), stage)
        bit_equal(o, r, f"eval_points[{stage}] f64 points")
        gold[f"raw_{stage}"] = r
        with torch.no_grad():
            r32 = renderer.eval_points(pA.float(), shim, grids, stage, "cpu")
            o32 = O.eval_points(scene, pA.float(), stage)
        bit_equal(o32, r32, f"eval_points[{stage}] f32 points")
        gold[f"raw32_{stage}"] = r32
    np.savez_compressed(os.path.join(out_dir, "nice_eval_points.npz"), **to_np(gold))

    # ---- case B: get_samples (bit-exact indices) + pose --------------------
    depth, color = synthetic_frame(H, W, 3)
    cam = torch.tensor([0.98, 0.05, -0.12, 0.08, 0.15, 0.2, -0.1])
    orig_get_device = torch.Tensor.get_device
    torch.Tensor.get_device = lambda self: "cpu"
    try:
        c2w_ref = rcommon.get_camera_from_tensor(cam)
    finally:
        torch.Tensor.get_device = orig_get_device
    bit_equal(O.camera_from_tensor(cam), c2w_ref, "get_camera_from_tensor")
    torch.manual_seed(5)
    ro_r, rd_r, d_r, c_r = rcommon.get_samples(8, H - 8, 10, W - 10, 96, H, W, fx, fy, cx, cy, c2w_ref,
                                               depth, color, "cpu")
    torch.manual_seed(5)
    ro_o, rd_o, d_o, c_o, idx = O.get_samples(8, H - 8, 10, W - 10, 96, fx, fy, cx, cy, c2w_ref, depth, color)
    for a, b, w in ((ro_o, ro_r, "rays_o"), (rd_o, rd_r, "rays_d"), (d_o, d_r, "depth"), (c_o, c_r, "color")):
        bit_equal(a, b, "get_samples." + w)
    ro_f, rd_f = rcommon.get_rays(H, W, fx, fy, cx, cy, c2w_ref, "cpu")
    ro_g, rd_g = O.get_rays(H, W, fx, fy, cx, cy, c2w_ref)
    bit_equal(rd_g, rd_f, "get_rays.rays_d"); bit_equal(ro_g, ro_f, "get_rays.rays_o")

    # ---- case C: render_batch_ray fwd + bwd, every stage -------------------
    goldC = {"cam": cam, "depth_img": depth, "color_img": color, "indices": idx,
             "intr": torch.tensor([H, W, fx, fy, cx, cy]), "crop": torch.tensor([8, H - 8, 10, W - 10]),
             "rays_o": ro_r, "rays_d": rd_r, "gt_depth": d_r, "gt_color": c_r}

    def run_ref(stage, gt, mode):
        gr = {k: v.clone().requires_grad_(mode == "map") for k, v in grids.items()}
        model.zero_grad()
        camt = cam.clone().requires_grad_(mode == "track")
        torch.Tensor.get_device = lambda self: "cpu"
        try:
            c2w = rcommon.get_camera_from_tensor(camt)
        finally:
            torch.Tensor.get_device = orig_get_device
        i, j = O.pixel_lattice(8, H - 8, 10, W - 10)
        ro, rd = rcommon.get_rays_from_uv(i[idx], j[idx], c2w, H, W, fx, fy, cx, cy, "cpu")
        dep, var, col = renderer.render_batch_ray(gr, shim, rd, ro, "cpu", stage, gt_depth=gt)
        return dep, var, col, gr, camt

    def run_orc(stage, gt, mode):
        gr = {k: v.clone().requires_grad_(mode == "map") for k, v in grids.items()}
        sdo = {k: v.clone().requires_grad_(mode == "map") for k, v in sd.items()}
        camt = cam.clone().requires_grad_(mode == "track")
        c2w = O.camera_from_tensor(camt)
        i, j = O.pixel_lattice(8, H - 8, 10, W - 10)
        ro, rd = O.rays_from_pixels(i[idx], j[idx], c2w, fx, fy, cx, cy)
        sc = O.Scene(sdo, gr, bound, nice=True, occupancy=True)
        dep, var, col = O.render_batch_ray(sc, rd, ro, stage, gt)
        return dep, var, col, gr, camt, sdo

    for stage in O.STAGES:
        for mode in ("map", "track"):
            gt = None if stage == "coarse" else d_r
            dr, vr, cr, gr_r, cam_r = run_ref(stage, gt, mode)
            do, vo, co, gr_o, cam_o, sdo = run_orc(stage, gt, mode)
            tag = f"render[{stage},{mode}]"
            bit_equal(do, dr, tag + ".depth"); bit_equal(vo, vr, tag + ".var"); bit_equal(co, cr, tag + ".color")
            gtd = d_r if gt is not None else torch.ones_like(d_r)
            if mode == "map":
                lr = O.mapping_loss(dr, cr, gtd, c_r, stage)
                lo_ = O.mapping_loss(do, co, gtd, c_r, stage)
            else:
                if stage != "color":
                    continue
                lr = O.tracking_loss(dr, vr, cr, gtd, c_r)
                lo_ = O.tracking_loss(do, vo, co, gtd, c_r)
            lr.backward(); lo_.backward()
            goldC[f"{stage}/{mode}/depth"], goldC[f"{stage}/{mode}/var"], goldC[f"{stage}/{mode}/color"] = dr, vr, cr
            goldC[f"{stage}/{mode}/loss"] = lr.detach()
            if mode == "track":
                bit_equal(cam_o.grad, cam_r.grad, tag + ".grad_cam")
                goldC[f"{stage}/track/grad_cam"] = cam_r.grad
            else:
                for k in grids:
                    if gr_r[k].grad is None:
                        assert gr_o[k].grad is None, (tag, k)
                        continue
                    bit_equal(gr_o[k].grad, gr_r[k].grad, f"{tag}.grad_{k}")
                    goldC[f"{stage}/map/grad_{k}"] = gr_r[k].grad
                for n_, p_ in model.named_parameters():
                    if p_.grad is None:
                        assert sdo[n_].grad is None, (tag, n_)
                        continue
                    bit_equal(sdo[n_].grad, p_.grad, f"{tag}.grad_{n_}")
                    goldC[f"{stage}/map/gradsd/{n_}"] = p_.grad.clone()
    # no-depth render of the non-coarse stages (N_surface -> 0, near = 0.01)
    with torch.no_grad():
        dr, vr, cr = renderer.render_batch_ray(grids, shim, rd_r, ro_r, "cpu", "color", gt_depth=None)
        do, vo, co = O.render_batch_ray(scene, rd_r, ro_r, "color", None)
    bit_equal(do, dr, "render[color,nodepth].depth"); bit_equal(co, cr, "render[color,nodepth].color")
    goldC["color/nodepth/depth"], goldC["color/nodepth/var"], goldC["color/nodepth/color"] = dr, vr, cr
    # ray origin outside the bound: far_bb negative -> clamps to 0, no NaN
    ro_out = ro_r.clone(); ro_out[:, 0] += 5.0
    with torch.no_grad():
        dr, vr, cr = renderer.render_batch_ray(grids, shim, rd_r, ro_out, "cpu", "color", gt_depth=d_r)
        do, vo, co = O.render_batch_ray(scene, rd_r, ro_out, "color", d_r)
    assert torch.isfinite(dr).all()
    bit_equal(do, dr, "render[color,outside].depth"); bit_equal(co, cr, "render[color,outside].color")
    goldC["color/outside/depth"], goldC["color/outside/var"], goldC["color/outside/color"] = dr, vr, cr
    np.savez_compressed(os.path.join(out_dir, "nice_render.npz"), **to_np(goldC))

    # ---- case D: iMAP* with the shipped trained weights --------------------
    ckpt = torch.load(os.path.join(REF, "output_imap/Replica/room0/ckpts/01999.tar"), map_location="cpu",
                      weights_only=False)
    icfg = rconfig.load_config("configs/Replica/room0.yaml", "configs/imap.yaml")
    imodel = rconfig.get_model(icfg, nice=False)
    imodel.load_state_dict(ckpt["decoder_state_dict"])
    isd = {k: v.detach().clone() for k, v in imodel.state_dict().items()}
    ibound = O.scene_bound(icfg["mapping"]["bound"], icfg["scale"], icfg["grid_len"]["bound_divisible"])
    r = icfg["rendering"]
    islam = types.SimpleNamespace(bound=ibound, H=H, W=W, fx=fx, fy=fy, cx=cx, cy=cy)
    irend = Renderer(icfg, None, islam)
    assert irend.nice is False and irend.occupancy is False
    iscene = O.Scene(isd, {}, ibound, nice=False, occupancy=False, n_samples=r["N_samples"],
                     n_surface=r["N_surface"], n_importance=r["N_importance"])
    c2w = ckpt["gt_c2w_list"][100][:3].float()
    idepth = depth * icfg["scale"]
    torch.manual_seed(9)
    ro, rd, gd, gc, iidx = O.get_samples(0, H, 0, W, 80, fx, fy, cx, cy, c2w, idepth, color)
    goldD = {"bound": ibound, "rays_o": ro, "rays_d": rd, "gt_depth": gd, "gt_color": gc,
             "n_samples": torch.tensor(r["N_samples"]), "n_surface": torch.tensor(r["N_surface"]),
             "n_importance": torch.tensor(r["N_importance"])}
    for k, v in isd.items():
        goldD["sd/" + k] = v
    imodel.zero_grad()
    dr, vr, cr = irend.render_batch_ray({}, imodel, rd, ro, "cpu", "color", gt_depth=gd)
    sdo = {k: v.clone().requires_grad_(True) for k, v in isd.items()}
    iscene.sd = sdo
    do, vo, co = O.render_batch_ray(iscene, rd, ro, "color", gd)
    bit_equal(do, dr, "imap.depth"); bit_equal(vo, vr, "imap.var"); bit_equal(co, cr, "imap.color")
    torch.manual_seed(21)
    sig_r = irend.regulation({}, imodel, rd, ro, gd, "cpu", "color")
    torch.manual_seed(21)
    t_rand = torch.rand(gd.shape[0], r["N_samples"])
    sig_o = O.regulation(iscene, rd, ro, gd, "color", t_rand=t_rand)
    bit_equal(sig_o, sig_r, "imap.regulation")
    lr = O.mapping_loss(dr, cr, gd, gc, "color", 0.05, nice=False) + 0.0005 * sig_r.abs().sum()
    lo_ = O.mapping_loss(do, co, gd, gc, "color", 0.05, nice=False) + 0.0005 * sig_o.abs().sum()
    lr.backward(); lo_.backward()
    for n_, p_ in imodel.named_parameters():
        bit_equal(sdo[n_].grad, p_.grad, "imap.grad_" + n_)
        goldD["gradsd/" + n_] = p_.grad.clone()
    goldD.update({"depth": dr, "var": vr, "color": cr, "reg_t_rand": t_rand, "reg_sigma": sig_r, "loss": lr.detach()})
    np.savez_compressed(os.path.join(out_dir, "imap_render.npz"), **to_np(goldD))
    print("oracle pinned against the reference; goldens written to", out_dir)


if __name__ == "__main__":
    main()
