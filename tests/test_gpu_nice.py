"""GPU parity of the NICE path: CUDA kernels (through the C ABI) vs the CPU
oracle on identical inputs, and vs the reference's golden vectors.

Tolerances (stated per check):
  * integer / index work and float64 geometry: bit-exact (pixel indices, rays,
    z-values, out-of-bound mask);
  * decoder outputs and rendered depth / colour: |err| <= 3e-4 + 1e-3*|ref|.
    The floor is the float32 Fourier argument p@B (|arg| up to a few hundred
    rad, 1 ulp ~ 2e-5) whose summation order differs between sgemm and FFMA;
  * gradients: max-abs error <= 2e-3 of the gradient's max-abs (atomic order).
"""
import pytest
import torch

from oracle import nice_oracle as O
from tests import helpers as T

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
OUT_TOL = dict(rtol=1e-3, atol=3e-4)
GRAD_REL = 2e-3


@pytest.fixture(scope="module")
def g():
    return T.load_nice()


@pytest.fixture(scope="module")
def scene(g):
    return T.cuda_nice(g, DEV)


def test_eval_points_stages(g, scene):
    model, grids, renderer = scene
    for stage in O.STAGES:
        for key, pts in ((f"raw_{stage}", g["points"]), (f"raw32_{stage}", g["points"].float())):
            with torch.no_grad():
                out = renderer.eval_points(pts.to(DEV), model, grids, stage, DEV).cpu()
            ref = g[key]
            assert torch.equal(out[:, 3] == 100, ref[:, 3] == 100), f"{key}: out-of-bound mask differs"
            torch.testing.assert_close(out, ref, **OUT_TOL, msg=lambda m: f"{key}: {m}")


def test_eval_points_contiguous_grids_match_channels_last(g, scene):
    model, grids, renderer = scene
    _, grids_nc, _ = T.cuda_nice(g, DEV, channels_last=False)
    p = g["points"].to(DEV)
    with torch.no_grad():
        a = renderer.eval_points(p, model, grids, "color", DEV)
        b = renderer.eval_points(p, model, grids_nc, "color", DEV)
    assert torch.equal(a, b)


def test_decoder_modules_direct(g, scene):
    """NICE.forward / sub-decoder forward (no bound override), decoder.py:312-342."""
    model, grids, _ = scene
    sd = T.state_dict(g)
    cpu_grids = {k: g[k] for k in O.GRID_KEYS}
    bounds = O.decoder_bounds(g["bound"])
    p = g["points"][:1500]
    with torch.no_grad():
        ref = O.nice_forward(sd, p.clone(), cpu_grids, bounds, "color")
        out = model(p.to(DEV)[None], grids, stage="color").cpu()
        torch.testing.assert_close(out, ref, **OUT_TOL)
        ref_mid = O.middle_occ(sd, p.clone(), cpu_grids, bounds)
        out_mid = model.middle_decoder(p.to(DEV)[None], grids).cpu()
        torch.testing.assert_close(out_mid, ref_mid, **OUT_TOL)
        ref_fine = O.fine_occ(sd, p.clone(), cpu_grids, bounds)
        out_fine = model.fine_decoder(p.to(DEV)[None], grids).cpu()
        torch.testing.assert_close(out_fine, ref_fine, **OUT_TOL)
        ref_c = O.coarse_mlp(sd, "coarse_decoder.", O.trilinear_feature(p.clone(), cpu_grids["grid_coarse"], bounds["coarse"]))
        out_c = model.coarse_decoder(p.to(DEV)[None], grids).cpu()
        torch.testing.assert_close(out_c, ref_c, **OUT_TOL)


def _rays(g, cam):
    import pointnerf_slam_b200 as P
    H, W, fx, fy, cx, cy = [float(v) for v in g["intr"]]
    H0, H1, W0, W1 = [int(v) for v in g["crop"]]
    c2w = P.get_camera_from_tensor(cam)
    return P.get_samples(H0, H1, W0, W1, g["indices"].numel(), int(H), int(W), fx, fy, cx, cy, c2w,
                         g["depth_img"].to(DEV), g["color_img"].to(DEV), DEV, indices=g["indices"].to(DEV))


def test_pose_rays_and_gathers_bit_exact(g):
    import pointnerf_slam_b200 as P
    cam = g["cam"].to(DEV)
    c2w = P.get_camera_from_tensor(cam)
    assert torch.equal(c2w.cpu(), O.camera_from_tensor(g["cam"]))
    ro, rd, gd, gc = _rays(g, cam)
    assert torch.equal(ro.cpu(), g["rays_o"]) and torch.equal(rd.cpu(), g["rays_d"])
    assert torch.equal(gd.cpu(), g["gt_depth"]) and torch.equal(gc.cpu(), g["gt_color"])
    H, W, fx, fy, cx, cy = [float(v) for v in g["intr"]]
    ro_f, rd_f = P.get_rays(int(H), int(W), fx, fy, cx, cy, c2w, DEV)
    ro_r, rd_r = O.get_rays(int(H), int(W), fx, fy, cx, cy, O.camera_from_tensor(g["cam"]))
    assert torch.equal(rd_f.cpu(), rd_r) and torch.equal(ro_f.cpu(), ro_r)


def test_pixel_indices_follow_torch_randint():
    """The draw is torch.randint on the device, as src/common.py:99 does."""
    import pointnerf_slam_b200 as P
    depth = torch.rand(68, 120, device=DEV)
    color = torch.rand(68, 120, 3, device=DEV)
    c2w = torch.eye(4, device=DEV)[:3]
    torch.manual_seed(123)
    idx = torch.randint(52 * 100, (77,), device=DEV)
    torch.manual_seed(123)
    ro, rd, d, c = P.get_samples(8, 60, 10, 110, 77, 68, 120, 60., 60., 59.5, 33.5, c2w, depth, color, DEV)
    assert torch.equal(d, depth[8:60, 10:110].reshape(-1)[idx])
    assert torch.equal(c, color[8:60, 10:110].reshape(-1, 3)[idx])


def test_z_values_bit_exact(g, scene):
    _, _, renderer = scene
    sc = T.oracle_scene(g)
    for gt in (g["gt_depth"], None):
        z = renderer.sample_z(g["rays_d"].to(DEV), g["rays_o"].to(DEV), gt.to(DEV) if gt is not None else None).cpu()
        ref = O.ray_z_values(sc, g["rays_o"], g["rays_d"], gt)
        assert z.dtype == ref.dtype == torch.float64
        assert torch.equal(z, ref)
    # origin outside the bound: far clamps to 0, descending stratified list, still sorted identically
    ro = g["rays_o"].clone(); ro[:, 0] += 5.0
    z = renderer.sample_z(g["rays_d"].to(DEV), ro.to(DEV), g["gt_depth"].to(DEV)).cpu()
    assert torch.equal(z, O.ray_z_values(sc, ro, g["rays_d"], g["gt_depth"]))


@pytest.mark.parametrize("stage", ["coarse", "middle", "fine", "color"])
def test_render_forward_vs_golden(g, scene, stage):
    model, grids, renderer = scene
    gt = None if stage == "coarse" else g["gt_depth"].to(DEV)
    with torch.no_grad():
        d, v, c = renderer.render_batch_ray(grids, model, g["rays_d"].to(DEV), g["rays_o"].to(DEV), DEV, stage, gt_depth=gt)
    assert d.dtype == torch.float64 and v.dtype == torch.float64 and c.dtype == torch.float32
    torch.testing.assert_close(d.cpu(), g[f"{stage}/map/depth"], **OUT_TOL)
    torch.testing.assert_close(v.cpu(), g[f"{stage}/map/var"], rtol=2e-3, atol=3e-4)
    torch.testing.assert_close(c.cpu(), g[f"{stage}/map/color"], **OUT_TOL)


def test_render_edge_cases(g, scene):
    model, grids, renderer = scene
    rd, ro = g["rays_d"].to(DEV), g["rays_o"].to(DEV)
    with torch.no_grad():
        d, v, c = renderer.render_batch_ray(grids, model, rd, ro, DEV, "color", gt_depth=None)
        torch.testing.assert_close(d.cpu(), g["color/nodepth/depth"], **OUT_TOL)
        torch.testing.assert_close(c.cpu(), g["color/nodepth/color"], **OUT_TOL)
        ro2 = ro.clone(); ro2[:, 0] += 5.0
        d, v, c = renderer.render_batch_ray(grids, model, rd, ro2, DEV, "color", gt_depth=g["gt_depth"].to(DEV))
        assert torch.isfinite(d).all()
        torch.testing.assert_close(d.cpu(), g["color/outside/depth"], **OUT_TOL)
        torch.testing.assert_close(c.cpu(), g["color/outside/color"], **OUT_TOL)
        # empty batch
        d, v, c = renderer.render_batch_ray(grids, model, rd[:0], ro[:0], DEV, "color", gt_depth=g["gt_depth"].to(DEV)[:0])
        assert d.shape == (0,) and c.shape == (0, 3)


@pytest.mark.parametrize("stage", ["coarse", "middle", "fine", "color"])
def test_mapping_gradients_vs_golden(g, stage):
    """Mapper.py:628-662 iteration: grads into grids and decoder parameters."""
    model, grids, renderer = T.cuda_nice(g, DEV)
    for k in grids:
        grids[k].requires_grad_(True)
    gt = None if stage == "coarse" else g["gt_depth"].to(DEV)
    d, v, c = renderer.render_batch_ray(grids, model, g["rays_d"].to(DEV), g["rays_o"].to(DEV), DEV, stage, gt_depth=gt)
    gtd = g["gt_depth"].to(DEV) if gt is not None else torch.ones_like(g["gt_depth"]).to(DEV)
    loss = O.mapping_loss(d, c, gtd, g["gt_color"].to(DEV), stage)
    torch.testing.assert_close(loss.cpu(), g[f"{stage}/map/loss"], rtol=1e-3, atol=1e-3)
    loss.backward()
    for k in O.GRID_KEYS:
        key = f"{stage}/map/grad_{k}"
        if key in g:
            assert grids[k].grad is not None, key
            assert T.rel_max(grids[k].grad, g[key]) < GRAD_REL, key
        else:
            assert grids[k].grad is None or float(grids[k].grad.abs().max()) == 0.0, f"{k} must get no gradient in {stage}"
    checked = 0
    for name, p in model.named_parameters():
        key = f"{stage}/map/gradsd/{name}"
        if key in g:
            assert p.grad is not None, key
            assert T.rel_max(p.grad, g[key]) < GRAD_REL, key
            checked += 1
        else:
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, name
    assert checked > 0
    if stage == "color":  # 4th output row of the colour decoder is overwritten => zero gradient
        assert float(model.color_decoder.output_linear.weight.grad[3].abs().max()) == 0.0


def test_tracking_gradient_to_camera(g, scene):
    """Tracker.py:269-335 iteration: grads to the camera 7-vector only."""
    model, grids, renderer = scene
    renderer.freeze_map = True
    try:
        cam = g["cam"].to(DEV).requires_grad_(True)
        ro, rd, gd, gc = _rays(g, cam)
        d, v, c = renderer.render_batch_ray(grids, model, rd, ro, DEV, "color", gt_depth=gd)
        loss = O.tracking_loss(d, v, c, gd, gc)
        loss.backward()
    finally:
        renderer.freeze_map = False
    torch.testing.assert_close(loss.cpu(), g["color/track/loss"], rtol=2e-3, atol=1e-3)
    assert T.rel_max(cam.grad, g["color/track/grad_cam"]) < 5e-3
    assert all(p.grad is None for p in model.parameters())


def test_ray_shards_sum_to_full_gradient(g):
    """Losses are plain sums, so summed shard gradients equal the un-sharded
    gradient when the batch-global max depth is shared."""
    model, grids, renderer = T.cuda_nice(g, DEV)
    for k in grids:
        grids[k].requires_grad_(True)
    rd, ro, gd, gc = g["rays_d"].to(DEV), g["rays_o"].to(DEV), g["gt_depth"].to(DEV), g["gt_color"].to(DEV)
    d, v, c = renderer.render_batch_ray(grids, model, rd, ro, DEV, "color", gt_depth=gd)
    O.mapping_loss(d, c, gd, gc, "color").backward()
    full = {k: grids[k].grad.clone() for k in grids if grids[k].grad is not None}
    for k in grids:
        grids[k].grad = None
    renderer.depth_max_override = gd.max().reshape(1)
    try:
        for sl in (slice(0, 40), slice(40, 96)):
            d, v, c = renderer.render_batch_ray(grids, model, rd[sl], ro[sl], DEV, "color", gt_depth=gd[sl])
            O.mapping_loss(d, c, gd[sl], gc[sl], "color").backward()
    finally:
        renderer.depth_max_override = None
    for k, f in full.items():
        assert T.rel_max(grids[k].grad, f) < 1e-5, k


def test_composite_standalone_matches_oracle(g):
    import pointnerf_slam_b200 as P
    torch.manual_seed(0)
    R, S = 50, 48
    raw = torch.randn(R, S, 4)
    z = torch.sort(torch.rand(R, S, dtype=torch.float64) * 3, -1)[0]
    rd = torch.randn(R, 3)
    for occ in (True, False):
        rr = raw.clone().requires_grad_(True)
        d, v, c, w = O.composite(rr, z, rd, occ)
        (d.sum() + 0.3 * v.sum() + (c * torch.tensor([1.0, -2.0, 0.5])).sum()).backward()
        rc = raw.clone().to(DEV).requires_grad_(True)
        d2, v2, c2, w2 = P.raw2outputs_nerf_color(rc, z.to(DEV), rd.to(DEV), occupancy=occ, device=DEV)
        (d2.sum() + 0.3 * v2.sum() + (c2 * torch.tensor([1.0, -2.0, 0.5], device=DEV)).sum()).backward()
        torch.testing.assert_close(d2.cpu(), d.detach(), rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(v2.cpu(), v.detach(), rtol=1e-4, atol=1e-6)
        torch.testing.assert_close(c2.cpu(), c.detach(), rtol=1e-4, atol=1e-5)
        torch.testing.assert_close(w2.cpu(), w.detach(), rtol=1e-4, atol=1e-7)
        assert T.rel_max(rc.grad, rr.grad) < 1e-4


def test_graphed_step_replays_the_eager_iteration():
    """graphs.GraphedStep: a captured tracking iteration gives the eager loss and pose gradient for the same pixels."""
    import pointnerf_slam_b200 as P
    d = T.load_nice()
    model, grids, renderer = T.cuda_nice(d)
    dev = torch.device("cuda", 0)
    renderer.freeze_map = True
    Hh, Ww = 60, 80
    depth = (1.0 + torch.rand(Hh, Ww, generator=torch.Generator().manual_seed(3))).to(dev)
    color = torch.rand(Hh, Ww, 3, generator=torch.Generator().manual_seed(4)).to(dev)
    c2w = torch.eye(4)[:3].clone()
    c2w[:, 3] = torch.tensor([0.1, 0.2, -0.1])
    cam = P.get_tensor_from_camera(c2w).to(dev).requires_grad_(True)
    idx = torch.randint(Hh * Ww, (200,), generator=torch.Generator().manual_seed(5)).to(dev)

    def iteration():
        c = P.get_camera_from_tensor(cam)
        o, dr, gd, gc = P.get_samples(0, Hh, 0, Ww, 200, Hh, Ww, 70.0, 70.0, 39.5, 29.5, c, depth, color, dev, indices=idx)
        dd, vv, cc = renderer.render_batch_ray(grids, model, dr, o, dev, "color", gt_depth=gd)
        loss = torch.abs(gd - dd).sum() + 0.5 * torch.abs(gc - cc).sum()
        cam.grad = None
        loss.backward()
        return loss

    ref_loss = iteration().item()
    ref_grad = cam.grad.clone()
    step = P.graphs.GraphedStep(iteration)
    assert step.launches > 5
    for _ in range(2):
        out = step()
    torch.cuda.synchronize()
    assert abs(out.item() - ref_loss) <= 1e-6 * abs(ref_loss) + 1e-9
    torch.testing.assert_close(cam.grad, ref_grad, rtol=1e-4, atol=1e-7)
    step.release()
    renderer.freeze_map = False


@pytest.mark.parametrize("stage", ["color", "fine"])
def test_mapping_loss_head_matches_oracle(stage):
    """losses.mapping_loss (Mapper.py:628-646): value and gradients vs the oracle's torch expression."""
    import pointnerf_slam_b200 as P
    g = torch.Generator().manual_seed(11)
    R = 5003
    depth = (1.0 + torch.rand(R, generator=g, dtype=torch.float64)).requires_grad_(True)
    color = torch.rand(R, 3, generator=g).requires_grad_(True)
    gt_depth = 1.0 + torch.rand(R, generator=g)
    gt_depth[torch.rand(R, generator=g) < 0.1] = 0.0
    gt_depth[7] = float(depth[7].detach().float())      # an exact tie: sign(0) = 0 in both
    gt_color = torch.rand(R, 3, generator=g)
    ref = O.mapping_loss(depth, color, gt_depth, gt_color, stage, 0.2)
    (3.0 * ref).backward()
    dd = depth.detach().cuda().requires_grad_(True)
    cc = color.detach().cuda().requires_grad_(True)
    out = P.losses.mapping_loss(dd, cc, gt_depth.cuda(), gt_color.cuda(), stage, 0.2)
    assert out.dtype == torch.float64 and out.dim() == 0
    (3.0 * out).backward()
    assert abs(out.item() - ref.item()) <= 2e-6 * abs(ref.item())     # float32 colour sum: summation order differs
    torch.testing.assert_close(dd.grad.cpu(), depth.grad, rtol=0, atol=0)
    if stage == "color":
        torch.testing.assert_close(cc.grad.cpu(), color.grad, rtol=0, atol=0)
    else:
        assert cc.grad is None
    again = P.losses.mapping_loss(dd, cc, gt_depth.cuda(), gt_color.cuda(), stage, 0.2)
    assert again.item() == out.item(), "fixed summation order: run-to-run deterministic"


@pytest.mark.parametrize("handle_dynamic,R", [(True, 1000), (False, 1000), (True, 203), (True, 5000)])
def test_tracking_loss_head_matches_oracle(handle_dynamic, R):
    """losses.tracking_loss (Tracker.py:306-330): value, median-based mask and gradients vs the oracle's torch expression."""
    import pointnerf_slam_b200 as P
    g = torch.Generator().manual_seed(21 + R)
    depth = (1.0 + torch.rand(R, generator=g, dtype=torch.float64)).requires_grad_(True)
    var = 1e-3 + 0.1 * torch.rand(R, generator=g, dtype=torch.float64)
    color = torch.rand(R, 3, generator=g).requires_grad_(True)
    gt_depth = (depth.detach() + 0.05 * torch.randn(R, generator=g, dtype=torch.float64)).float()
    gt_depth[torch.rand(R, generator=g) < 0.1] = 0.0
    big = torch.rand(R, generator=g) < 0.05            # outliers that the median test must drop
    gt_depth[big] += 3.0
    gt_color = torch.rand(R, 3, generator=g)
    ref = O.tracking_loss(depth, var, color, gt_depth, gt_color, 0.5, handle_dynamic)
    (2.0 * ref).backward()
    dd = depth.detach().cuda().requires_grad_(True)
    cc = color.detach().cuda().requires_grad_(True)
    out = P.losses.tracking_loss(dd, var.cuda(), cc, gt_depth.cuda(), gt_color.cuda(), 0.5, True, handle_dynamic)
    assert out.dtype == torch.float64 and out.dim() == 0
    (2.0 * out).backward()
    assert abs(out.item() - ref.item()) <= 2e-6 * abs(ref.item())
    torch.testing.assert_close(dd.grad.cpu(), depth.grad, rtol=1e-14, atol=0)
    torch.testing.assert_close(cc.grad.cpu(), color.grad, rtol=0, atol=0)
    if handle_dynamic:
        assert int((dd.grad == 0).sum()) > int((gt_depth == 0).sum()), "the median test must have dropped outliers"


@pytest.mark.parametrize("R", [8192, 8193, 40000])
def test_tracking_loss_any_number_of_rays(R):
    """Beyond 8192 rays (the fork's tracker passes every pixel with depth, Tracker.py:206-226) the median comes from the
    radix select: same mask and gradients as the oracle, including duplicate values around the median."""
    import pointnerf_slam_b200 as P
    g = torch.Generator().manual_seed(R)
    depth = (1.0 + torch.rand(R, generator=g, dtype=torch.float64)).requires_grad_(True)
    var = 1e-3 + 0.1 * torch.rand(R, generator=g, dtype=torch.float64)
    color = torch.rand(R, 3, generator=g).requires_grad_(True)
    gt_depth = (depth.detach() + 0.05 * torch.randn(R, generator=g, dtype=torch.float64)).float()
    gt_depth[torch.rand(R, generator=g) < 0.1] = 0.0
    gt_depth[torch.rand(R, generator=g) < 0.05] += 3.0
    gt_depth[100:140] = gt_depth[100]                      # ties
    with torch.no_grad():
        depth[100:140] = depth[100]; var[100:140] = var[100]
    gt_color = torch.rand(R, 3, generator=g)
    ref = O.tracking_loss(depth, var, color, gt_depth, gt_color, 0.5, True)
    ref.backward()
    dd, cc = depth.detach().cuda().requires_grad_(True), color.detach().cuda().requires_grad_(True)
    out = P.losses.tracking_loss(dd, var.cuda(), cc, gt_depth.cuda(), gt_color.cuda(), 0.5, True, True)
    out.backward()
    assert abs(out.item() - ref.item()) <= 2e-6 * abs(ref.item())
    torch.testing.assert_close(dd.grad.cpu(), depth.grad, rtol=1e-14, atol=0)
    torch.testing.assert_close(cc.grad.cpu(), color.grad, rtol=0, atol=0)


@pytest.mark.parametrize("R", [500, 9000])
def test_tracking_loss_nan_residual_empties_the_mask(R):
    """torch.median propagates NaN, so `tmp < 10*median` is false everywhere and the loss is 0 (Tracker.py:308-309)."""
    import pointnerf_slam_b200 as P
    g = torch.Generator().manual_seed(5)
    depth = (1.0 + torch.rand(R, generator=g, dtype=torch.float64))
    var = 1e-3 + 0.1 * torch.rand(R, generator=g, dtype=torch.float64)
    var[R // 3] = float("nan")
    color, gt_color = torch.rand(R, 3, generator=g), torch.rand(R, 3, generator=g)
    gt_depth = 1.0 + torch.rand(R, generator=g)
    ref = O.tracking_loss(depth, var, color, gt_depth, gt_color, 0.5, True)
    dd = depth.cuda().requires_grad_(True)
    out = P.losses.tracking_loss(dd, var.cuda(), color.cuda(), gt_depth.cuda(), gt_color.cuda(), 0.5, True, True)
    out.backward()
    assert ref.item() == 0.0 and out.item() == 0.0 and float(dd.grad.abs().sum()) == 0.0


def test_loss_heads_without_depth_supervision():
    """The fork's colour-only branches (Tracker.py:313-318, Mapper.py:633-637)."""
    import pointnerf_slam_b200 as P
    g = torch.Generator().manual_seed(8)
    R = 1500
    depth = (1.0 + torch.rand(R, generator=g, dtype=torch.float64)).requires_grad_(True)
    var = 1e-3 + 0.1 * torch.rand(R, generator=g, dtype=torch.float64)
    color = torch.rand(R, 3, generator=g).requires_grad_(True)
    gt_depth = (depth.detach() + 0.05 * torch.randn(R, generator=g, dtype=torch.float64)).float()
    gt_depth[torch.rand(R, generator=g) < 0.2] = 0.0
    gt_color = torch.rand(R, 3, generator=g)
    for name in ("tracking", "mapping"):
        color.grad = None
        if name == "tracking":
            ref = O.tracking_loss(depth, var, color, gt_depth, gt_color, 0.5, True, depth_supervision=False)
        else:
            ref = O.mapping_loss(depth, color, gt_depth, gt_color, "color", 0.2, depth_supervision=False)
        ref.backward()
        dd, cc = depth.detach().cuda().requires_grad_(True), color.detach().cuda().requires_grad_(True)
        if name == "tracking":
            out = P.losses.tracking_loss(dd, var.cuda(), cc, gt_depth.cuda(), gt_color.cuda(), 0.5, True, True, depth_supervision=False)
        else:
            out = P.losses.mapping_loss(dd, cc, gt_depth.cuda(), gt_color.cuda(), "color", 0.2, depth_supervision=False)
        out.backward()
        assert abs(out.item() - ref.item()) <= 2e-6 * abs(ref.item()), name
        torch.testing.assert_close(cc.grad.cpu(), color.grad, rtol=0, atol=0)
        assert dd.grad is None or float(dd.grad.abs().sum()) == 0.0
    assert depth.grad is None or float(depth.grad.abs().sum()) == 0.0


def test_regulation_carries_the_pose_gradient():
    """Renderer.regulation (Renderer.py:263-301) under bundle adjustment: the loss reaches the rays (and through them
    the camera tensor) via pts = o + d*z (Renderer.py:296-297; used at Mapper.py:650-655).  iMAP* decoder with the
    shipped checkpoint weights."""
    from tests.test_gpu_imap import build
    g = T.load("imap_render.npz")
    model, r = build(g)
    n = 64
    ro, rd, gd = g["rays_o"][:n], g["rays_d"][:n], g["gt_depth"][:n]
    t_rand = torch.rand(n, int(g["n_samples"]), generator=torch.Generator().manual_seed(3))
    roc, rdc = ro.clone().requires_grad_(True), rd.clone().requires_grad_(True)
    scene = O.Scene(T.state_dict(g), {}, g["bound"], nice=False, occupancy=False, n_samples=int(g["n_samples"]),
                    n_surface=int(g["n_surface"]), n_importance=int(g["n_importance"]))
    sig_ref = O.regulation(scene, rdc, roc, gd, "color", t_rand)
    (0.0005 * sig_ref.abs().sum()).backward()
    rog, rdg = ro.to(DEV).requires_grad_(True), rd.to(DEV).requires_grad_(True)
    sig = r.regulation({}, model, rdg, rog, gd.to(DEV), DEV, "color", t_rand=t_rand)
    (0.0005 * sig.abs().sum()).backward()
    assert T.rel_max(sig, sig_ref) < 1e-4
    assert rog.grad is not None and rdg.grad is not None
    assert T.rel_max(rog.grad, roc.grad) < 2e-3 and T.rel_max(rdg.grad, rdc.grad) < 2e-3
