"""The CPU oracle replays the reference's golden vectors (minted by
oracle/pin_against_reference.py from the reference's own code).  In the build
container the match is bit-exact; a small tolerance covers other CPUs' sgemm."""
import torch

from oracle import nice_oracle as O
from tests import helpers as T

TOL = dict(rtol=1e-5, atol=1e-6)


def test_bound_and_grid_shapes_room0():
    b = O.scene_bound([[-2.9, 8.9], [-3.2, 5.5], [-3.5, 3.3]], 1, 0.32)
    assert b[:, 1].tolist() == [8.94000015258789, 5.7600000381469725, 3.5399999618530273]
    sh = O.grid_shapes(b, {"coarse": 2, "middle": 0.32, "fine": 0.16, "color": 0.16})
    assert sh == {"grid_coarse": (7, 8, 11), "grid_middle": (21, 28, 37), "grid_fine": (43, 56, 74),
                  "grid_color": (43, 56, 74)}


def test_eval_points_all_stages():
    g = T.load("nice_eval_points.npz")
    scene = T.oracle_scene(g)
    for stage in O.STAGES:
        with torch.no_grad():
            out = O.eval_points(scene, g["points"].clone(), stage)
            out32 = O.eval_points(scene, g["points"].float(), stage)
        torch.testing.assert_close(out, g[f"raw_{stage}"], **TOL)
        torch.testing.assert_close(out32, g[f"raw32_{stage}"], **TOL)
        # the out-of-bound override is index-exact
        assert torch.equal(out[:, 3] == 100, g[f"raw_{stage}"][:, 3] == 100)


def test_render_forward_backward():
    g = T.load_nice()
    H, W, fx, fy, cx, cy = [float(v) for v in g["intr"]]
    H0, H1, W0, W1 = [int(v) for v in g["crop"]]
    i, j = O.pixel_lattice(H0, H1, W0, W1)
    idx = g["indices"]
    for stage in ("middle", "fine", "color"):
        grids = {k: g[k].clone().requires_grad_(True) for k in O.GRID_KEYS}
        sd = {k: v.clone().requires_grad_(True) for k, v in T.state_dict(g).items()}
        scene = T.oracle_scene(g, grids, sd)
        c2w = O.camera_from_tensor(g["cam"])
        ro, rd = O.rays_from_pixels(i[idx], j[idx], c2w, fx, fy, cx, cy)
        assert torch.equal(ro, g["rays_o"]) and torch.equal(rd, g["rays_d"])
        d, v, c = O.render_batch_ray(scene, rd, ro, stage, g["gt_depth"])
        torch.testing.assert_close(d, g[f"{stage}/map/depth"], **TOL)
        torch.testing.assert_close(v, g[f"{stage}/map/var"], **TOL)
        torch.testing.assert_close(c, g[f"{stage}/map/color"], **TOL)
        O.mapping_loss(d, c, g["gt_depth"], g["gt_color"], stage).backward()
        for k in O.GRID_KEYS:
            key = f"{stage}/map/grad_{k}"
            if key in g:
                torch.testing.assert_close(grids[k].grad, g[key], rtol=1e-4, atol=1e-6)
    # tracking gradient to the camera 7-vector
    cam = g["cam"].clone().requires_grad_(True)
    scene = T.oracle_scene(g)
    ro, rd = O.rays_from_pixels(i[idx], j[idx], O.camera_from_tensor(cam), fx, fy, cx, cy)
    d, v, c = O.render_batch_ray(scene, rd, ro, "color", g["gt_depth"])
    O.tracking_loss(d, v, c, g["gt_depth"], g["gt_color"]).backward()
    torch.testing.assert_close(cam.grad, g["color/track/grad_cam"], rtol=1e-4, atol=1e-4)


def test_imap_with_shipped_weights():
    g = T.load("imap_render.npz")
    sd = {k: v.clone().requires_grad_(True) for k, v in T.state_dict(g).items()}
    scene = O.Scene(sd, {}, g["bound"], nice=False, occupancy=False, n_samples=int(g["n_samples"]),
                    n_surface=int(g["n_surface"]), n_importance=int(g["n_importance"]))
    d, v, c = O.render_batch_ray(scene, g["rays_d"], g["rays_o"], "color", g["gt_depth"])
    torch.testing.assert_close(d, g["depth"], **TOL)
    torch.testing.assert_close(c, g["color"], rtol=1e-4, atol=1e-5)
    sig = O.regulation(scene, g["rays_d"], g["rays_o"], g["gt_depth"], "color", t_rand=g["reg_t_rand"])
    torch.testing.assert_close(sig, g["reg_sigma"], rtol=1e-4, atol=1e-4)
    loss = O.mapping_loss(d, c, g["gt_depth"], g["gt_color"], "color", 0.05, nice=False) + 0.0005 * sig.abs().sum()
    torch.testing.assert_close(loss, g["loss"], rtol=1e-6, atol=1e-6)
    loss.backward()
    for k, v in sd.items():
        gk = g["gradsd/" + k]
        assert T.rel_max(v.grad, gk) < 1e-3, k


def test_fixed_embed_order_is_an_admissible_order_of_the_reference_matmul():
    """EMBED_ORDER='fma' (used by the full-size GPU comparisons) evaluates p @ B as one fused x-y-z chain: per element it may
    differ from the box's BLAS by the last bit of the argument only, and its gradients are those of the matmul."""
    g = torch.Generator().manual_seed(0)
    p = (torch.rand(4000, 3, generator=g) * 10 - 4).requires_grad_(True)
    B = (torch.randn(3, O.EMBED, generator=g) * 25).requires_grad_(True)
    ref = O.fourier_embed(p, B)
    gw = torch.randn(ref.shape, generator=g)
    (ref * gw).sum().backward()
    gp, gB = p.grad.clone(), B.grad.clone()
    p.grad = B.grad = None
    assert O.EMBED_ORDER is None
    with O.fixed_embed_order():
        assert O.EMBED_ORDER == "fma"
        out = O.fourier_embed(p, B)
    assert O.EMBED_ORDER is None
    arg = (p.detach().double() @ B.detach().double())
    # rounding of the partial sums: relative to the magnitude of the terms, not of the (possibly cancelled) result
    ulp = torch.finfo(torch.float32).eps * (p.detach().double().abs() @ B.detach().double().abs()).clamp_min(1.0)
    # |sin a - sin b| <= |a - b|: both float32 evaluations lie within 2 ulps (of the terms) of the exact argument
    assert bool(((out.detach().double() - torch.sin(arg)).abs() <= 2 * ulp + 1e-6).all())
    assert bool(((ref.detach().double() - torch.sin(arg)).abs() <= 2 * ulp + 1e-6).all())
    (out * gw).sum().backward()
    assert torch.allclose(p.grad, gp, rtol=1e-3, atol=1e-2 * gp.abs().max().item())
    assert torch.allclose(B.grad, gB, rtol=1e-3, atol=1e-2 * gB.abs().max().item())
