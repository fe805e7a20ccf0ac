#!/usr/bin/env python
"""Developer timing probe (not the bench contract): room0-sized scene, N rays."""
import sys, os, time, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pointnerf_slam_b200 as P
from oracle import nice_oracle as O

def main(n_rays=5000, iters=10):
    dev = "cuda:0"
    cfg = {"rendering": {"lindisp": False, "perturb": 0.0, "N_samples": 32, "N_surface": 16, "N_importance": 0},
           "scale": 1, "occupancy": True, "coarse": True, "data": {"dim": 3},
           "grid_len": {"coarse": 2, "middle": 0.32, "fine": 0.16, "color": 0.16, "bound_divisible": 0.32},
           "model": {"c_dim": 32, "coarse_bound_enlarge": 2, "pos_embedding_method": "fourier"},
           "mapping": {"bound": [[-2.9, 8.9], [-3.2, 5.5], [-3.5, 3.3]]}}
    bound = P.load_bound(cfg)
    torch.manual_seed(0)
    model = P.get_model(cfg, nice=True).to(dev)
    P.attach_bounds(model, bound)
    grids = P.grid_init(cfg, bound, dev)
    H, W, fx, fy, cx, cy = 680, 1200, 600.0, 600.0, 599.5, 339.5
    slam = types.SimpleNamespace(bound=bound, H=H, W=W, fx=fx, fy=fy, cx=cx, cy=cy, nice=True)
    r = P.Renderer(cfg, None, slam)
    depth = (1 + 2 * torch.rand(H, W)).to(dev); depth[torch.rand(H, W, device=dev) < 0.02] = 0
    color = torch.rand(H, W, 3).to(dev)
    c2w = torch.eye(4)[:3].clone(); c2w[:, 3] = torch.tensor([2.0, 1.0, 0.0]); c2w = c2w.to(dev)
    ro, rd, gd, gc = P.get_samples(0, H, 0, W, n_rays, H, W, fx, fy, cx, cy, c2w, depth, color, dev)

    def timeit(fn, label):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(iters): fn()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        print(f"{label:40s} {ms:8.3f} ms  {n_rays/ms*1e3/1e6:8.3f} Mrays/s", flush=True)

    for stage in ("middle", "fine", "color"):
        def fwd():
            with torch.no_grad():
                r.render_batch_ray(grids, model, rd, ro, dev, stage, gt_depth=gd)
        timeit(fwd, f"fwd no-grad [{stage}]")
    for k in grids: grids[k].requires_grad_(True)
    for p in model.middle_decoder.parameters(): p.requires_grad_(False)
    for p in model.fine_decoder.parameters(): p.requires_grad_(False)
    for p in model.coarse_decoder.parameters(): p.requires_grad_(False)
    def mapping():
        d, v, c = r.render_batch_ray(grids, model, rd, ro, dev, "color", gt_depth=gd)
        loss = O.mapping_loss(d, c, gd, gc, "color")
        loss.backward()
        for k in grids: grids[k].grad = None
        model.zero_grad(set_to_none=True)
    timeit(mapping, "fwd+bwd mapping [color] grids+color dec")
    for k in grids: grids[k].requires_grad_(False)
    r.freeze_map = True
    cam = P.get_tensor_from_camera(c2w).to(dev).requires_grad_(True)
    idx = torch.randint(H * W, (n_rays,), device=dev)
    def tracking():
        c2 = P.get_camera_from_tensor(cam)
        o, d_, gdd, gcc = P.get_samples(0, H, 0, W, n_rays, H, W, fx, fy, cx, cy, c2, depth, color, dev, indices=idx)
        d, v, c = r.render_batch_ray(grids, model, d_, o, dev, "color", gt_depth=gdd)
        loss = O.tracking_loss(d, v, c, gdd, gcc)
        loss.backward(); cam.grad = None
    timeit(tracking, "fwd+bwd tracking [color] -> pose")

if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 5000)
