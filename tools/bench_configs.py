#!/usr/bin/env python
"""Throughput of the other BASELINE.json configurations on one GPU (developer/report probe):
  [2] tracking iteration: 1000 px, fwd+bwd to the camera 7-vector, stage color
  [5a] dense render_img: 680x1200 = 816,000 rays, no-grad, stage color
  [5b] eval_points on the 256^3 mesh-extraction lattice, stage fine
  iMAP*: mapping iteration 5000 rays, 32+12 samples, hidden-256 MLP, density compositing"""
import sys, os, time, types, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench as B
import pointnerf_slam_b200 as P
from oracle import nice_oracle as O

dev = torch.device("cuda", 0)
out = {}

def timeit(fn, iters, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

bound = P.load_bound(B.CFG)
torch.manual_seed(0)
model = P.get_model(B.CFG, nice=True).to(dev); P.attach_bounds(model, bound)
grids = P.grid_init(B.CFG, bound, dev, generator=torch.Generator().manual_seed(1))
slam = types.SimpleNamespace(bound=bound, H=B.H, W=B.W, fx=B.FX, fy=B.FY, cx=B.CX, cy=B.CY, nice=True)
r = P.Renderer(B.CFG, None, slam)
depth, color = [t.to(dev) for t in B.synthetic_frames(1, 100)[0]]
c2w = B.keyframe_poses(0)[0].to(dev)

# ---- [2] tracking
r.freeze_map = True
cam = P.get_tensor_from_camera(c2w).to(dev).requires_grad_(True)
def tracking():
    c = P.get_camera_from_tensor(cam)
    o, d, gd, gc = P.get_samples(100, B.H - 100, 100, B.W - 100, 1000, B.H, B.W, B.FX, B.FY, B.CX, B.CY, c, depth, color, dev)
    dd, vv, cc = r.render_batch_ray(grids, model, d, o, dev, "color", gt_depth=gd)
    m = gd > 0
    loss = torch.where(m, torch.abs(gd - dd) / torch.sqrt(vv.detach() + 1e-10), torch.zeros_like(dd)).sum() + \
        0.5 * (torch.abs(gc - cc) * m[:, None]).sum()
    loss.backward(); cam.grad = None
ms = timeit(tracking, 50)
out["tracking_1000px_fwd_bwd_pose"] = {"ms_per_iter": round(ms, 4), "rays_per_s": round(1000 / ms * 1e3, 1)}
# the same iteration captured once in a CUDA graph (the pose gradient stays in cam.grad between replays)
def tracking_keep_grad():
    c = P.get_camera_from_tensor(cam)
    o, d, gd, gc = P.get_samples(100, B.H - 100, 100, B.W - 100, 1000, B.H, B.W, B.FX, B.FY, B.CX, B.CY, c, depth, color, dev)
    dd, vv, cc = r.render_batch_ray(grids, model, d, o, dev, "color", gt_depth=gd)
    m = gd > 0
    loss = torch.where(m, torch.abs(gd - dd) / torch.sqrt(vv.detach() + 1e-10), torch.zeros_like(dd)).sum() + \
        0.5 * (torch.abs(gc - cc) * m[:, None]).sum()
    cam.grad = None
    loss.backward()
    return loss
gstep = P.graphs.GraphedStep(tracking_keep_grad)
ms = timeit(gstep, 50)
out["tracking_1000px_fwd_bwd_pose_cuda_graph"] = {"ms_per_iter": round(ms, 4), "rays_per_s": round(1000 / ms * 1e3, 1),
                                                   "library_kernels_per_replay": gstep.launches}
gstep.release()
# ... and with the package's fused tracker loss head (losses.tracking_loss, handle_dynamic as in the shipped configs)
def tracking_fused():
    c = P.get_camera_from_tensor(cam)
    o, d, gd, gc = P.get_samples(100, B.H - 100, 100, B.W - 100, 1000, B.H, B.W, B.FX, B.FY, B.CX, B.CY, c, depth, color, dev)
    dd, vv, cc = r.render_batch_ray(grids, model, d, o, dev, "color", gt_depth=gd)
    loss = P.losses.tracking_loss(dd, vv, cc, gd, gc, 0.5, True, True)
    cam.grad = None
    loss.backward()
    return loss
ms = timeit(tracking_fused, 50)
out["tracking_1000px_fused_loss_eager"] = {"ms_per_iter": round(ms, 4), "rays_per_s": round(1000 / ms * 1e3, 1)}
gstep = P.graphs.GraphedStep(tracking_fused)
ms = timeit(gstep, 50)
out["tracking_1000px_fused_loss_cuda_graph"] = {"ms_per_iter": round(ms, 4), "rays_per_s": round(1000 / ms * 1e3, 1),
                                                 "library_kernels_per_replay": gstep.launches}
gstep.release()
r.freeze_map = False

# ---- [5a] dense render
def dense():
    with torch.no_grad():
        r.render_img(grids, model, c2w, dev, "color", gt_depth=depth)
ms = timeit(dense, 3, warm=1)
out["render_img_816k_rays_color"] = {"ms": round(ms, 2), "rays_per_s": round(B.H * B.W / ms * 1e3, 1), "points_per_s": round(B.H * B.W * 48 / ms * 1e3, 1)}

# ---- [5b] 256^3 eval_points, stage fine
lo, hi = bound[:, 0] - 0.05, bound[:, 1] + 0.05
ax = [torch.linspace(float(lo[a]), float(hi[a]), 256) for a in range(3)]
pts = torch.stack(torch.meshgrid(*ax, indexing="ij"), -1).reshape(-1, 3).float().to(dev)
def mesh_query():
    with torch.no_grad():
        r.eval_points(pts, model, grids, "fine", dev)
ms = timeit(mesh_query, 3, warm=1)
out["eval_points_256cubed_fine"] = {"ms": round(ms, 2), "points_per_s": round(pts.shape[0] / ms * 1e3, 1),
                                    "frac_of_hbm_roofline_2076B_per_point": round(pts.shape[0] / ms * 1e3 * 2076 / (B.peaks()[0] * 1e9), 4)}
del pts

# ---- iMAP* mapping iteration
icfg = {"rendering": {"lindisp": False, "perturb": 0.0, "N_samples": 32, "N_surface": 0, "N_importance": 12}, "scale": 1,
        "occupancy": False, "data": {"dim": 3}, "model": {"c_dim": 32, "pos_embedding_method": "fourier"}, "coarse": False,
        "grid_len": B.CFG["grid_len"]}
imodel = P.get_model(icfg, nice=False).to(dev)
islam = types.SimpleNamespace(bound=bound, H=B.H, W=B.W, fx=B.FX, fy=B.FY, cx=B.CX, cy=B.CY, nice=False)
ir = P.Renderer(icfg, None, islam)
o, d, gd, gc = P.get_samples(0, B.H, 0, B.W, 5000, B.H, B.W, B.FX, B.FY, B.CX, B.CY, c2w, depth, color, dev)
def imap():
    dd, vv, cc = ir.render_batch_ray({}, imodel, d, o, dev, "color", gt_depth=gd)
    sig = ir.regulation({}, imodel, d, o, gd, dev, "color")
    m = gd > 0
    loss = torch.where(m, torch.abs(gd - dd), torch.zeros_like(dd)).sum() + 0.05 * torch.abs(gc - cc).sum() + 0.0005 * sig.abs().sum()
    loss.backward(); imodel.zero_grad(set_to_none=True)
ms = timeit(imap, 10)
out["imap_mapping_5000rays_fwd_bwd"] = {"ms_per_iter": round(ms, 3), "rays_per_s": round(5000 / ms * 1e3, 1)}
print(json.dumps(out, indent=1))
