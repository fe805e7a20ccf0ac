#!/usr/bin/env python
"""Where does the host time of one mapping iteration go?  (developer probe)"""
import cProfile, pstats, sys, os, time, io
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench as B

class A: steps = 30; warmup = 3; light = True; gpus = 1
# re-create the bench's step by monkeypatching run_ours's inner pieces is awkward; emulate instead
import types
import pointnerf_slam_b200 as P
dev = torch.device("cuda", 0)
bound = P.load_bound(B.CFG)
torch.manual_seed(0)
model = P.get_model(B.CFG, nice=True).to(dev); P.attach_bounds(model, bound)
grids = P.grid_init(B.CFG, bound, dev)
slam = types.SimpleNamespace(bound=bound, H=B.H, W=B.W, fx=B.FX, fy=B.FY, cx=B.CX, cy=B.CY, nice=True)
r = P.Renderer(B.CFG, None, slam)
for k in ("grid_middle", "grid_fine", "grid_color"): grids[k].requires_grad_(True)
for n_, p in model.named_parameters(): p.requires_grad_(n_.startswith("color_decoder."))
frames = [(d.to(dev), c.to(dev)) for d, c in B.synthetic_frames(5, 100)]
poses = B.keyframe_poses(0).to(dev)
cams = [P.get_tensor_from_camera(poses[k]).to(dev).requires_grad_(k > 0) for k in range(5)]
trained = [grids[k] for k in ("grid_middle", "grid_fine", "grid_color")] + [p for p in model.parameters() if p.requires_grad] + cams[1:]
def step():
    ro, rd, gd, gc = [], [], [], []
    for k in range(5):
        c2w = P.get_camera_from_tensor(cams[k])
        o, d, dd, cc = P.get_samples(0, B.H, 0, B.W, 1000, B.H, B.W, B.FX, B.FY, B.CX, B.CY, c2w, frames[k][0], frames[k][1], dev)
        ro.append(o); rd.append(d); gd.append(dd); gc.append(cc)
    ro, rd, gd, gc = torch.cat(ro), torch.cat(rd), torch.cat(gd), torch.cat(gc)
    depth, var, color = r.render_batch_ray(grids, model, rd, ro, dev, "color", gt_depth=gd)
    m = gd > 0
    loss = torch.where(m, torch.abs(gd - depth), torch.zeros_like(depth)).sum() + 0.2 * torch.abs(gc - color).sum()
    loss.backward()
    for t in trained: t.grad = None
for _ in range(5): step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(30): step()
t_submit = time.perf_counter() - t0
torch.cuda.synchronize()
t_total = time.perf_counter() - t0
print(f"host submit {t_submit/30*1e3:.3f} ms/step, with final sync {t_total/30*1e3:.3f} ms/step")
pr = cProfile.Profile(); pr.enable()
for _ in range(30): step()
pr.disable(); torch.cuda.synchronize()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(28); print(s.getvalue()[:6000])
