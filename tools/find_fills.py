#!/usr/bin/env python
"""Developer probe: which Python frames launch aten::fill_/zero_ during one keyframe's sampling backward."""
import sys, os, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench as B
import pointnerf_slam_b200 as P
from torch.profiler import profile, ProfilerActivity

dev = torch.device("cuda", 0)
depth, color = [t.to(dev) for t in B.synthetic_frames(1, 100)[0]]
cam = P.get_tensor_from_camera(B.keyframe_poses(0)[1]).to(dev).requires_grad_(True)
def it():
    c2w = P.get_camera_from_tensor(cam)
    o, d, gd, gc = P.get_samples(0, B.H, 0, B.W, 1000, B.H, B.W, B.FX, B.FY, B.CX, B.CY, c2w, depth, color, dev)
    ro = torch.cat([o, o]); rd = torch.cat([d, d])
    ((ro * rd).sum()).backward()
for _ in range(3): it()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], with_stack=True) as prof:
    it()
for e in prof.events():
    if e.name in ("aten::fill_", "aten::zero_", "aten::zeros", "aten::zeros_like", "aten::new_zeros"):
        st = [s for s in (e.stack or []) if "site-packages/torch" not in s][:4]
        print(e.name, e.input_shapes if hasattr(e, "input_shapes") else "", "|", " <- ".join(st) or (e.stack or [])[:3])
