import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests import helpers as T
from oracle import nice_oracle as O
g = T.load_nice(); DEV = "cuda:0"
model, grids, renderer = T.cuda_nice(g, DEV)
stage = sys.argv[1] if len(sys.argv) > 1 else "middle"
d, v, c = renderer.render_batch_ray(grids, model, g["rays_d"].to(DEV), g["rays_o"].to(DEV), DEV, stage, gt_depth=g["gt_depth"].to(DEV))
loss = O.mapping_loss(d, c, g["gt_depth"].to(DEV), g["gt_color"].to(DEV), stage); loss.backward()
for name, p in model.named_parameters():
    key = f"{stage}/map/gradsd/{name}"
    if key in g:
        ref = g[key]; got = p.grad.cpu()
        print(f"{name:45s} rel {T.rel_max(got, ref):9.2e}  |got| {got.abs().max():9.3e} |ref| {ref.abs().max():9.3e}")
