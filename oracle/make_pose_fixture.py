#!/usr/bin/env python
"""Extract real Replica-room0 camera poses from the reference's shipped checkpoint
(output_imap/Replica/room0/ckpts/01999.tar: gt_c2w_list, translations stored x scale 0.1)
into tests/golden/room0_poses.npz.  Build-container only (needs /root/reference)."""
import os
import numpy as np
import torch

REF = "/root/reference/output_imap/Replica/room0/ckpts/01999.tar"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "room0_poses.npz")
ck = torch.load(REF, map_location="cpu", weights_only=False)
c2w = ck["gt_c2w_list"][:400:10].clone().float()          # 40 poses, every 10th frame
c2w[:, :3, 3] /= 0.1                                       # undo configs/imap.yaml scale
np.savez_compressed(OUT, c2w=c2w.numpy(), frames=np.arange(0, 400, 10))
print(OUT, c2w.shape, c2w[:, :3, 3].min(0).values, c2w[:, :3, 3].max(0).values)
