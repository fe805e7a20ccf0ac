#!/usr/bin/env python
"""Benchmark of the differentiable ray-rendering hot path (BASELINE.json metric:
rays/sec forward+backward ``render_batch_ray``; % of the HBM roofline; vs the CPU
reference).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--config mapping|tracking|dense|mesh256|imap] [--scaling weak|strong]

Default workload (BASELINE.json configs[2]) -- one NICE-SLAM
mapping iteration, stage ``color``, on a synthetic Replica-room0-shaped scene:
5 keyframes x 1000 pixels = 5000 rays x 48 samples; pixel sampling and ray
generation from the (bundle-adjusted) camera tensors, ``render_batch_ray``, the
Mapper loss (src/Mapper.py:628-646), the backward into the middle / fine / colour
grids, the colour decoder and 4 camera poses (src/Mapper.py:402-477 upstream
configuration: fix_fine, BA) and the Mapper's optimiser step: Adam with the
stage learning rates on the frustum-selected voxels, the colour decoder and the
cameras (src/Mapper.py:129-200, 482-536, 657-674) -- in BOTH arms.

N > 1 is launched by torchrun (one rank per GPU, NCCL).  ``--scaling weak`` (default):
5000 rays per rank; ``--scaling strong``: the 5000 rays of ONE batch are sharded.
Either way the batch-global depth maximum is shared by an all-reduce(MAX) and the
gradients are summed over the ranks inside the timed step (``PN_BENCH_EXCHANGE``:
sparse | sparse_p2p | overlap | arena | dense).

The other BASELINE.json configurations print the same JSON schema:
``tracking`` (configs[1]: 1000 px to the camera pose), ``dense`` (configs[4]a: 816,000-ray
``render_img``; image rows shard over the ranks, outputs all-gathered), ``mesh256``
(configs[4]b: ``eval_points`` on the 256^3 lattice, slabs shard over the ranks) and
``imap`` (the iMAP* MLP mapping iteration this fork runs live).

``--impl reference`` times the reference's own torch implementation of the same
step on the host cores (the oracle port, pinned bit-equal against the reference;
the reference is pure Python and cannot travel to the GPU box).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

CFG = {"rendering": {"lindisp": False, "perturb": 0.0, "N_samples": 32, "N_surface": 16, "N_importance": 0},
       "scale": 1, "occupancy": True, "coarse": True, "data": {"dim": 3},
       "grid_len": {"coarse": 2, "middle": 0.32, "fine": 0.16, "color": 0.16, "bound_divisible": 0.32},
       "model": {"c_dim": 32, "coarse_bound_enlarge": 2, "pos_embedding_method": "fourier"},
       "mapping": {"bound": [[-2.9, 8.9], [-3.2, 5.5], [-3.5, 3.3]]}}
ICFG = {"rendering": {"lindisp": False, "perturb": 0.0, "N_samples": 32, "N_surface": 0, "N_importance": 12}, "scale": 1,
        "occupancy": False, "data": {"dim": 3}, "model": {"c_dim": 32, "pos_embedding_method": "fourier"}, "coarse": False,
        "grid_len": CFG["grid_len"]}
STAGE_LR = {  # configs/nice_slam.yaml:71-95
    "coarse": {"decoders_lr": 0.0, "coarse_lr": 0.001, "middle_lr": 0.0, "fine_lr": 0.0, "color_lr": 0.0},
    "middle": {"decoders_lr": 0.0, "coarse_lr": 0.0, "middle_lr": 0.1, "fine_lr": 0.0, "color_lr": 0.0},
    "fine": {"decoders_lr": 0.0, "coarse_lr": 0.0, "middle_lr": 0.005, "fine_lr": 0.005, "color_lr": 0.0},
    "color": {"decoders_lr": 0.005, "coarse_lr": 0.0, "middle_lr": 0.005, "fine_lr": 0.005, "color_lr": 0.005}}
H, W, FX, FY, CX, CY = 680, 1200, 600.0, 600.0, 599.5, 339.5
N_KEYFRAMES, PIX_PER_KF, S = 5, 1000, 48
W_COLOR = 0.2
# algorithmic bytes per unit: 1024 B per sample per grid gathered; backward re-gathers and RMWs
BYTES_PER_RAY_STEP = 589_872        # fwd+bwd mapping, stage colour
BYTES_PER_RAY_TRACK = 294_960       # fwd+bwd tracking (re-gather, no scatter)
BYTES_PER_RAY_DENSE = 147_504       # fwd only, stage colour
BYTES_PER_POINT_MESH = 2_076        # eval_points stage fine
BYTES_PER_SAMPLE_GATHER = 1024
# k-NN aggregation (config 4, builder-defined): per sample, forward = 8 x (16 B position record + 128 B feature row) + 128 B blended
# feature + 64 B index / distance lists; backward = 128 B feature gradient + 64 B lists + 8 x (128 B row + 12 B position) for the
# point gradient + 8 x 256 B read-modify-write of the gradient rows + 12 B point gradient
BYTES_PER_SAMPLE_KNN_FWD = 8 * (16 + 128) + 128 + 64
BYTES_PER_SAMPLE_KNN_BWD = 128 + 64 + 8 * (128 + 12) + 8 * 256 + 12
BYTES_PER_RAY_KNN = 48 * (BYTES_PER_SAMPLE_KNN_FWD + BYTES_PER_SAMPLE_KNN_BWD)
KNN_RADIUS = 0.16
METRIC = "rays/sec fwd+bwd render_batch_ray (NICE mapping iteration, stage color)"


ROOFLINE_NOTE = ("decoder layers run on tcgen05 (kind::tf32, 3xTF32 split, FP32 accumulate in TMEM); the decoder kernels are bound by "
                 "the shared-memory data pipe (tensor-core operand reads + operand stores: 0.5-0.6 wavefronts per cycle, 0.96 in the "
                 "weight-gradient kernel) and by the latency of the 5-9 dependent MMA round trips per tile, not by HBM; the grids "
                 "(46 MiB) are L2-resident, so DRAM traffic is far below the algorithmic bytes; HBM roofline is the BASELINE.md denominator")
ROOFLINE_NOTE_KNN = ("integer/float32 search + gather: no GEMM, SIMT kernels only; the candidate scan of the 27 neighbouring cells "
                     "(~150 position records per sample, L1/L2 hits) is not part of the algorithmic bytes; kernel_frac_of_hbm of knn_bwd "
                     "exceeds 1 because its algorithmic bytes (8 row read-modify-writes + 8 row re-reads per sample) mostly hit L2: "
                     "neighbouring samples share neighbours (measured DRAM traffic 70 MB per launch, profiles/r2_ncu_knn_summary.txt)")


def workload_name(scaling="weak"):
    return ("nice_mapping_iter_color: 5 keyframes x 1000 px, 48 samples/ray, room0 grids, BA, frustum-masked Adam step"
            + (" (one 5000-ray batch sharded over the ranks)" if scaling == "strong" else ""))


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p))["hbm_gbs"], "measured"
    return 6650.0, "fallback"


def synthetic_frames(n, seed):
    g = torch.Generator().manual_seed(seed)
    frames = []
    for _ in range(n):
        depth = 1.0 + 2.0 * torch.rand(H, W, generator=g)
        depth[torch.rand(H, W, generator=g) < 0.02] = 0.0
        frames.append((depth, torch.rand(H, W, 3, generator=g)))
    return frames


def keyframe_poses(rank):
    z = np.load(os.path.join(ROOT, "tests", "golden", "room0_poses.npz"))
    c2w = torch.from_numpy(z["c2w"]).float()
    sel = [(rank * N_KEYFRAMES + k) * 5 % c2w.shape[0] for k in range(N_KEYFRAMES)]
    return c2w[sel]


_NVML_POLL = r"""
import sys, time
import pynvml as N
N.nvmlInit()
h = N.nvmlDeviceGetHandleByIndex(int(sys.argv[1]))
mx = N.nvmlDeviceGetMaxClockInfo(h, N.NVML_CLOCK_SM)
while True:
    print(N.nvmlDeviceGetClockInfo(h, N.NVML_CLOCK_SM), mx, int(N.nvmlDeviceGetCurrentClocksThrottleReasons(h)),
          N.nvmlDeviceGetPowerUsage(h) // 1000, flush=True)
    time.sleep(0.001)
"""


class ClockSampler(threading.Thread):
    """SM clock / throttle reasons DURING the timed region.  A helper PROCESS polls NVML every millisecond (a thread
    of this process would starve behind the launch loop's GIL; `nvidia-smi -lms` cannot go below 100 ms while a default
    timed region lasts ~30 ms); `nvidia-smi -lms 100` is the fallback when pynvml is missing."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    BITS = {"sw_power_cap": 0x4, "hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.proc = index, None
        self.sm, self.mx, self.pw, self.reasons, self.source = [], [], [], set(), None
        self.first = threading.Event()

    def _read_nvml(self):
        self.proc = subprocess.Popen([sys.executable, "-c", _NVML_POLL, str(self.index)], stdout=subprocess.PIPE,
                                     stderr=subprocess.DEVNULL, text=True)
        for line in self.proc.stdout:
            f = line.split()
            if len(f) != 4:
                continue
            self.source = "nvml"
            self.sm.append(float(f[0])); self.mx.append(float(f[1])); self.pw.append(float(f[3]))
            for name, bit in self.BITS.items():
                if int(f[2]) & bit:
                    self.reasons.add(name)
            self.first.set()

    def _read_smi(self):
        self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                      "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        for line in self.proc.stdout:
            r = [x.strip() for x in line.split(",")]
            if len(r) >= 9 and r[1].replace(".", "").isdigit() and r[2].replace(".", "").isdigit():
                self.source = "nvidia-smi"
                self.sm.append(float(r[1])); self.mx.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        self.reasons.add(name)
                self.first.set()

    def run(self):
        try:
            self._read_nvml()
        except Exception:
            pass
        if not self.sm and not self._stopping:
            try:
                self._read_smi()
            except Exception:
                pass

    _stopping = False

    def begin(self):
        """Start polling; returns once the first sample has arrived (or after 3 s); samples taken before are dropped."""
        self.start()
        self.first.wait(3.0)
        del self.sm[:-1]
        del self.pw[:-1]

    def stop(self):
        self._stopping = True
        if self.proc is not None:
            self.proc.terminate()
        self.join(timeout=2)
        return {"sm_mhz": statistics.median(self.sm) if self.sm else None, "sm_max_mhz": max(self.mx) if self.mx else None,
                "reasons": sorted(self.reasons), "samples": len(self.sm), "source": self.source,
                "power_w_max": max(self.pw) if self.pw else None}


# ----------------------------------------------------------------------------
# ours
# ----------------------------------------------------------------------------
def build_scene(dev, nice=True):
    import pointnerf_slam_b200 as P
    cfg = CFG if nice else ICFG
    bound = P.load_bound(CFG)
    torch.manual_seed(0)
    model = P.get_model(cfg, nice=nice).to(dev)
    grids = {}
    if nice:
        P.attach_bounds(model, bound)
        grids = P.grid_init(CFG, bound, dev, generator=torch.Generator().manual_seed(1))
    slam = types.SimpleNamespace(bound=bound, H=H, W=W, fx=FX, fy=FY, cx=CX, cy=CY, nice=nice)
    return P, bound, model, grids, P.Renderer(cfg, None, slam)


def build_mapping(dev, rank, world, pix_per_kf=PIX_PER_KF, exchange="none", with_optimizer=True, arena=True, gen_seed=1234,
                  shared_cameras=False):
    """The bench's mapping workload: scene, synthetic keyframes, camera tensors, MappingIteration (+ StageOptimizer over
    the frustum-selected voxels).  Also used by tests/test_gpu_mapping_iteration.py: ONE definition of the step."""
    P, bound, model, grids, renderer = build_scene(dev, True)
    from pointnerf_slam_b200 import engine as E
    from pointnerf_slam_b200 import mapper as PM
    from pointnerf_slam_b200.mapping import MappingIteration
    # mapping configuration: grids + colour decoder trained, fine/middle/coarse decoders fixed, BA on 4 of 5 poses
    for k in ("grid_middle", "grid_fine", "grid_color"):
        grids[k].requires_grad_(True)
    for name, p in model.named_parameters():
        p.requires_grad_(name.startswith("color_decoder."))
    frames_host = synthetic_frames(N_KEYFRAMES, 100 + rank)
    frames = [(d.to(dev), c.to(dev)) for d, c in frames_host]
    poses = keyframe_poses(rank).to(dev)
    cams = [P.get_tensor_from_camera(poses[k]).to(dev).requires_grad_(k > 0) for k in range(N_KEYFRAMES)]
    gen = torch.Generator(device=dev).manual_seed(gen_seed + rank)
    # every gradient sink of the backward comes from one arena -> one memset per iteration instead of ~20 small fills
    n_arena = sum(g.numel() for k, g in grids.items() if k != "grid_coarse") + N_KEYFRAMES * pix_per_kf * S * 3 + 262144
    ar = E.GradArena(n_arena, dev) if arena else None
    opt = None
    masks = None
    if with_optimizer:
        # frustum feature selection from the current (newest) keyframe, Mapper.py:413-431 (frustum_feature_selection: True)
        masks = {k: PM.frustum_voxel_mask(poses[N_KEYFRAMES - 1], k, grids[k].shape[2:], frames[N_KEYFRAMES - 1][0], bound,
                                          H, W, FX, FY, CX, CY) for k in ("grid_middle", "grid_fine", "grid_color")}
        opt = PM.StageOptimizer({k: grids[k] for k in masks}, [p for p in model.parameters() if p.requires_grad], cams[1:],
                                masks=masks, stage_lr=STAGE_LR, lr_factor=1.0, BA_cam_lr=0.001)
        opt.set_stage("color")
    it = MappingIteration(renderer, model, grids, frames, cams, H, W, FX, FY, CX, CY, pix_per_kf, "color", W_COLOR, generator=gen,
                          arena=ar, exchange=exchange, optimizer=opt, shared_cameras=shared_cameras, world=world)
    return types.SimpleNamespace(P=P, bound=bound, model=model, grids=grids, renderer=renderer, frames=frames,
                                 frames_host=frames_host, poses=poses, cams=cams, gen=gen, arena=ar, optimizer=opt, masks=masks,
                                 iteration=it)


def run_ours(args):
    import pointnerf_slam_b200 as P
    from pointnerf_slam_b200 import _lib as L
    from pointnerf_slam_b200 import dist as D
    from pointnerf_slam_b200 import engine as E
    import torch.distributed as dist

    rank, world, local = D.init_from_env()
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    hbm, which = peaks()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2
    cfgname = args.config
    info = {}           # per-config description that goes into the JSON line
    cleanup = []

    # ---------------------------------------------------------------- workloads
    if cfgname == "mapping":
        strong = args.scaling == "strong"
        pix = PIX_PER_KF // world if strong else PIX_PER_KF
        exchange = os.environ.get("PN_BENCH_EXCHANGE", "sparse") if world > 1 else "none"
        # weak: every rank maps its own 5 keyframes (pose gradients stay local); strong: the rays of ONE 5-keyframe batch shard
        w = build_mapping(dev, rank if not strong else 0, world, pix, exchange, with_optimizer=not args.no_optimizer,
                          shared_cameras=strong)
        if strong:      # every rank holds the same keyframes; rank r draws its own pixels of them
            w.gen.manual_seed(1234 + rank)
        it = w.iteration
        trained = it.trained()
        pinned = [w.frames_host[0][0].pin_memory(), w.frames_host[0][1].pin_memory(), w.poses.cpu().pin_memory()]
        dev_in = [w.frames[0][0], w.frames[0][1], torch.empty_like(w.poses)]
        cams0 = [c.detach().clone() for c in w.cams]

        def step_body():
            # every timed iteration starts from the keyframes' poses: thousands of back-to-back Adam steps on the same five
            # keyframes (the sustained pass) would walk the bundle-adjusted cameras off the scene, and with them the set of
            # voxel rows an iteration touches (the sparse exchange's row capacity is sized for the real geometry)
            with torch.no_grad():
                torch._foreach_copy_(w.cams, cams0)
            return it()
        units_rank, unit, bytes_per_unit = N_KEYFRAMES * pix, "rays/s", BYTES_PER_RAY_STEP
        metric = METRIC
        n_samples = units_rank * S
        alg = {"grid_mlp_fwd:color": n_samples * 1024, "grid_mlp_fwd:fine": n_samples * 2048, "grid_mlp_fwd:middle": n_samples * 1024,
               "grid_mlp_bwd:color": n_samples * 3072, "grid_mlp_bwd:fine": n_samples * 3072, "grid_mlp_bwd:middle": n_samples * 3072,
               # k_wgrad_tc32 reads its operands once: h0..h4, emb, c, GH, the embedding-argument gradient, the output
               # gradient and the point (160+96+32+160+96+4+3 floats per sample) + 20 B of ReLU masks
               "grid_mlp_wgrad:color": n_samples * ((160 + 96 + 32 + 160 + 96 + 4 + 3) * 4 + 20)}
        info = {"workload": workload_name(args.scaling), "rays_per_step_per_gpu": units_rank, "samples_per_ray": S,
                "grids": {k: list(v.shape) for k, v in w.grids.items()},
                "parallelism": (f"ray-shard dp{world}" if strong else f"keyframe-shard dp{world}"),
                "optimizer_in_step": not args.no_optimizer, "gradient_exchange": exchange,
                "poses": "reset to the keyframes' poses at the start of every iteration (one multi-tensor copy inside the step)",
                "frustum_voxels": {k: int(m.sum()) for k, m in (w.masks or {}).items()}}

        def after_step():
            for t in trained:
                t.grad = None
        graph_gens = [w.gen]
        result_of = lambda out: out
        cleanup.append(lambda: setattr(E, "GRAD_ARENA", None))
    elif cfgname == "tracking":
        P_, bound, model, grids, renderer = build_scene(dev, True)
        renderer.freeze_map = True
        fh = synthetic_frames(1, 100 + rank)[0]
        depth, color = fh[0].to(dev), fh[1].to(dev)
        pose = keyframe_poses(rank)[0].to(dev)
        cam = P.get_tensor_from_camera(pose).to(dev).requires_grad_(True)
        gen = torch.Generator(device=dev).manual_seed(77 + rank)
        Hc, Wc = H - 200, W - 200

        from pointnerf_slam_b200.mapper import StageOptimizer
        from pointnerf_slam_b200.tracking import TrackingIteration
        cam0 = cam.detach().clone()
        opt = StageOptimizer({}, [], [cam])
        opt.set_lrs([0.0, 0.0, 0.0, 0.0, 0.0, 0.001])          # tracking cam_lr, configs/Replica/replica.yaml
        track = TrackingIteration(renderer, model, grids, depth, color, cam, H, W, FX, FY, CX, CY, 1000, 100, 100, 0.5, True, True,
                                  generator=gen, optimizer=opt)

        def step_body():
            cam.data.copy_(cam0)       # every timed iteration starts from the same pose (thousands of replays would drift off the scene)
            return track()
        pinned = [fh[0].pin_memory(), fh[1].pin_memory()]
        dev_in = [depth, color]
        units_rank, unit, bytes_per_unit = 1000, "rays/s", BYTES_PER_RAY_TRACK
        metric = "rays/sec fwd+bwd render_batch_ray (NICE tracking iteration to the camera pose, stage color)"
        n_samples = 1000 * S
        alg = {f"grid_mlp_fwd:{k}": n_samples * (2048 if k == "fine" else 1024) for k in ("color", "fine", "middle")}
        alg.update({f"grid_mlp_bwd:{k}": n_samples * (2048 if k == "fine" else 1024) for k in ("color", "fine", "middle")})
        info = {"workload": "nice_tracking_iter: 1000 px of the 480x1000 crop, 48 samples/ray, fwd+bwd to the camera 7-vector, "
                            "fused tracker loss (handle_dynamic), Adam step on the camera; replicas only (N ranks = N independent trackers)",
                "optimizer_in_step": True,
                "rays_per_step_per_gpu": 1000, "samples_per_ray": S, "parallelism": f"replicas x{world}"}
        after_step = lambda: None
        graph_gens = [gen]
        result_of = lambda out: out
    elif cfgname == "imap":
        P_, bound, model, grids, renderer = build_scene(dev, False)
        fh = synthetic_frames(1, 100 + rank)[0]
        depth, color = fh[0].to(dev), fh[1].to(dev)
        pose = keyframe_poses(rank)[0].to(dev)
        gen = torch.Generator(device=dev).manual_seed(78 + rank)
        t_rand = torch.rand(5000, 32, device=dev)

        def step_body():
            idx = torch.randint(H * W, (5000,), device=dev, generator=gen)
            o, d, gd, gc = P.get_samples(0, H, 0, W, 5000, H, W, FX, FY, CX, CY, pose, depth, color, dev, indices=idx)
            dd, vv, cc = renderer.render_batch_ray({}, model, d, o, dev, "color", gt_depth=gd)
            sig = renderer.regulation({}, model, d, o, gd, dev, "color", t_rand=t_rand)
            loss = P.losses.mapping_loss(dd, cc, gd, gc, "color", 0.05, nice=False) + 0.0005 * sig.abs().sum()
            model.zero_grad(set_to_none=True)
            loss.backward()
            return loss
        pinned = [fh[0].pin_memory(), fh[1].pin_memory()]
        dev_in = [depth, color]
        units_rank, unit, bytes_per_unit = 5000, "rays/s", None
        metric = "rays/sec fwd+bwd render_batch_ray (iMAP* mapping iteration: 256-wide MLP, 32+12 samples, regulation)"
        alg = {}
        info = {"workload": "imap_mapping_iter: 5000 rays, 32 + 12 importance samples (two passes), density compositing, regulation term",
                "rays_per_step_per_gpu": 5000, "parallelism": f"replicas x{world}"}
        after_step = lambda: None
        graph_gens = [gen]
        result_of = lambda out: out
    elif cfgname == "knn":
        # BASELINE config 4 (builder-defined semantics): 1,048,576 neural points U(bound), 32-ch features
        # N(0, 0.01^2), K = 8, the 5000-ray x 48-sample batch of the mapping step; fwd + bwd into features and camera pose
        from pointnerf_slam_b200 import knn as PK
        P_, bound, model, grids, renderer = build_scene(dev, True)
        fh = synthetic_frames(1, 100 + rank)[0]
        depth, color = fh[0].to(dev), fh[1].to(dev)
        pose = keyframe_poses(rank)[0].to(dev)
        cam = P.get_tensor_from_camera(pose).to(dev).requires_grad_(True)
        gen = torch.Generator(device=dev).manual_seed(79 + rank)
        gp = torch.Generator().manual_seed(4)
        b32 = bound.to(torch.float32)
        n_pts = 1 << 20
        xyz = (b32[:, 0] + (b32[:, 1] - b32[:, 0]) * torch.rand((n_pts, 3), generator=gp)).clamp(b32[:, 0], b32[:, 1]).contiguous().to(dev)
        feat = (0.01 * torch.randn((n_pts, 32), generator=gp)).to(dev).requires_grad_(True)
        field = PK.NeuralPointField(xyz, feat, bound, KNN_RADIUS, 1e-6)
        gw = torch.randn((5000 * S, 32), generator=gp).to(dev)          # stands in for the decoder's feature gradient

        def step_body():
            c = P.get_camera_from_tensor(cam)
            idx = torch.randint(H * W, (5000,), device=dev, generator=gen)
            o, d, gd, gc = P.get_samples(0, H, 0, W, 5000, H, W, FX, FY, CX, CY, c, depth, color, dev, indices=idx)
            z = renderer.sample_z(d, o, gd)
            f = field.aggregate_rays(o, d, z)
            loss = (f * gw).sum()
            cam.grad = None
            feat.grad = None
            loss.backward()
            if world > 1:
                dist.all_reduce(feat.grad)
            return loss

        def knn_cpu():
            with torch.no_grad():
                idx = torch.randint(H * W, (5000,), device=dev, generator=gen)
                o, d, gd, gc = P.get_samples(0, H, 0, W, 5000, H, W, FX, FY, CX, CY, pose, depth, color, dev, indices=idx)
                z = renderer.sample_z(d, o, gd)
                p = (o.double()[:, None, :] + d.double()[:, None, :] * z[..., None]).reshape(-1, 3).float().cpu()
            return knn_cpu_baseline(xyz.cpu(), feat.detach().cpu(), p, gw.cpu(), KNN_RADIUS)
        pinned = [fh[0].pin_memory(), fh[1].pin_memory()]
        dev_in = [depth, color]
        units_rank, unit, bytes_per_unit = 5000, "rays/s", BYTES_PER_RAY_KNN
        metric = "rays/sec fwd+bwd k-NN neural-point feature aggregation (1,048,576 points, K=8, 48 samples/ray)"
        n_samples = 5000 * S
        alg = {"knn_fwd": n_samples * BYTES_PER_SAMPLE_KNN_FWD, "knn_bwd": n_samples * BYTES_PER_SAMPLE_KNN_BWD}
        info = {"workload": "knn_aggregation: 5000 rays x 48 samples (the mapping step's sample placement) against 1,048,576 neural "
                            "points U(bound) with 32-ch features, K = 8 within radius %.2f m, fwd + bwd into the feature rows and "
                            "the camera 7-vector; BUILDER-DEFINED SEMANTICS (no reference implementation exists, SURVEY 8c)" % KNN_RADIUS,
                "rays_per_step_per_gpu": 5000, "samples_per_ray": S, "neural_points": n_pts,
                "parallelism": f"ray-shard dp{world}" + (" + NCCL all-reduce of the (P,32) feature gradient" if world > 1 else "")}
        after_step = lambda: None
        graph_gens = [gen]
        result_of = lambda out: out
    elif cfgname in ("dense", "mesh256"):
        P_, bound, model, grids, renderer = build_scene(dev, True)
        fh = synthetic_frames(1, 100)[0]
        depth, color = fh[0].to(dev), fh[1].to(dev)
        pose = keyframe_poses(0)[0].to(dev)
        if cfgname == "dense":
            def step_body():
                with torch.no_grad():
                    return D.render_img_sharded(renderer, grids, model, pose, dev, "color", depth, rank, world)
            units_total, unit, bytes_per_unit = H * W, "rays/s", BYTES_PER_RAY_DENSE
            metric = "rays/sec render_img (dense full-frame render, no grad, stage color)"
            pinned, dev_in = [fh[0].pin_memory()], [depth]
            n_chunks = -(-H * W // renderer.ray_batch_size)   # one launch per decoder and 100k-ray chunk: bytes of the AVERAGE launch
            n_s = (H * W // world) * S // n_chunks
            alg = {f"grid_mlp_fwd:{k}": n_s * (2048 if k == "fine" else 1024) for k in ("color", "fine", "middle")}
            info = {"workload": "dense_render: render_img of one 680x1200 frame = 816,000 rays x 48 samples in the reference's "
                                "100k-ray chunks; each chunk's rays shard over the ranks, outputs all-gathered",
                    "parallelism": f"ray-shard x{world} (strong scaling: one frame)"}
            result_of = lambda out: out[0]
        else:
            lo, hi = bound[:, 0] - 0.05, bound[:, 1] + 0.05       # Mesher.get_grid_uniform: bound +- 0.05 padding
            ax = [torch.linspace(float(lo[a]), float(hi[a]), 256) for a in range(3)]
            b, e = D.shard_bounds(256, rank, world)
            pts = torch.stack(torch.meshgrid(ax[0][b:e], ax[1], ax[2], indexing="ij"), -1).reshape(-1, 3).float().to(dev)

            def step_body():
                with torch.no_grad():
                    occ = renderer.eval_points(pts, model, grids, "fine", dev)[:, -1].contiguous()
                    return D.gather_shards(occ, 256 * 256 * 256, world)
            units_total, unit, bytes_per_unit = 256 ** 3, "points/s", BYTES_PER_POINT_MESH
            metric = "points/sec eval_points (256^3 mesh-extraction lattice, stage fine)"
            pinned, dev_in = [], []
            alg = {"grid_mlp_fwd:fine": pts.shape[0] * 2048, "grid_mlp_fwd:middle": pts.shape[0] * 1024}
            info = {"workload": "mesh256: eval_points on the 256^3 lattice of Mesher.get_grid_uniform, stage fine; x-slabs shard "
                                "over the ranks, occupancies all-gathered", "parallelism": f"point-shard x{world} (strong scaling)"}
            result_of = lambda out: out
        units_rank = units_total / world
        after_step = lambda: None
        graph_gens = []
    else:
        raise SystemExit(f"unknown config {cfgname}")

    # ---------------------------------------------------------------- stepping machinery
    graph = {"g": None, "out": None, "launches": 0}
    graphable = cfgname in ("mapping", "tracking", "imap", "knn") and not args.no_graph

    def h2d_inputs():
        for dst, src in zip(dev_in, pinned):
            dst.copy_(src, non_blocking=True)

    def eager_step(e2e=False):
        if e2e:
            h2d_inputs()
        out = step_body()
        res = None
        if e2e:
            r = result_of(out)
            res = r.item() if r.dim() == 0 else r.cpu()      # device -> host read of the step's result
        after_step()
        return res

    copy_stream = torch.cuda.Stream()
    staging = [[torch.empty_like(t) for t in dev_in] for _ in range(2)]
    ready = [torch.cuda.Event() for _ in range(2)]
    pipe = {"i": 0, "primed": False}

    def h2d_async(slot):
        with torch.cuda.stream(copy_stream):
            for dst, src in zip(staging[slot], pinned):
                dst.copy_(src, non_blocking=True)
            ready[slot].record(copy_stream)

    def step(e2e=False):
        if graph["g"] is None:
            return eager_step(e2e)
        if not e2e:
            graph["g"]()
            return None
        if e2e == "serial":                    # reference point: copy, then compute, on one stream
            h2d_inputs()
            graph["g"]()
            return graph["out"].item()
        # e2e: the copy of step i+1's inputs runs on a second stream while step i computes (what a mapper does with its
        # next keyframe); it lands in a staging buffer, a device-to-device copy moves it into the graph's static inputs,
        # and the step does not end before its prefetch has: every byte moves inside a timed region, one frame per step
        main = torch.cuda.current_stream()
        slot = pipe["i"] & 1
        if not pipe["primed"]:
            copy_stream.wait_stream(main)
            h2d_async(slot)
            pipe["primed"] = True
        main.wait_event(ready[slot])
        for dst, src in zip(dev_in, staging[slot]):
            dst.copy_(src)
        copy_stream.wait_stream(main)
        h2d_async(slot ^ 1)
        graph["g"]()
        out = graph["out"].item()
        main.wait_event(ready[slot ^ 1])
        pipe["i"] += 1
        return out

    # two-stream backward: a gain whenever no collective runs DURING the backward (one GPU, and the sparse exchanges, which
    # start after it); with NCCL all-reduces of finished gradients in flight it loses (overlap / arena modes)
    no_coll = world == 1 or (cfgname == "mapping" and os.environ.get("PN_BENCH_EXCHANGE", "sparse").startswith("sparse"))
    par_bwd = os.environ.get("PN_PARALLEL_BACKWARD", "1" if no_coll else "0") != "0"
    E.PARALLEL_BACKWARD = par_bwd

    def timed(k, e2e, profile):
        L.PROFILE = {} if profile else None
        E.PARALLEL_BACKWARD = (not profile) and par_bwd   # per-kernel event times: passes one after another
        evs = []
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        n0 = L.lib().pn_launch_count()
        for _ in range(k):
            flush.fill_(1)      # evict L2 between timed iterations (not timed)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            (eager_step if profile else step)(e2e)
            e1.record()
            evs.append((e0, e1))
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = sum(a.elapsed_time(b) for a, b in evs)
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        prof = L.PROFILE
        L.PROFILE = None
        E.PARALLEL_BACKWARD = par_bwd
        launched = L.lib().pn_launch_count() - n0
        if graph["g"] is not None and not profile:
            launched = graph["launches"] * k
        return t.item(), launched, prof

    warm = max(args.warmup, 3)
    for _ in range(warm):
        eager_step(False)
    # per-kernel durations (CUDA events around every C-ABI call) come from an eager pass over the same K steps;
    # the headline is timed on graph replays (events cannot be read back from inside a captured graph)
    ms_eager, launches_eager, prof = timed(args.steps, False, True)
    if graphable:
        try:
            gs = P.graphs.GraphedStep(step_body, generators=graph_gens, warmup=3)
            graph["g"], graph["out"], graph["launches"] = gs, gs.out, gs.launches
        except Exception as exc:   # a capture problem must not lose the measurement: fall back to eager launches
            print(f"[bench] CUDA-graph capture failed ({type(exc).__name__}: {exc}); timing eager launches", file=sys.stderr)
            graph["g"] = None
            graphable = False
            torch.cuda.synchronize()
    for _ in range(3):
        step(False)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.begin()
    ms_total, launches, _ = timed(args.steps, False, False)
    clocks = sampler.stop() if rank == 0 else None
    if args.light:
        ms_e2e = ms_e2e_serial = ms_total
    else:
        for _ in range(2):
            step(True)
        ms_e2e, _, _ = timed(args.steps, True, False)
        ms_e2e_serial, _, _ = timed(args.steps, "serial", False) if graph["g"] is not None else (ms_e2e, 0, None)

    # sustained mode: replay back to back for >= args.sustain seconds inside ONE timed region (clocks settle, power rises);
    # the untimed L2 flush of the headline loop is timed separately and subtracted
    sustained = None
    if args.sustain > 0 and not args.light:
        n_rep = max(args.steps, int(args.sustain * 1e3 / max(ms_total / args.steps, 1e-3)))
        s2 = ClockSampler(local)
        if rank == 0:
            s2.begin()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1, e2 = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e0.record()
        for _ in range(n_rep):
            flush.fill_(1)
            step(False)
        e1.record()
        for _ in range(n_rep):
            flush.fill_(1)
        e2.record()
        torch.cuda.synchronize()
        ms_s = torch.tensor([e0.elapsed_time(e1) - e1.elapsed_time(e2)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms_s, op=dist.ReduceOp.MAX)
        c2 = s2.stop() if rank == 0 else None
        sustained = {"steps": n_rep, "seconds": round(e0.elapsed_time(e1) * 1e-3, 3), "ms_per_step": round(ms_s.item() / n_rep, 4),
                     "value": round(units_rank * world * n_rep / (ms_s.item() * 1e-3), 1), "unit": unit,
                     "clocks": c2, "note": "back-to-back replays in one timed region, L2 flush between steps timed separately and subtracted"}

    units_step = units_rank * world
    value = units_step * args.steps / (ms_total * 1e-3)
    e2e_value = units_step * args.steps / (ms_e2e * 1e-3)
    kern = {k: [a.elapsed_time(b) for a, b in v] for k, v in (prof or {}).items()}
    share = {k: sum(v) / (ms_total if graph["g"] is not None else ms_eager) for k, v in kern.items()}
    top = max(kern, key=lambda k: sum(kern[k])) if kern else None
    roofline = None
    if top is not None:
        dur_ms = statistics.mean(kern[top])
        bytes_launch = alg.get(top)
        achieved = bytes_launch / (dur_ms * 1e-3) / 1e9 if bytes_launch else None
        roofline = {"bound": "hbm", "kernel": top, "achieved": round(achieved, 1) if achieved else None, "peak": hbm, "unit": "GB/s",
                    "frac": round(achieved / hbm, 4) if achieved else None,
                    "traffic": None,   # dram__bytes per launch is an ncu measurement: profiles/r2_dram_traffic_per_launch.json
                    "peak_source": which, "avg_launch_ms": round(dur_ms, 4), "alg_bytes_per_launch": bytes_launch,
                    "kernel_share_of_step": {k: round(v, 3) for k, v in sorted(share.items(), key=lambda kv: -kv[1])},
                    "kernel_ms": {k: round(statistics.mean(v), 4) for k, v in sorted(kern.items(), key=lambda kv: -sum(kv[1]))},
                    "kernel_frac_of_hbm": {k: round(alg[k] / (statistics.mean(v) * 1e-3) / 1e9 / hbm, 3) for k, v in kern.items() if alg.get(k)},
                    "note": ROOFLINE_NOTE_KNN if cfgname == "knn" else ROOFLINE_NOTE}
    step_frac = (value / world * bytes_per_unit / (hbm * 1e9)) if bytes_per_unit else None
    h2d_bytes = sum(t.numel() * t.element_size() for t in pinned)
    line = {"metric": metric, "value": round(value, 1), "unit": unit, "n_gpus": world, "steps": args.steps, "warmup": warm,
            "ms_per_step": round(ms_total / args.steps, 4), "higher_is_better": True,
            "scaling": "strong" if (args.scaling == "strong" or cfgname in ("dense", "mesh256")) else "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": dict(info, arithmetic="float32 results via 3xTF32 tensor-core products with FP32 accumulation (max error 5e-7 "
                           "relative, tests/test_gpu_tc.py); float64 geometry as in the reference",
                           l2="flushed between timed iterations (256 MiB fill, untimed); per-step CUDA events summed",
                           launch=("whole iteration captured once in a CUDA graph and replayed" if graph["g"] is not None else "eager launches"),
                           eager_ms_per_step=round(ms_eager / args.steps, 4)),
            "e2e": {"value": round(e2e_value, 1), "unit": unit, "h2d_bytes_per_step": h2d_bytes,
                    "d2h_bytes_per_step": 8 if cfgname in ("mapping", "tracking", "imap", "knn") else (H * W * 8 if cfgname == "dense" else 256 ** 3 * 4),
                    "ms_per_step": round(ms_e2e / args.steps, 4),
                    "h2d": "next step's inputs prefetched on a copy stream while the current step computes; a step ends only after "
                           "its prefetch has landed" if graph["g"] is not None else "inputs copied from pinned memory at the start of each step",
                    "serial_ms_per_step": round(ms_e2e_serial / args.steps, 4)},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline,
            "roofline_step": ({"bytes_per_unit": bytes_per_unit, "frac_of_hbm_per_gpu": round(step_frac, 4), "peak": hbm,
                               "peak_source": which} if step_frac is not None else None),
            "sustained": sustained}
    if cfgname == "mapping" and getattr(w.iteration, "_sparse", None) is not None:
        sp = w.iteration._sparse
        try:
            sp.check_overflow()
        except RuntimeError as exc:      # report it in the line instead of losing the measurement (the numbers are then suspect)
            print(f"[bench] {exc}", file=sys.stderr)
            line["config"]["exchange_overflow"] = str(exc)
        line["config"]["exchange_send_buffer_bytes"] = sp.bytes_per_rank()
        line["config"]["exchange_row_capacity"] = dict(sp.cap)
        line["config"]["touched_rows_this_rank"] = {k: int(c) for k, c in zip(sp.keys, sp.count.tolist())}
    if rank == 0 and world == 1 and not args.light and cfgname == "mapping":
        line["cpu_baseline"] = cpu_baseline(budget_s=15.0)
        line["cpu_baseline_1thread"] = cpu_baseline(budget_s=8.0, threads=1, pix=200)
        try:
            line["torch_eager_cuda"] = torch_eager_cuda_baseline(dev)
        except Exception as exc:
            line["torch_eager_cuda"] = {"unavailable": f"{type(exc).__name__}: {exc}"[:200]}
    if rank == 0 and world == 1 and not args.light and cfgname == "knn":
        line["cpu_baseline"] = knn_cpu()
    if rank == 0:
        emit(line)
    # Teardown.  A captured graph that contains NCCL kernels must be released before the process group
    # goes away, and no rank may sit in a collective at exit: drop the graph, synchronise, and leave
    # through os._exit (skipping NCCL's destructor-time rendezvous, which can wait forever on a
    # communicator that still has captured work).
    watchdog = threading.Timer(60.0, lambda: os._exit(0))   # never outlive the measurement by more than a minute
    watchdog.daemon = True
    watchdog.start()
    if graph["g"] is not None:
        graph["g"].release()
    graph["g"] = None
    graph["out"] = None
    for fn in cleanup:
        fn()
    torch.cuda.synchronize()
    if world > 1:
        try:
            dist.barrier()
        except Exception:
            pass
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


# ----------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port on the host cores
# ----------------------------------------------------------------------------
class OracleMapping:
    """The reference's mapping iteration restated with the oracle (pinned bit-equal against /root/reference): Mapper.py
    :413-431 (val[mask] parameters of the frustum-selected voxels), :509-518 (write-back before the forward), :551-662
    (sampling, render, loss, backward, Adam step with the stage learning rates), :665-674 (write-back after the step)."""

    def __init__(self, pix_per_kf, device="cpu", with_optimizer=True, seed=0, sd=None, grids=None, cams=None):
        """sd / grids / cams: optional state to start from (CPU tensors: decoder state dict, (1,32,Z,Y,X) grids, camera
        7-vectors) -- the parity tests hand over the CUDA arm's state; by default both arms are seeded independently."""
        from oracle import nice_oracle as O
        from oracle import mapper_oracle as M
        self.O, self.pix, self.dev = O, pix_per_kf, torch.device(device)
        self.bound = O.scene_bound(CFG["mapping"]["bound"], 1.0, 0.32)
        dev = self.dev
        sd = sd if sd is not None else O.init_nice_state(seed=seed)
        self.sd = {k: v.detach().clone().to(dev).requires_grad_(k.startswith("color_decoder.")) for k, v in sd.items()}
        if grids is None:
            grids = O.init_grids(self.bound, CFG["grid_len"], 32, 2, True, torch.Generator().manual_seed(1))
        self.frames_host = synthetic_frames(N_KEYFRAMES, 100)
        self.frames = [(d.to(dev), c.to(dev)) for d, c in self.frames_host]
        poses = keyframe_poses(0)
        from pointnerf_slam_b200.common import get_tensor_from_camera      # host-side input preparation (Mapper.py:470)
        cams = cams if cams is not None else [get_tensor_from_camera(poses[k]) for k in range(N_KEYFRAMES)]
        self.cams = [c.detach().clone().to(dev).requires_grad_(k > 0) for k, c in enumerate(cams)]
        self.gen = torch.Generator(device=dev).manual_seed(5)
        self.c = {k: v.detach().clone().contiguous().to(dev) for k, v in grids.items()}
        self.opt, self.masked = None, {}
        self.keys = ("grid_middle", "grid_fine", "grid_color")
        if with_optimizer:
            groups = {"grid_middle": [], "grid_fine": [], "grid_color": []}
            for k in self.keys:
                m3 = M.frustum_mask(poses[N_KEYFRAMES - 1].numpy(), k, self.c[k].shape[2:], self.frames_host[-1][0].numpy(), self.bound,
                                    H, W, FX, FY, CX, CY)
                mask = torch.from_numpy(m3).permute(2, 1, 0).unsqueeze(0).unsqueeze(0).repeat(1, 32, 1, 1, 1).to(dev)
                val_grad = self.c[k][mask].clone().requires_grad_(True)
                self.masked[k] = (val_grad, mask)
                groups[k].append(val_grad)
            lr = STAGE_LR["color"]
            self.opt = torch.optim.Adam([{"params": [v for k, v in self.sd.items() if v.requires_grad], "lr": lr["decoders_lr"]},
                                         {"params": groups["grid_middle"], "lr": lr["middle_lr"]},
                                         {"params": groups["grid_fine"], "lr": lr["fine_lr"]},
                                         {"params": groups["grid_color"], "lr": lr["color_lr"]},
                                         {"params": self.cams[1:], "lr": 0.001}])
        else:
            for k in self.keys:
                self.c[k].requires_grad_(True)

    def step(self, indices=None):
        O = self.O
        if self.opt is not None:
            for k, (val_grad, mask) in self.masked.items():       # Mapper.py:509-518
                val = self.c[k].detach()
                val[mask] = val_grad
                self.c[k] = val
        ro, rd, gd, gc = [], [], [], []
        for k in range(N_KEYFRAMES):
            idx = indices[k] if indices is not None else torch.randint(H * W, (self.pix,), generator=self.gen, device=self.dev)
            c2w = O.camera_from_tensor(self.cams[k])                          # Mapper.py:586
            o, d, dd, cc, _ = O.get_samples(0, H, 0, W, self.pix, FX, FY, CX, CY, c2w, self.frames[k][0], self.frames[k][1], idx)
            ro.append(o); rd.append(d); gd.append(dd); gc.append(cc)
        ro, rd, gd, gc = torch.cat(ro), torch.cat(rd), torch.cat(gd), torch.cat(gc)
        scene = O.Scene(self.sd, self.c, self.bound, nice=True, occupancy=True)
        depth, var, color = O.render_batch_ray(scene, rd, ro, "color", gd)
        loss = O.mapping_loss(depth, color, gd, gc, "color", W_COLOR)
        loss.backward()
        if self.opt is not None:
            self.opt.step()
            self.opt.zero_grad()
            for k, (val_grad, mask) in self.masked.items():       # Mapper.py:665-674
                val = self.c[k].detach()
                val[mask] = val_grad.clone().detach()
                self.c[k] = val
        return loss

    def zero_grad(self):
        for t in list(self.sd.values()) + list(self.c.values()) + self.cams:
            t.grad = None


def cpu_baseline(budget_s=15.0, threads=None, pix=PIX_PER_KF):
    """The same mapping iteration (5 x `pix` px) on the host cores, as many steps as fit the budget after one warm-up."""
    cores = threads or (os.cpu_count() or 1)
    prev = torch.get_num_threads()
    torch.set_num_threads(cores)
    om = OracleMapping(pix)
    om.step()
    n, t0 = 0, time.perf_counter()
    best = float("inf")
    while n < 1 or (time.perf_counter() - t0 < budget_s and n < 20):
        t1 = time.perf_counter()
        om.step()
        best = min(best, time.perf_counter() - t1)
        n += 1
    rays = N_KEYFRAMES * pix
    torch.set_num_threads(prev)
    return {"value": round(rays / best, 1), "unit": "rays/s", "cores": cores, "kind": "port",
            "sample": f"same mapping iteration (incl. Adam step) on {N_KEYFRAMES} x {pix} px = {rays} rays, torch CPU, best of {n} after 1 warm-up"}


def knn_cpu_baseline(xyz, feat, p, gw, radius, rays=200, budget_s=12.0):
    """k-NN aggregation fwd+bwd of the oracle port on the host cores: scipy cKDTree ball query (float64) + float32 re-ranking
    + torch blend and autograd, on the first `rays` rays of the step's sample points (the tree build is not timed: our index
    build is not part of the step either)."""
    from scipy.spatial import cKDTree
    from oracle import knn_oracle as KO
    tree = cKDTree(xyz.double().numpy())
    n = rays * S
    pn, xn = p[:n].contiguous(), xyz.numpy()
    best, reps, t_begin = float("inf"), 0, time.perf_counter()
    while reps < 1 or (time.perf_counter() - t_begin < budget_s and reps < 10):
        t0 = time.perf_counter()
        idx, _ = KO.knn_query(pn.numpy(), xn, radius, tree)
        pt = pn.clone().requires_grad_(True)
        ft = feat.clone().requires_grad_(True)
        (KO.aggregate(pt, xyz, ft, torch.from_numpy(idx), 1e-6) * gw[:n]).sum().backward()
        best = min(best, time.perf_counter() - t0)
        reps += 1
    return {"value": round(rays / best, 1), "unit": "rays/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"oracle port (cKDTree.query_ball_point + float32 re-ranking + torch blend, fwd+bwd) on {rays} rays x {S} samples "
                      f"of the same batch against the same 1,048,576 points, best of {reps}"}


def torch_eager_cuda_baseline(dev, steps=10):
    """The reference's op graph (oracle port, unmodified torch ops, allow_tf32 off) on the same GPU: the reference's real
    deployment is torch eager on CUDA, so this is the honest number to beat."""
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    om = OracleMapping(PIX_PER_KF, device=dev)
    for _ in range(3):
        om.step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        om.step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    return {"value": round(N_KEYFRAMES * PIX_PER_KF / ms * 1e3, 1), "unit": "rays/s", "ms_per_step": round(ms, 3), "steps": steps,
            "what": "oracle port of the reference's torch path (grid_sample, addmm, sort, cumprod, torch.optim.Adam on val[mask]) run "
                    "eagerly on cuda:0, allow_tf32=False, same 5 x 1000 px mapping iteration"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if args.config != "mapping":
        emit({"impl": "reference", "unavailable": f"the reference arm times the headline mapping iteration only (asked for {args.config})"})
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    om = OracleMapping(PIX_PER_KF)            # the same 5 x 1000 px step as our arm
    cap_s = float(os.environ.get("PN_BENCH_REF_CAP_S", "150"))
    t_begin = time.perf_counter()
    warm = 0
    for _ in range(max(args.warmup, 1)):
        om.step()
        warm += 1
        if time.perf_counter() - t_begin > cap_s / 3:
            break
    steps, t0 = 0, time.perf_counter()
    for _ in range(args.steps):
        om.step()
        steps += 1
        if time.perf_counter() - t_begin > cap_s:
            break
    dt = time.perf_counter() - t0
    rays = N_KEYFRAMES * PIX_PER_KF
    value = rays * steps / dt
    capped = steps < args.steps or warm < max(args.warmup, 1)
    sample = (f"{N_KEYFRAMES} x {PIX_PER_KF} px = {rays} rays per step (the full workload), {steps} timed steps after {warm} warm-up, "
              f"torch CPU ({torch.get_num_threads()} threads)" + (f"; stopped by the {cap_s:.0f} s wall-clock cap" if capped else ""))
    line = {"impl": "reference", "metric": METRIC, "value": round(value, 1), "unit": "rays/s", "n_gpus": args.gpus, "steps": steps,
            "warmup": warm, "ms_per_step": round(dt / steps * 1e3, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name("weak"), "rays_per_step_per_gpu": rays, "samples_per_ray": S,
                       "optimizer_in_step": True,
                       "note": "oracle port of the reference's torch path (pinned bit-equal against /root/reference), host cores; "
                               "steps/warmup as asked unless the wall-clock cap says otherwise"},
            "cpu_baseline": {"value": round(value, 1), "unit": "rays/s", "cores": torch.get_num_threads(), "kind": "port", "sample": sample},
            "e2e": {"value": round(value, 1), "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


_JSON_FD = None


def emit(line):
    """The one JSON line goes to the process's ORIGINAL stdout; everything else any library prints
    (NCCL's version banner, warnings) has been routed to stderr by main()."""
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="mapping", choices=["mapping", "tracking", "dense", "mesh256", "imap", "knn"])
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--sustain", type=float, default=2.0, help="seconds of back-to-back replays for the sustained-clock figure (0: skip)")
    ap.add_argument("--no-optimizer", action="store_true", help="mapping: leave the Adam step out of the timed iteration")
    ap.add_argument("--light", action="store_true", help="profiling runs: skip the e2e pass, the sustained pass and the CPU baselines")
    ap.add_argument("--no-graph", action="store_true", help="time eager launches instead of CUDA-graph replays")
    args = ap.parse_args()
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)     # keep stdout for the JSON line only
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
