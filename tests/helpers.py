"""Shared fixtures for the parity tests: rebuild the golden scenes on either side."""
import os
import types

import numpy as np
import torch

from oracle import nice_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
RENDER_CFG = {"rendering": {"lindisp": False, "perturb": 0.0, "N_samples": 32, "N_surface": 16, "N_importance": 0},
              "scale": 1, "occupancy": True}


def load(name):
    z = np.load(os.path.join(GOLDEN, name))
    return {k: torch.from_numpy(np.asarray(z[k])) for k in z.files}


def load_nice():
    """Scene (bound, grids, decoder state) + render inputs/outputs of the small NICE scene."""
    g = load("nice_eval_points.npz")
    g.update(load("nice_render.npz"))
    return g


def state_dict(g):
    return {k[3:]: v for k, v in g.items() if k.startswith("sd/")}


def oracle_scene(g, grids=None, sd=None, **kw):
    grids = grids if grids is not None else {k: g[k] for k in O.GRID_KEYS if k in g}
    return O.Scene(sd if sd is not None else state_dict(g), grids, g["bound"], **kw)


def cuda_nice(g, device="cuda:0", channels_last=True):
    """This package's NICE decoders + grids + Renderer for a golden scene."""
    import pointnerf_slam_b200 as P
    model = P.NICE(coarse=True).to(device)
    model.load_state_dict(state_dict(g))
    P.attach_bounds(model, g["bound"])
    grids = {}
    for k in O.GRID_KEYS:
        if k in g:
            t = g[k].to(device)
            grids[k] = t.contiguous(memory_format=torch.channels_last_3d) if channels_last else t.contiguous()
    slam = types.SimpleNamespace(bound=g["bound"], H=68, W=120, fx=60.0, fy=60.0, cx=59.5, cy=33.5, nice=True)
    renderer = P.Renderer(RENDER_CFG, None, slam)
    return model, grids, renderer


def rel_max(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def assert_close_q(a, b, rtol, atol, rtol_max, atol_max, q=0.999, what=""):
    """Two-level closeness for float32 pipelines whose oracle itself carries rounding noise (sin of arguments of hundreds of
    radians): at least a fraction q of the elements within the tight (rtol, atol), every element within the loose
    (rtol_max, atol_max).  The message reports the measured quantiles so that tolerances can be kept near them."""
    a, b = a.detach().double().cpu().reshape(-1), b.detach().double().cpu().reshape(-1)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    err = (a - b).abs()
    tight = err <= atol + rtol * b.abs()
    loose = err <= atol_max + rtol_max * b.abs()
    frac = tight.double().mean().item() if err.numel() else 1.0
    scale = b.abs().max().clamp_min(1e-30)
    qs = torch.quantile(err[:: max(1, err.numel() // 1_000_000)] / scale, torch.tensor([0.5, 0.99, 0.999], dtype=torch.float64)).tolist() \
        if err.numel() else [0, 0, 0]
    msg = (f"{what}: {100 * frac:.3f}% within tight tol (need {100 * q:.1f}%), max err / max|ref| = {(err.max() / scale).item():.2e}, "
           f"quantiles 50/99/99.9% = {qs[0]:.1e}/{qs[1]:.1e}/{qs[2]:.1e}")
    assert frac >= q and bool(loose.all()), msg
    return msg
