"""Model factory and scene set-up for the hot path.

``get_model`` mirrors src/config.py:63-79 + src/conv_onet/config.py:4-33;
``load_bound`` / ``grid_init`` / ``attach_bounds`` restate the orchestrator
duties of src/NICE_SLAM.py:200-315 (commented out in the fork, live upstream)
that the renderer depends on.  Grids are created directly in
``torch.channels_last_3d`` memory format -- logical shape (1, C, Z, Y, X) like
the reference, memory [Z][Y][X][C] as the kernels want -- so no transposition
happens on the hot path.
"""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np
import torch

from .decoder import MLP, NICE

decoder_dict = {'nice': NICE, 'imap': MLP}


def get_model(cfg, nice=True):
    """NICE decoders, or the iMAP* MLP (hidden 256, 4 blocks, no skips)."""
    dim = cfg['data']['dim']
    if nice:
        g = cfg['grid_len']
        return NICE(dim=dim, c_dim=cfg['model']['c_dim'], coarse=cfg['coarse'], coarse_grid_len=g['coarse'],
                    middle_grid_len=g['middle'], fine_grid_len=g['fine'], color_grid_len=g['color'],
                    pos_embedding_method=cfg['model']['pos_embedding_method'])
    return MLP(dim=dim, c_dim=0, color=True, hidden_size=256, skips=[], n_blocks=4,
               pos_embedding_method=cfg['model']['pos_embedding_method'])


def load_bound(cfg) -> torch.Tensor:
    """float64 (3,2) scene bound, upper edge enlarged to a multiple of
    ``bound_divisible``; the int32*float product is float32 on purpose
    (src/NICE_SLAM.py:208-213)."""
    bound = torch.from_numpy(np.array(cfg['mapping']['bound'], dtype=np.float64) * cfg['scale'])
    div = cfg['grid_len']['bound_divisible']
    bound[:, 1] = (((bound[:, 1] - bound[:, 0]) / div).int() + 1) * div + bound[:, 0]
    return bound


def attach_bounds(decoders, bound: torch.Tensor, coarse_bound_enlarge: float = 2) -> None:
    """src/NICE_SLAM.py:216-221."""
    decoders.bound = bound
    for name in ("middle_decoder", "fine_decoder", "color_decoder"):
        if hasattr(decoders, name):
            getattr(decoders, name).bound = bound
    if hasattr(decoders, "coarse_decoder"):
        decoders.coarse_decoder.bound = bound * coarse_bound_enlarge


def grid_shapes(cfg, bound: torch.Tensor) -> Dict[str, tuple]:
    """(Z,Y,X) per level: truncating division and x<->z swap (src/NICE_SLAM.py:276-311)."""
    extent = bound[:, 1] - bound[:, 0]
    out = {}
    for level in ("coarse", "middle", "fine", "color"):
        if level == "coarse":
            if not cfg.get('coarse', False):
                continue
            n = list(map(int, (extent * cfg['model']['coarse_bound_enlarge'] / cfg['grid_len'][level]).tolist()))
        else:
            n = list(map(int, (extent / cfg['grid_len'][level]).tolist()))
        out["grid_" + level] = (n[2], n[1], n[0])
    return out


def grid_init(cfg, bound: torch.Tensor, device, generator: Optional[torch.Generator] = None) -> Dict[str, torch.Tensor]:
    """Feature grids ~ N(0, 0.01^2) (fine: N(0, 1e-4^2)), src/NICE_SLAM.py:283-313."""
    c_dim = cfg['model']['c_dim']
    c = {}
    for key, (z, y, x) in grid_shapes(cfg, bound).items():
        std = 0.0001 if key == "grid_fine" else 0.01
        val = torch.zeros((1, c_dim, z, y, x)).normal_(mean=0, std=std, generator=generator)
        c[key] = val.to(device).contiguous(memory_format=torch.channels_last_3d)
    return c
