"""Decoder modules with the reference's names, shapes and call signatures
(reference: src/conv_onet/models/decoder.py), evaluated by fused CUDA kernels.

``state_dict`` keys are identical to the reference's (``middle_decoder.fc_c.0.weight``,
``fine_decoder.embedder._B`` ...), so its checkpoints (src/utils/Logger.py:25)
and the upstream pretrained-decoder loader (src/NICE_SLAM.py:225-255) work
unchanged.  The forward of every module is one kernel pass per sub-decoder:

    NICE.forward(p (1,N,3)|(N,3), c_grid, stage='middle')  -> (N,4)
    MLP.forward(p, c_grid=None)                            -> (N,) | (N,4)
    MLP_no_xyz.forward(p, c_grid)                          -> (N,)

Each sub-decoder needs a ``.bound`` attribute ((3,2) float64, set by the
orchestrator, src/NICE_SLAM.py:216-221) before it is called.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from . import engine as E


class GaussianFourierFeatureTransform(nn.Module):
    """Holds the (3, 93) projection ``_B`` (decoder.py:7-30); sin(p @ B) is
    evaluated inside the fused decoder kernel."""

    def __init__(self, num_input_channels, mapping_size=93, scale=25, learnable=True):
        super().__init__()
        B = torch.randn((num_input_channels, mapping_size)) * scale
        if learnable:
            self._B = nn.Parameter(B)
        else:
            self.register_buffer("_B", B, persistent=False)


class DenseLayer(nn.Linear):
    """Linear layer with xavier-uniform weights scaled by the activation gain
    and zero bias (decoder.py:70-79)."""

    def __init__(self, in_dim: int, out_dim: int, activation: str = "relu", *args, **kwargs) -> None:
        self.activation = activation
        super().__init__(in_dim, out_dim, *args, **kwargs)

    def reset_parameters(self) -> None:
        nn.init.xavier_uniform_(self.weight, gain=nn.init.calculate_gain(self.activation))
        if self.bias is not None:
            nn.init.zeros_(self.bias)


def _points(p: torch.Tensor) -> torch.Tensor:
    if p.dim() == 3:
        p = p.squeeze(0)
    if p.dim() != 2 or p.shape[-1] != 3:
        raise RuntimeError(f"decoder expects points of shape (1,N,3) or (N,3), got {tuple(p.shape)}")
    if p.dtype not in (torch.float32, torch.float64):
        p = p.float()
    return p


class MLP(nn.Module):
    """Grid decoder with Fourier-embedded xyz input (decoder.py:91-203), or the
    iMAP* single MLP when ``c_dim == 0``."""

    def __init__(self, name='', dim=3, c_dim=128, hidden_size=256, n_blocks=5, leaky=False, sample_mode='bilinear',
                 color=False, skips=[2], grid_len=0.16, pos_embedding_method='fourier', concat_feature=False):
        super().__init__()
        if pos_embedding_method != 'fourier':
            raise NotImplementedError("only the 'fourier' positional embedding (all shipped configs) is built")
        if leaky or sample_mode != 'bilinear':
            raise NotImplementedError("leaky ReLU / nearest sampling are unused by every shipped config")
        self.name, self.color, self.c_dim, self.grid_len = name, color, c_dim, grid_len
        self.no_grad_feature = False
        self.concat_feature, self.n_blocks, self.skips = concat_feature, n_blocks, list(skips)
        self.sample_mode = sample_mode
        self.bound = None
        if c_dim != 0:
            self.fc_c = nn.ModuleList([nn.Linear(c_dim, hidden_size) for _ in range(n_blocks)])
        embedding_size = 93
        self.embedder = GaussianFourierFeatureTransform(dim, mapping_size=embedding_size, scale=25)
        layers = [DenseLayer(embedding_size, hidden_size, activation="relu")]
        for i in range(n_blocks - 1):
            in_dim = hidden_size + embedding_size if i in self.skips else hidden_size
            layers.append(DenseLayer(in_dim, hidden_size, activation="relu"))
        self.pts_linears = nn.ModuleList(layers)
        self.output_linear = DenseLayer(hidden_size, 4 if color else 1, activation="linear")
        if c_dim != 0 and not (hidden_size == 32 and n_blocks == 5 and self.skips == [2] and c_dim in (32, 64)):
            raise NotImplementedError("grid decoders are built for hidden 32, 5 blocks, skip [2], c_dim 32/64 "
                                      "(the NICE configuration)")

    def forward(self, p, c_grid=None):
        p = _points(p)
        kind = "grid" if self.c_dim != 0 else "imap"
        plan = E.Plan(E.single_pass(self, kind, self.bound), None)
        raw = E.eval_plan(plan, p, c_grid if c_grid is not None else {})
        return raw if self.color else raw[:, 3]


class MLP_no_xyz(nn.Module):
    """Coarse-level decoder that sees only the interpolated feature
    (decoder.py:206-274)."""

    def __init__(self, name='', dim=3, c_dim=128, hidden_size=256, n_blocks=5, leaky=False, sample_mode='bilinear',
                 color=False, skips=[2], grid_len=0.16):
        super().__init__()
        if not (hidden_size == 32 and c_dim == 32 and n_blocks == 5 and list(skips) == [2] and not color):
            raise NotImplementedError("coarse decoder is built for hidden 32, c_dim 32, 5 blocks, skip [2]")
        self.name, self.color, self.c_dim, self.grid_len = name, color, c_dim, grid_len
        self.no_grad_feature = False
        self.n_blocks, self.skips, self.sample_mode = n_blocks, list(skips), sample_mode
        self.bound = None
        layers = [DenseLayer(hidden_size, hidden_size, activation="relu")]
        for i in range(n_blocks - 1):
            in_dim = hidden_size + c_dim if i in self.skips else hidden_size
            layers.append(DenseLayer(in_dim, hidden_size, activation="relu"))
        self.pts_linears = nn.ModuleList(layers)
        self.output_linear = DenseLayer(hidden_size, 1, activation="linear")

    def forward(self, p, c_grid, **kwargs):
        p = _points(p)
        plan = E.Plan(E.single_pass(self, "coarse", self.bound), None)
        return E.eval_plan(plan, p, c_grid)[:, 3]


class NICE(nn.Module):
    """Coarse / middle / fine / colour decoders (decoder.py:277-342)."""

    def __init__(self, dim=3, c_dim=32, coarse_grid_len=2.0, middle_grid_len=0.16, fine_grid_len=0.16,
                 color_grid_len=0.16, hidden_size=32, coarse=False, pos_embedding_method='fourier'):
        super().__init__()
        if coarse:
            self.coarse_decoder = MLP_no_xyz(name='coarse', dim=dim, c_dim=c_dim, color=False, hidden_size=hidden_size,
                                             grid_len=coarse_grid_len)
        self.middle_decoder = MLP(name='middle', dim=dim, c_dim=c_dim, color=False, skips=[2], n_blocks=5,
                                  hidden_size=hidden_size, grid_len=middle_grid_len,
                                  pos_embedding_method=pos_embedding_method)
        self.fine_decoder = MLP(name='fine', dim=dim, c_dim=c_dim * 2, color=False, skips=[2], n_blocks=5,
                                hidden_size=hidden_size, grid_len=fine_grid_len, concat_feature=True,
                                pos_embedding_method=pos_embedding_method)
        self.color_decoder = MLP(name='color', dim=dim, c_dim=c_dim, color=True, skips=[2], n_blocks=5,
                                 hidden_size=hidden_size, grid_len=color_grid_len,
                                 pos_embedding_method=pos_embedding_method)
        self.bound = None

    def forward(self, p, c_grid, stage='middle', **kwargs):
        p = _points(p)
        plan = E.Plan(E.stage_passes(self, stage, self.bound), None)
        return E.eval_plan(plan, p, c_grid)
