"""Importable name of the package that lives in ``pointnerf-slam_b200/``.

The product directory is named after the reference repository and contains a
hyphen; this shim points ``__path__`` at it so that
``import pointnerf_slam_b200.renderer`` etc. resolve to the real files.
"""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "pointnerf-slam_b200")
if not _os.path.isdir(_real):
    raise ImportError(f"package directory {_real} is missing")
__path__.insert(0, _real)
__doc__ = open(_os.path.join(_real, "__init__.py")).read().split('"""')[1]

from . import _lib, engine, common, decoder, config, renderer, graphs, losses  # noqa: E402,F401
from .renderer import Renderer  # noqa: E402,F401
from .decoder import NICE, MLP, MLP_no_xyz  # noqa: E402,F401
from .config import get_model, load_bound, grid_init, attach_bounds  # noqa: E402,F401
from .common import (get_samples, get_rays, get_rays_from_uv, get_camera_from_tensor,  # noqa: E402,F401
                     get_tensor_from_camera, raw2outputs_nerf_color, sample_pdf, normalize_3d_coordinate)
