"""B200-native differentiable ray-rendering path of pointNeRF-SLAM / NICE-SLAM.

Import as ``pointnerf_slam_b200`` (``pointnerf-slam_b200`` at the repo root is a
symbolic link to this directory: the name the project is known by, which Python
cannot import because of the hyphen).

Modules
  _lib      ctypes binding of libpnslam.so (C ABI in include/pnslam.h)
  engine    decoder passes, stashes, autograd boundaries
  renderer  Renderer  (drop-in for src/utils/Renderer.py)
  decoder   NICE / MLP / MLP_no_xyz  (drop-in for src/conv_onet/models/decoder.py)
  common    get_samples / get_rays / get_camera_from_tensor / ...  (src/common.py)
  config    get_model / load_bound / grid_init  (src/config.py, src/NICE_SLAM.py)
  dist      ray sharding + NCCL gradient all-reduce for the mapping step
  losses    mapping_loss / tracking_loss: the Mapper's and Tracker's loss heads with their gradients, one launch each
  mapping   MappingIteration: one iteration of the Mapper's hot loop (sample, render, loss, backward, exchange, Adam step)
  tracking  TrackingIteration: one iteration of the Tracker's pose optimisation
  graphs    GraphedStep: one tracking / mapping iteration captured in a CUDA graph
  knn       NeuralPointField / NeuralPointIndex: k-nearest neural-point feature aggregation (BASELINE config 4; builder-defined semantics)
  csrc/     CUDA kernels (sm_100a) and the C ABI
"""

from . import _lib, engine, common, decoder, config, renderer, graphs, losses, mapper, knn, mapping, tracking  # noqa: E402,F401
from .renderer import Renderer  # noqa: E402,F401
from .decoder import NICE, MLP, MLP_no_xyz  # noqa: E402,F401
from .config import get_model, load_bound, grid_init, attach_bounds  # noqa: E402,F401
from .common import (get_samples, get_samples_multi, KeyframeBatch, get_rays, get_rays_from_uv, get_camera_from_tensor,  # noqa: E402,F401
                     get_tensor_from_camera, raw2outputs_nerf_color, sample_pdf, normalize_3d_coordinate)
