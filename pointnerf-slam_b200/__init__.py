"""B200-native differentiable ray-rendering path of pointNeRF-SLAM / NICE-SLAM.

Import as ``pointnerf_slam_b200`` (the directory name carries a hyphen, which
Python cannot import; ``pointnerf_slam_b200/__init__.py`` at the repo root maps
the importable name onto this directory).

Modules
  _lib      ctypes binding of libpnslam.so (C ABI in include/pnslam.h)
  engine    decoder passes, stashes, autograd boundaries
  renderer  Renderer  (drop-in for src/utils/Renderer.py)
  decoder   NICE / MLP / MLP_no_xyz  (drop-in for src/conv_onet/models/decoder.py)
  common    get_samples / get_rays / get_camera_from_tensor / ...  (src/common.py)
  config    get_model / load_bound / grid_init  (src/config.py, src/NICE_SLAM.py)
  dist      ray sharding + NCCL gradient all-reduce for the mapping step
  losses    mapping_loss / tracking_loss: the Mapper's and Tracker's loss heads with their gradients, one launch each
  graphs    GraphedStep: one tracking / mapping iteration captured in a CUDA graph
  csrc/     CUDA kernels (sm_100a) and the C ABI
"""
