"""The C-ABI library loads on a CPU-only box and exports every symbol that
include/pnslam.h declares (no compute calls are made here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "pnslam.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pn_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_path():
    names = declared_symbols()
    for must in ("pn_grid_mlp_fwd", "pn_grid_mlp_bwd", "pn_grid_mlp_wgrad", "pn_composite_fwd", "pn_composite_bwd",
                 "pn_ray_zvals", "pn_sample_rays_fwd", "pn_rays_bwd", "pn_camera_from_tensor_fwd", "pn_last_error"):
        assert must in names


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as ge
    from pointnerf_slam_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        ge.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    missing = [n for n in declared_symbols() if not hasattr(lib, n)]
    assert not missing, f"libpnslam.so lacks {missing}"
    lib.pn_version.restype = ctypes.c_int
    assert lib.pn_version() >= 100


def test_product_path_refuses_cpu_tensors():
    """No CPU fallback: CPU inputs raise instead of silently computing."""
    import torch
    import pointnerf_slam_b200 as P
    model = P.NICE(coarse=True)
    P.attach_bounds(model, torch.tensor([[-1.0, 1.0]] * 3, dtype=torch.float64))
    c = {k: torch.zeros(1, 32, 4, 4, 4) for k in ("grid_middle", "grid_fine", "grid_color")}
    with pytest.raises(RuntimeError):
        model(torch.zeros(1, 8, 3), c, stage="color")


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "pointnerf_slam_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            txt = open(os.path.join(pkg, fn)).read()
            assert "oracle" not in txt.replace("# oracle", ""), f"{fn} mentions the oracle"
