"""tcgen05 building blocks: the 3xTF32 tensor-core product must be FP32-accurate."""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("K", [8, 32, 96, 128])
def test_tc_selftest_fp32_accuracy(K):
    from pointnerf_slam_b200 import _lib as L
    torch.manual_seed(K)
    X = (torch.randn(128, K) * 3).to(DEV)
    W = torch.randn(32, K).to(DEV)
    Y = torch.zeros(128, 32, device=DEV)
    L.check(L.lib().pn_tc_selftest(C.c_void_p(X.data_ptr()), C.c_void_p(W.data_ptr()), C.c_void_p(Y.data_ptr()), K,
                                   C.c_void_p(torch.cuda.current_stream().cuda_stream)), "pn_tc_selftest")
    torch.cuda.synchronize()
    ref = (X.double() @ W.double().T)
    scale = (X.double().abs() @ W.double().abs().T)          # magnitude of the summed terms
    err = ((Y.double() - ref).abs() / scale).max().item()
    fp32 = (((X @ W.T).double() - ref).abs() / scale).max().item()
    print(f"K={K}: 3xTF32 err {err:.2e}  (torch fp32 matmul {fp32:.2e})")
    assert err < 1e-6, f"3xTF32 product is not FP32-accurate: {err:.2e}"
