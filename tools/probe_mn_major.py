import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pointnerf_slam_b200 import _lib as L
lib = L.lib()
lib.pn_tc_selftest_mn.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
for K in (8, 32):
    for swap in (0, 1):
        torch.manual_seed(K)
        Xt = torch.randn(K, 128, device="cuda"); Wt = torch.randn(K, 32, device="cuda"); Y = torch.full((128, 32), -7.0, device="cuda")
        rc = lib.pn_tc_selftest_mn(Xt.data_ptr(), Wt.data_ptr(), Y.data_ptr(), K, swap, torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        ref = Xt.double().T @ Wt.double()
        print(f"K={K} swap={swap} rc={rc} max|Y|={Y.abs().max().item():.3e} err={(Y.double()-ref).abs().max().item():.3e} ref max {ref.abs().max().item():.3e}")
        # diagnose permutations: is Y equal to ref for some sub-block?
        print("   Y[0,:4]", Y[0,:4].tolist(), " ref[0,:4]", ref[0,:4].float().tolist())
