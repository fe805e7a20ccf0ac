#!/usr/bin/env python
"""Error distribution of the CUDA path against the oracle at full size (forward outputs and mapping gradients):
quantiles of |cuda - oracle| / (atol-free) scale, used to set the test tolerances (tests/helpers.assert_close_q)."""
import os, sys, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import tests.test_gpu_fullsize as F
from oracle import nice_oracle as O

w = F.world.__wrapped__() if hasattr(F.world, "__wrapped__") else None
