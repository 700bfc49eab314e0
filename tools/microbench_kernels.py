"""Device-side (CUDA event) timing of single library kernels for a list of linear shapes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hicom_b200 import ops

dev = "cuda"
for impl_name, impl in (("auto", ops.IMPL_AUTO),):
    for (M, N, K) in [(32, 2304, 2304), (32, 1152, 1152), (32, 3584, 1152), (32, 3584, 3584), (1024, 1152, 1152), (288, 1152, 1152)]:
        A = torch.randn(M, K, device=dev).bfloat16()
        W = (0.02 * torch.randn(N, K, device=dev)).bfloat16()
        b = torch.randn(N, device=dev).bfloat16()
        for _ in range(3):
            ops.linear(A, W, b, None, 1, False, impl)
        with ops.KernelTimer() as kt:
            for _ in range(20):
                ops.linear(A, W, b, None, 1, False, impl)
        for k, (c, ms) in kt.summary().items():
            print(f"{impl_name:8s} M={M:3d} N={N:5d} K={K:5d}  {k[:40]:40s} {ms / c * 1e3:8.1f} us")
