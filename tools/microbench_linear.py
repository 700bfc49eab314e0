"""Time hicom_b200.ops.linear variants with CUDA events (diagnostic tool)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hicom_b200 import ops


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3


def main():
    dev = "cuda"
    for (M, N, K) in [(32, 2304, 1152), (32, 2304, 2304), (1024, 1152, 1152), (10368, 3584, 1152), (10368, 3584, 3584),
                      (128, 256, 1152), (128, 256, 64), (18944, 256, 1152)]:
        A = torch.randn(M, K, device=dev).bfloat16()
        W = (0.02 * torch.randn(N, K, device=dev)).bfloat16()
        b = torch.randn(N, device=dev).bfloat16()
        for name, fn in [
            ("bias+gelu", lambda: ops.linear(A, W, b, None, 1, False, 0)),
            ("bias", lambda: ops.linear(A, W, b, None, 0, False, 0)),
            ("nobias", lambda: ops.linear(A, W, None, None, 0, False, 0)),
            ("nobias f32out", lambda: ops.linear(A, W, None, None, 0, True, 0)),
        ]:
            us = timeit(fn)
            print(f"M={M:6d} N={N:5d} K={K:5d} {name:14s} {us:9.1f} us  {2*M*N*K/us/1e6:8.1f} TFLOP/s", flush=True)


if __name__ == "__main__":
    main()
