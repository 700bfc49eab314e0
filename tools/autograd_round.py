"""One-process GPU check of the training path: tests/test_gpu_autograd.py, then a CUDA-event timing of
one forward+backward of the drop-in projector.  Writes gpurun_out/autograd_tests.log and gpurun_out/autograd_bench.json.

    python tools/autograd_round.py [--videos 8] [--skip-tests] [-k EXPR]
"""
import argparse
import io
import json
import os
import sys
from contextlib import redirect_stderr, redirect_stdout

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
OUT = os.path.join(ROOT, "gpurun_out")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--videos", type=int, default=8)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--skip-tests", action="store_true")
    ap.add_argument("-k", default=None)
    args = ap.parse_args()
    os.makedirs(OUT, exist_ok=True)
    if not args.skip_tests:
        import pytest
        buf = io.StringIO()
        argv = [os.path.join(ROOT, "tests", "test_gpu_autograd.py"), "-m", "gpu", "-q", "-rA", "--tb=short",
                "-p", "no:cacheprovider"] + (["-k", args.k] if args.k else [])
        with redirect_stdout(buf), redirect_stderr(buf):
            rc = pytest.main(argv)
        with open(os.path.join(OUT, "autograd_tests.log"), "w") as f:
            f.write(buf.getvalue() + f"\nexit code {int(rc)}\n")
        print(buf.getvalue()[-6000:], f"\npytest exit code {int(rc)}", flush=True)

    import torch
    import hicom_b200
    from hicom_b200 import autograd as ag
    from util import Cfg
    ag.enable(True)
    res = []
    for guide in ("coarse", "direct"):
        torch.manual_seed(0)
        m = hicom_b200.build_vision_projector(Cfg(use_guide=guide, hidden_size=3584, max_num_frames=16))
        with torch.no_grad():
            m.global_compressor.query.normal_(std=0.02)
        m = m.to(torch.bfloat16).cuda().train()
        B = args.videos
        X = (0.5 * torch.randn(B, 16, 27, 27, 1152, device="cuda")).bfloat16()
        E = (0.5 * torch.randn(B, 16, 27, 27, 1152, device="cuda")).bfloat16()
        G = (0.5 * torch.randn(B, 1152, device="cuda")).bfloat16()

        def step():
            m.zero_grad(set_to_none=True)
            out = m.forward_batched(X, E, G, "video")
            out.float().square().mean().backward()

        for _ in range(2):
            step()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(args.iters):
            step()
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / args.iters
        with torch.no_grad():
            for _ in range(2):
                m.forward_batched(X, E, G, "video")
            torch.cuda.synchronize()
            a.record()
            for _ in range(args.iters):
                m.forward_batched(X, E, G, "video")
            b.record()
            torch.cuda.synchronize()
        res.append({"use_guide": guide, "videos": B, "frames": 16 * B, "fwd_bwd_ms": round(ms, 3),
                    "inference_fwd_ms": round(a.elapsed_time(b) / args.iters, 3),
                    "train_frames_per_s": round(16 * B / ms * 1e3, 1)})
        print(json.dumps(res[-1]), flush=True)
    with open(os.path.join(OUT, "autograd_bench.json"), "w") as f:
        f.write(json.dumps({"what": "forward+backward of the drop-in projector, width 3584, 16 frames per video, bf16, "
                                    "backward GEMMs on " + ("the SIMT kernel" if os.environ.get("HICOM_GEMM_TC") == "0"
                                                               else "tcgen05 (large bf16) / SIMT"), "results": res}) + "\n")


if __name__ == "__main__":
    main()
