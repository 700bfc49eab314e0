"""One-process GPU check of the producer side: the parity tests of tests/test_gpu_producer.py, then a
CUDA-event timing of frames_embed = h + mlp(layernorm(h)) at the c2 scale (512 frames x 729 tokens, bf16).  Writes
gpurun_out/producer_tests.log and gpurun_out/producer_bench.json as it goes (a cut-off run keeps what it finished).

    python tools/producer_round.py [--frames 512] [--skip-tests]
"""
import argparse
import io
import json
import os
import sys
from contextlib import redirect_stderr, redirect_stdout

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "gpurun_out")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=512)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--skip-tests", action="store_true")
    args = ap.parse_args()
    os.makedirs(OUT, exist_ok=True)

    if not args.skip_tests:
        import pytest
        buf = io.StringIO()
        with redirect_stdout(buf), redirect_stderr(buf):
            rc = pytest.main([os.path.join(ROOT, "tests", "test_gpu_producer.py"), "-m", "gpu", "-q", "-rA",
                              "--tb=short", "-p", "no:cacheprovider"])
        with open(os.path.join(OUT, "producer_tests.log"), "w") as f:
            f.write(buf.getvalue() + f"\nexit code {int(rc)}\n")
        print(buf.getvalue()[-3000:], f"\npytest exit code {int(rc)}", flush=True)

    import torch
    from hicom_b200 import ops
    from hicom_b200.producer import SiglipHeadEmbed
    from oracle import siglip_head as SH

    dt = torch.bfloat16
    m = SiglipHeadEmbed()
    m.load_state_dict(SH.synth_head_state(0), strict=True)
    m = m.to(dt).cuda().eval()
    h = (0.7 * torch.randn(args.frames, 729, 1152, device="cuda")).to(dt)
    flops = 2.0 * 2 * args.frames * 729 * 1152 * 4304
    with torch.no_grad():
        for _ in range(3):
            m(h)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(args.iters):
            m(h)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / args.iters
        with ops.OpTimer() as t:
            for _ in range(args.iters):
                m(h)
        per_op = {k: round(v[1] / args.iters, 4) for k, v in t.summary().items()}
    line = {"what": "producer: frames_embed = h + head.mlp(head.layernorm(h)) (encoder.py:284-286)", "dtype": "bf16",
            "frames": args.frames, "ms": round(ms, 4), "frames_per_s": round(args.frames / ms * 1e3, 1),
            "tflops": round(flops / ms / 1e9, 1), "ms_per_op": per_op,
            "note": "inputs resident in HBM (860 MB for 512 frames > L2); CUDA events, 3 warm-up calls"}
    with open(os.path.join(OUT, "producer_bench.json"), "w") as f:
        f.write(json.dumps(line) + "\n")
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
