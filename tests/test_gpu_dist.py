"""GPU, >= 2 devices: the frame-sharded path over NCCL (``dist.forward_frame_sharded``) against the unsharded
forward — tokens of every rank, eager and as a captured CUDA graph.  Skipped on a one-GPU box (the single-GPU shard
identity is in test_gpu_fullsize.py / test_gpu_parity.py, the host logic in test_dist_cpu.py with gloo)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_frame_sharded_over_nccl_matches_unsharded(built_library):
    n = min(torch.cuda.device_count(), 8)
    n = 8 if n >= 8 else (4 if n >= 4 else 2)
    here = os.path.dirname(os.path.abspath(__file__))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", "29611", os.path.join(here, "dist_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("DIST_OK") == n, r.stdout[-3000:]
