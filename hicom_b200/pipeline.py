"""Host-buffer entry: run the compressor on videos that live in (pinned) host memory.

``compress_from_host`` is the end-to-end call a serving process makes when the features arrive on the
host: it cuts the batch into chunks, and double-buffers host->device copies, the compressor, and the
device->host copy of the tokens on two CUDA streams so PCIe transfers overlap the kernels.
"""
from __future__ import annotations

import contextlib
import os
from typing import Optional

import torch


def _numa_nodes_of(cpus):
    """NUMA nodes (sysfs) that own the given CPU ids; [] when sysfs does not say."""
    nodes = set()
    base = "/sys/devices/system/node"
    try:
        for name in os.listdir(base):
            if not name.startswith("node") or not name[4:].isdigit():
                continue
            for part in open(os.path.join(base, name, "cpulist")).read().strip().split(","):
                lo, _, hi = part.partition("-")
                if part and any(int(lo) <= c <= int(hi or lo) for c in cpus):
                    nodes.add(int(name[4:]))
    except Exception:
        return []
    return sorted(nodes)


@contextlib.contextmanager
def host_affinity(device=None, require: bool = False):
    """Pin the calling thread to the CPUs NVML reports as local to ``device`` for the duration of the block.

    Pinned host buffers allocated inside land on the GPU's NUMA node; a buffer on the far socket can cut the
    host->device rate by 2-3x.  Yields a report ``{"pinned": bool, "cpus": n, "numa_nodes": [...], "reason": ...}``:
    without NVML, or when the container's cpuset forbids the call, nothing changes and the report says why —
    ``require=True`` turns that into an error.
    """
    old = None
    report = {"pinned": False, "cpus": len(os.sched_getaffinity(0)), "numa_nodes": _numa_nodes_of(os.sched_getaffinity(0)),
              "reason": None}
    try:
        import pynvml
        index = torch.device(device if device is not None else torch.cuda.current_device()).index or 0
        visible = os.environ.get("CUDA_VISIBLE_DEVICES")
        if visible:
            index = int(visible.split(",")[index])
        pynvml.nvmlInit()
        old = os.sched_getaffinity(0)
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(index))
        now = os.sched_getaffinity(0)
        report.update(pinned=True, cpus=len(now), numa_nodes=_numa_nodes_of(now))
    except Exception as exc:
        old = None
        report["reason"] = repr(exc)[:160]
        if require:
            raise RuntimeError(f"host_affinity: cannot pin to the CPUs local to {device}: {exc!r}") from exc
    try:
        yield report
    finally:
        if old is not None:
            try:
                os.sched_setaffinity(0, old)
            except Exception:
                pass


# Device staging buffers (two slots per input, one per copy/compute stream), kept across calls: a serving process
# streams many batches of the same shape, and re-allocating 200 MB blocks per chunk makes the caching allocator
# stall for hundreds of milliseconds during its first dozen calls.
_STAGING = {}


def release_staging() -> None:
    """Drop the cached device staging buffers and streams of ``compress_from_host``."""
    _STAGING.clear()


def _staging(device, chunk, tensors):
    key = (str(device), chunk) + tuple((tuple(t.shape[1:]), t.dtype) if t is not None else None for t in tensors)
    st = _STAGING.get(key)
    if st is None:
        st = {"streams": [torch.cuda.Stream(device), torch.cuda.Stream(device)],
              "slots": [[None if t is None else torch.empty((chunk,) + tuple(t.shape[1:]), dtype=t.dtype, device=device)
                         for t in tensors] for _ in range(2)]}
        _STAGING[key] = st
    return st


@torch.no_grad()
def compress_from_host(projector, frames_feature: torch.Tensor, frames_embed: Optional[torch.Tensor],
                       guide_embed: Optional[torch.Tensor], modal: str = "video", out: Optional[torch.Tensor] = None,
                       chunk: int = 4, device=None) -> torch.Tensor:
    """``frames_feature`` (B,T,H,W,d) [+ ``frames_embed``] and ``guide_embed`` (B,…) on the HOST (pinned for
    async copies) -> tokens (B, n_tokens, Dh) on the host (``out`` if given, else a new pinned tensor)."""
    device = torch.device(device if device is not None else torch.cuda.current_device())
    B = frames_feature.shape[0]
    chunk = max(1, min(chunk, B))
    host = (frames_feature, frames_embed, guide_embed)
    st = _staging(device, chunk, host)
    streams = st["streams"]
    main = torch.cuda.current_stream(device)
    for s in streams:
        s.wait_stream(main)
    for i, b0 in enumerate(range(0, B, chunk)):
        b1 = min(B, b0 + chunk)
        s = streams[i % 2]
        with torch.cuda.stream(s):
            # stream order protects the slot: this copy is queued behind the compute that last read it
            x, e, g = (None if h is None else d[:b1 - b0].copy_(h[b0:b1], non_blocking=True)
                       for h, d in zip(host, st["slots"][i % 2]))
            tok = projector.forward_batched(x, e, g, modal)
            if out is None:
                out = torch.empty((B,) + tuple(tok.shape[1:]), dtype=tok.dtype).pin_memory()
            out[b0:b1].copy_(tok, non_blocking=True)
    for s in streams:
        main.wait_stream(s)
    main.synchronize()
    return out
