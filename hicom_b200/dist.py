"""Multi-GPU partitioning of the compressor across one 8xB200 box.

Two ways the path shards, one process per GPU (torch.distributed, NCCL over NVLink/NVSwitch):

* by VIDEO — videos are independent (the reference loops over them, hicom_arch.py:167-178):
  ``video_shard`` gives each rank a contiguous block; no data-path collective at all.
* by FRAME for one long video — rank r holds frames [t0, t0+Ts), Ts % temporal_kernel == 0.
  Local path: windows never cross a 4-frame boundary and the grid-pool taps stay inside the window,
  so there is no halo and no communication.  Global path: every rank computes split-softmax partials
  (m, l, o) over its frames (position rows offset by t0), normalises them on its own and applies the
  per-head value projection (linear, so it commutes with the merge); ONE all-gather exchanges the
  resulting 32 x d attention rows + the log-sum-exps of the scores (75 KB per video; the raw fp32
  partial would be 1.3 MB), and every rank combines them with softmax weights of the log-sum-exps
  (``ops.shard_combine``) and finishes the 32 query rows (replicated — cheaper than a second exchange).
"""
from __future__ import annotations

import contextlib
from typing import Tuple

import torch
import torch.distributed as dist


SHARD_OVERLAP = True


def video_shard(num_videos: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous [begin, end) block of videos for ``rank`` (sizes differ by at most one)."""
    base, extra = divmod(num_videos, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def frame_shard(num_frames: int, world: int, rank: int, temporal_kernel: int = 4) -> Tuple[int, int]:
    """Contiguous [t0, t1) block of frames for ``rank``; blocks are multiples of the temporal kernel so
    local windows never straddle ranks.  Raises if the video cannot be cut that way."""
    if num_frames % temporal_kernel:
        raise ValueError(f"{num_frames} frames are not a multiple of the temporal kernel {temporal_kernel}")
    groups = num_frames // temporal_kernel
    b, e = video_shard(groups, world, rank)
    return b * temporal_kernel, e * temporal_kernel


def pack_partials(m: torch.Tensor, l: torch.Tensor, o: torch.Tensor) -> torch.Tensor:
    """(B,S,J),(B,S,J),(B,S,J,d) -> one (B,S,J,d+2) fp32 message [o | m | l]."""
    return torch.cat([o, m.unsqueeze(-1), l.unsqueeze(-1)], dim=-1).contiguous()


def unpack_partials(buf: torch.Tensor):
    d = buf.shape[-1] - 2
    return buf[..., d].contiguous(), buf[..., d + 1].contiguous(), buf[..., :d].contiguous()


def gather_partials(m: torch.Tensor, l: torch.Tensor, o: torch.Tensor, group=None):
    """All-gather the split-softmax partials of every rank; returns (m,l,o) with the split axis grown to
    world*S, rank-major (any order would do: the merge is order-independent up to fp32 rounding)."""
    world = dist.get_world_size(group)
    if world == 1:
        return m, l, o
    msg = pack_partials(m, l, o)
    B, S, J, d2 = msg.shape
    out = torch.empty((world * B, S, J, d2), dtype=msg.dtype, device=msg.device)  # concatenated along dim 0
    dist.all_gather_into_tensor(out, msg, group=group)
    out = out.view(world, B, S, J, d2).permute(1, 0, 2, 3, 4).reshape(B, world * S, J, d2)
    return unpack_partials(out)


def gather_messages(msg: torch.Tensor, group=None) -> torch.Tensor:
    """All-gather every rank's (B, nbytes) frame-shard message -> (world, B, nbytes), rank-major."""
    world = dist.get_world_size(group)
    out = torch.empty((world,) + tuple(msg.shape), dtype=msg.dtype, device=msg.device)
    dist.all_gather_into_tensor(out.view(world * msg.shape[0], msg.shape[1]), msg.contiguous(), group=group)
    return out


@torch.no_grad()
def forward_frame_sharded(projector, frames_feature, frames_embed, guide_embed, t0: int, group=None,
                          modal: str = "video"):
    """One long video cut by frames: ``frames_feature`` (B,Ts,H,W,d) is THIS rank's block starting at global
    frame ``t0``.  Returns (local tokens of this block (B, Nw_block, Dh) | None, global tokens (B,Q,Dh) | None).
    Concatenating the local blocks in rank order and appending the global tokens reproduces
    ``forward_batched`` on the whole video (newline layouts other than flat/no_token are not sharded).
    Inference only (decorated ``no_grad``): the exchange of partials between ranks is not differentiated — training
    shards by video, where every rank runs the ordinary ``forward`` / ``forward_batched`` under DDP."""
    X = frames_feature
    B, Ts, H, W, d = X.shape
    lc, gc = projector.local_compressor, projector.global_compressor
    local_tokens = global_tokens = None
    # the local chain runs on the side stream, as in forward_batched, so that it overlaps the global chain and its exchange
    side = None
    if lc is not None and gc is not None and X.is_cuda and SHARD_OVERLAP:
        from .projector import _side_stream
        main = torch.cuda.current_stream(X.device)
        side = _side_stream(X.device)
        side.wait_stream(main)
    if lc is not None:
        if Ts % lc.temporal_kernel_size:
            raise ValueError("frame shard must be a multiple of the temporal kernel")
        with torch.cuda.stream(side) if side is not None else contextlib.nullcontext():
            att = lc.attend(X, frames_embed, guide_embed, modal, projector.local_logit_scale,
                            projector.local_logit_bias)
            Dh = lc.readout[-1].out_features
            local_tokens = torch.empty((B * att.shape[1], Dh), dtype=X.dtype, device=X.device)
            from .projector import _mlp_into
            _mlp_into(lc.readout, att, local_tokens, 0, att.shape[1], att.shape[1])
            local_tokens = local_tokens.view(B, att.shape[1], Dh)
    if gc is not None:
        Qg = gc.injected_query(guide_embed, B, X.dtype)
        m, l, o = gc.partials(X, gc.fold(Qg, projector.global_logit_scale), t0=t0,
                              logit_scale=projector.global_logit_scale)
        Dh = gc.readout[-1].out_features
        nq = gc.query.shape[0]  # Qg may hold one row per video (direct mode); finish() replicates it
        global_tokens = torch.empty((B * nq, Dh), dtype=X.dtype, device=X.device)
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            from . import ops
            if X.dtype == torch.float32:  # fp32 mode: exchange the raw partials (exact to fp32 rounding)
                if m.shape[1] > 1:  # reduce this rank's token splits first: one J*(d+2) fp32 message per video
                    m, l, o = ops.softmax_reduce(m, l, o)
                m, l, o = gather_partials(m, l, o, group)
                gc.finish(Qg, m, l, o, global_tokens, 0, nq)
            else:
                msgs = gather_messages(gc.shard_message(Qg, m, l, o), group)
                a = ops.shard_combine(msgs, Qg.shape[1], d, gc.attn_layer.num_heads, X.dtype)
                gc.finish_attended(Qg, a, global_tokens, 0, nq)
        else:
            gc.finish(Qg, m, l, o, global_tokens, 0, nq)
        global_tokens = global_tokens.view(B, nq, Dh)
    if side is not None:
        main.wait_stream(side)
        local_tokens.record_stream(main)
    return local_tokens, global_tokens
