"""Frame sharding of the operator across ranks.

Output rows of frame n depend only on value[n], sampling_locations[n], attention_weights[n]
(reference: the batch offset b_col*S*M*D, ms_deform_im2col_cuda.cuh:363), and so do all three
gradients, so N frames are cut into contiguous per-rank ranges with NO collective in forward or
backward.  One process per GPU; torch.distributed is only used by callers that want to reassemble
the full output (verification, evaluation)."""
from __future__ import annotations

from typing import Tuple

import torch


def frame_range(n_frames: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous, balanced [start, stop) of frames owned by `rank`; the first n_frames % world_size
    ranks get one extra frame.  Legal for any n_frames (the reference's 64-frame chunk rule,
    ms_deform_attn_cuda.cu:54-56, does not constrain the new kernels)."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError(f"bad rank/world_size {rank}/{world_size}")
    base, extra = divmod(int(n_frames), world_size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_frames(rank: int, world_size: int, *tensors: torch.Tensor):
    """Views of the rank's frames of every tensor (all share dim 0 = frames)."""
    n = tensors[0].shape[0]
    for t in tensors:
        if t.shape[0] != n:
            raise ValueError("all tensors must have the same number of frames in dim 0")
    a, b = frame_range(n, rank, world_size)
    return tuple(t[a:b] for t in tensors)


def gather_frames(local: torch.Tensor, n_frames: int, group=None) -> torch.Tensor:
    """Reassemble a frame-sharded tensor on every rank (verification / evaluation only; not on the
    operator's data path)."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    sizes = [b - a for a, b in (frame_range(n_frames, r, world) for r in range(world))]
    biggest = max(sizes)
    # all_gather wants equal shapes on every rank: pad the short shards, trim after
    padded = local.new_zeros((biggest,) + tuple(local.shape[1:]))
    padded[: local.shape[0]] = local
    parts = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(parts, padded, group=group)
    return torch.cat([p[:n] for p, n in zip(parts, sizes)], 0)
