"""Multi-GPU frame rendering: one process per GPU, interleaved image tiles per rank, one merge.

SURVEY.md §8(e): pixels are independent (lib.rs:63), so the frame shards with no exchange while
rendering; the only collective is the merge of the accumulation buffers at the end of the frame.
Every rank holds a full scene replica, renders the tiles ``k % world == rank`` (rtw_render_params
part_rank / part_count) into a zero-padded full-size buffer, and ``reduce(SUM)`` over NCCL (NVLink 5 /
NVSwitch) lands the frame on rank 0.  Because the random stream is keyed by (pixel, sample) and each
pixel is summed by exactly one rank, the merged image is bit-identical for every world size.

The function is backend-agnostic (it only needs ``scene.render_into(buffer)``), so the host-side
logic is covered on CPU with the gloo backend (tests/test_dist.py).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def partition(params, rank: int, world: int, tile_size: int = 32):
    """Return a copy of `params` restricted to this rank's tiles."""
    import copy

    p = copy.copy(params)
    p.tile_size = tile_size
    p.part_rank = rank
    p.part_count = world
    return p


def render_frame(render_into, params, accum: torch.Tensor, *, group=None, dst: int = 0, tile_size: int = 32):
    """Render this rank's share of the frame into `accum` (float32, h*w*3) and merge on `dst`.

    render_into(params, accum) -> stats   renders with the given (already partitioned) params into the
    tensor's memory (device memory for the CUDA backend, host memory for the CPU harness).
    Returns this rank's stats.
    """
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    p = partition(params, rank, world, tile_size) if world > 1 else params
    stats = render_into(p, accum)
    if world > 1:
        # pixels of other ranks are exact zeros in `accum`; x + 0 = x, so SUM is a gather
        dist.reduce(accum, dst=dst, op=dist.ReduceOp.SUM, group=group)
    return stats
