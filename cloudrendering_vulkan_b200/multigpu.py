"""spp split across GPUs: one process per GPU, grid + shadow volume replicated, fp32 accumulation buffers summed.

The reference is single-GPU; (pixel, frame) samples are independent (RNG state depends only on the frame seed and the
pixel index, PathTracer.comp:368), so the natural shard is the FRAME list: rank r renders frames {f : f mod world = r}
into a local fp32 SUM (CRV_ACCUM_SUM), one all-reduce (NCCL over NVLink on GPUs, gloo in the CPU tests) adds the
partial sums, and a divide by the total spp turns the sum into the mean the single-GPU running mean converges to.
No other collective exists on the path.
"""
from __future__ import annotations

import numpy as np


def frames_for_rank(all_seeds, rank: int, world: int) -> np.ndarray:
    """Seeds of the frames rank `rank` renders: frame f of the global list goes to rank f mod world."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    return np.ascontiguousarray(np.asarray(all_seeds, dtype=np.int32)[rank::world])


def frame_indices_for_rank(n_frames: int, rank: int, world: int) -> np.ndarray:
    return np.arange(n_frames, dtype=np.int64)[rank::world]


def reduce_accumulation(accum, spp_total: int, group=None, scale_fn=None):
    """In-place all-reduce(sum) of a torch accumulation tensor (any device) followed by the divide by spp_total.

    scale_fn(factor), when given, performs the divide on the device through the C ABI (crv_image_scale) instead of a
    torch op; returns `accum`.
    """
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(accum, op=dist.ReduceOp.SUM, group=group)
    if scale_fn is not None:
        scale_fn(1.0 / float(spp_total))
    else:
        accum.mul_(1.0 / float(spp_total))
    return accum
