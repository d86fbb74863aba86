"""spp split across GPUs: host side of the multi-GPU path, on top of crv_group_* (csrc/group.cu).

The reference is single-GPU; (pixel, frame) samples are independent (the RNG state depends only on the frame seed and the
pixel index, PathTracer.comp:368), so the natural shard is the FRAME list: the density grid and the shadow volume are
replicated, member r renders frames {f : f mod world = r} into a private fp32 SUM image (CRV_ACCUM_SUM), and one kernel per
member -- crv_group_all_reduce_sum_scale, peer loads and stores over NVLink -- adds the members' images in rank order,
divides by the total sample count and leaves the same final picture in every member's image.  No other exchange exists on
the path.  torch.distributed is only the messenger that carries the 96-byte memory handles between the ranks at set-up
(any backend: gloo in the CPU tests, nccl under torchrun).

    split = SampleSplit(renderer, rank, world)          # one per process; collective: exchanges the handles
    split.render_picture(all_seeds, stream)             # clear + my frames + reduce/scale, asynchronous
    picture = split.image.download(stream)              # identical on every member
"""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np

from . import capi
from .technique import Group


def frames_for_rank(all_seeds, rank: int, world: int) -> np.ndarray:
    """Seeds of the frames rank `rank` renders: frame f of the global list goes to rank f mod world."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    return np.ascontiguousarray(np.asarray(all_seeds, dtype=np.int32)[rank::world])


def frame_indices_for_rank(n_frames: int, rank: int, world: int) -> np.ndarray:
    return np.arange(n_frames, dtype=np.int64)[rank::world]


def all_gather_bytes(payload: bytes, world: int, dist_group=None) -> list:
    """every rank's `payload` (equal lengths), in rank order, through torch.distributed; a single process returns [payload]"""
    if world == 1:
        return [payload]
    import torch.distributed as dist
    out = [None] * world
    dist.all_gather_object(out, payload, group=dist_group)
    if any(len(b) != len(payload) for b in out):
        raise RuntimeError("all_gather_bytes: ranks sent payloads of different sizes")
    return out


def broadcast_bytes(payload: Optional[bytes], src: int, world: int, dist_group=None) -> bytes:
    if world == 1:
        return payload
    import torch.distributed as dist
    box = [payload]
    dist.broadcast_object_list(box, src=src, group=dist_group)
    return box[0]


def connect_group(group: Group, dist_group=None) -> None:
    """collective: publish this member's handle, map everybody else's"""
    group.connect(all_gather_bytes(group.export(), group.nranks, dist_group))


def init_group_nccl(group: Group, dist_group=None) -> None:
    """collective: rank 0 draws the ncclUniqueId, everybody joins the communicator behind crv_group_nccl_*"""
    uid = broadcast_bytes(Group.nccl_unique_id() if group.rank == 0 else None, 0, group.nranks, dist_group)
    group.nccl_init(uid)


class SampleSplit:
    """One member of the split around a HeadlessRenderer: re-targets the renderer's path tracer at the group's image and
    renders pictures whose frames are spread over the members."""

    def __init__(self, renderer, rank: int, world: int, nbuf: int = 1, dist_group=None, connect: bool = True, staging_bytes: int = 0):
        self.r = renderer
        self.rank, self.world = int(rank), int(world)
        img = renderer.resultImage
        self.group = Group(renderer.ctx, rank, world, img.width, img.height, nbuf, staging_bytes)
        if connect:
            connect_group(self.group, dist_group)
        self.use_buffer(0)

    def use_buffer(self, k: int):
        self.buffer = int(k)
        self.image = self.group.image(k)
        self.r.resultImage = self.image
        self.r.pathTracingTechnique.SetFrameReferences([self.image])

    def my_seeds(self, all_seeds) -> np.ndarray:
        return frames_for_rank(all_seeds, self.rank, self.world)

    def render_picture(self, all_seeds, stream=0, backend: str = "peer"):
        """clear, this member's frames as a plain sum, combine + divide by len(all_seeds): every member ends with the mean.
        backend: "push" (one-barrier combine: the fold stores the tracked sums into the members' inboxes; needs an even split,
        otherwise "peer"), "peer" (the group's combine kernel; only the tracked pixels travel when the frames split evenly),
        "peer_dense" (whole images), "nccl" (ncclAllReduce + scale, for comparison)"""
        seeds = self.my_seeds(all_seeds)
        if backend == "push" and len(seeds) > 0 and len(seeds) * self.world == len(all_seeds):
            # replicas of one scene, even split: tracker, fold that pushes the tracked sums to every member, one wait, one local kernel
            pt = self.r.pathTracingTechnique
            targets = self.group.push_begin(self.buffer)
            pt.RenderBatchPush(seeds, targets, 1.0 / float(len(seeds)), stream, 0)
            self.group.push_finish(1.0 / float(len(all_seeds)), pt.PixelLists(), stream, self.buffer)
            return
        if backend == "push":
            backend = "peer"
        self.image.clear(stream)
        if len(seeds):
            self.r.pathTracingTechnique.RenderBatch(seeds, 0, capi.ACCUM_SUM, stream, 0)
        self.reduce(len(all_seeds), stream, backend, n_mine=len(seeds))

    def reduce(self, spp_total: int, stream=0, backend: str = "peer", n_mine: int = 0):
        scale = 1.0 / float(spp_total)
        if backend == "peer" and n_mine > 0 and n_mine * self.world == spp_total:
            # replicas of one scene: the same pixel lists everywhere, and every member holds n_mine samples of every pixel
            lists = self.r.pathTracingTechnique.PixelLists()
            self.group.all_reduce_lists_sum_scale(scale, 1.0 / float(n_mine), lists, stream, self.buffer)
        elif backend in ("peer", "peer_dense"):
            self.group.all_reduce_sum_scale(scale, stream, self.buffer)
        elif backend == "nccl":
            self.group.nccl_all_reduce_sum_scale(scale, stream, self.buffer)
        else:
            raise ValueError(f"unknown reduce backend {backend!r}")

    def close(self):
        self.group.close()
