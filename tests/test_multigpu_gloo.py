"""N > 1 path on the CPU: two gloo ranks split the frame list, each accumulates a local fp32 SUM (computed by the
oracle here -- no GPU in this container), the sums are added in rank order and divided by spp_total -- the arithmetic of
crv_group_all_reduce_sum_scale (csrc/group.cu; its device side is tests/test_group.py).  Checks the host logic of
cloudrendering_vulkan_b200.multigpu (frame partition, the handle exchange between ranks) and the <= 1e-6 agreement with the
1-rank result."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _scene():
    sys.path.insert(0, ROOT)
    import cloudrendering_vulkan_b200 as crv
    from tests import oracle_binding as orc
    n, w, h, nsv = 16, 32, 24, 20
    cloud = orc.fbm_generate(n, 1234)
    cam = crv.CameraProperties(); cam.SetResolution(w, h); cam.SetFOV(90.0)
    cp = crv.CloudProperties(); cp.densityScaling = 120.0
    cp.SetFromGrid((n, n, n), (1.0 / n,) * 3, float(cloud.max()))
    par = crv.Parameters(); par.SetPhaseG(0.85)
    sp = crv.ShadowVolumeProperties(); sp.voxelAxisCount = nsv
    sp.SetLightDirection((1.0, -1.0, 0.0))
    sp.SetOrigin(tuple(cp.bounds[0])[:3], tuple(cp.bounds[1])[:3])
    sv = orc.shadow_volume(cloud, cp, sp, threads=1)
    return crv, orc, (w, h), cloud, sv, cam, cp, par, sp


def _local_sum(orc, scene, seeds):
    (w, h), cloud, sv, cam, cp, par, sp = scene
    acc = np.zeros((h, w, 4), np.float32)
    for s in seeds:
        orc.path_trace_frame(cloud, sv, cam, cp, par, sp, orc.push_constants_bytes(int(s), 0), 0.1, acc, accum_sum=1,
                             want_sample=False, threads=1)
    return acc


def _worker(rank, world, port, spp, out_dir):
    import torch
    import torch.distributed as dist
    crv, orc, size, cloud, sv, cam, cp, par, sp = _scene()
    from cloudrendering_vulkan_b200.multigpu import all_gather_bytes, broadcast_bytes, frames_for_rank
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # the messenger that carries the members' 96-byte memory handles (and the 128-byte NCCL id) between the ranks
    mine = bytes([rank]) * crv.capi.C.sizeof(crv.capi.GroupHandle)
    got = all_gather_bytes(mine, world)
    assert [b[0] for b in got] == list(range(world)) and all(len(b) == 96 for b in got)
    assert broadcast_bytes(b"x" * 128 if rank == 0 else None, 0, world) == b"x" * 128
    seeds = frames_for_rank(crv.msvc_rand_sequence(spp), rank, world)
    acc = _local_sum(orc, (size, cloud, sv, cam, cp, par, sp), seeds)
    # what the combine kernel computes: members' sums added in rank order, times 1 / spp_total
    parts = [torch.zeros(acc.shape, dtype=torch.float32) for _ in range(world)]
    dist.all_gather(parts, torch.from_numpy(acc))
    total = parts[0].numpy().copy()
    for p in parts[1:]:
        total = total + p.numpy()
    total = total * np.float32(1.0 / spp)
    if rank == 0:
        np.save(os.path.join(out_dir, "reduced.npy"), total)
    dist.barrier()
    dist.destroy_process_group()


def test_frame_partition():
    sys.path.insert(0, ROOT)
    from cloudrendering_vulkan_b200.multigpu import frame_indices_for_rank, frames_for_rank
    seeds = np.arange(100, 111, dtype=np.int32)
    for world in (1, 2, 3, 4, 8):
        parts = [frames_for_rank(seeds, r, world) for r in range(world)]
        assert sorted(np.concatenate(parts).tolist()) == seeds.tolist()           # disjoint cover
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1       # balanced
        assert all(np.array_equal(seeds[frame_indices_for_rank(len(seeds), r, world)], parts[r]) for r in range(world))
    with pytest.raises(ValueError):
        frames_for_rank(seeds, 2, 2)


def test_two_rank_reduce_matches_single_rank(tmp_path):
    import torch.multiprocessing as mp
    spp, world = 6, 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, spp, str(tmp_path)), nprocs=world, join=True)
    reduced = np.load(tmp_path / "reduced.npy")
    crv, orc, size, cloud, sv, cam, cp, par, sp = _scene()
    single = _local_sum(orc, (size, cloud, sv, cam, cp, par, sp), crv.msvc_rand_sequence(spp)) / np.float32(spp)
    assert np.allclose(reduced, single, rtol=1e-6, atol=1e-7)
    # and the sum-then-divide mean agrees with the reference's running mean to rounding
    mean = np.zeros_like(single)
    for f, s in enumerate(crv.msvc_rand_sequence(spp)):
        orc.path_trace_frame(cloud, sv, cam, cp, par, sp, orc.push_constants_bytes(int(s), f), 0.1, mean, want_sample=False, threads=1)
    assert np.allclose(reduced, mean, rtol=1e-5, atol=1e-7)
