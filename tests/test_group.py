"""Multi-GPU sample split through the C ABI (crv_group_*, csrc/group.cu).

The driver's GPU box has one B200, so the protocol -- handle exchange, start / end barriers through flags in mapped memory,
band reduce over all members' images, epochs, double buffering, the timeout -- is exercised with several members that share
device 0 (each on its own stream; their kernels wait for each other on the device exactly as they would across NVLink).
The same tests run with one member per GPU when more GPUs are visible (`gpurun --gpus 2`), NCCL comparison included.

Bar (SURVEY 8d "Multi-GPU"): N-member picture vs 1-GPU picture over the same seed list <= 1e-6 relative (summation order
only); here also: every member ends with the same bits.
"""
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def _devices(crv, n, distinct):
    have = crv.capi.lib.crv_device_count()
    if distinct:
        if have < n:
            pytest.skip(f"needs {n} GPUs, {have} visible")
        return list(range(n))
    return [0] * n


def _members(crv, n, distinct, w=96, h=64, ncloud=40, nsv=56, nbuf=1, **kw):
    """n replicas of one scene, one group member each, connected"""
    wl = crv.Workload("t", ncloud, w, h, 1, nsv, kw.get("density_scaling", 50.0), kw.get("absorption", 0.1), kw.get("g", 0.85))
    ms = []
    for rank, dev in enumerate(_devices(crv, n, distinct)):
        ctx = crv.Context(dev)
        r = crv.make_renderer(ctx, wl, cloud_layout="brick")
        r.set_cloud_synthetic(ncloud, 1234)
        ctx.synchronize()
        from cloudrendering_vulkan_b200.multigpu import SampleSplit
        own_image = r.resultImage
        split = SampleSplit(r, rank, n, nbuf=nbuf, connect=False)
        own_image.close()
        ms.append((ctx, r, split, crv.Stream(ctx)))
    handles = [m[2].group.export() for m in ms]
    for m in ms:
        m[2].group.connect(handles)
    return ms


def _close(ms):
    for ctx, r, split, stream in ms:
        ctx.synchronize()
        ctx.synchronize(int(stream))
    for ctx, r, split, stream in ms:
        stream.close()
        r.close()                # its result image is the group's (borrowed): closed with the group
        split.close()
        ctx.close()


def _single_gpu_picture(crv, seeds, mode, w=96, h=64, ncloud=40, nsv=56, **kw):
    ctx = crv.Context(0)
    wl = crv.Workload("t", ncloud, w, h, 1, nsv, kw.get("density_scaling", 50.0), kw.get("absorption", 0.1), kw.get("g", 0.85))
    r = crv.make_renderer(ctx, wl, cloud_layout="brick")
    r.set_cloud_synthetic(ncloud, 1234)
    r.pathTracingTechnique.RenderBatch(seeds, 0, mode, 0, 0)
    if mode == crv.ACCUM_SUM:
        r.resultImage.scale(1.0 / len(seeds))
    img = r.resultImage.download().copy()
    r.close()
    ctx.close()
    return img


@pytest.mark.parametrize("spp", [13, 12])
@pytest.mark.parametrize("n,distinct", [(2, False), (4, False), (3, False), (2, True), (4, True), (8, True)])
def test_split_picture_equals_single_gpu_picture(crv, n, distinct, spp):
    # 13: not a multiple of n, members hold different numbers of frames -> whole images are combined;
    # 12: even split -> only the tracked pixels cross between the members (crv_group_all_reduce_lists_sum_scale)
    if spp == 12 and n == 8:
        spp = 16
    seeds = crv.msvc_rand_sequence(spp)
    ms = _members(crv, n, distinct)
    for ctx, r, split, stream in ms:             # all renders and combines are enqueued before anybody waits
        split.render_picture(seeds, int(stream))
    imgs = []
    for ctx, r, split, stream in ms:
        imgs.append(split.image.download(int(stream)).copy())
        split.group.status(int(stream))
    for im in imgs[1:]:
        assert np.array_equal(_bits(imgs[0]), _bits(im)), "members hold different pictures"
    ref_sum = _single_gpu_picture(crv, seeds, crv.ACCUM_SUM)
    ref_mean = _single_gpu_picture(crv, seeds, crv.ACCUM_RUNNING_MEAN)
    assert ref_sum.max() > 0.1
    if spp % n == 0:                             # the dense combine of the same frames gives the same picture to rounding
        for ctx, r, split, stream in ms:
            split.render_picture(seeds, int(stream), "peer_dense")
        dense = ms[0][2].image.download(int(ms[0][3])).copy()
        for ctx, r, split, stream in ms[1:]:
            ctx.synchronize(int(stream))
        assert (np.abs(dense - imgs[0]) / np.maximum(np.abs(dense), 1e-3)).max() <= 1e-6
    err = np.abs(imgs[0] - ref_sum) / np.maximum(np.abs(ref_sum), 1e-3)
    assert err.max() <= 1e-6, f"split picture differs from the 1-GPU sum picture: {err.max():.3e}"
    err = np.abs(imgs[0] - ref_mean) / np.maximum(np.abs(ref_mean), 1e-3)
    assert err.max() <= 1e-5, f"split picture differs from the 1-GPU running mean: {err.max():.3e}"
    print(f"n={n} distinct={distinct}: reduce {ms[0][2].group.last_reduce_ms():.3f} ms")
    _close(ms)


def test_epochs_and_double_buffering(crv):
    """many pictures back to back, alternating between two accumulation buffers, no host synchronisation in between: the
    flag epochs keep the members in step and a buffer is never reused while a peer still reads it"""
    n, rounds = 3, 6
    ms = _members(crv, n, False, nbuf=2)
    pics = []
    all_seeds = crv.msvc_rand_sequence(4 * rounds)
    for k in range(rounds):
        for ctx, r, split, stream in ms:
            split.use_buffer(k % 2)
            split.render_picture(all_seeds[4 * k:4 * k + 4], int(stream))
        if k >= rounds - 2:                      # the last two pictures are still in the two buffers at the end
            pics.append(k)
    for ctx, r, split, stream in ms:
        ctx.synchronize(int(stream))
        split.group.status(int(stream))
    for k in pics:
        ref = _single_gpu_picture(crv, all_seeds[4 * k:4 * k + 4], crv.ACCUM_SUM)
        for ctx, r, split, stream in ms:
            got = split.group.image(k % 2).download(int(stream))
            assert np.abs(got - ref).max() <= 1e-6 * max(1.0, float(np.abs(ref).max()))
    _close(ms)


@pytest.mark.parametrize("n,distinct", [(1, False), (2, False), (3, False), (4, False), (2, True), (4, True), (8, True)])
def test_push_combine_equals_lists_combine_bitwise(crv, n, distinct):
    """crv_group_push_begin / crv_pt_render_batch_push / crv_group_push_finish (one barrier, the fold stores the tracked sums
    into the members' inboxes, no clear) against crv_group_all_reduce_lists_sum_scale (two barriers around a band reduce): the
    same arithmetic in the same order, so the same bits, on every member; then pictures back to back without host
    synchronisation (inbox slots alternate per picture), and a render call that needs several tracker launches."""
    spp = 16 if n == 8 else 12
    seeds = crv.msvc_rand_sequence(spp)
    ms = _members(crv, n, distinct)
    for ctx, r, split, stream in ms:
        split.render_picture(seeds, int(stream), "peer")
    lists_pic = []
    for ctx, r, split, stream in ms:
        lists_pic.append(split.image.download(int(stream)).copy())
        split.group.status(int(stream))
    for ctx, r, split, stream in ms:
        split.image.upload(np.full((64, 96, 4), np.nan, np.float32), int(stream))      # the push path must not read the image
        split.render_picture(seeds, int(stream), "push")
    for k, (ctx, r, split, stream) in enumerate(ms):
        got = split.image.download(int(stream)).copy()
        split.group.status(int(stream))
        assert np.array_equal(_bits(got), _bits(lists_pic[0])), f"member {k}: push combine differs from the lists combine"
    ref = _single_gpu_picture(crv, seeds, crv.ACCUM_SUM)
    err = np.abs(lists_pic[0] - ref) / np.maximum(np.abs(ref), 1e-3)
    assert err.max() <= 1e-6
    # five more pictures with different seeds, nobody waits in between; the last one is checked
    more = crv.msvc_rand_sequence(spp * 6)
    for k in range(1, 6):
        for ctx, r, split, stream in ms:
            split.render_picture(more[spp * k:spp * (k + 1)], int(stream), "push")
    last = []
    for ctx, r, split, stream in ms:
        last.append(split.image.download(int(stream)).copy())
        split.group.status(int(stream))
    ref = _single_gpu_picture(crv, more[spp * 5:spp * 6], crv.ACCUM_SUM)
    for got in last:
        assert np.array_equal(_bits(got), _bits(last[0]))
        assert (np.abs(got - ref) / np.maximum(np.abs(ref), 1e-3)).max() <= 1e-6
    # a sample buffer that holds one frame: every render call is spp / n tracker launches, the sums stay in the image until the last
    for ctx, r, split, stream in ms:
        r.pathTracingTechnique.SetTuning(sample_buffer_bytes=1)
        split.render_picture(seeds, int(stream), "push")
    for ctx, r, split, stream in ms:
        got = split.image.download(int(stream)).copy()
        split.group.status(int(stream))
        assert (np.abs(got - lists_pic[0]) / np.maximum(np.abs(lists_pic[0]), 1e-3)).max() <= 1e-6
    _close(ms)


@pytest.mark.parametrize("n,distinct", [(2, False), (3, False), (1, False), (2, True), (4, True)])
def test_staging_all_gather(crv, n, distinct):
    """a replicated host array: every member uploads 1/N and pushes it to its peers on the copy engines; two rounds into the
    same buffer (the second must wait for the peers' release) and a ragged size"""
    import ctypes as C
    elems = 1_000_003
    ctxs = [crv.Context(d) for d in _devices(crv, n, distinct)]
    groups = [crv.Group(c, k, n, 16, 16, nbuf=2, staging_bytes=4 * elems) for k, c in enumerate(ctxs)]
    hs = [g.export() for g in groups]
    for g in groups:
        g.connect(hs)
    streams = [crv.Stream(c) for c in ctxs]
    vols = [crv.Volume(c, elems, 1, 1, layout="linear") for c in ctxs]
    rng = np.random.default_rng(3)
    for rnd in range(3):
        data = rng.random(elems, dtype=np.float32)
        buf = rnd % 2
        for g, s, v in zip(groups, streams, vols):
            g.staging_all_gather(buf, data.ctypes.data, 4 * elems, int(s))
            v.upload_device(g.staging_ptr(buf), int(s))
            g.staging_release(buf, int(s))
        for g, s, v in zip(groups, streams, vols):
            assert np.array_equal(v.download(int(s)), data), f"round {rnd}: member {g.rank} holds another array"
            g.status(int(s))
    for g, s, v, c in zip(groups, streams, vols, ctxs):
        v.close()
        s.close()
        g.close()
        c.close()


def test_single_member_group_is_a_scale(crv, ctx):
    g = crv.Group(ctx, 0, 1, 32, 8)
    img = g.image(0)
    a = np.arange(32 * 8 * 4, dtype=np.float32).reshape(8, 32, 4)
    img.upload(a)
    g.all_reduce_sum_scale(0.25)
    assert np.array_equal(img.download(), a * np.float32(0.25))
    g.nccl_all_reduce_sum_scale(2.0)             # one member: no communicator needed
    assert np.array_equal(img.download(), a * np.float32(0.5))
    g.status()
    g.close()


def test_missing_peer_times_out_and_is_reported(crv, ctx):
    """a member whose peer never calls the combine gives up after the timeout instead of hanging the GPU"""
    g0, g1 = crv.Group(ctx, 0, 2, 16, 16), crv.Group(ctx, 1, 2, 16, 16)
    hs = [g0.export(), g1.export()]
    g0.connect(hs)
    g1.connect(hs)
    g0.set_timeout_ms(30)
    s = crv.Stream(ctx)
    g0.all_reduce_sum_scale(1.0, int(s))
    with pytest.raises(crv.CrvError) as e:
        g0.status(int(s))
    assert "gave up waiting" in e.value.message
    s.close()
    g0.close()
    g1.close()


def test_group_argument_errors(crv, ctx):
    with pytest.raises(crv.CrvError):
        crv.Group(ctx, 2, 2, 16, 16)             # rank outside the group
    with pytest.raises(crv.CrvError):
        crv.Group(ctx, 0, 17, 16, 16)            # more members than supported
    g = crv.Group(ctx, 0, 2, 16, 16)
    with pytest.raises(crv.CrvError):
        g.all_reduce_sum_scale(1.0)              # not connected yet
    other = crv.Group(ctx, 1, 2, 32, 16)         # another image size
    with pytest.raises(crv.CrvError):
        g.connect([g.export(), other.export()])
    with pytest.raises(crv.CrvError):
        g.image(3)
    other.close()
    g.close()


def test_nccl_route_matches_peer_kernel(crv):
    """north_star's wording: an NCCL reduce of the accumulation buffers -- same picture within summation order (needs 2 GPUs)"""
    n = 2
    seeds = crv.msvc_rand_sequence(8)
    ms = _members(crv, n, True)
    uid = crv.Group.nccl_unique_id()
    ts = [threading.Thread(target=m[2].group.nccl_init, args=(uid,)) for m in ms]      # ncclCommInitRank blocks until all joined
    [t.start() for t in ts]
    [t.join() for t in ts]
    out = {}
    for backend in ("peer", "nccl"):
        ts = [threading.Thread(target=m[2].render_picture, args=(seeds, int(m[3]), backend)) for m in ms]
        [t.start() for t in ts]
        [t.join() for t in ts]
        out[backend] = [m[2].image.download(int(m[3])).copy() for m in ms]
        print(backend, "reduce ms", ms[0][2].group.last_reduce_ms())
    assert np.array_equal(_bits(out["peer"][0]), _bits(out["peer"][1]))
    err = np.abs(out["nccl"][0] - out["peer"][0]) / np.maximum(np.abs(out["peer"][0]), 1e-3)
    assert err.max() <= 1e-6
    _close(ms)
