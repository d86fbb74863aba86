"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on the same seeded inputs.

Bars (BASELINE.json north_star / SURVEY.md 8d):
  * RNG: exact u32 equality
  * shadow volume: <= 1e-5 relative for T >= 1e-30 (absolute 1e-35 below)
  * per-sample radiance: <= 1e-4 relative (floor 1e-3) on pixels whose (steps, collisions) counters match the
    oracle's; pixels whose counters differ are "threshold-divergent" (libm vs CUDA log/sin/cos/pow differ by
    <= 2 ulp and a compare flipped) -- counted, reported, bounded, and covered by the statistical test
  * converged image: |mean_gpu - mean_cpu| <= 5 sigma / sqrt(spp)
  * everything that does not involve a transcendental difference is bit-exact: layouts, scheduling knobs,
    batch-vs-sequential accumulation.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

LAYOUTS = ["linear", "brick", "tex3d"]


def _bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def make_scene(crv, orc, ctx, n=40, w=96, h=64, nsv=56, layout="linear", g=0.85, absorption=0.0, density_scaling=50.0,
               light=(1.0, -1.0, 0.0), cloud=None):
    wl = crv.Workload("t", n, w, h, 1, nsv, density_scaling, absorption, g, light)
    r = crv.make_renderer(ctx, wl, cloud_layout=layout)
    if cloud is None:
        cloud = orc.fbm_generate(n, 1234)
    r.set_cloud_host(cloud, (n, n, n), (1.0 / n,) * 3)
    return r, cloud


# ---------------------------------------------------------------------------------------------------------
def test_rng_tap_bit_exact(crv, orc, ctx):
    for (w, h) in [(64, 64), (96, 40), (1920, 4)]:
        for seed in [38, 7719, 100, -5, 0]:
            gpu = crv.rng_tap(ctx, seed, w, h, 8)
            cpu = orc.rng_tap(seed, w, h, 8)
            assert np.array_equal(gpu, cpu), f"RNG stream differs for seed {seed} at {w}x{h}"


def test_tracker_log_equals_logf_exhaustively(crv, ctx):
    """the tracker's trimmed -log(zeta) gives the same bits as -logf(zeta) for every float in {0} U [2^-32, 1]"""
    import ctypes as C
    n = C.c_uint64(123)
    crv.capi.check(crv.capi.lib.crv_pt_selftest_log(ctx.handle, C.byref(n), None))
    assert n.value == 0, f"{n.value} inputs differ from logf"


@pytest.mark.parametrize("layout", LAYOUTS)
def test_volume_roundtrip_and_majorant(crv, ctx, layout):
    rng = np.random.default_rng(7)
    for shape in [(16, 16, 16), (13, 7, 21), (40, 9, 8), (1, 1, 1)]:
        a = rng.random(shape[0] * shape[1] * shape[2], dtype=np.float32)
        v = crv.Volume(ctx, *shape, layout=layout)
        v.upload(a)
        assert np.array_equal(v.download(), a)
        assert v.majorant() == a.max()
        v.close()


@pytest.mark.parametrize("layout", LAYOUTS)
def test_fbm_device_generator_matches_oracle(crv, orc, ctx, layout):
    n = 48
    v = crv.Volume(ctx, n, n, n, layout=layout)
    v.fbm_generate(1234)
    gpu = v.download()
    cpu = orc.fbm_generate(n, 1234)
    assert np.array_equal(_bits(gpu), _bits(cpu)), "device fBm generator is not bit-identical to the oracle's"
    assert v.majorant() == cpu.max()
    v.close()


# ---------------------------------------------------------------------------------------------------------
def _sv_check(gpu, cpu):
    big = cpu >= 1e-30
    rel = np.abs(gpu[big] - cpu[big]) / cpu[big]
    assert rel.max() <= 1e-5, f"shadow volume: max rel err {rel.max():.3e}"
    if (~big).any():
        assert np.abs(gpu[~big] - cpu[~big]).max() <= 1e-35
    return float(rel.max())


@pytest.mark.parametrize("layout", LAYOUTS)
@pytest.mark.parametrize("light", [(1.0, -1.0, 0.0), (0.3, -0.8, 0.5), (0.0, 0.0, 1.0)])
def test_shadow_volume_parity(crv, orc, ctx, layout, light):
    for nsv, ds in [(70, 50.0), (33, 400.0)]:
        r, cloud = make_scene(crv, orc, ctx, n=40, nsv=nsv, layout=layout, density_scaling=ds, light=light)
        gpu = r.shadowVolumeImage.download()
        cpu = orc.shadow_volume(cloud, r.cloudProperties, r.shadowVolumeProperties)
        _sv_check(gpu, cpu)
        assert gpu.min() < 0.9 and gpu.max() == 1.0       # the volume really holds shadow and lit voxels
        r.close()


def test_shadow_volume_nonuniform_grid(crv, orc, ctx):
    """ragged, non-cubic cloud grid with a non-zero boundary (clamp-to-edge texels matter)"""
    rng = np.random.default_rng(3)
    nx, ny, nz = 21, 34, 13
    cloud = rng.random(nx * ny * nz, dtype=np.float32)
    wl = crv.Workload("t", 0, 64, 64, 1, 45, 120.0, 0.0, None, (0.2, -1.0, 0.1))
    r = crv.make_renderer(ctx, wl)
    r.set_cloud_host(cloud, (nx, ny, nz), (0.01, 0.02, 0.03))
    cpu = orc.shadow_volume(cloud, r.cloudProperties, r.shadowVolumeProperties)
    _sv_check(r.shadowVolumeImage.download(), cpu)
    r.close()


@pytest.mark.parametrize("layout", LAYOUTS)
def test_shadow_volume_modes_bit_identical(crv, orc, ctx, layout):
    """the single-pass (interval + exact FMA division) and two-pass (parallel gather + streaming scan) builds are pure
    re-schedulings of the literal kernel: identical bits, for axis-aligned, oblique and upward light directions, ragged N
    and a non-cubic grid with a non-zero boundary"""
    rng = np.random.default_rng(11)
    cases = [((40, 40, 40), None, (1.0, -1.0, 0.0), 70, 50.0), ((40, 40, 40), None, (0.0, 0.0, 1.0), 33, 400.0),
             ((40, 40, 40), None, (-0.3, 0.8, -0.5), 61, 120.0), ((21, 34, 13), "rand", (0.2, -1.0, 0.1), 45, 120.0),
             ((8, 8, 8), "rand", (0.0, -1.0, 0.0), 129, 30.0), ((96, 96, 96), None, (1.0, -1.0, 0.0), 200, 50.0),
             ((96, 96, 96), None, (0.5, -0.2, 0.9), 203, 200.0)]
    for shape, kind, light, nsv, ds in cases:
        n = shape[0] * shape[1] * shape[2]
        cloud = rng.random(n, dtype=np.float32) if kind == "rand" else orc.fbm_generate(shape[0], 1234)
        wl = crv.Workload("t", 0, 64, 64, 1, nsv, ds, 0.0, None, light)
        r = crv.make_renderer(ctx, wl, cloud_layout=layout)
        vs = (0.01, 0.02, 0.03) if kind == "rand" else (1.0 / shape[0],) * 3
        r.set_cloud_host(cloud, shape, vs)
        sv = r.shadowVolumeTechnique
        assert sv.last_build_mode() == ("fused_rows" if layout == "brick" else "two_pass")     # what AUTO resolves to
        out = {}
        for mode in ["generic", "single_pass", "two_pass", "fused", "fused_v1", "fused_rows"]:
            sv.SetTuning(mode)
            r.shadowVolumeImage.upload(np.full(nsv ** 3, -7.0, np.float32))     # poison: every voxel must be rewritten
            r.UpdateShadowVolume()
            assert sv.last_build_mode() == mode
            out[mode] = r.shadowVolumeImage.download()
        for mode in ["single_pass", "two_pass", "fused", "fused_v1", "fused_rows"]:
            assert np.array_equal(_bits(out[mode]), _bits(out["generic"])), f"{mode} differs from the literal kernel ({shape}, {light}, N={nsv})"
        _sv_check(out["generic"], orc.shadow_volume(cloud, r.cloudProperties, r.shadowVolumeProperties))
        r.close()


# ---------------------------------------------------------------------------------------------------------
def _frame_parity(crv, orc, r, cloud, seed, absorption, min_match=0.99, expect_misses=True):
    sv_cpu = orc.shadow_volume(cloud, r.cloudProperties, r.shadowVolumeProperties)
    r.reset_accumulation()
    r.pathTracingTechnique.SetDebugCounters(True)
    r.pushConstants.seed = seed
    r.pathTracingTechnique.RecordDrawCommands(0, 0)
    gpu = r.resultImage.download()
    gs, gc = r.pathTracingTechnique.DownloadDebugCounters()
    pc = crv.PushConstants()
    pc.seed, pc.frameCount = seed, 0
    h, w = gpu.shape[:2]
    cpu = np.zeros((h, w, 4), np.float32)
    _, cs, cc = orc.path_trace_frame(cloud, sv_cpu, r.cameraProperties, r.cloudProperties, r.parameters,
                                     r.shadowVolumeProperties, pc, absorption, cpu)
    same = (gs == cs) & (gc == cc)
    err = np.abs(gpu - cpu) / np.maximum(np.abs(cpu), 1e-3)
    divergent = int((~same).sum())
    print(f"seed {seed}: {divergent}/{same.size} threshold-divergent paths ({divergent / same.size:.2e}); "
          f"max rel err on matching paths {err[same].max():.2e}; mean steps {cs.mean():.1f}, collisions {cc.mean():.2f}")
    assert same.mean() >= min_match, f"too many divergent paths: {divergent}/{same.size}"
    assert err[same].max() <= 1e-4
    assert cc.max() > 0, "scene must contain scattering paths"
    assert (cs == 0).any() or not expect_misses, "scene must contain misses"
    return divergent, same.size


@pytest.mark.parametrize("layout", LAYOUTS)
@pytest.mark.parametrize("g,absorption", [(None, 0.0), (0.85, 0.0), (0.85, 0.1), (-0.4, 0.3)])
def test_pt_single_frame_parity(crv, orc, ctx, layout, g, absorption):
    r, cloud = make_scene(crv, orc, ctx, layout=layout, g=g, absorption=absorption)
    for seed in [38, 7719, 21238]:
        _frame_parity(crv, orc, r, cloud, seed, absorption)
    r.close()


def test_pt_dense_long_paths(crv, orc, ctx):
    """config-4 style stress: densityScaling 200, albedo 0.99 -> long paths, heavy divergence inside warps"""
    r, cloud = make_scene(crv, orc, ctx, n=32, w=64, h=48, nsv=48, g=0.85, absorption=0.01, density_scaling=200.0)
    d, n = _frame_parity(crv, orc, r, cloud, 2437, 0.01, min_match=0.97)
    r.close()


def test_pt_camera_inside_cloud_and_zero_density(crv, orc, ctx):
    r, cloud = make_scene(crv, orc, ctx, n=32, w=64, h=48, nsv=40)
    # camera inside the box: entry point moves BACKWARDS to tmin (PathTracer.comp:387)
    r.cameraProperties.position[0], r.cameraProperties.position[1], r.cameraProperties.position[2] = 10.0, -20.0, 480.0
    r.apply_parameters()
    _frame_parity(crv, orc, r, cloud, 8855, 0.0, min_match=0.98, expect_misses=False)
    # densityScaling <= 0: every pixel is pure background (PathTracer.comp:379)
    r.cloudProperties.densityScaling = 0.0
    r.apply_parameters()
    r.reset_accumulation()
    r.pushConstants.seed = 5
    r.pathTracingTechnique.RecordDrawCommands(0, 0)
    img = r.resultImage.download()
    for (x, y) in [(0, 0), (31, 20), (63, 47)]:
        _, d = orc.camera_ray(r.cameraProperties, x, y)
        assert np.array_equal(_bits(img[y, x]), _bits(orc.background(d)))
    r.close()


def test_running_mean_over_existing_image(crv, orc, ctx):
    """frameCount > 0 over a non-zero image: the shader's accumulate formula (:420-424), bit for bit on misses"""
    r, cloud = make_scene(crv, orc, ctx, n=24, w=32, h=32, nsv=32)
    rng = np.random.default_rng(0)
    old = rng.random((32, 32, 4), dtype=np.float32)
    r.resultImage.upload(old)
    r.pushConstants.seed, r.pushConstants.frameCount = 1234, 7
    r.pathTracingTechnique.SetDebugCounters(True)
    r.pathTracingTechnique.RecordDrawCommands(0, 0)
    gpu = r.resultImage.download()
    steps, _ = r.pathTracingTechnique.DownloadDebugCounters()
    miss = steps == 0
    ys, xs = np.nonzero(miss)
    for y, x in list(zip(ys, xs))[:64]:
        _, d = orc.camera_ray(r.cameraProperties, int(x), int(y))
        want = (orc.background(d) + old[y, x] * np.float32(7)) / np.float32(8)
        assert np.array_equal(_bits(gpu[y, x]), _bits(want.astype(np.float32)))
    r.close()


# ---------------------------------------------------------------------------------------------------------
def test_batch_equals_sequential_frames_bitwise(crv, orc, ctx):
    r, _ = make_scene(crv, orc, ctx, g=0.85, absorption=0.1)
    for _ in range(6):
        r.DrawFrame()
    seq = r.resultImage.download().copy()
    r.reset_accumulation()
    r.render(6)
    bat = r.resultImage.download()
    assert np.array_equal(_bits(seq), _bits(bat)), "RenderBatch differs from successive RecordDrawCommands"
    # a second batch continues the running mean with frameCount = 6; tiny sample buffer forces several rounds
    r.pathTracingTechnique.SetTuning(sample_buffer_bytes=2 * 96 * 64 * 16)
    r.render(5)
    a = r.resultImage.download().copy()
    r.reset_accumulation()
    for _ in range(11):
        r.DrawFrame()
    assert np.array_equal(_bits(a), _bits(r.resultImage.download()))
    r.close()


def test_layouts_and_scheduling_are_bit_identical(crv, orc, ctx):
    """storage layout, occupancy and the wavefront threshold must not change a single bit of the image"""
    imgs = []
    cloud = orc.fbm_generate(40, 1234)
    for layout, ctas, thr in [("linear", 0, 0), ("brick", 1, 1), ("tex3d", 2, 32), ("linear", 3, 20)]:
        r, _ = make_scene(crv, orc, ctx, layout=layout, g=0.85, absorption=0.05, cloud=cloud)
        r.pathTracingTechnique.SetTuning(ctas_per_sm=ctas, track_threshold=thr)
        r.render(4)
        imgs.append(r.resultImage.download().copy())
        r.close()
    for im in imgs[1:]:
        assert np.array_equal(_bits(imgs[0]), _bits(im))


def test_schedulers_are_bit_identical(crv, orc, ctx):
    """the pool scheduler (paths in shared memory, one kind of work at a time) and the lane-owned scheduler run the same
    per-path arithmetic: identical image bits, identical counters, for every round length, in both math modes"""
    cloud = orc.fbm_generate(40, 1234)
    for math in [crv.PT_MATH_STRICT, crv.PT_MATH_DRIVER]:
        imgs, stats = [], []
        for sched, rounds in [("lanes", 0), ("pool", 0), ("pool", 1), ("pool", 3), ("pool", 16)]:
            r, _ = make_scene(crv, orc, ctx, layout="brick", g=0.85, absorption=0.05, cloud=cloud, density_scaling=120.0)
            r.extensions.math_mode = math
            r.apply_parameters()
            pt = r.pathTracingTechnique
            pt.SetScheduler(sched)
            pt.SetTuning(steps_per_decision=rounds)
            pt.ResetStats()
            r.render(5)
            imgs.append(r.resultImage.download().copy())
            st = pt.GetStats()
            stats.append((st.paths, st.tracking_steps, st.collisions))
            r.close()
        for im, st in zip(imgs[1:], stats[1:]):
            assert np.array_equal(_bits(imgs[0]), _bits(im)), f"scheduler changed the image (math mode {math})"
            assert st == stats[0], f"scheduler changed the counters: {st} vs {stats[0]}"


@pytest.mark.parametrize("frames", [1, 5, 32, 37])
def test_job_orders_are_bit_identical(crv, orc, ctx, frames):
    """crv_pt_set_job_order: frame-major over the classification's hit list, pixel-major over the list sorted by expected
    work (job / nframes by multiply-shift: a power of two, odd counts, one frame), and the automatic choice -- the order
    only decides which lane starts which path when: identical image bits and counters, in both accumulation modes and
    both schedulers"""
    cloud = orc.fbm_generate(40, 99)
    ref = {}
    for sched in ["pool", "lanes"]:
        for order in [0, 1, -1]:
            r, _ = make_scene(crv, orc, ctx, layout="brick", g=0.85, absorption=0.05, cloud=cloud, density_scaling=120.0)
            pt = r.pathTracingTechnique
            pt.SetScheduler(sched)
            pt.SetJobOrder(order)
            pt.ResetStats()
            r.render(frames)                                   # running mean, one batch
            mean = r.resultImage.download().copy()
            st = pt.GetStats()
            r.resultImage.clear()
            pt.RenderBatch(crv.msvc_rand_sequence(frames, 0), 0, crv.ACCUM_SUM)
            total = r.resultImage.download().copy()
            r.close()
            got = (_bits(mean), _bits(total), (st.paths, st.tracking_steps, st.collisions))
            if not ref:
                ref = got
                continue
            assert np.array_equal(ref[0], got[0]), f"job order {order} / {sched} changed the running-mean image"
            assert np.array_equal(ref[1], got[1]), f"job order {order} / {sched} changed the sum image"
            assert ref[2] == got[2], f"job order {order} / {sched} changed the counters: {got[2]} vs {ref[2]}"


@pytest.mark.parametrize("g,absorption,ds", [(None, 0.0, 50.0), (0.85, 0.1, 50.0), (0.85, 0.01, 200.0)])
def test_driver_math_mode_parity(crv, orc, ctx, g, absorption, ds):
    """CRV_PT_MATH_DRIVER (lg2 on the SFU, reciprocal multiply, contracted add for t += -log(zeta)/sigma): the same bars as
    the strict mode -- <= 1e-4 relative on paths whose counters match the oracle, divergent paths counted and bounded"""
    r, cloud = make_scene(crv, orc, ctx, layout="brick", g=g, absorption=absorption, density_scaling=ds)
    r.extensions.math_mode = crv.PT_MATH_DRIVER
    r.apply_parameters()
    tot_d = tot_n = 0
    for seed in [38, 7719, 21238, 2437]:
        d, n = _frame_parity(crv, orc, r, cloud, seed, absorption, min_match=0.97 if ds > 100 else 0.99)
        tot_d += d
        tot_n += n
    print(f"driver math: {tot_d}/{tot_n} threshold-divergent paths over 4 frames")
    r.close()


def test_driver_math_mode_converged(crv, orc, ctx):
    """32 frames in driver mode against the strict mode on the same seeds: the means agree far inside the Monte-Carlo noise"""
    spp = 32
    out = {}
    for math in [crv.PT_MATH_STRICT, crv.PT_MATH_DRIVER]:
        r, cloud = make_scene(crv, orc, ctx, n=32, w=64, h=48, nsv=40, layout="brick", g=0.85, absorption=0.1)
        r.extensions.math_mode = math
        r.apply_parameters()
        r.render(spp)
        out[math] = r.resultImage.download().copy()
        r.close()
    a, b = out[crv.PT_MATH_STRICT], out[crv.PT_MATH_DRIVER]
    rel = np.abs(a - b) / np.maximum(np.abs(a), 1e-3)
    print(f"driver vs strict, {spp} spp: {(rel > 1e-4).mean():.2e} of the pixels differ by more than 1e-4, max {rel.max():.2e}")
    assert (rel <= 1e-4).mean() > 0.97


def test_far_away_box_disables_the_approximate_pretest(crv, orc, ctx):
    """a cloud box ten million units from the origin: the FMA coordinate of the brick pre-test can no longer be bounded to a
    quarter texel, the host switches it off (every tentative collision goes through the exact sampler) and the image must
    still be the oracle's -- in both schedulers, with identical bits"""
    n, w, h, nsv = 24, 48, 40, 32
    cloud = orc.fbm_generate(n, 1234)
    imgs = []
    for sched in ["pool", "lanes"]:
        wl = crv.Workload("t", n, w, h, 1, nsv, 50.0, 0.0, 0.85, (1.0, -1.0, 0.0))
        r = crv.make_renderer(ctx, wl, cloud_layout="brick")
        r.set_cloud_host(cloud, (n, n, n), (1.0 / n,) * 3)
        off = 1.0e7
        cp = r.cloudProperties
        cp.bounds[0][0] += off
        cp.bounds[1][0] += off
        r.cameraProperties.position[0] += off
        r.apply_parameters()
        r.UpdateShadowVolume()                      # SetOrigin from the moved bounds, rebuild, rebind
        r.pathTracingTechnique.SetScheduler(sched)
        _frame_parity(crv, orc, r, cloud, 38, 0.0, min_match=0.98)
        imgs.append(r.resultImage.download().copy())
        r.close()
    assert np.array_equal(_bits(imgs[0]), _bits(imgs[1]))


def test_sum_mode_matches_mean(crv, orc, ctx):
    r, _ = make_scene(crv, orc, ctx, g=0.85)
    seeds = crv.msvc_rand_sequence(8)
    r.render(8, seeds=seeds)
    mean = r.resultImage.download().copy()
    r.reset_accumulation()
    r.render(8, seeds=seeds, accum_mode=crv.ACCUM_SUM)
    r.resultImage.scale(1.0 / 8)
    s = r.resultImage.download()
    assert np.allclose(s, mean, rtol=1e-5, atol=1e-7)
    r.close()


def test_converged_image_within_noise(crv, orc, ctx):
    """16 frames: GPU mean vs oracle mean within 5 sigma/sqrt(spp) per pixel (sigma from the oracle's samples)"""
    n, w, h, nsv, spp = 32, 64, 48, 40, 16
    r, cloud = make_scene(crv, orc, ctx, n=n, w=w, h=h, nsv=nsv, g=0.85, absorption=0.1)
    seeds = crv.msvc_rand_sequence(spp)
    r.render(spp, seeds=seeds)
    gpu = r.resultImage.download()
    sv_cpu = orc.shadow_volume(cloud, r.cloudProperties, r.shadowVolumeProperties)
    cpu = np.zeros((h, w, 4), np.float32)
    samples = []
    pc = crv.PushConstants()
    for f, s in enumerate(seeds):
        pc.seed, pc.frameCount = int(s), f
        smp, _, _ = orc.path_trace_frame(cloud, sv_cpu, r.cameraProperties, r.cloudProperties, r.parameters,
                                         r.shadowVolumeProperties, pc, 0.1, cpu)
        samples.append(smp)
    sigma = np.std(np.stack(samples), axis=0, ddof=1)
    bound = 5.0 * sigma / np.sqrt(spp) + 1e-5 * np.abs(cpu) + 1e-7
    outliers = np.abs(gpu - cpu) > bound
    frac = outliers.mean()
    print(f"converged: outlier fraction {frac:.2e}, max abs diff {np.abs(gpu - cpu).max():.3e}")
    assert frac <= 1e-3
    # and the images are identical wherever no path diverged: the overwhelming majority of pixels
    assert (np.abs(gpu - cpu) <= 1e-4 * np.maximum(np.abs(cpu), 1e-3)).mean() > 0.98
    r.close()


def test_stats_match_oracle_counters(crv, orc, ctx):
    r, cloud = make_scene(crv, orc, ctx, n=32, w=64, h=48, nsv=40, g=None)
    sv_cpu = orc.shadow_volume(cloud, r.cloudProperties, r.shadowVolumeProperties)
    r.pathTracingTechnique.SetSkipLevel(0)         # count the reference's work: every step executed
    r.pathTracingTechnique.ResetStats()
    r.pushConstants.seed = 38
    r.pathTracingTechnique.RecordDrawCommands(0, 0)
    st = r.pathTracingTechnique.GetStats()
    pc = crv.PushConstants()
    pc.seed, pc.frameCount = 38, 0
    _, cs, cc = orc.path_trace_frame(cloud, sv_cpu, r.cameraProperties, r.cloudProperties, r.parameters,
                                     r.shadowVolumeProperties, pc, 0.0, None)
    assert st.paths == 64 * 48
    assert abs(int(st.tracking_steps) - int(cs.sum())) <= 0.01 * cs.sum()
    assert abs(int(st.collisions) - int(cc.sum())) <= 0.02 * cc.sum() + 2
    assert 0 < st.density_fetches < st.tracking_steps, "the brick-max skip must remove fetches (and not all of them)"
    r.close()


SKIP_SCENES = {
    # name: make_scene kwargs
    "iso": dict(g=None),
    "hg_absorbing": dict(g=0.85, absorption=0.1),
    "back_scatter": dict(g=-0.4, absorption=0.3, light=(0.3, -0.8, 0.5)),
    "dense": dict(g=0.85, absorption=0.01, density_scaling=200.0),
    "ragged": dict(n=44, w=100, h=70, g=0.85, absorption=0.05),
}


@pytest.mark.parametrize("sched", ["pool", "lanes"])
@pytest.mark.parametrize("layout", ["brick", "linear"])
@pytest.mark.parametrize("name", sorted(SKIP_SCENES))
def test_skip_levels_are_bit_identical(crv, orc, ctx, name, layout, sched):
    """crv_pt_set_skip_level: background pixels for rays through empty bricks, segments cut at their last occupied brick
    and the lookup-free stretch in front of the first one must not change a single bit of the image, in either math mode"""
    for math in [crv.PT_MATH_STRICT, crv.PT_MATH_DRIVER]:
        imgs, stats = [], []
        for level in (0, 1):
            r, _ = make_scene(crv, orc, ctx, layout=layout, **SKIP_SCENES[name])
            r.extensions.math_mode = math
            r.apply_parameters()
            pt = r.pathTracingTechnique
            pt.SetScheduler(sched)
            pt.SetSkipLevel(level)
            pt.ResetStats()
            r.render(7)
            imgs.append(r.resultImage.download().copy())
            stats.append(pt.GetStats())
            r.close()
        assert np.array_equal(_bits(imgs[0]), _bits(imgs[1])), f"skip level changed the image ({name}, math {math})"
        full, fast = stats
        assert full.paths == fast.paths
        assert fast.collisions == full.collisions, "skipping must not lose a collision"
        assert fast.tracked_paths < full.tracked_paths, "the scene must contain rays that cross the box through empty bricks"
        assert fast.tracking_steps < full.tracking_steps
        print(f"{name}/{layout}/{sched}/math{math}: tracked {fast.tracked_paths}/{full.tracked_paths} paths, "
              f"steps {fast.tracking_steps}/{full.tracking_steps}")


def test_skip_level_camera_inside_and_sparse_cloud(crv, orc, ctx):
    """occupied spans with the camera inside the box (segments start at t_min < 0 behind the camera) and with a cloud that
    occupies a single brick far from the box centre; also a re-upload must re-classify (spans depend on the contents)"""
    n = 40
    cloud = orc.fbm_generate(n, 1234)
    for level_imgs in [[]]:
        for level in (0, 1):
            r, _ = make_scene(crv, orc, ctx, layout="brick", g=0.85, absorption=0.05, cloud=cloud)
            r.cameraProperties.position[2] = 300.0          # inside the box
            r.apply_parameters()
            r.pathTracingTechnique.SetSkipLevel(level)
            r.render(5)
            level_imgs.append(r.resultImage.download().copy())
            # sparse contents through the same technique object: one occupied corner brick
            sparse = np.zeros((n, n, n), np.float32)
            sparse[2:6, 30:34, 33:37] = 0.8
            r.set_cloud_host(sparse.ravel(), (n, n, n), (1.0 / n,) * 3)
            r.reset_accumulation()
            r.render(5)
            level_imgs.append(r.resultImage.download().copy())
            r.close()
        assert np.array_equal(_bits(level_imgs[0]), _bits(level_imgs[2]))
        assert np.array_equal(_bits(level_imgs[1]), _bits(level_imgs[3]))
        assert not np.array_equal(level_imgs[0], level_imgs[1]), "the re-uploaded cloud must change the image"


# ---------------------------------------------------------------------------------------------------------
def test_error_behaviour(crv, ctx):
    """status codes + message instead of the reference's abort() (Validation.cpp:4-10)"""
    cam, pc = crv.CameraProperties(), crv.PushConstants()
    pt = crv.RenderTechniquePT(ctx, None, cam, pc)
    with pytest.raises(crv.CrvError) as e:
        pt.RecordDrawCommands(0, 0)                      # nothing bound
    assert e.value.code == crv.capi.CRV_ERR_INVALID
    img = crv.Image(ctx, 64, 64)                         # wrong size for the default 800x600 camera
    pt.SetFrameReferences([img])
    with pytest.raises(crv.CrvError):
        pt.RecordDrawCommands(0, 0)
    with pytest.raises(crv.CrvError):
        crv.Volume(ctx, 0, 4, 4)
    with pytest.raises(crv.CrvError):
        pt.SetExtensions(crv.PathTracerExtensions(1.5))
    sp = crv.ShadowVolumeProperties()
    sv = crv.RenderTechniqueSV(ctx, sp, pc)
    with pytest.raises(crv.CrvError):
        sv.RecordDrawCommands(0, 0)
    sv.close()
    pt.close()
    img.close()


# ---------------------------------------------------------------------------------------------------------
# against the reference's OWN shaders (tests/golden/refshader_small.*: outputs of PathTracer.comp / ShadowVolume.comp
# compiled by g++, see tests/test_ref_shaders.py) -- no hand restatement between the CUDA path and the reference here
# ---------------------------------------------------------------------------------------------------------
def _refshader_renderer(crv, orc, ctx, case, layout):
    import json
    import os
    from tests.refshader_cases import case_cloud
    here = os.path.dirname(os.path.abspath(__file__))
    meta = json.load(open(os.path.join(here, "golden", "refshader_small.json")))[case["name"]]
    w, h = case["res"]
    wl = crv.Workload("t", 0, w, h, 1, case["nsv"], case["density_scaling"], 0.0, case["g"], case["light"], case["fov"])
    r = crv.make_renderer(ctx, wl, cloud_layout=layout)
    if case["cam_pos"] is not None:
        for i in range(3):
            r.cameraProperties.position[i] = case["cam_pos"][i]
        r.apply_parameters()
    cloud = case_cloud(orc, case)
    r.set_cloud_host(cloud, case["grid"], case["voxel"])
    # the product's own setters must have produced the very bytes the reference's setters did
    assert bytes(r.cameraProperties).hex() == meta["camera"] and bytes(r.cloudProperties).hex() == meta["cloud"]
    assert bytes(r.parameters).hex() == meta["parameters"] and bytes(r.shadowVolumeProperties).hex() == meta["shadow"]
    return r, cloud, meta


@pytest.mark.parametrize("layout", ["linear", "brick"])
@pytest.mark.parametrize("name", ["iso", "hg_dense", "inside", "ragged"])
def test_cuda_path_against_reference_shader_goldens(crv, orc, ctx, layout, name):
    import os
    from tests.refshader_cases import CASES
    case = [c for c in CASES if c["name"] == name][0]
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "refshader_small.npz"))
    r, cloud, meta = _refshader_renderer(crv, orc, ctx, case, layout)
    rel = _sv_check(r.shadowVolumeImage.download(), gold[f"{name}_sv"])
    r.reset_accumulation()
    pt = r.pathTracingTechnique
    pt.SetDebugCounters(True)
    for f, seed in enumerate(meta["seeds"]):          # progressive frames, as RenderLoop issues them
        r.pushConstants.seed = seed
        r.pushConstants.frameCount = f
        pt.RecordDrawCommands(0, 0)
    gpu = r.resultImage.download()
    gs, gc = pt.DownloadDebugCounters()
    ref_img, rs, rc = gold[f"{name}_image"], gold[f"{name}_steps"], gold[f"{name}_coll"]
    same = (gs == rs) & (gc == rc)
    err = np.abs(gpu - ref_img) / np.maximum(np.abs(ref_img), 1e-3)
    ok = same & (err.max(axis=2) <= 1e-4)
    print(f"{name}/{layout}: sv max rel {rel:.2e}; {int((~same).sum())}/{same.size} threshold-divergent in the last frame; "
          f"{int(ok.sum())}/{same.size} pixels within 1e-4 of PathTracer.comp's accumulated image")
    # a path that diverged in an EARLIER frame also moves the accumulated pixel: both kinds are bounded together
    assert ok.mean() >= 0.97, f"{name}: only {ok.mean():.4f} of the pixels match the reference shader's image"
    assert same.mean() >= 0.97
    r.close()


def test_cuda_path_against_live_reference_shaders(crv, orc, ctx):
    """a bigger scene through the compiled reference shaders themselves (prebuilt oracle/_ref/libcrv_refshaders.so)"""
    if orc.ref_shaders_lib() is None:
        pytest.skip("oracle/_ref/libcrv_refshaders.so not present")
    r, cloud = make_scene(crv, orc, ctx, n=64, w=160, h=96, nsv=80, layout="brick", g=0.85, density_scaling=80.0)
    sv_ref = orc.refsh_shadow_volume(cloud, r.cloudProperties, r.shadowVolumeProperties)
    _sv_check(r.shadowVolumeImage.download(), sv_ref)
    r.reset_accumulation()
    pt = r.pathTracingTechnique
    pt.SetDebugCounters(True)
    r.pushConstants.seed = 7719
    pt.RecordDrawCommands(0, 0)
    gpu = r.resultImage.download()
    gs, gc = pt.DownloadDebugCounters()
    ref = np.zeros_like(gpu)
    rs, rc = orc.refsh_path_trace(cloud, sv_ref, r.cameraProperties, r.cloudProperties, r.parameters,
                                  r.shadowVolumeProperties, orc.push_constants_bytes(7719, 0), ref)
    same = (gs == rs) & (gc == rc)
    err = np.abs(gpu - ref) / np.maximum(np.abs(ref), 1e-3)
    print(f"live reference shaders: {int((~same).sum())}/{same.size} threshold-divergent, max rel err {err[same].max():.2e}")
    assert same.mean() >= 0.99 and err[same].max() <= 1e-4 and rc.sum() > 1000
    r.close()


# ---------------------------------------------------------------------------------------------------------
# edge cases: empty / degenerate / ragged inputs, both schedulers
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("sched", ["pool", "lanes"])
def test_empty_cloud_is_pure_background(crv, orc, ctx, sched):
    """an all-zero grid: majorant 0 -> maxExtinction = max(0, 0.01) (main.cpp:144-163); every path crosses the box without
    a collision and every pixel is exactly the oracle's background-only sample; the shadow volume is all ones"""
    n, w, h, nsv = 16, 40, 24, 24
    cloud = np.zeros(n ** 3, np.float32)
    wl = crv.Workload("t", n, w, h, 1, nsv, 50.0, 0.0, 0.85, (1.0, -1.0, 0.0))
    r = crv.make_renderer(ctx, wl, cloud_layout="brick")
    r.pathTracingTechnique.SetScheduler(sched)
    r.set_cloud_host(cloud, (n, n, n), (1.0 / n,) * 3)
    assert r.cloudProperties.maxExtinction == np.float32(0.01)
    sv = r.shadowVolumeImage.download()
    assert (sv == 1.0).all()
    r.pathTracingTechnique.SetDebugCounters(True)
    r.pushConstants.seed = 38
    r.pathTracingTechnique.RecordDrawCommands(0, 0)
    gpu = r.resultImage.download()
    gs, gc = r.pathTracingTechnique.DownloadDebugCounters()
    pc = crv.PushConstants()
    pc.seed, pc.frameCount = 38, 0
    cpu = np.zeros((h, w, 4), np.float32)
    _, cs, cc = orc.path_trace_frame(cloud, sv, r.cameraProperties, r.cloudProperties, r.parameters, r.shadowVolumeProperties, pc, 0.0, cpu)
    assert gc.sum() == 0 and cc.sum() == 0
    assert np.array_equal(gs, cs), "tracking-iteration counts through empty space must be identical"
    assert np.array_equal(_bits(gpu), _bits(cpu)), "no transcendental reaches the pixel value without a collision"
    r.close()


@pytest.mark.parametrize("shape,res", [((1, 1, 1), (2, 2)), ((3, 5, 2), (34, 2)), ((9, 8, 17), (30, 38))])
def test_degenerate_sizes(crv, orc, ctx, shape, res):
    """one-voxel and ragged grids, 2x2 / two-row / non-multiple-of-32 images (fewer jobs than one warp, off-grid dispatch stride; CameraProperties keeps half sizes, so resolutions are even: UniformBuffers.h:46-50)"""
    rng = np.random.default_rng(21)
    cloud = (rng.random(shape[0] * shape[1] * shape[2], dtype=np.float32) * 0.8 + 0.2).astype(np.float32)
    w, h = res
    wl = crv.Workload("t", 0, w, h, 1, 19, 300.0, 0.0, None, (0.3, -0.8, 0.5))
    imgs = []
    for sched in ["pool", "lanes"]:
        r = crv.make_renderer(ctx, wl, cloud_layout="brick")
        r.pathTracingTechnique.SetScheduler(sched)
        r.set_cloud_host(cloud, shape, (0.5 / shape[0], 0.5 / shape[1], 0.5 / shape[2]))
        _sv_check(r.shadowVolumeImage.download(), orc.shadow_volume(cloud, r.cloudProperties, r.shadowVolumeProperties))
        sv_cpu = orc.shadow_volume(cloud, r.cloudProperties, r.shadowVolumeProperties)
        r.reset_accumulation()
        r.pathTracingTechnique.SetDebugCounters(True)
        r.pushConstants.seed = 7719
        r.pathTracingTechnique.RecordDrawCommands(0, 0)
        gpu = r.resultImage.download()
        gs, gc = r.pathTracingTechnique.DownloadDebugCounters()
        pc = crv.PushConstants()
        pc.seed, pc.frameCount = 7719, 0
        cpu = np.zeros((h, w, 4), np.float32)
        _, cs, cc = orc.path_trace_frame(cloud, sv_cpu, r.cameraProperties, r.cloudProperties, r.parameters,
                                         r.shadowVolumeProperties, pc, 0.0, cpu)
        same = (gs == cs) & (gc == cc)
        err = np.abs(gpu - cpu) / np.maximum(np.abs(cpu), 1e-3)
        assert same.mean() >= 0.9 and (not same.any() or err[same].max() <= 1e-4)
        r.render(3)                                   # and a batch on top of the frame, for the hit-list / fold path
        imgs.append(r.resultImage.download().copy())
        r.close()
    assert np.array_equal(_bits(imgs[0]), _bits(imgs[1]))


def test_more_frames_than_one_launch_carries(crv, orc, ctx):
    """300 frames: more than the 256 seeds one launch carries by value -> several launches, still the sequential result"""
    r, _ = make_scene(crv, orc, ctx, n=16, w=40, h=24, nsv=24, layout="brick", g=0.85, absorption=0.1)
    seeds = crv.msvc_rand_sequence(300)
    r.render(300, seeds=seeds)
    a = r.resultImage.download().copy()
    r.reset_accumulation()
    for f in range(0, 300, 75):
        r.render(75, seeds=seeds[f:f + 75])
    assert np.array_equal(_bits(a), _bits(r.resultImage.download()))
    r.close()


@pytest.mark.gpu
@pytest.mark.parametrize("layout", ["brick", "linear"])
@pytest.mark.parametrize("n", [40, 44, 42])
def test_upload_path_builds_the_same_skip_tables(crv, ctx, layout, n):
    """the separable upload pass (four voxels per thread when nx % 4 == 0, one otherwise) against the per-brick gather kernels
    the device generator uses: same image bits AND the same fetch counter -- a fetch happens exactly when a table entry
    fails to reject, so equal counters over thousands of lookups mean equal tables where it matters"""
    w, h, spp = 96, 64, 4
    wl = crv.Workload("t", n, w, h, spp, 56, 50.0, 0.1, 0.85)
    seeds = crv.msvc_rand_sequence(spp)
    out = []
    host = None
    for upload in (False, True):
        r = crv.make_renderer(ctx, wl, cloud_layout=layout)
        if upload:
            r.set_cloud_host(host, (n, n, n), (1.0 / n,) * 3)
        else:
            r.set_cloud_synthetic(n, 1234)
            host = r.cloudImage.download().copy()
        r.pathTracingTechnique.ResetStats()
        r.pathTracingTechnique.RenderBatch(seeds, 0, crv.ACCUM_RUNNING_MEAN, 0, 0)
        st = r.pathTracingTechnique.GetStats()
        out.append((r.resultImage.download().copy(), int(st.tracking_steps), int(st.collisions), int(st.density_fetches)))
        r.close()
    assert np.array_equal(_bits(out[0][0]), _bits(out[1][0]))
    assert out[0][1:] == out[1][1:], (out[0][1:], out[1][1:])
    assert out[0][3] > 0
