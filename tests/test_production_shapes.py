"""The BASELINE.json configurations at their named shapes, against the CPU oracle, under the driver-run `-m gpu` suite.

  c1  128^3 cloud, 512x512, 16 spp, N_sv 500: EVERY path of all 16 frames, isotropic and g = 0.85, strict and driver math
  c2 / c4  512^3 cloud (coarse threshold table at shift 1), 1920x1080: >= 24 strided rows of one frame
  c3  1024^3 cloud (shift 2), 3840x2160: >= 24 strided rows of one frame
  c5  1024^3 shadow volume for three of the animated sun directions, strided columns, <= 1e-5
plus the production table shapes (shift 1, 2) forced onto a 40^3 grid, bit-identical to shift 0, and a re-bind of the
technique to a volume with the same brick COUNT but another shape (the q-table cache was keyed on the count).

Bars as in test_gpu_parity.py: per-sample radiance <= 1e-4 relative (floor 1e-3) on paths whose (steps, collisions) counters
equal the oracle's; the others are threshold-divergent (libm vs CUDA transcendentals moved a compare): counted and bounded.
The cloud is generated on the device (bit-identical to the oracle's generator, test_fbm_device_generator_matches_oracle)
and downloaded for the oracle; the oracle samples the GPU's shadow volume (its own parity is the c5 test and test_gpu_parity).
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def _renderer(crv, ctx, name, math=None, layout="brick"):
    wl = crv.CONFIGS[name]
    r = crv.make_renderer(ctx, wl, cloud_layout=layout, math_mode=crv.PT_MATH_STRICT if math is None else math)
    r.set_cloud_synthetic(wl.n_cloud, wl.fbm_seed)
    return wl, r


def _rows_parity(crv, orc, r, wl, cloud_h, sv_h, seed, ys, frame_count=0):
    """one dispatch with counters on the GPU, the oracle on image rows `ys`: (checked, divergent, max rel err on matching paths)"""
    pt = r.pathTracingTechnique
    r.reset_accumulation()
    pt.SetDebugCounters(True)
    r.pushConstants.seed = int(seed)
    r.pushConstants.frameCount = frame_count
    pt.RecordDrawCommands(0, 0)
    g_img = r.resultImage.download()
    gs, gc = pt.DownloadDebugCounters()
    pt.SetDebugCounters(False)
    r.reset_accumulation()
    pt.RecordDrawCommands(0, 0)                       # production path (skip level 1): same bits
    assert np.array_equal(_bits(g_img), _bits(r.resultImage.download())), "skip level 1 changed the frame"
    pc = crv.PushConstants()
    pc.seed, pc.frameCount = int(seed), 0
    H, W = g_img.shape[:2]
    n_tot = n_div = 0
    max_err = 0.0
    c_img = np.zeros((H, W, 4), np.float32)
    for y in ys:
        _, cs, cc = orc.path_trace_frame(cloud_h, sv_h, r.cameraProperties, r.cloudProperties, r.parameters,
                                         r.shadowVolumeProperties, pc, wl.absorption, c_img, rect=(0, y, W, y + 1), want_sample=False)
        same = (gs[y] == cs[y]) & (gc[y] == cc[y])
        err = np.abs(g_img[y] - c_img[y]) / np.maximum(np.abs(c_img[y]), 1e-3)
        n_tot += W
        n_div += int((~same).sum())
        if same.any():
            max_err = max(max_err, float(err[same].max()))
    return n_tot, n_div, max_err, int(gc[ys].sum())


@pytest.mark.parametrize("name", ["c1_iso", "c1_hg"])
@pytest.mark.parametrize("math", ["strict", "driver"])
def test_c1_every_path_of_every_frame(crv, orc, ctx, name, math):
    """SURVEY 8d "strict parity config": 128^3, 512x512, 16 spp, N_sv = 500, absorption 0 (pure reference semantics)"""
    wl, r = _renderer(crv, ctx, name, crv.PT_MATH_STRICT if math == "strict" else crv.PT_MATH_DRIVER)
    pt = r.pathTracingTechnique
    cloud_h = np.ascontiguousarray(r.cloudImage.download(), np.float32)
    sv_h = np.ascontiguousarray(r.shadowVolumeImage.download(), np.float32)
    sv_cpu = orc.shadow_volume(cloud_h, r.cloudProperties, r.shadowVolumeProperties)
    big = sv_cpu >= 1e-30
    assert (np.abs(sv_h[big] - sv_cpu[big]) / sv_cpu[big]).max() <= 1e-5, "c1 shadow volume (N_sv = 500) beyond 1e-5"
    seeds = crv.msvc_rand_sequence(wl.spp)
    H, W = wl.height, wl.width
    pt.SetDebugCounters(True)
    total = divergent = 0
    worst = 0.0
    mean_gpu = np.zeros((H, W, 4), np.float64)
    mean_cpu = np.zeros((H, W, 4), np.float64)
    for f in range(wl.spp):
        r.reset_accumulation()
        r.pushConstants.seed = int(seeds[f])
        pt.RecordDrawCommands(0, 0)                   # frameCount 0: the frame's sample itself
        g = r.resultImage.download()
        gs, gc = pt.DownloadDebugCounters()
        pc = crv.PushConstants()
        pc.seed, pc.frameCount = int(seeds[f]), 0
        c = np.zeros((H, W, 4), np.float32)
        _, cs, cc = orc.path_trace_frame(cloud_h, sv_cpu, r.cameraProperties, r.cloudProperties, r.parameters,
                                         r.shadowVolumeProperties, pc, wl.absorption, c, want_sample=False)
        same = (gs == cs) & (gc == cc)
        err = np.abs(g - c) / np.maximum(np.abs(c), 1e-3)
        total += same.size
        divergent += int((~same).sum())
        worst = max(worst, float(err[same].max()))
        mean_gpu += g
        mean_cpu += c
    print(f"{name}/{math}: {divergent} threshold-divergent of {total} paths, max rel err on the others {worst:.2e}")
    assert total == 16 * 512 * 512
    assert worst <= 1e-4
    assert divergent <= 16, f"{divergent} threshold-divergent paths (a few per 4.19 M are expected: libm vs CUDA log/sin/cos)"
    # the 16-spp picture: divergent paths are still valid samples of the same estimator
    diff = np.abs(mean_gpu - mean_cpu) / 16.0
    assert (diff > 1e-4 * np.maximum(np.abs(mean_cpu) / 16.0, 1e-3)).sum() <= 4 * max(divergent, 1)
    r.close()


@pytest.mark.parametrize("name,expect_shift", [("c2", 1), ("c4", 1)])
def test_512_cubed_strided_rows(crv, orc, ctx, name, expect_shift):
    """512^3 cloud: the coarse threshold table no longer fits at brick granularity (shift 1, 16-voxel cells)"""
    for math in (crv.PT_MATH_STRICT, crv.PT_MATH_DRIVER):
        wl, r = _renderer(crv, ctx, name, math)
        cloud_h = np.ascontiguousarray(r.cloudImage.download(), np.float32)
        sv_h = np.ascontiguousarray(r.shadowVolumeImage.download(), np.float32)
        ys = list(range(7, wl.height, wl.height // 27))[:27]
        assert len(ys) >= 24
        seeds = crv.msvc_rand_sequence(3)
        n, d, e, coll = _rows_parity(crv, orc, r, wl, cloud_h, sv_h, seeds[2], ys)
        assert r.pathTracingTechnique.GetCoarseShift() == expect_shift
        print(f"{name} math {math}: {d} divergent of {n}, max rel err {e:.2e}, {coll} collisions on the checked rows")
        assert coll > 1000, "the checked rows must cross the cloud"
        assert e <= 1e-4
        assert d <= max(4, n // 20000), f"{d} threshold-divergent paths of {n}"
        del cloud_h, sv_h
        r.close()


@pytest.fixture(scope="module")
def big(crv, ctx):
    """the 1024^3 scene (c3: 4K picture, N_sv 1024), built once for the module"""
    wl, r = _renderer(crv, ctx, "c3", crv.PT_MATH_STRICT)
    cloud_h = np.ascontiguousarray(r.cloudImage.download(), np.float32)
    yield wl, r, cloud_h
    r.close()


def test_1024_cubed_4k_strided_rows(crv, orc, ctx, big):
    """c3: 1024^3 cloud (shift 2, 32-voxel cells), 3840x2160"""
    wl, r, cloud_h = big
    sv_h = np.ascontiguousarray(r.shadowVolumeImage.download(), np.float32)
    ys = list(range(11, wl.height, wl.height // 26))[:26]
    assert len(ys) >= 24
    seeds = crv.msvc_rand_sequence(2)
    for math in (crv.PT_MATH_STRICT, crv.PT_MATH_DRIVER):
        r.extensions.math_mode = math
        r.apply_parameters()
        n, d, e, coll = _rows_parity(crv, orc, r, wl, cloud_h, sv_h, seeds[1], ys)
        assert r.pathTracingTechnique.GetCoarseShift() == 2
        print(f"c3 math {math}: {d} divergent of {n}, max rel err {e:.2e}, {coll} collisions on the checked rows")
        assert coll > 1000
        assert e <= 1e-4
        assert d <= max(4, n // 20000)
    r.extensions.math_mode = crv.PT_MATH_STRICT
    r.apply_parameters()


def test_c5_shadow_volume_1024_animated_sun(crv, orc, ctx, big):
    """c5: ShadowVolume rebuild at N_sv = N_cloud = 1024 for three of the animated sun directions
    L_k = normalize(cos th_k, -1, sin th_k), th_k = 2 pi k / 64: strided columns against ShadowVolume.comp's
    restatement, <= 1e-5 relative where T >= 1e-30"""
    wl, r, cloud_h = big
    n = wl.n_sv
    xs = np.arange(5, n, 37, dtype=np.uint32)
    ys = np.arange(3, n, 41, dtype=np.uint32)
    for k in (1, 23, 47):
        th = 2.0 * np.pi * k / 64.0
        r.lightDirection = (float(np.cos(th)), -1.0, float(np.sin(th)))
        r.UpdateShadowVolume()
        ms = r.shadowVolumeTechnique.last_build_ms()
        gpu = r.shadowVolumeImage.download().reshape(n, n, n)[:, ys][:, :, xs]        # [z, y, x]
        cpu = orc.shadow_volume_columns(cloud_h, r.cloudProperties, r.shadowVolumeProperties, xs, ys)   # [y, x, z]
        gpu = np.ascontiguousarray(np.transpose(gpu, (1, 2, 0)))
        big_t = cpu >= 1e-30
        rel = np.abs(gpu[big_t] - cpu[big_t]) / cpu[big_t]
        assert (cpu < 0.999).mean() > 0.02, "the checked columns must cross the cloud"
        assert rel.max() <= 1e-5, f"sun {k}: shadow volume max rel err {rel.max():.3e}"
        assert np.abs(gpu[~big_t] - cpu[~big_t]).max() <= 1e-35 if (~big_t).any() else True
        print(f"sun {k}: {cpu.size} voxels checked, max rel err {rel.max():.2e}, rebuild {ms:.3f} ms ({r.shadowVolumeTechnique.last_build_mode()})")
    r.lightDirection = wl.light_dir
    r.UpdateShadowVolume()


def test_coarse_table_shapes_are_bit_identical(crv, orc, ctx):
    """the production cell sizes (16 voxels at 512^3, 32 at 1024^3) forced onto a 40^3 / ragged grid: same image bits and
    counters as at brick granularity, both schedulers of the table (pool tracker, PPM shares the builder)"""
    for shape_kw in (dict(n=40), dict(n=44, w=100, h=70)):
        imgs, stats, shifts = [], [], []
        for shift in (0, 1, 2, 3):
            wl = crv.Workload("t", shape_kw["n"], shape_kw.get("w", 96), shape_kw.get("h", 64), 1, 56, 120.0, 0.05, 0.85)
            r = crv.make_renderer(ctx, wl, cloud_layout="brick")
            r.set_cloud_host(orc.fbm_generate(shape_kw["n"], 1234), (shape_kw["n"],) * 3, (1.0 / shape_kw["n"],) * 3)
            pt = r.pathTracingTechnique
            pt.SetCoarseShift(shift)
            pt.ResetStats()
            r.render(6)
            imgs.append(r.resultImage.download().copy())
            st = pt.GetStats()
            stats.append((st.paths, st.tracking_steps, st.collisions, st.tracked_paths))
            shifts.append(pt.GetCoarseShift())
            # and back to automatic on the same object: the tables are rebuilt
            pt.SetCoarseShift(-1)
            r.reset_accumulation()
            r.render(6)
            assert pt.GetCoarseShift() == 0
            assert np.array_equal(_bits(imgs[0]), _bits(r.resultImage.download()))
            r.close()
        assert shifts == [0, 1, 2, 3]
        for im, st in zip(imgs[1:], stats[1:]):
            assert np.array_equal(_bits(imgs[0]), _bits(im)), "the coarse cell size changed the image"
            assert st == stats[0], f"the coarse cell size changed the counters: {st} vs {stats[0]}"


def test_rebind_to_same_brick_count_other_shape(crv, orc, ctx):
    """64x32x16 and 16x128x16 voxels both have 8*4*2 = 2*16*2 = 64 bricks, but other padded tables and
    strides: re-binding the same technique must rebuild them (it used to key the cache on the brick count alone)"""
    rng = np.random.default_rng(5)

    def cloud(shape):
        nx, ny, nz = shape
        z, y, x = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
        d = ((x / nx - 0.5) ** 2 + (y / ny - 0.5) ** 2 + (z / nz - 0.5) ** 2)
        return (np.clip(0.9 - 6.0 * d, 0, 1) * (0.6 + 0.4 * rng.random((nz, ny, nx)))).astype(np.float32).ravel()

    shapes = [(64, 32, 16), (16, 128, 16), (64, 32, 16)]
    clouds = [cloud(s) for s in shapes]
    clouds[2] = clouds[0]
    wl = crv.Workload("t", 64, 96, 64, 1, 40, 150.0, 0.05, 0.85)
    shared = crv.make_renderer(ctx, wl, cloud_layout="brick")
    for shape, data in zip(shapes, clouds):
        vs = tuple(1.0 / max(shape) for _ in range(3))
        shared.set_cloud_host(data, shape, vs)
        shared.render(4)
        got = shared.resultImage.download().copy()
        fresh = crv.make_renderer(ctx, wl, cloud_layout="brick")
        fresh.set_cloud_host(data, shape, vs)
        fresh.render(4)
        want = fresh.resultImage.download().copy()
        fresh.close()
        assert want.max() > 0.05
        assert np.array_equal(_bits(got), _bits(want)), f"re-bound technique differs from a fresh one at shape {shape}"
    shared.close()
