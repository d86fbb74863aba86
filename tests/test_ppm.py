"""PPM photon tracer: RenderTechniquePPM's tracing pass, shaders/PPM_PT.comp.

CPU (`-m "not gpu"`): the oracle's restatement against golden vectors produced by the reference's OWN shader compiled by g++
(tests/golden/ppm_small.*, generator make_golden_ppm.py), a live differential where oracle/_ref is present, and the
PhotonMapProperties mirror against the reference's own struct.  GPU: the CUDA kernel against the oracle -- the collision map
(integer) exactly, the photons of every cell as a set within the 1e-4 relative tolerance of the path tracer (positions move
with the last bits of log(); the angles come from acosf / atan2f)."""
import ctypes as C
import json
import os

import numpy as np
import pytest

from tests.ppm_cases import CASES, case_scene

HERE = os.path.dirname(os.path.abspath(__file__))


def _bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


# ---- CPU ------------------------------------------------------------------------------------------------------------
def test_photon_structs_match_the_reference_headers(crv, orc):
    meta = json.load(open(os.path.join(HERE, "golden", "ppm_small.json")))
    assert meta["sizeof"] == [C.sizeof(crv.capi.PhotonMapPropertiesPOD), C.sizeof(crv.capi.PhotonPOD)] == [76, 48]
    pm = crv.PhotonMapProperties()                          # defaults (UniformBuffers.h:214-222)
    assert pm.GetTotalSize() == 100 ** 3 and pm.photonSize == 48 and pm.stepSize == 10.0 and pm.voxelSize == 1.0
    for case in CASES:
        _, cp, _, _ = case_scene(orc, case, crv)
        pm = crv.PhotonMapProperties()
        pm.SetBounds(cp.bounds)                             # frozen bytes of the reference's SetBounds on the same bounds
        assert bytes(pm).hex() == meta[case["name"]]["ref_props_default_cells"]
        assert meta[case["name"]]["total_size_default"] == pm.GetTotalSize()
    ref = orc.ref_structs_lib()
    if ref is not None:                                     # live, on non-cubic bounds (largest extent wins, :230-241)
        b = (C.c_float * 8)(-3.0, 2.0, 1.0, 0.0, 5.0, 40.0, 9.0, 0.0)
        buf = (C.c_ubyte * 76)()
        ref.ref_photon_map_properties(b, 1, buf, None)
        pm = crv.PhotonMapProperties()
        pm.SetBounds([[b[0], b[1], b[2], b[3]], [b[4], b[5], b[6], b[7]]])
        assert bytes(pm) == bytes(buf)


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_restatement_matches_reference_shader_goldens(crv, orc, case):
    gold = np.load(os.path.join(HERE, "golden", "ppm_small.npz"))
    meta = json.load(open(os.path.join(HERE, "golden", "ppm_small.json")))[case["name"]]
    cloud, cp, par, pm = case_scene(orc, case, crv)
    assert bytes(pm).hex() == meta["props_used"]
    ph, co, steps = orc.ppm_trace(cloud, cp, par, pm, case["seed"], groups_x=case["groups"], threads=1)
    cells = gold[f"{case['name']}_cells"]
    assert np.array_equal(np.nonzero(co[:, 0])[0], cells)
    assert np.array_equal(co[cells, 0], gold[f"{case['name']}_counts"]) and int(co[:, 0].sum()) == meta["deposits"] > 500
    assert (co[:, 1:] == 0).all()
    assert np.array_equal(_bits(ph[cells][:, :, :10]), _bits(gold[f"{case['name']}_photons"])), "photon records differ from PPM_PT.comp's"
    assert steps > 10 * meta["deposits"]


def test_restatement_matches_live_reference_shader(crv, orc):
    if orc.ref_shaders_lib() is None:
        pytest.skip("oracle/_ref/libcrv_refshaders.so not present")
    case = dict(CASES[0], n=32, cells=8, groups=40, seed=21238, density_scaling=800.0)
    cloud, cp, par, pm = case_scene(orc, case, crv)
    a_ph, a_co, _ = orc.ppm_trace(cloud, cp, par, pm, case["seed"], groups_x=case["groups"], threads=1)
    b_ph, b_co = orc.refsh_ppm_trace(cloud, cp, par, pm, case["seed"], groups_x=case["groups"], threads=1)
    assert np.array_equal(a_co, b_co) and a_co[:, 0].max() > 32, "the scene must overflow a cell (slots beyond 32 are dropped)"
    assert np.array_equal(_bits(a_ph[:, :, :10]), _bits(b_ph[:, :, :10]))


# ---- GPU ------------------------------------------------------------------------------------------------------------
def _sorted_cell(ph, n):
    p = ph[:n, :10]
    return p[np.lexsort((p[:, 2], p[:, 1], p[:, 0]))]


def _direction(p):
    """the unit vector a photon record encodes: theta = acos(dir.z), phi = 2 * atan(dir.y, dir.x) (PPM_PT.comp:180-187)"""
    th, ph = p[:, 9].astype(np.float64), p[:, 8].astype(np.float64) / 2.0
    return np.stack([np.sin(th) * np.cos(ph), np.sin(th) * np.sin(ph), np.cos(th)], axis=1)


def _compare_maps(g_ph, g_co, c_ph, c_co, extent=1000.0):
    """-> (cells whose count differs, photons compared, worst error of matching cells on a common 1e-4 scale).  Power:
    relative; position: absolute against the box size, 1e-6 of it (a few ulp of a coordinate) counting as 1e-4 -- a
    coordinate near zero has no meaningful relative error; the two angles: as the direction they encode (acos is
    ill-conditioned at the poles), absolute."""
    diff_cells = np.nonzero(g_co[:, 0] != c_co[:, 0])[0]
    worst, compared = 0.0, 0
    for cell in np.nonzero(c_co[:, 0])[0]:
        n = int(c_co[cell, 0])
        if n > 32 or g_co[cell, 0] != n:
            continue                                        # which 32 of an overflowing cell survive is order-dependent
        a, b = _sorted_cell(g_ph[cell], n), _sorted_cell(c_ph[cell], n)
        perr = np.abs(a[:, :3] - b[:, :3]).max() / extent * 100.0            # position: 1e-6 of the box counts like 1e-4
        werr = (np.abs(a[:, 4:8] - b[:, 4:8]) / np.maximum(np.abs(b[:, 4:8]), 1e-30)).max()
        derr = np.abs(_direction(a) - _direction(b)).max()
        worst = max(worst, float(perr), float(werr), float(derr))
        compared += n
    return diff_cells, compared, worst


@pytest.mark.gpu
@pytest.mark.parametrize("layout", ["linear", "brick", "tex3d"])
@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_cuda_photon_tracer_against_the_oracle(crv, orc, ctx, case, layout):
    cloud, cp, par, pm = case_scene(orc, case, crv)
    n = case["n"]
    vol = crv.Volume(ctx, n, n, n, layout=layout)
    vol.upload(cloud)
    pc = crv.PushConstants()
    pc.seed = case["seed"]
    ppm = crv.RenderTechniquePPM(ctx, pm, pc)
    ppm.AllocateResources()
    ppm.QueueUpdateCloudData(cp)
    ppm.QueueUpdateCloudDataSampler(vol)
    ppm.QueueUpdateParameters(par)
    ppm.UpdateDescriptorSets()
    ppm.RecordPhotonTracing(0, case["groups"])
    g_ph, g_co = ppm.Download()
    st = ppm.LastTraceStats()
    c_ph, c_co, steps = orc.ppm_trace(cloud, cp, par, pm, case["seed"], groups_x=case["groups"])
    diff, compared, worst = _compare_maps(g_ph, g_co, c_ph, c_co)
    print(f"{case['name']}/{layout}: {int(c_co[:, 0].sum())} deposits, {len(diff)} cells with a different count, "
          f"{compared} photons compared, max rel err {worst:.2e}; steps gpu {st['tracking_steps']} cpu {steps}; {st['ms']:.3f} ms")
    assert len(diff) <= max(2, 0.01 * np.count_nonzero(c_co[:, 0])), "threshold-divergent photons move deposits between cells"
    assert abs(int(g_co[:, 0].sum()) - int(c_co[:, 0].sum())) <= 0.01 * c_co[:, 0].sum() + 2
    assert compared > 300 and worst <= 1e-4
    assert abs(st["tracking_steps"] - steps) <= 0.01 * steps and st["deposits"] == int(g_co[:, 0].sum())
    # a second dispatch with another seed leaves exactly what a fresh technique produces: the lazy clear is complete
    pc.seed = 2437
    ppm.RecordPhotonTracing(0, case["groups"])
    h_ph, h_co = ppm.Download()
    e_ph, e_co, _ = orc.ppm_trace(cloud, cp, par, pm, 2437, groups_x=case["groups"])
    assert np.array_equal(np.nonzero(h_co[:, 0])[0], np.nonzero(e_co[:, 0])[0]) or len(_compare_maps(h_ph, h_co, e_ph, e_co)[0]) <= 3
    empty = h_co[:, 0] == 0
    assert (h_ph[empty] == 0).all(), "stale photons of the previous dispatch survived"
    used = np.minimum(h_co[:, 0], 32)
    tail = np.arange(32)[None, :] >= used[:, None]
    assert (h_ph[tail] == 0).all()
    ppm.close()
    vol.close()


@pytest.mark.gpu
def test_photon_tracer_default_dispatch_and_errors(crv, orc, ctx):
    """groups_x = 0 is the reference's vkCmdDispatch(200, 1, 1): 12 800 photons into the default 100^3 map"""
    case = CASES[0]
    cloud, cp, par, _ = case_scene(orc, case, crv)
    pm = crv.PhotonMapProperties()
    pm.SetBounds(cp.bounds)
    n = case["n"]
    vol = crv.Volume(ctx, n, n, n, layout="brick")
    vol.upload(cloud)
    pc = crv.PushConstants()
    ppm = crv.RenderTechniquePPM(ctx, pm, pc)
    with pytest.raises(crv.CrvError):
        ppm.RecordPhotonTracing()                           # nothing bound
    ppm.QueueUpdateCloudData(cp)
    ppm.QueueUpdateCloudDataSampler(vol)
    ppm.QueueUpdateParameters(par)
    ppm.UpdateDescriptorSets()
    with pytest.raises(crv.CrvError):
        ppm.RecordPhotonTracing()                           # AllocateResources not called
    ppm.AllocateResources()
    ppm.RecordPhotonTracing()
    st = ppm.LastTraceStats()
    _, co = ppm.Download()
    _, c_co, steps = orc.ppm_trace(cloud, cp, par, pm, pc.seed, groups_x=200)
    assert st["deposits"] == int(co[:, 0].sum()) and abs(st["deposits"] - int(c_co[:, 0].sum())) <= 0.01 * c_co[:, 0].sum() + 2
    assert abs(st["tracking_steps"] - steps) <= 0.01 * steps
    ppm.close()
    vol.close()


# =====================================================================================================================
# estimate pass (PPM_PE.comp)
# =====================================================================================================================
from tests.ppm_cases import ESTIMATE, estimate_scene  # noqa: E402


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_estimate_restatement_matches_reference_shader_goldens(crv, orc, case):
    """the oracle's PPM_PE restatement over the oracle's photon map == the image PPM_PE.comp itself (compiled by g++) made
    from PPM_PT.comp's photon map: bit for bit, on top of a non-zero image with frameCount 3"""
    E = ESTIMATE
    gold = np.load(os.path.join(HERE, "golden", "ppm_small.npz"))[f"{case['name']}_estimate"]
    meta = json.load(open(os.path.join(HERE, "golden", "ppm_small.json")))[case["name"]]
    cloud, cp, par, pm, cam, sp, sv = estimate_scene(orc, case, crv)
    ph, co, _ = orc.ppm_trace(cloud, cp, par, pm, case["seed"], groups_x=case["groups"] * E["groups_scale"], threads=1)
    img = np.full((E["height"], E["width"], 4), E["old_value"], np.float32)
    steps = orc.ppm_estimate(cloud, sv, cam, cp, par, pm, sp, E["seed"], E["frame_count"], E["radius"], ph, co, img)
    assert np.array_equal(_bits(img), _bits(gold)) and steps > 10000
    assert meta["estimate_pixels_lit_by_photons"] > 50, "the scene must exercise the photon gather"


def _ppm_technique(crv, ctx, case, orc, layout="brick"):
    cloud, cp, par, pm, cam, sp, sv = estimate_scene(orc, case, crv)
    n, E = case["n"], ESTIMATE
    vol = crv.Volume(ctx, n, n, n, layout=layout)
    vol.upload(cloud)
    svol = crv.Volume(ctx, E["nsv"], E["nsv"], E["nsv"], layout=layout)
    svol.upload(sv)
    img = crv.Image(ctx, E["width"], E["height"])
    pc = crv.PushConstants()
    ppm = crv.RenderTechniquePPM(ctx, pm, pc, 10.0)
    ppm.AllocateResources()
    ppm.SetFrameReferences([img])
    ppm.QueueUpdateCloudData(cp)
    ppm.QueueUpdateCloudDataSampler(vol)
    ppm.QueueUpdateParameters(par)
    ppm.QueueUpdateCameraProperties(cam)
    ppm.QueueUpdateShadowVolume(sp)
    ppm.QueueUpdateShadowVolumeSampler(svol)
    ppm.UpdateDescriptorSets()
    return ppm, pc, img, (vol, svol), (cloud, cp, par, pm, cam, sp, sv)


@pytest.mark.gpu
@pytest.mark.parametrize("layout", ["linear", "brick", "tex3d"])
@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_cuda_photon_estimate_against_the_oracle(crv, orc, ctx, case, layout):
    """PPM_PE on the device over the ORACLE's photon map (uploaded), against the oracle's estimate: <= 1e-4 relative per
    pixel; pixels where a compare flipped on an ulp of sin/cos (a sky-march sample more or less) are counted and bounded"""
    E = ESTIMATE
    ppm, pc, img, vols, (cloud, cp, par, pm, cam, sp, sv) = _ppm_technique(crv, ctx, case, orc, layout)
    ph, co, _ = orc.ppm_trace(cloud, cp, par, pm, case["seed"], groups_x=case["groups"] * E["groups_scale"], threads=1)
    ppm.Upload(ph, co)
    img.upload(np.full((E["height"], E["width"], 4), E["old_value"], np.float32))
    pc.seed, pc.frameCount, pc.pmRadius = E["seed"], E["frame_count"], E["radius"]
    ppm.RecordPhotonEstimate(0, 0)
    gpu = img.download()
    st = ppm.LastEstimateStats()
    cpu = np.full((E["height"], E["width"], 4), E["old_value"], np.float32)
    steps = orc.ppm_estimate(cloud, sv, cam, cp, par, pm, sp, E["seed"], E["frame_count"], E["radius"], ph, co, cpu)
    err = np.abs(gpu - cpu) / np.maximum(np.abs(cpu), 1e-3)
    bad = err.max(axis=2) > 1e-4
    print(f"{case['name']}/{layout}: {int(bad.sum())}/{bad.size} pixels beyond 1e-4 (max {err.max():.2e}); march steps gpu {st['march_steps']} cpu {steps}; {st['ms']:.3f} ms")
    assert st["march_steps"] == steps, "the ray march is pure IEEE arithmetic: identical step counts"
    assert bad.mean() <= 0.01
    assert np.isfinite(gpu).all()
    ppm.close()
    img.close()
    for v in vols:
        v.close()


@pytest.mark.gpu
def test_ppm_progressive_frames(crv, orc, ctx):
    """RecordDrawCommands for a few frames: the radius follows UpdateRadius (RenderTechniquePPM.cpp:302-312), the image is the
    running mean of the frames, and the photon map matters (an empty map gives a darker image on average)"""
    case = CASES[0]
    ppm, pc, img, vols, (cloud, cp, par, pm, cam, sp, sv) = _ppm_technique(crv, ctx, case, orc)
    seeds = crv.msvc_rand_sequence(4)
    radius, radii = None, []
    for f in range(4):
        pc.seed, pc.frameCount = int(seeds[f]), f
        ppm.RecordDrawCommands(0, 0)
        radius = 10.0 if f <= 1 else np.float32(radius) * np.float32(np.float32((f - 1) + np.float32(0.8)) / np.float32(f)) ** np.float32(0.33333)
        radii.append(pc.pmRadius)
        assert abs(pc.pmRadius - float(radius)) <= 1e-5 * float(radius)
    assert radii[0] == radii[1] == 10.0 and radii[3] < radii[2] < 10.0
    a = img.download().copy()
    assert np.isfinite(a).all() and a.max() > 0
    st = ppm.LastTraceStats()
    assert st["deposits"] > 1000
    # the same last frame over an empty map differs where photons were gathered
    img.clear()
    pc.frameCount, pc.pmRadius = 0, 80.0
    ppm.RecordPhotonEstimate(0, 0)
    with_photons = img.download().copy()
    cells = pm.GetTotalSize()
    ppm.Upload(np.zeros((cells, 32, 12), np.float32), np.zeros((cells, 4), np.uint32))
    img.clear()
    ppm.RecordPhotonEstimate(0, 0)
    without = img.download()
    # (not pixel-wise brighter: a step that gathers photons also draws one more random number, :357, so the sky directions
    # of the rest of that pixel's march change)
    assert (np.abs(with_photons - without) > 1e-6).sum() > 100 and with_photons.mean() > without.mean()
    ppm.close()
    img.close()
    for v in vols:
        v.close()
