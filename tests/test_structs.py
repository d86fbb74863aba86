"""Parameter-struct mirrors vs the REFERENCE: golden bytes made from the reference's own UniformBuffers.h
(tests/golden/structs_ref.json, see make_golden.py) and, when oracle/_ref is present, a live differential test."""
import ctypes as C
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "structs_ref.json")))


def test_sizes(crv):
    assert GOLD["sizes"] == [24, 68, 60, 28, 152]
    assert [C.sizeof(t) for t in (crv.PushConstants, crv.CameraProperties, crv.CloudProperties, crv.Parameters,
                                  crv.ShadowVolumeProperties)] == GOLD["sizes"]


def test_defaults(crv):
    assert bytes(crv.PushConstants()).hex() == GOLD["push_constants_default"]
    assert bytes(crv.CloudProperties()).hex() == GOLD["cloud_default"]
    pc = crv.PushConstants()
    assert (pc.seed, pc.frameCount) == (100, 1)


def _camera(crv, case):
    c = crv.CameraProperties()
    c.SetResolution(case["w"], case["h"])
    rot = list(case["rot"]) if case["rot"] else None
    if rot:
        c.SetRotation(rot)
    fov = c.SetFOV(case["fov"]) if case["fov"] else None
    return c, rot, fov


@pytest.mark.parametrize("case", GOLD["camera"], ids=lambda c: f"{c['w']}x{c['h']}")
def test_camera_bytes(crv, case):
    c, rot, fov = _camera(crv, case)
    assert bytes(c).hex() == case["bytes"]
    if rot:
        assert rot == case["rot_out"]           # pitch clamped in place to +-80 (UniformBuffers.h:31)
    if fov:
        assert fov == case["fov_out"]           # fov clamped to 60..150 (UniformBuffers.h:54)
    assert (c.GetWidth(), c.GetHeight()) == (case["w"] // 2 * 2, case["h"] // 2 * 2)


@pytest.mark.parametrize("case", GOLD["shadow"], ids=lambda c: f"n{c['n']}")
def test_shadow_bytes(crv, case):
    s = crv.ShadowVolumeProperties()
    s.voxelAxisCount = case["n"]
    if case["light"]:
        s.SetLightDirection(case["light"])
    if case["origin"]:
        s.SetOrigin(case["origin"][0], case["origin"][1])
    assert bytes(s).hex() == case["bytes"]


@pytest.mark.parametrize("case", GOLD["parameters"], ids=lambda c: f"g{c['g']}")
def test_parameters_bytes(crv, case):
    p = crv.Parameters()
    if case["g"] is not None:
        p.SetPhaseG(case["g"])
    assert bytes(p).hex() == case["bytes"]


def test_default_phase_is_isotropic_and_invalid_g_rejected(crv):
    p = crv.Parameters()
    assert p.isotropic == 1 and p.maxRayBounces == 5 and p.lightIntensity == 5.0
    p.SetPhaseG(0.85)
    before = bytes(p)
    p.SetPhaseG(1.5)                 # rejected (UniformBuffers.h:95)
    assert bytes(p) == before
    p.SetPhaseG(0.99)                # clamped to 0.95
    assert abs(p.GetPhaseG() - 0.95) < 1e-7 and p.isotropic == 0
    p.SetPhaseG(0.0)                 # isotropic again; derived constants keep their last anisotropic values
    assert p.isotropic == 1 and abs(p.phaseOnePlusG2 - (1 + 0.95 ** 2)) < 1e-6


@pytest.mark.parametrize("case", GOLD["basis"], ids=lambda c: str(c["n"]))
def test_orthonormal_basis(crv, orc, case):
    r, u = crv.orthonormal_basis(case["n"])
    assert r.tobytes().hex() == case["right"] and u.tobytes().hex() == case["up"]
    # the shader's createOrthonormalBasis (PathTracer.comp:197-206) is the same frame
    xp, yp = orc.orthonormal_basis(case["n"])
    assert np.array_equal(xp, r) and np.array_equal(yp, u)


def test_survey_probe_values(crv):
    """cloud 1000^3, L = (1,-1,0)/sqrt2 -> right, up, b0 and voxelSize(N=500)"""
    s = crv.ShadowVolumeProperties()
    s.SetLightDirection((1, -1, 0))
    s.SetOrigin((-500, -500, 0), (500, 500, 1000))
    assert np.allclose(s.right[:3], (0.5, 0.5, -0.70710677), atol=1e-6)
    assert np.allclose(s.up[:3], (0.5, 0.5, 0.70710677), atol=1e-6)
    assert np.allclose(list(s.bounds[0]), (-1478.4, -253.65, 500.0, 1.0), atol=0.05)
    assert abs(s.voxelSize - 3.47104) < 1e-4


def test_set_cloud_properties(crv, orc):
    """SetCloudProperties (main.cpp:144-163): product mirror == oracle restatement, incl. the 0.01 floor"""
    for counts, vs, maj, ds in [((128, 128, 128), (1 / 128,) * 3, 0.91, 50.0), ((100, 100, 100), (0.01,) * 3, 0.0, 200.0),
                                ((21, 34, 13), (0.01, 0.02, 0.03), 0.004, 120.0)]:
        c = crv.CloudProperties()
        c.densityScaling = ds
        c.SetFromGrid(counts, vs, maj)
        buf = (C.c_ubyte * 60)()
        orc.lib.orc_set_cloud_properties((C.c_uint32 * 3)(*counts), (C.c_double * 3)(*vs), C.c_float(maj), C.c_float(ds), buf)
        assert bytes(c) == bytes(buf)
        assert c.maxExtinction == max(np.float32(maj), np.float32(0.01))
    c = crv.CloudProperties()
    c.SetFromGrid((100, 100, 100), (0.01,) * 3, 0.5)
    assert list(c.bounds[0]) == [-500.0, -500.0, 0.0, 0.0] and list(c.bounds[1]) == [500.0, 500.0, 1000.0, 0.0]


def test_msvc_rand_sequence(crv, orc):
    want = [38, 7719, 21238, 2437, 8855, 11797, 8365, 32285, 10450, 30612]
    assert list(crv.msvc_rand_sequence(10)) == want
    assert list(orc.msvc_rand(10)) == want


def test_live_differential_against_reference_headers(crv, orc):
    ref = orc.ref_structs_lib()
    if ref is None:
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    rng = np.random.default_rng(11)
    for _ in range(200):
        w, h = int(rng.integers(2, 4000)), int(rng.integers(2, 2400))
        fov = float(rng.uniform(30, 170))
        rot = [float(rng.uniform(-120, 120)), float(rng.uniform(-360, 360))]
        buf = (C.c_ubyte * 68)()
        r = (C.c_float * 2)(*rot)
        f = C.c_float(fov)
        ref.ref_camera(1, w, h, 1, r, 1, C.byref(f), buf)
        c = crv.CameraProperties()
        c.SetResolution(w, h)
        c.SetRotation(rot)
        c.SetFOV(fov)
        assert bytes(c) == bytes(buf)
        n = int(rng.integers(2, 2048))
        light = rng.normal(size=3).astype(np.float32) * 3
        lo = rng.uniform(-900, 0, 3).astype(np.float32)
        hi = lo + rng.uniform(1, 2000, 3).astype(np.float32)
        buf = (C.c_ubyte * 152)()
        ref.ref_shadow(n, 1, (C.c_float * 3)(*light), 1, (C.c_float * 3)(*lo), (C.c_float * 3)(*hi), buf)
        s = crv.ShadowVolumeProperties()
        s.voxelAxisCount = n
        s.SetLightDirection(light)
        s.SetOrigin(lo, hi)
        assert bytes(s) == bytes(buf)
        g = float(rng.uniform(-1.2, 1.2))
        buf = (C.c_ubyte * 28)()
        ref.ref_parameters(1, C.c_float(g), buf)
        p = crv.Parameters()
        p.SetPhaseG(g)
        assert bytes(p) == bytes(buf)
