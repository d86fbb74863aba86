"""The CPU oracle against known-answer vectors derived from the shader text (tests/golden/oracle_kat.json), its frozen
small-scene outputs (oracle_small.npz) and properties the shaders imply.  No GPU, no product compute."""
import hashlib
import json
import math
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
KAT = json.load(open(os.path.join(HERE, "golden", "oracle_kat.json")))
SMALL = np.load(os.path.join(HERE, "golden", "oracle_small.npz"))
SMALL_STRUCTS = json.load(open(os.path.join(HERE, "golden", "structs_small.json")))


def _buf(hexstr):
    import ctypes as C
    b = bytes.fromhex(hexstr)
    return (C.c_ubyte * len(b)).from_buffer_copy(b)


# ---- RNG (PathTracer.comp:220-241,368) ------------------------------------------------------------------
def test_rng_survey_vector(orc):
    """SURVEY.md 8a a3, computed by hand from the algorithm: seed 38, gid (1,1), dispatch stride 32*17"""
    k = KAT["rng_survey"]
    assert orc.lib.orc_rng_seed_state(38, 1, 1, 17) == 20710 == k["state0"]
    assert orc.lib.orc_rng_warmup_count(20710) == 6 == k["warmup"]
    st, fl = orc.rng_draws(38, 1, 1, 17, 4)
    assert list(st) == [2451026975, 4128812277, 2471091190, 2933830221] == k["states"]
    assert [float(x) for x in fl] == [0.5706741809844971, 0.9613140225410461, 0.5753457546234131, 0.683085560798645]


def test_rng_pure_python_restatement(orc):
    """independent re-derivation of xorshift32 (13,17,5) + warm-up in Python integers"""
    def draws(seed, gx, gy, ngx, n):
        s0 = (seed & 0xffffffff) * (gx + gy * 32 * ngx) & 0xffffffff
        s = s0
        out = []
        for i in range(s0 % 7 + 2 + n):
            s ^= (s << 13) & 0xffffffff
            s ^= s >> 17
            s ^= (s << 5) & 0xffffffff
            if i >= s0 % 7 + 2:
                out.append(s)
        return out
    for k in KAT["rng_more"]:
        st, fl = orc.rng_draws(k["seed"], k["gid"][0], k["gid"][1], k["num_groups_x"], 6)
        assert list(st) == k["states"] == draws(k["seed"], k["gid"][0], k["gid"][1], k["num_groups_x"], 6)
        assert list(fl.view(np.uint32)) == k["float_bits"]
        assert all(np.float32(s) / np.float32(4294967296.0) == f for s, f in zip(st, fl))


def test_rng_stuck_state_and_range(orc):
    st, fl = orc.rng_draws(100, 0, 0, 17, 4)          # pixel 0: seed * 0 = 0 -> all-zero stuck state
    assert list(st) == [0, 0, 0, 0] and list(fl) == [0.0] * 4
    tap = orc.rng_tap(38, 64, 8, 4)
    f = tap.astype(np.float32) / np.float32(4294967296.0)
    assert f.min() >= 0.0 and f.max() <= 1.0          # 1.0 is reachable (u32 -> f32 rounds up), never above


# ---- helper functions ---------------------------------------------------------------------------------
def test_background_kat_and_stops(orc):
    for k in KAT["background"]:
        assert list(orc.background((0.3, k["y"], -0.2)).view(np.uint32)) == k["rgba_bits"]
    assert np.allclose(orc.background((0, -1.0, 0)), (0.0, 0.0, 0.02, 1.0))
    assert np.allclose(orc.background((0, 0.0, 0)), (0.7, 0.9, 1.0, 1.0), atol=1e-6)
    assert np.allclose(orc.background((0, 1.0, 0)), (0.01, 0.1, 0.7, 1.0), atol=1e-6)
    assert np.allclose(orc.background((0, 5.0, 0)), orc.background((0, 1.0, 0)))


def test_intersect_cloud_cases(crv, orc):
    """cases modelled on tests::intersectCloud (Tests.cpp:31-76): front hit, miss, inside, axis-parallel rays"""
    cp = crv.CloudProperties()
    cp.SetFromGrid((100, 100, 100), (0.01,) * 3, 0.5)             # box [-500,500]^2 x [0,1000]
    hit, tmax, tmin = orc.intersect_cloud(cp, (0, 0, -800), (0, 0, 1))
    assert hit and tmin == 800.0 and tmax == 1800.0
    hit, tmax, tmin = orc.intersect_cloud(cp, (0, 0, -800), (0, 1, 0))
    assert not hit
    hit, tmax, tmin = orc.intersect_cloud(cp, (10, -20, 480), (0.6, 0.0, 0.8))
    assert hit and tmin < 0 < tmax                                # origin inside: tmin behind the ray
    hit, tmax, tmin = orc.intersect_cloud(cp, (0, 0, 2000), (0, 0, 1))
    assert hit and tmax < 0                                        # box behind: caller rejects on tmax < 0 (:379)
    hit, _, _ = orc.intersect_cloud(cp, (600, 0, 500), (0, 0, 1))
    assert not hit                                                 # parallel to the slab, outside it (inf compares)
    d = np.array([0.3, -0.2, 0.93], np.float32)
    d /= np.linalg.norm(d)
    hit, tmax, tmin = orc.intersect_cloud(cp, (-100, 50, -800), d)
    p_in = np.array([-100, 50, -800], np.float32) + d * np.float32(tmin)
    assert hit and abs(p_in[2]) < 1e-3


def test_orthonormal_basis_is_orthonormal(orc):
    rng = np.random.default_rng(0)
    for _ in range(100):
        n = rng.normal(size=3).astype(np.float32)
        n /= np.linalg.norm(n)
        a, b = orc.orthonormal_basis(n)
        assert abs(np.dot(a, b)) < 1e-6 and abs(np.dot(a, n)) < 1e-6 and abs(np.dot(b, n)) < 1e-6
        assert abs(np.linalg.norm(a) - 1) < 1e-6 and abs(np.linalg.norm(b) - 1) < 1e-6


def test_hg_pdf_and_cdf_inversion(crv, orc):
    p = crv.Parameters()
    p.SetPhaseG(0.85)
    g = p.GetPhaseG()
    for c in np.linspace(-1, 1, 21):
        want = (1 - g * g) / (4 * math.pi * (1 + g * g - 2 * g * c) ** 1.5)
        assert abs(orc.calculate_pdf(p, float(c)) - want) <= 5e-5 * want      # fp32 cancellation in 1+g^2-2g*c near c = 1
    # the pdf integrates to 1 over the sphere, invertCDF maps [0,1] onto [-1,1] monotonically
    cs = np.linspace(-1, 1, 20001)
    pdf = np.array([orc.calculate_pdf(p, float(c)) for c in cs])
    assert abs(np.trapezoid(pdf, cs) * 2 * math.pi - 1) < 1e-3
    xs = np.linspace(0, 1, 101)
    inv = np.array([orc.invert_cdf(p, float(x)) for x in xs])
    assert abs(inv[0] + 1) < 1e-5 and abs(inv[-1] - 1) < 1e-5 and np.all(np.diff(inv) > 0)
    iso = crv.Parameters()
    assert abs(orc.sample_phase(iso, (0, 0, 1), (1, 0, 0)) - 1 / (4 * math.pi)) < 1e-8      # fp32 constant


def test_scatter_ray_draw_counts(crv, orc):
    """isotropic: 2 draws (1 when sin^2 <= 0); HG: phi first then cos(theta); directions stay ~unit length"""
    def advance(s, n):
        for _ in range(n):
            s ^= (s << 13) & 0xffffffff
            s ^= s >> 17
            s ^= (s << 5) & 0xffffffff
        return s
    iso, hg = crv.Parameters(), crv.Parameters()
    hg.SetPhaseG(0.85)
    d, s1 = orc.scatter_ray(iso, 12345, (0, 0, 1))
    assert s1 == advance(12345, 2) and abs(np.linalg.norm(d) - 1) < 1e-6
    d, s1 = orc.scatter_ray(hg, 12345, (0.6, 0.0, 0.8))
    assert s1 == advance(12345, 2) and abs(np.linalg.norm(d) - 1) < 1e-5
    d, s1 = orc.scatter_ray(iso, 0, (0, 0, 1))        # stuck state: xi = 0 -> z = -1, sin^2 = 0 -> single draw
    assert s1 == 0 and list(d) == [0.0, 0.0, -1.0]
    # forward-peaked: mean cosine of many HG samples ~ g
    s, cosines = 777, []
    for _ in range(4000):
        d, s = orc.scatter_ray(hg, s, (0, 0, 1))
        cosines.append(d[2])
    assert abs(np.mean(cosines) - 0.85) < 0.02


def test_texture3d_vulkan_semantics(orc):
    """LINEAR + CLAMP_TO_EDGE: texel centres at (i+0.5)/n reproduce texels, edges clamp, midpoints average"""
    rng = np.random.default_rng(1)
    nx, ny, nz = 5, 4, 3
    v = rng.random(nx * ny * nz, dtype=np.float32)
    T = lambda i, j, k: v[i + nx * (j + ny * k)]
    for (i, j, k) in [(0, 0, 0), (4, 3, 2), (2, 1, 1)]:
        assert orc.texture3d(v, nx, ny, nz, ((i + .5) / nx, (j + .5) / ny, (k + .5) / nz)) == T(i, j, k)
    assert orc.texture3d(v, nx, ny, nz, (0.0, 0.0, 0.0)) == T(0, 0, 0)
    assert orc.texture3d(v, nx, ny, nz, (-3.0, 7.0, 1.0)) == T(0, 3, 2)
    mid = orc.texture3d(v, nx, ny, nz, (1.0 / nx * 2, .5 / ny, .5 / nz))       # halfway between texel 1 and 2 in x
    assert abs(mid - 0.5 * (T(1, 0, 0) + T(2, 0, 0))) < 1e-6
    # interpolant stays inside the hull of the 8 texels
    for _ in range(200):
        s = rng.random(3)
        val = orc.texture3d(v, nx, ny, nz, s)
        assert v.min() - 1e-6 <= val <= v.max() + 1e-6


# ---- frozen generator and small-scene outputs -----------------------------------------------------------
@pytest.mark.parametrize("n", [16, 32])
def test_fbm_generator_frozen(orc, n):
    a = orc.fbm_generate(n, 1234)
    k = KAT[f"fbm_{n}"]
    assert hashlib.sha256(a.tobytes()).hexdigest() == k["sha256"]
    assert int(a.max().view(np.uint32)) == k["max_bits"] and int((a > 0).sum()) == k["nonzero"]
    assert 0.0 <= a.min() and a.max() <= 1.0
    # thread count must not change a bit
    assert np.array_equal(a, orc.fbm_generate(n, 1234, threads=1))


def test_small_scene_frozen_outputs(crv, orc):
    """regression pin of the oracle itself: SV and three PT frames on the 16^3 scene (bit-exact, any thread count)"""
    cam, cp, sp = _buf(SMALL_STRUCTS["camera"]), _buf(SMALL_STRUCTS["cloud"]), _buf(SMALL_STRUCTS["shadow"])
    cloud = orc.fbm_generate(16, 1234)
    sv = orc.shadow_volume(cloud, cp, sp, threads=3)
    assert np.array_equal(sv.view(np.uint32), SMALL["sv"].view(np.uint32))
    for tag, g, absorption, seed in [("iso", None, 0.0, 38), ("hg", 0.85, 0.0, 7719), ("hg_abs", 0.85, 0.1, 21238)]:
        par = crv.Parameters()
        if g is not None:
            par.SetPhaseG(g)
        img = np.zeros((24, 32, 4), np.float32)
        smp, steps, coll = orc.path_trace_frame(cloud, sv, cam, cp, par, sp, orc.push_constants_bytes(seed, 0), absorption, img)
        assert np.array_equal(steps, SMALL[f"{tag}_steps"]) and np.array_equal(coll, SMALL[f"{tag}_coll"])
        assert np.array_equal(smp.view(np.uint32), SMALL[f"{tag}_sample"].view(np.uint32))
        assert np.array_equal(img, smp)                     # frameCount 0 over a zero image: mean == sample
        assert coll.max() > 0 and (steps == 0).any()


def test_shadow_volume_properties(crv, orc):
    """size-independent properties: T in (0,1], T == 1 before the column enters the cloud, empty cloud -> all ones,
    doubling densityScaling squares the transmittance of fully-traversed columns (to rounding)"""
    cam, cp, sp = _buf(SMALL_STRUCTS["camera"]), _buf(SMALL_STRUCTS["cloud"]), _buf(SMALL_STRUCTS["shadow"])
    cloud = orc.fbm_generate(16, 1234)
    sv = orc.shadow_volume(cloud, cp, sp).reshape(20, 20, 20)      # [z][y][x]
    assert sv.max() == 1.0 and sv.min() > 0.0
    assert np.all(sv[0] == 1.0)                                     # first slice lies outside the cloud box
    assert np.array_equal(orc.shadow_volume(np.zeros_like(cloud), cp, sp), np.ones(8000, np.float32))
    cp2 = crv.CloudProperties.from_buffer_copy(bytes(cp))
    cp2.densityScaling = 2 * cp2.densityScaling
    sv2 = orc.shadow_volume(cloud, cp2, sp).reshape(20, 20, 20)
    assert np.allclose(sv2, sv * sv, rtol=2e-5, atol=1e-12)


def test_accumulate_formula(crv, orc):
    """(sample + old*frameCount)/(frameCount+1) (:420-424); sum mode adds"""
    cam, cp, sp = _buf(SMALL_STRUCTS["camera"]), _buf(SMALL_STRUCTS["cloud"]), _buf(SMALL_STRUCTS["shadow"])
    cloud = orc.fbm_generate(16, 1234)
    sv = SMALL["sv"]
    par = crv.Parameters()
    rng = np.random.default_rng(5)
    old = rng.random((24, 32, 4), dtype=np.float32)
    img = old.copy()
    smp, _, _ = orc.path_trace_frame(cloud, sv, cam, cp, par, sp, orc.push_constants_bytes(38, 9), 0.0, img)
    want = (smp + old * np.float32(9)) / np.float32(10)
    assert np.array_equal(img.view(np.uint32), want.astype(np.float32).view(np.uint32))
    img = old.copy()
    orc.path_trace_frame(cloud, sv, cam, cp, par, sp, orc.push_constants_bytes(38, 9), 0.0, img, accum_sum=1)
    assert np.array_equal(img, old + smp)


def test_rows_sample_equals_full_frame(crv, orc):
    cam, cp, sp = _buf(SMALL_STRUCTS["camera"]), _buf(SMALL_STRUCTS["cloud"]), _buf(SMALL_STRUCTS["shadow"])
    cloud = orc.fbm_generate(16, 1234)
    par = crv.Parameters()
    pc = orc.push_constants_bytes(38, 0)
    n, steps, coll = orc.path_trace_rows(cloud, SMALL["sv"], cam, cp, par, sp, pc, 0.0, 5, want_counters=True)
    assert n == 32 * 5
    assert np.array_equal(steps[::5], SMALL["iso_steps"][::5]) and np.array_equal(coll[::5], SMALL["iso_coll"][::5])
    assert steps[1::5].sum() == 0
