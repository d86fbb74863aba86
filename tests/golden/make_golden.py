"""Generates the committed golden fixtures under tests/golden/.

Two sources:
  * REFERENCE code: oracle/_ref/libcrv_refstructs.so, compiled from the reference's own UniformBuffers.h /
    Grid3D.h (+ vendored glm) where they lie under /root/reference.  -> structs_ref.json, grid_ref_*.xyz
    These pin the product's struct mirrors and .xyz reader/writer to real reference behaviour.
  * the CPU oracle (oracle/libcrv_oracle.so): known-answer vectors derived from the shader text, frozen so
    that later edits of the oracle cannot drift silently.  -> oracle_kat.json, oracle_small.npz
    (The reference's own tests pin nothing on this path; the oracle is pinned against the reference's shader
    sources by make_golden_refshaders.py / tests/test_ref_shaders.py, see oracle/crv_oracle.h.)

Run in the build container (needs /root/reference for the first group):  python tests/golden/make_golden.py
"""
from __future__ import annotations

import ctypes as C
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from tests import oracle_binding as orc  # noqa: E402

CAMERA_CASES = [  # (w, h, rotation or None, fov or None)
    (800, 600, None, None), (1920, 1080, None, 90.0), (512, 512, None, 90.0), (3840, 2160, None, 90.0),
    (800, 600, (30.0, 45.0), 120.0), (512, 512, (-95.0, 10.0), 60.0), (1920, 1080, (12.5, -170.0), 200.0),
    (64, 48, (80.0, 359.0), 10.0),
]
SHADOW_CASES = [  # (n_axis, light or None, (lo, hi) or None)
    (500, None, None),
    (500, (1.0, -1.0, 0.0), ((-500.0, -500.0, 0.0), (500.0, 500.0, 1000.0))),
    (512, (1.0, -1.0, 0.0), ((-500.0, -500.0, 0.0), (500.0, 500.0, 1000.0))),
    (1024, (0.3, -0.8, 0.5), ((-500.0, -500.0, 0.0), (500.0, 500.0, 1000.0))),
    (64, (0.0, 0.0, -1.0), ((-1.0, -2.0, 0.0), (1.0, 2.0, 3.0))),
    (33, (0.0, 1.0, 0.0), ((-50.0, -25.0, 0.0), (50.0, 25.0, 10.0))),
    (40, (-2.0, -0.1, -7.0), None),
]
PHASE_CASES = [None, 0.85, -0.3, 0.99, -1.0, 1.5, 0.00001, 0.0]
BASIS_CASES = [(0.0, 0.0, 1.0), (0.0, 0.0, -1.0), (1.0, 0.0, 0.0), (0.70710677, -0.70710677, 0.0),
               (0.26726124, -0.5345225, 0.80178374), (-0.6, 0.0, -0.8)]


def ref_fixtures():
    ref = orc.ref_structs_lib()
    if ref is None:
        print("oracle/_ref not built: skipping reference-derived fixtures")
        return
    out = {"sizes": [ref.ref_sizeof(i) for i in range(5)], "camera": [], "shadow": [], "parameters": [], "basis": []}
    buf = (C.c_ubyte * 24)()
    ref.ref_push_constants_default(buf)
    out["push_constants_default"] = bytes(buf).hex()
    buf = (C.c_ubyte * 60)()
    ref.ref_cloud_default(buf)
    out["cloud_default"] = bytes(buf).hex()
    for (w, h, rot, fov) in CAMERA_CASES:
        buf = (C.c_ubyte * 68)()
        r = (C.c_float * 2)(*(rot or (0.0, 0.0)))
        f = C.c_float(fov or 0.0)
        ref.ref_camera(1, w, h, 1 if rot else 0, r, 1 if fov else 0, C.byref(f), buf)
        out["camera"].append({"w": w, "h": h, "rot": rot, "fov": fov, "rot_out": [r[0], r[1]], "fov_out": f.value,
                              "bytes": bytes(buf).hex()})
    for (n, light, origin) in SHADOW_CASES:
        buf = (C.c_ubyte * 152)()
        l3 = (C.c_float * 3)(*(light or (0, 0, 0)))
        lo = (C.c_float * 3)(*(origin[0] if origin else (0, 0, 0)))
        hi = (C.c_float * 3)(*(origin[1] if origin else (0, 0, 0)))
        ref.ref_shadow(n, 1 if light else 0, l3, 1 if origin else 0, lo, hi, buf)
        out["shadow"].append({"n": n, "light": light, "origin": origin, "bytes": bytes(buf).hex()})
    for g in PHASE_CASES:
        buf = (C.c_ubyte * 28)()
        ref.ref_parameters(0 if g is None else 1, C.c_float(g or 0.0), buf)
        out["parameters"].append({"g": g, "bytes": bytes(buf).hex()})
    for n in BASIS_CASES:
        r3, u3 = (C.c_float * 3)(), (C.c_float * 3)()
        ref.ref_basis((C.c_float * 3)(*n), r3, u3)
        out["basis"].append({"n": n, "right": np.array(r3[:], np.float32).tobytes().hex(),
                             "up": np.array(u3[:], np.float32).tobytes().hex()})
    json.dump(out, open(os.path.join(HERE, "structs_ref.json"), "w"), indent=1)

    # .xyz files written by the REFERENCE Grid3D::Save (x-fastest raw dump) + what its Load makes of them
    nx, ny, nz = 5, 4, 3
    data = (np.arange(nx * ny * nz, dtype=np.float32) * 0.25 + 1.0)
    ref.ref_grid_save.argtypes = [C.c_char_p, C.c_uint, C.c_uint, C.c_uint, C.c_double, C.c_double, C.c_double, C.c_void_p]
    path = os.path.join(HERE, "grid_ref_saved.xyz")
    ref.ref_grid_save(path.encode(), nx, ny, nz, 0.01, 0.02, 0.04, data.ctypes.data_as(C.c_void_p))
    counts, vs, maj = (C.c_uint * 3)(), (C.c_double * 3)(), C.c_float()
    loaded = np.zeros(nx * ny * nz, np.float32)
    ref.ref_grid_load.argtypes = [C.c_char_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
    rc = ref.ref_grid_load(path.encode(), counts, vs, loaded.ctypes.data_as(C.c_void_p), loaded.size, C.byref(maj))
    assert rc == 0
    json.dump({"counts": list(counts), "voxel_size": list(vs), "majorant": maj.value, "saved_input": data.tolist(),
               "loaded": loaded.tolist()}, open(os.path.join(HERE, "grid_ref_loaded.json"), "w"), indent=1)
    print("reference-derived fixtures written")


def small_scene():
    """The tiny scene all oracle golden vectors use (no product code involved: structs come from the reference shim
    when present, else from the frozen hex in structs_small.json)."""
    import cloudrendering_vulkan_b200 as crv
    n, w, h, nsv = 16, 32, 24, 20
    cloud = orc.fbm_generate(n, 1234)
    cam = crv.CameraProperties(); cam.SetResolution(w, h); cam.SetFOV(90.0)
    cp = crv.CloudProperties(); cp.densityScaling = 120.0
    cp.SetFromGrid((n, n, n), (1.0 / n,) * 3, float(cloud.max()))
    sp = crv.ShadowVolumeProperties(); sp.voxelAxisCount = nsv
    sp.SetLightDirection((1.0, -1.0, 0.0))
    sp.SetOrigin(tuple(cp.bounds[0])[:3], tuple(cp.bounds[1])[:3])
    return crv, n, w, h, nsv, cloud, cam, cp, sp


def oracle_fixtures():
    kat = {}
    # the survey's hand-derived vector: seed 38, gid (1,1), 17 groups -> state0 20710, warm-up 6
    st, fl = orc.rng_draws(38, 1, 1, 17, 4)
    kat["rng_survey"] = {"seed": 38, "gid": [1, 1], "num_groups_x": 17, "state0": int(orc.lib.orc_rng_seed_state(38, 1, 1, 17)),
                         "warmup": int(orc.lib.orc_rng_warmup_count(20710)), "states": [int(x) for x in st],
                         "floats": [float(x) for x in fl]}
    kat["rng_more"] = []
    for (seed, gx, gy, ngx) in [(7719, 959, 539, 61), (100, 0, 0, 17), (-5, 3, 2, 26), (21238, 1919, 1079, 61), (2437, 17, 0, 121)]:
        st, fl = orc.rng_draws(seed, gx, gy, ngx, 6)
        kat["rng_more"].append({"seed": seed, "gid": [gx, gy], "num_groups_x": ngx, "states": [int(x) for x in st],
                                "float_bits": [int(x) for x in fl.view(np.uint32)]})
    kat["background"] = [{"y": y, "rgba_bits": [int(v) for v in orc.background((0.0, y, 0.0)).view(np.uint32)]}
                         for y in [-1.0, -0.5, -0.04, -0.02, 0.0, 0.1, 0.25, 0.5, 0.75, 1.0, 2.0]]
    for n in [16, 32]:
        a = orc.fbm_generate(n, 1234)
        kat[f"fbm_{n}"] = {"sha256": hashlib.sha256(a.tobytes()).hexdigest(), "max_bits": int(a.max().view(np.uint32)),
                           "nonzero": int((a > 0).sum())}
    json.dump(kat, open(os.path.join(HERE, "oracle_kat.json"), "w"), indent=1)

    crv, n, w, h, nsv, cloud, cam, cp, sp = small_scene()
    sv = orc.shadow_volume(cloud, cp, sp)
    out = {"sv": sv}
    for tag, g, absorption, seed in [("iso", None, 0.0, 38), ("hg", 0.85, 0.0, 7719), ("hg_abs", 0.85, 0.1, 21238)]:
        par = crv.Parameters()
        if g is not None:
            par.SetPhaseG(g)
        pc = crv.PushConstants(); pc.seed = seed; pc.frameCount = 0
        img = np.zeros((h, w, 4), np.float32)
        smp, steps, coll = orc.path_trace_frame(cloud, sv, cam, cp, par, sp, pc, absorption, img)
        out[f"{tag}_sample"] = smp
        out[f"{tag}_steps"] = steps
        out[f"{tag}_coll"] = coll
    np.savez_compressed(os.path.join(HERE, "oracle_small.npz"), **out)
    json.dump({"camera": bytes(cam).hex(), "cloud": bytes(cp).hex(), "shadow": bytes(sp).hex()},
              open(os.path.join(HERE, "structs_small.json"), "w"), indent=1)
    print("oracle fixtures written")


if __name__ == "__main__":
    ref_fixtures()
    oracle_fixtures()
