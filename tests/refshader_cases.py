"""The small scenes the reference-shader golden vectors (tests/golden/refshader_small.*) are made of.  TEST INFRASTRUCTURE.

`build_case` makes the parameter blocks with the REFERENCE's setters (oracle/_ref/libcrv_refstructs.so) when that shim is
present; the frozen bytes in refshader_small.json are what the tests use where it is not (the GPU box)."""
from __future__ import annotations

import ctypes as C
import struct

import numpy as np

CASES = [
    # reference defaults: isotropic phase, camera outside looking at the box, sun (1,-1,0)
    dict(name="iso", grid=(16, 16, 16), voxel=(1 / 16,) * 3, res=(32, 24), fov=90.0, nsv=20, light=(1.0, -1.0, 0.0),
         density_scaling=120.0, g=None, frames=2, cam_pos=None),
    # Henyey-Greenstein forward peak, oblique sun, dense medium (many collisions per path)
    dict(name="hg_dense", grid=(16, 16, 16), voxel=(1 / 16,) * 3, res=(40, 28), fov=90.0, nsv=24, light=(0.3, -0.8, 0.5),
         density_scaling=600.0, g=0.85, frames=2, cam_pos=None),
    # backward-scattering phase, camera INSIDE the cloud box (entry moves back to tmin, PathTracer.comp:387)
    dict(name="inside", grid=(16, 16, 16), voxel=(1 / 16,) * 3, res=(24, 24), fov=120.0, nsv=17, light=(0.0, -1.0, 0.0),
         density_scaling=200.0, g=-0.3, frames=1, cam_pos=(10.0, -20.0, 400.0)),
    # ragged non-cubic grid with non-zero boundary texels (clamp-to-edge matters), odd image size -> off-image invocations
    dict(name="ragged", grid=(21, 12, 9), voxel=(0.04, 0.06, 0.09), res=(34, 18), fov=75.0, nsv=33, light=(0.2, -1.0, 0.1),
         density_scaling=90.0, g=0.6, frames=2, cam_pos=None, noise=True),
]


def case_cloud(orc, case):
    nx, ny, nz = case["grid"]
    if case.get("noise"):
        return np.random.default_rng(5).random(nx * ny * nz, dtype=np.float32)
    assert nx == ny == nz
    return orc.fbm_generate(nx, 1234)


def build_case(orc, case):
    """(cloud, camera68, cloud60, params28, shadow152) with the blocks made by the reference's own setters"""
    ref = orc.ref_structs_lib()
    cloud = case_cloud(orc, case)
    cam = orc._raw_buf(68)
    f = C.c_float(case["fov"])
    ref.ref_camera(1, case["res"][0], case["res"][1], 0, (C.c_float * 2)(0, 0), 1, C.byref(f), cam)
    if case["cam_pos"] is not None:            # CameraProperties::position is a public member (UniformBuffers.h:17)
        C.memmove(cam, struct.pack("<3f", *case["cam_pos"]), 12)
    cp = orc._raw_buf(60)
    orc.lib.orc_set_cloud_properties((C.c_uint32 * 3)(*case["grid"]), (C.c_double * 3)(*case["voxel"]),
                                     C.c_float(float(cloud.max())), C.c_float(case["density_scaling"]), cp)
    par = orc._raw_buf(28)
    ref.ref_parameters(0 if case["g"] is None else 1, C.c_float(case["g"] or 0.0), par)
    b = np.frombuffer(bytes(cp), np.float32)
    sp = orc._raw_buf(152)
    ref.ref_shadow(int(case["nsv"]), 1, orc._f3(case["light"]), 1, orc._f3(b[0:3]), orc._f3(b[4:7]), sp)
    return cloud, cam, cp, par, sp


def frozen_case(orc, case, meta):
    """same, from the bytes frozen in refshader_small.json (no reference code needed)"""
    m = meta[case["name"]]
    mk = lambda key, n: (C.c_ubyte * n).from_buffer_copy(bytes.fromhex(m[key]))
    return case_cloud(orc, case), mk("camera", 68), mk("cloud", 60), mk("parameters", 28), mk("shadow", 152)
