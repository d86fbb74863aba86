"""Small photon-tracing scenes shared by tests/golden/make_golden_ppm.py and tests/test_ppm.py."""
import numpy as np

CASES = [
    {"name": "iso", "n": 16, "density_scaling": 300.0, "g": None, "absorption": 0.0, "cells": 12, "seed": 38, "groups": 6,
     "light": (1.0, -1.0, 0.0)},
    {"name": "hg_absorbing", "n": 24, "density_scaling": 500.0, "g": 0.85, "absorption": 0.2, "cells": 16, "seed": 7719, "groups": 10,
     "light": (0.3, -0.8, 0.5)},
]


def case_scene(orc, case, crv=None):
    """(cloud, CloudProperties, Parameters, PhotonMapProperties) built with the PRODUCT's struct setters (themselves pinned
    against the reference's headers by tests/test_structs.py and test_ppm.py)"""
    if crv is None:
        import cloudrendering_vulkan_b200 as crv
    n = case["n"]
    cloud = orc.fbm_generate(n, 1234)
    cp = crv.CloudProperties()
    cp.densityScaling = case["density_scaling"]
    cp.SetFromGrid((n, n, n), (1.0 / n,) * 3, float(cloud.max()))
    par = crv.Parameters()
    if case["g"] is not None:
        par.SetPhaseG(case["g"])
    pm = crv.PhotonMapProperties()
    for i in range(3):
        pm.voxelCount[i] = case["cells"]
    pm.SetBounds(cp.bounds)
    pm.absorption = case["absorption"]
    L = np.array(case["light"] + (0.0,), np.float32)
    L = L / np.float32(np.sqrt(np.float32((L * L).sum())))
    for i in range(4):
        pm.lightDirection[i] = float(L[i])
    return cloud, cp, par, pm


ESTIMATE = {"width": 48, "height": 32, "nsv": 24, "radius": 60.0, "frame_count": 3, "seed": 7719, "old_value": 0.25, "groups_scale": 4}


def estimate_scene(orc, case, crv=None):
    """camera / shadow-volume structs and the CPU shadow volume for the photon-estimate cases"""
    if crv is None:
        import cloudrendering_vulkan_b200 as crv
    cloud, cp, par, pm = case_scene(orc, case, crv)
    cam = crv.CameraProperties()
    cam.SetResolution(ESTIMATE["width"], ESTIMATE["height"])
    cam.SetFOV(90.0)
    sp = crv.ShadowVolumeProperties()
    sp.voxelAxisCount = ESTIMATE["nsv"]
    sp.SetLightDirection(case["light"])
    sp.SetOrigin(tuple(cp.bounds[0])[:3], tuple(cp.bounds[1])[:3])
    sv = orc.shadow_volume(cloud, cp, sp)
    return cloud, cp, par, pm, cam, sp, sv
