"""cloudrendering_vulkan_b200 -- B200-native replacement of CloudRendering-Vulkan's PathTracer + ShadowVolume hot path.

Layout: csrc/ (hand-written sm_100a CUDA kernels + the C ABI of include/crv_c_api.h, built into
libcrv_b200.so), capi.py (ctypes binding), structs.py / technique.py (host-side mirror of UniformBuffers.h
and the RenderTechnique interface), scene.py (headless driver + BASELINE workloads).  Importing this package
loads the shared library and fails loudly if it has not been built: there is no CPU path.
"""
from . import capi
from .capi import PT_MATH_DRIVER, PT_MATH_STRICT, ACCUM_RUNNING_MEAN, ACCUM_SUM, LAYOUT_BRICK, LAYOUT_LINEAR, LAYOUT_TEX3D, CrvError
from .scene import CONFIGS, HeadlessRenderer, Workload, make_renderer
from .structs import (CameraProperties, CloudProperties, Grid3D, Parameters, PathTracerExtensions, PhotonMapProperties, PushConstants,
                      ShadowVolumeProperties, msvc_rand_sequence, orthonormal_basis)
from .technique import Context, Group, Image, Stream, RenderTechnique, RenderTechniquePPM, RenderTechniquePT, RenderTechniqueSV, Volume, rng_tap

__all__ = [
    "capi", "CrvError", "ACCUM_RUNNING_MEAN", "ACCUM_SUM", "PT_MATH_STRICT", "PT_MATH_DRIVER", "LAYOUT_LINEAR", "LAYOUT_BRICK", "LAYOUT_TEX3D",
    "CONFIGS", "HeadlessRenderer", "Workload", "make_renderer",
    "CameraProperties", "CloudProperties", "Grid3D", "Parameters", "PathTracerExtensions", "PhotonMapProperties", "PushConstants",
    "ShadowVolumeProperties", "msvc_rand_sequence", "orthonormal_basis",
    "Context", "Group", "Image", "Stream", "RenderTechnique", "RenderTechniquePPM", "RenderTechniquePT", "RenderTechniqueSV", "Volume", "rng_tap",
]
