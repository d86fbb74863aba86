"""ctypes binding of include/crv_c_api.h (the drop-in C ABI of the PathTracer + ShadowVolume hot path).

The shared library is built in-tree (cloudrendering_vulkan_b200/libcrv_b200.so, see csrc/Makefile) and is the
only compute backend: if it is missing this module raises at import time -- there is no Python/CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcrv_b200.so")

CRV_OK = 0
CRV_ERR_INVALID = -1
CRV_ERR_NO_DEVICE = -2
CRV_ERR_CUDA = -3
CRV_ERR_OOM = -4
CRV_ERR_IO = -5

LAYOUT_LINEAR = 0
LAYOUT_BRICK = 1
LAYOUT_TEX3D = 2
LAYOUTS = {"linear": LAYOUT_LINEAR, "brick": LAYOUT_BRICK, "tex3d": LAYOUT_TEX3D}

ACCUM_RUNNING_MEAN = 0
ACCUM_SUM = 1
PT_MATH_STRICT = 0
PT_MATH_DRIVER = 1


# ---- parameter blocks: byte-for-byte UniformBuffers.h:6-167 ---------------------------------------------
class PushConstantsPOD(C.Structure):
    _fields_ = [("time", C.c_double), ("seed", C.c_int32), ("frameCount", C.c_uint32), ("pmRadius", C.c_float),
                ("currentBuffer", C.c_uint32)]


class CameraPropertiesPOD(C.Structure):
    _fields_ = [("position", C.c_float * 3), ("halfWidth", C.c_int32), ("forward", C.c_float * 3),
                ("halfHeight", C.c_int32), ("right", C.c_float * 3), ("nearPlane", C.c_float),
                ("up", C.c_float * 3), ("pixelSizeY", C.c_float), ("pixelSizeX", C.c_float)]


class CloudPropertiesPOD(C.Structure):
    _fields_ = [("bounds", (C.c_float * 4) * 2), ("voxelCount", C.c_uint32 * 4), ("maxExtinction", C.c_float),
                ("baseScaling", C.c_float), ("densityScaling", C.c_float)]


class ParametersPOD(C.Structure):
    _fields_ = [("maxRayBounces", C.c_uint32), ("lightIntensity", C.c_float), ("phaseG", C.c_float),
                ("phaseOnePlusG2", C.c_float), ("phaseOneMinusG2", C.c_float), ("phaseOneOver2G", C.c_float),
                ("isotropic", C.c_uint8), ("_pad", C.c_uint8 * 3)]


class ShadowVolumePropertiesPOD(C.Structure):
    _fields_ = [("bounds", (C.c_float * 4) * 2), ("lightDirection", C.c_float * 4), ("right", C.c_float * 4),
                ("up", C.c_float * 4), ("basisChange", (C.c_float * 4) * 4), ("voxelAxisCount", C.c_uint32),
                ("voxelSize", C.c_float)]


class PhotonPOD(C.Structure):
    """UniformBuffers.h:203-210 / PPM_PT.comp:14-21"""
    _fields_ = [("position", C.c_float * 4), ("power", C.c_float * 4), ("phi", C.c_float), ("theta", C.c_float),
                ("_padding", C.c_float * 2)]


class PhotonMapPropertiesPOD(C.Structure):
    """UniformBuffers.h:212-249"""
    _fields_ = [("lightDirection", C.c_float * 4), ("bounds", (C.c_float * 4) * 2), ("voxelCount", C.c_int32 * 3),
                ("voxelSize", C.c_float), ("photonSize", C.c_uint32), ("stepSize", C.c_float), ("absorption", C.c_float)]


class PTExtensionsPOD(C.Structure):
    _fields_ = [("absorption", C.c_float), ("math_mode", C.c_uint32), ("reserved", C.c_uint32 * 2)]


class PTStats(C.Structure):
    _fields_ = [("paths", C.c_uint64), ("tracking_steps", C.c_uint64), ("collisions", C.c_uint64),
                ("density_fetches", C.c_uint64), ("kernel_launches", C.c_uint64), ("tracked_paths", C.c_uint64)]


class PixelLists(C.Structure):
    _fields_ = [("tracked_xy", C.c_void_p), ("other_pixels", C.c_void_p), ("counts", C.c_void_p), ("width", C.c_uint32), ("reserved", C.c_uint32)]


class PushTargets(C.Structure):
    _fields_ = [("inbox", C.c_void_p * 16), ("flag", C.c_void_p * 16), ("done", C.c_void_p), ("nranks", C.c_uint32), ("rank", C.c_uint32),
                ("epoch", C.c_uint32), ("reserved", C.c_uint32)]


class GroupHandle(C.Structure):
    _fields_ = [("pid", C.c_uint64), ("device", C.c_int32), ("rank", C.c_uint32), ("base", C.c_uint64), ("bytes", C.c_uint64),
                ("ipc", C.c_ubyte * 64)]


assert C.sizeof(GroupHandle) == 96
assert C.sizeof(PushConstantsPOD) == 24
assert C.sizeof(CameraPropertiesPOD) == 68
assert C.sizeof(CloudPropertiesPOD) == 60
assert C.sizeof(ParametersPOD) == 28
assert C.sizeof(ShadowVolumePropertiesPOD) == 152
assert C.sizeof(PhotonPOD) == 48 and C.sizeof(PhotonMapPropertiesPOD) == 76

_vp = C.c_void_p
_u32 = C.c_uint32
_i32 = C.c_int32
_fp = C.POINTER(C.c_float)
_P = C.POINTER

# name -> (restype, argtypes); every symbol include/crv_c_api.h declares (tests/test_capi_symbols.py checks it)
SIGNATURES = {
    "crv_version": (C.c_char_p, []),
    "crv_last_error": (C.c_char_p, []),
    "crv_device_count": (C.c_int, []),
    "crv_ctx_create": (C.c_int, [C.c_int, _P(_vp)]),
    "crv_ctx_destroy": (C.c_int, [_vp]),
    "crv_ctx_synchronize": (C.c_int, [_vp, _vp]),
    "crv_ctx_launch_count": (C.c_int, [_vp, _P(C.c_uint64)]),
    "crv_ctx_sm_count": (C.c_int, [_vp, _P(C.c_int)]),
    "crv_stream_create": (C.c_int, [_vp, _P(_vp)]),
    "crv_stream_destroy": (C.c_int, [_vp, _vp]),
    "crv_group_create": (C.c_int, [_vp, _u32, _u32, _u32, _u32, _u32, C.c_uint64, _P(_vp)]),
    "crv_group_staging_ptr": (C.c_int, [_vp, _u32, _P(_vp)]),
    "crv_group_staging_all_gather": (C.c_int, [_vp, _u32, _vp, C.c_uint64, _vp]),
    "crv_group_staging_release": (C.c_int, [_vp, _u32, _vp]),
    "crv_group_destroy": (C.c_int, [_vp]),
    "crv_group_export": (C.c_int, [_vp, _P(GroupHandle)]),
    "crv_group_connect": (C.c_int, [_vp, _P(GroupHandle), _u32]),
    "crv_group_image": (C.c_int, [_vp, _u32, _P(_vp)]),
    "crv_group_all_reduce_sum_scale": (C.c_int, [_vp, _u32, C.c_float, _vp]),
    "crv_group_set_timeout_ms": (C.c_int, [_vp, _u32]),
    "crv_group_all_reduce_lists_sum_scale": (C.c_int, [_vp, _u32, C.c_float, C.c_float, _P(PixelLists), _vp]),
    "crv_pt_pixel_lists": (C.c_int, [_vp, _P(PixelLists)]),
    "crv_group_push_begin": (C.c_int, [_vp, _u32, _P(PushTargets)]),
    "crv_group_push_finish": (C.c_int, [_vp, _u32, C.c_float, _P(PixelLists), _vp]),
    "crv_pt_render_batch_push": (C.c_int, [_vp, _P(_i32), C.c_uint, _P(PushTargets), C.c_float, _vp, C.c_uint]),
    "crv_group_status": (C.c_int, [_vp, _vp]),
    "crv_group_last_reduce_ms": (C.c_int, [_vp, _P(C.c_float)]),
    "crv_group_nccl_unique_id": (C.c_int, [_vp]),
    "crv_group_nccl_init": (C.c_int, [_vp, _vp]),
    "crv_group_nccl_all_reduce_sum_scale": (C.c_int, [_vp, _u32, C.c_float, _vp]),
    "crv_volume_create": (C.c_int, [_vp, _u32, _u32, _u32, C.c_int, _P(_vp)]),
    "crv_volume_destroy": (C.c_int, [_vp]),
    "crv_volume_upload": (C.c_int, [_vp, _vp, C.c_int, _vp]),
    "crv_volume_download": (C.c_int, [_vp, _vp, _vp]),
    "crv_volume_dims": (C.c_int, [_vp, _P(_u32), _P(_u32), _P(_u32), _P(C.c_int)]),
    "crv_volume_majorant": (C.c_int, [_vp, _P(C.c_float), _vp]),
    "crv_volume_device_ptr": (C.c_int, [_vp, _P(_vp)]),
    "crv_volume_mark_dirty": (C.c_int, [_vp]),
    "crv_image_create": (C.c_int, [_vp, _u32, _u32, _P(_vp)]),
    "crv_image_wrap": (C.c_int, [_vp, _vp, _u32, _u32, _P(_vp)]),
    "crv_image_destroy": (C.c_int, [_vp]),
    "crv_image_clear": (C.c_int, [_vp, _vp]),
    "crv_image_download": (C.c_int, [_vp, _vp, _vp]),
    "crv_image_upload": (C.c_int, [_vp, _vp, _vp]),
    "crv_image_device_ptr": (C.c_int, [_vp, _P(_vp)]),
    "crv_image_scale": (C.c_int, [_vp, C.c_float, _vp]),
    "crv_sv_create": (C.c_int, [_vp, _P(ShadowVolumePropertiesPOD), _P(PushConstantsPOD), _P(_vp)]),
    "crv_sv_destroy": (C.c_int, [_vp]),
    "crv_sv_set_frame_references": (C.c_int, [_vp, _vp]),
    "crv_sv_clear_frame_references": (C.c_int, [_vp]),
    "crv_sv_queue_update_cloud_data": (C.c_int, [_vp, _P(CloudPropertiesPOD)]),
    "crv_sv_queue_update_cloud_data_sampler": (C.c_int, [_vp, _vp]),
    "crv_sv_queue_update_camera_properties": (C.c_int, [_vp, _P(CameraPropertiesPOD)]),
    "crv_sv_queue_update_parameters": (C.c_int, [_vp, _P(ParametersPOD)]),
    "crv_sv_queue_update_shadow_volume": (C.c_int, [_vp, _P(ShadowVolumePropertiesPOD)]),
    "crv_sv_queue_update_shadow_volume_sampler": (C.c_int, [_vp, _vp]),
    "crv_sv_update_descriptor_sets": (C.c_int, [_vp]),
    "crv_sv_record_draw_commands": (C.c_int, [_vp, _vp, C.c_uint]),
    "crv_sv_last_build_ms": (C.c_int, [_vp, _P(C.c_float)]),
    "crv_sv_set_tuning": (C.c_int, [_vp, C.c_int]),
    "crv_sv_last_build_mode": (C.c_int, [_vp, _P(C.c_int)]),
    "crv_pt_create": (C.c_int, [_vp, _P(CameraPropertiesPOD), _P(PushConstantsPOD), _P(_vp)]),
    "crv_pt_destroy": (C.c_int, [_vp]),
    "crv_pt_set_frame_references": (C.c_int, [_vp, _P(_vp), C.c_uint]),
    "crv_pt_clear_frame_references": (C.c_int, [_vp]),
    "crv_pt_queue_update_cloud_data": (C.c_int, [_vp, _P(CloudPropertiesPOD), C.c_uint]),
    "crv_pt_queue_update_cloud_data_sampler": (C.c_int, [_vp, _vp, C.c_uint]),
    "crv_pt_queue_update_camera_properties": (C.c_int, [_vp, _P(CameraPropertiesPOD), C.c_uint]),
    "crv_pt_queue_update_parameters": (C.c_int, [_vp, _P(ParametersPOD), C.c_uint]),
    "crv_pt_queue_update_shadow_volume": (C.c_int, [_vp, _P(ShadowVolumePropertiesPOD), C.c_uint]),
    "crv_pt_queue_update_shadow_volume_sampler": (C.c_int, [_vp, _vp, C.c_uint]),
    "crv_pt_update_descriptor_sets": (C.c_int, [_vp]),
    "crv_pt_record_draw_commands": (C.c_int, [_vp, _vp, C.c_uint]),
    "crv_pt_set_extensions": (C.c_int, [_vp, _P(PTExtensionsPOD)]),
    "crv_pt_render_batch": (C.c_int, [_vp, _P(_i32), C.c_uint, _u32, C.c_int, _vp, C.c_uint]),
    "crv_pt_get_stats": (C.c_int, [_vp, _P(PTStats), _vp]),
    "crv_pt_reset_stats": (C.c_int, [_vp, _vp]),
    "crv_pt_last_render_ms": (C.c_int, [_vp, _P(C.c_float), _P(C.c_uint)]),
    "crv_pt_set_debug_counters": (C.c_int, [_vp, C.c_int]),
    "crv_pt_download_debug_counters": (C.c_int, [_vp, _vp, _vp, _vp]),
    "crv_pt_selftest_log": (C.c_int, [_vp, _P(C.c_uint64), _vp]),
    "crv_pt_rng_tap": (C.c_int, [_vp, _i32, _u32, _u32, _u32, _vp, _vp]),
    "crv_pt_set_tuning": (C.c_int, [_vp, C.c_int, C.c_int, C.c_uint64, C.c_int]),
    "crv_pt_set_scheduler": (C.c_int, [_vp, C.c_int]),
    "crv_pt_set_skip_level": (C.c_int, [_vp, C.c_int]),
    "crv_pt_set_coarse_shift": (C.c_int, [_vp, C.c_int]),
    "crv_pt_set_job_order": (C.c_int, [_vp, C.c_int]),
    "crv_pt_get_coarse_shift": (C.c_int, [_vp, _P(C.c_int)]),
    "crv_ppm_create": (C.c_int, [_vp, _P(PhotonMapPropertiesPOD), _P(PushConstantsPOD), C.c_float, _P(_vp)]),
    "crv_ppm_set_frame_references": (C.c_int, [_vp, _P(_vp), C.c_uint]),
    "crv_ppm_queue_update_camera_properties": (C.c_int, [_vp, _P(CameraPropertiesPOD)]),
    "crv_ppm_queue_update_shadow_volume": (C.c_int, [_vp, _P(ShadowVolumePropertiesPOD)]),
    "crv_ppm_queue_update_shadow_volume_sampler": (C.c_int, [_vp, _vp]),
    "crv_ppm_record_draw_commands": (C.c_int, [_vp, _vp, C.c_uint]),
    "crv_ppm_update_radius": (C.c_int, [_vp, C.c_uint]),
    "crv_ppm_record_photon_estimate": (C.c_int, [_vp, _vp, C.c_uint]),
    "crv_ppm_upload": (C.c_int, [_vp, _vp, _vp, _vp]),
    "crv_ppm_last_estimate_stats": (C.c_int, [_vp, _P(C.c_uint64), _P(C.c_float)]),
    "crv_ppm_destroy": (C.c_int, [_vp]),
    "crv_ppm_allocate_resources": (C.c_int, [_vp]),
    "crv_ppm_free_resources": (C.c_int, [_vp]),
    "crv_ppm_queue_update_cloud_data": (C.c_int, [_vp, _P(CloudPropertiesPOD)]),
    "crv_ppm_queue_update_cloud_data_sampler": (C.c_int, [_vp, _vp]),
    "crv_ppm_queue_update_parameters": (C.c_int, [_vp, _P(ParametersPOD)]),
    "crv_ppm_update_descriptor_sets": (C.c_int, [_vp]),
    "crv_ppm_record_photon_tracing": (C.c_int, [_vp, _vp, C.c_uint32]),
    "crv_ppm_download": (C.c_int, [_vp, _vp, _vp, _vp]),
    "crv_ppm_last_trace_stats": (C.c_int, [_vp, _P(C.c_uint64), _P(C.c_uint64), _P(C.c_uint64), _P(C.c_float)]),
    "crv_photon_map_init": (None, [_P(PhotonMapPropertiesPOD)]),
    "crv_photon_map_set_bounds": (None, [_P(PhotonMapPropertiesPOD), _vp]),
    "crv_photon_map_total_size": (C.c_uint32, [_P(PhotonMapPropertiesPOD)]),
    "crv_image_blit_srgb8": (C.c_int, [_vp, _vp, _vp]),
    "crv_image_save_png": (C.c_int, [_vp, C.c_char_p, _vp]),
    "crv_image_save_pfm": (C.c_int, [_vp, C.c_char_p, _vp]),
    "crv_image_save_checkpoint": (C.c_int, [_vp, C.c_uint32, C.c_int32, C.c_char_p, _vp]),
    "crv_image_load_checkpoint": (C.c_int, [_vp, _P(C.c_uint32), _P(C.c_int32), C.c_char_p, _vp]),
    "crv_write_png_rgba8": (C.c_int, [C.c_char_p, C.c_uint32, C.c_uint32, _vp]),
    "crv_push_constants_init": (None, [_P(PushConstantsPOD)]),
    "crv_camera_init": (None, [_P(CameraPropertiesPOD)]),
    "crv_camera_set_resolution": (None, [_P(CameraPropertiesPOD), C.c_int, C.c_int]),
    "crv_camera_set_rotation": (None, [_P(CameraPropertiesPOD), _fp]),
    "crv_camera_set_fov": (None, [_P(CameraPropertiesPOD), _fp]),
    "crv_cloud_init": (None, [_P(CloudPropertiesPOD)]),
    "crv_cloud_set_from_grid": (None, [_P(CloudPropertiesPOD), _P(_u32), _P(C.c_double), C.c_float]),
    "crv_parameters_init": (None, [_P(ParametersPOD)]),
    "crv_parameters_set_phase_g": (None, [_P(ParametersPOD), C.c_float]),
    "crv_shadow_init": (None, [_P(ShadowVolumePropertiesPOD)]),
    "crv_shadow_set_light_direction": (None, [_P(ShadowVolumePropertiesPOD), _fp]),
    "crv_shadow_set_origin": (None, [_P(ShadowVolumePropertiesPOD), _fp, _fp]),
    "crv_orthonormal_basis": (None, [_fp, _fp, _fp]),
    "crv_msvc_rand_sequence": (None, [_u32, C.c_uint, _P(_i32)]),
    "crv_grid_load": (C.c_int, [C.c_char_p, _P(_u32), _P(C.c_double), _vp, C.c_size_t]),
    "crv_grid_save": (C.c_int, [C.c_char_p, _P(_u32), _P(C.c_double), _vp]),
    "crv_grid_save_zfastest": (C.c_int, [C.c_char_p, _P(_u32), _P(C.c_double), _vp]),
    "crv_fbm_generate": (C.c_int, [_vp, _u32, _vp]),
}


class CrvError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"crv error {code}: {message}")
        self.code = code
        self.message = message


def _load() -> C.CDLL:
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C cloudrendering_vulkan_b200/csrc`). There is no CPU fallback for this path.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    return lib


lib = _load()


def check(rc: int) -> None:
    """Raise CrvError for a negative status (the C ABI never throws/aborts; this wrapper does)."""
    if rc != CRV_OK:
        raise CrvError(rc, lib.crv_last_error().decode("utf-8", "replace"))
