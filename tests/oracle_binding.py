"""ctypes binding of the CPU oracle (oracle/libcrv_oracle.so).  TEST INFRASTRUCTURE: only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this module."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB_PATH = os.path.join(ORACLE_DIR, "libcrv_oracle.so")
REF_LIB_PATH = os.path.join(ORACLE_DIR, "_ref", "libcrv_refstructs.so")


def _ensure_built():
    if not os.path.exists(LIB_PATH):
        env = dict(os.environ)
        env.pop("CXX", None)
        subprocess.run(["make", "libcrv_oracle.so"], cwd=ORACLE_DIR, env=env, check=True, stdout=subprocess.DEVNULL)


_ensure_built()
lib = C.CDLL(LIB_PATH)

_vp, _u32, _i32, _f = C.c_void_p, C.c_uint32, C.c_int32, C.c_float
_fp = C.POINTER(C.c_float)
lib.orc_rng_seed_state.restype = _u32
lib.orc_rng_seed_state.argtypes = [_i32, _u32, _u32, _u32]
lib.orc_rng_warmup_count.restype = _u32
lib.orc_rng_warmup_count.argtypes = [_u32]
lib.orc_rng_draws.argtypes = [_i32, _u32, _u32, _u32, _u32, _vp, _vp]
lib.orc_rng_tap.argtypes = [_i32, _u32, _u32, _u32, _vp]
lib.orc_camera_ray.argtypes = [_vp, C.c_int, C.c_int, _fp, _fp]
lib.orc_intersect_cloud.restype = C.c_int
lib.orc_intersect_cloud.argtypes = [_vp, _fp, _fp, _fp, _fp]
lib.orc_orthonormal_basis.argtypes = [_fp, _fp, _fp]
lib.orc_calculate_pdf.restype = _f
lib.orc_calculate_pdf.argtypes = [_vp, _f]
lib.orc_invert_cdf.restype = _f
lib.orc_invert_cdf.argtypes = [_vp, _f]
lib.orc_sample_phase.restype = _f
lib.orc_sample_phase.argtypes = [_vp, _fp, _fp]
lib.orc_background.argtypes = [_fp, _fp]
lib.orc_blit_srgb8.argtypes = [_vp, C.c_size_t, _vp]
lib.orc_ppm_estimate.restype = C.c_uint64
lib.orc_ppm_estimate.argtypes = [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _u32, _f, _vp, _vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int]
lib.orc_ppm_trace.restype = C.c_uint64
lib.orc_ppm_trace.argtypes = [_vp, _vp, _vp, _vp, _i32, _u32, _vp, _vp, C.c_uint64, C.c_int]
lib.orc_texture3d.restype = _f
lib.orc_texture3d.argtypes = [_vp, _u32, _u32, _u32, _fp]
lib.orc_sample_cloud.restype = _f
lib.orc_sample_cloud.argtypes = [_vp, _vp, _fp]
lib.orc_sample_shadow_volume.restype = _f
lib.orc_sample_shadow_volume.argtypes = [_vp, _vp, _fp]
lib.orc_scatter_ray.argtypes = [_vp, C.POINTER(_u32), _fp, _fp]
lib.orc_shadow_volume.argtypes = [_vp, _vp, _vp, _vp, C.c_int]
lib.orc_shadow_volume_columns.argtypes = [_vp, _vp, _vp, _vp, _u32, _vp, _u32, _vp, C.c_int]
lib.orc_path_trace_frame.argtypes = [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _f, C.c_int, _vp, _vp, _vp, _vp, C.c_int]
lib.orc_path_trace_rect.argtypes = [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _f, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                    _vp, _vp, _vp, _vp, C.c_int]
lib.orc_path_trace_rows.argtypes = [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _f, C.c_int, _vp, _vp, _vp, _vp, C.c_int]
lib.orc_fbm_generate.argtypes = [_u32, _u32, _vp, C.c_int]
lib.orc_max_threads.restype = C.c_int


def _f3(v):
    return (C.c_float * 3)(float(v[0]), float(v[1]), float(v[2]))


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _raw(struct):
    """address of a ctypes struct (the oracle reads raw bytes at the shaders' std140 offsets)"""
    return C.cast(C.byref(struct), C.c_void_p)


def max_threads() -> int:
    return lib.orc_max_threads()


def rng_draws(seed, gx, gy, ngx, n):
    st = np.empty(n, np.uint32)
    fl = np.empty(n, np.float32)
    lib.orc_rng_draws(seed, gx, gy, ngx, n, _ptr(st), _ptr(fl))
    return st, fl


def rng_tap(seed, width, height, n_draws):
    out = np.empty((height, width, n_draws), np.uint32)
    lib.orc_rng_tap(seed, width, height, n_draws, _ptr(out))
    return out


def camera_ray(cam, x, y):
    p, d = (C.c_float * 3)(), (C.c_float * 3)()
    lib.orc_camera_ray(_raw(cam), x, y, p, d)
    return np.array(p[:], np.float32), np.array(d[:], np.float32)


def intersect_cloud(cloud, pos, direction):
    tmax, tmin = C.c_float(), C.c_float()
    hit = lib.orc_intersect_cloud(_raw(cloud), _f3(pos), _f3(direction), C.byref(tmax), C.byref(tmin))
    return bool(hit), tmax.value, tmin.value


def orthonormal_basis(n):
    a, b = (C.c_float * 3)(), (C.c_float * 3)()
    lib.orc_orthonormal_basis(_f3(n), a, b)
    return np.array(a[:], np.float32), np.array(b[:], np.float32)


def calculate_pdf(params, c):
    return lib.orc_calculate_pdf(_raw(params), c)


def invert_cdf(params, xi):
    return lib.orc_invert_cdf(_raw(params), xi)


def sample_phase(params, incoming, sample):
    return lib.orc_sample_phase(_raw(params), _f3(incoming), _f3(sample))


def background(direction):
    out = (C.c_float * 4)()
    lib.orc_background(_f3(direction), out)
    return np.array(out[:], np.float32)


def ppm_trace(cloud, cloud_props, params, pm_props, seed, groups_x=200, threads=0):
    """one photon-tracing dispatch of PPM_PT.comp (restatement) -> (photons (cells,32,12) f32, collisions (cells,4) u32, steps)"""
    cells = int(np.prod(np.frombuffer(bytes(pm_props), np.int32, 3, 48)))
    ph = np.empty((cells, 32, 12), np.float32)
    co = np.empty((cells, 4), np.uint32)
    steps = lib.orc_ppm_trace(_ptr(np.ascontiguousarray(cloud, np.float32)), _raw(cloud_props), _raw(params), _raw(pm_props), int(seed),
                              int(groups_x), _ptr(ph), _ptr(co), cells, threads)
    return ph, co, int(steps)


def refsh_ppm_trace(cloud, cloud_props, params, pm_props, seed, groups_x=200, threads=0):
    """the same dispatch through the reference's own PPM_PT.comp compiled by g++ (oracle/_ref)"""
    rl = ref_shaders_lib()
    cells = int(np.prod(np.frombuffer(bytes(pm_props), np.int32, 3, 48)))
    nx, ny, nz = _cloud_dims(cloud_props)
    ph = np.empty((cells, 32, 12), np.float32)
    co = np.empty((cells, 4), np.uint32)
    rl.refsh_ppm_trace.argtypes = [_vp, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp, _vp, _u32, _vp, _vp, C.c_uint64, C.c_int]
    rl.refsh_ppm_block_size.restype = _u32
    rl.refsh_ppm_photon_size.restype = _u32
    rl.refsh_ppm_trace(_ptr(np.ascontiguousarray(cloud, np.float32)), nx, ny, nz, _raw(cloud_props), _raw(params), _raw(pm_props),
                       push_constants_bytes(seed, 0), int(groups_x), _ptr(ph), _ptr(co), cells, threads)
    return ph, co


def ppm_estimate(cloud, sv, cam, cloud_props, params, pm_props, sv_props, seed, frame_count, pm_radius, photons, collisions, image,
                 rows=None, row_step=1, threads=0):
    """one photon-estimate dispatch of PPM_PE.comp (restatement) over a given photon map; accumulates into image (H,W,4)"""
    w, h = image_size(cam)
    y0, y1 = rows if rows is not None else (0, h)
    assert image.dtype == np.float32 and image.flags["C_CONTIGUOUS"] and image.size == w * h * 4
    return int(lib.orc_ppm_estimate(_ptr(np.ascontiguousarray(cloud, np.float32)), _ptr(np.ascontiguousarray(sv, np.float32)), _raw(cam),
                                    _raw(cloud_props), _raw(params), _raw(pm_props), _raw(sv_props), int(seed), int(frame_count),
                                    float(pm_radius), _ptr(np.ascontiguousarray(photons, np.float32)),
                                    _ptr(np.ascontiguousarray(collisions, np.uint32)), _ptr(image), int(y0), int(y1), int(row_step), threads))


def refsh_ppm_estimate(cloud, sv, cam, cloud_props, params, pm_props, sv_props, seed, frame_count, pm_radius, photons, collisions, image,
                       rows=None, row_step=1, threads=0):
    """the same dispatch through the reference's own PPM_PE.comp compiled by g++ (oracle/_ref)"""
    import struct
    rl = ref_shaders_lib()
    w, h = image_size(cam)
    y0, y1 = rows if rows is not None else (0, h)
    nx, ny, nz = _cloud_dims(cloud_props)
    n = int(np.frombuffer(bytes(sv_props), np.uint32, 1, 144)[0])
    push = (C.c_ubyte * 24).from_buffer_copy(struct.pack("<diIfI", 0.0, int(seed), int(frame_count), float(pm_radius), 0))
    rl.refsh_ppm_estimate.argtypes = [_vp, C.c_int, C.c_int, C.c_int, _vp, C.c_int, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                      C.c_int, C.c_int, C.c_int, C.c_int]
    rl.refsh_ppm_estimate(_ptr(np.ascontiguousarray(cloud, np.float32)), nx, ny, nz, _ptr(np.ascontiguousarray(sv, np.float32)), n, _raw(cam),
                          _raw(cloud_props), _raw(params), _raw(pm_props), _raw(sv_props), push,
                          _ptr(np.ascontiguousarray(photons, np.float32)), _ptr(np.ascontiguousarray(collisions, np.uint32)), _ptr(image),
                          int(y0), int(y1), int(row_step), threads)


def blit_srgb8(rgba):
    """the reference's present step (vkCmdBlitImage to B8G8R8A8_SRGB): HxWx4 float -> HxWx4 bytes B,G,R,A"""
    a = np.ascontiguousarray(rgba, np.float32)
    out = np.empty(a.shape, np.uint8)
    lib.orc_blit_srgb8(_ptr(a), a.size // 4, _ptr(out))
    return out


def texture3d(vol, nx, ny, nz, uvw):
    v = np.ascontiguousarray(vol, np.float32)
    return lib.orc_texture3d(_ptr(v), nx, ny, nz, _f3(uvw))


def sample_cloud(vol, cloud, pos):
    v = np.ascontiguousarray(vol, np.float32)
    return lib.orc_sample_cloud(_ptr(v), _raw(cloud), _f3(pos))


def sample_shadow_volume(sv, svprops, pos):
    v = np.ascontiguousarray(sv, np.float32)
    return lib.orc_sample_shadow_volume(_ptr(v), _raw(svprops), _f3(pos))


def scatter_ray(params, rng_state, incoming):
    st = _u32(rng_state)
    out = (C.c_float * 3)()
    lib.orc_scatter_ray(_raw(params), C.byref(st), _f3(incoming), out)
    return np.array(out[:], np.float32), st.value


def shadow_volume(cloud, cloud_props, sv_props, threads=0):
    n = int(np.frombuffer(bytes(sv_props), np.uint32, 1, 144)[0])
    v = np.ascontiguousarray(cloud, np.float32)
    out = np.empty(n * n * n, np.float32)
    lib.orc_shadow_volume(_ptr(v), _raw(cloud_props), _raw(sv_props), _ptr(out), threads)
    return out


def shadow_volume_columns(cloud, cloud_props, sv_props, xs, ys, threads=0):
    """ShadowVolume.comp for the columns (x, y) in xs x ys only: array [len(ys), len(xs), N] (z last)"""
    n = int(np.frombuffer(bytes(sv_props), np.uint32, 1, 144)[0])
    v = cloud if (isinstance(cloud, np.ndarray) and cloud.dtype == np.float32 and cloud.flags["C_CONTIGUOUS"]) else np.ascontiguousarray(cloud, np.float32)
    xa, ya = np.ascontiguousarray(xs, np.uint32), np.ascontiguousarray(ys, np.uint32)
    out = np.empty((len(ya), len(xa), n), np.float32)
    lib.orc_shadow_volume_columns(_ptr(v), _raw(cloud_props), _raw(sv_props), _ptr(xa), len(xa), _ptr(ya), len(ya), _ptr(out), threads)
    return out


def path_trace_rows(cloud, sv, cam, cloud_props, params, sv_props, push, absorption, row_step, image=None,
                    want_counters=False, threads=0):
    """Rows 0, row_step, ... of one frame in ONE call (cloud/sv must already be contiguous float32).
    Returns (n_paths, steps, collisions); counters are None unless requested."""
    w, h = image_size(cam)
    steps = np.zeros((h, w), np.uint32) if want_counters else None
    coll = np.zeros((h, w), np.uint32) if want_counters else None
    lib.orc_path_trace_rows(_ptr(cloud), _ptr(sv), _raw(cam), _raw(cloud_props), _raw(params), _raw(sv_props), _raw(push),
                            absorption, int(row_step), _ptr(image), None, _ptr(steps), _ptr(coll), threads)
    return w * len(range(0, h, int(row_step))), steps, coll


def path_trace_frame(cloud, sv, cam, cloud_props, params, sv_props, push, absorption, image, accum_sum=0,
                     want_sample=True, threads=0, rect=None):
    """One dispatch of PathTracer.comp on the CPU. `image` (H,W,4) is updated in place.
    Returns (sample, steps, collisions)."""
    w, h = image_size(cam)
    v = np.ascontiguousarray(cloud, np.float32)
    s = np.ascontiguousarray(sv, np.float32)
    sample = np.zeros((h, w, 4), np.float32) if want_sample else None
    steps = np.zeros((h, w), np.uint32)
    coll = np.zeros((h, w), np.uint32)
    if image is not None:
        assert image.dtype == np.float32 and image.flags["C_CONTIGUOUS"] and image.size == w * h * 4
    if rect is None:
        lib.orc_path_trace_frame(_ptr(v), _ptr(s), _raw(cam), _raw(cloud_props), _raw(params), _raw(sv_props), _raw(push),
                                 absorption, accum_sum, _ptr(image), _ptr(sample), _ptr(steps), _ptr(coll), threads)
    else:
        x0, y0, x1, y1 = rect
        lib.orc_path_trace_rect(_ptr(v), _ptr(s), _raw(cam), _raw(cloud_props), _raw(params), _raw(sv_props), _raw(push),
                                absorption, accum_sum, x0, y0, x1, y1, _ptr(image), _ptr(sample), _ptr(steps), _ptr(coll), threads)
    return sample, steps, coll


def fbm_generate(n, seed=1234, threads=0):
    out = np.empty(n * n * n, np.float32)
    lib.orc_fbm_generate(n, seed, _ptr(out), threads)
    return out


def ref_structs_lib():
    """The shim around the reference's own headers (oracle/_ref), or None when it was never built."""
    if not os.path.exists(REF_LIB_PATH):
        return None
    return C.CDLL(REF_LIB_PATH)


# ---- scene setup without any product code: reference headers (oracle/_ref) + main.cpp restatements ----------
lib.orc_set_cloud_properties.argtypes = [C.POINTER(_u32), C.POINTER(C.c_double), _f, _f, _vp]
lib.orc_msvc_rand.argtypes = [_u32, _u32, _vp]


class RawStruct(C.Structure):
    """a parameter block as raw bytes (what the UBO holds); the oracle reads it at the shaders' offsets"""
    pass


def _raw_buf(nbytes):
    return (C.c_ubyte * nbytes)()


def msvc_rand(n, seed0=0):
    out = np.empty(n, np.int32)
    lib.orc_msvc_rand(seed0, n, _ptr(out))
    return out


def image_size(cam):
    b = bytes(cam)
    return 2 * int(np.frombuffer(b, np.int32, 1, 12)[0]), 2 * int(np.frombuffer(b, np.int32, 1, 28)[0])


def push_constants_bytes(seed, frame_count):
    import struct
    return (C.c_ubyte * 24).from_buffer_copy(struct.pack("<diIfI", 0.0, int(seed), int(frame_count), 0.0, 0))


def reference_scene_structs(n_cloud, width, height, fov, n_sv, light, density_scaling, phase_g, majorant):
    """(camera68, cloud60, params28, shadow152) byte buffers made by the REFERENCE's own setters (oracle/_ref) and the
    restated SetCloudProperties; raises if the shim was never built."""
    ref = ref_structs_lib()
    if ref is None:
        raise RuntimeError("oracle/_ref/libcrv_refstructs.so missing: run `make -C oracle` where /root/reference exists")
    cam = _raw_buf(68)
    f = C.c_float(fov)
    ref.ref_camera(1, int(width), int(height), 0, (C.c_float * 2)(0, 0), 1, C.byref(f), cam)
    cloud = _raw_buf(60)
    lib.orc_set_cloud_properties((_u32 * 3)(n_cloud, n_cloud, n_cloud), (C.c_double * 3)(*([1.0 / n_cloud] * 3)),
                                 _f(majorant), _f(density_scaling), cloud)
    par = _raw_buf(28)
    ref.ref_parameters(0 if phase_g is None else 1, C.c_float(phase_g or 0.0), par)
    b = np.frombuffer(bytes(cloud), np.float32)
    sh = _raw_buf(152)
    ref.ref_shadow(int(n_sv), 1, _f3(light), 1, _f3(b[0:3]), _f3(b[4:7]), sh)
    return cam, cloud, par, sh


# ---- the reference's OWN shaders compiled by g++ (oracle/_ref/libcrv_refshaders.so; see oracle/ref_shaders_host.cpp) ----
REF_SHADERS_PATH = os.path.join(ORACLE_DIR, "_ref", "libcrv_refshaders.so")
_refsh = None


def ref_shaders_lib():
    """PathTracer.comp / ShadowVolume.comp themselves, compiled where they lie; None when never built."""
    global _refsh
    if _refsh is None and os.path.exists(REF_SHADERS_PATH):
        _refsh = C.CDLL(REF_SHADERS_PATH)
        _refsh.refsh_path_trace.argtypes = [_vp, C.c_int, C.c_int, C.c_int, _vp, C.c_int, _vp, _vp, _vp, _vp, _vp, _vp,
                                            C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _vp, _vp, C.c_int]
        _refsh.refsh_shadow_volume.argtypes = [_vp, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp, C.c_int]
        _refsh.refsh_block_sizes.argtypes = [_vp]
        _refsh.refsh_max_threads.restype = C.c_int
    return _refsh


def refsh_block_sizes():
    out = np.zeros(7, np.uint32)
    ref_shaders_lib().refsh_block_sizes(_ptr(out))
    return [int(v) for v in out]


def _cloud_dims(cloud_props):
    return [int(v) for v in np.frombuffer(bytes(cloud_props), np.uint32, 3, 32)]


def refsh_shadow_volume(cloud, cloud_props, sv_props, threads=0):
    """one dispatch of the reference's ShadowVolume.comp on the CPU"""
    n = int(np.frombuffer(bytes(sv_props), np.uint32, 1, 144)[0])
    nx, ny, nz = _cloud_dims(cloud_props)
    v = np.ascontiguousarray(cloud, np.float32)
    out = np.empty(n * n * n, np.float32)
    ref_shaders_lib().refsh_shadow_volume(_ptr(v), nx, ny, nz, _raw(cloud_props), _raw(sv_props), _ptr(out), threads)
    return out


def refsh_path_trace(cloud, sv, cam, cloud_props, params, sv_props, push, image, rect=None, row_step=1, want_counters=True,
                     threads=0):
    """one dispatch of the reference's PathTracer.comp on the CPU; `image` (H,W,4) is accumulated in place by the
    shader itself.  Returns (steps, collisions) as counted by the compat layer (log() / texture(shadowSampler) calls)."""
    w, h = image_size(cam)
    n = int(np.frombuffer(bytes(sv_props), np.uint32, 1, 144)[0])
    nx, ny, nz = _cloud_dims(cloud_props)
    assert cloud.dtype == np.float32 and cloud.flags["C_CONTIGUOUS"] and sv.dtype == np.float32 and sv.flags["C_CONTIGUOUS"]
    assert image.dtype == np.float32 and image.flags["C_CONTIGUOUS"] and image.size == w * h * 4
    x0, y0, x1, y1 = rect if rect is not None else (0, 0, w, h)
    steps = np.zeros((h, w), np.uint32) if want_counters else None
    coll = np.zeros((h, w), np.uint32) if want_counters else None
    ref_shaders_lib().refsh_path_trace(_ptr(cloud), nx, ny, nz, _ptr(sv), n, _raw(cam), _raw(cloud_props), _raw(params),
                                       _raw(sv_props), _raw(push), _ptr(image), x0, y0, x1, y1, int(row_step),
                                       _ptr(steps), _ptr(coll), threads)
    return steps, coll
