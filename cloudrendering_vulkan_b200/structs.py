"""Host parameter blocks of the hot path with the reference's setter names (UniformBuffers.h:6-167).

Each class is the ctypes POD (same bytes the shaders' UBOs hold) plus the member functions of the reference
struct; the arithmetic lives in include/crv/uniform_buffers.h behind the C ABI so that Python, C and C++
callers produce identical bytes.  Grid3D mirrors Grid3D.h (x-fastest storage, ".xyz" files).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi
from .capi import lib


def _f3(v):
    return (C.c_float * 3)(float(v[0]), float(v[1]), float(v[2]))


class PushConstants(capi.PushConstantsPOD):
    """UniformBuffers.h:6-13 (seed = 100, frameCount = 1 by default)."""

    def __init__(self):
        super().__init__()
        lib.crv_push_constants_init(C.byref(self))


class CameraProperties(capi.CameraPropertiesPOD):
    """UniformBuffers.h:15-69."""

    def __init__(self):
        super().__init__()
        lib.crv_camera_init(C.byref(self))

    def SetRotation(self, rotation):
        """rotation = [pitch_deg, yaw_deg]; clamped in place like the reference's glm::vec2& argument."""
        r = (C.c_float * 2)(float(rotation[0]), float(rotation[1]))
        lib.crv_camera_set_rotation(C.byref(self), r)
        rotation[0], rotation[1] = r[0], r[1]

    def GetWidth(self) -> int:
        return self.halfWidth * 2

    def GetHeight(self) -> int:
        return self.halfHeight * 2

    def SetResolution(self, width: int, height: int):
        lib.crv_camera_set_resolution(C.byref(self), int(width), int(height))

    def SetFOV(self, fov: float) -> float:
        """Returns the clamped fov (the reference clamps its float& argument to 60..150)."""
        f = C.c_float(fov)
        lib.crv_camera_set_fov(C.byref(self), C.byref(f))
        return f.value

    def GetNearPlane(self) -> float:
        return self.nearPlane

    def GetFwd(self):
        return tuple(self.forward)


class CloudProperties(capi.CloudPropertiesPOD):
    """UniformBuffers.h:71-78; SetFromGrid is SetCloudProperties (main.cpp:144-163)."""

    def __init__(self):
        super().__init__()
        lib.crv_cloud_init(C.byref(self))

    def SetFromGrid(self, counts, voxel_size, majorant: float):
        c = (C.c_uint32 * 3)(*[int(x) for x in counts])
        v = (C.c_double * 3)(*[float(x) for x in voxel_size])
        lib.crv_cloud_set_from_grid(C.byref(self), c, v, C.c_float(majorant))


class Parameters(capi.ParametersPOD):
    """UniformBuffers.h:80-112 (isotropic by default: g = 1e-8)."""

    def __init__(self):
        super().__init__()
        lib.crv_parameters_init(C.byref(self))

    def SetPhaseG(self, g: float):
        lib.crv_parameters_set_phase_g(C.byref(self), C.c_float(g))

    def GetPhaseG(self) -> float:
        return self.phaseG


class ShadowVolumeProperties(capi.ShadowVolumePropertiesPOD):
    """UniformBuffers.h:114-167."""

    def __init__(self):
        super().__init__()
        lib.crv_shadow_init(C.byref(self))

    def SetLightDirection(self, direction):
        lib.crv_shadow_set_light_direction(C.byref(self), _f3(direction))

    def SetOrigin(self, lower, upper):
        lib.crv_shadow_set_origin(C.byref(self), _f3(lower), _f3(upper))

    def GetLightDirection(self):
        return tuple(self.lightDirection)


class PhotonMapProperties(capi.PhotonMapPropertiesPOD):
    """UniformBuffers.h:212-249: a cube of voxelCount^3 cells over the cloud's largest extent."""

    def __init__(self):
        super().__init__()
        lib.crv_photon_map_init(C.byref(self))

    def SetBounds(self, bounds):
        """bounds: the CloudProperties.bounds array (2 x vec4), as main.cpp:162 passes it"""
        b = ((C.c_float * 4) * 2)()
        for i in range(2):
            for j in range(4):
                b[i][j] = float(bounds[i][j])
        lib.crv_photon_map_set_bounds(C.byref(self), C.cast(b, C.c_void_p))

    def GetTotalSize(self) -> int:
        return int(lib.crv_photon_map_total_size(C.byref(self)))


class PathTracerExtensions(capi.PTExtensionsPOD):
    """Not in the reference: absorption (1 - albedo); 0 keeps the reference's RNG stream untouched.
    math_mode: capi.PT_MATH_STRICT (default) or capi.PT_MATH_DRIVER (crv_c_api.h: crv_pt_math_mode)."""

    def __init__(self, absorption: float = 0.0, math_mode: int = 0):
        super().__init__()
        self.absorption = absorption
        self.math_mode = math_mode


def msvc_rand_sequence(n: int, seed0: int = 0) -> np.ndarray:
    """Frame seeds: std::rand() after srand(seed0) with the MSVC LCG (main.cpp:722,949)."""
    out = np.empty(n, dtype=np.int32)
    lib.crv_msvc_rand_sequence(C.c_uint32(seed0), C.c_uint(n), out.ctypes.data_as(C.POINTER(C.c_int32)))
    return out


def orthonormal_basis(n):
    r, u = (C.c_float * 3)(), (C.c_float * 3)()
    lib.crv_orthonormal_basis(_f3(n), r, u)
    return np.array(r[:], np.float32), np.array(u[:], np.float32)


class Grid3D:
    """Grid3D<float> (Grid3D.h:5-192): x-fastest dense grid + the ".xyz" file format."""

    def __init__(self, size_x: int, size_y: int, size_z: int, voxel_size_x=1.0, voxel_size_y=1.0, voxel_size_z=1.0):
        self.count = (int(size_x), int(size_y), int(size_z))
        self.voxel_size = (float(voxel_size_x), float(voxel_size_y), float(voxel_size_z))
        self.data = np.zeros(self.count[0] * self.count[1] * self.count[2], dtype=np.float32)

    @staticmethod
    def Load(filename: str):
        """Returns None when the file cannot be opened, like the reference's nullptr (Grid3D.h:54-58)."""
        counts = (C.c_uint32 * 3)()
        vs = (C.c_double * 3)()
        rc = lib.crv_grid_load(filename.encode(), counts, vs, None, 0)
        if rc == capi.CRV_ERR_IO:
            return None
        capi.check(rc)
        g = Grid3D(counts[0], counts[1], counts[2], vs[0], vs[1], vs[2])
        capi.check(lib.crv_grid_load(filename.encode(), counts, vs, g.data.ctypes.data_as(C.c_void_p), g.data.size))
        return g

    def __call__(self, x: int, y: int, z: int) -> float:
        nx, ny, nz = self.count
        if 0 <= x < nx and 0 <= y < ny and 0 <= z < nz:
            return float(self.data[x + y * nx + z * nx * ny])
        raise IndexError("Index out of 3D grid bounds")          # std::out_of_range (Grid3D.h:119)

    def set(self, x: int, y: int, z: int, value: float):
        nx, ny, nz = self.count
        if not (0 <= x < nx and 0 <= y < ny and 0 <= z < nz):
            raise IndexError("Index out of 3D grid bounds")
        self.data[x + y * nx + z * nx * ny] = value

    def GetByteSize(self) -> int:
        return self.data.nbytes

    def GetElementSize(self) -> int:
        return 4

    def GetSize(self) -> int:
        return self.data.size

    def GetVoxelSize(self):
        return self.voxel_size

    def GetVoxelCount(self):
        return self.count

    def GetData(self) -> np.ndarray:
        return self.data

    def _hdr(self):
        return (C.c_uint32 * 3)(*self.count), (C.c_double * 3)(*self.voxel_size)

    def Save(self, filename: str):
        """Header + raw x-fastest data, exactly as Grid3D::Save (Grid3D.h:167-178; not the inverse of Load)."""
        c, v = self._hdr()
        capi.check(lib.crv_grid_save(filename.encode(), c, v, self.data.ctypes.data_as(C.c_void_p)))

    def SaveZFastest(self, filename: str):
        """GridData::save convention (GridData.h:12-25): z-fastest on disk, what Load expects."""
        c, v = self._hdr()
        capi.check(lib.crv_grid_save_zfastest(filename.encode(), c, v, self.data.ctypes.data_as(C.c_void_p)))

    def GetMajorant(self) -> float:
        """max over all voxels starting from 0 (Grid3D.h:181-192)."""
        return float(np.fmax.reduce(self.data, initial=np.float32(0))) if self.data.size else 0.0
