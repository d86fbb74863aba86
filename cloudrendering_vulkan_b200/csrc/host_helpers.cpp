// host_helpers.cpp -- C entry points for the host-side data model of the hot path: the parameter-struct
// setters of UniformBuffers.h (through include/crv/uniform_buffers.h), SetCloudProperties (main.cpp:144-163),
// the MSVC rand() frame-seed sequence (main.cpp:722,949) and the Grid3D ".xyz" file format (Grid3D.h,
// GridData.h).  Pure host code: compiled by g++ with -ffp-contract=off, no CUDA in this file.
#include <cstring>
#include <memory>
#include <new>

#include "../../include/crv/grid3d.h"
#include "../../include/crv/uniform_buffers.h"

void crv_set_error(const char* fmt, ...);

extern "C" {

void crv_push_constants_init(crv_push_constants* p) { *p = crv::PushConstants(); }
void crv_camera_init(crv_camera_properties* c) { *c = crv::CameraProperties(); }
void crv_camera_set_resolution(crv_camera_properties* c, int width, int height)
{
    static_cast<crv::CameraProperties*>(c)->SetResolution(width, height);
}
void crv_camera_set_rotation(crv_camera_properties* c, float* rotation_xy)
{
    static_cast<crv::CameraProperties*>(c)->SetRotation(rotation_xy);
}
void crv_camera_set_fov(crv_camera_properties* c, float* fov) { static_cast<crv::CameraProperties*>(c)->SetFOV(*fov); }
void crv_cloud_init(crv_cloud_properties* c) { *c = crv::CloudProperties(); }
void crv_cloud_set_from_grid(crv_cloud_properties* c, const uint32_t counts[3], const double voxel_size[3], float majorant)
{
    static_cast<crv::CloudProperties*>(c)->SetFromGrid(counts, voxel_size, majorant);
}
void crv_parameters_init(crv_parameters* p) { *p = crv::Parameters(); }
void crv_parameters_set_phase_g(crv_parameters* p, float g) { static_cast<crv::Parameters*>(p)->SetPhaseG(g); }
void crv_shadow_init(crv_shadow_volume_properties* s) { *s = crv::ShadowVolumeProperties(); }
void crv_shadow_set_light_direction(crv_shadow_volume_properties* s, const float dir[3])
{
    static_cast<crv::ShadowVolumeProperties*>(s)->SetLightDirection(dir);
}
void crv_shadow_set_origin(crv_shadow_volume_properties* s, const float lower[3], const float upper[3])
{
    static_cast<crv::ShadowVolumeProperties*>(s)->SetOrigin(lower, upper);
}
void crv_orthonormal_basis(const float n[3], float right[3], float up[3])
{
    crv::detail::V3 r, u;
    crv::detail::orthonormal_basis(crv::detail::V3{n[0], n[1], n[2]}, r, u);
    right[0] = r.x; right[1] = r.y; right[2] = r.z;
    up[0] = u.x; up[1] = u.y; up[2] = u.z;
}
void crv_msvc_rand_sequence(uint32_t seed0, unsigned n, int32_t* out) { crv::MsvcRandSequence(seed0, n, out); }

// ---- Grid3D .xyz -----------------------------------------------------------------------------------
int crv_grid_load(const char* path, uint32_t counts[3], double voxel_size[3], float* data, size_t capacity)
{
    if (!path || !counts || !voxel_size) { crv_set_error("crv_grid_load: null argument"); return CRV_ERR_INVALID; }
    if (!data) {
        // header-only query
        std::ifstream in(path, std::ifstream::in | std::ifstream::binary);
        if (!in.good()) { crv_set_error("crv_grid_load: cannot open '%s'", path); return CRV_ERR_IO; }
        in.read(reinterpret_cast<char*>(counts), sizeof(uint32_t) * 3);
        in.read(reinterpret_cast<char*>(voxel_size), sizeof(double) * 3);
        if (!in.good()) { crv_set_error("crv_grid_load: '%s' is shorter than the 36-byte header", path); return CRV_ERR_IO; }
        return CRV_OK;
    }
    try {
        std::unique_ptr<crv::Grid3D<float>> g(crv::Grid3D<float>::Load(path));
        if (!g) { crv_set_error("crv_grid_load: cannot open '%s'", path); return CRV_ERR_IO; }   // Load returns nullptr (Grid3D.h:54-58)
        const auto c = g->GetVoxelCount();
        const auto v = g->GetVoxelSize();
        for (int i = 0; i < 3; ++i) { counts[i] = c[i]; voxel_size[i] = v[i]; }
        if (g->GetSize() > capacity) { crv_set_error("crv_grid_load: buffer holds %zu floats, grid has %zu", capacity, g->GetSize()); return CRV_ERR_INVALID; }
        std::memcpy(data, g->data(), g->GetByteSize());
    } catch (const std::bad_alloc&) {
        crv_set_error("crv_grid_load: out of host memory");
        return CRV_ERR_OOM;
    }
    return CRV_OK;
}

static int grid_save(const char* path, const uint32_t counts[3], const double vs[3], const float* data, bool zfast)
{
    if (!path || !counts || !vs || !data) { crv_set_error("crv_grid_save: null argument"); return CRV_ERR_INVALID; }
    try {
        crv::Grid3D<float> g(counts[0], counts[1], counts[2], vs[0], vs[1], vs[2]);
        std::memcpy(g.GetData(), data, g.GetByteSize());
        if (zfast) g.SaveZFastest(path); else g.Save(path);
    } catch (const std::bad_alloc&) {
        crv_set_error("crv_grid_save: out of host memory");
        return CRV_ERR_OOM;
    }
    std::ifstream chk(path, std::ifstream::in | std::ifstream::binary);
    if (!chk.good()) { crv_set_error("crv_grid_save: cannot write '%s'", path); return CRV_ERR_IO; }
    return CRV_OK;
}

int crv_grid_save(const char* path, const uint32_t counts[3], const double voxel_size[3], const float* data)
{
    return grid_save(path, counts, voxel_size, data, false);
}
int crv_grid_save_zfastest(const char* path, const uint32_t counts[3], const double voxel_size[3], const float* data_xfastest)
{
    return grid_save(path, counts, voxel_size, data_xfastest, true);
}

}  // extern "C"
