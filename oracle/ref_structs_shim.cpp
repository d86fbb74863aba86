// oracle/ref_structs_shim.cpp -- TEST INFRASTRUCTURE, not product code.
//
// Thin extern "C" shim around the REAL reference headers, compiled where they lie
// (-I/root/reference/CloudRendering-Vulkan -I/root/reference/include).  It lets the tests
// compare the product's POD mirrors / setters (include/crv/uniform_buffers.h, grid3d.h)
// byte-for-byte against what the reference's own C++ produces:
//   UniformBuffers.h:6-167   PushConstants, CameraProperties, CloudProperties, Parameters,
//                            ShadowVolumeProperties (+ setters)
//   Grid3D.h:5-192           Grid3D<float> Load / Save / GetMajorant
// Output goes to oracle/_ref/libcrv_refstructs.so (git-ignored, travels with gpurun).
// Nothing from the reference is copied into this repository; the one function the headers
// leave undefined (utilities::GetOrthonormalBasis, Utilities.cpp:7-16, which cannot be
// compiled here because Utilities.cpp pulls in Vulkan/GLFW/ImGui through stdafx.h) is
// restated below and is itself cross-checked against PathTracer.comp:197-206.
#define GLM_ENABLE_EXPERIMENTAL
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <fstream>
#include <string>
#include <vector>

#include <glm/glm.hpp>
#include <glm/gtx/transform.hpp>

namespace utilities {
void GetOrthonormalBasis(const glm::vec3& fwd, glm::vec3& outRight, glm::vec3& outUp);
}

#include "UniformBuffers.h"
#include "Grid3D.h"

// Restatement of Utilities.cpp:7-16 (branchless tangent frame, same formula as the shader).
void utilities::GetOrthonormalBasis(const glm::vec3& n, glm::vec3& r, glm::vec3& u)
{
    const float s = (n.z >= 0.0f) ? 1.0f : -1.0f;
    const float a = n.y / (1.0f + std::abs(n.z));
    const float b = n.y * a;
    const float c = -n.x * a;
    r = glm::vec3(n.z + s * b, s * c, -n.x);
    u = glm::vec3(c, 1.0f - b, -s * n.y);
}

extern "C" {

// sizes: 0 PushConstants, 1 CameraProperties, 2 CloudProperties, 3 Parameters, 4 ShadowVolumeProperties
int ref_sizeof(int which)
{
    switch (which) {
    case 0: return (int)sizeof(PushConstants);
    case 1: return (int)sizeof(CameraProperties);
    case 2: return (int)sizeof(CloudProperties);
    case 3: return (int)sizeof(Parameters);
    case 4: return (int)sizeof(ShadowVolumeProperties);
    }
    return -1;
}

void ref_push_constants_default(void* out)
{
    PushConstants p;
    std::memset(out, 0, sizeof(p));
    std::memcpy(out, &p, sizeof(p));
}

// Default-construct, then (optionally) SetResolution, SetRotation, SetFOV in the order main.cpp uses
// (SetResolution main.cpp:383 -> SetFOV main.cpp:932/384).  fov/rot are in/out like the reference's references.
void ref_camera(int set_res, int w, int h, int set_rot, float* rot2, int set_fov, float* fov, void* out)
{
    CameraProperties c;
    if (set_res) c.SetResolution(w, h);
    if (set_rot) { glm::vec2 r(rot2[0], rot2[1]); c.SetRotation(r); rot2[0] = r.x; rot2[1] = r.y; }
    if (set_fov) c.SetFOV(*fov);
    std::memcpy(out, &c, sizeof(c));
}

void ref_cloud_default(void* out)
{
    CloudProperties c;
    std::memcpy(out, &c, sizeof(c));
}

// Parameters has 3 trailing padding bytes after `bool isotropic`; zero them like the reference's
// zero-initialised global (main.cpp:40) so that byte comparison is meaningful.
void ref_parameters(int set_g, float g, void* out)
{
    static Parameters p_static;          // static storage: padding is zero, as for g_parameters
    p_static = Parameters();
    if (set_g) p_static.SetPhaseG(g);
    std::memset(out, 0, sizeof(Parameters));
    std::memcpy(out, &p_static, offsetof(Parameters, lightIntensity) + 4 * 5 + 1);  // through isotropic's byte
}

// SetLightDirection followed by SetOrigin, the order UpdateShadowVolume uses (main.cpp:232-235).
void ref_shadow(unsigned n_axis, int set_light, const float* l3, int set_origin, const float* lo3,
                const float* hi3, void* out)
{
    ShadowVolumeProperties s;
    s.voxelAxisCount = n_axis;
    if (set_light) s.SetLightDirection(glm::vec3(l3[0], l3[1], l3[2]));
    if (set_origin) s.SetOrigin(glm::vec3(lo3[0], lo3[1], lo3[2]), glm::vec3(hi3[0], hi3[1], hi3[2]));
    std::memcpy(out, &s, sizeof(s));
}

void ref_basis(const float* n3, float* r3, float* u3)
{
    glm::vec3 r, u;
    utilities::GetOrthonormalBasis(glm::vec3(n3[0], n3[1], n3[2]), r, u);
    r3[0] = r.x; r3[1] = r.y; r3[2] = r.z;
    u3[0] = u.x; u3[1] = u.y; u3[2] = u.z;
}

// Grid3D<float>: Load a .xyz file through the reference loader; copies x-fastest data out.
// returns 0 ok, 1 load failed (nullptr), 2 buffer too small
int ref_grid_load(const char* path, unsigned* counts3, double* voxel3, float* data, size_t capacity,
                  float* majorant)
{
    Grid3D<float>* g = Grid3D<float>::Load(path);
    if (!g) return 1;
    glm::uvec3 c = g->GetVoxelCount();
    glm::dvec3 v = g->GetVoxelSize();
    counts3[0] = c.x; counts3[1] = c.y; counts3[2] = c.z;
    voxel3[0] = v.x; voxel3[1] = v.y; voxel3[2] = v.z;
    int rc = 0;
    if (g->GetSize() > capacity) rc = 2;
    else std::memcpy(data, g->GetData(), g->GetByteSize());
    *majorant = g->GetMajorant();
    delete g;
    return rc;
}

// Build a Grid3D from x-fastest data and Save() it with the reference writer.
int ref_grid_save(const char* path, unsigned nx, unsigned ny, unsigned nz, double vx, double vy,
                  double vz, const float* data)
{
    Grid3D<float> g(nx, ny, nz, vx, vy, vz);
    std::memcpy(g.GetData(), data, g.GetByteSize());
    g.Save(path);
    return 0;
}

// PhotonMapProperties / Photon (UniformBuffers.h:203-249) for the PPM photon tracer:
// which = 0: sizeof(PhotonMapProperties), 1: sizeof(Photon)
int ref_ppm_sizeof(int which) { return which == 0 ? (int)sizeof(PhotonMapProperties) : (int)sizeof(Photon); }
// default-constructed block, then (optionally) SetBounds(cloud bounds) -- out receives sizeof(PhotonMapProperties) bytes
void ref_photon_map_properties(const float cloud_bounds[8], int set_bounds, void* out, unsigned* total_size)
{
    PhotonMapProperties p;
    if (set_bounds) {
        glm::vec4 b[2] = {glm::vec4(cloud_bounds[0], cloud_bounds[1], cloud_bounds[2], cloud_bounds[3]),
                          glm::vec4(cloud_bounds[4], cloud_bounds[5], cloud_bounds[6], cloud_bounds[7])};
        p.SetBounds(b);
    }
    std::memcpy(out, &p, sizeof(p));
    if (total_size) *total_size = p.GetTotalSize();
}

}  // extern "C"
