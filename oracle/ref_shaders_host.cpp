// oracle/ref_shaders_host.cpp -- CPU dispatch of the reference's OWN compute shaders.  TEST INFRASTRUCTURE ONLY.
//
// The #include lines below pull in /root/reference/shaders/PathTracer.comp, ShadowVolume.comp, PPM_PT.comp and PPM_PE.comp after the
// lexical rewrite of oracle/glsl_to_cpp.py (generated into the git-ignored oracle/_ref/, never committed); the GLSL
// types and built-ins come from oracle/glsl_compat.h.  This file is only the "driver": it binds the descriptor set
// (RenderTechniquePT.cpp:20-35 bindings 0-6, RenderTechniqueSV.cpp:13-22 bindings 0-3), pushes the constants
// (RenderTechniquePT.cpp:132) and runs one shader_main() per invocation of the dispatch grid
// (RenderTechniquePT.cpp:141: (W/32+1) x (H/32+1) groups of 32x32; RenderTechniqueSV.cpp:110: (N/32+1)^2 groups).
//
// Built by oracle/Makefile into oracle/_ref/libcrv_refshaders.so only where /root/reference exists; the GPU box uses
// the prebuilt file.  It pins the restatement (crv_oracle.cpp): tests/test_ref_shaders.py demands bit-identical
// images, shadow volumes and per-pixel iteration counts between the two.
#include <omp.h>

#include <cstring>

#include "glsl_compat.h"

namespace glsl {
namespace pt {
#include "_ref/PathTracer.comp.inc"
}
namespace sv {
#include "_ref/ShadowVolume.comp.inc"
}
namespace ppm {
#include "_ref/PPM_PT.comp.inc"
}
namespace ppe {
#include "_ref/PPM_PE.comp.inc"
}
}  // namespace glsl

namespace {

template <class Block> void bind_block(Block& dst, const void* src, size_t nbytes, size_t declared)
{
    static_assert(sizeof(Block) >= 20, "uniform block unexpectedly small");
    std::memset(&dst, 0, sizeof(Block));
    std::memcpy(&dst, src, nbytes < declared ? nbytes : declared);
}

int resolve_threads(int n) { return n > 0 ? n : omp_get_max_threads(); }

}  // namespace

extern "C" {

// sizes of the uniform blocks as the SHADER declares them (std140 / push-constant layout reproduced by the C++ structs
// the rewrite emits): checked against the host structs of UniformBuffers.h by tests/test_ref_shaders.py
void refsh_block_sizes(uint32_t out[7])
{
    out[0] = sizeof(glsl::pt::CameraProperties_block);
    out[1] = sizeof(glsl::pt::CloudProperties_block);
    out[2] = sizeof(glsl::pt::Parameters_block);
    out[3] = sizeof(glsl::pt::ShadowVolumeProperties_block);
    out[4] = offsetof(glsl::pt::PushConstants_block, pmRadius) + sizeof(float);
    out[5] = sizeof(glsl::sv::ShadowVolumeProperties_block);
    out[6] = sizeof(glsl::sv::CloudProperties_block);
}

// One vkCmdDispatch of PathTracer.comp.  image = W*H rgba32f, updated in place by the shader's own accumulate.
// Only invocations with x in [x0,x1), y in [y0,y1), (y - y0) % row_step == 0 run (x1/y1 may extend into the
// off-image part of the dispatch grid, whose stores are discarded as on the GPU).
// steps / collisions (optional, W*H): calls of log() (one per delta-tracking iteration, :338) and of
// texture(shadowSampler) (one per real collision, :126) made by the invocation -- counted by the compat layer,
// the shader text is not instrumented.
void refsh_path_trace(const float* cloud, int nx, int ny, int nz, const float* shadow, int nsv, const void* camera68,
                      const void* cloud60, const void* params28, const void* svprops152, const void* push24, float* image,
                      int x0, int y0, int x1, int y1, int row_step, uint32_t* steps, uint32_t* collisions, int num_threads)
{
    using namespace glsl::pt;
    bind_block(cameraProperties, camera68, 68, sizeof(cameraProperties));
    bind_block(cloudProperties, cloud60, 60, sizeof(cloudProperties));
    bind_block(parameters, params28, 28, sizeof(parameters));
    bind_block(shadowVolumeProperties, svprops152, 152, sizeof(shadowVolumeProperties));
    bind_block(pushConstants, push24, 20, sizeof(pushConstants));
    const int W = cameraProperties.halfWidth * 2, H = cameraProperties.halfHeight * 2;      // CameraProperties::GetWidth/Height
    cloudSampler.texels = cloud; cloudSampler.nx = nx; cloudSampler.ny = ny; cloudSampler.nz = nz; cloudSampler.slot = 0;
    shadowSampler.texels = shadow; shadowSampler.nx = shadowSampler.ny = shadowSampler.nz = nsv; shadowSampler.slot = 1;
    resultImage.texels = image; resultImage.width = W; resultImage.height = H;
    const glsl::uvec3 groups((uint32_t)(W / 32 + 1), (uint32_t)(H / 32 + 1), 1u);
    const int gx_end = (int)(groups.x * 32u), gy_end = (int)(groups.y * 32u);
    if (x1 > gx_end) x1 = gx_end;
    if (y1 > gy_end) y1 = gy_end;
    if (row_step < 1) row_step = 1;
    const int nrows = y1 > y0 ? (y1 - y0 + row_step - 1) / row_step : 0;
#pragma omp parallel for schedule(dynamic, 1) num_threads(resolve_threads(num_threads))
    for (int r = 0; r < nrows; ++r) {
        const int y = y0 + r * row_step;
        for (int x = x0; x < x1; ++x) {
            glsl::invocation().global_id = glsl::uvec3((uint32_t)x, (uint32_t)y, 0u);
            glsl::invocation().num_groups = groups;
            glsl::Counters& c = glsl::counters();
            const uint64_t l0 = c.log_calls, t0 = c.texture_calls[1];
            shader_main();
            if (x < W && y < H) {
                if (steps) steps[(size_t)y * W + x] = (uint32_t)(c.log_calls - l0);
                if (collisions) collisions[(size_t)y * W + x] = (uint32_t)(c.texture_calls[1] - t0);
            }
        }
    }
}

// One vkCmdDispatch of ShadowVolume.comp; out = N*N*N r32f, x fastest.
void refsh_shadow_volume(const float* cloud, int nx, int ny, int nz, const void* cloud60, const void* svprops152, float* out,
                         int num_threads)
{
    using namespace glsl::sv;
    bind_block(cloudProperties, cloud60, 60, sizeof(cloudProperties));
    bind_block(shadowVolumeProperties, svprops152, 152, sizeof(shadowVolumeProperties));
    const int N = (int)shadowVolumeProperties.voxelAxisCount;
    cloudSampler.texels = cloud; cloudSampler.nx = nx; cloudSampler.ny = ny; cloudSampler.nz = nz; cloudSampler.slot = 0;
    shadowVolume.texels = out; shadowVolume.nx = shadowVolume.ny = shadowVolume.nz = N;
    const glsl::uvec3 groups((uint32_t)(N / 32 + 1), (uint32_t)(N / 32 + 1), 1u);
    const int g_end = (int)(groups.x * 32u);
#pragma omp parallel for schedule(dynamic, 1) num_threads(resolve_threads(num_threads))
    for (int y = 0; y < g_end; ++y) {
        for (int x = 0; x < g_end; ++x) {
            if (x >= N || y >= N) continue;      // every store of an off-volume column is discarded: skip its work
            glsl::invocation().global_id = glsl::uvec3((uint32_t)x, (uint32_t)y, 0u);
            glsl::invocation().num_groups = groups;
            shader_main();
        }
    }
}

int refsh_max_threads(void) { return omp_get_max_threads(); }


// One photon-tracing dispatch of PPM_PT.comp (RenderTechniquePPM.cpp:233-253: both buffers zero-filled, then
// vkCmdDispatch(groups_x, 1, 1) of 64x1x1 invocations; the reference uses groups_x = 200).  photons = cells*32 records of
// 48 bytes, collisions = cells uvec4; bindings 0-5 as RenderTechniquePPM.cpp:22-38 declares them.
uint32_t refsh_ppm_block_size(void) { return (uint32_t)sizeof(glsl::ppm::PhotonMapProperties_block); }
uint32_t refsh_ppm_photon_size(void) { return (uint32_t)sizeof(glsl::ppm::Photon); }

void refsh_ppm_trace(const float* cloud, int nx, int ny, int nz, const void* cloud60, const void* params28, const void* pmprops76,
                     const void* push24, uint32_t groups_x, void* photons, uint32_t* collisions, uint64_t cells, int num_threads)
{
    namespace P = glsl::ppm;
    bind_block(P::cloudProperties, cloud60, 60, sizeof(P::CloudProperties_block));
    bind_block(P::parameters, params28, 28, sizeof(P::Parameters_block));
    bind_block(P::photonMapProperties, pmprops76, 76, sizeof(P::PhotonMapProperties_block));
    std::memset(&P::pushConstants, 0, sizeof(P::pushConstants));
    std::memcpy(&P::pushConstants, push24, 24 < sizeof(P::pushConstants) ? 24 : sizeof(P::pushConstants));
    P::cloudSampler.texels = cloud;
    P::cloudSampler.nx = nx; P::cloudSampler.ny = ny; P::cloudSampler.nz = nz;
    P::cloudSampler.slot = 0;
    P::photons = static_cast<P::Photon*>(photons);
    P::collisions = reinterpret_cast<glsl::uvec4*>(collisions);
    std::memset(photons, 0, cells * 32 * sizeof(P::Photon));            // vkCmdFillBuffer(..., 0)
    std::memset(collisions, 0, cells * sizeof(glsl::uvec4));
    const int total = (int)(groups_x * 64u);
#pragma omp parallel for schedule(dynamic, 64) num_threads(resolve_threads(num_threads))
    for (int i = 0; i < total; ++i) {
        glsl::invocation().global_id = glsl::uvec3((uint32_t)i, 0u, 0u);
        glsl::invocation().num_groups = glsl::uvec3(groups_x, 1u, 1u);
        P::shader_main();
    }
}

// One photon-estimate dispatch of PPM_PE.comp (RenderTechniquePPM.cpp:260-273: (W/32+1) x (H/32+1) groups of 32x32) over a
// given photon map; bindings 0-9 as RenderTechniquePPM.cpp:59-81 declares them.  image = W*H rgba32f, accumulated in place by
// the shader.  Only rows y0 <= y < y1 with (y - y0) % row_step == 0 run.
void refsh_ppm_estimate(const float* cloud, int nx, int ny, int nz, const float* sv, int nsv, const void* camera68, const void* cloud60,
                        const void* params28, const void* pmprops76, const void* svprops152, const void* push24, const void* photons,
                        const uint32_t* collisions, float* image, int y0, int y1, int row_step, int num_threads)
{
    namespace P = glsl::ppe;
    bind_block(P::cameraProperties, camera68, 68, sizeof(P::CameraProperties_block));
    bind_block(P::cloudProperties, cloud60, 60, sizeof(P::CloudProperties_block));
    bind_block(P::parameters, params28, 28, sizeof(P::Parameters_block));
    bind_block(P::photonMapProperties, pmprops76, 76, sizeof(P::PhotonMapProperties_block));
    bind_block(P::shadowVolumeProperties, svprops152, 152, sizeof(P::ShadowVolumeProperties_block));
    std::memset(&P::pushConstants, 0, sizeof(P::pushConstants));
    std::memcpy(&P::pushConstants, push24, 24 < sizeof(P::pushConstants) ? 24 : sizeof(P::pushConstants));
    P::cloudSampler.texels = cloud; P::cloudSampler.nx = nx; P::cloudSampler.ny = ny; P::cloudSampler.nz = nz; P::cloudSampler.slot = 0;
    P::shadowVolumeSampler.texels = sv; P::shadowVolumeSampler.nx = nsv; P::shadowVolumeSampler.ny = nsv; P::shadowVolumeSampler.nz = nsv;
    P::shadowVolumeSampler.slot = 1;
    const int W = P::cameraProperties.halfWidth * 2, H = P::cameraProperties.halfHeight * 2;
    P::resultImage.texels = image; P::resultImage.width = W; P::resultImage.height = H;
    P::photons = static_cast<P::Photon*>(const_cast<void*>(photons));
    P::collisions = reinterpret_cast<glsl::uvec4*>(const_cast<uint32_t*>(collisions));
    const uint32_t gx = (uint32_t)(W / 32 + 1), gy = (uint32_t)(H / 32 + 1);
    if (y1 > (int)(gy * 32u)) y1 = (int)(gy * 32u);
    if (row_step < 1) row_step = 1;
    const int nrows = (y1 - y0 + row_step - 1) / row_step;
#pragma omp parallel for schedule(dynamic, 1) num_threads(resolve_threads(num_threads))
    for (int r = 0; r < nrows; ++r) {
        const int y = y0 + r * row_step;
        for (uint32_t x = 0; x < gx * 32u; ++x) {
            glsl::invocation().global_id = glsl::uvec3(x, (uint32_t)y, 0u);
            glsl::invocation().num_groups = glsl::uvec3(gx, gy, 1u);
            P::shader_main();
        }
    }
}

}  // extern "C"
