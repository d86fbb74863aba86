// sv_sweep.h -- kernel-parameter block of the shadow-volume build (see sv_sweep.cu)
#pragma once
#include "crv_internal.h"

struct SVConst {
    // ShadowVolumeProperties (ShadowVolume.comp:9-19): bounds[0].xyz, right, up, lightDirection, voxelSize, N
    float b0[3];
    float right[3], up[3], light[3];
    float voxel_size;
    int n;
    // CloudProperties (ShadowVolume.comp:22-29): bounds and (b1 - b0), scaling factors
    float cb0[3], cb1[3], cext[3];
    float density_scaling, base_scaling;
    VolView cloud;      // sampled (binding 2)
    VolView out;        // stored  (binding 0)
    // ---- fast path only (sv_sweep_fast_kernel) ----
    float rcext[3];             // RN(1 / cext[i])            (host, IEEE)
    float rbase;                // RN(1 / base_scaling)
    const float* brick_absmax;  // dilated per-brick max |v| of the cloud (volume.cu), to skip empty-space gathers
    float* dist_z;              // dist_z[z] = accumulatedDistance after slice z (n entries, device; written by pass 0)
    int search_iters;           // ceil(log2(n + 1)): trip count of the per-column z-interval bisection
    ushort2* zrange;            // two-pass build: [z_in, z_end) of every column, n*n entries (device)
    // fused build: approximate brick coordinate of voxel (x,y,z), affine in all three: v = pre_c + x*pre_r + y*pre_u + z*pre_l
    // (~ ((s*size - 0.5) / 8) per axis), against the wide-dilated brick table; null when the error bound does not hold
    const float* brick_absmax_wide;
    float pre_c[3], pre_r[3], pre_u[3], pre_l[3];
    const int* occ_ext;         // crv_volume::occ_ext of the cloud: brick-index extremes of its occupied bricks (device, 6 ints)
};

constexpr int SV_MAX_FAST_N = 4096;   // entries of the constant reciprocal table (4096^3 floats = 275 GB: beyond HBM anyway)

// fills the constant reciprocal table of the fast kernels, once per device
int crv_sv_init_tables(crv_ctx* ctx);
// true when the parameter block satisfies the fast kernel's preconditions (finite values, voxelSize > 0,
// divisors in the range where the FMA division of exact_div.cuh is exact)
bool crv_sv_fast_path_ok(const SVConst& c);

// crv_sv_mode of crv_c_api.h (AUTO is resolved by the caller)
enum { SV_MODE_GENERIC = 1, SV_MODE_SINGLE_PASS = 2, SV_MODE_TWO_PASS = 3, SV_MODE_FUSED = 4, SV_MODE_FUSED_V1 = 5, SV_MODE_FUSED_ROWS = 6 };
// dist_z[z] = accumulatedDistance after slice z (the shader's repeated fp32 addition), for the fused build
int crv_launch_sv_dist(crv_ctx* ctx, float* dist_z, int n, float voxel_size, cudaStream_t s);
int crv_launch_sv_sweep(crv_ctx* ctx, const SVConst& c, int cloud_layout, int out_layout, int mode, cudaStream_t s);
