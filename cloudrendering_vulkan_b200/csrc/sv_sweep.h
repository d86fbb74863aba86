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
};

int crv_launch_sv_sweep(crv_ctx* ctx, const SVConst& c, int cloud_layout, int out_layout, cudaStream_t s);
