// sv_sweep.cu -- shadow-volume (deep shadow map) build: the sm_100a replacement of ShadowVolume.comp.
//
// Reference: shaders/ShadowVolume.comp:77-94 (helpers :34-56), dispatched as (N/32+1)^2 groups of 32x32
// (RenderTechniqueSV.cpp:110).  One thread owns one (x,y) column of the light-aligned N^3 cube and marches z:
//     p      = (b0 + (x*right + y*up + z*light) * voxelSize).xyz                       (:50-56)
//     if p inside the cloud AABB (inclusive, :34-42):  acc += texture(cloud, p) * densityScaling / baseScaling
//     dist  += voxelSize                                    (repeated fp32 addition, :91)
//     out[x,y,z] = exp(-dist * acc / (z+1))                                             (:92)
//
// Parity: a transmittance of exp(-tau) with tau up to ~70 only stays within 1e-5 relative if tau itself is
// right to about one ulp, so everything up to the exponent is evaluated in exactly the oracle's operation
// order with IEEE + - * / (no contraction: file compiled with --fmad=false) and the serial z order of the
// running sums is kept.  Only expf differs from glibc (<= 2 ulp).
//
// Shape of the work (N = 1024, cloud 1024^3): 4*N^3 = 4.3 GB of output written exactly once, coalesced
// along x (the warp's 32 columns are consecutive x), + one compulsory read of the cloud (4.3 GB) through
// L1/L2; roofline = HBM.  Algorithmic bytes per voxel: 4 B store + 4*Ncloud^3/N^3 B load.
// The z loop is processed in chunks of U slices: all trilinear gathers of a chunk are issued before the
// serial accumulate/exp/store pass, which gives 8*U independent loads in flight per thread.
#include "crv_internal.h"
#include "sv_sweep.h"

namespace {

template <int CL, int OL, int U>
__global__ void __launch_bounds__(256) sv_sweep_kernel(const __grid_constant__ SVConst c)
{
    const int x = blockIdx.x * 32 + threadIdx.x;
    const int y = blockIdx.y * 8 + threadIdx.y;
    const int N = c.n;
    if (x >= N || y >= N) return;      // the reference's surplus threads only produce discarded OOB stores

    // (x*right + y*up), loop invariant per column
    const float fx = (float)x, fy = (float)y;
    const float ax = fx * c.right[0] + fy * c.up[0];
    const float ay = fx * c.right[1] + fy * c.up[1];
    const float az = fx * c.right[2] + fy * c.up[2];

    float acc = 0.0f;      // accumulatedValue.x
    float dist = 0.0f;     // accumulatedDistance

    for (int z0 = 0; z0 < N; z0 += U) {
        float tex[U];
        bool in[U];
        // ---- gather pass: positions, AABB tests and trilinear fetches of U slices ----
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int z = z0 + u;
            in[u] = false;
            tex[u] = 0.0f;
            if (z < N) {
                const float fz = (float)z;
                const float px = c.b0[0] + ((ax + fz * c.light[0]) * c.voxel_size);
                const float py = c.b0[1] + ((ay + fz * c.light[1]) * c.voxel_size);
                const float pz = c.b0[2] + ((az + fz * c.light[2]) * c.voxel_size);
                const bool inside = !(px < c.cb0[0] || py < c.cb0[1] || pz < c.cb0[2] ||
                                      px > c.cb1[0] || py > c.cb1[1] || pz > c.cb1[2]);
                if (inside) {
                    // sampleCloud (:44-48): ((p,0) - b0) / (b1 - b0)
                    const float sx = (px - c.cb0[0]) / c.cext[0];
                    const float sy = (py - c.cb0[1]) / c.cext[1];
                    const float sz = (pz - c.cb0[2]) / c.cext[2];
                    const Footprint f = make_footprint(c.cloud, sx, sy, sz);
                    tex[u] = trilinear_fetch<CL>(c.cloud, f);
                    in[u] = true;
                }
            }
        }
        // ---- serial pass: running sums, transmittance, store ----
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int z = z0 + u;
            if (z < N) {
                if (in[u]) acc = acc + (tex[u] * c.density_scaling) / c.base_scaling;
                dist = dist + c.voxel_size;
                float t;
                if (acc == 0.0f) {
                    t = 1.0f;      // exp(((-dist)*0)/(z+1)) = exp(-0) = 1 exactly: skip the divide + exp
                } else {
                    t = expf(((-dist) * acc) / (float)(z + 1));
                }
                VolAcc<OL>::store(c.out, x, y, z, t);
            }
        }
    }
}

template <int CL, int OL>
void launch_sv(const SVConst& c, cudaStream_t s)
{
    const dim3 block(32, 8, 1);
    const dim3 grid((c.n + 31) / 32, (c.n + 7) / 8, 1);
    sv_sweep_kernel<CL, OL, 4><<<grid, block, 0, s>>>(c);
}

template <int CL>
void launch_sv_out(const SVConst& c, int out_layout, cudaStream_t s)
{
    if (out_layout == CRV_LAYOUT_BRICK) launch_sv<CL, CRV_LAYOUT_BRICK>(c, s);
    else if (out_layout == CRV_LAYOUT_TEX3D) launch_sv<CL, CRV_LAYOUT_TEX3D>(c, s);
    else launch_sv<CL, CRV_LAYOUT_LINEAR>(c, s);
}

}  // namespace

int crv_launch_sv_sweep(crv_ctx* ctx, const SVConst& c, int cloud_layout, int out_layout, cudaStream_t s)
{
    if (cloud_layout == CRV_LAYOUT_BRICK) launch_sv_out<CRV_LAYOUT_BRICK>(c, out_layout, s);
    else if (cloud_layout == CRV_LAYOUT_TEX3D) launch_sv_out<CRV_LAYOUT_TEX3D>(c, out_layout, s);
    else launch_sv_out<CRV_LAYOUT_LINEAR>(c, out_layout, s);
    ctx->launches++;
    CRV_CUDA(cudaGetLastError());
    return CRV_OK;
}
