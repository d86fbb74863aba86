// crv_internal.h -- shared between the translation units of libcrv_b200.so (host objects, device-side
// volume views and the trilinear sampler).  Not part of the public boundary (include/crv_c_api.h is).
#pragma once

#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <string>

#include "../../include/crv_c_api.h"

// ------------------------------------------------------------------------------------------------
// error plumbing: no exceptions/abort across the C boundary (crv_c_api.h "Conventions")
// ------------------------------------------------------------------------------------------------
void crv_set_error(const char* fmt, ...);

#define CRV_CUDA(expr)                                                                              \
    do {                                                                                            \
        cudaError_t _e = (expr);                                                                    \
        if (_e != cudaSuccess) {                                                                    \
            crv_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return (_e == cudaErrorMemoryAllocation) ? CRV_ERR_OOM : CRV_ERR_CUDA;                  \
        }                                                                                           \
    } while (0)

#define CRV_REQUIRE(cond, msg)                       \
    do {                                             \
        if (!(cond)) {                               \
            crv_set_error("%s", msg);                \
            return CRV_ERR_INVALID;                  \
        }                                            \
    } while (0)

// ------------------------------------------------------------------------------------------------
// Device-side view of a 3D R32F volume (cloud density or shadow volume).
//   LINEAR : data[x + nx*(y + ny*z)]                                   (Grid3D order, Grid3D.h:112-116)
//   BRICK  : 8x8x8 bricks of 2 KiB; brick (bx,by,bz) at ((bz*nby + by)*nbx + bx)*512,
//            voxel inside at (lz*64 + ly*8 + lx).  Dimensions are padded up to a multiple of 8.
//   TEX3D  : cudaArray behind a point-sampled texture object (+ surface object for writes)
// ------------------------------------------------------------------------------------------------
constexpr int CRV_BRICK = 8;          // voxels per brick edge (also the granularity of the max grid)
constexpr int CRV_BRICK_SHIFT = 3;
constexpr int CRV_BRICK_VOX = 512;

struct VolView {
    const float* data;            // LINEAR / BRICK
    cudaTextureObject_t tex;      // TEX3D
    cudaSurfaceObject_t surf;     // TEX3D (writes)
    int nx, ny, nz;               // logical size
    int nbx, nby, nbz;            // bricks per axis = ceil(n/8)
    float fnx, fny, fnz;          // float(n): the `size` of the Vulkan unnormalised coordinate u = s*size
};

template <int LAYOUT> struct VolAcc;

template <> struct VolAcc<CRV_LAYOUT_LINEAR> {
    static __device__ __forceinline__ size_t index(const VolView& v, int i, int j, int k)
    {
        return static_cast<size_t>(static_cast<unsigned>(i) + static_cast<unsigned>(v.nx) * (static_cast<unsigned>(j) + static_cast<unsigned>(v.ny) * static_cast<unsigned>(k)));
    }
    static __device__ __forceinline__ float load(const VolView& v, int i, int j, int k) { return __ldg(v.data + index(v, i, j, k)); }
    static __device__ __forceinline__ void store(const VolView& v, int i, int j, int k, float val)
    {
        const_cast<float*>(v.data)[index(v, i, j, k)] = val;
    }
};

template <> struct VolAcc<CRV_LAYOUT_BRICK> {
    static __device__ __forceinline__ size_t index(const VolView& v, int i, int j, int k)
    {
        const unsigned b = (static_cast<unsigned>(k >> CRV_BRICK_SHIFT) * v.nby + static_cast<unsigned>(j >> CRV_BRICK_SHIFT)) * v.nbx + static_cast<unsigned>(i >> CRV_BRICK_SHIFT);
        const unsigned l = ((k & 7) << 6) | ((j & 7) << 3) | (i & 7);
        return static_cast<size_t>(b) * CRV_BRICK_VOX + l;
    }
    static __device__ __forceinline__ float load(const VolView& v, int i, int j, int k) { return __ldg(v.data + index(v, i, j, k)); }
    static __device__ __forceinline__ void store(const VolView& v, int i, int j, int k, float val)
    {
        const_cast<float*>(v.data)[index(v, i, j, k)] = val;
    }
};

template <> struct VolAcc<CRV_LAYOUT_TEX3D> {
    // point filter + unnormalised coordinates: texel i is fetched at i + 0.5.  NO hardware filtering is
    // used anywhere: CUDA's linear filter has 9-bit weights, far outside the 1e-4/1e-5 parity budget.
    static __device__ __forceinline__ float load(const VolView& v, int i, int j, int k)
    {
        return tex3D<float>(v.tex, static_cast<float>(i) + 0.5f, static_cast<float>(j) + 0.5f, static_cast<float>(k) + 0.5f);
    }
    static __device__ __forceinline__ void store(const VolView& v, int i, int j, int k, float val)
    {
        surf3Dwrite(val, v.surf, i * static_cast<int>(sizeof(float)), j, k);
    }
};

// Integer texel footprint + weights of texture(sampler3D, s) with LINEAR / CLAMP_TO_EDGE
// (Initializers.cpp:408-418; Vulkan texel filtering: u = s*size, i0 = floor(u-0.5), alpha = frac(u-0.5),
// clamp the integer index).  Same operation order as the oracle, no contraction.
struct Footprint {
    int i0, i1, j0, j1, k0, k1;   // clamped texel indices
    float a, b, c;                // weights along x, y, z
};

__device__ __forceinline__ void footprint_axis(float s, float fn, int n, int& i0c, int& i1c, float& a)
{
    const float u = __fsub_rn(__fmul_rn(s, fn), 0.5f);
    const float fu = floorf(u);
    a = __fsub_rn(u, fu);
    const int i0 = __float2int_rd(u);                 // saturating, NaN -> 0
    i0c = min(max(i0, 0), n - 1);
    i1c = max(min(i0, n - 2) + 1, 0);                 // clamp(i0 + 1, 0, n-1) without overflow
}

__device__ __forceinline__ Footprint make_footprint(const VolView& v, float sx, float sy, float sz)
{
    Footprint f;
    footprint_axis(sx, v.fnx, v.nx, f.i0, f.i1, f.a);
    footprint_axis(sy, v.fny, v.ny, f.j0, f.j1, f.b);
    footprint_axis(sz, v.fnz, v.nz, f.k0, f.k1, f.c);
    return f;
}

__device__ __forceinline__ float lerp_rn(float v0, float v1, float a)
{
    return __fadd_rn(v0, __fmul_rn(a, __fsub_rn(v1, v0)));      // v0 + a*(v1 - v0), never contracted
}

template <int LAYOUT>
__device__ __forceinline__ float trilinear_fetch(const VolView& v, const Footprint& f)
{
    const float t000 = VolAcc<LAYOUT>::load(v, f.i0, f.j0, f.k0);
    const float t100 = VolAcc<LAYOUT>::load(v, f.i1, f.j0, f.k0);
    const float t010 = VolAcc<LAYOUT>::load(v, f.i0, f.j1, f.k0);
    const float t110 = VolAcc<LAYOUT>::load(v, f.i1, f.j1, f.k0);
    const float t001 = VolAcc<LAYOUT>::load(v, f.i0, f.j0, f.k1);
    const float t101 = VolAcc<LAYOUT>::load(v, f.i1, f.j0, f.k1);
    const float t011 = VolAcc<LAYOUT>::load(v, f.i0, f.j1, f.k1);
    const float t111 = VolAcc<LAYOUT>::load(v, f.i1, f.j1, f.k1);
    const float x00 = lerp_rn(t000, t100, f.a);
    const float x10 = lerp_rn(t010, t110, f.a);
    const float x01 = lerp_rn(t001, t101, f.a);
    const float x11 = lerp_rn(t011, t111, f.a);
    const float y0 = lerp_rn(x00, x10, f.b);
    const float y1 = lerp_rn(x01, x11, f.b);
    return lerp_rn(y0, y1, f.c);
}

// ------------------------------------------------------------------------------------------------
// host objects behind the opaque handles
// ------------------------------------------------------------------------------------------------
struct crv_ctx {
    int device = 0;
    int sm_count = 0;
    int cc_major = 0, cc_minor = 0;
    uint64_t launches = 0;          // kernels launched through this context
};

struct crv_volume {
    crv_ctx* ctx = nullptr;
    int layout = CRV_LAYOUT_LINEAR;
    uint32_t nx = 0, ny = 0, nz = 0;
    uint32_t nbx = 0, nby = 0, nbz = 0;
    float* data = nullptr;                 // LINEAR / BRICK storage
    size_t data_elems = 0;
    cudaArray_t array = nullptr;           // TEX3D storage
    cudaTextureObject_t tex = 0;
    cudaSurfaceObject_t surf = 0;
    float* brick_max = nullptr;            // nbx*nby*nbz: max(v, 0) over each brick's dilated voxel set
    float* brick_absmax = nullptr;         // nbx*nby*nbz: max |v| over the same set (feeds the q-table)
    bool brick_max_valid = false;
    float majorant = 0.0f;
    bool majorant_valid = false;
    VolView view() const;
};

struct crv_image {
    crv_ctx* ctx = nullptr;
    uint32_t width = 0, height = 0;
    float* rgba = nullptr;                 // width*height*4 floats
    bool owned = true;
};

// kernels / helpers implemented across the .cu files --------------------------------------------------
int crv_launch_retile(crv_ctx* ctx, const float* linear_src, crv_volume* dst, cudaStream_t s);     // volume.cu
int crv_launch_untile(crv_ctx* ctx, const crv_volume* src, float* linear_dst, cudaStream_t s);     // volume.cu
int crv_launch_brick_max(crv_ctx* ctx, crv_volume* vol, cudaStream_t s);                           // volume.cu
int crv_launch_majorant(crv_ctx* ctx, crv_volume* vol, cudaStream_t s);                            // volume.cu
int crv_launch_fbm(crv_ctx* ctx, float* linear_dst, uint32_t n, uint32_t seed, cudaStream_t s);    // fbm.cu
