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

// The element offset of voxel (i,j,k) is separable in both global layouts: off = off_x(i) + off_y(j) + off_z(k)
// (< 2^32 elements, guaranteed by crv_volume_create), which lets the 8 texel addresses of a trilinear
// footprint be formed from 6 per-axis terms and 12 adds instead of 8 full index computations.
template <> struct VolAcc<CRV_LAYOUT_LINEAR> {
    static __device__ __forceinline__ unsigned off_x(const VolView&, int i) { return (unsigned)i; }
    static __device__ __forceinline__ unsigned off_y(const VolView& v, int j) { return (unsigned)j * (unsigned)v.nx; }
    static __device__ __forceinline__ unsigned off_z(const VolView& v, int k) { return (unsigned)k * ((unsigned)v.nx * (unsigned)v.ny); }
    // elements between slices k and k+1 while both lie in the same aligned group of 8 slices
    static __device__ __forceinline__ unsigned z_step(const VolView& v) { return (unsigned)v.nx * (unsigned)v.ny; }
    static __device__ __forceinline__ size_t index(const VolView& v, int i, int j, int k)
    {
        return static_cast<size_t>(off_x(v, i) + off_y(v, j) + off_z(v, k));
    }
    static __device__ __forceinline__ float load(const VolView& v, int i, int j, int k) { return __ldg(v.data + index(v, i, j, k)); }
    static __device__ __forceinline__ void store(const VolView& v, int i, int j, int k, float val)
    {
        const_cast<float*>(v.data)[index(v, i, j, k)] = val;
    }
    // write-once data that another kernel reads later: do not let it displace the sampled volume in L2
    static __device__ __forceinline__ void store_streaming(const VolView& v, int i, int j, int k, float val)
    {
        __stcs(const_cast<float*>(v.data) + index(v, i, j, k), val);
    }
    // read of a volume this kernel also writes (coherent path, streaming)
    static __device__ __forceinline__ float load_rw(const VolView& v, int i, int j, int k) { return __ldcs(v.data + index(v, i, j, k)); }
};

template <> struct VolAcc<CRV_LAYOUT_BRICK> {
    static __device__ __forceinline__ unsigned off_x(const VolView&, int i) { return ((unsigned)(i >> CRV_BRICK_SHIFT) << 9) | (unsigned)(i & 7); }
    static __device__ __forceinline__ unsigned off_y(const VolView& v, int j)
    {
        return (unsigned)(j >> CRV_BRICK_SHIFT) * ((unsigned)v.nbx << 9) + ((unsigned)(j & 7) << 3);
    }
    static __device__ __forceinline__ unsigned off_z(const VolView& v, int k)
    {
        return (unsigned)(k >> CRV_BRICK_SHIFT) * (((unsigned)v.nbx * (unsigned)v.nby) << 9) + ((unsigned)(k & 7) << 6);
    }
    static __device__ __forceinline__ unsigned z_step(const VolView&) { return 64u; }
    static __device__ __forceinline__ size_t index(const VolView& v, int i, int j, int k)
    {
        return static_cast<size_t>(off_x(v, i) + off_y(v, j) + off_z(v, k));
    }
    static __device__ __forceinline__ float load(const VolView& v, int i, int j, int k) { return __ldg(v.data + index(v, i, j, k)); }
    static __device__ __forceinline__ void store(const VolView& v, int i, int j, int k, float val)
    {
        const_cast<float*>(v.data)[index(v, i, j, k)] = val;
    }
    static __device__ __forceinline__ void store_streaming(const VolView& v, int i, int j, int k, float val)
    {
        __stcs(const_cast<float*>(v.data) + index(v, i, j, k), val);
    }
    static __device__ __forceinline__ float load_rw(const VolView& v, int i, int j, int k) { return __ldcs(v.data + index(v, i, j, k)); }
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
    static __device__ __forceinline__ void store_streaming(const VolView& v, int i, int j, int k, float val) { store(v, i, j, k, val); }
    static __device__ __forceinline__ float load_rw(const VolView& v, int i, int j, int k)
    {
        return surf3Dread<float>(v.surf, i * static_cast<int>(sizeof(float)), j, k);
    }
};

// Integer texel footprint + weights of texture(sampler3D, s) with LINEAR / CLAMP_TO_EDGE
// (Initializers.cpp:408-418; Vulkan texel filtering: u = s*size, i0 = floor(u-0.5), alpha = frac(u-0.5),
// clamp the integer index).  Same operation order as the oracle, no contraction.
struct Footprint {
    int i0, i1, j0, j1, k0, k1;   // clamped texel indices
    float a, b, c;                // weights along x, y, z
};

// INSIDE = true: the caller guarantees 0 <= s <= 1 (so -1 <= floor(u) <= n-1) and the clamps shrink to one
// max and one min; otherwise any float (NaN, +-inf, far outside) is handled.
template <bool INSIDE = false>
__device__ __forceinline__ void footprint_axis(float s, float fn, int n, int& i0c, int& i1c, float& a)
{
    const float u = __fsub_rn(__fmul_rn(s, fn), 0.5f);
    const int i0 = __float2int_rd(u);                 // floor; saturating, NaN -> 0
    a = __fsub_rn(u, INSIDE ? (float)i0 : floorf(u));
    if (INSIDE) {
        i0c = max(i0, 0);
        i1c = min(i0 + 1, n - 1);
    } else {
        i0c = min(max(i0, 0), n - 1);
        i1c = max(min(i0, n - 2) + 1, 0);             // clamp(i0 + 1, 0, n-1) without overflow
    }
}

template <bool INSIDE = false>
__device__ __forceinline__ Footprint make_footprint(const VolView& v, float sx, float sy, float sz)
{
    Footprint f;
    footprint_axis<INSIDE>(sx, v.fnx, v.nx, f.i0, f.i1, f.a);
    footprint_axis<INSIDE>(sy, v.fny, v.ny, f.j0, f.j1, f.b);
    footprint_axis<INSIDE>(sz, v.fnz, v.nz, f.k0, f.k1, f.c);
    return f;
}

__device__ __forceinline__ float lerp_rn(float v0, float v1, float a)
{
    return __fadd_rn(v0, __fmul_rn(a, __fsub_rn(v1, v0)));      // v0 + a*(v1 - v0), never contracted
}

__device__ __forceinline__ float trilerp(const Footprint& f, float t000, float t100, float t010, float t110, float t001,
                                         float t101, float t011, float t111)
{
    const float x00 = lerp_rn(t000, t100, f.a);
    const float x10 = lerp_rn(t010, t110, f.a);
    const float x01 = lerp_rn(t001, t101, f.a);
    const float x11 = lerp_rn(t011, t111, f.a);
    const float y0 = lerp_rn(x00, x10, f.b);
    const float y1 = lerp_rn(x01, x11, f.b);
    return lerp_rn(y0, y1, f.c);
}

template <int LAYOUT>
__device__ __forceinline__ float trilinear_fetch(const VolView& v, const Footprint& f)
{
    if constexpr (LAYOUT == CRV_LAYOUT_TEX3D) {
        const float t000 = VolAcc<LAYOUT>::load(v, f.i0, f.j0, f.k0);
        const float t100 = VolAcc<LAYOUT>::load(v, f.i1, f.j0, f.k0);
        const float t010 = VolAcc<LAYOUT>::load(v, f.i0, f.j1, f.k0);
        const float t110 = VolAcc<LAYOUT>::load(v, f.i1, f.j1, f.k0);
        const float t001 = VolAcc<LAYOUT>::load(v, f.i0, f.j0, f.k1);
        const float t101 = VolAcc<LAYOUT>::load(v, f.i1, f.j0, f.k1);
        const float t011 = VolAcc<LAYOUT>::load(v, f.i0, f.j1, f.k1);
        const float t111 = VolAcc<LAYOUT>::load(v, f.i1, f.j1, f.k1);
        return trilerp(f, t000, t100, t010, t110, t001, t101, t011, t111);
    } else {
        using A = VolAcc<LAYOUT>;
        const unsigned x0 = A::off_x(v, f.i0), x1 = A::off_x(v, f.i1);
        const unsigned y0 = A::off_y(v, f.j0), y1 = A::off_y(v, f.j1);
        const unsigned z0 = A::off_z(v, f.k0), z1 = A::off_z(v, f.k1);
        const unsigned r00 = y0 + z0, r10 = y1 + z0, r01 = y0 + z1, r11 = y1 + z1;
        const float t000 = __ldg(v.data + (r00 + x0));
        const float t100 = __ldg(v.data + (r00 + x1));
        const float t010 = __ldg(v.data + (r10 + x0));
        const float t110 = __ldg(v.data + (r10 + x1));
        const float t001 = __ldg(v.data + (r01 + x0));
        const float t101 = __ldg(v.data + (r01 + x1));
        const float t011 = __ldg(v.data + (r11 + x0));
        const float t111 = __ldg(v.data + (r11 + x1));
        return trilerp(f, t000, t100, t010, t110, t001, t101, t011, t111);
    }
}

// ------------------------------------------------------------------------------------------------
// host objects behind the opaque handles
// ------------------------------------------------------------------------------------------------
struct crv_ctx {
    int device = 0;
    int sm_count = 0;
    int cc_major = 0, cc_minor = 0;
    uint64_t launches = 0;          // kernels launched through this context
    bool sv_tables_ready = false;   // constant tables of sv_sweep.cu uploaded to this device
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
    float* brick_absmax_wide = nullptr;    // nbx*nby*nbz: max |v| over voxels [8b-1, 8b+9]^3 (approximate brick test)
    int* occ_ext = nullptr;                // 6 ints, device: min / max brick index per axis over the bricks with brick_absmax_wide != 0
                                           // ("occupied"); min = INT_MAX, max = -1 when there is none.  Valid with the brick tables.
    bool brick_max_valid = false;
    float majorant = 0.0f;
    bool majorant_valid = false;
    uint64_t generation = 0;               // contents version: every upload / build / mark_dirty draws a new process-wide number
    VolView view() const;
};

struct crv_image {
    crv_ctx* ctx = nullptr;
    uint32_t width = 0, height = 0;
    float* rgba = nullptr;                 // width*height*4 floats
    bool owned = true;
    unsigned char* bgra8 = nullptr;        // width*height*4 bytes: staging of the present step (allocated on first use)
};

// kernels / helpers implemented across the .cu files --------------------------------------------------
int crv_launch_retile(crv_ctx* ctx, const float* linear_src, crv_volume* dst, cudaStream_t s);     // volume.cu
int crv_launch_retile_tables(crv_ctx* ctx, const float* linear_src, crv_volume* dst, bool store, float* scratch, cudaStream_t s);  // volume.cu
int crv_launch_untile(crv_ctx* ctx, const crv_volume* src, float* linear_dst, cudaStream_t s);     // volume.cu
int crv_launch_brick_max(crv_ctx* ctx, crv_volume* vol, cudaStream_t s);                           // volume.cu
int crv_launch_majorant(crv_ctx* ctx, crv_volume* vol, cudaStream_t s);                            // volume.cu
int crv_launch_blit_srgb8(crv_ctx* ctx, const float* rgba, unsigned char* bgra8, size_t npix, cudaStream_t s);  // volume.cu
int crv_launch_fbm(crv_ctx* ctx, float* linear_dst, uint32_t n, uint32_t seed, cudaStream_t s);    // fbm.cu
