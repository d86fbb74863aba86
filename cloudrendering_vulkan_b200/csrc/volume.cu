// volume.cu -- layout conversion, majorant and brick-max kernels for 3D R32F volumes.
//
// Replaces the staging-buffer -> image copy of the reference (main.cpp:288-295, Utilities.cpp:94-111) and
// Grid3D::GetMajorant (Grid3D.h:181-192, an O(n) CPU scan) with device passes.  All of these are pure
// streaming kernels (HBM-bound, 4 B/voxel in, <= 4 B/voxel out); they run once per cloud upload.
#include "crv_internal.h"

VolView crv_volume::view() const
{
    VolView v;
    v.data = data;
    v.tex = tex;
    v.surf = surf;
    v.nx = (int)nx; v.ny = (int)ny; v.nz = (int)nz;
    v.nbx = (int)nbx; v.nby = (int)nby; v.nbz = (int)nbz;
    v.fnx = (float)nx; v.fny = (float)ny; v.fnz = (float)nz;
    return v;
}

namespace {

// linear x-fastest -> native layout of `dst` (BRICK or TEX3D surface).  One thread per voxel, x fastest, so
// reads are coalesced; BRICK writes are 32 B runs (8 voxels), i.e. whole sectors.
template <int LAYOUT>
__global__ void retile_kernel(const float* __restrict__ src, VolView dst)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y * blockDim.y + threadIdx.y;
    const int k = blockIdx.z;
    if (i >= dst.nx || j >= dst.ny) return;
    const size_t lin = (size_t)i + (size_t)dst.nx * ((size_t)j + (size_t)dst.ny * k);
    VolAcc<LAYOUT>::store(dst, i, j, k, src[lin]);
}

template <int LAYOUT>
__global__ void untile_kernel(VolView src, float* __restrict__ dstp)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y * blockDim.y + threadIdx.y;
    const int k = blockIdx.z;
    if (i >= src.nx || j >= src.ny) return;
    const size_t lin = (size_t)i + (size_t)src.nx * ((size_t)j + (size_t)src.ny * k);
    dstp[lin] = VolAcc<LAYOUT>::load(src, i, j, k);
}

// Brick-max grid: for brick (bx,by,bz) the max over voxels [8b, 8b+8] per axis (clamped to the volume):
// the brick's own 8 voxels plus ONE voxel of dilation on the upper side.  A trilinear footprint whose
// clamped lower corner (i0c,j0c,k0c) lies in the brick touches only voxels of that set, so every texel
// of the footprint is <= this max (used by the path tracer to reject tentative collisions without
// fetching, see pt_wavefront.cu).  One warp per brick: 9x9x9 = 729 voxels, 23 per lane.
// Two outputs per brick: max |v| (bounds the lerp and its rounding overshoot, feeds the q-table) and
// max(v, 0) (feeds the majorant, same start value T() = 0 as Grid3D::GetMajorant).
template <int LAYOUT>
__global__ void brick_max_kernel(VolView v, float* __restrict__ out_abs, float* __restrict__ out_max)
{
    const int warps_per_block = blockDim.x >> 5;
    const long long brick = (long long)blockIdx.x * warps_per_block + (threadIdx.x >> 5);
    const long long nbricks = (long long)v.nbx * v.nby * v.nbz;
    if (brick >= nbricks) return;
    const int lane = threadIdx.x & 31;
    const int bx = (int)(brick % v.nbx);
    const int by = (int)((brick / v.nbx) % v.nby);
    const int bz = (int)(brick / ((long long)v.nbx * v.nby));
    float m = 0.0f, ma = 0.0f;
    for (int t = lane; t < 729; t += 32) {
        const int lx = t % 9, ly = (t / 9) % 9, lz = t / 81;
        const int i = min(bx * 8 + lx, v.nx - 1), j = min(by * 8 + ly, v.ny - 1), k = min(bz * 8 + lz, v.nz - 1);
        const float val = VolAcc<LAYOUT>::load(v, i, j, k);
        m = fmaxf(m, val);
        ma = fmaxf(ma, fabsf(val));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        ma = fmaxf(ma, __shfl_xor_sync(0xffffffffu, ma, o));
    }
    if (lane == 0) { out_abs[brick] = ma; out_max[brick] = m; }
}

// Wider dilation for the path tracer's APPROXIMATE brick test (pt_wavefront.cu): max |v| over voxels
// [8b-1, 8b+9] per axis (clamped).  If an approximate texel coordinate (off by < 1 texel) has its lower footprint
// corner in brick b, every texel of the exact footprint lies in this set.  11^3 = 1331 voxels, 42 per lane.
template <int LAYOUT>
__global__ void brick_absmax_wide_kernel(VolView v, float* __restrict__ out_abs)
{
    const int warps_per_block = blockDim.x >> 5;
    const long long brick = (long long)blockIdx.x * warps_per_block + (threadIdx.x >> 5);
    const long long nbricks = (long long)v.nbx * v.nby * v.nbz;
    if (brick >= nbricks) return;
    const int lane = threadIdx.x & 31;
    const int bx = (int)(brick % v.nbx);
    const int by = (int)((brick / v.nbx) % v.nby);
    const int bz = (int)(brick / ((long long)v.nbx * v.nby));
    float ma = 0.0f;
    for (int t = lane; t < 1331; t += 32) {
        const int lx = t % 11 - 1, ly = (t / 11) % 11 - 1, lz = t / 121 - 1;
        const int i = min(max(bx * 8 + lx, 0), v.nx - 1), j = min(max(by * 8 + ly, 0), v.ny - 1), k = min(max(bz * 8 + lz, 0), v.nz - 1);
        ma = fmaxf(ma, fabsf(VolAcc<LAYOUT>::load(v, i, j, k)));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ma = fmaxf(ma, __shfl_xor_sync(0xffffffffu, ma, o));
    if (lane == 0) out_abs[brick] = ma;
}

// ------------------------------------------------------------------------------------------------------
// The same three tables from the LINEAR x-fastest source at upload time, as a separable max filter fused with the layout
// conversion: the dilated voxel sets are products of intervals ([8b, 8b+8] and [8b-1, 8b+9] per axis, clamped), so
//   pass X (retile_tables_kernel, the pass that reads every voxel anyway): per (brick column bx, y, z) the three window
//           maxima along x -> three arrays of nbx * ny * nz
//   pass YZ (tables_yz_kernel): per brick the maxima of those over its y and z windows.
// max is exact and order-independent: the tables hold the same floats as brick_max_kernel / brick_absmax_wide_kernel
// (1.4 ms of strided gathers at 512^3) produce.
// ------------------------------------------------------------------------------------------------------
template <int LAYOUT, bool STORE>
__global__ void __launch_bounds__(256) retile_tables_kernel(const float* __restrict__ src, VolView dst, float* __restrict__ ax_abs,
                                                            float* __restrict__ ax_max, float* __restrict__ ax_wide)
{
    const int i = blockIdx.x * 64 + threadIdx.x;            // blockDim = (64, 4): two warps per row, 8-lane groups = brick rows
    const int j = blockIdx.y * 4 + threadIdx.y;
    const int k = blockIdx.z;
    const bool row = j < dst.ny;
    const size_t base = (size_t)dst.nx * ((size_t)(row ? j : 0) + (size_t)dst.ny * k);
    const bool in = row && i < dst.nx;
    const float val = in ? src[base + i] : 0.0f;
    if (STORE && in) VolAcc<LAYOUT>::store(dst, i, j, k, val);
    // out-of-range lanes contribute the identity of both statistics (the clamped sets only repeat edge voxels)
    float a = fabsf(val), m = fmaxf(val, 0.0f);
    const unsigned lane = threadIdx.x & 31u;
    const float a_self = a, m_self = m;
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) {
        a = fmaxf(a, __shfl_xor_sync(0xffffffffu, a, o));
        m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    }
    // neighbours of the 8-voxel group: x = g-1, g+8, g+9 (g = first voxel of the group); from the adjacent groups of
    // the warp where they exist, otherwise straight from memory
    const unsigned g0 = lane & ~7u;
    const int gx = i & ~7;
    float n1a = __shfl_sync(0xffffffffu, a_self, (g0 + 8u) & 31u), n1m = __shfl_sync(0xffffffffu, m_self, (g0 + 8u) & 31u);
    float n2a = __shfl_sync(0xffffffffu, a_self, (g0 + 9u) & 31u);
    float p1a = __shfl_sync(0xffffffffu, a_self, (g0 + 31u) & 31u);
    if (g0 == 24u) {
        const float v1 = (row && gx + 8 < dst.nx) ? src[base + gx + 8] : 0.0f;
        const float v2 = (row && gx + 9 < dst.nx) ? src[base + gx + 9] : 0.0f;
        n1a = fabsf(v1); n1m = fmaxf(v1, 0.0f); n2a = fabsf(v2);
    }
    if (g0 == 0u) p1a = (row && gx >= 1 && gx - 1 < dst.nx) ? fabsf(src[base + gx - 1]) : 0.0f;
    if ((lane & 7u) == 0u && row && gx < dst.nx) {
        const size_t o = (size_t)(gx >> 3) + (size_t)dst.nbx * ((size_t)j + (size_t)dst.ny * k);
        const float na = fmaxf(a, n1a);
        ax_abs[o] = na;
        ax_max[o] = fmaxf(m, n1m);
        ax_wide[o] = fmaxf(fmaxf(na, n2a), p1a);
    }
}

// The same pass with four voxels per thread (128-bit loads and stores; a brick row of 8 voxels is a pair of lanes): a third
// of the instructions per voxel.  Needs nx % 4 == 0 and 16-byte aligned rows; LINEAR and BRICK destinations.
template <int LAYOUT, bool STORE>
__global__ void __launch_bounds__(256) retile_tables4_kernel(const float* __restrict__ src, VolView dst, float* __restrict__ ax_abs,
                                                             float* __restrict__ ax_max, float* __restrict__ ax_wide)
{
    const unsigned lane = threadIdx.x;                       // blockDim = (32, 8): one warp per row of 128 voxels
    const int i = (blockIdx.x * 32 + (int)lane) * 4;
    const int j = blockIdx.y * 8 + threadIdx.y;
    const int k = blockIdx.z;
    const bool row = j < dst.ny;
    const size_t base = (size_t)dst.nx * ((size_t)(row ? j : 0) + (size_t)dst.ny * k);
    const bool in = row && i < dst.nx;                       // nx % 4 == 0: all four voxels or none
    float4 v = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    if (in) v = __ldcs(reinterpret_cast<const float4*>(src + base + i));
    if (STORE && in) *reinterpret_cast<float4*>(const_cast<float*>(dst.data) + VolAcc<LAYOUT>::index(dst, i, j, k)) = v;
    // out-of-range lanes contribute the identity of both statistics (the clamped sets only repeat edge voxels)
    const float first_a = fabsf(v.x), second_a = fabsf(v.y), last_a = fabsf(v.w), first_m = fmaxf(v.x, 0.0f);
    float a = fmaxf(fmaxf(first_a, second_a), fmaxf(fabsf(v.z), last_a));
    float m = fmaxf(fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w)), 0.0f);
    a = fmaxf(a, __shfl_xor_sync(0xffffffffu, a, 1));
    m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
    // neighbours of the 8-voxel group: x = g-1, g+8, g+9 (g = first voxel of the group); from the adjacent lanes of the
    // warp where they exist, otherwise straight from memory
    const unsigned p0 = lane & ~1u;
    const int gx = i & ~7;
    float n1a = __shfl_sync(0xffffffffu, first_a, (p0 + 2u) & 31u), n1m = __shfl_sync(0xffffffffu, first_m, (p0 + 2u) & 31u);
    float n2a = __shfl_sync(0xffffffffu, second_a, (p0 + 2u) & 31u);
    float p1a = __shfl_sync(0xffffffffu, last_a, (p0 + 31u) & 31u);
    if (p0 == 30u) {
        const float v1 = (row && gx + 8 < dst.nx) ? src[base + gx + 8] : 0.0f;
        const float v2 = (row && gx + 9 < dst.nx) ? src[base + gx + 9] : 0.0f;
        n1a = fabsf(v1); n1m = fmaxf(v1, 0.0f); n2a = fabsf(v2);
    }
    if (p0 == 0u) p1a = (row && gx >= 1 && gx - 1 < dst.nx) ? fabsf(src[base + gx - 1]) : 0.0f;
    if ((lane & 1u) == 0u && row && gx < dst.nx) {
        const size_t o = (size_t)(gx >> 3) + (size_t)dst.nbx * ((size_t)j + (size_t)dst.ny * k);
        const float na = fmaxf(a, n1a);
        ax_abs[o] = na;
        ax_max[o] = fmaxf(m, n1m);
        ax_wide[o] = fmaxf(fmaxf(na, n2a), p1a);
    }
}

// one thread per brick, lanes along bx (coalesced rows of the pass-X arrays)
__global__ void __launch_bounds__(128) tables_yz_kernel(const float* __restrict__ ax_abs, const float* __restrict__ ax_max,
                                                        const float* __restrict__ ax_wide, int nbx, int nby, int nbz, int ny, int nz,
                                                        float* __restrict__ out_abs, float* __restrict__ out_max, float* __restrict__ out_wide)
{
    const int bx = blockIdx.x * blockDim.x + threadIdx.x;
    const int by = blockIdx.y, bz = blockIdx.z;
    if (bx >= nbx) return;
    float na = 0.0f, nm = 0.0f, wa = 0.0f;
    for (int dz = -1; dz <= 9; ++dz) {
        const int k = min(max(bz * 8 + dz, 0), nz - 1);
        for (int dy = -1; dy <= 9; ++dy) {
            const int j = min(max(by * 8 + dy, 0), ny - 1);
            const size_t o = (size_t)bx + (size_t)nbx * ((size_t)j + (size_t)ny * k);
            wa = fmaxf(wa, ax_wide[o]);
            if (dz >= 0 && dz <= 8 && dy >= 0 && dy <= 8) {
                na = fmaxf(na, ax_abs[o]);
                nm = fmaxf(nm, ax_max[o]);
            }
        }
    }
    const size_t b = (size_t)bx + (size_t)nbx * ((size_t)by + (size_t)nby * bz);
    out_abs[b] = na;
    out_max[b] = nm;
    out_wide[b] = wa;
}

// majorant = max over the brick-max grid (each voxel belongs to at least one brick's set)
__global__ void majorant_kernel(const float* __restrict__ brick_max, long long n, float* __restrict__ out)
{
    float m = 0.0f;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        m = fmaxf(m, brick_max[i]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    // non-negative floats order like their bit patterns: integer atomicMax is exact
    if ((threadIdx.x & 31) == 0) atomicMax(reinterpret_cast<int*>(out), __float_as_int(m));
}

}  // namespace

int crv_launch_retile(crv_ctx* ctx, const float* linear_src, crv_volume* dst, cudaStream_t s)
{
    const dim3 block(64, 4, 1);
    const dim3 grid((dst->nx + block.x - 1) / block.x, (dst->ny + block.y - 1) / block.y, dst->nz);
    const VolView v = dst->view();
    if (dst->layout == CRV_LAYOUT_BRICK)
        retile_kernel<CRV_LAYOUT_BRICK><<<grid, block, 0, s>>>(linear_src, v);
    else if (dst->layout == CRV_LAYOUT_TEX3D)
        retile_kernel<CRV_LAYOUT_TEX3D><<<grid, block, 0, s>>>(linear_src, v);
    else
        retile_kernel<CRV_LAYOUT_LINEAR><<<grid, block, 0, s>>>(linear_src, v);
    ctx->launches++;
    CRV_CUDA(cudaGetLastError());
    return CRV_OK;
}

// Occupancy extent: min / max brick index per axis over the bricks whose wide maximum is not zero (crv_volume::occ_ext).  The
// path tracer turns it into a world-space box (segments end where they leave it), the shadow-volume build into the range of
// chunks of a column bundle that can meet density at all.
__global__ void occ_init_kernel(int* __restrict__ ext)
{
    if (threadIdx.x < 6) ext[threadIdx.x] = threadIdx.x < 3 ? 0x7fffffff : -1;
}
__global__ void occ_extent_kernel(const float* __restrict__ wide, int nbx, int nby, int nbz, int* __restrict__ ext)
{
    const int bx = blockIdx.x * blockDim.x + threadIdx.x, by = blockIdx.y * blockDim.y + threadIdx.y, bz = blockIdx.z;
    bool occ = false;
    if (bx < nbx && by < nby) occ = wide[((size_t)bz * nby + by) * nbx + bx] != 0.0f;
    // one atomic per warp and bound: the warp's lanes share bz and span a few (bx, by)
    const unsigned m = __ballot_sync(0xffffffffu, occ);
    if (m == 0u) return;
    int lox = occ ? bx : 0x7fffffff, hix = occ ? bx : -1, loy = occ ? by : 0x7fffffff, hiy = occ ? by : -1;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lox = min(lox, __shfl_xor_sync(0xffffffffu, lox, o));
        hix = max(hix, __shfl_xor_sync(0xffffffffu, hix, o));
        loy = min(loy, __shfl_xor_sync(0xffffffffu, loy, o));
        hiy = max(hiy, __shfl_xor_sync(0xffffffffu, hiy, o));
    }
    if (((threadIdx.y * blockDim.x + threadIdx.x) & 31u) == 0u) {
        atomicMin(ext + 0, lox); atomicMin(ext + 1, loy); atomicMin(ext + 2, bz);
        atomicMax(ext + 3, hix); atomicMax(ext + 4, hiy); atomicMax(ext + 5, bz);
    }
}

static int launch_occ_extent(crv_ctx* ctx, crv_volume* vol, cudaStream_t s)
{
    if (!vol->occ_ext) CRV_CUDA(cudaMalloc(&vol->occ_ext, 6 * sizeof(int)));
    occ_init_kernel<<<1, 32, 0, s>>>(vol->occ_ext);
    const dim3 block(32, 8, 1), grid((vol->nbx + 31) / 32, (vol->nby + 7) / 8, vol->nbz);
    occ_extent_kernel<<<grid, block, 0, s>>>(vol->brick_absmax_wide, (int)vol->nbx, (int)vol->nby, (int)vol->nbz, vol->occ_ext);
    ctx->launches += 2;
    CRV_CUDA(cudaGetLastError());
    return CRV_OK;
}

// Upload path: layout conversion (when `store`) + all three brick tables in two passes over a linear x-fastest source.
// `scratch` holds the pass-X arrays: 3 * nbx * ny * nz floats.
int crv_launch_retile_tables(crv_ctx* ctx, const float* linear_src, crv_volume* dst, bool store, float* scratch, cudaStream_t s)
{
    const long long nbricks = (long long)dst->nbx * dst->nby * dst->nbz;
    if (!dst->brick_max) CRV_CUDA(cudaMalloc(&dst->brick_max, sizeof(float) * nbricks));
    if (!dst->brick_absmax) CRV_CUDA(cudaMalloc(&dst->brick_absmax, sizeof(float) * nbricks));
    if (!dst->brick_absmax_wide) CRV_CUDA(cudaMalloc(&dst->brick_absmax_wide, sizeof(float) * nbricks));
    const size_t na = (size_t)dst->nbx * dst->ny * dst->nz;
    float *ax_abs = scratch, *ax_max = scratch + na, *ax_wide = scratch + 2 * na;
    const dim3 block(64, 4, 1);
    const dim3 grid((dst->nx + 63) / 64, (dst->ny + 3) / 4, dst->nz);
    const VolView v = dst->view();
    const bool wide = dst->nx % 4 == 0 && (reinterpret_cast<uintptr_t>(linear_src) & 15u) == 0 && dst->layout != CRV_LAYOUT_TEX3D;
    const dim3 block4(32, 8, 1);
    const dim3 grid4((dst->nx / 4 + 31) / 32, (dst->ny + 7) / 8, dst->nz);
    if (wide && !store) retile_tables4_kernel<CRV_LAYOUT_LINEAR, false><<<grid4, block4, 0, s>>>(linear_src, v, ax_abs, ax_max, ax_wide);
    else if (wide && dst->layout == CRV_LAYOUT_BRICK) retile_tables4_kernel<CRV_LAYOUT_BRICK, true><<<grid4, block4, 0, s>>>(linear_src, v, ax_abs, ax_max, ax_wide);
    else if (wide) retile_tables4_kernel<CRV_LAYOUT_LINEAR, true><<<grid4, block4, 0, s>>>(linear_src, v, ax_abs, ax_max, ax_wide);
    else if (!store) retile_tables_kernel<CRV_LAYOUT_LINEAR, false><<<grid, block, 0, s>>>(linear_src, v, ax_abs, ax_max, ax_wide);
    else if (dst->layout == CRV_LAYOUT_BRICK) retile_tables_kernel<CRV_LAYOUT_BRICK, true><<<grid, block, 0, s>>>(linear_src, v, ax_abs, ax_max, ax_wide);
    else if (dst->layout == CRV_LAYOUT_TEX3D) retile_tables_kernel<CRV_LAYOUT_TEX3D, true><<<grid, block, 0, s>>>(linear_src, v, ax_abs, ax_max, ax_wide);
    else retile_tables_kernel<CRV_LAYOUT_LINEAR, true><<<grid, block, 0, s>>>(linear_src, v, ax_abs, ax_max, ax_wide);
    const dim3 grid2((dst->nbx + 127) / 128, dst->nby, dst->nbz);
    tables_yz_kernel<<<grid2, 128, 0, s>>>(ax_abs, ax_max, ax_wide, (int)dst->nbx, (int)dst->nby, (int)dst->nbz, (int)dst->ny, (int)dst->nz,
                                            dst->brick_absmax, dst->brick_max, dst->brick_absmax_wide);
    ctx->launches += 2;
    CRV_CUDA(cudaGetLastError());
    {
        const int rc = launch_occ_extent(ctx, dst, s);
        if (rc != CRV_OK) return rc;
    }
    dst->brick_max_valid = true;
    return CRV_OK;
}

int crv_launch_untile(crv_ctx* ctx, const crv_volume* src, float* linear_dst, cudaStream_t s)
{
    const dim3 block(64, 4, 1);
    const dim3 grid((src->nx + block.x - 1) / block.x, (src->ny + block.y - 1) / block.y, src->nz);
    const VolView v = src->view();
    if (src->layout == CRV_LAYOUT_BRICK)
        untile_kernel<CRV_LAYOUT_BRICK><<<grid, block, 0, s>>>(v, linear_dst);
    else if (src->layout == CRV_LAYOUT_TEX3D)
        untile_kernel<CRV_LAYOUT_TEX3D><<<grid, block, 0, s>>>(v, linear_dst);
    else
        untile_kernel<CRV_LAYOUT_LINEAR><<<grid, block, 0, s>>>(v, linear_dst);
    ctx->launches++;
    CRV_CUDA(cudaGetLastError());
    return CRV_OK;
}

int crv_launch_brick_max(crv_ctx* ctx, crv_volume* vol, cudaStream_t s)
{
    const long long nbricks = (long long)vol->nbx * vol->nby * vol->nbz;
    if (!vol->brick_max) CRV_CUDA(cudaMalloc(&vol->brick_max, sizeof(float) * nbricks));
    if (!vol->brick_absmax) CRV_CUDA(cudaMalloc(&vol->brick_absmax, sizeof(float) * nbricks));
    const int threads = 256, wpb = threads / 32;
    const unsigned blocks = (unsigned)((nbricks + wpb - 1) / wpb);
    const VolView v = vol->view();
    if (vol->layout == CRV_LAYOUT_BRICK)
        brick_max_kernel<CRV_LAYOUT_BRICK><<<blocks, threads, 0, s>>>(v, vol->brick_absmax, vol->brick_max);
    else if (vol->layout == CRV_LAYOUT_TEX3D)
        brick_max_kernel<CRV_LAYOUT_TEX3D><<<blocks, threads, 0, s>>>(v, vol->brick_absmax, vol->brick_max);
    else
        brick_max_kernel<CRV_LAYOUT_LINEAR><<<blocks, threads, 0, s>>>(v, vol->brick_absmax, vol->brick_max);
    if (!vol->brick_absmax_wide) CRV_CUDA(cudaMalloc(&vol->brick_absmax_wide, sizeof(float) * nbricks));
    if (vol->layout == CRV_LAYOUT_BRICK)
        brick_absmax_wide_kernel<CRV_LAYOUT_BRICK><<<blocks, threads, 0, s>>>(v, vol->brick_absmax_wide);
    else if (vol->layout == CRV_LAYOUT_TEX3D)
        brick_absmax_wide_kernel<CRV_LAYOUT_TEX3D><<<blocks, threads, 0, s>>>(v, vol->brick_absmax_wide);
    else
        brick_absmax_wide_kernel<CRV_LAYOUT_LINEAR><<<blocks, threads, 0, s>>>(v, vol->brick_absmax_wide);
    ctx->launches += 2;
    CRV_CUDA(cudaGetLastError());
    {
        const int rc = launch_occ_extent(ctx, vol, s);
        if (rc != CRV_OK) return rc;
    }
    vol->brick_max_valid = true;
    return CRV_OK;
}

int crv_launch_majorant(crv_ctx* ctx, crv_volume* vol, cudaStream_t s)
{
    if (!vol->brick_max_valid) {
        const int rc = crv_launch_brick_max(ctx, vol, s);
        if (rc != CRV_OK) return rc;
    }
    float* d_out = nullptr;
    CRV_CUDA(cudaMalloc(&d_out, sizeof(float)));
    CRV_CUDA(cudaMemsetAsync(d_out, 0, sizeof(float), s));
    const long long n = (long long)vol->nbx * vol->nby * vol->nbz;
    const int blocks = (int)((n + 255) / 256 < 1024 ? (n + 255) / 256 : 1024);
    majorant_kernel<<<blocks, 256, 0, s>>>(vol->brick_max, n, d_out);
    ctx->launches++;
    CRV_CUDA(cudaGetLastError());
    CRV_CUDA(cudaMemcpyAsync(&vol->majorant, d_out, sizeof(float), cudaMemcpyDeviceToHost, s));
    CRV_CUDA(cudaStreamSynchronize(s));
    CRV_CUDA(cudaFree(d_out));
    vol->majorant_valid = true;
    return CRV_OK;
}

// ------------------------------------------------------------------------------------------------------
// Present step.  After the dispatch the reference blits the RGBA32F result image 1:1 into the swapchain image,
// VK_FORMAT_B8G8R8A8_SRGB (RenderTechniquePT.cpp:143-167, VulkanSwapchain.cpp:68-79): per channel clamp to [0,1],
// linear -> sRGB transfer function on R, G, B (alpha stays linear), round(x * 255) to UNORM8, stored B, G, R, A.
// Same size and texel centres, so VK_FILTER_LINEAR reduces to a copy.
// ------------------------------------------------------------------------------------------------------
namespace {
__device__ __forceinline__ float srgb_oetf(float l)
{
    const float c = fminf(fmaxf(l, 0.0f), 1.0f);                 // NaN -> 0
    return c <= 0.0031308f ? c * 12.92f : 1.055f * powf(c, 1.0f / 2.4f) - 0.055f;
}
__device__ __forceinline__ unsigned unorm8(float v) { return (unsigned)__float2int_rn(fminf(fmaxf(v, 0.0f), 1.0f) * 255.0f); }

__global__ void blit_srgb8_kernel(const float4* __restrict__ src, uchar4* __restrict__ dst, size_t n)
{
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 p = src[i];
    dst[i] = make_uchar4((unsigned char)unorm8(srgb_oetf(p.z)), (unsigned char)unorm8(srgb_oetf(p.y)),
                         (unsigned char)unorm8(srgb_oetf(p.x)), (unsigned char)unorm8(p.w));
}
}  // namespace

int crv_launch_blit_srgb8(crv_ctx* ctx, const float* rgba, unsigned char* bgra8, size_t npix, cudaStream_t s)
{
    blit_srgb8_kernel<<<(unsigned)((npix + 255) / 256), 256, 0, s>>>(reinterpret_cast<const float4*>(rgba), reinterpret_cast<uchar4*>(bgra8), npix);
    ctx->launches++;
    CRV_CUDA(cudaGetLastError());
    return CRV_OK;
}
