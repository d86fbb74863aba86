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
#include <cmath>

#include "crv_internal.h"
#include "exact_div.cuh"
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

// ------------------------------------------------------------------------------------------------------
// Fast kernels.  Same results bit for bit as sv_sweep_kernel (test_shadow_volume_modes_bit_identical).
// The literal kernel is issue- and latency-bound, not memory-bound (profiles/r1a_sv_ncu.txt: 81 warp
// instructions per 32 voxels, DRAM at 20 %).  Shared ingredients:
//  * Each coordinate of p(z) = b0 + ((a + z*light) * voxelSize) is a chain of correctly rounded operations that
//    are each weakly monotonic in z (voxelSize > 0), so per axis the set {z : cb0 <= p(z) <= cb1} is an interval
//    and the column is inside the cloud AABB exactly for z in [z_in, z_end).  The interval is found once per
//    column by bisection on the SAME fp32 expressions; the per-voxel position / AABB test disappears outside it
//    (81 % of the voxels for a cube inside its bounding sphere's cube).
//  * Before the column meets non-zero density acc == 0 and the shader's value is exp(-0) = 1: store only.
//  * All divisions have launch-uniform divisors -> exact FMA division with host-computed reciprocals
//    (exact_div.cuh); the reciprocal of float(z+1) comes from an n-entry table.
//  * Column tiles are issued centre-out so that the cheap corner tiles (columns that never enter the cloud)
//    fill the tail of the last wave; stores are streaming (written once, read later by another kernel).
// ------------------------------------------------------------------------------------------------------
constexpr int SV_ROWS = 8;

// c_rcp_z[z] = RN(1 / float(z + 1)), filled once per device by crv_sv_init_tables (host IEEE division).  The index is
// warp-uniform, so the lookup is one constant-bank load.
__constant__ float c_rcp_z[SV_MAX_FAST_N];

// [z_in, z_end): the z-interval in which column `a` (= x*right + y*up) is inside the cloud AABB (:34-42)
__device__ __forceinline__ void column_range(const SVConst& c, const float a[3], int& z_in, int& z_end)
{
    const int N = c.n;
    z_in = 0;
    z_end = N;
#pragma unroll
    for (int ax = 0; ax < 3; ++ax) {
        const bool inc = c.light[ax] >= 0.0f;
        int lo_a = 0, hi_a = N, lo_b = 0, hi_b = N;
        for (int it = 0; it < c.search_iters; ++it) {
            if (lo_a < hi_a) {      // first z with  (inc ? p >= cb0 : p <= cb1)
                const int mid = (lo_a + hi_a) >> 1;
                const float p = c.b0[ax] + ((a[ax] + (float)mid * c.light[ax]) * c.voxel_size);
                const bool t = inc ? !(p < c.cb0[ax]) : !(p > c.cb1[ax]);
                if (t) hi_a = mid; else lo_a = mid + 1;
            }
            if (lo_b < hi_b) {      // first z with  (inc ? p > cb1 : p < cb0)
                const int mid = (lo_b + hi_b) >> 1;
                const float p = c.b0[ax] + ((a[ax] + (float)mid * c.light[ax]) * c.voxel_size);
                const bool t = inc ? (p > c.cb1[ax]) : (p < c.cb0[ax]);
                if (t) hi_b = mid; else lo_b = mid + 1;
            }
        }
        z_in = max(z_in, lo_a);
        z_end = min(z_end, lo_b);
    }
    if (z_end < z_in) z_end = z_in;
}

// texture(cloud, p(z)) for column `a` (:44-56), p known to be inside the cloud AABB (so 0 <= s <= 1).
// SKIP: consult the dilated brick-max table first -- if every texel the footprint can touch is zero the
// trilinear value is exactly 0 and the 8 loads (and the DRAM traffic of empty space) are skipped.
template <int CL, bool SKIP>
__device__ __forceinline__ float column_sample(const SVConst& c, const float a[3], int z)
{
    const float fz = (float)z;
    const float px = c.b0[0] + ((a[0] + fz * c.light[0]) * c.voxel_size);
    const float py = c.b0[1] + ((a[1] + fz * c.light[1]) * c.voxel_size);
    const float pz = c.b0[2] + ((a[2] + fz * c.light[2]) * c.voxel_size);
    // sampleCloud: ((p,0) - b0) / (b1 - b0); a -0 numerator gives +0 here instead of -0: same u = s*size - 0.5
    const float sx = div_rn_by(px - c.cb0[0], c.cext[0], c.rcext[0]);
    const float sy = div_rn_by(py - c.cb0[1], c.cext[1], c.rcext[1]);
    const float sz = div_rn_by(pz - c.cb0[2], c.cext[2], c.rcext[2]);
    const Footprint f = make_footprint<true>(c.cloud, sx, sy, sz);
    if (SKIP) {
        const unsigned b = ((unsigned)(f.k0 >> CRV_BRICK_SHIFT) * c.cloud.nby + (unsigned)(f.j0 >> CRV_BRICK_SHIFT)) * c.cloud.nbx +
                           (unsigned)(f.i0 >> CRV_BRICK_SHIFT);
        if (__ldg(c.brick_absmax + b) == 0.0f) return 0.0f;      // lerps of +-0 texels: +-0, and acc + (+-0) = acc
    }
    return trilinear_fetch<CL>(c.cloud, f);
}

// centre-out tile permutation: i = 0,1,2,3,... -> mid, mid-1, mid+1, mid-2, ... (a permutation of [0, n))
__device__ __forceinline__ int centre_out(int i, int n)
{
    const int k = (i + 1) >> 1;
    return (i & 1) ? (n >> 1) - k : (n >> 1) + k;
}

// ---- single pass: one thread per column, gathers of U slices batched ahead of the serial scan -------------
template <int CL, int OL, int U>
__global__ void __launch_bounds__(32 * SV_ROWS) sv_sweep_fast_kernel(const __grid_constant__ SVConst c)
{
    const int N = c.n;
    const int x = centre_out(blockIdx.x, gridDim.x) * 32 + threadIdx.x;
    const int y = centre_out(blockIdx.y, gridDim.y) * SV_ROWS + threadIdx.y;
    const bool live = x < N && y < N;
    const unsigned mask = __ballot_sync(0xffffffffu, live);
    if (!live) return;

    const float fx = (float)x, fy = (float)y;
    const float a[3] = {fx * c.right[0] + fy * c.up[0], fx * c.right[1] + fy * c.up[1], fx * c.right[2] + fy * c.up[2]};
    int z_in, z_end;
    column_range(c, a, z_in, z_end);

    float acc = 0.0f;      // accumulatedValue.x
    float dist = 0.0f;     // accumulatedDistance
    float fz1 = 0.0f;      // float(z + 1), exact
    for (int z0 = 0; z0 < N; z0 += U) {
        float tex[U];
#pragma unroll
        for (int u = 0; u < U; ++u) tex[u] = 0.0f;
        // gather pass (skipped by the whole warp when no lane is inside the AABB in this chunk)
        const bool mine = (z0 + U > z_in) && (z0 < z_end);
        if (__any_sync(mask, mine)) {
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int z = z0 + u;
                if (z >= z_in && z < z_end) tex[u] = column_sample<CL, false>(c, a, z);
            }
        }
        // serial pass: running sums, transmittance, store
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int z = z0 + u;
            if (z < N) {
                // outside [z_in, z_end) the shader adds nothing; inside, adding (0*ds)/bs = 0 is a no-op too
                if (tex[u] != 0.0f) acc = acc + div_rn_by(tex[u] * c.density_scaling, c.base_scaling, c.rbase);
                dist = dist + c.voxel_size;
                fz1 = fz1 + 1.0f;
                float t = 1.0f;    // acc == 0: exp(((-dist)*0)/(z+1)) = exp(-0) = 1 exactly
                if (acc != 0.0f) t = expf(div_rn_by((-dist) * acc, fz1, c_rcp_z[z]));
                VolAcc<OL>::store_streaming(c.out, x, y, z, t);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------------
// Two-pass build.  The single-pass kernels stay latency-bound: every column alternates between a batch of
// dependent-address gathers and its serial scan, and at N = 512 the whole grid is a single wave
// (profiles/r1b_svfast_512_ncu.txt: 38 % of the stall samples wait on the gathers, DRAM at 19 %).  The gathers do
// not depend on the running sum, so they are split off:
//   pass 0  (prologue of sv_gather_kernel) [z_in, z_end) of the CTA's 64 columns, kept for pass 2
//   pass 1  sv_gather_kernel : for the voxels inside the cloud AABB only, the addend
//           (texture*densityScaling)/baseScaling, stored IN PLACE in the output volume.  Fully parallel: one CTA
//           = one 8x8 column bundle, 8 slices at a time (a compact block of samples, so the texel loads of a
//           warp share cache lines), all loads independent.
//   pass 2  sv_scan_kernel   : per column, serial in z: acc += addend (read back only inside the interval,
//           coalesced and issued U slices ahead of the recurrence), exp, streaming store.  A pure stream.
// Same operations in the same order per column as the shader => same bits as the literal kernel.
// Extra traffic over the algorithmic 4 N^3 + 4 Ncloud^3: one write + one read of the in-AABB voxels.
// ------------------------------------------------------------------------------------------------------
// one of the six monotone predicates of column_range: first z in [0, N] at which it becomes true
__device__ __forceinline__ int range_bisect(const SVConst& c, float a_ax, int ax, bool leave)
{
    const int N = c.n;
    const bool inc = c.light[ax] >= 0.0f;
    int lo = 0, hi = N;
    for (int it = 0; it < c.search_iters; ++it) {
        if (lo < hi) {
            const int mid = (lo + hi) >> 1;
            const float p = c.b0[ax] + ((a_ax + (float)mid * c.light[ax]) * c.voxel_size);
            const bool t = leave ? (inc ? (p > c.cb1[ax]) : (p < c.cb0[ax])) : (inc ? !(p < c.cb0[ax]) : !(p > c.cb1[ax]));
            if (t) hi = mid; else lo = mid + 1;
        }
    }
    return lo;
}

// The same first-true index, found from the real-number crossing instead of by bisection: the predicate is monotone in z
// (every operation of the position chain is weakly monotone), so the candidate ceil(z*) is put right by stepping while the
// predicate disagrees -- the result is the bisection's, for two or three evaluations instead of log2(N) + 1.
__device__ __forceinline__ int range_seek(const SVConst& c, float a_ax, int ax, bool leave)
{
    const int N = c.n;
    const bool inc = c.light[ax] >= 0.0f;
    auto pred = [&](int z) -> bool {
        const float p = c.b0[ax] + ((a_ax + (float)z * c.light[ax]) * c.voxel_size);
        return leave ? (inc ? (p > c.cb1[ax]) : (p < c.cb0[ax])) : (inc ? !(p < c.cb0[ax]) : !(p > c.cb1[ax]));
    };
    // position reaches `bound` at z* = ((bound - b0) / voxel_size - a_ax) / light
    const float bound = (leave == inc) ? c.cb1[ax] : c.cb0[ax];
    const float zs = __fdividef(__fdividef(bound - c.b0[ax], c.voxel_size) - a_ax, c.light[ax]);
    int k = (zs == zs) ? (int)fminf(fmaxf(ceilf(zs), 0.0f), (float)N) : 0;      // light == 0: NaN / inf, any start will do
    int guard = 0;
    while (k > 0 && pred(k - 1) && guard < 8) { --k; ++guard; }
    while (k < N && !pred(k) && guard < 16) { ++k; ++guard; }
    if (guard >= 8) return range_bisect(c, a_ax, ax, leave);                    // estimate far off (degenerate direction): bisect
    return k;
}

template <int CL, int OL>
__global__ void __launch_bounds__(512) sv_gather_kernel(const __grid_constant__ SVConst c)
{
    __shared__ int s_lo, s_hi;
    __shared__ int s_search[6][64];
    __shared__ int s_zin[64], s_zend[64];
    const int N = c.n;
    const int t = threadIdx.x;

    // ---- pass 0 fused: [z_in, z_end) of the bundle's 64 columns, 6 bisections per column spread over 384 threads ----
    if (t == 0) { s_lo = N; s_hi = 0; }
    if (t < 384) {
        const int col = t & 63, search = t >> 6;
        const int cx = blockIdx.x * 8 + (col & 7), cy = blockIdx.y * 8 + (col >> 3);
        const int ax = search >> 1;
        const float a_ax = (float)cx * c.right[ax] + (float)cy * c.up[ax];
        s_search[search][col] = range_bisect(c, a_ax, ax, (search & 1) != 0);
    } else if (t == 511 && blockIdx.x == 0 && blockIdx.y == 0) {
        float dist = 0.0f;      // accumulatedDistance after every slice: the shader's repeated addition (:91), once
        for (int z = 0; z < N; ++z) {
            dist = dist + c.voxel_size;
            c.dist_z[z] = dist;
        }
    }
    __syncthreads();
    if (t < 64) {
        const int cx = blockIdx.x * 8 + (t & 7), cy = blockIdx.y * 8 + (t >> 3);
        int z_in = max(max(s_search[0][t], s_search[2][t]), s_search[4][t]);
        int z_end = min(min(s_search[1][t], s_search[3][t]), s_search[5][t]);
        if (z_end < z_in) z_end = z_in;
        if (cx >= N || cy >= N) z_in = z_end = 0;
        else c.zrange[(size_t)cy * N + cx] = make_ushort2((unsigned short)z_in, (unsigned short)z_end);
        s_zin[t] = z_in;
        s_zend[t] = z_end;
        if (z_in < z_end) {
            atomicMin(&s_lo, z_in);
            atomicMax(&s_hi, z_end);
        }
    }
    __syncthreads();
    const int lo = s_lo & ~7, hi = s_hi;
    if (lo >= hi) return;       // the bundle never enters the cloud

    // ---- pass 1: the in-AABB addends, 8x8 columns x 8 slices per step (two steps in flight) ----
    const int lx = t & 7, ly = (t >> 3) & 7, lz = t >> 6;
    const int x = blockIdx.x * 8 + lx, y = blockIdx.y * 8 + ly;
    const int z_in = s_zin[t & 63], z_end = s_zend[t & 63];
    const float fx = (float)x, fy = (float)y;
    const float a[3] = {fx * c.right[0] + fy * c.up[0], fx * c.right[1] + fy * c.up[1], fx * c.right[2] + fy * c.up[2]};
    for (int z0 = lo + lz; z0 < hi; z0 += 16) {
        const int z1 = z0 + 8;
        const bool in0 = z0 >= z_in && z0 < z_end, in1 = z1 >= z_in && z1 < z_end;
        float t0 = 0.0f, t1 = 0.0f;
        if (in0) t0 = column_sample<CL, true>(c, a, z0);
        if (in1) t1 = column_sample<CL, true>(c, a, z1);
        if (in0) VolAcc<OL>::store(c.out, x, y, z0, div_rn_by(t0 * c.density_scaling, c.base_scaling, c.rbase));
        if (in1) VolAcc<OL>::store(c.out, x, y, z1, div_rn_by(t1 * c.density_scaling, c.base_scaling, c.rbase));
    }
}

// Column accessor for an aligned chunk of 8 slices: one base pointer + a per-slice step in the global layouts.
template <int OL>
struct ColumnChunk {
    float* p;
    unsigned step;
    __device__ __forceinline__ ColumnChunk(const VolView& v, unsigned off_xy, int, int, int z0)
        : p(const_cast<float*>(v.data) + (off_xy + VolAcc<OL>::off_z(v, z0))), step(VolAcc<OL>::z_step(v)) {}
    __device__ __forceinline__ void store(const VolView&, int u, float val) const { __stcs(p + u * step, val); }
    __device__ __forceinline__ float load(const VolView&, int u) const { return __ldcs(p + u * step); }
};
template <>
struct ColumnChunk<CRV_LAYOUT_TEX3D> {
    int x, y, z0;
    __device__ __forceinline__ ColumnChunk(const VolView&, unsigned, int x_, int y_, int z0_) : x(x_), y(y_), z0(z0_) {}
    __device__ __forceinline__ void store(const VolView& v, int u, float val) const { VolAcc<CRV_LAYOUT_TEX3D>::store(v, x, y, z0 + u, val); }
    __device__ __forceinline__ float load(const VolView& v, int u) const { return VolAcc<CRV_LAYOUT_TEX3D>::load_rw(v, x, y, z0 + u); }
};

// One thread per column; chunks of 8 slices (one brick row in z).  A chunk in which no lane of the warp has met
// density yet (acc == 0 everywhere, nobody inside its interval) is all ones: 8 streaming stores and nothing else --
// the running distance is rebuilt from the dist table when the warp leaves that regime (same repeated-addition values).
template <int OL, bool FULL>
__device__ __forceinline__ void scan_chunk(const SVConst& c, const ColumnChunk<OL>& col, int z0, int z_in, int z_end, bool mine,
                                           float& acc, float& dist, float& fz1)
{
    constexpr int U = 8;
    float inc[U];
#pragma unroll
    for (int u = 0; u < U; ++u) inc[u] = 0.0f;
    if (mine) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int z = z0 + u;
            if (z >= z_in && z < z_end) inc[u] = col.load(c.out, u);
        }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
        if (FULL || z0 + u < c.n) {
            acc = acc + inc[u];        // + 0 outside the interval leaves acc unchanged (acc is never -0)
            dist = dist + c.voxel_size;
            fz1 = fz1 + 1.0f;
            // no branch on acc == 0 (it would serialise the 8 independent exp chains): ((-dist)*0)/(z+1) = +-0 and
            // expf(+-0) = 1 exactly, the value the shader stores there
            col.store(c.out, u, expf(div_rn_by((-dist) * acc, fz1, c_rcp_z[z0 + u])));
        }
    }
}

template <int OL>
__global__ void __launch_bounds__(32 * SV_ROWS) sv_scan_kernel(const __grid_constant__ SVConst c)
{
    constexpr int U = 8;
    const int N = c.n;
    const int x = centre_out(blockIdx.x, gridDim.x) * 32 + threadIdx.x;
    const int y = centre_out(blockIdx.y, gridDim.y) * SV_ROWS + threadIdx.y;
    const bool live = x < N && y < N;
    const unsigned mask = __ballot_sync(0xffffffffu, live);
    if (!live) return;
    const ushort2 r = c.zrange[(size_t)y * N + x];
    const int z_in = r.x, z_end = r.y;
    unsigned off_xy = 0;
    if constexpr (OL != CRV_LAYOUT_TEX3D) off_xy = VolAcc<OL>::off_x(c.out, x) + VolAcc<OL>::off_y(c.out, y);

    float acc = 0.0f;      // accumulatedValue.x
    float dist = 0.0f;     // accumulatedDistance
    float fz1 = 0.0f;      // float(z + 1), exact
    bool stale = false;    // warp-uniform: dist / fz1 were not advanced through the last all-ones chunks
    for (int z0 = 0; z0 < N; z0 += U) {
        const bool full = z0 + U <= N;
        const bool mine = (z0 + U > z_in) && (z0 < z_end);
        const ColumnChunk<OL> col(c.out, off_xy, x, y, z0);
        if (!__any_sync(mask, mine || acc != 0.0f)) {
            if (full) {
#pragma unroll
                for (int u = 0; u < U; ++u) col.store(c.out, u, 1.0f);
            } else {
                for (int u = 0; z0 + u < N; ++u) col.store(c.out, u, 1.0f);
            }
            stale = true;
            continue;
        }
        if (stale) {       // dist_z[z] = accumulatedDistance after slice z: the same repeated fp32 additions (:91)
            dist = __ldg(c.dist_z + (z0 - 1));
            fz1 = (float)z0;
            stale = false;
        }
        if (full) scan_chunk<OL, true>(c, col, z0, z_in, z_end, mine, acc, dist, fz1);
        else scan_chunk<OL, false>(c, col, z0, z_in, z_end, mine, acc, dist, fz1);
    }
}

// ------------------------------------------------------------------------------------------------------
// Fused build (the default).  The two-pass form pays for its parallel gathers with a write + read of every
// in-AABB addend and leaves the scan with one thread per column -- 0.86 waves at N = 512, DRAM at 53 %
// (profiles/r1n_sv_ncu.txt).  Here one CTA owns an 8x8 column bundle from z = 0 to N and walks it in chunks of
// 8 slices, 512 threads = one voxel each:
//   * chunks before any column of the bundle enters the cloud AABB are all ones: one coalesced 2 KiB store;
//   * inside, every thread gathers its voxel's addend (same code as sv_gather_kernel) into shared memory, 64
//     threads add the 8 addends of their column IN THE SHADER'S SERIAL ORDER and leave the 8 running sums
//     there, then every thread does exp + store for its voxel with its own sum;
//   * after the last column has left the AABB the sums are final: exp + store only.
// Nothing is written twice and nothing is read back, the output is stored in whole 8x8x8 blocks (exactly one
// brick of the BRICK layout, 32-byte rows otherwise), and the parallelism is n^2/64 CTAs of 512 threads at any N.
// accumulatedDistance comes from the dist_z table: the shader's repeated fp32 addition (:91), done once per
// (N, voxelSize) by sv_dist_kernel.  Per column: the same operations in the same order => the same bits.
// ------------------------------------------------------------------------------------------------------
// Store cursor of one thread of the fused build: voxel (x, y, z0 + lz) of successive 8-slice chunks.
template <int OL>
struct ChunkCursor {
    float* p;
    unsigned step;
    __device__ __forceinline__ ChunkCursor(const VolView& v, int x, int y, int lz)
        : p(const_cast<float*>(v.data) + (VolAcc<OL>::off_x(v, x) + VolAcc<OL>::off_y(v, y) + VolAcc<OL>::off_z(v, lz))),
          step(VolAcc<OL>::off_z(v, 8)) {}
    __device__ __forceinline__ void store(const VolView&, int, float val) const { __stcs(p, val); }
    __device__ __forceinline__ void advance() { p += step; }
};
template <>
struct ChunkCursor<CRV_LAYOUT_TEX3D> {
    int x, y, lz;
    __device__ __forceinline__ ChunkCursor(const VolView&, int x_, int y_, int lz_) : x(x_), y(y_), lz(lz_) {}
    __device__ __forceinline__ void store(const VolView& v, int z0, float val) const { VolAcc<CRV_LAYOUT_TEX3D>::store(v, x, y, z0 + lz, val); }
    __device__ __forceinline__ void advance() {}
};

__global__ void sv_dist_kernel(float* __restrict__ dist_z, int n, float voxel_size)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    float dist = 0.0f;
    for (int z = 0; z < n; ++z) {
        dist = dist + voxel_size;
        dist_z[z] = dist;
    }
}

template <int CL, int OL>
__global__ void __launch_bounds__(512, 4) sv_fused_kernel(const __grid_constant__ SVConst c)
{
    __shared__ int s_lo, s_hi;
    __shared__ int s_search[6][64];
    __shared__ int s_zin[64], s_zend[64];
    __shared__ float s_add[16][64];     // two chunks' addends, then the running sums after each slice
    __shared__ float s_acc[64];         // running sum of every column after the previous chunk
    const int N = c.n;
    const int t = threadIdx.x;
    const int bx = centre_out(blockIdx.x, gridDim.x), by = centre_out(blockIdx.y, gridDim.y);

    // ---- [z_in, z_end) of the bundle's 64 columns: 6 bisections per column spread over 384 threads ----
    if (t == 0) { s_lo = N; s_hi = 0; }
    if (t < 384) {
        const int col = t & 63, search = t >> 6;
        const int cx = bx * 8 + (col & 7), cy = by * 8 + (col >> 3);
        const int ax = search >> 1;
        const float a_ax = (float)cx * c.right[ax] + (float)cy * c.up[ax];
        s_search[search][col] = range_bisect(c, a_ax, ax, (search & 1) != 0);
    }
    __syncthreads();
    if (t < 64) {
        const int cx = bx * 8 + (t & 7), cy = by * 8 + (t >> 3);
        int z_in = max(max(s_search[0][t], s_search[2][t]), s_search[4][t]);
        int z_end = min(min(s_search[1][t], s_search[3][t]), s_search[5][t]);
        if (z_end < z_in) z_end = z_in;
        if (cx >= N || cy >= N) z_in = z_end = 0;
        s_zin[t] = z_in;
        s_zend[t] = z_end;
        s_acc[t] = 0.0f;
        if (z_in < z_end) {
            atomicMin(&s_lo, z_in);
            atomicMax(&s_hi, z_end);
        }
    }
    __syncthreads();
    const int hi = s_hi;
    const int lo = hi > 0 ? (s_lo & ~7) : N;      // bundle never enters the cloud: every chunk is "before"

    const int col = t & 63, lz = t >> 6;
    const int x = bx * 8 + (t & 7), y = by * 8 + ((t >> 3) & 7);
    const bool inb = x < N && y < N;
    const int z_in = s_zin[col], z_end = s_zend[col];
    const float fx = (float)x, fy = (float)y;
    const float a[3] = {fx * c.right[0] + fy * c.up[0], fx * c.right[1] + fy * c.up[1], fx * c.right[2] + fy * c.up[2]};
    // Approximate brick coordinate of this thread's voxel as a function of z (one FMA per axis).  It tracks the sampler's
    // lower texel to within a quarter texel (bound checked on the host), and the wide table covers every texel a
    // footprint within one texel of it can touch: a zero there proves the trilinear value is +-0 without the exact
    // position / division / footprint chain.  Most of the cloud's bounding box is empty.
    const bool pre = c.brick_absmax_wide != nullptr;
    const float pa[3] = {__fmaf_rn(fy, c.pre_u[0], __fmaf_rn(fx, c.pre_r[0], c.pre_c[0])),
                         __fmaf_rn(fy, c.pre_u[1], __fmaf_rn(fx, c.pre_r[1], c.pre_c[1])),
                         __fmaf_rn(fy, c.pre_u[2], __fmaf_rn(fx, c.pre_r[2], c.pre_c[2]))};

    // One pointer per thread, advanced by a whole chunk per iteration: in both global layouts the offset of slice
    // z0 + lz (z0 a multiple of 8) is off_z(z0) + off_z(lz), and off_z(z0) is z0/8 steps of off_z(8).
    ChunkCursor<OL> cur(c.out, x, y, lz);
    int z0 = 0;
    // ---- before any column of the bundle enters the cloud AABB: exp(-0) = 1 ----
    for (; z0 + 8 <= lo; z0 += 8) {
        if (inb) cur.store(c.out, z0, 1.0f);
        cur.advance();
    }
    // ---- while some column is inside its interval: gather, serial sums, exp -- two chunks (16 slices) per barrier pair ----
    auto addend = [&](int z) -> float {
        if (!(z >= z_in && z < z_end)) return 0.0f;       // + 0 outside the interval leaves the sum unchanged (it is never -0)
        if (pre) {
            const float fz = (float)z;
            // floor by a round-down add of 2^23; inside the AABB the coordinate is >= -1/16 - 1/32, so only -1 needs a clamp
            const int bi = max(__float_as_int(__fadd_rd(__fmaf_rn(fz, c.pre_l[0], pa[0]), 8388608.0f)) - 0x4B000000, 0);
            const int bj = max(__float_as_int(__fadd_rd(__fmaf_rn(fz, c.pre_l[1], pa[1]), 8388608.0f)) - 0x4B000000, 0);
            const int bk = max(__float_as_int(__fadd_rd(__fmaf_rn(fz, c.pre_l[2], pa[2]), 8388608.0f)) - 0x4B000000, 0);
            const unsigned b = ((unsigned)min(bk, c.cloud.nbz - 1) * c.cloud.nby + (unsigned)min(bj, c.cloud.nby - 1)) * c.cloud.nbx +
                               (unsigned)min(bi, c.cloud.nbx - 1);
            if (__ldg(c.brick_absmax_wide + b) == 0.0f) return 0.0f;
        }
        return div_rn_by(column_sample<CL, true>(c, a, z) * c.density_scaling, c.base_scaling, c.rbase);
    };
    // ((-dist)*0)/(z+1) = +-0 and expf(+-0) = 1 exactly, the value the shader stores while a column has met no
    // density: a warp (4x8 columns of one slice) none of whose columns has any skips the divide + exp altogether
    auto transmittance = [&](int z, float acc) -> float {
        float tr = 1.0f;
        if (__any_sync(0xffffffffu, acc != 0.0f) && z < N) tr = expf(div_rn_by((-__ldg(c.dist_z + z)) * acc, (float)(z + 1), c_rcp_z[z]));
        return tr;
    };
    for (; z0 < hi; z0 += 16) {
        const int za = z0 + lz, zb = z0 + 8 + lz;
        s_add[lz][col] = addend(za);
        s_add[8 + lz][col] = addend(zb);
        __syncthreads();
        if (t < 64) {
            float s = s_acc[t];
#pragma unroll
            for (int u = 0; u < 16; ++u) {
                s = s + s_add[u][t];              // accumulatedValue after slice z0 + u, the shader's order (:89)
                s_add[u][t] = s;
            }
            s_acc[t] = s;
        }
        __syncthreads();
        const float tra = transmittance(za, s_add[lz][col]);
        const float trb = transmittance(zb, s_add[8 + lz][col]);
        if (inb && za < N) cur.store(c.out, z0, tra);
        cur.advance();
        if (inb && zb < N) cur.store(c.out, z0 + 8, trb);
        cur.advance();
    }
    // ---- every column has left the AABB: the sums are final ----
    {
        const float acc = s_acc[col];
        const bool any = __any_sync(0xffffffffu, acc != 0.0f);
        for (; z0 < N; z0 += 8) {
            const int z = z0 + lz;
            float tr = 1.0f;
            if (any && z < N) tr = expf(div_rn_by((-__ldg(c.dist_z + z)) * acc, (float)(z + 1), c_rcp_z[z]));
            if (inb && z < N) cur.store(c.out, z0, tr);
            cur.advance();
        }
    }
}

// ------------------------------------------------------------------------------------------------------
// Fused build, brick output (the default): the same per-column arithmetic as sv_fused_kernel, organised around what the
// profile of that kernel showed (profiles/r2c_sv_ncu.txt: 0.93 warp instructions per voxel, issue-active 70 %, DRAM at 41 %):
// it was bound by instructions spent on voxels that end up holding exactly 1.0f.
//   * Every 8-slice chunk of the bundle (an 8^3 block of shadow voxels = one output brick) is classified first, one warp
//     per chunk: the chunk's sample positions lie in a small box of cloud bricks (<= 4 per axis); if every one of them has
//     a zero wide maximum, every addend of the chunk is +-0 -- what the per-voxel pre-test of sv_fused_kernel would find 512
//     times -- and the running sums pass through unchanged.  Chunks outside [lo, hi) are empty by the AABB interval.
//   * Empty chunks while no column has met density are all ones: a warp stores the whole 2 KiB brick with four 128-bit
//     stores per lane (st.global.cs.v4), 16 bricks per CTA iteration, no per-voxel work at all.  That is ~85 % of the volume.
//   * Empty chunks after density has been met need exp((-dist*acc)/(z+1)) per voxel with the sums as they stand: no
//     shared memory, no barrier.
//   * Only the remaining chunks (inside the cloud's occupied bricks) run the gather / serial sum / exp sequence.
// Per column the shader's operations happen in the shader's order, so the bits equal the other build modes'.
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void store_ones_brick(float* brick, unsigned lane)
{
    float4* p = reinterpret_cast<float4*>(brick) + lane;
    const float4 one = make_float4(1.0f, 1.0f, 1.0f, 1.0f);
    __stcs(p, one);
    __stcs(p + 32, one);
    __stcs(p + 64, one);
    __stcs(p + 96, one);
}

template <int CL>
__global__ void __launch_bounds__(512, 3) sv_fused_brick_kernel(const __grid_constant__ SVConst c)
{
    constexpr int OL = CRV_LAYOUT_BRICK;
    constexpr int MAX_CHUNK_WORDS = SV_MAX_FAST_N / 8 / 32;
    __shared__ int s_lo, s_hi;
    __shared__ int s_search[6][64];
    __shared__ int s_zin[64], s_zend[64];
    constexpr int G = 4;                // chunks per gather / serial-sum / exp iteration (G * 8 slices between two barriers)
    __shared__ float s_add[8 * G][64];  // G chunks' addends, then the running sums after each slice
    __shared__ float s_acc[64];         // running sum of every column after the previous chunk
    __shared__ uint32_t s_full[MAX_CHUNK_WORDS];      // bit k: chunk k may add density
    const int N = c.n;
    const int t = threadIdx.x;
    const unsigned lane = t & 31u, warp = t >> 5;
    const int bx = centre_out(blockIdx.x, gridDim.x), by = centre_out(blockIdx.y, gridDim.y);
    const int nchunks = (N + 7) >> 3;

    // ---- which chunks of this bundle can meet density at all?  Its samples' brick coordinates are affine in z: per axis the
    //      z range in which the bundle's 8 x 8 columns can touch the box of occupied bricks (crv_volume::occ_ext) is closed
    //      form; outside the intersection of the three ranges every chunk is empty.  Most bundles have no such range. ----
    __shared__ int s_clo, s_chi;
    if (t == 0) { s_lo = N; s_hi = 0; }
    if (t < MAX_CHUNK_WORDS) s_full[t] = 0u;
    const bool pre = c.brick_absmax_wide != nullptr;
    if (t == 0) {
        float zlo = 0.0f, zhi = (float)N;
        if (pre) {
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                const int nb = a == 0 ? c.cloud.nbx : a == 1 ? c.cloud.nby : c.cloud.nbz;
                const int e0 = __ldg(c.occ_ext + a), e1 = __ldg(c.occ_ext + 3 + a);
                if (e1 < e0) { zlo = 1.0f; zhi = 0.0f; break; }                   // nothing occupied anywhere
                const float base = c.pre_c[a] + (float)(bx * 8) * c.pre_r[a] + (float)(by * 8) * c.pre_u[a];
                const float er = 7.0f * c.pre_r[a], eu = 7.0f * c.pre_u[a];
                const float mn = base + fminf(er, 0.0f) + fminf(eu, 0.0f) - 0.25f, mx = base + fmaxf(er, 0.0f) + fmaxf(eu, 0.0f) + 0.25f;
                // coordinates beyond the volume clamp to its boundary bricks: an occupied boundary brick extends to infinity
                const float lo_b = e0 == 0 ? -3.0e38f : (float)e0, hi_b = e1 == nb - 1 ? 3.0e38f : (float)(e1 + 1);
                const float l = c.pre_l[a];
                // need  mx + z*l >= lo_b  and  mn + z*l < hi_b
                if (l > 0.0f) {
                    zlo = fmaxf(zlo, (lo_b - mx) / l);
                    zhi = fminf(zhi, (hi_b - mn) / l);
                } else if (l < 0.0f) {
                    zlo = fmaxf(zlo, (hi_b - mn) / l);
                    zhi = fminf(zhi, (lo_b - mx) / l);
                } else if (mx < lo_b || mn >= hi_b) {
                    zlo = 1.0f; zhi = 0.0f;
                }
            }
        }
        // one chunk of slack on either side covers the rounding of the divisions above
        const bool none = !(zlo <= zhi);
        s_clo = none ? N : max((int)floorf(fmaxf(zlo, 0.0f)) - 8, 0) & ~7;
        s_chi = none ? 0 : min((int)ceilf(fminf(zhi, (float)N)) + 8, N);
    }
    __syncthreads();
    // ---- classification of the chunks in that range: one warp per chunk, lanes over the cloud bricks its samples can touch ----
    for (int k = (s_clo >> 3) + (int)warp; k * 8 < s_chi; k += 16) {
        bool occupied = !pre;
        if (pre) {
            // approximate brick coordinate of the chunk's corner voxel and its extent over the chunk (affine in x, y, z)
            int b0i[3], cnt[3];
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                const float v0 = c.pre_c[a] + (float)(bx * 8) * c.pre_r[a] + (float)(by * 8) * c.pre_u[a] + (float)(k * 8) * c.pre_l[a];
                const float er = 7.0f * c.pre_r[a], eu = 7.0f * c.pre_u[a], el = 7.0f * c.pre_l[a];
                const float vmin = v0 + fminf(er, 0.0f) + fminf(eu, 0.0f) + fminf(el, 0.0f) - 0.01f;
                const float vmax = v0 + fmaxf(er, 0.0f) + fmaxf(eu, 0.0f) + fmaxf(el, 0.0f) + 0.01f;
                const int nb = a == 0 ? c.cloud.nbx : a == 1 ? c.cloud.nby : c.cloud.nbz;
                const int i0 = min(max(__float2int_rd(vmin), 0), nb - 1), i1 = min(max(__float2int_rd(vmax), 0), nb - 1);
                b0i[a] = i0;
                cnt[a] = i1 - i0 + 1;
            }
            // lanes as a 4 x 4 x 2 block of bricks, twice along z (the extent of a chunk is < 3 bricks per axis: at most 4 touched)
            const int ix = (int)(lane & 3u), iy = (int)((lane >> 2) & 3u);
            bool mine = false;
            if (ix < cnt[0] && iy < cnt[1]) {
                const unsigned row = (unsigned)(b0i[1] + iy) * c.cloud.nbx + (unsigned)(b0i[0] + ix);
                const unsigned slab = (unsigned)c.cloud.nby * c.cloud.nbx;
                for (int iz = (int)(lane >> 4); iz < cnt[2]; iz += 2)
                    mine = mine || !(__ldg(c.brick_absmax_wide + ((unsigned)(b0i[2] + iz) * slab + row)) == 0.0f);
            }
            occupied = __any_sync(0xffffffffu, mine) || cnt[0] > 4 || cnt[1] > 4;
        }
        if (occupied && lane == 0) atomicOr(&s_full[k >> 5], 1u << (k & 31));
    }
    __syncthreads();
    bool any_full = false;
#pragma unroll
    for (int w = 0; w < MAX_CHUNK_WORDS; ++w) any_full = any_full || s_full[w] != 0u;

    // ---- [z_in, z_end) of the bundle's 64 columns -- only a bundle with such chunks needs them: 6 searches per column spread
    //      over 384 threads; chunks outside every column's interval are empty after all ----
    if (any_full) {
        if (t < 384) {
            const int col = t & 63, search = t >> 6;
            const int cx = bx * 8 + (col & 7), cy = by * 8 + (col >> 3);
            const int ax = search >> 1;
            const float a_ax = (float)cx * c.right[ax] + (float)cy * c.up[ax];
            s_search[search][col] = range_seek(c, a_ax, ax, (search & 1) != 0);
        }
        __syncthreads();
        if (t < 64) {
            const int cx = bx * 8 + (t & 7), cy = by * 8 + (t >> 3);
            int z_in = max(max(s_search[0][t], s_search[2][t]), s_search[4][t]);
            int z_end = min(min(s_search[1][t], s_search[3][t]), s_search[5][t]);
            if (z_end < z_in) z_end = z_in;
            if (cx >= N || cy >= N) z_in = z_end = 0;
            s_zin[t] = z_in;
            s_zend[t] = z_end;
            if (z_in < z_end) {
                atomicMin(&s_lo, z_in);
                atomicMax(&s_hi, z_end);
            }
        }
        __syncthreads();
        const int hi = s_hi, lo = hi > 0 ? (s_lo & ~7) : N;
        if (t < MAX_CHUNK_WORDS) {              // keep the chunks [lo / 8, ceil(hi / 8))
            const int first = lo >> 3, last = (hi + 7) >> 3, w0 = t * 32;
            uint32_t keep = 0u;
            for (int bit = 0; bit < 32; ++bit) if (w0 + bit >= first && w0 + bit < last) keep |= 1u << bit;
            s_full[t] &= keep;
        }
    }
    if (t < 64) {
        s_acc[t] = 0.0f;
        if (!any_full) { s_zin[t] = 0; s_zend[t] = 0; }
    }
    __syncthreads();

    const int col = t & 63, lz = t >> 6;
    const int x = bx * 8 + (t & 7), y = by * 8 + ((t >> 3) & 7);
    const bool inb = x < N && y < N;
    const bool whole_bundle = bx * 8 + 8 <= N && by * 8 + 8 <= N;
    const int z_in = s_zin[col], z_end = s_zend[col];
    const float fx = (float)x, fy = (float)y;
    const float a[3] = {fx * c.right[0] + fy * c.up[0], fx * c.right[1] + fy * c.up[1], fx * c.right[2] + fy * c.up[2]};
    const float pa[3] = {__fmaf_rn(fy, c.pre_u[0], __fmaf_rn(fx, c.pre_r[0], c.pre_c[0])),
                         __fmaf_rn(fy, c.pre_u[1], __fmaf_rn(fx, c.pre_r[1], c.pre_c[1])),
                         __fmaf_rn(fy, c.pre_u[2], __fmaf_rn(fx, c.pre_r[2], c.pre_c[2]))};
    auto transmittance = [&](int z, float acc) -> float {
        float tr = 1.0f;
        if (__any_sync(0xffffffffu, acc != 0.0f) && z < N) tr = expf(div_rn_by((-__ldg(c.dist_z + z)) * acc, (float)(z + 1), c_rcp_z[z]));
        return tr;
    };
    // element offset of this thread's voxel in chunk 0, and of the bundle's brick 0; a chunk further is one brick row of the volume
    const unsigned chunk_step = VolAcc<OL>::off_z(c.out, 8);
    float* const vox0 = const_cast<float*>(c.out.data) + (VolAcc<OL>::off_x(c.out, x) + VolAcc<OL>::off_y(c.out, y) + VolAcc<OL>::off_z(c.out, lz));
    float* const brick0 = const_cast<float*>(c.out.data) + (VolAcc<OL>::off_x(c.out, bx * 8) + VolAcc<OL>::off_y(c.out, by * 8));

    bool any_acc = false;             // CTA-uniform: some column of the bundle has a non-zero running sum
    int k = 0;                        // current chunk
    while (k < nchunks) {
        // length of the run of empty chunks starting at k (CTA-uniform)
        int run = 0;
        {
            int j = k;
            while (j < nchunks) {
                const uint32_t w = s_full[j >> 5] >> (j & 31);
                if (w != 0u) { j += __ffs(w) - 1; break; }
                j = (j | 31) + 1;
            }
            run = min(j, nchunks) - k;
        }
        if (run > 0) {
            if (!any_acc && whole_bundle) {
                // all ones: whole bricks, one warp each
                const int whole = (k + run) * 8 <= N ? run : run - 1;        // the last chunk may be cut by N
                for (int j = (int)warp; j < whole; j += 16) store_ones_brick(brick0 + (size_t)(k + j) * chunk_step, lane);
                if (whole < run) {
                    const int z = (k + whole) * 8 + lz;
                    if (z < N) __stcs(vox0 + (size_t)(k + whole) * chunk_step, 1.0f);
                }
            } else if (!any_acc) {
                for (int j = 0; j < run; ++j) {
                    const int z = (k + j) * 8 + lz;
                    if (inb && z < N) __stcs(vox0 + (size_t)(k + j) * chunk_step, 1.0f);
                }
            } else {
                // the sums stand: exp + store only
                const float acc = s_acc[col];
                const bool any = __any_sync(0xffffffffu, acc != 0.0f);
                for (int j = 0; j < run; ++j) {
                    const int z = (k + j) * 8 + lz;
                    float tr = 1.0f;
                    if (any && z < N) tr = expf(div_rn_by((-__ldg(c.dist_z + z)) * acc, (float)(z + 1), c_rcp_z[z]));
                    if (inb && z < N) __stcs(vox0 + (size_t)(k + j) * chunk_step, tr);
                }
            }
            k += run;
            continue;
        }
        // ---- chunk k may add density: gather, serial sums, exp for chunks k .. k + G - 1 between one pair of barriers.  The G
        //      brick look-ups of a thread are issued together, then the samples of those that are not proven empty: a bundle
        //      inside the cloud is a chain of such iterations, and their latency is what the build waits for at the end ----
        const int z0 = k * 8;
        float addv[G];
        bool need[G];
        float wide[G];
#pragma unroll
        for (int u = 0; u < G; ++u) {
            const int z = z0 + 8 * u + lz;
            need[u] = z >= z_in && z < z_end;
            wide[u] = 1.0f;
            if (need[u] && pre) {
                const float fz = (float)z;
                const int bi = max(__float_as_int(__fadd_rd(__fmaf_rn(fz, c.pre_l[0], pa[0]), 8388608.0f)) - 0x4B000000, 0);
                const int bj = max(__float_as_int(__fadd_rd(__fmaf_rn(fz, c.pre_l[1], pa[1]), 8388608.0f)) - 0x4B000000, 0);
                const int bk = max(__float_as_int(__fadd_rd(__fmaf_rn(fz, c.pre_l[2], pa[2]), 8388608.0f)) - 0x4B000000, 0);
                const unsigned b = ((unsigned)min(bk, c.cloud.nbz - 1) * c.cloud.nby + (unsigned)min(bj, c.cloud.nby - 1)) * c.cloud.nbx +
                                   (unsigned)min(bi, c.cloud.nbx - 1);
                wide[u] = __ldg(c.brick_absmax_wide + b);
            }
        }
#pragma unroll
        for (int u = 0; u < G; ++u) {
            addv[u] = 0.0f;       // + 0 outside the interval / in empty bricks leaves the sum unchanged (it is never -0)
            if (need[u] && !(wide[u] == 0.0f))
                addv[u] = div_rn_by(column_sample<CL, true>(c, a, z0 + 8 * u + lz) * c.density_scaling, c.base_scaling, c.rbase);
            s_add[8 * u + lz][col] = addv[u];
        }
        __syncthreads();
        bool nz = false;
        if (t < 64) {
            float sum = s_acc[t];
#pragma unroll
            for (int u = 0; u < 8 * G; ++u) {
                sum = sum + s_add[u][t];          // accumulatedValue after slice z0 + u, the shader's order (:89)
                s_add[u][t] = sum;
            }
            s_acc[t] = sum;
            nz = sum != 0.0f;
        }
        any_acc = __syncthreads_or(nz) != 0 || any_acc;
#pragma unroll
        for (int u = 0; u < G; ++u) {
            const int z = z0 + 8 * u + lz;
            const float tr = transmittance(z, s_add[8 * u + lz][col]);
            if (inb && z < N) __stcs(vox0 + (size_t)(k + u) * chunk_step, tr);
        }
        k += G;
    }
}

// ------------------------------------------------------------------------------------------------------
// Fused build, brick output, warp-owned rows (SV_MODE_FUSED_ROWS): the chunk classification of sv_fused_brick_kernel with a
// walk that needs no CTA barrier.  One CTA of 8 warps per bundle; warp w owns the bundle's row y = w (8 columns) and walks it
// from the bundle's first chunk that may add density to z = N on its own: lane = x + 8 * zz holds the voxel (x, y, 4 h + zz) of
// half-chunk h, so a warp's store covers four whole 32-byte sectors of the brick, and the running sums of its 8 columns live in
// registers (the 4 lanes of a column hold the same value).  The shader's serial order is kept by fetching the column's four
// addends with shuffles and adding them one after the other in every lane of the column.  The chunks in front of the first
// occupied one are ones for every column of the bundle: whole bricks, one warp each, before the walk.
// ------------------------------------------------------------------------------------------------------
#ifndef CRV_SV_ROWS_GROUP
#define CRV_SV_ROWS_GROUP 4
#endif
#ifndef CRV_SV_ROWS_BATCH
#define CRV_SV_ROWS_BATCH 2        // samples whose texel loads are issued together (0: sample by sample, with the second emptiness proof)
#endif
#ifndef CRV_SV_ROWS_CTAS
#define CRV_SV_ROWS_CTAS 4
#endif
template <int CL>
__global__ void __launch_bounds__(256, CRV_SV_ROWS_CTAS) sv_fused_rows_kernel(const __grid_constant__ SVConst c)
{
    constexpr int OL = CRV_LAYOUT_BRICK;
    constexpr int NW = 8;               // warps per CTA = rows of the bundle
    constexpr int MAX_CHUNK_WORDS = SV_MAX_FAST_N / 8 / 32;
    __shared__ int s_lo, s_hi;
    __shared__ int s_search[6][64];
    __shared__ int s_zin[64], s_zend[64];
    __shared__ uint32_t s_full[MAX_CHUNK_WORDS];      // bit k: chunk k may add density
    __shared__ int s_clo, s_chi;
    const int N = c.n;
    const int t = threadIdx.x;
    const unsigned lane = t & 31u, warp = t >> 5;
    const int bx = centre_out(blockIdx.x, gridDim.x), by = centre_out(blockIdx.y, gridDim.y);
    const int nchunks = (N + 7) >> 3;

    // ---- the z range in which the bundle can touch the box of occupied bricks (as in sv_fused_brick_kernel) ----
    if (t == 0) { s_lo = N; s_hi = 0; }
    if (t < MAX_CHUNK_WORDS) s_full[t] = 0u;
    const bool pre = c.brick_absmax_wide != nullptr;
    if (t == 0) {
        float zlo = 0.0f, zhi = (float)N;
        if (pre) {
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                const int nb = a == 0 ? c.cloud.nbx : a == 1 ? c.cloud.nby : c.cloud.nbz;
                const int e0 = __ldg(c.occ_ext + a), e1 = __ldg(c.occ_ext + 3 + a);
                if (e1 < e0) { zlo = 1.0f; zhi = 0.0f; break; }                   // nothing occupied anywhere
                const float base = c.pre_c[a] + (float)(bx * 8) * c.pre_r[a] + (float)(by * 8) * c.pre_u[a];
                const float er = 7.0f * c.pre_r[a], eu = 7.0f * c.pre_u[a];
                const float mn = base + fminf(er, 0.0f) + fminf(eu, 0.0f) - 0.25f, mx = base + fmaxf(er, 0.0f) + fmaxf(eu, 0.0f) + 0.25f;
                const float lo_b = e0 == 0 ? -3.0e38f : (float)e0, hi_b = e1 == nb - 1 ? 3.0e38f : (float)(e1 + 1);
                const float l = c.pre_l[a];
                if (l > 0.0f) {
                    zlo = fmaxf(zlo, (lo_b - mx) / l);
                    zhi = fminf(zhi, (hi_b - mn) / l);
                } else if (l < 0.0f) {
                    zlo = fmaxf(zlo, (hi_b - mn) / l);
                    zhi = fminf(zhi, (lo_b - mx) / l);
                } else if (mx < lo_b || mn >= hi_b) {
                    zlo = 1.0f; zhi = 0.0f;
                }
            }
        }
        const bool none = !(zlo <= zhi);
        s_clo = none ? N : max((int)floorf(fmaxf(zlo, 0.0f)) - 8, 0) & ~7;
        s_chi = none ? 0 : min((int)ceilf(fminf(zhi, (float)N)) + 8, N);
    }
    __syncthreads();
    // ---- classification of the chunks in that range: one warp per chunk ----
    for (int k = (s_clo >> 3) + (int)warp; k * 8 < s_chi; k += NW) {
        bool occupied = !pre;
        if (pre) {
            int b0i[3], cnt[3];
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                const float v0 = c.pre_c[a] + (float)(bx * 8) * c.pre_r[a] + (float)(by * 8) * c.pre_u[a] + (float)(k * 8) * c.pre_l[a];
                const float er = 7.0f * c.pre_r[a], eu = 7.0f * c.pre_u[a], el = 7.0f * c.pre_l[a];
                const float vmin = v0 + fminf(er, 0.0f) + fminf(eu, 0.0f) + fminf(el, 0.0f) - 0.01f;
                const float vmax = v0 + fmaxf(er, 0.0f) + fmaxf(eu, 0.0f) + fmaxf(el, 0.0f) + 0.01f;
                const int nb = a == 0 ? c.cloud.nbx : a == 1 ? c.cloud.nby : c.cloud.nbz;
                const int i0 = min(max(__float2int_rd(vmin), 0), nb - 1), i1 = min(max(__float2int_rd(vmax), 0), nb - 1);
                b0i[a] = i0;
                cnt[a] = i1 - i0 + 1;
            }
            const int ix = (int)(lane & 3u), iy = (int)((lane >> 2) & 3u);
            bool mine = false;
            if (ix < cnt[0] && iy < cnt[1]) {
                const unsigned row = (unsigned)(b0i[1] + iy) * c.cloud.nbx + (unsigned)(b0i[0] + ix);
                const unsigned slab = (unsigned)c.cloud.nby * c.cloud.nbx;
                for (int iz = (int)(lane >> 4); iz < cnt[2]; iz += 2)
                    mine = mine || !(__ldg(c.brick_absmax_wide + ((unsigned)(b0i[2] + iz) * slab + row)) == 0.0f);
            }
            occupied = __any_sync(0xffffffffu, mine) || cnt[0] > 4 || cnt[1] > 4;
        }
        if (occupied && lane == 0) atomicOr(&s_full[k >> 5], 1u << (k & 31));
    }
    __syncthreads();
    bool any_full = false;
#pragma unroll
    for (int w = 0; w < MAX_CHUNK_WORDS; ++w) any_full = any_full || s_full[w] != 0u;

    // ---- [z_in, z_end) of the bundle's 64 columns (6 searches per column) ----
    if (any_full) {
        for (int i = t; i < 384; i += 256) {
            const int col = i & 63, search = i >> 6;
            const int cx = bx * 8 + (col & 7), cy = by * 8 + (col >> 3);
            const int ax = search >> 1;
            const float a_ax = (float)cx * c.right[ax] + (float)cy * c.up[ax];
            s_search[search][col] = range_seek(c, a_ax, ax, (search & 1) != 0);
        }
        __syncthreads();
        if (t < 64) {
            const int cx = bx * 8 + (t & 7), cy = by * 8 + (t >> 3);
            int z_in = max(max(s_search[0][t], s_search[2][t]), s_search[4][t]);
            int z_end = min(min(s_search[1][t], s_search[3][t]), s_search[5][t]);
            if (z_end < z_in) z_end = z_in;
            if (cx >= N || cy >= N) z_in = z_end = 0;
            s_zin[t] = z_in;
            s_zend[t] = z_end;
            if (z_in < z_end) {
                atomicMin(&s_lo, z_in);
                atomicMax(&s_hi, z_end);
            }
        }
        __syncthreads();
        const int hi = s_hi, lo = hi > 0 ? (s_lo & ~7) : N;
        if (t < MAX_CHUNK_WORDS) {              // keep the chunks [lo / 8, ceil(hi / 8))
            const int first = lo >> 3, last = (hi + 7) >> 3, w0 = t * 32;
            uint32_t keep = 0u;
            for (int bit = 0; bit < 32; ++bit) if (w0 + bit >= first && w0 + bit < last) keep |= 1u << bit;
            s_full[t] &= keep;
        }
    } else if (t < 64) {
        s_zin[t] = 0;
        s_zend[t] = 0;
    }
    __syncthreads();

    // ---- the chunks in front of the first one that may add density: ones in every column ----
    int first_full = nchunks;
#pragma unroll
    for (int w = MAX_CHUNK_WORDS - 1; w >= 0; --w)
        if (s_full[w] != 0u) first_full = w * 32 + __ffs(s_full[w]) - 1;
    first_full = min(first_full, nchunks);
    const int lx = (int)(lane & 7u), zz = (int)(lane >> 3);
    const int x = bx * 8 + lx, y = by * 8 + (int)warp;
    const bool inb = x < N && y < N;
    const bool whole_bundle = bx * 8 + 8 <= N && by * 8 + 8 <= N;
    const unsigned chunk_step = VolAcc<OL>::off_z(c.out, 8);
    // this lane's voxel of half-chunk 0 of chunk 0; half-chunk 1 lies 4 slices (4 * 64 elements of the brick) further
    float* const vox0 = const_cast<float*>(c.out.data) + (VolAcc<OL>::off_x(c.out, x) + VolAcc<OL>::off_y(c.out, y) + VolAcc<OL>::off_z(c.out, zz));
    float* const brick0 = const_cast<float*>(c.out.data) + (VolAcc<OL>::off_x(c.out, bx * 8) + VolAcc<OL>::off_y(c.out, by * 8));
    constexpr unsigned HALF = 4u * 64u;
    if (whole_bundle) {
        const int whole = first_full * 8 <= N ? first_full : first_full - 1;            // the last chunk may be cut by N
        for (int j = (int)warp; j < whole; j += NW) store_ones_brick(brick0 + (size_t)j * chunk_step, lane);
        if (whole < first_full) {
#pragma unroll
            for (int h = 0; h < 2; ++h)
                if (whole * 8 + 4 * h + zz < N) __stcs(vox0 + (size_t)whole * chunk_step + h * HALF, 1.0f);
        }
    } else {
        for (int j = 0; j < first_full; ++j) {
#pragma unroll
            for (int h = 0; h < 2; ++h)
                if (inb && j * 8 + 4 * h + zz < N) __stcs(vox0 + (size_t)j * chunk_step + h * HALF, 1.0f);
        }
    }
    if (first_full >= nchunks) return;

    // ---- the walk of this warp's row ----
    const int col = lx + 8 * (int)warp;
    const int z_in = s_zin[col], z_end = s_zend[col];
    const float fx = (float)x, fy = (float)y;
    const float a[3] = {fx * c.right[0] + fy * c.up[0], fx * c.right[1] + fy * c.up[1], fx * c.right[2] + fy * c.up[2]};
    const float pa[3] = {__fmaf_rn(fy, c.pre_u[0], __fmaf_rn(fx, c.pre_r[0], c.pre_c[0])),
                         __fmaf_rn(fy, c.pre_u[1], __fmaf_rn(fx, c.pre_r[1], c.pre_c[1])),
                         __fmaf_rn(fy, c.pre_u[2], __fmaf_rn(fx, c.pre_r[2], c.pre_c[2]))};
    float acc = 0.0f;                 // accumulatedValue of this lane's column after the previous chunk
    bool row_any = false;             // warp-uniform: some column of the row has a non-zero sum
    constexpr int G = CRV_SV_ROWS_GROUP;          // chunks that may add density handled per iteration: 2 G samples per lane in flight
    // first proof of emptiness for the 2 G samples of a group starting at slice zb0: approximate brick coordinate against the
    // wide table (one L2 load per sample, all issued together)
    auto lookup = [&](int zb0, bool (&need)[2 * G], float (&wide)[2 * G]) {
#pragma unroll
        for (int u = 0; u < 2 * G; ++u) {
            const int z = zb0 + 4 * u;
            need[u] = z >= z_in && z < z_end;
            wide[u] = 1.0f;
            if (need[u] && pre) {
                const float fz = (float)z;
                const int bi = max(__float_as_int(__fadd_rd(__fmaf_rn(fz, c.pre_l[0], pa[0]), 8388608.0f)) - 0x4B000000, 0);
                const int bj = max(__float_as_int(__fadd_rd(__fmaf_rn(fz, c.pre_l[1], pa[1]), 8388608.0f)) - 0x4B000000, 0);
                const int bk = max(__float_as_int(__fadd_rd(__fmaf_rn(fz, c.pre_l[2], pa[2]), 8388608.0f)) - 0x4B000000, 0);
                const unsigned b = ((unsigned)min(bk, c.cloud.nbz - 1) * c.cloud.nby + (unsigned)min(bj, c.cloud.nby - 1)) * c.cloud.nbx +
                                   (unsigned)min(bi, c.cloud.nbx - 1);
                wide[u] = __ldg(c.brick_absmax_wide + b);
            }
        }
    };
    int k = first_full;
    while (k < nchunks) {
        const bool full = ((s_full[k >> 5] >> (k & 31)) & 1u) != 0u;
        float* const dst = vox0 + (size_t)k * chunk_step;
        const int zb = k * 8 + zz;                                   // this lane's slice in half-chunk 0
        if (!full) {
            if (!row_any) {
#pragma unroll
                for (int h = 0; h < 2; ++h)
                    if (inb && zb + 4 * h < N) __stcs(dst + h * HALF, 1.0f);
            } else {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int z = zb + 4 * h;
                    if (z < N) {
                        const float tr = expf(div_rn_by((-__ldg(c.dist_z + z)) * acc, (float)(z + 1), c_rcp_z[z]));
                        if (inb) __stcs(dst + h * HALF, tr);
                    }
                }
            }
            ++k;
            continue;
        }
        // chunks k .. k + G - 1 (an empty one among them just adds zeros): all look-ups, then the samples, then the serial sums
        float addv[2 * G], wide[2 * G];
        bool need[2 * G];
        lookup(zb, need, wide);
#if CRV_SV_ROWS_BATCH > 0
        if constexpr (CL != CRV_LAYOUT_TEX3D) {
            // B samples at a time in three steps -- footprints, all 8 B texel loads, lerps: a lane waits for DRAM once per batch.
            // The second proof of emptiness (exact footprint against the dilated brick maxima) is left out: the voxels it
            // would spare after the wide table has passed them read eight zeros and add +-0, the same bits.
            using A = VolAcc<CL>;
            constexpr int B = CRV_SV_ROWS_BATCH;
            static_assert((2 * G) % B == 0, "batches tile the group");
#pragma unroll
            for (int h = 0; h < 2 * G; h += B) {
                unsigned x0[B], x1[B], r00[B], r10[B], r01[B], r11[B];
                float wa[B], wb[B], wc[B];
                bool live[B];
#pragma unroll
                for (int v = 0; v < B; ++v) {
                    const int u = h + v;
                    live[v] = need[u] && !(wide[u] == 0.0f);
                    x0[v] = x1[v] = r00[v] = r10[v] = r01[v] = r11[v] = 0u;
                    wa[v] = wb[v] = wc[v] = 0.0f;
                    if (live[v]) {
                        const float fz = (float)(zb + 4 * u);
                        const float px = c.b0[0] + ((a[0] + fz * c.light[0]) * c.voxel_size);
                        const float py = c.b0[1] + ((a[1] + fz * c.light[1]) * c.voxel_size);
                        const float pz = c.b0[2] + ((a[2] + fz * c.light[2]) * c.voxel_size);
                        const float sx = div_rn_by(px - c.cb0[0], c.cext[0], c.rcext[0]);
                        const float sy = div_rn_by(py - c.cb0[1], c.cext[1], c.rcext[1]);
                        const float sz = div_rn_by(pz - c.cb0[2], c.cext[2], c.rcext[2]);
                        const Footprint f = make_footprint<true>(c.cloud, sx, sy, sz);
                        x0[v] = A::off_x(c.cloud, f.i0);
                        x1[v] = A::off_x(c.cloud, f.i1);
                        const unsigned y0 = A::off_y(c.cloud, f.j0), y1 = A::off_y(c.cloud, f.j1);
                        const unsigned q0 = A::off_z(c.cloud, f.k0), q1 = A::off_z(c.cloud, f.k1);
                        r00[v] = y0 + q0; r10[v] = y1 + q0; r01[v] = y0 + q1; r11[v] = y1 + q1;
                        wa[v] = f.a; wb[v] = f.b; wc[v] = f.c;
                    }
                }
                float tx[B][8];
#pragma unroll
                for (int v = 0; v < B; ++v) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) tx[v][j] = 0.0f;
                    if (live[v]) {
                        const float* d = c.cloud.data;
                        tx[v][0] = __ldg(d + (r00[v] + x0[v])); tx[v][1] = __ldg(d + (r00[v] + x1[v]));
                        tx[v][2] = __ldg(d + (r10[v] + x0[v])); tx[v][3] = __ldg(d + (r10[v] + x1[v]));
                        tx[v][4] = __ldg(d + (r01[v] + x0[v])); tx[v][5] = __ldg(d + (r01[v] + x1[v]));
                        tx[v][6] = __ldg(d + (r11[v] + x0[v])); tx[v][7] = __ldg(d + (r11[v] + x1[v]));
                    }
                }
#pragma unroll
                for (int v = 0; v < B; ++v) {
                    const int u = h + v;
                    addv[u] = 0.0f;
                    if (live[v]) {
                        Footprint f;
                        f.a = wa[v]; f.b = wb[v]; f.c = wc[v];
                        const float val = trilerp(f, tx[v][0], tx[v][1], tx[v][2], tx[v][3], tx[v][4], tx[v][5], tx[v][6], tx[v][7]);
                        addv[u] = div_rn_by(val * c.density_scaling, c.base_scaling, c.rbase);
                    }
                }
            }
        } else
#endif
        {
#pragma unroll
            for (int u = 0; u < 2 * G; ++u) {
                addv[u] = 0.0f;       // + 0 outside the interval / in empty bricks leaves the sum unchanged (it is never -0)
                if (need[u] && !(wide[u] == 0.0f))
                    addv[u] = div_rn_by(column_sample<CL, true>(c, a, zb + 4 * u) * c.density_scaling, c.base_scaling, c.rbase);
            }
        }
        // accumulatedValue after each slice, in the shader's order (:89): the addend of slice 4 u + j of this lane's column sits
        // in lane lx + 8 j
        float mine[2 * G];
#pragma unroll
        for (int u = 0; u < 2 * G; ++u) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                acc = acc + __shfl_sync(0xffffffffu, addv[u], lx + 8 * j);
                if (j == zz) mine[u] = acc;
            }
        }
        row_any = row_any || __any_sync(0xffffffffu, acc != 0.0f);
#pragma unroll
        for (int u = 0; u < 2 * G; ++u) {
            const int z = zb + 4 * u;
            float tr = 1.0f;
            if (row_any && z < N) tr = expf(div_rn_by((-__ldg(c.dist_z + z)) * mine[u], (float)(z + 1), c_rcp_z[z]));
            if (inb && z < N) __stcs(dst + (size_t)(u >> 1) * chunk_step + (u & 1) * HALF, tr);
        }
        k += G;
    }
}

template <int CL, int OL>
void launch_sv(const SVConst& c, int mode, int* launches, cudaStream_t s)
{
    if (mode == SV_MODE_FUSED && OL == CRV_LAYOUT_BRICK) {
        sv_fused_brick_kernel<CL><<<dim3((c.n + 7) / 8, (c.n + 7) / 8, 1), 512, 0, s>>>(c);
        *launches = 1;
        return;
    }
    if (mode == SV_MODE_FUSED_ROWS && OL == CRV_LAYOUT_BRICK) {
        sv_fused_rows_kernel<CL><<<dim3((c.n + 7) / 8, (c.n + 7) / 8, 1), 256, 0, s>>>(c);
        *launches = 1;
        return;
    }
    if (mode == SV_MODE_FUSED || mode == SV_MODE_FUSED_V1 || mode == SV_MODE_FUSED_ROWS) {
        sv_fused_kernel<CL, OL><<<dim3((c.n + 7) / 8, (c.n + 7) / 8, 1), 512, 0, s>>>(c);
        *launches = 1;
        return;
    }
    const dim3 block(32, SV_ROWS, 1);
    const dim3 grid((c.n + 31) / 32, (c.n + SV_ROWS - 1) / SV_ROWS, 1);
    if (mode == SV_MODE_TWO_PASS) {
        sv_gather_kernel<CL, OL><<<dim3((c.n + 7) / 8, (c.n + 7) / 8, 1), 512, 0, s>>>(c);
        sv_scan_kernel<OL><<<grid, block, 0, s>>>(c);
        *launches = 2;
    } else if (mode == SV_MODE_SINGLE_PASS) {
        sv_sweep_fast_kernel<CL, OL, 4><<<grid, block, 0, s>>>(c);
        *launches = 1;
    } else {
        sv_sweep_kernel<CL, OL, 4><<<grid, block, 0, s>>>(c);
        *launches = 1;
    }
}

template <int CL>
void launch_sv_out(const SVConst& c, int out_layout, int mode, int* launches, cudaStream_t s)
{
    if (out_layout == CRV_LAYOUT_BRICK) launch_sv<CL, CRV_LAYOUT_BRICK>(c, mode, launches, s);
    else if (out_layout == CRV_LAYOUT_TEX3D) launch_sv<CL, CRV_LAYOUT_TEX3D>(c, mode, launches, s);
    else launch_sv<CL, CRV_LAYOUT_LINEAR>(c, mode, launches, s);
}

}  // namespace

bool crv_sv_fast_path_ok(const SVConst& c)
{
    auto fin = [](float v) { return std::isfinite(v); };
    auto div_ok = [&](float d) { return fin(d) && std::fabs(d) >= 1e-12f && std::fabs(d) <= 1e12f; };
    bool ok = fin(c.voxel_size) && c.voxel_size > 0.0f && c.voxel_size < 1e12f && fin(c.density_scaling) && div_ok(c.base_scaling) &&
              c.n > 0 && c.n <= SV_MAX_FAST_N && c.zrange != nullptr && c.dist_z != nullptr && c.brick_absmax != nullptr;
    for (int i = 0; i < 3; ++i)
        ok = ok && fin(c.b0[i]) && fin(c.right[i]) && fin(c.up[i]) && fin(c.light[i]) && fin(c.cb0[i]) && fin(c.cb1[i]) && div_ok(c.cext[i]) &&
             std::fabs(c.b0[i]) < 1e12f && std::fabs(c.cb0[i]) < 1e12f;
    return ok;
}

int crv_sv_init_tables(crv_ctx* ctx)
{
    if (ctx->sv_tables_ready) return CRV_OK;
    static float tab[SV_MAX_FAST_N];
    for (int z = 0; z < SV_MAX_FAST_N; ++z) tab[z] = 1.0f / (float)(z + 1);      // host IEEE division
    CRV_CUDA(cudaMemcpyToSymbol(c_rcp_z, tab, sizeof(tab)));
    ctx->sv_tables_ready = true;
    return CRV_OK;
}

int crv_launch_sv_dist(crv_ctx* ctx, float* dist_z, int n, float voxel_size, cudaStream_t s)
{
    sv_dist_kernel<<<1, 32, 0, s>>>(dist_z, n, voxel_size);
    ctx->launches++;
    CRV_CUDA(cudaGetLastError());
    return CRV_OK;
}

int crv_launch_sv_sweep(crv_ctx* ctx, const SVConst& c, int cloud_layout, int out_layout, int mode, cudaStream_t s)
{
    int launches = 0;
    if (cloud_layout == CRV_LAYOUT_BRICK) launch_sv_out<CRV_LAYOUT_BRICK>(c, out_layout, mode, &launches, s);
    else if (cloud_layout == CRV_LAYOUT_TEX3D) launch_sv_out<CRV_LAYOUT_TEX3D>(c, out_layout, mode, &launches, s);
    else launch_sv_out<CRV_LAYOUT_LINEAR>(c, out_layout, mode, &launches, s);
    ctx->launches += launches;
    CRV_CUDA(cudaGetLastError());
    return CRV_OK;
}
