// fbm.cu -- synthetic fBm cloud generator on the device (bench/test scaffolding, NOT a reference feature).
//
// The reference ships no cloud data (models/PUT_YOUR_MODELS_HERE.txt is empty), so every benchmark runs on the
// frozen synthetic scene of SURVEY.md section 8d.  The definition (all fp32, no contraction, so that the
// device and the CPU oracle's generator agree bit for bit):
//   mix32(h)      : h ^= h>>16; h *= 0x7feb352d; h ^= h>>15; h *= 0x846ca68b; h ^= h>>16
//   lattice(i,j,k,s) = float(mix32(i*0x8da6b343 ^ j*0xd8163841 ^ k*0xcb1ab31f ^ s*0x9e3779b9) >> 8) * 2^-24
//   noise(p,s)    : trilinear blend of the 8 lattice values around p, weights faded by t*t*(3-2t),
//                   blended along x, then y, then z as a + t*(b-a)
//   fbm(p)        : sum_{o<5} 0.5^o * noise(p * 4*2^o, seed+o) / 1.9375
//   p             = ((i+0.5)/n, (j+0.5)/n, (k+0.5)/n);  r2 = |p - 0.5|^2 / 0.25
//   density       = clamp((fbm - 0.35 - 0.65*r2) / 0.35, 0, 1)
// The cloud box is 1000^3 world units for every n (voxel size 1/n, baseScaling 1000; main.cpp:930,147-150).
#include "crv_internal.h"

namespace {

__device__ __forceinline__ uint32_t mix32(uint32_t h)
{
    h ^= h >> 16;
    h *= 0x7feb352du;
    h ^= h >> 15;
    h *= 0x846ca68bu;
    h ^= h >> 16;
    return h;
}

__device__ __forceinline__ float lattice(uint32_t i, uint32_t j, uint32_t k, uint32_t s)
{
    const uint32_t h = mix32((i * 0x8da6b343u) ^ (j * 0xd8163841u) ^ (k * 0xcb1ab31fu) ^ (s * 0x9e3779b9u));
    return (float)(h >> 8) * 5.9604644775390625e-08f;
}

__device__ __forceinline__ float fade(float t) { return (t * t) * (3.0f - 2.0f * t); }
__device__ __forceinline__ float mixf(float a, float b, float t) { return a + t * (b - a); }

__device__ float value_noise(float x, float y, float z, uint32_t s)
{
    const float fx = floorf(x), fy = floorf(y), fz = floorf(z);
    const uint32_t i = (uint32_t)(int)fx, j = (uint32_t)(int)fy, k = (uint32_t)(int)fz;
    const float u = fade(x - fx), v = fade(y - fy), w = fade(z - fz);
    const float x00 = mixf(lattice(i, j, k, s), lattice(i + 1, j, k, s), u);
    const float x10 = mixf(lattice(i, j + 1, k, s), lattice(i + 1, j + 1, k, s), u);
    const float x01 = mixf(lattice(i, j, k + 1, s), lattice(i + 1, j, k + 1, s), u);
    const float x11 = mixf(lattice(i, j + 1, k + 1, s), lattice(i + 1, j + 1, k + 1, s), u);
    return mixf(mixf(x00, x10, v), mixf(x01, x11, v), w);
}

__global__ void __launch_bounds__(256) fbm_kernel(float* __restrict__ out, uint32_t n, uint32_t seed)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t j = blockIdx.y * blockDim.y + threadIdx.y;
    const uint32_t k = blockIdx.z;
    if (i >= n || j >= n) return;
    const float fn = (float)n;
    const float px = ((float)i + 0.5f) / fn, py = ((float)j + 0.5f) / fn, pz = ((float)k + 0.5f) / fn;
    float sum = 0.0f, amp = 1.0f, freq = 4.0f;
#pragma unroll
    for (uint32_t o = 0; o < 5; ++o) {
        sum = sum + amp * value_noise(px * freq, py * freq, pz * freq, seed + o);
        freq = freq * 2.0f;
        amp = amp * 0.5f;
    }
    const float fbm = sum / 1.9375f;
    const float dx = px - 0.5f, dy = py - 0.5f, dz = pz - 0.5f;
    const float r2 = ((dx * dx + dy * dy) + dz * dz) / 0.25f;
    float d = ((fbm - 0.35f) - 0.65f * r2) / 0.35f;
    d = fminf(fmaxf(d, 0.0f), 1.0f);
    out[(size_t)i + (size_t)n * ((size_t)j + (size_t)n * k)] = d;
}

}  // namespace

int crv_launch_fbm(crv_ctx* ctx, float* linear_dst, uint32_t n, uint32_t seed, cudaStream_t s)
{
    const dim3 block(64, 4, 1), grid((n + 63) / 64, (n + 3) / 4, n);
    fbm_kernel<<<grid, block, 0, s>>>(linear_dst, n, seed);
    ctx->launches++;
    CRV_CUDA(cudaGetLastError());
    return CRV_OK;
}
