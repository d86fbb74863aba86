// oracle/crv_oracle.cpp -- CPU ORACLE.  TEST INFRASTRUCTURE ONLY (see crv_oracle.h for the contract).
//
// Function-for-function scalar fp32 restatement of the reference shaders; every function cites the shader
// lines it follows.  Build: g++ -O2 -ffp-contract=off -fno-fast-math -fopenmp (oracle/Makefile).
// PARITY STATUS: pinned bit for bit against the reference's own shader sources compiled by g++ (oracle/_ref/
// libcrv_refshaders.so, tests/test_ref_shaders.py); driver-defined behaviour (no SPIR-V / Vulkan in this image) is
// unpinned and fixed as documented in crv_oracle.h.
#include "crv_oracle.h"

#include <cmath>
#include <cstring>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

struct vec3 { float x, y, z; };
struct vec4 { float x, y, z, w; };

inline vec3 operator+(vec3 a, vec3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline vec3 operator-(vec3 a, vec3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline vec3 operator*(vec3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline vec3 operator*(float s, vec3 a) { return {s * a.x, s * a.y, s * a.z}; }
inline vec3 operator-(vec3 a) { return {-a.x, -a.y, -a.z}; }
inline vec4 operator+(vec4 a, vec4 b) { return {a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w}; }
inline vec4 operator-(vec4 a, vec4 b) { return {a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w}; }
inline vec4 operator*(vec4 a, float s) { return {a.x * s, a.y * s, a.z * s, a.w * s}; }
inline vec4 operator*(float s, vec4 a) { return {s * a.x, s * a.y, s * a.z, s * a.w}; }
inline vec4 operator/(vec4 a, vec4 b) { return {a.x / b.x, a.y / b.y, a.z / b.z, a.w / b.w}; }
// dot products are summed left to right
inline float dot(vec3 a, vec3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
inline float length(vec3 a) { return std::sqrt(dot(a, a)); }

// ---- uniform blocks, read at the std140 offsets the shaders declare -------------------------------------
struct Camera {          // PathTracer.comp:43-55
    vec3 position; int halfWidth; vec3 forward; int halfHeight; vec3 right; float nearPlane;
    vec3 up; float pixelSizeY; float pixelSizeX;
};
struct Cloud {           // PathTracer.comp:58-66
    vec4 bounds[2]; uint32_t voxelCount[4]; float maxExtinction, baseScaling, densityScaling;
};
struct Params {          // PathTracer.comp:68-78
    uint32_t maxRayBounces; float lightIntensity, phaseG, phaseOnePlusG2, phaseOneMinusG2, phaseOneOver2G;
    bool isotropic;
};
struct Shadow {          // PathTracer.comp:81-91
    vec4 bounds[2]; vec4 lightDirection, right, up; vec4 basisChange[4]; uint32_t voxelAxisCount; float voxelSize;
};
struct Push {            // PathTracer.comp:93-99
    double time; int32_t seed; uint32_t frameCount; float pmRadius;
};

template <typename T> inline T rd(const void* base, size_t off)
{
    T v;
    std::memcpy(&v, static_cast<const char*>(base) + off, sizeof(T));
    return v;
}
inline vec3 rd3(const void* b, size_t off) { return {rd<float>(b, off), rd<float>(b, off + 4), rd<float>(b, off + 8)}; }
inline vec4 rd4(const void* b, size_t off) { return {rd<float>(b, off), rd<float>(b, off + 4), rd<float>(b, off + 8), rd<float>(b, off + 12)}; }

Camera load_camera(const void* p)
{
    Camera c;
    c.position = rd3(p, 0);   c.halfWidth = rd<int32_t>(p, 12);
    c.forward = rd3(p, 16);   c.halfHeight = rd<int32_t>(p, 28);
    c.right = rd3(p, 32);     c.nearPlane = rd<float>(p, 44);
    c.up = rd3(p, 48);        c.pixelSizeY = rd<float>(p, 60);
    c.pixelSizeX = rd<float>(p, 64);
    return c;
}
Cloud load_cloud(const void* p)
{
    Cloud c;
    c.bounds[0] = rd4(p, 0); c.bounds[1] = rd4(p, 16);
    for (int i = 0; i < 4; ++i) c.voxelCount[i] = rd<uint32_t>(p, 32 + 4 * i);
    c.maxExtinction = rd<float>(p, 48); c.baseScaling = rd<float>(p, 52); c.densityScaling = rd<float>(p, 56);
    return c;
}
Params load_params(const void* p)
{
    Params q;
    q.maxRayBounces = rd<uint32_t>(p, 0); q.lightIntensity = rd<float>(p, 4); q.phaseG = rd<float>(p, 8);
    q.phaseOnePlusG2 = rd<float>(p, 12); q.phaseOneMinusG2 = rd<float>(p, 16); q.phaseOneOver2G = rd<float>(p, 20);
    // C++ bool = 1 byte + 3 uninitialised padding bytes on the host side (UniformBuffers.h:90): read the byte only
    q.isotropic = rd<uint8_t>(p, 24) != 0;
    return q;
}
Shadow load_shadow(const void* p)
{
    Shadow s;
    s.bounds[0] = rd4(p, 0); s.bounds[1] = rd4(p, 16);
    s.lightDirection = rd4(p, 32); s.right = rd4(p, 48); s.up = rd4(p, 64);
    for (int i = 0; i < 4; ++i) s.basisChange[i] = rd4(p, 80 + 16 * i);
    s.voxelAxisCount = rd<uint32_t>(p, 144); s.voxelSize = rd<float>(p, 148);
    return s;
}
Push load_push(const void* p)
{
    Push q;
    q.time = rd<double>(p, 0); q.seed = rd<int32_t>(p, 8); q.frameCount = rd<uint32_t>(p, 12); q.pmRadius = rd<float>(p, 16);
    return q;
}

// ---- constants (PathTracer.comp:17-37) ------------------------------------------------------------------
const float PI = 3.14159265359f;
const float INV_4Pi = 1.0f / (4.0f * PI);
const vec4 BG_COLORS[5] = {{0.00f, 0.0f, 0.02f, 1.0f}, {0.01f, 0.05f, 0.2f, 1.0f}, {0.7f, 0.9f, 1.0f, 1.0f},
                           {0.1f, 0.3f, 1.0f, 1.0f}, {0.01f, 0.1f, 0.7f, 1.0f}};
const float BG_DISTS[5] = {-1.0f, -0.04f, 0.0f, 0.5f, 1.0f};

// ---- RNG (PathTracer.comp:220-241) ----------------------------------------------------------------------
struct Rng {
    uint32_t state;
    uint32_t rand_xorshift()                 // :221-228
    {
        state ^= (state << 13);
        state ^= (state >> 17);
        state ^= (state << 5);
        return state;
    }
    float generateRandomNumber()             // :230-233 ; u32 -> f32 is round-to-nearest-even, range [0,1]
    {
        return static_cast<float>(rand_xorshift()) / 4294967296.0f;
    }
    void initializeRandom(uint32_t seed)     // :235-241 ; the bound seed % 7 + 2 uses the ORIGINAL seed
    {
        state = seed;
        const uint32_t n = seed % 7u + 2u;
        for (uint32_t i = 0; i < n; ++i) generateRandomNumber();
    }
};
// :368  int seed * uint(gid.x + gid.y * 32 * numGroups.x), all mod 2^32
inline uint32_t seed_state(int32_t seed, uint32_t gx, uint32_t gy, uint32_t ngx)
{
    return static_cast<uint32_t>(seed) * (gx + gy * 32u * ngx);
}

// ---- texture(sampler3D) : LINEAR, CLAMP_TO_EDGE, unnormalizedCoordinates = false -------------------------
// Vulkan texel filtering: u = s*size; i0 = floor(u-0.5); i1 = i0+1; alpha = frac(u-0.5); wrap by clamping the
// integer texel index to [0,size-1]; blend along x, then y, then z as v0 + a*(v1 - v0).
inline int clampi(int v, int hi) { return v < 0 ? 0 : (v > hi ? hi : v); }
inline int to_int_sat(float f)
{
    // NaN -> 0 and saturation, so that a NaN coordinate cannot index out of bounds (result is NaN anyway)
    if (!(f == f)) return 0;
    if (f < -2147483000.0f) return -2147483647 - 1;
    if (f > 2147483000.0f) return 2147483647;
    return static_cast<int>(f);
}
float texture3d(const float* vol, uint32_t nx, uint32_t ny, uint32_t nz, vec3 s)
{
    const float u = s.x * static_cast<float>(nx) - 0.5f;
    const float v = s.y * static_cast<float>(ny) - 0.5f;
    const float w = s.z * static_cast<float>(nz) - 0.5f;
    const float fu = std::floor(u), fv = std::floor(v), fw = std::floor(w);
    const float a = u - fu, b = v - fv, c = w - fw;
    const int i0 = to_int_sat(fu), j0 = to_int_sat(fv), k0 = to_int_sat(fw);
    const int i0c = clampi(i0, (int)nx - 1), i1c = clampi(i0 == 2147483647 ? i0 : i0 + 1, (int)nx - 1);
    const int j0c = clampi(j0, (int)ny - 1), j1c = clampi(j0 == 2147483647 ? j0 : j0 + 1, (int)ny - 1);
    const int k0c = clampi(k0, (int)nz - 1), k1c = clampi(k0 == 2147483647 ? k0 : k0 + 1, (int)nz - 1);
    const size_t sy = nx, sz = static_cast<size_t>(nx) * ny;
    auto T = [&](int i, int j, int k) -> float { return vol[i + sy * j + sz * k]; };
    const float x00 = T(i0c, j0c, k0c) + a * (T(i1c, j0c, k0c) - T(i0c, j0c, k0c));
    const float x10 = T(i0c, j1c, k0c) + a * (T(i1c, j1c, k0c) - T(i0c, j1c, k0c));
    const float x01 = T(i0c, j0c, k1c) + a * (T(i1c, j0c, k1c) - T(i0c, j0c, k1c));
    const float x11 = T(i0c, j1c, k1c) + a * (T(i1c, j1c, k1c) - T(i0c, j1c, k1c));
    const float y0 = x00 + b * (x10 - x00);
    const float y1 = x01 + b * (x11 - x01);
    return y0 + c * (y1 - y0);
}

// ---- shader helper functions ----------------------------------------------------------------------------
// :104-113  (lowp evaluated in fp32; mix(a,b,t) = a*(1-t) + b*t; smoothstep = Hermite of the clamped ratio)
vec4 sampleBackground(vec3 dir)
{
    const float dist = dir.y;
    vec4 col = BG_COLORS[0];
    for (int i = 1; i < 5; ++i) {
        float t = (dist - BG_DISTS[i - 1]) / (BG_DISTS[i] - BG_DISTS[i - 1]);
        t = std::fmin(std::fmax(t, 0.0f), 1.0f);         // clamp = min(max(t,0),1); fmax drops a NaN operand
        const float s = (t * t) * (3.0f - 2.0f * t);
        col = col * (1.0f - s) + BG_COLORS[i] * s;
    }
    return col;
}

// :115-119
float sampleCloud(const float* vol, const Cloud& cp, vec3 pos)
{
    const vec4 n = (vec4{pos.x, pos.y, pos.z, 0.0f} - cp.bounds[0]) / (cp.bounds[1] - cp.bounds[0]);
    return texture3d(vol, cp.voxelCount[0], cp.voxelCount[1], cp.voxelCount[2], vec3{n.x, n.y, n.z});
}

// mat4 * vec4 as the linear combination of columns, summed left to right
inline vec4 matvec(const vec4 m[4], vec4 v)
{
    return ((m[0] * v.x + m[1] * v.y) + m[2] * v.z) + m[3] * v.w;
}

// :121-127   (M * (vec4(pos,1) - b0)) / (M * (b1 - b0))
float sampleShadowVolume(const float* sv, const Shadow& sp, vec3 pos)
{
    const vec4 num = matvec(sp.basisChange, vec4{pos.x, pos.y, pos.z, 1.0f} - sp.bounds[0]);
    const vec4 den = matvec(sp.basisChange, sp.bounds[1] - sp.bounds[0]);
    const vec4 n = num / den;
    return texture3d(sv, sp.voxelAxisCount, sp.voxelAxisCount, sp.voxelAxisCount, vec3{n.x, n.y, n.z});
}

struct Ray { vec3 pos, dir; };

// :144-151
void getCameraRay(const Camera& c, int x, int y, Ray& ray)
{
    ray.dir = (c.forward * c.nearPlane + (c.right * c.pixelSizeX) * static_cast<float>(x - c.halfWidth)) -
              (c.up * c.pixelSizeY) * static_cast<float>(y - c.halfHeight);
    ray.pos = ray.dir + c.position;
    const float len = length(ray.dir);          // normalize(v) = v / sqrt(dot(v,v))
    ray.dir = vec3{ray.dir.x / len, ray.dir.y / len, ray.dir.z / len};
}

// :154-184 (isNegativeSign :139-142).  Comparisons are kept literally: NaNs (0 * inf) fall through them the
// same way they do in the shader.  Out-parameter order is (tmax, tmin).
bool intersectCloud(const Cloud& cp, const Ray& ray, float& tmax, float& tmin)
{
    float tymin = 0, tymax = 0, tzmin = 0, tzmax = 0;
    const vec3 invdir{1.0f / ray.dir.x, 1.0f / ray.dir.y, 1.0f / ray.dir.z};
    const int sign[3] = {int(invdir.x < 0), int(invdir.y < 0), int(invdir.z < 0)};

    tmin = (cp.bounds[sign[0]].x - ray.pos.x) * invdir.x;
    tmax = (cp.bounds[1 - sign[0]].x - ray.pos.x) * invdir.x;
    tymin = (cp.bounds[sign[1]].y - ray.pos.y) * invdir.y;
    tymax = (cp.bounds[1 - sign[1]].y - ray.pos.y) * invdir.y;

    if ((tmin > tymax) || (tymin > tmax)) return false;
    if (tymin > tmin) tmin = tymin;
    if (tymax < tmax) tmax = tymax;

    tzmin = (cp.bounds[sign[2]].z - ray.pos.z) * invdir.z;
    tzmax = (cp.bounds[1 - sign[2]].z - ray.pos.z) * invdir.z;

    if ((tmin > tzmax) || (tzmin > tmax)) return false;
    if (tzmin > tmin) tmin = tzmin;
    if (tzmax < tmax) tmax = tzmax;
    return true;
}

// :197-206
void createOrthonormalBasis(vec3 n, vec3& xp, vec3& yp)
{
    const float sz = n.z >= 0.0f ? 1.0f : -1.0f;
    const float a = n.y / (1.0f + std::fabs(n.z));
    const float b = n.y * a;
    const float c = (-n.x) * a;
    xp = vec3{n.z + sz * b, sz * c, -n.x};
    yp = vec3{c, 1.0f - b, (-sz) * n.y};
}

// :247-250
float calculatePDF(const Params& p, float costheta)
{
    return (INV_4Pi * p.phaseOneMinusG2) / std::pow(p.phaseOnePlusG2 - (2.0f * p.phaseG) * costheta, 1.5f);
}
// :252-256
float invertCDF(const Params& p, float xi)
{
    const float sqrTerm = p.phaseOneMinusG2 / ((1.0f - p.phaseG) + (2.0f * p.phaseG) * xi);
    return p.phaseOneOver2G * (p.phaseOnePlusG2 - sqrTerm * sqrTerm);
}
// :259-274
float samplePhase(const Params& p, vec3 incoming, vec3 sample)
{
    if (p.isotropic) return INV_4Pi;
    return calculatePDF(p, dot(incoming, sample));
}

// :277-311
void scatterRay(const Params& p, Rng& rng, vec3 incoming, vec3& out)
{
    if (p.isotropic) {
        float xi = rng.generateRandomNumber();
        out.z = xi * 2.0f - 1.0f;
        float sinTheta = 1.0f - out.z * out.z;          // actually sin^2
        if (sinTheta > 0.0f) {
            sinTheta = std::sqrt(sinTheta > 0.0f ? sinTheta : 0.0f);
            xi = rng.generateRandomNumber();
            const float phi = (xi * 2.0f) * PI;
            out.x = sinTheta * std::cos(phi);
            out.y = sinTheta * std::sin(phi);
        } else {
            out.x = out.y = 0.0f;                        // and NO second draw
        }
    } else {
        const float phi = (rng.generateRandomNumber() * 2.0f) * PI;     // phi is drawn first
        const float cosTheta = invertCDF(p, rng.generateRandomNumber());
        const float s2 = 1.0f - cosTheta * cosTheta;
        const float sinTheta = std::sqrt(s2 > 0.0f ? s2 : 0.0f);        // max(0, s2): NaN -> 0
        vec3 t0, t1;
        createOrthonormalBasis(incoming, t0, t1);
        out = ((sinTheta * std::cos(phi)) * t0 + (sinTheta * std::sin(phi)) * t1) + cosTheta * incoming;  // not renormalised
    }
}

struct Counters { uint32_t steps = 0, collisions = 0; };

// :318-361
bool findScatterPoint(const float* vol, const Cloud& cp, Rng& rng, Ray& ray, Counters& cnt)
{
    float tmax = 0, tmin = 0;
    if (!intersectCloud(cp, ray, tmax, tmin)) return false;

    const vec3 exitPoint = ray.pos + ray.dir * tmax;
    const float dist = length(ray.pos - exitPoint);
    float t = 0;
    vec3 currentPoint{0, 0, 0};
    const float sigma = (cp.maxExtinction * cp.densityScaling) / cp.baseScaling;
    do {
        const float zeta = rng.generateRandomNumber();
        ++cnt.steps;
        t += (-std::log(zeta)) / sigma;
        if (!(t < dist)) return false;                     // `t >= dist` (:339): left the cloud.  A NaN also leaves
                                                           // (the shader would spin forever on it)

        currentPoint = ray.pos + t * ray.dir;
        const float extinction = sampleCloud(vol, cp, currentPoint);

        const float xi = rng.generateRandomNumber();
        if (xi < extinction / cp.maxExtinction) break;
    } while (true);

    ray.pos = currentPoint;
    return true;
}

// ==========================================================================================================
// PPM photon tracer: shaders/PPM_PT.comp.  Helper bodies shared with PathTracer.comp are reused
// (rand_xorshift / generateRandomNumber / intersectCloud / invertCDF / sampleCloud are byte-identical there); what
// differs is restated here with its PPM_PT.comp lines: scatterRay ends in normalize() (:296), the delta-tracking exit
// test is `t > dist` (:414), and the path deposits photons instead of gathering sun light.
// ==========================================================================================================
struct PhotonMapProps {   // PPM_PT.comp:57-67 (std140 offsets 0 / 16,32 / 48 / 60 / 64 / 68 / 72)
    vec4 lightDirection; vec4 bounds[2]; int32_t voxelCount[3]; float voxelSize; uint32_t photonSize; float stepSize, absorption;
};
PhotonMapProps load_photon_map_props(const void* p)
{
    PhotonMapProps q;
    q.lightDirection = rd4(p, 0); q.bounds[0] = rd4(p, 16); q.bounds[1] = rd4(p, 32);
    for (int i = 0; i < 3; ++i) q.voxelCount[i] = rd<int32_t>(p, 48 + 4 * i);
    q.voxelSize = rd<float>(p, 60); q.photonSize = rd<uint32_t>(p, 64); q.stepSize = rd<float>(p, 68); q.absorption = rd<float>(p, 72);
    return q;
}
struct Photon { float position[4], power[4], phi, theta, padding[2]; };     // PPM_PT.comp:14-21, 48 bytes (std430)
const uint32_t ELEMENTS_PER_CELL = 32;                                       // :26

inline vec3 normalize3(vec3 v) { const float l = std::sqrt(dot(v, v)); return {v.x / l, v.y / l, v.z / l}; }
inline float clampf(float x, float lo, float hi) { const float a = x < lo ? lo : x; return hi < a ? hi : a; }   // min(max(x, lo), hi)

float SphericalTheta(vec3 v) { return std::acos(clampf(v.z, -1.0f, 1.0f)); }                 // :180-182
float SphericalPhi(vec3 v)                                                                   // :184-187
{
    const float p = std::atan2(v.y, v.x) * 2.0f;
    return (p < 0.0f) ? (p + 2.0f * PI) : p;
}
vec3 dirFromPolar(float theta, float phi)                                                    // :189-195
{
    return {std::sin(theta) * std::cos(phi), std::sin(theta) * std::sin(phi), std::cos(theta)};
}

// :262-297: PathTracer.comp's scatterRay followed by normalize(sampleDirection)
void scatterRayPPM(const Params& p, Rng& rng, vec3 incoming, vec3& out)
{
    scatterRay(p, rng, incoming, out);
    out = normalize3(out);
}

struct PhotonSink {
    Photon* photons;
    uint32_t* collisions;     // uvec4 per cell, .x is the counter
    vec4 currentColor;
};

// :303-321
void depositPhoton(const PhotonMapProps& pm, PhotonSink& sink, vec3 pos, vec3 dir)
{
    const vec3 normPos = {(pos.x - pm.bounds[0].x) / pm.voxelSize, (pos.y - pm.bounds[0].y) / pm.voxelSize, (pos.z - pm.bounds[0].z) / pm.voxelSize};
    const int32_t ix = (int32_t)normPos.x, iy = (int32_t)normPos.y, iz = (int32_t)normPos.z;      // ivec3(float): truncation
    const int32_t idx = ix + (iy + iz * pm.voxelCount[1]) * pm.voxelCount[0];
    Photon ph;
    ph.position[0] = pos.x; ph.position[1] = pos.y; ph.position[2] = pos.z; ph.position[3] = 0.0f;
    ph.power[0] = sink.currentColor.x; ph.power[1] = sink.currentColor.y; ph.power[2] = sink.currentColor.z; ph.power[3] = sink.currentColor.w;
    ph.theta = SphericalTheta(dir);
    ph.phi = SphericalPhi(dir);
    ph.padding[0] = ph.padding[1] = 0.0f;
    uint32_t offset;
#pragma omp atomic capture
    { offset = sink.collisions[4 * (size_t)idx]; sink.collisions[4 * (size_t)idx] += 1u; }
    if (offset < ELEMENTS_PER_CELL) sink.photons[(size_t)idx * ELEMENTS_PER_CELL + offset] = ph;
}

// :324-361 (the commented-out face selection is dead; pdf = 1 / projected area, a point drawn in the box moved to tmin)
void generateSunRay(const Cloud& cp, const PhotonMapProps& pm, Rng& rng, Ray& ray, float& pdf)
{
    const vec4 bd = cp.bounds[1] - cp.bounds[0];
    const vec3 L = {pm.lightDirection.x, pm.lightDirection.y, pm.lightDirection.z};
    const vec3 areas = {(bd.z * bd.y) * std::fabs(L.x), (bd.x * bd.z) * std::fabs(L.y), (bd.y * bd.x) * std::fabs(L.z)};   // zxy * yzx * abs(L)
    const float totalArea = (areas.x * 1.0f + areas.y * 1.0f) + areas.z * 1.0f;                 // dot(areas, vec3(1))
    pdf = 1.0f / totalArea;
    ray.dir = L;
    const float r0 = rng.generateRandomNumber(), r1 = rng.generateRandomNumber(), r2 = rng.generateRandomNumber();   // left to right
    ray.pos = vec3{cp.bounds[0].x, cp.bounds[0].y, cp.bounds[0].z} + vec3{bd.x * r0, bd.y * r1, bd.z * r2};
    float tmax = 0, tmin = 0;
    intersectCloud(cp, ray, tmax, tmin);
    ray.pos = ray.pos + ray.dir * tmin;
}

// :363-387
void generateAmbientRay(const Cloud& cp, Rng& rng, Ray& ray, float& pdf)
{
    const float theta = (2.0f * PI) * rng.generateRandomNumber();                               // 2 * PI * rnd: int*float first
    const float phi = std::acos(1.0f - 2.0f * rng.generateRandomNumber());
    ray.dir = normalize3(dirFromPolar(theta, phi));
    const vec4 bd4 = cp.bounds[1] - cp.bounds[0];
    const vec3 bd = {bd4.x, bd4.y, bd4.z};
    const vec3 areas = {(bd.z * bd.y) * std::fabs(ray.dir.x), (bd.x * bd.z) * std::fabs(ray.dir.y), (bd.y * bd.x) * std::fabs(ray.dir.z)};
    const float totalArea = (areas.x * 1.0f + areas.y * 1.0f) + areas.z * 1.0f;
    const float r0 = rng.generateRandomNumber(), r1 = rng.generateRandomNumber(), r2 = rng.generateRandomNumber();
    const vec3 q = vec3{cp.bounds[0].x, cp.bounds[0].y, cp.bounds[0].z} + vec3{bd.x * r0, bd.y * r1, bd.z * r2};
    ray.pos = {q.x + 0.1f, q.y + 0.1f, q.z + 0.1f};
    pdf = INV_4Pi / totalArea;
    float tmax = 0, tmin = 0;
    intersectCloud(cp, ray, tmax, tmin);
    ray.pos = ray.pos + ray.dir * tmin;
}

// :395-446; returns true to keep scattering.  `steps` counts delta-tracking iterations (log calls), like the path tracer's.
bool interactWithMedium(const float* vol, const Cloud& cp, const PhotonMapProps& pm, Rng& rng, PhotonSink& sink, Ray& ray,
                        bool& firstScatter, uint64_t& steps)
{
    float tmax = 0, tmin = 0;
    if (!intersectCloud(cp, ray, tmax, tmin)) return false;
    const vec3 exitPoint = ray.pos + ray.dir * tmax;
    const float dist = length(ray.pos - exitPoint);
    float t = 0;
    vec3 currentPoint{0, 0, 0};
    const float sigma = (cp.maxExtinction * cp.densityScaling) / cp.baseScaling;
    do {
        const float zeta = rng.generateRandomNumber();
        ++steps;
        t += (-std::log(zeta)) / sigma;
        if (t > dist) return false;                        // :414 -- strictly greater, unlike PathTracer.comp:339
        if (!(t == t)) return false;                       // a NaN would spin forever in the shader
        currentPoint = ray.pos + t * ray.dir;
        const float extinction = sampleCloud(vol, cp, currentPoint);
        const float xi = rng.generateRandomNumber();
        if (xi < extinction / cp.maxExtinction) break;
    } while (true);
    ray.pos = currentPoint;
    if (firstScatter) firstScatter = false;                // :434-441: the first scattering event deposits nothing
    else depositPhoton(pm, sink, currentPoint, ray.dir);
    const float xi = rng.generateRandomNumber();
    return xi > pm.absorption;                             // :445-446
}

// :451-488, one invocation
void tracePhoton(const float* vol, const Cloud& cp, const Params& par, const PhotonMapProps& pm, int32_t seed, uint32_t gid_x,
                 uint32_t emittedPhotons, Photon* photons, uint32_t* collisions, uint64_t& steps)
{
    if (cp.densityScaling <= 0.0f) return;
    Rng rng;
    rng.initializeRandom((uint32_t)seed * gid_x);          // gid.y = 0 in the (groups,1,1) dispatch
    PhotonSink sink{photons, collisions, {0, 0, 0, 0}};
    bool firstScatter = true;
    Ray ray{{0, 0, 0}, {0, 0, 0}};
    float pdf = 0;
    const float prob = 0.5f;
    if (rng.generateRandomNumber() < prob) {
        generateSunRay(cp, pm, rng, ray, pdf);
        const float s = (1.0f * par.lightIntensity) / (pdf * (float)emittedPhotons);    // SUNLIGHT_COLOR * I / (pdf * N)
        sink.currentColor = {s / prob, s / prob, s / prob, s / prob};
    } else {
        generateAmbientRay(cp, rng, ray, pdf);
        const vec4 bg = sampleBackground(-ray.dir);
        const float d = pdf * (float)emittedPhotons;
        const vec4 c = {(bg.x / d) * 0.5f, (bg.y / d) * 0.5f, (bg.z / d) * 0.5f, (bg.w / d) * 0.5f};
        const float q = 1.0f - prob;
        sink.currentColor = {c.x / q, c.y / q, c.z / q, c.w / q};
    }
    sink.currentColor = sink.currentColor * 0.5f;
    vec3 newDir{0, 0, 0};
    while (interactWithMedium(vol, cp, pm, rng, sink, ray, firstScatter, steps)) {
        scatterRayPPM(par, rng, ray.dir, newDir);
        ray.dir = newDir;
    }
}

// ----------------------------------------------------------------------------------------------------------
// PPM photon estimate: shaders/PPM_PE.comp (the second dispatch of RenderTechniquePPM::RecordDrawCommands).  One invocation
// per pixel: the camera ray is marched through the cloud back to front in steps of stepSize; every step adds the sun term
// through the shadow volume, a sky term along one phase-sampled direction, and the photon-map density estimate.
// getCameraRay / intersectCloud / sampleCloud / sampleShadowVolume / samplePhase / scatterRay / sampleBackground are the
// same text as PathTracer.comp's (scatterRay without PPM_PT's normalize, PPM_PE.comp:391-425).
// ----------------------------------------------------------------------------------------------------------
bool isInCloud(const Cloud& cp, vec3 p)                 // PPM_PE.comp:186-194
{
    return !(p.x < cp.bounds[0].x || p.y < cp.bounds[0].y || p.z < cp.bounds[0].z || p.x > cp.bounds[1].x || p.y > cp.bounds[1].y ||
             p.z > cp.bounds[1].z);
}

struct PhotonMapView {
    const Photon* photons;
    const uint32_t* collisions;
    uint64_t cells;
};

// :302-375.  Reads beyond the 32 slots of an overflowed cell run into the following cells' slots like the shader's loop
// (`i < collisions[idx].x`); beyond the buffer the read returns zeros (robust buffer access).
vec4 samplePhotonMap(const float* vol, const Cloud& cp, const Params& par, const PhotonMapProps& pm, const PhotonMapView& map, float pmRadius,
                     Rng& rng, vec3 pos, vec3 dir, float scatter)
{
    (void)vol; (void)cp;
    if (scatter == 0.0f) return {0, 0, 0, 0};
    const vec3 gp = {pos.x - pm.bounds[0].x, pos.y - pm.bounds[0].y, pos.z - pm.bounds[0].z};
    auto imax = [](int a, int b) { return a < b ? b : a; };
    auto imin = [](int a, int b) { return b < a ? b : a; };
    const int xMin = imax((int)((gp.x - pmRadius) / pm.voxelSize), 0), xMax = imin((int)((gp.x + pmRadius) / pm.voxelSize), pm.voxelCount[0] - 1);
    const int yMin = imax((int)((gp.y - pmRadius) / pm.voxelSize), 0), yMax = imin((int)((gp.y + pmRadius) / pm.voxelSize), pm.voxelCount[1] - 1);
    const int zMin = imax((int)((gp.z - pmRadius) / pm.voxelSize), 0), zMax = imin((int)((gp.z + pmRadius) / pm.voxelSize), pm.voxelCount[2] - 1);
    vec4 acc = {0, 0, 0, 0};
    const float sqrRadius = pmRadius * pmRadius;
    for (int z = zMin; z <= zMax; ++z)
        for (int y = yMin; y <= yMax; ++y)
            for (int x = xMin; x <= xMax; ++x) {
                const int idx = x + (y + z * pm.voxelCount[1]) * pm.voxelCount[0];
                const uint32_t n = map.collisions[4 * (size_t)idx];
                for (uint32_t i = 0; i < n; ++i) {
                    const uint64_t o = (uint64_t)idx * ELEMENTS_PER_CELL + i;
                    if (o >= map.cells * ELEMENTS_PER_CELL) continue;           // zeros: position (0,0,0), power 0 -> adds nothing
                    const Photon& ph = map.photons[o];
                    const vec3 dv = {ph.position[0] - pos.x, ph.position[1] - pos.y, ph.position[2] - pos.z};
                    if (dot(dv, dv) <= sqrRadius) {
                        const vec3 pdir = dirFromPolar(ph.theta, ph.phi);       // (theta, phi) in the ESTIMATE's argument order (:343)
                        const float w = samplePhase(par, dir, pdir);
                        acc = acc + vec4{w * ph.power[0], w * ph.power[1], w * ph.power[2], w * ph.power[3]};
                    }
                }
            }
    if (((((0.0f + acc.x) + acc.y) + acc.z) + acc.w) / 4.0f <= 0.0f) return {0, 0, 0, 0};      // average() :129-137
    (void)(5.0f * rng.generateRandomNumber());      // sampleDistance: the six density samples that follow feed nothing (:357-366)
    const float d = ((PI * 4.0f / 3.0f) * pmRadius * sqrRadius) * scatter;                    // PI4_3 * r * r^2 * scatter, left to right
    return {acc.x / d, acc.y / d, acc.z / d, acc.w / d};
}

// PPM_PE.comp:430-545 without the final accumulate; returns `result`
vec4 estimatePixel(const float* vol, const float* sv, const Camera& cam, const Cloud& cp, const Params& par, const Shadow& sp,
                   const PhotonMapProps& pm, const PhotonMapView& map, int32_t seed, float pmRadius, uint32_t gx, uint32_t gy, uint32_t ngx,
                   uint64_t& marchSteps)
{
    Rng rng;
    rng.initializeRandom(seed_state(seed, gx, gy, ngx));
    Ray ray;
    getCameraRay(cam, (int)gx, (int)gy, ray);
    vec4 result = sampleBackground(ray.dir);
    float tmax = 0, tmin = 0;
    if (intersectCloud(cp, ray, tmax, tmin) && cp.densityScaling > 0.0f) {
        const vec3 closestPoint = ray.pos + ray.dir * tmin;
        ray.pos = ray.pos + ray.dir * (tmax - pm.stepSize * rng.generateRandomNumber());
        const float maxDistance = length(ray.pos - closestPoint);
        float currentDistance = 0;
        const vec3 L = {sp.lightDirection.x, sp.lightDirection.y, sp.lightDirection.z};
        while (currentDistance < maxDistance) {
            ++marchSteps;
            const float currentStep = pm.stepSize;
            ray.pos = ray.pos - ray.dir * currentStep;
            currentDistance += currentStep;
            const float extinction = (sampleCloud(vol, cp, ray.pos) * cp.densityScaling) / cp.baseScaling;
            const float albedo = 1.0f - pm.absorption;
            const float scatter = albedo * extinction;
            const float pdf = samplePhase(par, -ray.dir, L);
            float accumulatedDensity = sampleShadowVolume(sv, sp, ray.pos);
            float direct = ((1.0f * par.lightIntensity) * pdf) * accumulatedDensity;              // all four channels equal
            vec4 directRadiance = {direct, direct, direct, direct};
            vec3 ambientRayDir{0, 0, 0};
            scatterRay(par, rng, ray.dir, ambientRayDir);
            accumulatedDensity = 0;
            float accumulatedDistance = 0;
            uint32_t samples = 0;
            vec3 pos = ray.pos;
            const float sampleDist = 100.0f;
            vec4 ambientRadiance = sampleBackground(ambientRayDir);
            while (isInCloud(cp, pos)) {
                pos = pos + ambientRayDir * sampleDist;
                accumulatedDensity += sampleCloud(vol, cp, pos);
                accumulatedDistance += sampleDist;
                samples++;
            }
            accumulatedDensity *= cp.densityScaling / cp.baseScaling;
            const float e = std::exp(((-accumulatedDistance) * accumulatedDensity) / (float)samples);   // samples == 0: 0/0 -> NaN -> not added
            ambientRadiance = ambientRadiance * e;
            if (ambientRadiance.x > 0.0f) directRadiance = directRadiance + ambientRadiance;
            const vec4 pmr = samplePhotonMap(vol, cp, par, pm, map, pmRadius, rng, ray.pos, -ray.dir, scatter);
            const vec4 indirectRadiance = albedo * pmr;
            const float attenuation = std::exp((-currentStep) * extinction);
            result = result * attenuation;
            result = result + (scatter * (directRadiance + indirectRadiance)) * currentStep;
        }
    }
    return result;
}

// :366-425 without the final accumulate; returns `result`
vec4 tracePixel(const float* vol, const float* sv, const Camera& cam, const Cloud& cp, const Params& par,
                const Shadow& sp, int32_t seed, float absorption, uint32_t gx, uint32_t gy, uint32_t ngx, Counters& cnt)
{
    Rng rng;
    rng.initializeRandom(seed_state(seed, gx, gy, ngx));

    Ray ray;
    vec4 result{0, 0, 0, 0};
    getCameraRay(cam, (int)gx, (int)gy, ray);

    float tmax = 0, tmin = 0;
    if (!intersectCloud(cp, ray, tmax, tmin) || tmax < 0 || cp.densityScaling <= 0) {
        result = sampleBackground(ray.dir);
    } else {
        ray.pos = ray.pos + ray.dir * (tmax < tmin ? tmax : tmin);
        Ray currentRay = ray;
        const vec3 lightDir{sp.lightDirection.x, sp.lightDirection.y, sp.lightDirection.z};
        while (true) {
            if (!findScatterPoint(vol, cp, rng, ray, cnt)) {
                result = result + sampleBackground(ray.dir);
                break;
            }
            ++cnt.collisions;
            const float pdf = samplePhase(par, -ray.dir, lightDir);
            const float accumulatedDensity = sampleShadowVolume(sv, sp, ray.pos);
            const float c = ((1.0f * par.lightIntensity) * pdf) * accumulatedDensity;   // SUNLIGHT_COLOR = vec4(1)
            result = result + vec4{c, c, c, c};

            // EXTENSION (not in PathTracer.comp; idiom of PPM_PT.comp:445-446): survive iff xi > absorption.
            // No draw at all when absorption == 0, so the reference RNG stream is untouched.
            if (absorption > 0.0f) {
                const float xi = rng.generateRandomNumber();
                if (!(xi > absorption)) break;             // absorbed: no background term
            }

            scatterRay(par, rng, currentRay.dir, ray.dir);
            currentRay = ray;
        }
    }
    return result;
}

void trace_rect(const float* vol, const float* sv, const void* camera68, const void* cloud60, const void* params28,
                const void* svprops152, const void* push24, float absorption, int accum_sum, int x0, int y0, int x1,
                int y1, float* image, float* sample, uint32_t* steps, uint32_t* collisions, int num_threads,
                int row_step = 1)
{
    const Camera cam = load_camera(camera68);
    const Cloud cp = load_cloud(cloud60);
    const Params par = load_params(params28);
    const Shadow sp = load_shadow(svprops152);
    const Push pc = load_push(push24);
    const int W = cam.halfWidth * 2, H = cam.halfHeight * 2;
    const uint32_t ngx = static_cast<uint32_t>(W / 32 + 1);      // RenderTechniquePT.cpp:141
    if (x0 < 0) x0 = 0;
    if (y0 < 0) y0 = 0;
    if (x1 > W) x1 = W;
    if (y1 > H) y1 = H;
    if (row_step < 1) row_step = 1;
    const int rw = x1 - x0, rh = (y1 - y0 + row_step - 1) / row_step;     // rows y0, y0+row_step, ...
    if (rw <= 0 || rh <= 0) return;
    const int64_t count = static_cast<int64_t>(rw) * rh;
#ifdef _OPENMP
    if (num_threads <= 0) num_threads = omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 64) num_threads(num_threads)
#endif
    for (int64_t idx = 0; idx < count; ++idx) {
        const int x = x0 + static_cast<int>(idx % rw), y = y0 + static_cast<int>(idx / rw) * row_step;
        Counters cnt;
        const vec4 r = tracePixel(vol, sv, cam, cp, par, sp, pc.seed, absorption, (uint32_t)x, (uint32_t)y, ngx, cnt);
        const size_t p = static_cast<size_t>(y) * W + x;
        if (sample) { sample[4 * p] = r.x; sample[4 * p + 1] = r.y; sample[4 * p + 2] = r.z; sample[4 * p + 3] = r.w; }
        if (steps) steps[p] = cnt.steps;
        if (collisions) collisions[p] = cnt.collisions;
        if (image) {
            float* px = image + 4 * p;
            const float in[4] = {r.x, r.y, r.z, r.w};
            for (int c = 0; c < 4; ++c) {
                if (accum_sum) {
                    px[c] = px[c] + in[c];
                } else {
                    // :420-424  result += old * frameCount; result /= frameCount + 1
                    float v = in[c] + px[c] * static_cast<float>(pc.frameCount);
                    v = v / static_cast<float>(pc.frameCount + 1u);
                    px[c] = v;
                }
            }
        }
    }
}

// ---- synthetic fBm cloud (frozen generator) ----------------------------------------
inline uint32_t mix32(uint32_t h)
{
    h ^= h >> 16; h *= 0x7feb352du; h ^= h >> 15; h *= 0x846ca68bu; h ^= h >> 16;
    return h;
}
inline float lattice(uint32_t ix, uint32_t iy, uint32_t iz, uint32_t s)
{
    const uint32_t h = mix32((ix * 0x8da6b343u) ^ (iy * 0xd8163841u) ^ (iz * 0xcb1ab31fu) ^ (s * 0x9e3779b9u));
    return static_cast<float>(h >> 8) * (1.0f / 16777216.0f);
}
inline float fade(float t) { return (t * t) * (3.0f - 2.0f * t); }
inline float lerp(float a, float b, float t) { return a + t * (b - a); }
float value_noise(float px, float py, float pz, uint32_t s)
{
    const float fx = std::floor(px), fy = std::floor(py), fz = std::floor(pz);
    const uint32_t ix = (uint32_t)(int)fx, iy = (uint32_t)(int)fy, iz = (uint32_t)(int)fz;
    const float u = fade(px - fx), v = fade(py - fy), w = fade(pz - fz);
    const float x00 = lerp(lattice(ix, iy, iz, s), lattice(ix + 1, iy, iz, s), u);
    const float x10 = lerp(lattice(ix, iy + 1, iz, s), lattice(ix + 1, iy + 1, iz, s), u);
    const float x01 = lerp(lattice(ix, iy, iz + 1, s), lattice(ix + 1, iy, iz + 1, s), u);
    const float x11 = lerp(lattice(ix, iy + 1, iz + 1, s), lattice(ix + 1, iy + 1, iz + 1, s), u);
    return lerp(lerp(x00, x10, v), lerp(x01, x11, v), w);
}
float fbm_density(uint32_t i, uint32_t j, uint32_t k, uint32_t n, uint32_t seed)
{
    const float fn = static_cast<float>(n);
    const float px = (static_cast<float>(i) + 0.5f) / fn;
    const float py = (static_cast<float>(j) + 0.5f) / fn;
    const float pz = (static_cast<float>(k) + 0.5f) / fn;
    float sum = 0.0f, amp = 1.0f, freq = 4.0f;
    for (uint32_t o = 0; o < 5; ++o) {
        sum = sum + amp * value_noise(px * freq, py * freq, pz * freq, seed + o);
        freq = freq * 2.0f;
        amp = amp * 0.5f;
    }
    const float fbm = sum / 1.9375f;
    const float dx = px - 0.5f, dy = py - 0.5f, dz = pz - 0.5f;
    const float r2 = ((dx * dx + dy * dy) + dz * dz) / 0.25f;
    float d = ((fbm - 0.35f) - 0.65f * r2) / 0.35f;
    d = d < 0.0f ? 0.0f : (d > 1.0f ? 1.0f : d);
    return d;
}

}  // namespace

// =========================================================================================================
extern "C" {

uint32_t orc_rng_seed_state(int32_t seed, uint32_t gx, uint32_t gy, uint32_t ngx) { return seed_state(seed, gx, gy, ngx); }
uint32_t orc_rng_warmup_count(uint32_t state0) { return state0 % 7u + 2u; }

void orc_rng_draws(int32_t seed, uint32_t gx, uint32_t gy, uint32_t ngx, uint32_t n, uint32_t* states, float* floats)
{
    Rng r;
    r.initializeRandom(seed_state(seed, gx, gy, ngx));
    for (uint32_t i = 0; i < n; ++i) {
        const float f = r.generateRandomNumber();
        if (states) states[i] = r.state;
        if (floats) floats[i] = f;
    }
}

void orc_rng_tap(int32_t seed, uint32_t width, uint32_t height, uint32_t n_draws, uint32_t* out)
{
    const uint32_t ngx = width / 32u + 1u;
    for (uint32_t y = 0; y < height; ++y)
        for (uint32_t x = 0; x < width; ++x) {
            Rng r;
            r.initializeRandom(seed_state(seed, x, y, ngx));
            uint32_t* o = out + (static_cast<size_t>(y) * width + x) * n_draws;
            for (uint32_t i = 0; i < n_draws; ++i) o[i] = r.rand_xorshift();
        }
}

void orc_camera_ray(const void* camera68, int x, int y, float pos[3], float dir[3])
{
    Ray r;
    getCameraRay(load_camera(camera68), x, y, r);
    pos[0] = r.pos.x; pos[1] = r.pos.y; pos[2] = r.pos.z;
    dir[0] = r.dir.x; dir[1] = r.dir.y; dir[2] = r.dir.z;
}

int orc_intersect_cloud(const void* cloud60, const float pos[3], const float dir[3], float* tmax, float* tmin)
{
    Ray r{{pos[0], pos[1], pos[2]}, {dir[0], dir[1], dir[2]}};
    float a = 0, b = 0;
    const bool hit = intersectCloud(load_cloud(cloud60), r, a, b);
    *tmax = a; *tmin = b;
    return hit ? 1 : 0;
}

void orc_orthonormal_basis(const float n[3], float xp[3], float yp[3])
{
    vec3 a, b;
    createOrthonormalBasis(vec3{n[0], n[1], n[2]}, a, b);
    xp[0] = a.x; xp[1] = a.y; xp[2] = a.z;
    yp[0] = b.x; yp[1] = b.y; yp[2] = b.z;
}

float orc_calculate_pdf(const void* params28, float c) { return calculatePDF(load_params(params28), c); }
float orc_invert_cdf(const void* params28, float xi) { return invertCDF(load_params(params28), xi); }
float orc_sample_phase(const void* params28, const float in[3], const float s[3])
{
    return samplePhase(load_params(params28), vec3{in[0], in[1], in[2]}, vec3{s[0], s[1], s[2]});
}
uint64_t orc_ppm_trace(const float* cloud, const void* cloud60, const void* params28, const void* pmprops76, int32_t seed,
                       uint32_t groups_x, void* photons, uint32_t* collisions, uint64_t cells, int num_threads)
{
    const Cloud cp = load_cloud(cloud60);
    const Params par = load_params(params28);
    const PhotonMapProps pm = load_photon_map_props(pmprops76);
    std::memset(photons, 0, cells * ELEMENTS_PER_CELL * sizeof(Photon));       // RenderTechniquePPM.cpp:236-237
    std::memset(collisions, 0, cells * 4 * sizeof(uint32_t));
    const int total = (int)(groups_x * 64u);
    uint64_t steps = 0;
#pragma omp parallel for schedule(dynamic, 64) reduction(+ : steps) num_threads(num_threads > 0 ? num_threads : omp_get_max_threads())
    for (int i = 0; i < total; ++i)
        tracePhoton(cloud, cp, par, pm, seed, (uint32_t)i, (uint32_t)total, static_cast<Photon*>(photons), collisions, steps);
    return steps;
}
// One photon-estimate dispatch over rows y0 <= y < y1 (step row_step) of the image; accumulates into `image` like the shader
// (:541-545).  Returns the number of ray-march steps.
uint64_t orc_ppm_estimate(const float* cloud, const float* sv, const void* camera68, const void* cloud60, const void* params28,
                          const void* pmprops76, const void* svprops152, int32_t seed, uint32_t frame_count, float pm_radius,
                          const void* photons, const uint32_t* collisions, float* image, int y0, int y1, int row_step, int num_threads)
{
    const Camera cam = load_camera(camera68);
    const Cloud cp = load_cloud(cloud60);
    const Params par = load_params(params28);
    const PhotonMapProps pm = load_photon_map_props(pmprops76);
    const Shadow sp = load_shadow(svprops152);
    const int W = cam.halfWidth * 2, H = cam.halfHeight * 2;
    const uint32_t ngx = (uint32_t)(W / 32 + 1);
    const PhotonMapView map{static_cast<const Photon*>(photons), collisions,
                            (uint64_t)pm.voxelCount[0] * (uint64_t)pm.voxelCount[1] * (uint64_t)pm.voxelCount[2]};
    if (y1 > H) y1 = H;
    if (row_step < 1) row_step = 1;
    const int nrows = y1 > y0 ? (y1 - y0 + row_step - 1) / row_step : 0;
    uint64_t steps = 0;
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : steps) num_threads(num_threads > 0 ? num_threads : omp_get_max_threads())
    for (int r = 0; r < nrows; ++r) {
        const int y = y0 + r * row_step;
        for (int x = 0; x < W; ++x) {
            vec4 res = estimatePixel(cloud, sv, cam, cp, par, sp, pm, map, seed, pm_radius, (uint32_t)x, (uint32_t)y, ngx, steps);
            float* px = image + ((size_t)y * W + x) * 4;
            const vec4 old = {px[0], px[1], px[2], px[3]};
            res = res + old * (float)frame_count;
            const float d = (float)(frame_count + 1u);
            px[0] = res.x / d; px[1] = res.y / d; px[2] = res.z / d; px[3] = res.w / d;
        }
    }
    return steps;
}
void orc_blit_srgb8(const float* rgba, size_t npix, unsigned char* bgra8)
{
    auto oetf = [](float l) {
        const float c = !(l > 0.0f) ? 0.0f : (l > 1.0f ? 1.0f : l);       // clamp; NaN -> 0
        return c <= 0.0031308f ? c * 12.92f : 1.055f * powf(c, 1.0f / 2.4f) - 0.055f;
    };
    auto unorm8 = [](float v) {
        const float c = !(v > 0.0f) ? 0.0f : (v > 1.0f ? 1.0f : v);
        return (unsigned char)lrintf(c * 255.0f);                        // round to nearest even, like cvt.rni
    };
    for (size_t i = 0; i < npix; ++i) {
        bgra8[4 * i + 0] = unorm8(oetf(rgba[4 * i + 2]));
        bgra8[4 * i + 1] = unorm8(oetf(rgba[4 * i + 1]));
        bgra8[4 * i + 2] = unorm8(oetf(rgba[4 * i + 0]));
        bgra8[4 * i + 3] = unorm8(rgba[4 * i + 3]);
    }
}
void orc_background(const float dir[3], float rgba[4])
{
    const vec4 c = sampleBackground(vec3{dir[0], dir[1], dir[2]});
    rgba[0] = c.x; rgba[1] = c.y; rgba[2] = c.z; rgba[3] = c.w;
}
float orc_texture3d(const float* vol, uint32_t nx, uint32_t ny, uint32_t nz, const float uvw[3])
{
    return texture3d(vol, nx, ny, nz, vec3{uvw[0], uvw[1], uvw[2]});
}
float orc_sample_cloud(const float* vol, const void* cloud60, const float pos[3])
{
    return sampleCloud(vol, load_cloud(cloud60), vec3{pos[0], pos[1], pos[2]});
}
float orc_sample_shadow_volume(const float* sv, const void* svprops152, const float pos[3])
{
    return sampleShadowVolume(sv, load_shadow(svprops152), vec3{pos[0], pos[1], pos[2]});
}
void orc_scatter_ray(const void* params28, uint32_t* rng_state, const float incoming[3], float out_dir[3])
{
    Rng r;
    r.state = *rng_state;
    vec3 o{0, 0, 0};
    scatterRay(load_params(params28), r, vec3{incoming[0], incoming[1], incoming[2]}, o);
    *rng_state = r.state;
    out_dir[0] = o.x; out_dir[1] = o.y; out_dir[2] = o.z;
}

// ShadowVolume.comp:77-94 (helpers :34-56).  One (x,y) column per loop iteration, z serial.
void orc_shadow_volume(const float* cloud, const void* cloud60, const void* svprops152, float* out, int num_threads)
{
    const Cloud cp = load_cloud(cloud60);
    const Shadow sp = load_shadow(svprops152);
    const uint32_t N = sp.voxelAxisCount;
#ifdef _OPENMP
    if (num_threads <= 0) num_threads = omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 4) num_threads(num_threads)
#endif
    for (int64_t col = 0; col < static_cast<int64_t>(N) * N; ++col) {
        const uint32_t x = static_cast<uint32_t>(col % N), y = static_cast<uint32_t>(col / N);
        float accumulatedValue = 0.0f;          // .x of the vec4; the other channels never reach the r32f store
        float accumulatedDistance = 0.0f;
        for (uint32_t z = 0; z < N; ++z) {
            // :50-56  (b0 + (x*right + y*up + z*light) * voxelSize).xyz
            const vec4 p4 = sp.bounds[0] + ((static_cast<float>(x) * sp.right + static_cast<float>(y) * sp.up) +
                                            static_cast<float>(z) * sp.lightDirection) * sp.voxelSize;
            const vec3 p{p4.x, p4.y, p4.z};
            // :34-42 inclusive AABB
            const bool inside = !(p.x < cp.bounds[0].x || p.y < cp.bounds[0].y || p.z < cp.bounds[0].z ||
                                  p.x > cp.bounds[1].x || p.y > cp.bounds[1].y || p.z > cp.bounds[1].z);
            if (inside) accumulatedValue = accumulatedValue + (sampleCloud(cloud, cp, p) * cp.densityScaling) / cp.baseScaling;
            accumulatedDistance = accumulatedDistance + sp.voxelSize;
            const float e = ((-accumulatedDistance) * accumulatedValue) / static_cast<float>(z + 1u);
            out[x + static_cast<size_t>(N) * (y + static_cast<size_t>(N) * z)] = std::exp(e);
        }
    }
}

// the same for selected columns only (large shadow volumes in the tests): out[(iy * nxs + ix) * N + z] = T(xs[ix], ys[iy], z)
void orc_shadow_volume_columns(const float* cloud, const void* cloud60, const void* svprops152, const uint32_t* xs, uint32_t nxs,
                               const uint32_t* ys, uint32_t nys, float* out, int num_threads)
{
    const Cloud cp = load_cloud(cloud60);
    const Shadow sp = load_shadow(svprops152);
    const uint32_t N = sp.voxelAxisCount;
#ifdef _OPENMP
    if (num_threads <= 0) num_threads = omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 4) num_threads(num_threads)
#endif
    for (int64_t col = 0; col < static_cast<int64_t>(nxs) * nys; ++col) {
        const uint32_t x = xs[col % nxs], y = ys[col / nxs];
        float accumulatedValue = 0.0f;
        float accumulatedDistance = 0.0f;
        for (uint32_t z = 0; z < N; ++z) {
            const vec4 p4 = sp.bounds[0] + ((static_cast<float>(x) * sp.right + static_cast<float>(y) * sp.up) +
                                            static_cast<float>(z) * sp.lightDirection) * sp.voxelSize;
            const vec3 p{p4.x, p4.y, p4.z};
            const bool inside = !(p.x < cp.bounds[0].x || p.y < cp.bounds[0].y || p.z < cp.bounds[0].z ||
                                  p.x > cp.bounds[1].x || p.y > cp.bounds[1].y || p.z > cp.bounds[1].z);
            if (inside) accumulatedValue = accumulatedValue + (sampleCloud(cloud, cp, p) * cp.densityScaling) / cp.baseScaling;
            accumulatedDistance = accumulatedDistance + sp.voxelSize;
            const float e = ((-accumulatedDistance) * accumulatedValue) / static_cast<float>(z + 1u);
            out[static_cast<size_t>(col) * N + z] = std::exp(e);
        }
    }
}

void orc_path_trace_frame(const float* cloud, const float* sv, const void* camera68, const void* cloud60,
                          const void* params28, const void* svprops152, const void* push24, float absorption,
                          int accum_sum, float* image, float* sample, uint32_t* steps, uint32_t* collisions,
                          int num_threads)
{
    trace_rect(cloud, sv, camera68, cloud60, params28, svprops152, push24, absorption, accum_sum, 0, 0, 1 << 30, 1 << 30,
               image, sample, steps, collisions, num_threads);
}

void orc_path_trace_rect(const float* cloud, const float* sv, const void* camera68, const void* cloud60,
                         const void* params28, const void* svprops152, const void* push24, float absorption,
                         int accum_sum, int x0, int y0, int x1, int y1, float* image, float* sample,
                         uint32_t* steps, uint32_t* collisions, int num_threads)
{
    trace_rect(cloud, sv, camera68, cloud60, params28, svprops152, push24, absorption, accum_sum, x0, y0, x1, y1, image,
               sample, steps, collisions, num_threads);
}

// every row_step-th row of the image (bounded, unbiased sample of a frame for the CPU baseline)
void orc_path_trace_rows(const float* cloud, const float* sv, const void* camera68, const void* cloud60,
                         const void* params28, const void* svprops152, const void* push24, float absorption,
                         int row_step, float* image, float* sample, uint32_t* steps, uint32_t* collisions, int num_threads)
{
    trace_rect(cloud, sv, camera68, cloud60, params28, svprops152, push24, absorption, 0, 0, 0, 1 << 30, 1 << 30, image,
               sample, steps, collisions, num_threads, row_step);
}

void orc_fbm_generate(uint32_t n, uint32_t seed, float* out, int num_threads)
{
#ifdef _OPENMP
    if (num_threads <= 0) num_threads = omp_get_max_threads();
#pragma omp parallel for schedule(static) num_threads(num_threads)
#endif
    for (int64_t k = 0; k < static_cast<int64_t>(n); ++k)
        for (uint32_t j = 0; j < n; ++j)
            for (uint32_t i = 0; i < n; ++i)
                out[i + static_cast<size_t>(n) * (j + static_cast<size_t>(n) * k)] = fbm_density(i, j, (uint32_t)k, n, seed);
}

int orc_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

}  // extern "C"

// ---- host-side scene setup that lives in main.cpp rather than in the headers (restated for the reference arm) ----
extern "C" {

// SetCloudProperties (main.cpp:144-163) on a default CloudProperties (UniformBuffers.h:71-78) with the given
// densityScaling: size = voxelSize * count * baseScaling (double, narrowed to float), box centred in x/y on z = 0.
void orc_set_cloud_properties(const uint32_t counts[3], const double voxel_size[3], float majorant, float density_scaling,
                              void* out60)
{
    const float base_scaling = 1000.0f;
    float size[3];
    for (int i = 0; i < 3; ++i) size[i] = static_cast<float>(voxel_size[i] * counts[i] * base_scaling);
    float b0[4] = {-size[0] / 2, -size[1] / 2, 0.0f, 0.0f};
    const float add[4] = {0.0f, 0.0f, size[2], 0.0f};
    float b1[4];
    for (int i = 0; i < 4; ++i) b1[i] = -b0[i] + add[i];
    const uint32_t vc[4] = {counts[0], counts[1], counts[2], 0u};
    const float max_ext = majorant > 0.01f ? majorant : 0.01f;
    char* o = static_cast<char*>(out60);
    std::memcpy(o, b0, 16);
    std::memcpy(o + 16, b1, 16);
    std::memcpy(o + 32, vc, 16);
    std::memcpy(o + 48, &max_ext, 4);
    std::memcpy(o + 52, &base_scaling, 4);
    std::memcpy(o + 56, &density_scaling, 4);
}

// std::rand() of the MSVC runtime after srand(seed0): the per-frame seed source (main.cpp:722,949)
void orc_msvc_rand(uint32_t seed0, uint32_t n, int32_t* out)
{
    uint32_t s = seed0;
    for (uint32_t i = 0; i < n; ++i) {
        s = s * 214013u + 2531011u;
        out[i] = static_cast<int32_t>((s >> 16) & 0x7fffu);
    }
}

}  // extern "C"
