// crv/uniform_buffers.h -- host parameter blocks of the PathTracer/ShadowVolume hot path.
//
// C++ mirrors of UniformBuffers.h:6-167 (reference) with the same member-function names and the
// same results BYTE FOR BYTE (tests/test_structs.py checks them against the reference's own
// headers compiled in oracle/_ref, and against committed golden bytes).  No glm: the few vector
// operations the setters need are spelled out in fp32 in the evaluation order glm uses, because the
// bytes of these structs are what the kernels consume.
//
// Each struct derives from the plain-C struct of crv_c_api.h and adds no data, so a
// crv::CameraProperties* can be handed straight to the C ABI.
//
// Compile host code that includes this header with -ffp-contract=off (the Makefile does).
#pragma once

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>

#include "../crv_c_api.h"

namespace crv {
namespace detail {

struct V3 { float x, y, z; };
struct V4 { float x, y, z, w; };
struct M4 { V4 c[4]; };   // column-major like glm::mat4

inline V4 mul(const V4& a, float s) { return V4{a.x * s, a.y * s, a.z * s, a.w * s}; }
inline V4 add(const V4& a, const V4& b) { return V4{a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w}; }
inline V4 sub(const V4& a, const V4& b) { return V4{a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w}; }
inline V4 mulv(const V4& a, const V4& b) { return V4{a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w}; }
inline V4 neg(const V4& a) { return V4{-a.x, -a.y, -a.z, -a.w}; }

// glm::radians: degrees * 0.01745329251994329576923690768489 (rounded to float first)
inline float radians(float deg) { return deg * 0.01745329251994329576923690768489f; }

// glm::dot(vec3): (x*x + y*y) + z*z ; glm::dot(vec4): (x*x + y*y) + (z*z + w*w)
inline float dot3(const V3& a, const V3& b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
inline float dot4(const V4& a, const V4& b) { return (a.x * b.x + a.y * b.y) + (a.z * b.z + a.w * b.w); }

// glm::normalize(v) = v * inversesqrt(dot(v,v)), inversesqrt(x) = 1/sqrt(x)
inline V3 normalize3(const V3& v)
{
    const float inv = 1.0f / std::sqrt(dot3(v, v));
    return V3{v.x * inv, v.y * inv, v.z * inv};
}
// glm::distance(a,b) = length(b - a) = sqrt(dot(d,d))
inline float distance3(const V3& a, const V3& b)
{
    const V3 d{b.x - a.x, b.y - a.y, b.z - a.z};
    return std::sqrt(dot3(d, d));
}
inline float distance4(const V4& a, const V4& b)
{
    const V4 d = sub(b, a);
    return std::sqrt(dot4(d, d));
}

// utilities::GetOrthonormalBasis (Utilities.cpp:7-16); same frame as PathTracer.comp:197-206
inline void orthonormal_basis(const V3& n, V3& right, V3& up)
{
    const float sz = (n.z >= 0.0f) ? 1.0f : -1.0f;
    const float a = n.y / (1.0f + std::fabs(n.z));
    const float b = n.y * a;
    const float c = (-n.x) * a;
    right = V3{n.z + sz * b, sz * c, -n.x};
    up = V3{c, 1.0f - b, (-sz) * n.y};
}

// glm::rotate(angle, axis) applied to the identity (gtx/transform.inl:12-15 -> ext/matrix_transform.inl:18-46)
inline M4 rotate_identity(float angle, const V3& axis_in)
{
    const float c = std::cos(angle);
    const float s = std::sin(angle);
    const V3 axis = normalize3(axis_in);
    const V3 t{(1.0f - c) * axis.x, (1.0f - c) * axis.y, (1.0f - c) * axis.z};
    float r[3][3];
    r[0][0] = c + t.x * axis.x;
    r[0][1] = t.x * axis.y + s * axis.z;
    r[0][2] = t.x * axis.z - s * axis.y;
    r[1][0] = t.y * axis.x - s * axis.z;
    r[1][1] = c + t.y * axis.y;
    r[1][2] = t.y * axis.z + s * axis.x;
    r[2][0] = t.z * axis.x + s * axis.y;
    r[2][1] = t.z * axis.y - s * axis.x;
    r[2][2] = c + t.z * axis.z;
    const V4 e0{1, 0, 0, 0}, e1{0, 1, 0, 0}, e2{0, 0, 1, 0}, e3{0, 0, 0, 1};
    M4 out;
    for (int j = 0; j < 3; ++j) out.c[j] = add(add(mul(e0, r[j][0]), mul(e1, r[j][1])), mul(e2, r[j][2]));
    out.c[3] = e3;
    return out;
}

// glm mat4*mat4: Result[j] = ((A0*B[j][0] + A1*B[j][1]) + A2*B[j][2]) + A3*B[j][3]
inline M4 matmul(const M4& a, const M4& b)
{
    M4 out;
    for (int j = 0; j < 4; ++j) {
        const V4& bj = b.c[j];
        out.c[j] = add(add(add(mul(a.c[0], bj.x), mul(a.c[1], bj.y)), mul(a.c[2], bj.z)), mul(a.c[3], bj.w));
    }
    return out;
}
// glm mat4*vec4: (m0*v0 + m1*v1) + (m2*v2 + m3*v3)
inline V4 matvec(const M4& m, const V4& v)
{
    return add(add(mul(m.c[0], v.x), mul(m.c[1], v.y)), add(mul(m.c[2], v.z), mul(m.c[3], v.w)));
}

// glm::inverse(mat4): cofactor expansion (detail/func_matrix.inl compute_inverse<4,4>); the grouping of
// every product/difference follows glm so the bytes agree.
inline M4 inverse(const M4& mm)
{
    // m[c][r] accessor in glm's index order
    auto m = [&](int c, int r) -> float { return (&mm.c[c].x)[r]; };
    // 2x2 sub-determinants of rows (ra,rb) taken from column pairs (2,3), (1,3), (1,2)
    auto sub2 = [&](int ca, int cb, int ra, int rb) -> float { return m(ca, ra) * m(cb, rb) - m(cb, ra) * m(ca, rb); };
    float f[6][3];
    const int rows[6][2] = {{2, 3}, {1, 3}, {1, 2}, {0, 3}, {0, 2}, {0, 1}};
    for (int k = 0; k < 6; ++k) {
        f[k][0] = sub2(2, 3, rows[k][0], rows[k][1]);
        f[k][1] = sub2(1, 3, rows[k][0], rows[k][1]);
        f[k][2] = sub2(1, 2, rows[k][0], rows[k][1]);
    }
    V4 fac[6];
    for (int k = 0; k < 6; ++k) fac[k] = V4{f[k][0], f[k][0], f[k][1], f[k][2]};
    V4 vec[4];
    for (int r = 0; r < 4; ++r) vec[r] = V4{m(1, r), m(0, r), m(0, r), m(0, r)};
    const V4 inv0 = add(sub(mulv(vec[1], fac[0]), mulv(vec[2], fac[1])), mulv(vec[3], fac[2]));
    const V4 inv1 = add(sub(mulv(vec[0], fac[0]), mulv(vec[2], fac[3])), mulv(vec[3], fac[4]));
    const V4 inv2 = add(sub(mulv(vec[0], fac[1]), mulv(vec[1], fac[3])), mulv(vec[3], fac[5]));
    const V4 inv3 = add(sub(mulv(vec[0], fac[2]), mulv(vec[1], fac[4])), mulv(vec[2], fac[5]));
    const V4 sa{+1, -1, +1, -1}, sb{-1, +1, -1, +1};
    M4 inv;
    inv.c[0] = mulv(inv0, sa);
    inv.c[1] = mulv(inv1, sb);
    inv.c[2] = mulv(inv2, sa);
    inv.c[3] = mulv(inv3, sb);
    const V4 row0{inv.c[0].x, inv.c[1].x, inv.c[2].x, inv.c[3].x};
    const V4 d = mulv(mm.c[0], row0);
    const float det = (d.x + d.y) + (d.z + d.w);
    const float inv_det = 1.0f / det;
    for (int j = 0; j < 4; ++j) inv.c[j] = mul(inv.c[j], inv_det);
    return inv;
}

}  // namespace detail

// ---------------------------------------------------------------------------------------------
// UniformBuffers.h:6-13
struct PushConstants : crv_push_constants {
    PushConstants()
    {
        time = 0;
        seed = 100;
        frameCount = 1;
        pmRadius = 0;
        currentBuffer = 0;
    }
};
static_assert(sizeof(PushConstants) == 24, "PushConstants must be 24 bytes (UniformBuffers.h:6-13)");
static_assert(offsetof(crv_push_constants, seed) == 8 && offsetof(crv_push_constants, frameCount) == 12 &&
                  offsetof(crv_push_constants, pmRadius) == 16,
              "PushConstants offsets");

// UniformBuffers.h:15-69
struct CameraProperties : crv_camera_properties {
    CameraProperties()
    {
        position[0] = 0; position[1] = 0; position[2] = -800;
        halfWidth = 800 / 2;
        forward[0] = 0; forward[1] = 0; forward[2] = 1;
        halfHeight = 600 / 2;
        right[0] = 1; right[1] = 0; right[2] = 0;
        nearPlane = 50.0f;
        up[0] = 0; up[1] = 1; up[2] = 0;
        pixelSizeY = (2 * std::tan(detail::radians(120 / 2.f)) * nearPlane) / (halfHeight * 2);
        pixelSizeX = pixelSizeY * static_cast<float>(halfHeight) / static_cast<float>(halfWidth);
    }

    // :29-35  pitch clamped to +-80 deg, forward = Rx(pitch) * Ry(yaw) * (0,0,1,0), basis from forward
    void SetRotation(float rotation[2])
    {
        rotation[0] = std::clamp(rotation[0], -80.f, 80.f);
        const detail::M4 rx = detail::rotate_identity(detail::radians(rotation[0]), detail::V3{1, 0, 0});
        const detail::M4 ry = detail::rotate_identity(detail::radians(rotation[1]), detail::V3{0, 1, 0});
        const detail::V4 f = detail::matvec(detail::matmul(rx, ry), detail::V4{0, 0, 1, 0});
        forward[0] = f.x; forward[1] = f.y; forward[2] = f.z;
        detail::V3 r, u;
        detail::orthonormal_basis(detail::V3{f.x, f.y, f.z}, r, u);
        right[0] = r.x; right[1] = r.y; right[2] = r.z;
        up[0] = u.x; up[1] = u.y; up[2] = u.z;
    }
    int GetWidth() const { return halfWidth * 2; }
    int GetHeight() const { return halfHeight * 2; }
    void SetResolution(int newWidth, int newHeight)
    {
        halfWidth = newWidth / 2;
        halfHeight = newHeight / 2;
    }
    // :52-58  clamp 60..150 deg; the image plane is square in world units whatever the aspect (quirk kept)
    void SetFOV(float& fov)
    {
        fov = std::clamp(fov, 60.f, 150.f);
        const float height = 2 * std::tan(detail::radians(fov / 2.f)) * nearPlane;
        pixelSizeY = height / (halfHeight * 2);
        pixelSizeX = pixelSizeY * static_cast<float>(halfHeight) / static_cast<float>(halfWidth);
    }
    float GetNearPlane() const { return nearPlane; }
    const float* GetFwd() const { return forward; }
};
static_assert(sizeof(CameraProperties) == 68, "CameraProperties must be 68 bytes (UniformBuffers.h:15-69)");

// UniformBuffers.h:71-78 (+ SetCloudProperties, main.cpp:144-163)
struct CloudProperties : crv_cloud_properties {
    CloudProperties()
    {
        std::memset(bounds, 0, sizeof(bounds));
        std::memset(voxelCount, 0, sizeof(voxelCount));
        maxExtinction = 0.2f;
        baseScaling = 1000.f;
        densityScaling = 200.f;
    }
    // main.cpp:144-163: size = voxelSize * count * baseScaling (double, then narrowed to float),
    // box centred in x/y, standing on z = 0; maxExtinction = max(majorant, 0.01)
    void SetFromGrid(const uint32_t counts[3], const double voxel_size[3], float majorant)
    {
        float size[3];
        for (int i = 0; i < 3; ++i) size[i] = static_cast<float>(voxel_size[i] * counts[i] * baseScaling);
        maxExtinction = std::max(majorant, 0.01f);
        voxelCount[0] = counts[0]; voxelCount[1] = counts[1]; voxelCount[2] = counts[2]; voxelCount[3] = 0;
        bounds[0][0] = -size[0] / 2;
        bounds[0][1] = -size[1] / 2;
        bounds[0][2] = 0;
        bounds[0][3] = 0;
        const float add[4] = {0, 0, size[2], 0};
        for (int i = 0; i < 4; ++i) bounds[1][i] = -bounds[0][i] + add[i];
    }
};
static_assert(sizeof(CloudProperties) == 60, "CloudProperties must be 60 bytes (UniformBuffers.h:71-78)");
static_assert(offsetof(crv_cloud_properties, voxelCount) == 32 && offsetof(crv_cloud_properties, maxExtinction) == 48 &&
                  offsetof(crv_cloud_properties, densityScaling) == 56,
              "CloudProperties offsets");

// UniformBuffers.h:80-112
struct Parameters : crv_parameters {
    Parameters()
    {
        maxRayBounces = 5;
        lightIntensity = 5;
        phaseG = 0.00000001f;
        phaseOnePlusG2 = 1.0f + phaseG * phaseG;
        phaseOneMinusG2 = 1.0f - phaseG * phaseG;
        phaseOneOver2G = 0.5f / phaseG;
        isotropic = std::fabs(phaseG) < 0.0001f;
        _pad[0] = _pad[1] = _pad[2] = 0;
    }
    // :93-106  reject |g| > 1, clamp to +-0.95, derived constants only refreshed when anisotropic
    void SetPhaseG(float value)
    {
        if (value > 1 || value < -1) return;
        phaseG = std::min(std::max(value, -0.95f), 0.95f);
        isotropic = (std::fabs(phaseG) < 0.0001f);
        if (!isotropic) {
            phaseOnePlusG2 = 1.0f + phaseG * phaseG;
            phaseOneMinusG2 = 1.0f - phaseG * phaseG;
            phaseOneOver2G = 0.5f / phaseG;
        }
    }
    float GetPhaseG() const { return phaseG; }
};
static_assert(sizeof(Parameters) == 28, "Parameters must be 28 bytes (UniformBuffers.h:80-112)");

// UniformBuffers.h:114-167
struct ShadowVolumeProperties : crv_shadow_volume_properties {
    ShadowVolumeProperties()
    {
        std::memset(bounds, 0, sizeof(bounds));
        const float l[4] = {1, -1, 0, 0}, r[4] = {1, 0, 0, 0}, u[4] = {0, 1, 0, 0};
        std::memcpy(lightDirection, l, 16);
        std::memcpy(right, r, 16);
        std::memcpy(up, u, 16);
        std::memcpy(basisChange[0], l, 16);
        std::memcpy(basisChange[1], r, 16);
        std::memcpy(basisChange[2], u, 16);
        std::memset(basisChange[3], 0, 16);
        voxelAxisCount = 500;
        voxelSize = 1.0f;
    }
    // :131-143  normalise, tangent frame, then UpdateOrigin with the *current* bounds (its centre
    // argument (b0-b1)/2 is the reference's; SetOrigin always follows, main.cpp:232-235)
    void SetLightDirection(const float dir[3])
    {
        using namespace detail;
        const V3 n = normalize3(V3{dir[0], dir[1], dir[2]});
        V3 r, u;
        orthonormal_basis(n, r, u);
        const float l4[4] = {n.x, n.y, n.z, 0}, r4[4] = {r.x, r.y, r.z, 0}, u4[4] = {u.x, u.y, u.z, 0};
        std::memcpy(lightDirection, l4, 16);
        std::memcpy(right, r4, 16);
        std::memcpy(up, u4, 16);
        const V4 b0{bounds[0][0], bounds[0][1], bounds[0][2], bounds[0][3]};
        const V4 b1{bounds[1][0], bounds[1][1], bounds[1][2], bounds[1][3]};
        const V4 half = V4{(b0.x - b1.x) / 2.f, (b0.y - b1.y) / 2.f, (b0.z - b1.z) / 2.f, (b0.w - b1.w) / 2.f};
        UpdateOrigin(distance4(b0, b1) / 2.f, V3{half.x, half.y, half.z});
    }
    // :145-151  bounding sphere of the cloud box; samples are points on the grid: voxelSize = 2R/(N-1)
    void SetOrigin(const float lower[3], const float upper[3])
    {
        using namespace detail;
        const V3 lo{lower[0], lower[1], lower[2]}, hi{upper[0], upper[1], upper[2]};
        const float radius = distance3(lo, hi) / 2.f;
        voxelSize = (2 * radius) / (voxelAxisCount - 1);
        UpdateOrigin(radius, V3{(lo.x + hi.x) / 2.f, (lo.y + hi.y) / 2.f, (lo.z + hi.z) / 2.f});
    }
    const float* GetLightDirection() const { return lightDirection; }

private:
    // :159-166
    void UpdateOrigin(float radius, const detail::V3& center)
    {
        using namespace detail;
        const V4 l{lightDirection[0], lightDirection[1], lightDirection[2], lightDirection[3]};
        const V4 r{right[0], right[1], right[2], right[3]};
        const V4 u{up[0], up[1], up[2], up[3]};
        const V4 corner = mul(sub(sub(neg(l), r), u), radius);
        const V4 c4{center.x, center.y, center.z, 1};
        const V4 b0 = add(c4, corner), b1 = sub(c4, corner);
        std::memcpy(bounds[0], &b0, 16);
        std::memcpy(bounds[1], &b1, 16);
        M4 basis;
        basis.c[0] = r; basis.c[1] = u; basis.c[2] = l; basis.c[3] = V4{0, 0, 0, 1};
        const M4 inv = inverse(basis);
        std::memcpy(basisChange, &inv, 64);
    }
};
static_assert(sizeof(ShadowVolumeProperties) == 152, "ShadowVolumeProperties must be 152 bytes (UniformBuffers.h:114-167)");
static_assert(offsetof(crv_shadow_volume_properties, voxelAxisCount) == 144 &&
                  offsetof(crv_shadow_volume_properties, voxelSize) == 148,
              "ShadowVolumeProperties offsets");

// UniformBuffers.h:203-210
struct Photon : crv_photon {};
static_assert(sizeof(Photon) == 48, "Photon must match the std430 record of PPM_PT.comp:14-21");

// UniformBuffers.h:212-249.  The setters are implemented once, behind the C ABI (host_helpers.cpp), and pinned against
// the reference's own struct by tests/test_ppm.py.
struct PhotonMapProperties : crv_photon_map_properties {
    PhotonMapProperties() { crv_photon_map_init(this); }
    void SetBounds(const float bounds[2][4]) { crv_photon_map_set_bounds(this, bounds); }
    uint32_t GetTotalSize() const { return crv_photon_map_total_size(this); }
};
static_assert(sizeof(PhotonMapProperties) == 76, "PhotonMapProperties must match the std140 block of PPM_PT.comp:57-67");

// Extension block (not in the reference): see crv_c_api.h
struct PathTracerExtensions : crv_pt_extensions {
    PathTracerExtensions()
    {
        absorption = 0.0f;
        math_mode = CRV_PT_MATH_STRICT;
        reserved[0] = reserved[1] = 0;
    }
};

// MSVC rand(): s = s*214013 + 2531011; return (s >> 16) & 0x7fff  -- the per-frame seed source
// (main.cpp:722 g_pushConstants.seed = std::rand(), after srand(0) at main.cpp:949)
inline void MsvcRandSequence(uint32_t seed0, unsigned n, int32_t* out)
{
    uint32_t s = seed0;
    for (unsigned i = 0; i < n; ++i) {
        s = s * 214013u + 2531011u;
        out[i] = static_cast<int32_t>((s >> 16) & 0x7fffu);
    }
}

}  // namespace crv
