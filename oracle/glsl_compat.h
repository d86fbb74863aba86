// oracle/glsl_compat.h -- a minimal GLSL 4.50 compute-shader environment in C++.  TEST INFRASTRUCTURE ONLY.
//
// Purpose: let the reference's OWN shader sources (/root/reference/shaders/PathTracer.comp, ShadowVolume.comp, PPM_PT.comp,
// PPM_PE.comp) be
// compiled by g++ WHERE THEY LIE and executed on the CPU, so that the hand-written restatement in crv_oracle.cpp
// (and through it the CUDA kernels) is pinned against the reference's real code instead of against a reading of it.
// oracle/glsl_to_cpp.py does the purely lexical part (layout qualifiers, in/out/inout, float-literal suffixes,
// swizzle calls); everything the shader text *means* is supplied here.
//
// What GLSL / Vulkan leave to the driver is fixed exactly as crv_oracle.h states it (this is the one thing a real
// driver could do differently, and it is not pinned by the reference): no FMA contraction (build flag), correctly
// rounded + - * / sqrt, glibc libm for log/exp/pow/sin/cos, dot products and matrix*vector summed left to right,
// normalize(v) = v / sqrt(dot(v,v)), mix(a,b,t) = a*(1-t) + b*t, smoothstep = Hermite of the clamped ratio,
// lowp evaluated as fp32, LINEAR / CLAMP_TO_EDGE filtering by the Vulkan "texel filtering" rules
// (Initializers.cpp:389-418: unnormalizedCoordinates = false, one mip level).
#pragma once
#include <cmath>
#include <cstddef>
#include <cstdint>

namespace glsl {

typedef uint32_t uint;

// a GLSL bool inside a uniform block occupies 4 bytes (std140)
struct bool32 {
    uint32_t v;
    operator bool() const { return v != 0u; }
};

// ---------------------------------------------------------------------------------------------------------
// vectors.  Arithmetic operators are hidden friends (non-template per instantiation), so the scalar operand
// takes GLSL's implicit int -> uint -> float conversions.
// ---------------------------------------------------------------------------------------------------------
template <class T> struct tvec2 {
    T x, y;
    tvec2() : x(0), y(0) {}
    explicit tvec2(T s) : x(s), y(s) {}
    tvec2(T a, T b) : x(a), y(b) {}
    template <class U> explicit tvec2(const tvec2<U>& o) : x((T)o.x), y((T)o.y) {}
    T& operator[](int i) { return (&x)[i]; }
    const T& operator[](int i) const { return (&x)[i]; }
};

template <class T> struct tvec3 {
    T x, y, z;
    tvec3() : x(0), y(0), z(0) {}
    explicit tvec3(T s) : x(s), y(s), z(s) {}
    tvec3(T a, T b, T c) : x(a), y(b), z(c) {}
    template <class U> explicit tvec3(const tvec3<U>& o) : x((T)o.x), y((T)o.y), z((T)o.z) {}
    template <class U, class V> tvec3(const tvec2<U>& a, V c) : x((T)a.x), y((T)a.y), z((T)c) {}
    T& operator[](int i) { return (&x)[i]; }
    const T& operator[](int i) const { return (&x)[i]; }
    tvec2<T> xy() const { return tvec2<T>(x, y); }
    tvec3 zxy() const { return tvec3(z, x, y); }
    tvec3 yzx() const { return tvec3(y, z, x); }

    friend tvec3 operator+(const tvec3& a, const tvec3& b) { return tvec3(a.x + b.x, a.y + b.y, a.z + b.z); }
    friend tvec3 operator-(const tvec3& a, const tvec3& b) { return tvec3(a.x - b.x, a.y - b.y, a.z - b.z); }
    friend tvec3 operator*(const tvec3& a, const tvec3& b) { return tvec3(a.x * b.x, a.y * b.y, a.z * b.z); }
    friend tvec3 operator/(const tvec3& a, const tvec3& b) { return tvec3(a.x / b.x, a.y / b.y, a.z / b.z); }
    friend tvec3 operator*(const tvec3& a, T s) { return tvec3(a.x * s, a.y * s, a.z * s); }
    friend tvec3 operator*(T s, const tvec3& a) { return tvec3(s * a.x, s * a.y, s * a.z); }
    friend tvec3 operator/(const tvec3& a, T s) { return tvec3(a.x / s, a.y / s, a.z / s); }
    friend tvec3 operator/(T s, const tvec3& a) { return tvec3(s / a.x, s / a.y, s / a.z); }
    friend tvec3 operator+(const tvec3& a, T s) { return tvec3(a.x + s, a.y + s, a.z + s); }
    friend tvec3 operator-(const tvec3& a, T s) { return tvec3(a.x - s, a.y - s, a.z - s); }
    friend tvec3 operator-(const tvec3& a) { return tvec3(-a.x, -a.y, -a.z); }
    tvec3& operator+=(const tvec3& b) { return *this = *this + b; }
    tvec3& operator-=(const tvec3& b) { return *this = *this - b; }
    tvec3& operator*=(T s) { return *this = *this * s; }
    tvec3& operator/=(T s) { return *this = *this / s; }
};

template <class T> struct tvec4 {
    T x, y, z, w;
    tvec4() : x(0), y(0), z(0), w(0) {}
    explicit tvec4(T s) : x(s), y(s), z(s), w(s) {}
    tvec4(T a, T b, T c, T d) : x(a), y(b), z(c), w(d) {}
    template <class U> explicit tvec4(const tvec4<U>& o) : x((T)o.x), y((T)o.y), z((T)o.z), w((T)o.w) {}
    template <class U, class V> tvec4(const tvec3<U>& a, V d) : x((T)a.x), y((T)a.y), z((T)a.z), w((T)d) {}
    T& operator[](int i) { return (&x)[i]; }
    const T& operator[](int i) const { return (&x)[i]; }
    tvec3<T> xyz() const { return tvec3<T>(x, y, z); }
    tvec3<T> zxy() const { return tvec3<T>(z, x, y); }
    tvec3<T> yzx() const { return tvec3<T>(y, z, x); }
    tvec2<T> xy() const { return tvec2<T>(x, y); }

    friend tvec4 operator+(const tvec4& a, const tvec4& b) { return tvec4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
    friend tvec4 operator-(const tvec4& a, const tvec4& b) { return tvec4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }
    friend tvec4 operator*(const tvec4& a, const tvec4& b) { return tvec4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
    friend tvec4 operator/(const tvec4& a, const tvec4& b) { return tvec4(a.x / b.x, a.y / b.y, a.z / b.z, a.w / b.w); }
    friend tvec4 operator*(const tvec4& a, T s) { return tvec4(a.x * s, a.y * s, a.z * s, a.w * s); }
    friend tvec4 operator*(T s, const tvec4& a) { return tvec4(s * a.x, s * a.y, s * a.z, s * a.w); }
    friend tvec4 operator/(const tvec4& a, T s) { return tvec4(a.x / s, a.y / s, a.z / s, a.w / s); }
    friend tvec4 operator+(const tvec4& a, T s) { return tvec4(a.x + s, a.y + s, a.z + s, a.w + s); }
    friend tvec4 operator-(const tvec4& a, T s) { return tvec4(a.x - s, a.y - s, a.z - s, a.w - s); }
    friend tvec4 operator-(const tvec4& a) { return tvec4(-a.x, -a.y, -a.z, -a.w); }
    tvec4& operator+=(const tvec4& b) { return *this = *this + b; }
    tvec4& operator-=(const tvec4& b) { return *this = *this - b; }
    tvec4& operator*=(T s) { return *this = *this * s; }
    tvec4& operator/=(T s) { return *this = *this / s; }
};

typedef tvec2<float> vec2;
typedef tvec3<float> vec3;
typedef tvec4<float> vec4;
typedef tvec2<int32_t> ivec2;
typedef tvec3<int32_t> ivec3;
typedef tvec4<int32_t> ivec4;
typedef tvec2<uint32_t> uvec2;
typedef tvec3<uint32_t> uvec3;
typedef tvec4<uint32_t> uvec4;

// column-major 4x4; M * v = ((c0*v.x + c1*v.y) + c2*v.z) + c3*v.w
struct mat4 {
    vec4 c[4];
    friend vec4 operator*(const mat4& m, const vec4& v) { return ((m.c[0] * v.x + m.c[1] * v.y) + m.c[2] * v.z) + m.c[3] * v.w; }
};

// ---------------------------------------------------------------------------------------------------------
// built-in functions (GLSL 4.50 section 8), scalar float and the vector forms the two shaders use
// ---------------------------------------------------------------------------------------------------------
struct Counters {
    uint64_t log_calls;          // one per delta-tracking iteration (PathTracer.comp:338)
    uint64_t texture_calls[2];   // per sampler slot: 0 = cloud, 1 = shadow volume
};
inline Counters& counters()
{
    static thread_local Counters c = {0, {0, 0}};
    return c;
}

inline float abs(float a) { return std::fabs(a); }
inline float sqrt(float a) { return std::sqrt(a); }
inline float exp(float a) { return std::exp(a); }
inline float log(float a)
{
    counters().log_calls++;
    return std::log(a);
}
inline float pow(float a, float b) { return std::pow(a, b); }
inline float sin(float a) { return std::sin(a); }
inline float cos(float a) { return std::cos(a); }
inline float tan(float a) { return std::tan(a); }
inline float acos(float a) { return std::acos(a); }
inline float atan(float y, float x) { return std::atan2(y, x); }
inline float floor(float a) { return std::floor(a); }
inline float min(float a, float b) { return b < a ? b : a; }        // "y if y < x, otherwise x"
inline float max(float a, float b) { return a < b ? b : a; }        // "y if x < y, otherwise x"
inline float clamp(float x, float lo, float hi) { return min(max(x, lo), hi); }
inline float mix(float a, float b, float t) { return a * (1.0f - t) + b * t; }
inline vec4 mix(const vec4& a, const vec4& b, float t) { return a * (1.0f - t) + b * t; }
inline float smoothstep(float e0, float e1, float x)
{
    const float t = clamp((x - e0) / (e1 - e0), 0.0f, 1.0f);
    return t * t * (3.0f - 2.0f * t);
}
inline float dot(const vec3& a, const vec3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline float dot(const vec4& a, const vec4& b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }
inline float length(const vec3& a) { return std::sqrt(dot(a, a)); }
inline float distance(const vec3& a, const vec3& b) { return length(a - b); }
inline vec3 normalize(const vec3& a) { return a / std::sqrt(dot(a, a)); }
inline vec3 abs(const vec3& a) { return vec3(std::fabs(a.x), std::fabs(a.y), std::fabs(a.z)); }
// storage-buffer atomics (PPM_PT.comp:316): the CPU dispatch runs invocations on several threads
inline uint atomicAdd(uint& mem, uint v)
{
    uint old;
#pragma omp atomic capture
    { old = mem; mem += v; }
    return old;
}
inline vec4 exp(const vec4& a) { return vec4(std::exp(a.x), std::exp(a.y), std::exp(a.z), std::exp(a.w)); }

// ---------------------------------------------------------------------------------------------------------
// opaque types
// ---------------------------------------------------------------------------------------------------------
// sampler3D over an R32F image stored x-fastest (Utilities.cpp:94-111 uploads Grid3D tightly packed)
struct sampler3D {
    const float* texels = nullptr;
    int nx = 0, ny = 0, nz = 0;
    int slot = 0;
};
struct image2D {          // rgba32f
    float* texels = nullptr;
    int width = 0, height = 0;
};
struct image3D {          // r32f, writeonly
    float* texels = nullptr;
    int nx = 0, ny = 0, nz = 0;
};

namespace detail {
inline int clampi(int v, int hi) { return v < 0 ? 0 : (v > hi ? hi : v); }
inline int to_int_sat(float f)
{
    if (!(f == f)) return 0;
    if (f < -2147483000.0f) return -2147483647 - 1;
    if (f > 2147483000.0f) return 2147483647;
    return static_cast<int>(f);
}
}  // namespace detail

// texture(): (r, 0, 0, 1) for a one-channel format.  u = s*size; i0 = floor(u - 0.5), i1 = i0 + 1, weight = frac(u - 0.5);
// indices clamped to [0, size-1]; blended along x, then y, then z as v0 + a*(v1 - v0).
inline vec4 texture(const sampler3D& s, const vec3& p)
{
    counters().texture_calls[s.slot]++;
    const float u = p.x * static_cast<float>(s.nx) - 0.5f;
    const float v = p.y * static_cast<float>(s.ny) - 0.5f;
    const float w = p.z * static_cast<float>(s.nz) - 0.5f;
    const float fu = std::floor(u), fv = std::floor(v), fw = std::floor(w);
    const float a = u - fu, b = v - fv, c = w - fw;
    const int i0 = detail::to_int_sat(fu), j0 = detail::to_int_sat(fv), k0 = detail::to_int_sat(fw);
    const int i0c = detail::clampi(i0, s.nx - 1), i1c = detail::clampi(i0 == 2147483647 ? i0 : i0 + 1, s.nx - 1);
    const int j0c = detail::clampi(j0, s.ny - 1), j1c = detail::clampi(j0 == 2147483647 ? j0 : j0 + 1, s.ny - 1);
    const int k0c = detail::clampi(k0, s.nz - 1), k1c = detail::clampi(k0 == 2147483647 ? k0 : k0 + 1, s.nz - 1);
    const size_t sy = (size_t)s.nx, sz = (size_t)s.nx * (size_t)s.ny;
    const float* t = s.texels;
    const float t000 = t[i0c + sy * j0c + sz * k0c], t100 = t[i1c + sy * j0c + sz * k0c];
    const float t010 = t[i0c + sy * j1c + sz * k0c], t110 = t[i1c + sy * j1c + sz * k0c];
    const float t001 = t[i0c + sy * j0c + sz * k1c], t101 = t[i1c + sy * j0c + sz * k1c];
    const float t011 = t[i0c + sy * j1c + sz * k1c], t111 = t[i1c + sy * j1c + sz * k1c];
    const float x00 = t000 + a * (t100 - t000);
    const float x10 = t010 + a * (t110 - t010);
    const float x01 = t001 + a * (t101 - t001);
    const float x11 = t011 + a * (t111 - t011);
    const float y0 = x00 + b * (x10 - x00);
    const float y1 = x01 + b * (x11 - x01);
    return vec4(y0 + c * (y1 - y0), 0.0f, 0.0f, 1.0f);
}

// out-of-bounds image loads return zero and stores are discarded (robustBufferAccess-style behaviour the reference
// relies on for the invocations beyond the image edge, RenderTechniquePT.cpp:141 dispatches W/32+1 groups)
inline vec4 imageLoad(const image2D& im, const ivec2& p)
{
    if (p.x < 0 || p.y < 0 || p.x >= im.width || p.y >= im.height) return vec4(0.0f);
    const float* t = im.texels + ((size_t)p.y * (size_t)im.width + (size_t)p.x) * 4;
    return vec4(t[0], t[1], t[2], t[3]);
}
inline void imageStore(const image2D& im, const ivec2& p, const vec4& v)
{
    if (p.x < 0 || p.y < 0 || p.x >= im.width || p.y >= im.height) return;
    float* t = im.texels + ((size_t)p.y * (size_t)im.width + (size_t)p.x) * 4;
    t[0] = v.x; t[1] = v.y; t[2] = v.z; t[3] = v.w;
}
inline void imageStore(const image3D& im, const ivec3& p, const vec4& v)
{
    if (p.x < 0 || p.y < 0 || p.z < 0 || p.x >= im.nx || p.y >= im.ny || p.z >= im.nz) return;
    im.texels[(size_t)p.x + (size_t)im.nx * ((size_t)p.y + (size_t)im.ny * (size_t)p.z)] = v.x;
}

// compute-shader built-in variables (per invocation)
struct Invocation {
    uvec3 global_id;
    uvec3 num_groups;
};
inline Invocation& invocation()
{
    static thread_local Invocation inv;
    return inv;
}
#define gl_GlobalInvocationID (::glsl::invocation().global_id)
#define gl_NumWorkGroups (::glsl::invocation().num_groups)

}  // namespace glsl
