// host_helpers.cpp -- C entry points for the host-side data model of the hot path: the parameter-struct
// setters of UniformBuffers.h (through include/crv/uniform_buffers.h), SetCloudProperties (main.cpp:144-163),
// the MSVC rand() frame-seed sequence (main.cpp:722,949) and the Grid3D ".xyz" file format (Grid3D.h,
// GridData.h).  Pure host code: compiled by g++ with -ffp-contract=off, no CUDA in this file.
#include <cmath>
#include <cstdio>
#include <vector>
#include <cstring>
#include <memory>
#include <new>

#include "../../include/crv/grid3d.h"
#include "../../include/crv/uniform_buffers.h"

void crv_set_error(const char* fmt, ...);

extern "C" {

void crv_push_constants_init(crv_push_constants* p) { *p = crv::PushConstants(); }
void crv_camera_init(crv_camera_properties* c) { *c = crv::CameraProperties(); }
void crv_camera_set_resolution(crv_camera_properties* c, int width, int height)
{
    static_cast<crv::CameraProperties*>(c)->SetResolution(width, height);
}
void crv_camera_set_rotation(crv_camera_properties* c, float* rotation_xy)
{
    static_cast<crv::CameraProperties*>(c)->SetRotation(rotation_xy);
}
void crv_camera_set_fov(crv_camera_properties* c, float* fov) { static_cast<crv::CameraProperties*>(c)->SetFOV(*fov); }
void crv_cloud_init(crv_cloud_properties* c) { *c = crv::CloudProperties(); }
void crv_cloud_set_from_grid(crv_cloud_properties* c, const uint32_t counts[3], const double voxel_size[3], float majorant)
{
    static_cast<crv::CloudProperties*>(c)->SetFromGrid(counts, voxel_size, majorant);
}
void crv_parameters_init(crv_parameters* p) { *p = crv::Parameters(); }
void crv_parameters_set_phase_g(crv_parameters* p, float g) { static_cast<crv::Parameters*>(p)->SetPhaseG(g); }
void crv_shadow_init(crv_shadow_volume_properties* s) { *s = crv::ShadowVolumeProperties(); }
void crv_shadow_set_light_direction(crv_shadow_volume_properties* s, const float dir[3])
{
    static_cast<crv::ShadowVolumeProperties*>(s)->SetLightDirection(dir);
}
void crv_shadow_set_origin(crv_shadow_volume_properties* s, const float lower[3], const float upper[3])
{
    static_cast<crv::ShadowVolumeProperties*>(s)->SetOrigin(lower, upper);
}
void crv_orthonormal_basis(const float n[3], float right[3], float up[3])
{
    crv::detail::V3 r, u;
    crv::detail::orthonormal_basis(crv::detail::V3{n[0], n[1], n[2]}, r, u);
    right[0] = r.x; right[1] = r.y; right[2] = r.z;
    up[0] = u.x; up[1] = u.y; up[2] = u.z;
}
void crv_msvc_rand_sequence(uint32_t seed0, unsigned n, int32_t* out) { crv::MsvcRandSequence(seed0, n, out); }

// ---- PhotonMapProperties (UniformBuffers.h:212-249) ----------------------------------------------------
void crv_photon_map_init(crv_photon_map_properties* p)
{
    std::memset(p, 0, sizeof(*p));
    const float inv = 1.0f / std::sqrt(1.0f * 1.0f + 1.0f * 1.0f);        // glm::normalize(vec4(1,-1,0,0)): v * inversesqrt(dot(v,v))
    p->lightDirection[0] = 1.0f * inv; p->lightDirection[1] = -1.0f * inv;
    for (int i = 0; i < 4; ++i) p->bounds[1][i] = 100.0f;
    p->voxelCount[0] = p->voxelCount[1] = p->voxelCount[2] = 100;
    p->voxelSize = (p->bounds[1][0] - p->bounds[0][0]) / (float)p->voxelCount[0];
    p->photonSize = (uint32_t)sizeof(crv_photon);
    p->stepSize = 10.0f;
    p->absorption = 0.0f;
}
void crv_photon_map_set_bounds(crv_photon_map_properties* p, const float bounds[2][4])
{
    float size[4];
    for (int i = 0; i < 4; ++i) { p->bounds[0][i] = bounds[0][i]; size[i] = bounds[1][i] - bounds[0][i]; }
    int hi = 0;
    float best = 1.175494351e-38f;                                         // FLT_MIN, like the reference's scan
    for (int i = 0; i < 3; ++i)
        if (best < std::fabs(size[i])) { hi = i; best = std::fabs(size[i]); }
    for (int i = 0; i < 4; ++i) p->bounds[1][i] = p->bounds[0][i] + size[hi];
    p->voxelSize = (p->bounds[1][0] - p->bounds[0][0]) / (float)p->voxelCount[0];
}
uint32_t crv_photon_map_total_size(const crv_photon_map_properties* p)
{
    return (uint32_t)(p->voxelCount[0] * p->voxelCount[1] * p->voxelCount[2]);
}

// ---- Grid3D .xyz -----------------------------------------------------------------------------------
int crv_grid_load(const char* path, uint32_t counts[3], double voxel_size[3], float* data, size_t capacity)
{
    if (!path || !counts || !voxel_size) { crv_set_error("crv_grid_load: null argument"); return CRV_ERR_INVALID; }
    if (!data) {
        // header-only query
        std::ifstream in(path, std::ifstream::in | std::ifstream::binary);
        if (!in.good()) { crv_set_error("crv_grid_load: cannot open '%s'", path); return CRV_ERR_IO; }
        in.read(reinterpret_cast<char*>(counts), sizeof(uint32_t) * 3);
        in.read(reinterpret_cast<char*>(voxel_size), sizeof(double) * 3);
        if (!in.good()) { crv_set_error("crv_grid_load: '%s' is shorter than the 36-byte header", path); return CRV_ERR_IO; }
        return CRV_OK;
    }
    try {
        std::unique_ptr<crv::Grid3D<float>> g(crv::Grid3D<float>::Load(path));
        if (!g) { crv_set_error("crv_grid_load: cannot open '%s'", path); return CRV_ERR_IO; }   // Load returns nullptr (Grid3D.h:54-58)
        const auto c = g->GetVoxelCount();
        const auto v = g->GetVoxelSize();
        for (int i = 0; i < 3; ++i) { counts[i] = c[i]; voxel_size[i] = v[i]; }
        if (g->GetSize() > capacity) { crv_set_error("crv_grid_load: buffer holds %zu floats, grid has %zu", capacity, g->GetSize()); return CRV_ERR_INVALID; }
        std::memcpy(data, g->data(), g->GetByteSize());
    } catch (const std::bad_alloc&) {
        crv_set_error("crv_grid_load: out of host memory");
        return CRV_ERR_OOM;
    }
    return CRV_OK;
}

static int grid_save(const char* path, const uint32_t counts[3], const double vs[3], const float* data, bool zfast)
{
    if (!path || !counts || !vs || !data) { crv_set_error("crv_grid_save: null argument"); return CRV_ERR_INVALID; }
    try {
        crv::Grid3D<float> g(counts[0], counts[1], counts[2], vs[0], vs[1], vs[2]);
        std::memcpy(g.GetData(), data, g.GetByteSize());
        if (zfast) g.SaveZFastest(path); else g.Save(path);
    } catch (const std::bad_alloc&) {
        crv_set_error("crv_grid_save: out of host memory");
        return CRV_ERR_OOM;
    }
    std::ifstream chk(path, std::ifstream::in | std::ifstream::binary);
    if (!chk.good()) { crv_set_error("crv_grid_save: cannot write '%s'", path); return CRV_ERR_IO; }
    return CRV_OK;
}

int crv_grid_save(const char* path, const uint32_t counts[3], const double voxel_size[3], const float* data)
{
    return grid_save(path, counts, voxel_size, data, false);
}
int crv_grid_save_zfastest(const char* path, const uint32_t counts[3], const double voxel_size[3], const float* data_xfastest)
{
    return grid_save(path, counts, voxel_size, data_xfastest, true);
}


// ---- PNG: 8-bit RGBA, filter 0 on every row, zlib stream of stored deflate blocks (no compression library needed) ----
struct CrcTable {
    uint32_t v[256];
    constexpr CrcTable() : v{}
    {
        for (uint32_t i = 0; i < 256; ++i) {
            uint32_t c = i;
            for (int k = 0; k < 8; ++k) c = (c & 1u) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
            v[i] = c;
        }
    }
};
static constexpr CrcTable kCrcTable{};          // built by the compiler: nothing to initialise (or race on) at run time

static uint32_t png_crc(uint32_t crc, const unsigned char* p, size_t n)
{
    crc = ~crc;
    for (size_t i = 0; i < n; ++i) crc = kCrcTable.v[(crc ^ p[i]) & 0xffu] ^ (crc >> 8);
    return ~crc;
}

static bool png_chunk(FILE* f, const char type[4], const unsigned char* data, uint32_t n)
{
    const unsigned char len[4] = {(unsigned char)(n >> 24), (unsigned char)(n >> 16), (unsigned char)(n >> 8), (unsigned char)n};
    uint32_t crc = png_crc(0, reinterpret_cast<const unsigned char*>(type), 4);
    if (n) crc = png_crc(crc, data, n);
    const unsigned char c4[4] = {(unsigned char)(crc >> 24), (unsigned char)(crc >> 16), (unsigned char)(crc >> 8), (unsigned char)crc};
    return std::fwrite(len, 1, 4, f) == 4 && std::fwrite(type, 1, 4, f) == 4 && (n == 0 || std::fwrite(data, 1, n, f) == n) &&
           std::fwrite(c4, 1, 4, f) == 4;
}

int crv_write_png_rgba8(const char* path, uint32_t width, uint32_t height, const unsigned char* rgba)
{
    if (!path || !rgba || width == 0 || height == 0) { crv_set_error("crv_write_png_rgba8: bad argument"); return CRV_ERR_INVALID; }
    // raw scanlines: filter byte 0 + width*4 bytes
    const size_t row = (size_t)width * 4 + 1, raw_n = row * height;
    std::vector<unsigned char> z;
    z.reserve(raw_n + raw_n / 65535 * 5 + 16);
    z.push_back(0x78); z.push_back(0x01);                         // zlib header, no preset dictionary
    uint32_t a = 1, b = 0;                                        // adler32 of the raw data
    std::vector<unsigned char> raw(raw_n);
    for (uint32_t y = 0; y < height; ++y) {
        raw[y * row] = 0;
        std::memcpy(&raw[y * row + 1], rgba + (size_t)y * width * 4, (size_t)width * 4);
    }
    for (size_t i = 0; i < raw_n; ++i) { a = (a + raw[i]) % 65521u; b = (b + a) % 65521u; }
    for (size_t off = 0; off < raw_n; off += 65535) {
        const size_t n = raw_n - off < 65535 ? raw_n - off : 65535;
        z.push_back(off + n == raw_n ? 1 : 0);                    // BFINAL, BTYPE = 00 (stored)
        z.push_back((unsigned char)(n & 0xff)); z.push_back((unsigned char)(n >> 8));
        z.push_back((unsigned char)(~n & 0xff)); z.push_back((unsigned char)((~n >> 8) & 0xff));
        z.insert(z.end(), raw.begin() + off, raw.begin() + off + n);
    }
    const uint32_t adler = (b << 16) | a;
    z.push_back((unsigned char)(adler >> 24)); z.push_back((unsigned char)(adler >> 16));
    z.push_back((unsigned char)(adler >> 8)); z.push_back((unsigned char)adler);

    FILE* f = std::fopen(path, "wb");
    if (!f) { crv_set_error("crv_write_png_rgba8: cannot write '%s'", path); return CRV_ERR_IO; }
    static const unsigned char sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    const unsigned char ihdr[13] = {(unsigned char)(width >> 24), (unsigned char)(width >> 16), (unsigned char)(width >> 8), (unsigned char)width,
                                    (unsigned char)(height >> 24), (unsigned char)(height >> 16), (unsigned char)(height >> 8), (unsigned char)height,
                                    8, 6, 0, 0, 0};               // 8 bit, colour type 6 (RGBA), deflate, adaptive filtering, no interlace
    bool ok = std::fwrite(sig, 1, 8, f) == 8 && png_chunk(f, "IHDR", ihdr, 13) && png_chunk(f, "IDAT", z.data(), (uint32_t)z.size()) &&
              png_chunk(f, "IEND", nullptr, 0);
    ok = (std::fclose(f) == 0) && ok;
    if (!ok) { crv_set_error("crv_write_png_rgba8: short write to '%s'", path); return CRV_ERR_IO; }
    return CRV_OK;
}

}  // extern "C"
