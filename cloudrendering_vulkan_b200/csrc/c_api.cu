// c_api.cu -- the C ABI of include/crv_c_api.h: contexts, volumes, images and the two render techniques.
//
// Host-side mirror of RenderTechniqueSV / RenderTechniquePT (RenderTechniqueSV.cpp:4-113,
// RenderTechniquePT.cpp:12-168): the Vulkan pipeline / descriptor-set plumbing becomes a table of bound
// resources per image index, vkCmdDispatch becomes a kernel launch on the caller's stream.
// All host arithmetic that feeds the kernels is fp32 in the shader's evaluation order (compiled with
// -ffp-contract=off), because the kernels take these values as if the shader had computed them.
#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cmath>
#include <cstdarg>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "crv_internal.h"
#include "pt_wavefront.h"
#include "sv_sweep.h"

// ---------------------------------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------------------------------
static thread_local char g_err[1024] = "";

// CUDA loads a kernel at its first launch by default, and that load may synchronise the device.  A group member whose stream
// waits for a peer (crv_group_*) must never be synchronised by the host thread that still has to enqueue that peer's work
// (crv_headless --gpus N, the tests: several members in one process), so the library asks for eager loading -- all of its
// kernels are loaded when the context first touches the module -- unless the environment already says otherwise.  The
// variable is read when CUDA initialises, which is after this library has been loaded in every host of this repo.
namespace {
struct EagerModuleLoading {
    EagerModuleLoading() { setenv("CUDA_MODULE_LOADING", "EAGER", 0); }
} g_eager_module_loading;
}  // namespace

void crv_set_error(const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

#define CRV_ENTER(ctxp)                                                     \
    do {                                                                    \
        if (!(ctxp)) { crv_set_error("null handle"); return CRV_ERR_INVALID; } \
        CRV_CUDA(cudaSetDevice((ctxp)->device));                            \
    } while (0)

#define CRV_TRY(expr)                  \
    do {                               \
        const int _rc = (expr);        \
        if (_rc != CRV_OK) return _rc; \
    } while (0)

// ---------------------------------------------------------------------------------------------------
// technique objects
// ---------------------------------------------------------------------------------------------------
struct crv_sv {
    crv_ctx* ctx = nullptr;
    const crv_shadow_volume_properties* ctor_props = nullptr;   // borrowed, read at record time (.cpp:110)
    const crv_push_constants* pc = nullptr;
    // pending writes (m_writeQueue) and committed bindings (the descriptor set)
    const crv_cloud_properties* q_cloud = nullptr;
    crv_volume* q_cloud_vol = nullptr;
    const crv_shadow_volume_properties* q_sv = nullptr;
    crv_volume* q_out = nullptr;
    bool has_cloud = false, has_sv = false;
    crv_cloud_properties cloud{};
    crv_shadow_volume_properties sv{};
    crv_volume* cloud_vol = nullptr;
    crv_volume* out = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    bool timed = false;
    float* dist_z = nullptr;         // accumulatedDistance after every slice (two-pass build), rcp_n entries
    uint32_t rcp_n = 0;              // N the per-N buffers (dist_z, zrange) were sized for
    float dist_for_vs = -1.0f;       // voxelSize dist_z currently holds the running distances of (fused build)
    ushort2* zrange = nullptr;       // per-column [z_in, z_end) of the two-pass build (rcp_n^2 entries)
    int mode = CRV_SV_MODE_AUTO;     // crv_sv_mode
    int last_mode = 0;
};

struct PTBindings {
    const crv_cloud_properties* q_cloud = nullptr;
    crv_volume* q_cloud_vol = nullptr;
    const crv_camera_properties* q_cam = nullptr;
    const crv_parameters* q_params = nullptr;
    const crv_shadow_volume_properties* q_sv = nullptr;
    crv_volume* q_sv_vol = nullptr;
    bool has_cloud = false, has_cam = false, has_params = false, has_sv = false;
    crv_cloud_properties cloud{};
    crv_camera_properties cam{};
    crv_parameters params{};
    crv_shadow_volume_properties sv{};
    crv_volume* cloud_vol = nullptr;
    crv_volume* sv_vol = nullptr;
};

struct crv_pt {
    crv_ctx* ctx = nullptr;
    const crv_camera_properties* ctor_cam = nullptr;            // borrowed, dispatch size (.cpp:141)
    const crv_push_constants* pc = nullptr;                     // borrowed, read at record time (.cpp:132)
    std::vector<crv_image*> images;
    std::vector<PTBindings> sets;
    crv_pt_extensions ext{};
    // hit list of the last classification (pixels whose camera ray enters the box)
    uint32_t* hitlist = nullptr;
    float* start = nullptr;             // per hit pixel start state (7 floats, SoA)
    uint32_t* nhit = nullptr;
    size_t hitlist_elems = 0;
    uint32_t* misslist = nullptr;      // pixels whose camera ray misses the box
    float* missdy = nullptr;           // dir.y of those rays
    // the same list sorted by descending expected work (crv_pt_set_job_order): built on demand, once per classification
    float* hit_keys = nullptr;
    uint32_t* hitlist_sorted = nullptr;
    float* start_sorted = nullptr;
    uint32_t* sorted_from = nullptr;
    uint32_t* sort_scratch = nullptr;
    bool sorted_valid = false;
    int job_order = -1;                // -1 automatic, 0 frame-major, 1 pixel-major over the sorted list
    // the classification depends on the camera and the cloud box only: kept across render calls while those bytes stand
    // ... and, through the occupied spans, on the cloud image's contents and the skip level
    bool geom_valid = false;
    unsigned char geom_key[68 + 32 + 4 + 8 + 8 + 4 + 4] = {};
    float* occ_aabb = nullptr;         // world-space box around them (6 floats), PTConst::occ_aabb
    int skip_level = 1;                // crv_pt_set_skip_level
    // q-table cache
    float* qtable_wide = nullptr;
    float* qtable = nullptr;
    size_t qtable_elems = 0;
    uint32_t q_dims[3] = {0, 0, 0};   // brick counts per axis the tables were sized for
    uint8_t* qcoarse = nullptr;       // pool scheduler: byte thresholds per coarse cell (shared-memory resident)
    uint32_t qcoarse_bytes = 0;       // padded to 16
    int qcoarse_shift = 0;            // coarse cell = (8 << shift)^3 voxels
    int forced_shift = -1;            // crv_pt_set_coarse_shift: >= 0 asks for at least this shift (tests: the 512^3 / 1024^3 table shapes on small grids)
    int q_forced_for = -1;            // forced_shift the tables were built for
    const crv_volume* q_for_vol = nullptr;
    float q_for_max_ext = -1.0f;
    uint64_t q_for_generation = 0;
    // per-launch buffers
    float2* samples = nullptr;
    size_t samples_bytes = 0;
    unsigned int* job_counter = nullptr;
    float* rcp_fc1 = nullptr;          // reciprocals of the running mean's divisors for the frames of the current render call
    size_t rcp_elems = 0;
    unsigned long long* stats = nullptr;
    uint32_t* dbg_steps = nullptr;
    uint32_t* dbg_coll = nullptr;
    size_t dbg_elems = 0;
    bool dbg_enabled = false;
    uint64_t launches = 0;
    // device timing of the path-tracing kernel launches of the last render call (CUDA events on its stream)
    std::vector<cudaEvent_t> ev;
    size_t ev_used = 0;
    // tuning
    int ctas_per_sm = 0;
    int track_threshold = 20;
    int steps_per_decision = 0;        // 0 = the scheduler's own default (pool: its unrolled round; lanes: 6)
    int scheduler = CRV_PT_SCHED_POOL;
    // Sample records are sized by the hit count once it is known on the host (copied back asynchronously after a
    // classification, never waited for), by the pixel count until then: c2's 256 frames need 290 MB, not 4.2 GB.
    uint64_t sample_budget = 2ull << 30;
    uint32_t* counts_host = nullptr;       // pinned: [hits, misses] of the last classification
    cudaEvent_t counts_ev = nullptr;
    bool counts_pending = false, nhit_valid = false;
    uint32_t nhit_known = 0;
};

// volumes carry a contents generation so cached q-tables / classifications notice re-uploads.  The numbers come from one
// process-wide atomic counter (a technique may be re-bound to another volume object that happens to live at the same address),
// the generation itself lives in the volume: no shared registry, contexts on different host threads do not meet here.
static std::atomic<uint64_t> g_generation{1};
static uint64_t vol_generation(const crv_volume* v) { return v->generation; }
static void vol_touch(crv_volume* v) { v->generation = g_generation.fetch_add(1, std::memory_order_relaxed) + 1; }

// stream-ordered scratch that is released on every path out of the scope, error returns included
struct AsyncScratch {
    void* p = nullptr;
    cudaStream_t s = nullptr;
    explicit AsyncScratch(cudaStream_t stream) : s(stream) {}
    cudaError_t alloc(size_t bytes) { return cudaMallocAsync(&p, bytes, s); }
    template <typename T> T* as() const { return static_cast<T*>(p); }
    ~AsyncScratch() { if (p) cudaFreeAsync(p, s); }
    AsyncScratch(const AsyncScratch&) = delete;
    AsyncScratch& operator=(const AsyncScratch&) = delete;
};

extern "C" {

// ---------------------------------------------------------------------------------------------------
// library / context
// ---------------------------------------------------------------------------------------------------
const char* crv_version(void) { return "crv-b200 0.1 (sm_100a)"; }
const char* crv_last_error(void) { return g_err; }

int crv_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int crv_ctx_create(int device, crv_ctx** out)
{
    CRV_REQUIRE(out != nullptr, "crv_ctx_create: out is null");
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) {
        cudaGetLastError();
        crv_set_error("no CUDA device available: libcrv_b200 has no CPU fallback");
        return CRV_ERR_NO_DEVICE;
    }
    if (device < 0 || device >= n) { crv_set_error("device %d out of range (0..%d)", device, n - 1); return CRV_ERR_INVALID; }
    cudaDeviceProp prop;
    CRV_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        crv_set_error("device %d is sm_%d%d; this library carries sm_100a code only (no fallback path)", device, prop.major, prop.minor);
        return CRV_ERR_NO_DEVICE;
    }
    CRV_CUDA(cudaSetDevice(device));
    {   // staging buffers (cudaMallocAsync) are re-used across calls instead of going back to the driver at every sync:
        // a 537 MB upload otherwise pays a fresh device allocation per call
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
            unsigned long long keep = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        cudaGetLastError();
    }
    crv_ctx* c = new (std::nothrow) crv_ctx();
    if (!c) return CRV_ERR_OOM;
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    c->cc_major = prop.major;
    c->cc_minor = prop.minor;
    *out = c;
    return CRV_OK;
}

int crv_ctx_destroy(crv_ctx* ctx)
{
    delete ctx;
    return CRV_OK;
}

int crv_ctx_synchronize(crv_ctx* ctx, void* stream)
{
    CRV_ENTER(ctx);
    CRV_CUDA(cudaStreamSynchronize(as_stream(stream)));
    return CRV_OK;
}

int crv_stream_create(crv_ctx* ctx, void** out)
{
    CRV_ENTER(ctx);
    CRV_REQUIRE(out != nullptr, "crv_stream_create: out is null");
    cudaStream_t s = nullptr;
    CRV_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    *out = s;
    return CRV_OK;
}

int crv_stream_destroy(crv_ctx* ctx, void* stream)
{
    CRV_ENTER(ctx);
    if (stream) CRV_CUDA(cudaStreamDestroy(as_stream(stream)));
    return CRV_OK;
}

int crv_ctx_launch_count(crv_ctx* ctx, uint64_t* out)
{
    CRV_REQUIRE(ctx && out, "crv_ctx_launch_count: null argument");
    *out = ctx->launches;
    return CRV_OK;
}

int crv_ctx_sm_count(crv_ctx* ctx, int* out)
{
    CRV_REQUIRE(ctx && out, "crv_ctx_sm_count: null argument");
    *out = ctx->sm_count;
    return CRV_OK;
}

// ---------------------------------------------------------------------------------------------------
// volumes
// ---------------------------------------------------------------------------------------------------
int crv_volume_create(crv_ctx* ctx, uint32_t nx, uint32_t ny, uint32_t nz, int layout, crv_volume** out)
{
    CRV_ENTER(ctx);
    CRV_REQUIRE(out != nullptr, "crv_volume_create: out is null");
    *out = nullptr;
    CRV_REQUIRE(nx > 0 && ny > 0 && nz > 0, "crv_volume_create: empty volume");
    CRV_REQUIRE(layout >= CRV_LAYOUT_LINEAR && layout <= CRV_LAYOUT_TEX3D, "crv_volume_create: unknown layout");
    const uint64_t nbx = (nx + 7) / 8, nby = (ny + 7) / 8, nbz = (nz + 7) / 8;
    CRV_REQUIRE(nbx * nby * nbz * 512ull < (1ull << 32), "crv_volume_create: more than 2^32 (padded) voxels is not supported");
    crv_volume* v = new (std::nothrow) crv_volume();
    if (!v) return CRV_ERR_OOM;
    v->ctx = ctx;
    v->layout = layout;
    v->nx = nx; v->ny = ny; v->nz = nz;
    v->nbx = (uint32_t)nbx; v->nby = (uint32_t)nby; v->nbz = (uint32_t)nbz;
    cudaError_t e = cudaSuccess;
    if (layout == CRV_LAYOUT_LINEAR) {
        v->data_elems = (size_t)nx * ny * nz;
        e = cudaMalloc(&v->data, v->data_elems * sizeof(float));
    } else if (layout == CRV_LAYOUT_BRICK) {
        v->data_elems = (size_t)(nbx * nby * nbz) * 512;
        e = cudaMalloc(&v->data, v->data_elems * sizeof(float));
    } else {
        const cudaChannelFormatDesc desc = cudaCreateChannelDesc<float>();
        e = cudaMalloc3DArray(&v->array, &desc, make_cudaExtent(nx, ny, nz), cudaArraySurfaceLoadStore);
        if (e == cudaSuccess) {
            cudaResourceDesc rd{};
            rd.resType = cudaResourceTypeArray;
            rd.res.array.array = v->array;
            cudaTextureDesc td{};
            td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeClamp;
            td.filterMode = cudaFilterModePoint;          // never the 9-bit hardware lerp
            td.readMode = cudaReadModeElementType;
            td.normalizedCoords = 0;
            e = cudaCreateTextureObject(&v->tex, &rd, &td, nullptr);
            if (e == cudaSuccess) e = cudaCreateSurfaceObject(&v->surf, &rd);
        }
    }
    if (e == cudaSuccess) e = cudaMalloc(&v->occ_ext, 6 * sizeof(int));
    if (e != cudaSuccess) {
        crv_set_error("crv_volume_create(%u,%u,%u,layout %d): %s", nx, ny, nz, layout, cudaGetErrorString(e));
        crv_volume_destroy(v);
        return e == cudaErrorMemoryAllocation ? CRV_ERR_OOM : CRV_ERR_CUDA;
    }
    vol_touch(v);
    *out = v;
    return CRV_OK;
}

int crv_volume_destroy(crv_volume* v)
{
    if (!v) return CRV_OK;
    cudaSetDevice(v->ctx->device);
    if (v->tex) cudaDestroyTextureObject(v->tex);
    if (v->surf) cudaDestroySurfaceObject(v->surf);
    if (v->array) cudaFreeArray(v->array);
    if (v->data) cudaFree(v->data);
    if (v->brick_max) cudaFree(v->brick_max);
    if (v->brick_absmax) cudaFree(v->brick_absmax);
    if (v->brick_absmax_wide) cudaFree(v->brick_absmax_wide);
    if (v->occ_ext) cudaFree(v->occ_ext);
    delete v;
    return CRV_OK;
}

int crv_volume_upload(crv_volume* v, const float* data, int is_device, void* stream)
{
    CRV_REQUIRE(v && data, "crv_volume_upload: null argument");
    CRV_ENTER(v->ctx);
    cudaStream_t s = as_stream(stream);
    const size_t n = (size_t)v->nx * v->ny * v->nz;
    const cudaMemcpyKind kind = is_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    // scratch of the fused table pass (volume.cu): 3 window-maxima arrays of nbx * ny * nz floats, stream-ordered
    const size_t scratch_bytes = 3 * (size_t)v->nbx * v->ny * v->nz * sizeof(float);
    bool tables_done = false;
    if (v->layout == CRV_LAYOUT_LINEAR) {
        CRV_CUDA(cudaMemcpyAsync(v->data, data, n * sizeof(float), kind, s));
        AsyncScratch scratch(s);
        CRV_CUDA(scratch.alloc(scratch_bytes));
        CRV_TRY(crv_launch_retile_tables(v->ctx, v->data, v, false, scratch.as<float>(), s));
        tables_done = true;
    } else if (v->layout == CRV_LAYOUT_TEX3D) {
        cudaMemcpy3DParms p{};
        p.srcPtr = make_cudaPitchedPtr(const_cast<float*>(data), v->nx * sizeof(float), v->nx, v->ny);
        p.dstArray = v->array;
        p.extent = make_cudaExtent(v->nx, v->ny, v->nz);
        p.kind = kind;
        CRV_CUDA(cudaMemcpy3DAsync(&p, s));
    } else {
        AsyncScratch scratch(s), tmp(s);
        CRV_CUDA(scratch.alloc(scratch_bytes));
        if (is_device) {
            CRV_TRY(crv_launch_retile_tables(v->ctx, data, v, true, scratch.as<float>(), s));
        } else {
            CRV_CUDA(tmp.alloc(n * sizeof(float)));
            CRV_CUDA(cudaMemcpyAsync(tmp.p, data, n * sizeof(float), cudaMemcpyHostToDevice, s));
            CRV_TRY(crv_launch_retile_tables(v->ctx, tmp.as<float>(), v, true, scratch.as<float>(), s));
        }
        tables_done = true;
    }
    v->brick_max_valid = tables_done;
    v->majorant_valid = false;
    vol_touch(v);
    return CRV_OK;
}

int crv_volume_download(crv_volume* v, float* host_out, void* stream)
{
    CRV_REQUIRE(v && host_out, "crv_volume_download: null argument");
    CRV_ENTER(v->ctx);
    cudaStream_t s = as_stream(stream);
    const size_t n = (size_t)v->nx * v->ny * v->nz;
    if (v->layout == CRV_LAYOUT_LINEAR) {
        CRV_CUDA(cudaMemcpyAsync(host_out, v->data, n * sizeof(float), cudaMemcpyDeviceToHost, s));
    } else {
        AsyncScratch tmp(s);
        CRV_CUDA(tmp.alloc(n * sizeof(float)));
        CRV_TRY(crv_launch_untile(v->ctx, v, tmp.as<float>(), s));
        CRV_CUDA(cudaMemcpyAsync(host_out, tmp.p, n * sizeof(float), cudaMemcpyDeviceToHost, s));
    }
    CRV_CUDA(cudaStreamSynchronize(s));
    return CRV_OK;
}

int crv_volume_dims(const crv_volume* v, uint32_t* nx, uint32_t* ny, uint32_t* nz, int* layout)
{
    CRV_REQUIRE(v != nullptr, "crv_volume_dims: null volume");
    if (nx) *nx = v->nx;
    if (ny) *ny = v->ny;
    if (nz) *nz = v->nz;
    if (layout) *layout = v->layout;
    return CRV_OK;
}

int crv_volume_majorant(crv_volume* v, float* out, void* stream)
{
    CRV_REQUIRE(v && out, "crv_volume_majorant: null argument");
    CRV_ENTER(v->ctx);
    if (!v->majorant_valid) CRV_TRY(crv_launch_majorant(v->ctx, v, as_stream(stream)));
    *out = v->majorant;
    return CRV_OK;
}

int crv_volume_device_ptr(crv_volume* v, float** out)
{
    CRV_REQUIRE(v && out, "crv_volume_device_ptr: null argument");
    *out = (v->layout == CRV_LAYOUT_LINEAR) ? v->data : nullptr;
    return CRV_OK;
}

int crv_volume_mark_dirty(crv_volume* v)
{
    CRV_REQUIRE(v != nullptr, "crv_volume_mark_dirty: null volume");
    v->brick_max_valid = false;
    v->majorant_valid = false;
    vol_touch(v);
    return CRV_OK;
}

int crv_fbm_generate(crv_volume* v, uint32_t seed, void* stream)
{
    CRV_REQUIRE(v != nullptr, "crv_fbm_generate: null volume");
    CRV_ENTER(v->ctx);
    CRV_REQUIRE(v->nx == v->ny && v->ny == v->nz, "crv_fbm_generate: the synthetic scene is a cube");
    cudaStream_t s = as_stream(stream);
    if (v->layout == CRV_LAYOUT_LINEAR) {
        CRV_TRY(crv_launch_fbm(v->ctx, v->data, v->nx, seed, s));
        v->brick_max_valid = false;
        v->majorant_valid = false;
        vol_touch(v);
        return CRV_OK;
    }
    const size_t n = (size_t)v->nx * v->ny * v->nz;
    AsyncScratch tmp(s);
    CRV_CUDA(tmp.alloc(n * sizeof(float)));
    CRV_TRY(crv_launch_fbm(v->ctx, tmp.as<float>(), v->nx, seed, s));
    return crv_volume_upload(v, tmp.as<float>(), 1, stream);
}

// ---------------------------------------------------------------------------------------------------
// images
// ---------------------------------------------------------------------------------------------------
int crv_image_create(crv_ctx* ctx, uint32_t width, uint32_t height, crv_image** out)
{
    CRV_ENTER(ctx);
    CRV_REQUIRE(out != nullptr, "crv_image_create: out is null");
    *out = nullptr;
    CRV_REQUIRE(width > 0 && height > 0, "crv_image_create: empty image");
    crv_image* img = new (std::nothrow) crv_image();
    if (!img) return CRV_ERR_OOM;
    img->ctx = ctx; img->width = width; img->height = height; img->owned = true;
    const cudaError_t e = cudaMalloc(&img->rgba, (size_t)width * height * 4 * sizeof(float));
    if (e != cudaSuccess) {
        crv_set_error("crv_image_create(%u,%u): %s", width, height, cudaGetErrorString(e));
        delete img;
        return e == cudaErrorMemoryAllocation ? CRV_ERR_OOM : CRV_ERR_CUDA;
    }
    CRV_CUDA(cudaMemset(img->rgba, 0, (size_t)width * height * 4 * sizeof(float)));
    *out = img;
    return CRV_OK;
}

int crv_image_wrap(crv_ctx* ctx, float* device_rgba, uint32_t width, uint32_t height, crv_image** out)
{
    CRV_ENTER(ctx);
    CRV_REQUIRE(out && device_rgba && width > 0 && height > 0, "crv_image_wrap: bad argument");
    CRV_REQUIRE((reinterpret_cast<uintptr_t>(device_rgba) & 15u) == 0, "crv_image_wrap: pointer must be 16-byte aligned");
    crv_image* img = new (std::nothrow) crv_image();
    if (!img) return CRV_ERR_OOM;
    img->ctx = ctx; img->width = width; img->height = height; img->rgba = device_rgba; img->owned = false;
    *out = img;
    return CRV_OK;
}

int crv_image_destroy(crv_image* img)
{
    if (!img) return CRV_OK;
    cudaSetDevice(img->ctx->device);
    if (img->owned && img->rgba) cudaFree(img->rgba);
    if (img->bgra8) cudaFree(img->bgra8);
    delete img;
    return CRV_OK;
}

int crv_image_clear(crv_image* img, void* stream)
{
    CRV_REQUIRE(img != nullptr, "crv_image_clear: null image");
    CRV_ENTER(img->ctx);
    CRV_CUDA(cudaMemsetAsync(img->rgba, 0, (size_t)img->width * img->height * 16, as_stream(stream)));
    return CRV_OK;
}

int crv_image_download(crv_image* img, float* host_rgba, void* stream)
{
    CRV_REQUIRE(img && host_rgba, "crv_image_download: null argument");
    CRV_ENTER(img->ctx);
    CRV_CUDA(cudaMemcpyAsync(host_rgba, img->rgba, (size_t)img->width * img->height * 16, cudaMemcpyDeviceToHost, as_stream(stream)));
    CRV_CUDA(cudaStreamSynchronize(as_stream(stream)));
    return CRV_OK;
}

int crv_image_upload(crv_image* img, const float* host_rgba, void* stream)
{
    CRV_REQUIRE(img && host_rgba, "crv_image_upload: null argument");
    CRV_ENTER(img->ctx);
    CRV_CUDA(cudaMemcpyAsync(img->rgba, host_rgba, (size_t)img->width * img->height * 16, cudaMemcpyHostToDevice, as_stream(stream)));
    return CRV_OK;
}

// ---- present step and image files ------------------------------------------------------------
int crv_image_blit_srgb8(crv_image* img, uint8_t* host_bgra8, void* stream)
{
    CRV_REQUIRE(img && host_bgra8, "crv_image_blit_srgb8: null argument");
    CRV_ENTER(img->ctx);
    const size_t npix = (size_t)img->width * img->height;
    if (!img->bgra8) CRV_CUDA(cudaMalloc(&img->bgra8, npix * 4));
    cudaStream_t s = as_stream(stream);
    CRV_TRY(crv_launch_blit_srgb8(img->ctx, img->rgba, img->bgra8, npix, s));
    CRV_CUDA(cudaMemcpyAsync(host_bgra8, img->bgra8, npix * 4, cudaMemcpyDeviceToHost, s));
    CRV_CUDA(cudaStreamSynchronize(s));
    return CRV_OK;
}

int crv_image_save_png(crv_image* img, const char* path, void* stream)
{
    CRV_REQUIRE(img && path, "crv_image_save_png: null argument");
    std::vector<uint8_t> px((size_t)img->width * img->height * 4);
    CRV_TRY(crv_image_blit_srgb8(img, px.data(), stream));
    for (size_t i = 0; i < px.size(); i += 4) std::swap(px[i], px[i + 2]);        // B8G8R8A8 -> PNG's R, G, B, A
    return crv_write_png_rgba8(path, img->width, img->height, px.data());
}

int crv_image_save_pfm(crv_image* img, const char* path, void* stream)
{
    CRV_REQUIRE(img && path, "crv_image_save_pfm: null argument");
    std::vector<float> px((size_t)img->width * img->height * 4);
    CRV_TRY(crv_image_download(img, px.data(), stream));
    FILE* f = std::fopen(path, "wb");
    if (!f) { crv_set_error("crv_image_save_pfm: cannot write '%s'", path); return CRV_ERR_IO; }
    std::fprintf(f, "PF\n%u %u\n-1.0\n", img->width, img->height);              // RGB float, little endian, bottom-up rows
    bool ok = true;
    for (uint32_t y = img->height; y-- > 0 && ok;)
        for (uint32_t x = 0; x < img->width && ok; ++x) ok = std::fwrite(&px[((size_t)y * img->width + x) * 4], sizeof(float), 3, f) == 3;
    ok = (std::fclose(f) == 0) && ok;
    if (!ok) { crv_set_error("crv_image_save_pfm: short write to '%s'", path); return CRV_ERR_IO; }
    return CRV_OK;
}

// Accumulation checkpoint: the progressive state of the path is the result image plus the frame counter the next
// dispatch will push (PathTracer.comp:420-424); the seed of the last frame is kept so a resumed run can continue the
// reference's rand() sequence.  Layout: magic "CRVACC1\0", u32 width, u32 height, u32 frameCount, i32 seed, then
// width*height*4 little-endian floats.
struct CheckpointHeader {
    char magic[8];
    uint32_t width, height, frame_count;
    int32_t seed;
};

int crv_image_save_checkpoint(crv_image* img, uint32_t frame_count, int32_t seed, const char* path, void* stream)
{
    CRV_REQUIRE(img && path, "crv_image_save_checkpoint: null argument");
    std::vector<float> px((size_t)img->width * img->height * 4);
    CRV_TRY(crv_image_download(img, px.data(), stream));
    CheckpointHeader h{};
    std::memcpy(h.magic, "CRVACC1", 8);
    h.width = img->width; h.height = img->height; h.frame_count = frame_count; h.seed = seed;
    FILE* f = std::fopen(path, "wb");
    if (!f) { crv_set_error("crv_image_save_checkpoint: cannot write '%s'", path); return CRV_ERR_IO; }
    bool ok = std::fwrite(&h, sizeof(h), 1, f) == 1 && std::fwrite(px.data(), sizeof(float), px.size(), f) == px.size();
    ok = (std::fclose(f) == 0) && ok;
    if (!ok) { crv_set_error("crv_image_save_checkpoint: short write to '%s'", path); return CRV_ERR_IO; }
    return CRV_OK;
}

int crv_image_load_checkpoint(crv_image* img, uint32_t* frame_count, int32_t* seed, const char* path, void* stream)
{
    CRV_REQUIRE(img && path, "crv_image_load_checkpoint: null argument");
    FILE* f = std::fopen(path, "rb");
    if (!f) { crv_set_error("crv_image_load_checkpoint: cannot open '%s'", path); return CRV_ERR_IO; }
    CheckpointHeader h{};
    std::vector<float> px((size_t)img->width * img->height * 4);
    bool ok = std::fread(&h, sizeof(h), 1, f) == 1;
    if (ok && (std::memcmp(h.magic, "CRVACC1", 8) != 0 || h.width != img->width || h.height != img->height)) {
        std::fclose(f);
        crv_set_error("crv_image_load_checkpoint: '%s' is not a %ux%u accumulation checkpoint", path, img->width, img->height);
        return CRV_ERR_INVALID;
    }
    ok = ok && std::fread(px.data(), sizeof(float), px.size(), f) == px.size();
    std::fclose(f);
    if (!ok) { crv_set_error("crv_image_load_checkpoint: '%s' is truncated", path); return CRV_ERR_IO; }
    CRV_TRY(crv_image_upload(img, px.data(), stream));
    CRV_ENTER(img->ctx);
    CRV_CUDA(cudaStreamSynchronize(as_stream(stream)));          // px goes out of scope
    if (frame_count) *frame_count = h.frame_count;
    if (seed) *seed = h.seed;
    return CRV_OK;
}

int crv_image_device_ptr(crv_image* img, float** out)
{
    CRV_REQUIRE(img && out, "crv_image_device_ptr: null argument");
    *out = img->rgba;
    return CRV_OK;
}

int crv_image_scale(crv_image* img, float scale, void* stream)
{
    CRV_REQUIRE(img != nullptr, "crv_image_scale: null image");
    CRV_ENTER(img->ctx);
    return crv_launch_scale(img->ctx, img->rgba, (size_t)img->width * img->height * 4, scale, as_stream(stream));
}

// ---------------------------------------------------------------------------------------------------
// RenderTechniqueSV
// ---------------------------------------------------------------------------------------------------
int crv_sv_create(crv_ctx* ctx, const crv_shadow_volume_properties* props, const crv_push_constants* pc, crv_sv** out)
{
    CRV_ENTER(ctx);
    CRV_REQUIRE(out != nullptr, "crv_sv_create: out is null");
    *out = nullptr;
    CRV_REQUIRE(props != nullptr, "crv_sv_create: ShadowVolumeProperties pointer is null");
    crv_sv* sv = new (std::nothrow) crv_sv();
    if (!sv) return CRV_ERR_OOM;
    sv->ctx = ctx;
    sv->ctor_props = props;
    sv->pc = pc;
    CRV_CUDA(cudaEventCreate(&sv->ev0));
    CRV_CUDA(cudaEventCreate(&sv->ev1));
    *out = sv;
    return CRV_OK;
}

int crv_sv_destroy(crv_sv* sv)
{
    if (!sv) return CRV_OK;
    cudaSetDevice(sv->ctx->device);
    if (sv->ev0) cudaEventDestroy(sv->ev0);
    if (sv->ev1) cudaEventDestroy(sv->ev1);
    if (sv->dist_z) cudaFree(sv->dist_z);
    if (sv->zrange) cudaFree(sv->zrange);
    delete sv;
    return CRV_OK;
}

int crv_sv_set_frame_references(crv_sv* sv, crv_volume* shadow_image)
{
    CRV_REQUIRE(sv && shadow_image, "crv_sv_set_frame_references: null argument");
    sv->out = shadow_image;         // SetFrameReferences writes binding 0 immediately (.cpp:49-65)
    return CRV_OK;
}
int crv_sv_clear_frame_references(crv_sv* sv)
{
    CRV_REQUIRE(sv != nullptr, "crv_sv_clear_frame_references: null technique");
    return CRV_OK;                  // the reference's ClearFrameReferences is empty (.cpp:67-69)
}
int crv_sv_queue_update_cloud_data(crv_sv* sv, const crv_cloud_properties* cloud)
{
    CRV_REQUIRE(sv && cloud, "crv_sv_queue_update_cloud_data: null argument");
    sv->q_cloud = cloud;
    return CRV_OK;
}
int crv_sv_queue_update_cloud_data_sampler(crv_sv* sv, crv_volume* cloud_image)
{
    CRV_REQUIRE(sv && cloud_image, "crv_sv_queue_update_cloud_data_sampler: null argument");
    sv->q_cloud_vol = cloud_image;
    return CRV_OK;
}
int crv_sv_queue_update_camera_properties(crv_sv* sv, const crv_camera_properties*) { CRV_REQUIRE(sv != nullptr, "null technique"); return CRV_OK; }
int crv_sv_queue_update_parameters(crv_sv* sv, const crv_parameters*) { CRV_REQUIRE(sv != nullptr, "null technique"); return CRV_OK; }
int crv_sv_queue_update_shadow_volume(crv_sv* sv, const crv_shadow_volume_properties* p)
{
    CRV_REQUIRE(sv && p, "crv_sv_queue_update_shadow_volume: null argument");
    sv->q_sv = p;
    return CRV_OK;
}
int crv_sv_queue_update_shadow_volume_sampler(crv_sv* sv, crv_volume* shadow_image)
{
    CRV_REQUIRE(sv && shadow_image, "crv_sv_queue_update_shadow_volume_sampler: null argument");
    sv->q_out = shadow_image;       // rebinds the STORAGE image, binding 0 (.cpp:92-95)
    return CRV_OK;
}
int crv_sv_update_descriptor_sets(crv_sv* sv)
{
    CRV_REQUIRE(sv != nullptr, "crv_sv_update_descriptor_sets: null technique");
    if (sv->q_cloud) { sv->cloud = *sv->q_cloud; sv->has_cloud = true; sv->q_cloud = nullptr; }
    if (sv->q_cloud_vol) { sv->cloud_vol = sv->q_cloud_vol; sv->q_cloud_vol = nullptr; }
    if (sv->q_sv) { sv->sv = *sv->q_sv; sv->has_sv = true; sv->q_sv = nullptr; }
    if (sv->q_out) { sv->out = sv->q_out; sv->q_out = nullptr; }
    return CRV_OK;
}

int crv_sv_record_draw_commands(crv_sv* sv, void* stream, unsigned /*image_index*/)
{
    CRV_REQUIRE(sv != nullptr, "crv_sv_record_draw_commands: null technique");
    CRV_ENTER(sv->ctx);
    CRV_REQUIRE(sv->has_cloud && sv->cloud_vol, "RenderTechniqueSV: cloud data / sampler not bound (bindings 2,3)");
    CRV_REQUIRE(sv->has_sv, "RenderTechniqueSV: ShadowVolumeProperties not bound (binding 1)");
    CRV_REQUIRE(sv->out != nullptr, "RenderTechniqueSV: no shadow-volume image (binding 0)");
    const uint32_t n = sv->sv.voxelAxisCount;
    CRV_REQUIRE(n > 0 && sv->ctor_props->voxelAxisCount == n, "RenderTechniqueSV: dispatch size and bound voxelAxisCount disagree");
    CRV_REQUIRE(sv->out->nx == n && sv->out->ny == n && sv->out->nz == n, "RenderTechniqueSV: shadow image is not voxelAxisCount^3");
    CRV_REQUIRE(sv->out != sv->cloud_vol, "RenderTechniqueSV: cloud and shadow images alias");

    SVConst c{};
    for (int i = 0; i < 3; ++i) {
        c.b0[i] = sv->sv.bounds[0][i];
        c.right[i] = sv->sv.right[i];
        c.up[i] = sv->sv.up[i];
        c.light[i] = sv->sv.lightDirection[i];
        c.cb0[i] = sv->cloud.bounds[0][i];
        c.cb1[i] = sv->cloud.bounds[1][i];
        c.cext[i] = sv->cloud.bounds[1][i] - sv->cloud.bounds[0][i];
    }
    c.voxel_size = sv->sv.voxelSize;
    c.n = (int)n;
    c.density_scaling = sv->cloud.densityScaling;
    c.base_scaling = sv->cloud.baseScaling;
    c.cloud = sv->cloud_vol->view();
    c.out = sv->out->view();
    cudaStream_t s = as_stream(stream);
    // fast-path extras: host-side IEEE reciprocals (this file is compiled without contraction / fast-math)
    CRV_TRY(crv_sv_init_tables(sv->ctx));
    if (sv->rcp_n != n) {
        if (sv->dist_z) CRV_CUDA(cudaFree(sv->dist_z));
        if (sv->zrange) CRV_CUDA(cudaFree(sv->zrange));
        sv->dist_z = nullptr;
        sv->zrange = nullptr;
        sv->rcp_n = 0;
        CRV_CUDA(cudaMalloc(&sv->zrange, sizeof(ushort2) * (size_t)n * n));
        CRV_CUDA(cudaMalloc(&sv->dist_z, sizeof(float) * n));
        sv->rcp_n = n;
        sv->dist_for_vs = -1.0f;
    }
    if (!sv->cloud_vol->brick_max_valid) CRV_TRY(crv_launch_brick_max(sv->ctx, sv->cloud_vol, s));
    for (int i = 0; i < 3; ++i) c.rcext[i] = 1.0f / c.cext[i];
    c.rbase = 1.0f / c.base_scaling;
    c.brick_absmax = sv->cloud_vol->brick_absmax;
    c.occ_ext = sv->cloud_vol->occ_ext;
    {   // fused build's approximate brick coordinate (doubles here, one rounding to fp32 each)
        bool ok = true;
        for (int i = 0; i < 3; ++i) {
            const double nvox = i == 0 ? c.cloud.nx : i == 1 ? c.cloud.ny : c.cloud.nz;
            const double k = nvox / (8.0 * (double)c.cext[i]);
            const double vs = (double)c.voxel_size;
            c.pre_c[i] = (float)(((double)c.b0[i] - (double)c.cb0[i]) * k - 1.0 / 16.0);
            c.pre_r[i] = (float)((double)c.right[i] * vs * k);
            c.pre_u[i] = (float)((double)c.up[i] * vs * k);
            c.pre_l[i] = (float)((double)c.light[i] * vs * k);
            // the exact chain rounds absolute coordinates: its distance from the real-number position, in texels, must stay
            // far below the one texel the wide table tolerates (same criterion as the path tracer's pre-test)
            const double mag = std::fabs((double)c.b0[i]) + std::fabs((double)c.cb0[i]) + std::fabs((double)c.cb1[i]) + 2.0 * vs * (double)n;
            if (!(mag * std::fabs(8.0 * k) * 8.0 * 1.1920929e-07 < 0.25) || !(nvox < 2097152.0)) ok = false;
        }
        c.brick_absmax_wide = ok ? sv->cloud_vol->brick_absmax_wide : nullptr;
    }
    c.dist_z = sv->dist_z;
    c.search_iters = 1;
    while ((1u << c.search_iters) < n + 1u) c.search_iters++;
    c.zrange = sv->zrange;
    // AUTO: the fused builds store whole 8^3 blocks, which is one contiguous brick of the BRICK layout -- warp-owned rows:
    // 0.124 ms at 512^3, 0.86 ms at 1024^3 (CTA-wide chunks 0.135 / 0.92, two-pass 0.26 / 1.81); in the other layouts the
    // two-pass build's row stores do as well or better
    int mode = sv->mode != CRV_SV_MODE_AUTO ? sv->mode : (sv->out->layout == CRV_LAYOUT_BRICK ? CRV_SV_MODE_FUSED_ROWS : CRV_SV_MODE_TWO_PASS);
    if (!crv_sv_fast_path_ok(c)) mode = CRV_SV_MODE_GENERIC;       // preconditions of the fast kernels not met
    sv->last_mode = mode;
    CRV_CUDA(cudaEventRecord(sv->ev0, s));
    if ((mode == CRV_SV_MODE_FUSED || mode == CRV_SV_MODE_FUSED_V1 || mode == CRV_SV_MODE_FUSED_ROWS) && sv->dist_for_vs != c.voxel_size) {       // inside the timed region when it is needed
        CRV_TRY(crv_launch_sv_dist(sv->ctx, sv->dist_z, c.n, c.voxel_size, s));
        sv->dist_for_vs = c.voxel_size;
    } else if (mode == CRV_SV_MODE_TWO_PASS) {
        sv->dist_for_vs = c.voxel_size;                                        // its first kernel rewrites the table
    }
    CRV_TRY(crv_launch_sv_sweep(sv->ctx, c, sv->cloud_vol->layout, sv->out->layout, mode, s));
    CRV_CUDA(cudaEventRecord(sv->ev1, s));
    sv->timed = true;
    sv->out->brick_max_valid = false;
    sv->out->majorant_valid = false;
    vol_touch(sv->out);
    return CRV_OK;
}

int crv_sv_set_tuning(crv_sv* sv, int mode)
{
    CRV_REQUIRE(sv != nullptr, "crv_sv_set_tuning: null technique");
    CRV_REQUIRE(mode >= CRV_SV_MODE_AUTO && mode <= CRV_SV_MODE_FUSED_ROWS, "crv_sv_set_tuning: unknown mode");
    sv->mode = mode;
    return CRV_OK;
}

int crv_sv_last_build_mode(crv_sv* sv, int* out)
{
    CRV_REQUIRE(sv && out, "crv_sv_last_build_mode: null argument");
    *out = sv->last_mode;
    return CRV_OK;
}

int crv_sv_last_build_ms(crv_sv* sv, float* ms)
{
    CRV_REQUIRE(sv && ms, "crv_sv_last_build_ms: null argument");
    CRV_ENTER(sv->ctx);
    CRV_REQUIRE(sv->timed, "crv_sv_last_build_ms: nothing recorded yet");
    CRV_CUDA(cudaEventSynchronize(sv->ev1));
    CRV_CUDA(cudaEventElapsedTime(ms, sv->ev0, sv->ev1));
    return CRV_OK;
}

// ---------------------------------------------------------------------------------------------------
// RenderTechniquePT
// ---------------------------------------------------------------------------------------------------
int crv_pt_create(crv_ctx* ctx, const crv_camera_properties* cam, const crv_push_constants* pc, crv_pt** out)
{
    CRV_ENTER(ctx);
    CRV_REQUIRE(out != nullptr, "crv_pt_create: out is null");
    *out = nullptr;
    CRV_REQUIRE(cam && pc, "crv_pt_create: CameraProperties / PushConstants pointer is null");
    crv_pt* pt = new (std::nothrow) crv_pt();
    if (!pt) return CRV_ERR_OOM;
    pt->ctx = ctx;
    pt->ctor_cam = cam;
    pt->pc = pc;
    pt->sets.resize(1);
    const int rc = [&]() -> int {
        CRV_CUDA(cudaMalloc(&pt->job_counter, sizeof(unsigned int)));
        CRV_CUDA(cudaMalloc(&pt->nhit, 2 * sizeof(uint32_t)));      // [0] hits, [1] misses
        CRV_CUDA(cudaMallocHost(&pt->counts_host, 2 * sizeof(uint32_t)));
        CRV_CUDA(cudaEventCreateWithFlags(&pt->counts_ev, cudaEventDisableTiming));
        CRV_CUDA(cudaMalloc(&pt->occ_aabb, 6 * sizeof(float)));
        CRV_CUDA(cudaMalloc(&pt->stats, sizeof(unsigned long long) * PT_STAT_COUNT));
        CRV_CUDA(cudaMemset(pt->stats, 0, sizeof(unsigned long long) * PT_STAT_COUNT));
        return CRV_OK;
    }();
    if (rc != CRV_OK) {
        crv_pt_destroy(pt);           // releases whatever was allocated before the failure
        return rc;
    }
    *out = pt;
    return CRV_OK;
}

// Coarse threshold table (pt_wavefront.h: PTConst::qcoarse): cell size and padded byte count for a cloud volume, and the
// approximate-cell-coordinate constants of the tracker's pre-test for the geometry in c (cb0 / cb1 / cext / cloud set).
static void coarse_table_shape(const crv_volume* cv, int min_shift, int& shift, uint32_t& bytes)
{
    size_t cells = 0;
    for (shift = min_shift > 0 ? min_shift : 0;; ++shift) {
        const size_t ncx = ((cv->nbx - 1) >> shift) + 1, ncy = ((cv->nby - 1) >> shift) + 1, ncz = ((cv->nbz - 1) >> shift) + 1;
        cells = (ncx + 2) * (ncy + 2) * (ncz + 2);
        if (cells + 1 <= (size_t)crv_pt_pool_smem_budget()) break;
    }
    bytes = (uint32_t)((cells + 1 + 15) / 16 * 16);
}

static void coarse_pretest_geometry(PTConst& c, int shift)
{
    c.use_wide = 1;
    for (int i = 0; i < 3; ++i) {
        const double n = i == 0 ? c.cloud.nx : i == 1 ? c.cloud.ny : c.cloud.nz;
        const double sc = n / (double)c.cext[i], of = -(double)c.cb0[i] * sc - 0.5;
        const double cell = 8.0 * (double)(1 << shift);
        c.ascc[i] = (float)(sc / cell);
        c.aofc[i] = (float)(of / cell + 1.0);
        const double span = std::fabs((double)c.cb0[i]) + std::fabs((double)c.cb1[i]);
        if (!(span * std::fabs(sc) * 8.0 * 1.1920929e-07 < 0.25) || !(n < 2097152.0)) c.use_wide = 0;
    }
    c.cnx = (((uint32_t)c.cloud.nbx - 1u) >> shift) + 3u;
    const uint32_t cny = (((uint32_t)c.cloud.nby - 1u) >> shift) + 3u, cnz = (((uint32_t)c.cloud.nbz - 1u) >> shift) + 3u;
    c.cnxy = c.cnx * cny;
    c.cbias = 0x4B000000u * (1u + c.cnx + c.cnxy);
    c.cmax = c.cnxy * cnz - 1u;
    if (!c.use_wide) {          // scale 0 and a bias that selects the never-reject byte behind the table
        for (int i = 0; i < 3; ++i) c.ascc[i] = c.aofc[i] = 0.0f;
        c.cmax += 1u;
        c.cbias -= c.cmax;
    }
}

// ---------------------------------------------------------------------------------------------------
// RenderTechniquePPM, photon-tracing pass
// ---------------------------------------------------------------------------------------------------
struct crv_ppm {
    crv_ctx* ctx = nullptr;
    const crv_photon_map_properties* ctor_props = nullptr;   // borrowed, like m_photonMapProperties (RenderTechniquePPM.cpp:12)
    crv_push_constants* pc = nullptr;                        // borrowed; UpdateRadius writes pmRadius into it (:302-312)
    float initial_radius = 10.0f;                            // main.cpp:813
    float alpha = 0.8f;                                      // RenderTechniquePPM.h:66
    // estimate-pass bindings
    crv_camera_properties cam{};
    crv_shadow_volume_properties sv{};
    crv_volume* sv_vol = nullptr;
    bool has_cam = false, has_sv = false;
    const crv_camera_properties* q_cam = nullptr;
    const crv_shadow_volume_properties* q_sv = nullptr;
    crv_volume* q_sv_vol = nullptr;
    std::vector<crv_image*> images;
    cudaEvent_t ev2 = nullptr, ev3 = nullptr;
    bool timed_estimate = false;
    // committed / queued bindings
    crv_photon_map_properties props{};
    crv_cloud_properties cloud{};
    crv_parameters params{};
    crv_volume* cloud_vol = nullptr;
    bool has_cloud = false, has_params = false;
    const crv_cloud_properties* q_cloud = nullptr;
    const crv_parameters* q_params = nullptr;
    crv_volume* q_cloud_vol = nullptr;
    // resources
    crv_photon* photons = nullptr;
    uint32_t* collisions = nullptr;
    unsigned long long cells = 0;
    float* qtable = nullptr;
    size_t qtable_elems = 0;
    uint32_t q_dims[3] = {0, 0, 0};   // brick counts per axis the tables were sized for
    uint8_t* qcoarse = nullptr;               // shared-memory threshold table, like the path tracer's
    uint32_t qcoarse_bytes = 0;
    int qcoarse_shift = 0;
    unsigned long long* stats = nullptr;      // 3 u64
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    bool timed = false;
};

int crv_ppm_create(crv_ctx* ctx, const crv_photon_map_properties* props, crv_push_constants* pc, float initial_radius, crv_ppm** out)
{
    CRV_REQUIRE(ctx && props && pc && out, "crv_ppm_create: null argument");
    CRV_ENTER(ctx);
    *out = nullptr;
    crv_ppm* p = new (std::nothrow) crv_ppm();
    if (!p) { crv_set_error("out of host memory"); return CRV_ERR_OOM; }
    p->ctx = ctx;
    p->ctor_props = props;
    p->pc = pc;
    p->props = *props;
    p->initial_radius = initial_radius;
    pc->pmRadius = initial_radius;                  // RenderTechniquePPM.cpp:15
    if (cudaMalloc(&p->stats, 4 * sizeof(unsigned long long)) != cudaSuccess || cudaEventCreate(&p->ev0) != cudaSuccess ||
        cudaEventCreate(&p->ev1) != cudaSuccess || cudaEventCreate(&p->ev2) != cudaSuccess || cudaEventCreate(&p->ev3) != cudaSuccess) {
        cudaGetLastError();
        crv_ppm_destroy(p);
        crv_set_error("crv_ppm_create: device allocation failed");
        return CRV_ERR_CUDA;
    }
    *out = p;
    return CRV_OK;
}

int crv_ppm_free_resources(crv_ppm* ppm)
{
    CRV_REQUIRE(ppm != nullptr, "crv_ppm_free_resources: null technique");
    CRV_ENTER(ppm->ctx);
    if (ppm->photons) CRV_CUDA(cudaFree(ppm->photons));
    if (ppm->collisions) CRV_CUDA(cudaFree(ppm->collisions));
    ppm->photons = nullptr;
    ppm->collisions = nullptr;
    ppm->cells = 0;
    return CRV_OK;
}

int crv_ppm_destroy(crv_ppm* ppm)
{
    if (!ppm) return CRV_OK;
    cudaSetDevice(ppm->ctx->device);
    if (ppm->photons) cudaFree(ppm->photons);
    if (ppm->collisions) cudaFree(ppm->collisions);
    if (ppm->qtable) cudaFree(ppm->qtable);
    if (ppm->qcoarse) cudaFree(ppm->qcoarse);
    if (ppm->stats) cudaFree(ppm->stats);
    if (ppm->ev0) cudaEventDestroy(ppm->ev0);
    if (ppm->ev1) cudaEventDestroy(ppm->ev1);
    if (ppm->ev2) cudaEventDestroy(ppm->ev2);
    if (ppm->ev3) cudaEventDestroy(ppm->ev3);
    delete ppm;
    return CRV_OK;
}

int crv_ppm_allocate_resources(crv_ppm* ppm)
{
    CRV_REQUIRE(ppm != nullptr, "crv_ppm_allocate_resources: null technique");
    CRV_TRY(crv_ppm_free_resources(ppm));
    CRV_ENTER(ppm->ctx);
    const crv_photon_map_properties& pr = *ppm->ctor_props;
    CRV_REQUIRE(pr.voxelCount[0] > 0 && pr.voxelCount[1] > 0 && pr.voxelCount[2] > 0, "RenderTechniquePPM: empty photon map");
    const unsigned long long cells = (unsigned long long)pr.voxelCount[0] * pr.voxelCount[1] * pr.voxelCount[2];
    CRV_REQUIRE(cells <= 0x7fffffffull, "RenderTechniquePPM: photon map has more than 2^31 cells");
    CRV_CUDA(cudaMalloc(&ppm->photons, cells * PPM_ELEMENTS_PER_CELL * sizeof(crv_photon)));
    CRV_CUDA(cudaMalloc(&ppm->collisions, cells * 4 * sizeof(uint32_t)));
    CRV_CUDA(cudaMemset(ppm->photons, 0, cells * PPM_ELEMENTS_PER_CELL * sizeof(crv_photon)));
    CRV_CUDA(cudaMemset(ppm->collisions, 0, cells * 4 * sizeof(uint32_t)));
    ppm->cells = cells;
    return CRV_OK;
}

int crv_ppm_queue_update_cloud_data(crv_ppm* ppm, const crv_cloud_properties* cloud)
{
    CRV_REQUIRE(ppm && cloud, "crv_ppm_queue_update_cloud_data: null argument");
    ppm->q_cloud = cloud;
    return CRV_OK;
}
int crv_ppm_queue_update_cloud_data_sampler(crv_ppm* ppm, crv_volume* cloud_image)
{
    CRV_REQUIRE(ppm && cloud_image, "crv_ppm_queue_update_cloud_data_sampler: null argument");
    ppm->q_cloud_vol = cloud_image;
    return CRV_OK;
}
int crv_ppm_queue_update_parameters(crv_ppm* ppm, const crv_parameters* params)
{
    CRV_REQUIRE(ppm && params, "crv_ppm_queue_update_parameters: null argument");
    ppm->q_params = params;
    return CRV_OK;
}
int crv_ppm_set_frame_references(crv_ppm* ppm, crv_image* const* result_images, unsigned count)
{
    CRV_REQUIRE(ppm && (result_images || count == 0), "crv_ppm_set_frame_references: null argument");
    ppm->images.assign(result_images, result_images + count);
    return CRV_OK;
}
int crv_ppm_queue_update_camera_properties(crv_ppm* ppm, const crv_camera_properties* cam)
{
    CRV_REQUIRE(ppm && cam, "crv_ppm_queue_update_camera_properties: null argument");
    ppm->q_cam = cam;
    return CRV_OK;
}
int crv_ppm_queue_update_shadow_volume(crv_ppm* ppm, const crv_shadow_volume_properties* p)
{
    CRV_REQUIRE(ppm && p, "crv_ppm_queue_update_shadow_volume: null argument");
    ppm->q_sv = p;
    return CRV_OK;
}
int crv_ppm_queue_update_shadow_volume_sampler(crv_ppm* ppm, crv_volume* shadow_image)
{
    CRV_REQUIRE(ppm && shadow_image, "crv_ppm_queue_update_shadow_volume_sampler: null argument");
    ppm->q_sv_vol = shadow_image;
    return CRV_OK;
}
int crv_ppm_update_descriptor_sets(crv_ppm* ppm)
{
    CRV_REQUIRE(ppm != nullptr, "crv_ppm_update_descriptor_sets: null technique");
    if (ppm->q_cam) { ppm->cam = *ppm->q_cam; ppm->has_cam = true; ppm->q_cam = nullptr; }
    if (ppm->q_sv) { ppm->sv = *ppm->q_sv; ppm->has_sv = true; ppm->q_sv = nullptr; }
    if (ppm->q_sv_vol) { ppm->sv_vol = ppm->q_sv_vol; ppm->q_sv_vol = nullptr; }
    if (ppm->q_cloud) { ppm->cloud = *ppm->q_cloud; ppm->has_cloud = true; ppm->q_cloud = nullptr; }
    if (ppm->q_cloud_vol) { ppm->cloud_vol = ppm->q_cloud_vol; ppm->q_cloud_vol = nullptr; }
    if (ppm->q_params) { ppm->params = *ppm->q_params; ppm->has_params = true; ppm->q_params = nullptr; }
    ppm->props = *ppm->ctor_props;              // g_photonMapPropertiesBuffer->SetData() (main.cpp:300)
    return CRV_OK;
}

// kernel-parameter block shared by the two dispatches: cloud, phase and photon-map parameters, tables
static int ppm_build_const(crv_ppm* ppm, cudaStream_t s, PPMConst& k)
{
    CRV_REQUIRE(ppm->has_cloud && ppm->cloud_vol, "RenderTechniquePPM: cloud data / sampler not bound");
    CRV_REQUIRE(ppm->has_params, "RenderTechniquePPM: Parameters not bound");
    CRV_REQUIRE(ppm->photons && ppm->collisions, "RenderTechniquePPM: AllocateResources has not been called");
    const crv_photon_map_properties& pr = ppm->props;
    CRV_REQUIRE((unsigned long long)pr.voxelCount[0] * pr.voxelCount[1] * pr.voxelCount[2] == ppm->cells,
                "RenderTechniquePPM: PhotonMapProperties.voxelCount changed since AllocateResources");
    CRV_REQUIRE(ppm->cloud.maxExtinction > 0.0f, "RenderTechniquePPM: maxExtinction must be positive");
    std::memset(&k, 0, sizeof(k));
    PTConst& c = k.base;
    const crv_cloud_properties& cp = ppm->cloud;
    for (int i = 0; i < 3; ++i) {
        c.cb0[i] = cp.bounds[0][i];
        c.cb1[i] = cp.bounds[1][i];
        c.cext[i] = cp.bounds[1][i] - cp.bounds[0][i];
        k.pm_light[i] = pr.lightDirection[i];
        k.pm_b0[i] = pr.bounds[0][i];
        k.pm_count[i] = pr.voxelCount[i];
    }
    c.max_ext = cp.maxExtinction;
    c.sigma = (cp.maxExtinction * cp.densityScaling) / cp.baseScaling;
    c.density_scaling = cp.densityScaling;
    c.light_intensity = 1.0f * ppm->params.lightIntensity;
    c.two_g = 2.0f * ppm->params.phaseG;
    c.one_minus_g = 1.0f - ppm->params.phaseG;
    c.one_plus_g2 = ppm->params.phaseOnePlusG2;
    c.one_minus_g2 = ppm->params.phaseOneMinusG2;
    c.one_over_2g = ppm->params.phaseOneOver2G;
    c.isotropic = ppm->params.isotropic != 0;
    {
        auto ordinary = [](float v) { return std::isfinite(v) && std::fabs(v) >= 1e-18f && std::fabs(v) <= 1e18f; };
        bool ok = ordinary(c.sigma) && ordinary(c.max_ext);
        for (int i = 0; i < 3; ++i) ok = ok && ordinary(c.cext[i]);
        CRV_REQUIRE(ok || !(c.density_scaling > 0.0f), "RenderTechniquePPM: extinction scale and cloud extent must be finite and within 1e-18..1e18");
        c.r_sigma = 1.0f / c.sigma;
        c.r_max_ext = 1.0f / c.max_ext;
        for (int i = 0; i < 3; ++i) c.rcext[i] = 1.0f / c.cext[i];
    }
    crv_volume* cv = ppm->cloud_vol;
    c.cloud = cv->view();
    CRV_REQUIRE((uint32_t)c.cloud.nx == cp.voxelCount[0] && (uint32_t)c.cloud.ny == cp.voxelCount[1] && (uint32_t)c.cloud.nz == cp.voxelCount[2],
                "RenderTechniquePPM: cloud image size differs from CloudProperties.voxelCount");
    const size_t nbricks = (size_t)cv->nbx * cv->nby * cv->nbz;
    if (!cv->brick_max_valid) CRV_TRY(crv_launch_brick_max(ppm->ctx, cv, s));
    // (keyed on the three brick counts: the padded coarse table depends on them separately)
    if (ppm->qtable_elems != nbricks || ppm->q_dims[0] != cv->nbx || ppm->q_dims[1] != cv->nby || ppm->q_dims[2] != cv->nbz) {
        if (ppm->qtable) CRV_CUDA(cudaFree(ppm->qtable));
        ppm->qtable = nullptr;
        ppm->qtable_elems = 0;
        CRV_CUDA(cudaMalloc(&ppm->qtable, nbricks * sizeof(float)));
        if (ppm->qcoarse) CRV_CUDA(cudaFree(ppm->qcoarse));
        ppm->qcoarse = nullptr;
        coarse_table_shape(cv, 0, ppm->qcoarse_shift, ppm->qcoarse_bytes);
        CRV_CUDA(cudaMalloc(&ppm->qcoarse, ppm->qcoarse_bytes));
        ppm->qtable_elems = nbricks;
        ppm->q_dims[0] = cv->nbx; ppm->q_dims[1] = cv->nby; ppm->q_dims[2] = cv->nbz;
    }
    CRV_CUDA(cudaMemsetAsync(ppm->qcoarse, 128, ppm->qcoarse_bytes, s));
    CRV_TRY(crv_launch_qcoarse(ppm->ctx, cv->brick_absmax_wide, ppm->qcoarse, (int)cv->nbx, (int)cv->nby, (int)cv->nbz, ppm->qcoarse_shift, c.max_ext, s));
    c.qcoarse = ppm->qcoarse;
    c.qcoarse_bytes = ppm->qcoarse_bytes;
    coarse_pretest_geometry(c, ppm->qcoarse_shift);
    CRV_TRY(crv_launch_qtable(ppm->ctx, cv->brick_absmax, ppm->qtable, (long long)nbricks, c.max_ext, s));   // 1 MB: rebuilt every dispatch
    c.qtable = ppm->qtable;
    k.pm_voxel_size = pr.voxelSize;
    k.pm_absorption = pr.absorption;
    k.seed = ppm->pc->seed;                    // vkCmdPushConstants snapshot (RenderTechniquePPM.cpp:243)
    k.photons = ppm->photons;
    k.collisions = ppm->collisions;
    k.stats = ppm->stats;
    k.cells = ppm->cells;

    k.pm_step_size = pr.stepSize;
    k.base_scaling = cp.baseScaling;
    return CRV_OK;
}

int crv_ppm_record_photon_tracing(crv_ppm* ppm, void* stream, uint32_t groups_x)
{
    CRV_REQUIRE(ppm != nullptr, "crv_ppm_record_photon_tracing: null technique");
    CRV_ENTER(ppm->ctx);
    if (groups_x == 0) groups_x = 200;          // vkCmdDispatch(commandBuffer, 200, 1, 1) (RenderTechniquePPM.cpp:253)
    CRV_REQUIRE(groups_x <= (1u << 24), "RenderTechniquePPM: more than 2^30 photons per dispatch");
    cudaStream_t s = as_stream(stream);
    PPMConst k;
    CRV_TRY(ppm_build_const(ppm, s, k));
    k.n_photons = groups_x * 64u;
    CRV_CUDA(cudaEventRecord(ppm->ev0, s));
    CRV_CUDA(cudaMemsetAsync(ppm->stats, 0, 4 * sizeof(unsigned long long), s));
    CRV_TRY(crv_launch_ppm_clear(ppm->ctx, ppm->photons, ppm->collisions, ppm->cells, s));
    CRV_TRY(crv_launch_ppm_trace(ppm->ctx, k, ppm->cloud_vol->layout, s));
    CRV_CUDA(cudaEventRecord(ppm->ev1, s));
    ppm->timed = true;
    return CRV_OK;
}

// The estimate half of RecordDrawCommands (RenderTechniquePPM.cpp:260-273): PPM_PE.comp over the current photon map into
// result image `image_index`, with seed / frameCount / pmRadius read from the PushConstants now.
int crv_ppm_record_photon_estimate(crv_ppm* ppm, void* stream, unsigned image_index)
{
    CRV_REQUIRE(ppm != nullptr, "crv_ppm_record_photon_estimate: null technique");
    CRV_ENTER(ppm->ctx);
    CRV_REQUIRE(image_index < ppm->images.size() && ppm->images[image_index], "RenderTechniquePPM: no result image (binding 0)");
    CRV_REQUIRE(ppm->has_cam, "RenderTechniquePPM: CameraProperties not bound (binding 1)");
    CRV_REQUIRE(ppm->has_sv && ppm->sv_vol, "RenderTechniquePPM: shadow volume / sampler not bound (bindings 8,9)");
    cudaStream_t s = as_stream(stream);
    PPMConst k;
    CRV_TRY(ppm_build_const(ppm, s, k));
    PTConst& c = k.base;
    crv_image* img = ppm->images[image_index];
    const int W = ppm->cam.halfWidth * 2, H = ppm->cam.halfHeight * 2;
    CRV_REQUIRE(W > 0 && H > 0 && (int)img->width == W && (int)img->height == H, "RenderTechniquePPM: result image size differs from CameraProperties resolution");
    const uint32_t nsv = ppm->sv.voxelAxisCount;
    CRV_REQUIRE(ppm->sv_vol->nx == nsv && ppm->sv_vol->ny == nsv && ppm->sv_vol->nz == nsv, "RenderTechniquePPM: shadow image is not voxelAxisCount^3");
    float d[4];
    for (int i = 0; i < 3; ++i) {
        c.cam_pos[i] = ppm->cam.position[i];
        c.fwd_near[i] = ppm->cam.forward[i] * ppm->cam.nearPlane;
        c.right_psx[i] = ppm->cam.right[i] * ppm->cam.pixelSizeX;
        c.up_psy[i] = ppm->cam.up[i] * ppm->cam.pixelSizeY;
        c.light[i] = ppm->sv.lightDirection[i];
        c.m0[i] = ppm->sv.basisChange[0][i];
        c.m1[i] = ppm->sv.basisChange[1][i];
        c.m2[i] = ppm->sv.basisChange[2][i];
        c.m3[i] = ppm->sv.basisChange[3][i];
    }
    for (int i = 0; i < 4; ++i) {
        c.sb0[i] = ppm->sv.bounds[0][i];
        d[i] = ppm->sv.bounds[1][i] - ppm->sv.bounds[0][i];
    }
    for (int i = 0; i < 3; ++i) {
        c.sden[i] = ((ppm->sv.basisChange[0][i] * d[0] + ppm->sv.basisChange[1][i] * d[1]) + ppm->sv.basisChange[2][i] * d[2]) + ppm->sv.basisChange[3][i] * d[3];
        CRV_REQUIRE(std::isfinite(c.sden[i]) && std::fabs(c.sden[i]) >= 1e-18f && std::fabs(c.sden[i]) <= 1e18f,
                    "RenderTechniquePPM: shadow-volume extent must be finite and within 1e-18..1e18");
        c.rsden[i] = 1.0f / c.sden[i];
    }
    c.half_w = ppm->cam.halfWidth;
    c.half_h = ppm->cam.halfHeight;
    c.width = W;
    c.height = H;
    c.row_stride = 32u * (uint32_t)(W / 32 + 1);
    c.inv4pi_omg2 = (1.0f / (4.0f * 3.14159265359f)) * ppm->params.phaseOneMinusG2;
    c.sv = ppm->sv_vol->view();
    k.pm_radius = ppm->pc->pmRadius;
    k.frame_count = ppm->pc->frameCount;
    k.image = reinterpret_cast<float4*>(img->rgba);
    CRV_CUDA(cudaEventRecord(ppm->ev2, s));
    CRV_TRY(crv_launch_ppm_estimate(ppm->ctx, k, ppm->cloud_vol->layout, ppm->sv_vol->layout, s));
    CRV_CUDA(cudaEventRecord(ppm->ev3, s));
    ppm->timed_estimate = true;
    return CRV_OK;
}

// UpdateRadius (RenderTechniquePPM.cpp:302-312): the progressive radius, written into the caller's PushConstants
int crv_ppm_update_radius(crv_ppm* ppm, unsigned frame_number)
{
    CRV_REQUIRE(ppm != nullptr, "crv_ppm_update_radius: null technique");
    if (frame_number <= 1) ppm->pc->pmRadius = ppm->initial_radius;
    else ppm->pc->pmRadius = ppm->pc->pmRadius * std::pow(((float)(frame_number - 1u) + ppm->alpha) / (float)frame_number, .33333f);
    return CRV_OK;
}

// RenderTechniquePPM::RecordDrawCommands (:229-273) without the swapchain blit: radius update, tracing, estimate
int crv_ppm_record_draw_commands(crv_ppm* ppm, void* stream, unsigned image_index)
{
    CRV_REQUIRE(ppm != nullptr, "crv_ppm_record_draw_commands: null technique");
    CRV_TRY(crv_ppm_update_radius(ppm, ppm->pc->frameCount));
    CRV_TRY(crv_ppm_record_photon_tracing(ppm, stream, 0));
    return crv_ppm_record_photon_estimate(ppm, stream, image_index);
}

// test / tooling hook: replace the photon map by host data (cells*32 photons, cells*4 u32)
int crv_ppm_upload(crv_ppm* ppm, const crv_photon* host_photons, const uint32_t* host_collisions, void* stream)
{
    CRV_REQUIRE(ppm && host_photons && host_collisions, "crv_ppm_upload: null argument");
    CRV_ENTER(ppm->ctx);
    CRV_REQUIRE(ppm->photons && ppm->collisions, "crv_ppm_upload: no resources allocated");
    cudaStream_t s = as_stream(stream);
    CRV_CUDA(cudaMemcpyAsync(ppm->photons, host_photons, ppm->cells * PPM_ELEMENTS_PER_CELL * sizeof(crv_photon), cudaMemcpyHostToDevice, s));
    CRV_CUDA(cudaMemcpyAsync(ppm->collisions, host_collisions, ppm->cells * 4 * sizeof(uint32_t), cudaMemcpyHostToDevice, s));
    CRV_CUDA(cudaStreamSynchronize(s));
    return CRV_OK;
}

int crv_ppm_last_estimate_stats(crv_ppm* ppm, uint64_t* march_steps, float* ms)
{
    CRV_REQUIRE(ppm != nullptr, "crv_ppm_last_estimate_stats: null technique");
    CRV_ENTER(ppm->ctx);
    CRV_REQUIRE(ppm->timed_estimate, "crv_ppm_last_estimate_stats: no estimate dispatch recorded yet");
    CRV_CUDA(cudaEventSynchronize(ppm->ev3));
    unsigned long long v = 0;
    CRV_CUDA(cudaMemcpy(&v, ppm->stats + 3, sizeof(v), cudaMemcpyDeviceToHost));
    if (march_steps) *march_steps = v;
    if (ms) CRV_CUDA(cudaEventElapsedTime(ms, ppm->ev2, ppm->ev3));
    return CRV_OK;
}

int crv_ppm_download(crv_ppm* ppm, crv_photon* host_photons, uint32_t* host_collisions, void* stream)
{
    CRV_REQUIRE(ppm != nullptr, "crv_ppm_download: null technique");
    CRV_ENTER(ppm->ctx);
    CRV_REQUIRE(ppm->photons && ppm->collisions, "crv_ppm_download: no resources allocated");
    cudaStream_t s = as_stream(stream);
    if (host_photons) CRV_CUDA(cudaMemcpyAsync(host_photons, ppm->photons, ppm->cells * PPM_ELEMENTS_PER_CELL * sizeof(crv_photon), cudaMemcpyDeviceToHost, s));
    if (host_collisions) CRV_CUDA(cudaMemcpyAsync(host_collisions, ppm->collisions, ppm->cells * 4 * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    CRV_CUDA(cudaStreamSynchronize(s));
    return CRV_OK;
}

int crv_ppm_last_trace_stats(crv_ppm* ppm, uint64_t* steps, uint64_t* deposits, uint64_t* dropped, float* ms)
{
    CRV_REQUIRE(ppm != nullptr, "crv_ppm_last_trace_stats: null technique");
    CRV_ENTER(ppm->ctx);
    CRV_REQUIRE(ppm->timed, "crv_ppm_last_trace_stats: no dispatch recorded yet");
    CRV_CUDA(cudaEventSynchronize(ppm->ev1));
    unsigned long long v[3];
    CRV_CUDA(cudaMemcpy(v, ppm->stats, sizeof(v), cudaMemcpyDeviceToHost));
    if (steps) *steps = v[0];
    if (deposits) *deposits = v[1];
    if (dropped) *dropped = v[2];
    if (ms) CRV_CUDA(cudaEventElapsedTime(ms, ppm->ev0, ppm->ev1));
    return CRV_OK;
}

int crv_pt_destroy(crv_pt* pt)
{
    if (!pt) return CRV_OK;
    cudaSetDevice(pt->ctx->device);
    if (pt->qtable) cudaFree(pt->qtable);
    if (pt->qtable_wide) cudaFree(pt->qtable_wide);
    if (pt->qcoarse) cudaFree(pt->qcoarse);
    if (pt->rcp_fc1) cudaFree(pt->rcp_fc1);
    if (pt->hitlist) cudaFree(pt->hitlist);
    if (pt->start) cudaFree(pt->start);
    if (pt->nhit) cudaFree(pt->nhit);
    if (pt->counts_host) cudaFreeHost(pt->counts_host);
    if (pt->counts_ev) cudaEventDestroy(pt->counts_ev);
    if (pt->occ_aabb) cudaFree(pt->occ_aabb);
    if (pt->misslist) cudaFree(pt->misslist);
    if (pt->missdy) cudaFree(pt->missdy);
    if (pt->hit_keys) cudaFree(pt->hit_keys);
    if (pt->hitlist_sorted) cudaFree(pt->hitlist_sorted);
    if (pt->start_sorted) cudaFree(pt->start_sorted);
    if (pt->sorted_from) cudaFree(pt->sorted_from);
    if (pt->sort_scratch) cudaFree(pt->sort_scratch);
    if (pt->samples) cudaFree(pt->samples);
    if (pt->job_counter) cudaFree(pt->job_counter);
    if (pt->stats) cudaFree(pt->stats);
    if (pt->dbg_steps) cudaFree(pt->dbg_steps);
    if (pt->dbg_coll) cudaFree(pt->dbg_coll);
    for (cudaEvent_t e : pt->ev) cudaEventDestroy(e);
    delete pt;
    return CRV_OK;
}

int crv_pt_set_frame_references(crv_pt* pt, crv_image* const* result_images, unsigned count)
{
    CRV_REQUIRE(pt && result_images && count > 0, "crv_pt_set_frame_references: bad argument");
    pt->images.assign(result_images, result_images + count);
    if (pt->sets.size() < count) pt->sets.resize(count);
    return CRV_OK;
}
int crv_pt_clear_frame_references(crv_pt* pt)
{
    CRV_REQUIRE(pt != nullptr, "crv_pt_clear_frame_references: null technique");
    pt->images.clear();
    return CRV_OK;
}

static PTBindings* pt_set(crv_pt* pt, unsigned idx)
{
    if (!pt) return nullptr;
    if (idx >= pt->sets.size()) {
        if (idx >= 64) return nullptr;
        pt->sets.resize(idx + 1);
    }
    return &pt->sets[idx];
}

int crv_pt_queue_update_cloud_data(crv_pt* pt, const crv_cloud_properties* cloud, unsigned idx)
{
    PTBindings* b = pt_set(pt, idx);
    CRV_REQUIRE(b && cloud, "crv_pt_queue_update_cloud_data: bad argument");
    b->q_cloud = cloud;
    return CRV_OK;
}
int crv_pt_queue_update_cloud_data_sampler(crv_pt* pt, crv_volume* vol, unsigned idx)
{
    PTBindings* b = pt_set(pt, idx);
    CRV_REQUIRE(b && vol, "crv_pt_queue_update_cloud_data_sampler: bad argument");
    b->q_cloud_vol = vol;
    return CRV_OK;
}
int crv_pt_queue_update_camera_properties(crv_pt* pt, const crv_camera_properties* cam, unsigned idx)
{
    PTBindings* b = pt_set(pt, idx);
    CRV_REQUIRE(b && cam, "crv_pt_queue_update_camera_properties: bad argument");
    b->q_cam = cam;
    return CRV_OK;
}
int crv_pt_queue_update_parameters(crv_pt* pt, const crv_parameters* params, unsigned idx)
{
    PTBindings* b = pt_set(pt, idx);
    CRV_REQUIRE(b && params, "crv_pt_queue_update_parameters: bad argument");
    b->q_params = params;
    return CRV_OK;
}
int crv_pt_queue_update_shadow_volume(crv_pt* pt, const crv_shadow_volume_properties* p, unsigned idx)
{
    PTBindings* b = pt_set(pt, idx);
    CRV_REQUIRE(b && p, "crv_pt_queue_update_shadow_volume: bad argument");
    b->q_sv = p;
    return CRV_OK;
}
int crv_pt_queue_update_shadow_volume_sampler(crv_pt* pt, crv_volume* vol, unsigned idx)
{
    PTBindings* b = pt_set(pt, idx);
    CRV_REQUIRE(b && vol, "crv_pt_queue_update_shadow_volume_sampler: bad argument");
    b->q_sv_vol = vol;
    return CRV_OK;
}

int crv_pt_update_descriptor_sets(crv_pt* pt)
{
    CRV_REQUIRE(pt != nullptr, "crv_pt_update_descriptor_sets: null technique");
    for (PTBindings& b : pt->sets) {
        if (b.q_cloud) { b.cloud = *b.q_cloud; b.has_cloud = true; b.q_cloud = nullptr; }
        if (b.q_cloud_vol) { b.cloud_vol = b.q_cloud_vol; b.q_cloud_vol = nullptr; }
        if (b.q_cam) { b.cam = *b.q_cam; b.has_cam = true; b.q_cam = nullptr; }
        if (b.q_params) { b.params = *b.q_params; b.has_params = true; b.q_params = nullptr; }
        if (b.q_sv) { b.sv = *b.q_sv; b.has_sv = true; b.q_sv = nullptr; }
        if (b.q_sv_vol) { b.sv_vol = b.q_sv_vol; b.q_sv_vol = nullptr; }
    }
    return CRV_OK;
}

int crv_pt_set_extensions(crv_pt* pt, const crv_pt_extensions* ext)
{
    CRV_REQUIRE(pt && ext, "crv_pt_set_extensions: null argument");
    CRV_REQUIRE(ext->absorption >= 0.0f && ext->absorption <= 1.0f, "crv_pt_set_extensions: absorption must be in [0,1]");
    CRV_REQUIRE(ext->math_mode == CRV_PT_MATH_STRICT || ext->math_mode == CRV_PT_MATH_DRIVER, "crv_pt_set_extensions: unknown math_mode");
    pt->ext = *ext;
    return CRV_OK;
}

int crv_pt_set_tuning(crv_pt* pt, int ctas_per_sm, int track_threshold, uint64_t sample_buffer_bytes, int steps_per_decision)
{
    CRV_REQUIRE(pt != nullptr, "crv_pt_set_tuning: null technique");
    pt->ctas_per_sm = ctas_per_sm > 0 ? ctas_per_sm : 0;
    if (track_threshold > 0) pt->track_threshold = track_threshold > 32 ? 32 : track_threshold;
    if (steps_per_decision > 0) pt->steps_per_decision = steps_per_decision > 64 ? 64 : steps_per_decision;
    if (sample_buffer_bytes > 0) pt->sample_budget = sample_buffer_bytes;
    return CRV_OK;
}

int crv_pt_set_skip_level(crv_pt* pt, int level)
{
    CRV_REQUIRE(pt != nullptr, "crv_pt_set_skip_level: null technique");
    CRV_REQUIRE(level == 0 || level == 1, "crv_pt_set_skip_level: level must be 0 or 1");
    pt->skip_level = level;
    return CRV_OK;
}

int crv_pt_set_coarse_shift(crv_pt* pt, int min_shift)
{
    CRV_REQUIRE(pt != nullptr, "crv_pt_set_coarse_shift: null technique");
    CRV_REQUIRE(min_shift >= -1 && min_shift <= 6, "crv_pt_set_coarse_shift: shift must be -1 (automatic) or 0..6");
    pt->forced_shift = min_shift;
    return CRV_OK;
}

int crv_pt_set_job_order(crv_pt* pt, int order)
{
    CRV_REQUIRE(pt != nullptr, "crv_pt_set_job_order: null technique");
    CRV_REQUIRE(order >= -1 && order <= 1, "crv_pt_set_job_order: order must be -1 (automatic), 0 (frame-major) or 1 (pixel-major, sorted)");
    pt->job_order = order;
    return CRV_OK;
}

int crv_pt_get_coarse_shift(crv_pt* pt, int* shift)
{
    CRV_REQUIRE(pt && shift, "crv_pt_get_coarse_shift: null argument");
    *shift = pt->qcoarse_shift;
    return CRV_OK;
}

int crv_pt_set_scheduler(crv_pt* pt, int scheduler)
{
    CRV_REQUIRE(pt != nullptr, "crv_pt_set_scheduler: null technique");
    CRV_REQUIRE(scheduler == CRV_PT_SCHED_POOL || scheduler == CRV_PT_SCHED_LANES, "crv_pt_set_scheduler: unknown scheduler");
    pt->scheduler = scheduler;
    return CRV_OK;
}

// Build the kernel-parameter block from the committed bindings.  fp32, shader evaluation order.
static int pt_build_const(crv_pt* pt, unsigned idx, cudaStream_t s, PTConst& c)
{
    CRV_REQUIRE(idx < pt->sets.size() && idx < pt->images.size(), "RenderTechniquePT: image index out of range");
    PTBindings& b = pt->sets[idx];
    crv_image* img = pt->images[idx];
    CRV_REQUIRE(img != nullptr, "RenderTechniquePT: no result image (binding 0)");
    CRV_REQUIRE(b.has_cam, "RenderTechniquePT: CameraProperties not bound (binding 1)");
    CRV_REQUIRE(b.has_cloud && b.cloud_vol, "RenderTechniquePT: cloud data / sampler not bound (bindings 2,3)");
    CRV_REQUIRE(b.has_params, "RenderTechniquePT: Parameters not bound (binding 4)");
    CRV_REQUIRE(b.has_sv && b.sv_vol, "RenderTechniquePT: shadow volume / sampler not bound (bindings 5,6)");
    const int W = pt->ctor_cam->halfWidth * 2, H = pt->ctor_cam->halfHeight * 2;
    CRV_REQUIRE(W > 0 && H > 0, "RenderTechniquePT: empty resolution");
    CRV_REQUIRE((int)img->width == W && (int)img->height == H, "RenderTechniquePT: result image size differs from CameraProperties resolution");
    const uint32_t nsv = b.sv.voxelAxisCount;
    CRV_REQUIRE(b.sv_vol->nx == nsv && b.sv_vol->ny == nsv && b.sv_vol->nz == nsv, "RenderTechniquePT: shadow image is not voxelAxisCount^3");
    CRV_REQUIRE(b.cloud.maxExtinction > 0.0f, "RenderTechniquePT: maxExtinction must be positive");

    std::memset(&c, 0, sizeof(c));
    for (int i = 0; i < 3; ++i) {
        c.cam_pos[i] = b.cam.position[i];
        c.fwd_near[i] = b.cam.forward[i] * b.cam.nearPlane;
        c.right_psx[i] = b.cam.right[i] * b.cam.pixelSizeX;
        c.up_psy[i] = b.cam.up[i] * b.cam.pixelSizeY;
        c.cb0[i] = b.cloud.bounds[0][i];
        c.cb1[i] = b.cloud.bounds[1][i];
        c.cext[i] = b.cloud.bounds[1][i] - b.cloud.bounds[0][i];
        c.light[i] = b.sv.lightDirection[i];
        c.m0[i] = b.sv.basisChange[0][i];
        c.m1[i] = b.sv.basisChange[1][i];
        c.m2[i] = b.sv.basisChange[2][i];
        c.m3[i] = b.sv.basisChange[3][i];
    }
    c.half_w = b.cam.halfWidth;
    c.half_h = b.cam.halfHeight;
    c.width = W;
    c.height = H;
    c.row_stride = 32u * (uint32_t)(W / 32 + 1);
    c.max_ext = b.cloud.maxExtinction;
    c.sigma = (b.cloud.maxExtinction * b.cloud.densityScaling) / b.cloud.baseScaling;
    c.density_scaling = b.cloud.densityScaling;
    c.light_intensity = 1.0f * b.params.lightIntensity;
    c.two_g = 2.0f * b.params.phaseG;
    c.one_minus_g = 1.0f - b.params.phaseG;
    c.one_plus_g2 = b.params.phaseOnePlusG2;
    c.one_minus_g2 = b.params.phaseOneMinusG2;
    c.one_over_2g = b.params.phaseOneOver2G;
    {
        const float PI = 3.14159265359f;
        const float INV_4Pi = 1.0f / (4.0f * PI);
        c.inv4pi_omg2 = INV_4Pi * b.params.phaseOneMinusG2;
    }
    c.isotropic = b.params.isotropic != 0;
    c.absorption = pt->ext.absorption;
    float d[4];
    for (int i = 0; i < 4; ++i) {
        c.sb0[i] = b.sv.bounds[0][i];
        d[i] = b.sv.bounds[1][i] - b.sv.bounds[0][i];
    }
    for (int i = 0; i < 3; ++i)
        c.sden[i] = ((b.sv.basisChange[0][i] * d[0] + b.sv.basisChange[1][i] * d[1]) + b.sv.basisChange[2][i] * d[2]) + b.sv.basisChange[3][i] * d[3];

    c.cloud = b.cloud_vol->view();
    c.sv = b.sv_vol->view();
    CRV_REQUIRE((uint32_t)c.cloud.nx == b.cloud.voxelCount[0] && (uint32_t)c.cloud.ny == b.cloud.voxelCount[1] && (uint32_t)c.cloud.nz == b.cloud.voxelCount[2],
                "RenderTechniquePT: cloud image size differs from CloudProperties.voxelCount");

    // q-table: rebuilt when the cloud image, its contents or maxExtinction changed
    crv_volume* cv = b.cloud_vol;
    const size_t nbricks = (size_t)cv->nbx * cv->nby * cv->nbz;
    if (!cv->brick_max_valid) CRV_TRY(crv_launch_brick_max(pt->ctx, cv, s));
    // (the padded tables and the coarse cell size depend on the three brick counts separately, not on their product)
    const bool same_dims = pt->qtable_elems == nbricks && pt->q_dims[0] == cv->nbx && pt->q_dims[1] == cv->nby && pt->q_dims[2] == cv->nbz &&
                           pt->q_forced_for == pt->forced_shift;
    if (pt->q_for_vol != cv || pt->q_for_generation != vol_generation(cv) || pt->q_for_max_ext != c.max_ext || !same_dims) {
        if (!same_dims) {
            if (pt->qtable) CRV_CUDA(cudaFree(pt->qtable));
            if (pt->qtable_wide) CRV_CUDA(cudaFree(pt->qtable_wide));
            if (pt->qcoarse) CRV_CUDA(cudaFree(pt->qcoarse));
            pt->qcoarse = nullptr;
            pt->qtable = pt->qtable_wide = nullptr;
            pt->qtable_elems = 0;
            CRV_CUDA(cudaMalloc(&pt->qtable, nbricks * sizeof(float)));
            CRV_CUDA(cudaMalloc(&pt->qtable_wide, ((size_t)(cv->nbx + 2) * (cv->nby + 2) * (cv->nbz + 2) + 1) * sizeof(float)));   // + the +inf entry
            // coarse table: the smallest cell size whose padded table fits the pool scheduler's shared-memory budget
            coarse_table_shape(cv, pt->forced_shift, pt->qcoarse_shift, pt->qcoarse_bytes);
            CRV_CUDA(cudaMalloc(&pt->qcoarse, pt->qcoarse_bytes));
            pt->qtable_elems = nbricks;
            pt->q_dims[0] = cv->nbx; pt->q_dims[1] = cv->nby; pt->q_dims[2] = cv->nbz;
            pt->q_forced_for = pt->forced_shift;
        }
        CRV_CUDA(cudaMemsetAsync(pt->qcoarse, 128, pt->qcoarse_bytes, s));
        CRV_TRY(crv_launch_qcoarse(pt->ctx, cv->brick_absmax_wide, pt->qcoarse, (int)cv->nbx, (int)cv->nby, (int)cv->nbz, pt->qcoarse_shift, c.max_ext, s));
        pt->launches += 1;
        CRV_TRY(crv_launch_qtable(pt->ctx, cv->brick_absmax, pt->qtable, (long long)nbricks, c.max_ext, s));
        CRV_TRY(crv_launch_qtable_padded(pt->ctx, cv->brick_absmax_wide, pt->qtable_wide, (int)cv->nbx, (int)cv->nby, (int)cv->nbz, c.max_ext, s));
        pt->launches += 2;
        pt->q_for_vol = cv;
        pt->q_for_generation = vol_generation(cv);
        pt->q_for_max_ext = c.max_ext;
    }
    c.qtable = pt->qtable;
    c.qtable_wide = pt->qtable_wide;
    c.qcoarse = pt->qcoarse;
    c.qcoarse_bytes = pt->qcoarse_bytes;

    // exact FMA division (exact_div.cuh): reciprocals by IEEE division on the host; the divisors must be ordinary numbers
    {
        auto ordinary = [](float v) { return std::isfinite(v) && std::fabs(v) >= 1e-18f && std::fabs(v) <= 1e18f; };
        bool ok = ordinary(c.sigma) && ordinary(c.max_ext);
        for (int i = 0; i < 3; ++i) ok = ok && ordinary(c.cext[i]) && ordinary(c.sden[i]);
        CRV_REQUIRE(ok || !(c.density_scaling > 0.0f),
                    "RenderTechniquePT: extinction scale, cloud extent and shadow-volume extent must be finite and within 1e-18..1e18");
        c.r_sigma = 1.0f / c.sigma;
        c.neg_ln2_over_sigma = (float)(-0.69314718055994530942 / (double)c.sigma);
        c.math_mode = (int)pt->ext.math_mode;
        c.r_max_ext = 1.0f / c.max_ext;
        // approximate brick pre-test: v = q*asc8 + aof8 ~ ((((q - cb0)/cext)*size - 0.5) / 8) + 1, the padded brick
        // coordinate, evaluated with FMAs.  Its error is a few ulps of the largest term, |cb0/cext|*size (in texels);
        // the wide table tolerates 1 texel, we allow 1/4.
        for (int i = 0; i < 3; ++i) {
            c.rcext[i] = 1.0f / c.cext[i];
            c.rsden[i] = 1.0f / c.sden[i];
            const double n = i == 0 ? c.cloud.nx : i == 1 ? c.cloud.ny : c.cloud.nz;
            const double sc = n / (double)c.cext[i], of = -(double)c.cb0[i] * sc - 0.5;
            c.asc8[i] = (float)(sc / 8.0);             // lane-owned scheduler: the same coordinate at brick granularity
            c.aof8[i] = (float)(of / 8.0 + 1.0);
        }
        coarse_pretest_geometry(c, pt->qcoarse_shift);  // pool scheduler: cell coordinate, strides, bias; decides c.use_wide
        // occupied spans (pt_classify_kernel): exact-in-spirit brick coordinate; only where the fp32 error bound above holds
        float min_texel = 3.0e38f;
        for (int i = 0; i < 3; ++i) {
            const double n = i == 0 ? c.cloud.nx : i == 1 ? c.cloud.ny : c.cloud.nz;
            const double sc = n / (double)c.cext[i], of = -(double)c.cb0[i] * sc - 0.5;
            c.bsc[i] = (float)(sc / 8.0);
            c.bof[i] = (float)(of / 8.0);
            min_texel = std::fmin(min_texel, (float)std::fabs((double)c.cext[i] / n));
        }
        c.span_margin = 0.25f * min_texel;
        c.occ_wide = c.use_wide ? cv->brick_absmax_wide : nullptr;
        c.occ_aabb = pt->occ_aabb;
        c.skip_level = (c.use_wide && pt->skip_level > 0) ? 1 : 0;
        c.pnbx = (uint32_t)c.cloud.nbx + 2u;
        c.pnbxy = c.pnbx * ((uint32_t)c.cloud.nby + 2u);
        c.pbias = 0x4B000000u * (1u + c.pnbx + c.pnbxy);          // mod 2^32, like the device-side sum it cancels
        c.pmax = c.pnbxy * ((uint32_t)c.cloud.nbz + 2u) - 1u;
        if (!c.use_wide) {
            // every approximate coordinate collapses to 0 and the bias moves the index to the +inf entry behind the
            // table: the pre-test never rejects and every tentative collision goes through the exact sampler
            for (int i = 0; i < 3; ++i) c.asc8[i] = c.aof8[i] = 0.0f;
            c.pmax += 1u;
            c.pbias -= c.pmax;
        }
    }

    c.npix = (uint32_t)W * (uint32_t)H;
    c.image = reinterpret_cast<float4*>(img->rgba);
    c.job_counter = pt->job_counter;
    c.stats = pt->stats;
    c.track_threshold = pt->track_threshold;
    c.steps_per_decision = pt->steps_per_decision > 0 ? pt->steps_per_decision
                         : (pt->scheduler == CRV_PT_SCHED_POOL ? crv_pt_pool_round_steps() : 6);
    c.scheduler = pt->scheduler;
    return CRV_OK;
}

// launch the tracker bracketed by a CUDA-event pair on the same stream
static int pt_timed_launch(crv_pt* pt, const PTConst& c, const PTBindings& b, cudaStream_t s)
{
    while (pt->ev.size() < pt->ev_used + 2) {
        cudaEvent_t e;
        CRV_CUDA(cudaEventCreate(&e));
        pt->ev.push_back(e);
    }
    CRV_CUDA(cudaEventRecord(pt->ev[pt->ev_used], s));
    CRV_TRY(crv_launch_pt(pt->ctx, c, b.cloud_vol->layout, b.sv_vol->layout, pt->ctas_per_sm, s));
    CRV_CUDA(cudaEventRecord(pt->ev[pt->ev_used + 1], s));
    pt->ev_used += 2;
    return CRV_OK;
}

static int pt_render(crv_pt* pt, const int32_t* seeds, unsigned n, uint32_t first_fc, int mode, cudaStream_t s, unsigned idx,
                     const crv_push_targets* push = nullptr, float others_scale = 0.0f)
{
    CRV_REQUIRE(n > 0 && seeds, "RenderTechniquePT: no frames to render");
    CRV_REQUIRE(mode == CRV_ACCUM_RUNNING_MEAN || mode == CRV_ACCUM_SUM, "RenderTechniquePT: unknown accumulation mode");
    PTConst c;
    CRV_TRY(pt_build_const(pt, idx, s, c));
    PTBindings& b = pt->sets[idx];
    c.accum_mode = mode;
    pt->ev_used = 0;

    // debug counters (single-frame renders only)
    if (n == 1 && pt->dbg_enabled) {
        if (pt->dbg_elems != c.npix) {
            if (pt->dbg_steps) CRV_CUDA(cudaFree(pt->dbg_steps));
            if (pt->dbg_coll) CRV_CUDA(cudaFree(pt->dbg_coll));
            pt->dbg_steps = pt->dbg_coll = nullptr;
            pt->dbg_elems = 0;
            CRV_CUDA(cudaMalloc(&pt->dbg_steps, sizeof(uint32_t) * c.npix));
            CRV_CUDA(cudaMalloc(&pt->dbg_coll, sizeof(uint32_t) * c.npix));
            pt->dbg_elems = c.npix;
        }
        c.dbg_steps = pt->dbg_steps;
        c.dbg_collisions = pt->dbg_coll;
        c.skip_level = 0;              // the per-pixel counters are the reference's: every step is executed
        // the pool scheduler accumulates per-path counts round by round
        CRV_CUDA(cudaMemsetAsync(pt->dbg_steps, 0, sizeof(uint32_t) * c.npix, s));
        CRV_CUDA(cudaMemsetAsync(pt->dbg_coll, 0, sizeof(uint32_t) * c.npix, s));
    }

    // classification: miss pixels get all n background samples folded now, hit pixels go to the hit list
    CRV_REQUIRE(c.width < 65536 && c.height < 65536, "RenderTechniquePT: resolution above 65535 is not supported");
    if (pt->hitlist_elems != c.npix) {
        if (pt->hitlist) CRV_CUDA(cudaFree(pt->hitlist));
        if (pt->start) CRV_CUDA(cudaFree(pt->start));
        if (pt->misslist) CRV_CUDA(cudaFree(pt->misslist));
        if (pt->missdy) CRV_CUDA(cudaFree(pt->missdy));
        if (pt->hit_keys) CRV_CUDA(cudaFree(pt->hit_keys));
        if (pt->hitlist_sorted) CRV_CUDA(cudaFree(pt->hitlist_sorted));
        if (pt->start_sorted) CRV_CUDA(cudaFree(pt->start_sorted));
        if (pt->sorted_from) CRV_CUDA(cudaFree(pt->sorted_from));
        pt->hitlist = nullptr;
        pt->start = nullptr;
        pt->misslist = nullptr;
        pt->missdy = nullptr;
        pt->hit_keys = nullptr;
        pt->hitlist_sorted = nullptr;
        pt->start_sorted = nullptr;
        pt->sorted_from = nullptr;
        pt->hitlist_elems = 0;
        pt->geom_valid = false;
        pt->sorted_valid = false;
        CRV_CUDA(cudaMalloc(&pt->hitlist, sizeof(uint32_t) * c.npix));
        CRV_CUDA(cudaMalloc(&pt->start, sizeof(float) * 8 * (size_t)c.npix));
        CRV_CUDA(cudaMalloc(&pt->misslist, sizeof(uint32_t) * c.npix));
        CRV_CUDA(cudaMalloc(&pt->missdy, sizeof(float) * c.npix));
        CRV_CUDA(cudaMalloc(&pt->hit_keys, sizeof(float) * c.npix));
        CRV_CUDA(cudaMalloc(&pt->hitlist_sorted, sizeof(uint32_t) * c.npix));
        CRV_CUDA(cudaMalloc(&pt->start_sorted, sizeof(float) * 8 * (size_t)c.npix));
        CRV_CUDA(cudaMalloc(&pt->sorted_from, sizeof(uint32_t) * c.npix));
        if (!pt->sort_scratch) CRV_CUDA(cudaMalloc(&pt->sort_scratch, sizeof(uint32_t) * (PT_SORT_BUCKETS + 1)));
        pt->hitlist_elems = c.npix;
    }
    c.hitlist = pt->hitlist;
    c.start = pt->start;
    c.nhit = pt->nhit;
    c.first_frame_count = first_fc;
    if (pt->rcp_elems < n) {
        if (pt->rcp_fc1) CRV_CUDA(cudaFree(pt->rcp_fc1));
        pt->rcp_fc1 = nullptr;
        pt->rcp_elems = 0;
        const size_t cap = n < 256 ? 256 : n;
        CRV_CUDA(cudaMalloc(&pt->rcp_fc1, sizeof(float) * cap));
        pt->rcp_elems = cap;
    }
    const bool rcp_table = n > 1 && mode == CRV_ACCUM_RUNNING_MEAN;      // a plain sum divides by nothing
    if (rcp_table) {                  // a single frame computes its one reciprocal in the fold kernel (same IEEE division)
        CRV_TRY(crv_launch_rcp_frames(pt->ctx, pt->rcp_fc1, first_fc, n, s));
        pt->launches++;
    }
    c.rcp_fc1 = rcp_table ? pt->rcp_fc1 : nullptr;
    {
        unsigned char key[sizeof(pt->geom_key)];
        std::memcpy(key, &b.cam, 68);
        std::memcpy(key + 68, b.cloud.bounds, 32);
        const uint32_t off = b.cloud.densityScaling <= 0.0f ? 1u : 0u;
        std::memcpy(key + 100, &off, 4);
        const uint64_t vol_id = (uint64_t)(uintptr_t)b.cloud_vol, vol_gen = vol_generation(b.cloud_vol);
        const int32_t lvl = c.skip_level, spans = c.occ_wide != nullptr;
        std::memcpy(key + 104, &vol_id, 8);
        std::memcpy(key + 112, &vol_gen, 8);
        std::memcpy(key + 120, &lvl, 4);
        std::memcpy(key + 124, &spans, 4);
        if (!pt->geom_valid || std::memcmp(key, pt->geom_key, sizeof(key)) != 0) {
            CRV_CUDA(cudaMemsetAsync(pt->nhit, 0, 2 * sizeof(uint32_t), s));
            CRV_CUDA(cudaMemsetAsync(pt->job_counter, 0, sizeof(unsigned int), s));
            if (c.occ_wide)
                CRV_TRY(crv_launch_occ_aabb(pt->ctx, b.cloud_vol->occ_ext, pt->occ_aabb, c.cb0, c.cext, c.cloud.nx, c.cloud.ny, c.cloud.nz, s));
            CRV_TRY(crv_launch_classify(pt->ctx, c, pt->hitlist, pt->start, pt->nhit, pt->misslist, pt->missdy, pt->hit_keys, s));
            pt->launches++;
            pt->sorted_valid = false;
            std::memcpy(pt->geom_key, key, sizeof(key));
            pt->geom_valid = true;
            // the counts travel to the host behind the classification; a later render call picks them up if they have arrived
            CRV_CUDA(cudaMemcpyAsync(pt->counts_host, pt->nhit, 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
            CRV_CUDA(cudaEventRecord(pt->counts_ev, s));
            pt->counts_pending = true;
            pt->nhit_valid = false;
        } else if (pt->counts_pending && cudaEventQuery(pt->counts_ev) == cudaSuccess) {
            pt->nhit_known = pt->counts_host[0];
            pt->nhit_valid = true;
            pt->counts_pending = false;
        }
        (void)cudaGetLastError();          // cudaErrorNotReady from the query is not an error
    }

    // rounds of `per_round` frames through the sample buffer (sized for the worst case: every pixel hits)
    const size_t frame_bytes = (size_t)(pt->nhit_valid ? (pt->nhit_known > 0 ? pt->nhit_known : 1u) : c.npix) * sizeof(float2);
    uint64_t per_round = pt->sample_budget / frame_bytes;
    if (per_round < 1) per_round = 1;
    // job ids are 32-bit, and every warp may bump the job counter once more after the queue ran dry (<= 32 per warp, a few
    // thousand warps): keep a margin so that the counter cannot wrap back into the valid range
    const uint64_t max_jobs_frames = (0xffffffffull - (1ull << 22)) / c.npix;
    CRV_REQUIRE(max_jobs_frames >= 1, "RenderTechniquePT: more than 2^32 - 2^22 pixels per frame is not supported");
    if (per_round > max_jobs_frames) per_round = max_jobs_frames;
    if (per_round > (uint64_t)PT_MAX_FRAMES_PER_LAUNCH) per_round = PT_MAX_FRAMES_PER_LAUNCH;   // seeds travel by value
    if (per_round > n) per_round = n;
    const size_t need = (size_t)per_round * frame_bytes;
    if (pt->samples_bytes < need) {
        if (pt->samples) CRV_CUDA(cudaFree(pt->samples));
        pt->samples = nullptr;
        pt->samples_bytes = 0;
        CRV_CUDA(cudaMalloc(&pt->samples, need));
        pt->samples_bytes = need;
    }
    c.samples = pt->samples;
    // Job order: with several frames per launch the jobs run pixel-major over the hit list sorted by descending expected
    // work, so the paths still in flight when the job queue runs dry are the short ones (the tail of the persistent kernel
    // is what a member of a sample split with few frames per GPU pays for).  Identical samples either way.
    const bool sorted_jobs = pt->job_order == 1 || (pt->job_order < 0 && per_round >= 4);
    if (sorted_jobs && !pt->sorted_valid) {
        const float diag = std::sqrt((c.cext[0] * c.cext[0] + c.cext[1] * c.cext[1]) + c.cext[2] * c.cext[2]);
        CRV_TRY(crv_launch_sort_hits(pt->ctx, pt->nhit, pt->hit_keys, c.sigma * 3.0f * diag, pt->hitlist, pt->start, pt->hitlist_sorted,
                                     pt->start_sorted, pt->sorted_from, pt->sort_scratch, s));
        pt->launches += 3;
        pt->sorted_valid = true;
    }
    c.job_order = sorted_jobs ? 1 : 0;
    c.hitlist_sorted = pt->hitlist_sorted;
    c.start_sorted = pt->start_sorted;
    c.sorted_from = pt->sorted_from;
    for (unsigned f0 = 0; f0 < n; f0 += (unsigned)per_round) {
        const unsigned nf = (n - f0 < per_round) ? (n - f0) : (unsigned)per_round;
        c.nframes = nf;
        const U32Div dv = make_u32_div(nf);
        c.div_magic = dv.magic;
        c.div_shift = dv.shift;
        std::memcpy(c.seeds, seeds + f0, sizeof(int32_t) * nf);
        c.first_frame_count = first_fc + f0;
        CRV_TRY(pt_timed_launch(pt, c, b, s));        // the job counter is zero: reset by the previous fold (or by the classification)
        FoldConst fc{};
        fc.image = c.image;
        fc.samples = pt->samples;
        fc.hitlist = pt->hitlist;
        fc.nhit = pt->nhit;
        fc.npix = c.npix;
        fc.width = (uint32_t)c.width;
        fc.nframes = nf;
        fc.first_frame_count = first_fc + f0;
        fc.accum_mode = mode;
        fc.rcp_fc1 = c.rcp_fc1 ? pt->rcp_fc1 + f0 : nullptr;
        fc.misslist = pt->misslist;
        fc.missdy = pt->missdy;
        fc.job_counter = pt->job_counter;
        fc.stats = pt->stats;
        if (push) {
            fc.fresh = f0 == 0;
            fc.last_round = f0 + nf >= n;
            fc.frames_total = n;
            fc.others_scale = others_scale;
            fc.push_n = push->nranks;
            fc.push_rank = push->rank;
            fc.push_epoch = push->epoch;
            for (uint32_t p = 0; p < push->nranks; ++p) {
                fc.push_inbox[p] = reinterpret_cast<float4*>(push->inbox[p]);
                fc.push_flag[p] = push->flag[p];
            }
            fc.push_done = push->done;
            CRV_TRY(crv_launch_fold_push(pt->ctx, fc, s));
        } else {
            CRV_TRY(crv_launch_fold(pt->ctx, fc, s));
        }
        pt->launches += 2;
    }
    return CRV_OK;
}

int crv_pt_record_draw_commands(crv_pt* pt, void* stream, unsigned image_index)
{
    CRV_REQUIRE(pt != nullptr, "crv_pt_record_draw_commands: null technique");
    CRV_ENTER(pt->ctx);
    const int32_t seed = pt->pc->seed;                 // vkCmdPushConstants snapshot (.cpp:132)
    return pt_render(pt, &seed, 1, pt->pc->frameCount, CRV_ACCUM_RUNNING_MEAN, as_stream(stream), image_index);
}

int crv_pt_render_batch(crv_pt* pt, const int32_t* seeds, unsigned n, uint32_t first_frame_count, int accum_mode,
                        void* stream, unsigned image_index)
{
    CRV_REQUIRE(pt != nullptr, "crv_pt_render_batch: null technique");
    CRV_ENTER(pt->ctx);
    return pt_render(pt, seeds, n, first_frame_count, accum_mode, as_stream(stream), image_index);
}

int crv_pt_render_batch_push(crv_pt* pt, const int32_t* seeds, unsigned n, const crv_push_targets* targets, float others_scale,
                             void* stream, unsigned image_index)
{
    CRV_REQUIRE(pt != nullptr, "crv_pt_render_batch_push: null technique");
    CRV_REQUIRE(targets && targets->nranks >= 1 && targets->nranks <= 16 && targets->rank < targets->nranks && targets->done,
                "crv_pt_render_batch_push: incomplete push targets (crv_group_push_begin fills them)");
    for (uint32_t p = 0; p < targets->nranks; ++p)
        CRV_REQUIRE(targets->inbox[p] && targets->flag[p], "crv_pt_render_batch_push: incomplete push targets (crv_group_push_begin fills them)");
    CRV_ENTER(pt->ctx);
    return pt_render(pt, seeds, n, 0, CRV_ACCUM_SUM, as_stream(stream), image_index, targets, others_scale);
}

int crv_pt_get_stats(crv_pt* pt, crv_pt_stats* out, void* stream)
{
    CRV_REQUIRE(pt && out, "crv_pt_get_stats: null argument");
    CRV_ENTER(pt->ctx);
    unsigned long long h[PT_STAT_COUNT];
    CRV_CUDA(cudaMemcpyAsync(h, pt->stats, sizeof(h), cudaMemcpyDeviceToHost, as_stream(stream)));
    CRV_CUDA(cudaStreamSynchronize(as_stream(stream)));
    out->paths = h[PT_STAT_PATHS];
    out->tracking_steps = h[PT_STAT_STEPS];
    out->collisions = h[PT_STAT_COLLISIONS];
    out->density_fetches = h[PT_STAT_FETCHES];
    out->kernel_launches = pt->launches;
    out->tracked_paths = h[PT_STAT_TRACKED];
    return CRV_OK;
}

int crv_pt_pixel_lists(crv_pt* pt, crv_pixel_lists* out)
{
    CRV_REQUIRE(pt && out, "crv_pt_pixel_lists: null argument");
    CRV_REQUIRE(pt->geom_valid && pt->hitlist && pt->misslist, "crv_pt_pixel_lists: nothing rendered yet");
    out->tracked_xy = pt->hitlist;
    out->other_pixels = pt->misslist;
    out->counts = pt->nhit;
    out->width = (uint32_t)(pt->ctor_cam->halfWidth * 2);
    out->reserved = 0;
    return CRV_OK;
}

int crv_pt_last_render_ms(crv_pt* pt, float* tracker_ms, unsigned* tracker_launches)
{
    CRV_REQUIRE(pt && tracker_ms, "crv_pt_last_render_ms: null argument");
    CRV_ENTER(pt->ctx);
    CRV_REQUIRE(pt->ev_used >= 2, "crv_pt_last_render_ms: nothing rendered yet");
    CRV_CUDA(cudaEventSynchronize(pt->ev[pt->ev_used - 1]));
    float total = 0.0f;
    for (size_t i = 0; i + 1 < pt->ev_used; i += 2) {
        float ms = 0.0f;
        CRV_CUDA(cudaEventElapsedTime(&ms, pt->ev[i], pt->ev[i + 1]));
        total += ms;
    }
    *tracker_ms = total;
    if (tracker_launches) *tracker_launches = (unsigned)(pt->ev_used / 2);
    return CRV_OK;
}

int crv_pt_reset_stats(crv_pt* pt, void* stream)
{
    CRV_REQUIRE(pt != nullptr, "crv_pt_reset_stats: null technique");
    CRV_ENTER(pt->ctx);
    CRV_CUDA(cudaMemsetAsync(pt->stats, 0, sizeof(unsigned long long) * PT_STAT_COUNT, as_stream(stream)));
    return CRV_OK;
}

int crv_pt_set_debug_counters(crv_pt* pt, int enable)
{
    CRV_REQUIRE(pt != nullptr, "crv_pt_set_debug_counters: null technique");
    pt->dbg_enabled = enable != 0;
    return CRV_OK;
}

int crv_pt_download_debug_counters(crv_pt* pt, uint32_t* steps, uint32_t* collisions, void* stream)
{
    CRV_REQUIRE(pt != nullptr, "crv_pt_download_debug_counters: null technique");
    CRV_ENTER(pt->ctx);
    CRV_REQUIRE(pt->dbg_steps && pt->dbg_coll, "crv_pt_download_debug_counters: no single-frame render with counters enabled yet");
    cudaStream_t s = as_stream(stream);
    if (steps) CRV_CUDA(cudaMemcpyAsync(steps, pt->dbg_steps, sizeof(uint32_t) * pt->dbg_elems, cudaMemcpyDeviceToHost, s));
    if (collisions) CRV_CUDA(cudaMemcpyAsync(collisions, pt->dbg_coll, sizeof(uint32_t) * pt->dbg_elems, cudaMemcpyDeviceToHost, s));
    CRV_CUDA(cudaStreamSynchronize(s));
    return CRV_OK;
}

int crv_pt_selftest_log(crv_ctx* ctx, uint64_t* mismatches, void* stream)
{
    CRV_ENTER(ctx);
    CRV_REQUIRE(mismatches != nullptr, "crv_pt_selftest_log: null argument");
    cudaStream_t s = as_stream(stream);
    unsigned long long* d = nullptr;
    CRV_CUDA(cudaMalloc(&d, sizeof(unsigned long long)));
    CRV_CUDA(cudaMemsetAsync(d, 0, sizeof(unsigned long long), s));
    int rc = crv_launch_log_selftest(ctx, d, s);
    unsigned long long h = 0;
    if (rc == CRV_OK && cudaMemcpyAsync(&h, d, sizeof(h), cudaMemcpyDeviceToHost, s) != cudaSuccess) rc = CRV_ERR_CUDA;
    cudaStreamSynchronize(s);
    cudaFree(d);
    *mismatches = h;
    return rc;
}

int crv_pt_rng_tap(crv_ctx* ctx, int32_t seed, uint32_t width, uint32_t height, uint32_t n_draws, uint32_t* host_out, void* stream)
{
    CRV_ENTER(ctx);
    CRV_REQUIRE(host_out && width > 0 && height > 0 && n_draws > 0, "crv_pt_rng_tap: bad argument");
    cudaStream_t s = as_stream(stream);
    uint32_t* d = nullptr;
    const size_t bytes = (size_t)width * height * n_draws * sizeof(uint32_t);
    CRV_CUDA(cudaMalloc(&d, bytes));
    int rc = crv_launch_rng_tap(ctx, seed, width, height, n_draws, d, s);
    if (rc == CRV_OK) {
        const cudaError_t e = cudaMemcpyAsync(host_out, d, bytes, cudaMemcpyDeviceToHost, s);
        if (e != cudaSuccess) { crv_set_error("crv_pt_rng_tap: %s", cudaGetErrorString(e)); rc = CRV_ERR_CUDA; }
    }
    cudaStreamSynchronize(s);
    cudaFree(d);
    return rc;
}

}  // extern "C"
