// group.cu -- the multi-GPU half of the path: the spp-split's accumulation buffers are combined and divided by
// the sample count in ONE kernel per GPU that reads and writes its peers' memory over NVLink.
//
// The reference is a single-GPU program; north_star partitions its path by samples: every GPU holds a replica of the density
// grid and the shadow volume, renders its share of the frame list into a private fp32 SUM image (CRV_ACCUM_SUM) and the
// images are combined.  (pixel, frame) samples are independent (PathTracer.comp:368: the seed depends on the frame seed and
// the pixel only) and the running mean (:420-424) is the only coupling, so the combination is sum / spp_total.
//
// crv_group: one member per GPU (one per process under torchrun, or all in one process for crv_headless --gpus N).  Every
// member owns one cudaMalloc'd block -- a flag page followed by `nbuf` RGBA32F images, the render targets -- and maps the
// blocks of its peers (cudaIpc between processes, peer access inside one).  crv_group_all_reduce_sum_scale enqueues, on every
// member's stream:
//   start barrier  every member tells every peer "my image is complete" (a one-warp kernel: a 32-bit word with a system-scope
//                  release into each peer's flag page) and waits for the same word from everybody.  The waits are stream memory
//                  operations (cuStreamWaitValue32): no SM is held while a member waits, and the members' kernels need not be
//                  co-resident (several members may even share one device, as the tests do);
//   band reduce    ONE kernel: member r owns rows band r (1/N of the image): it loads that band from all N images (peer loads,
//                  issued together), adds them in rank order, multiplies by `scale` and stores the result into all N images --
//                  a reduce-scatter, the scale epilogue and an all-gather in one pass, 2 x 33 MB over NVLink at 1080p;
//   end barrier    the kernel's last CTA tells every peer "my band is in your image"; the stream then waits for the peers' words:
//                  past this point the member's image holds the final picture and no peer touches it before the next call.
// The result is deterministic (fixed summation order, computed once per pixel) and identical on all members.
// Where the driver lacks stream memory operations the flags are written / polled by one-warp kernels; their spins carry a
// deadline (default 10 s, crv_group_set_timeout_ms) and a member whose peer never arrives sets a sticky error that
// crv_group_status reports, instead of hanging the GPU.
//
// The NCCL route north_star names is kept beside it for comparison (crv_group_nccl_*: ncclAllReduce in place through the
// libnccl.so.2 the process has loaded, then the scale kernel); bench.py times both.
#include <dlfcn.h>
#include <unistd.h>

#include <chrono>
#include <cstring>
#include <new>
#include <thread>
#include <vector>

#include "crv_internal.h"

namespace {

constexpr int GROUP_MAX = 16;
constexpr size_t FLAG_PAGE = 4096;                 // start[16], end[16], done counter, error word
constexpr int RED_THREADS = 512;

constexpr int GROUP_MAX_BUF = 8;
struct GroupFlags {
    uint32_t start[GROUP_MAX];
    uint32_t end[GROUP_MAX];
    uint32_t done;                                 // CTAs of the local kernel that have stored their part of the band
    uint32_t error;                                // sticky: 1 = start barrier timed out, 2 = end barrier timed out, 3 = staging wait
    uint32_t pad[30];
    // staging all-gather: arrived[b][p] = epoch of the last slice member p has pushed into MY staging buffer b;
    // released[b][p] = last epoch of ITS staging buffer b that member p has finished reading
    uint32_t arrived[GROUP_MAX_BUF][GROUP_MAX];
    uint32_t released[GROUP_MAX_BUF][GROUP_MAX];
    // push combine: pushed[b][p] = epoch of the last picture of buffer b whose sums member p has stored into MY inboxes;
    // push_done[b] = CTA counter of my own fold of buffer b (zero between launches)
    uint32_t pushed[GROUP_MAX_BUF][GROUP_MAX];
    uint32_t push_done[GROUP_MAX_BUF];
};
static_assert(sizeof(GroupFlags) <= FLAG_PAGE, "flag page");

struct ReduceArgs {
    unsigned char* peer[GROUP_MAX];                // base of every member's block (own included), in THIS device's address space
    size_t image_offset;                           // byte offset of the image inside a block
    uint32_t n4;                                   // float4 elements per image
    uint32_t rank, nranks;
    uint32_t epoch;
    float scale;
    unsigned long long timeout_ns;
    // pixel-list variant (crv_group_all_reduce_lists_sum_scale): only the listed pixels travel
    const uint32_t* tracked_xy;                    // x | y << 16, counts[0] entries (this member's own copy of the list)
    const uint32_t* other_pixels;                  // pixel indices, counts[1] entries
    const uint32_t* counts;
    uint32_t width;
    float others_scale;
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v)
{
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p)
{
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// wait until *flag has reached `epoch` (wrap-safe); false on timeout
__device__ __forceinline__ bool wait_flag(const uint32_t* flag, uint32_t epoch, unsigned long long timeout_ns)
{
    const unsigned long long t0 = global_ns();
    unsigned spins = 0;
    while ((int32_t)(ld_acquire_sys(flag) - epoch) < 0) {
        if ((++spins & 1023u) == 0u && global_ns() - t0 > timeout_ns) return false;
        __nanosleep(64);
    }
    return true;
}

// band reduce of member `rank`: its rows band of all N images -> sum in rank order -> * scale -> into all N images.
// The barriers around it are stream operations (flags_write / flags_wait below), so the kernel never waits for anybody and
// can use the whole GPU: all peers' loads of several elements are in flight before the first add.
template <int N>
__global__ void __launch_bounds__(RED_THREADS) group_reduce_kernel(const __grid_constant__ ReduceArgs a)
{
    const int nr = N > 0 ? N : (int)a.nranks;
    const uint32_t per = (a.n4 + a.nranks - 1u) / a.nranks;
    const uint32_t lo = min(a.rank * per, a.n4), hi = min(lo + per, a.n4);
    constexpr int U = N > 0 ? (N <= 2 ? 4 : N <= 4 ? 2 : 1) : 1;      // elements per thread and iteration: U * N loads in flight
    const uint32_t stride = gridDim.x * RED_THREADS;
    for (uint32_t i0 = lo + blockIdx.x * RED_THREADS + threadIdx.x; i0 < hi; i0 += stride * U) {
        float4 v[U][N > 0 ? N : GROUP_MAX];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint32_t i = i0 + u * stride;
#pragma unroll
            for (int p = 0; p < (N > 0 ? N : GROUP_MAX); ++p)
                if (p < nr && i < hi) v[u][p] = __ldcv(reinterpret_cast<const float4*>(a.peer[p] + a.image_offset) + i);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint32_t i = i0 + u * stride;
            if (i >= hi) continue;
            float4 s = v[u][0];
#pragma unroll
            for (int p = 1; p < (N > 0 ? N : GROUP_MAX); ++p)
                if (p < nr) { s.x += v[u][p].x; s.y += v[u][p].y; s.z += v[u][p].z; s.w += v[u][p].w; }
            s.x *= a.scale; s.y *= a.scale; s.z *= a.scale; s.w *= a.scale;
#pragma unroll
            for (int p = 0; p < (N > 0 ? N : GROUP_MAX); ++p)
                if (p < nr) reinterpret_cast<float4*>(a.peer[p] + a.image_offset)[i] = s;
        }
    }
    // end barrier, signalling half: the last CTA to finish tells every peer "my band is in your image"
    __shared__ int s_last;
    GroupFlags* mine = reinterpret_cast<GroupFlags*>(a.peer[a.rank]);
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(&mine->done, 1u) == gridDim.x - 1u);
    __syncthreads();
    if (!s_last) return;
    if (threadIdx.x == 0) mine->done = 0u;         // re-armed for the next call
    if ((int)threadIdx.x < nr && threadIdx.x != a.rank) {
        __threadfence_system();
        st_release_sys(&reinterpret_cast<GroupFlags*>(a.peer[threadIdx.x])->end[a.rank], a.epoch);
    }
}

// The same with pixel lists: every member holds the same SET of listed pixels (its tracker's hit list; the order of the list
// may differ from member to member, so ownership goes by pixel index: member r combines the listed pixels with index % N == r)
// and only those cross NVLink; every other pixel is finished locally (value * others_scale).
template <int N>
__global__ void __launch_bounds__(RED_THREADS) group_reduce_lists_kernel(const __grid_constant__ ReduceArgs a)
{
    const int nr = N > 0 ? N : (int)a.nranks;
    const uint32_t ntracked = a.counts[0], nothers = a.counts[1];
    const uint32_t stride = gridDim.x * RED_THREADS, t0 = blockIdx.x * RED_THREADS + threadIdx.x;
    for (uint32_t j = t0; j < ntracked; j += stride) {
        const uint32_t xy = __ldg(a.tracked_xy + j);
        const uint32_t i = (xy >> 16) * a.width + (xy & 0xffffu);
        if (i % a.nranks != a.rank) continue;
        float4 v[N > 0 ? N : GROUP_MAX];
#pragma unroll
        for (int p = 0; p < (N > 0 ? N : GROUP_MAX); ++p)
            if (p < nr) v[p] = __ldcv(reinterpret_cast<const float4*>(a.peer[p] + a.image_offset) + i);
        float4 s = v[0];
#pragma unroll
        for (int p = 1; p < (N > 0 ? N : GROUP_MAX); ++p)
            if (p < nr) { s.x += v[p].x; s.y += v[p].y; s.z += v[p].z; s.w += v[p].w; }
        s.x *= a.scale; s.y *= a.scale; s.z *= a.scale; s.w *= a.scale;
#pragma unroll
        for (int p = 0; p < (N > 0 ? N : GROUP_MAX); ++p)
            if (p < nr) reinterpret_cast<float4*>(a.peer[p] + a.image_offset)[i] = s;
    }
    float4* img = reinterpret_cast<float4*>(a.peer[a.rank] + a.image_offset);
    for (uint32_t j = t0; j < nothers; j += stride) {
        const uint32_t i = __ldg(a.other_pixels + j);
        float4 v = img[i];
        v.x *= a.others_scale; v.y *= a.others_scale; v.z *= a.others_scale; v.w *= a.others_scale;
        img[i] = v;
    }
    __shared__ int s_last;
    GroupFlags* mine = reinterpret_cast<GroupFlags*>(a.peer[a.rank]);
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(&mine->done, 1u) == gridDim.x - 1u);
    __syncthreads();
    if (!s_last) return;
    if (threadIdx.x == 0) mine->done = 0u;
    if ((int)threadIdx.x < nr && threadIdx.x != a.rank) {
        __threadfence_system();
        st_release_sys(&reinterpret_cast<GroupFlags*>(a.peer[threadIdx.x])->end[a.rank], a.epoch);
    }
}

__global__ void group_scale_lists_kernel(float4* __restrict__ img, const uint32_t* __restrict__ tracked_xy, const uint32_t* __restrict__ other_pixels,
                                         const uint32_t* __restrict__ counts, uint32_t width, float scale, float others_scale)
{
    const uint32_t stride = gridDim.x * blockDim.x, t0 = blockIdx.x * blockDim.x + threadIdx.x;
    for (uint32_t j = t0; j < counts[0]; j += stride) {
        const uint32_t xy = tracked_xy[j];
        float4& v = img[(xy >> 16) * width + (xy & 0xffffu)];
        v.x *= scale; v.y *= scale; v.z *= scale; v.w *= scale;
    }
    for (uint32_t j = t0; j < counts[1]; j += stride) {
        float4& v = img[other_pixels[j]];
        v.x *= others_scale; v.y *= others_scale; v.z *= others_scale; v.w *= others_scale;
    }
}

// Second half of the push combine (crv_group_push_finish): every member's fold has stored its tracked sums into the inbox
// this member keeps for it and the flags have been waited for, so everything is local: tracked pixel = (inbox[0] + inbox[1]
// + ...) * scale in rank order -- the arithmetic of group_reduce_lists_kernel, so the same bits, on every member.
struct FinishArgs {
    const float4* inbox0;                          // inbox of member 0; member p's follows at p * stride4
    size_t stride4;
    float4* image;
    const uint32_t* tracked_xy;
    const uint32_t* counts;
    uint32_t width, nranks;
    float scale;
};
template <int N>
__global__ void __launch_bounds__(RED_THREADS) group_push_finish_kernel(const __grid_constant__ FinishArgs a)
{
    const int nr = N > 0 ? N : (int)a.nranks;
    const uint32_t ntracked = a.counts[0];
    const uint32_t stride = gridDim.x * RED_THREADS;
    for (uint32_t j = blockIdx.x * RED_THREADS + threadIdx.x; j < ntracked; j += stride) {
        const uint32_t xy = __ldg(a.tracked_xy + j);
        const uint32_t i = (xy >> 16) * a.width + (xy & 0xffffu);
        float4 v[N > 0 ? N : GROUP_MAX];
#pragma unroll
        for (int p = 0; p < (N > 0 ? N : GROUP_MAX); ++p)
            if (p < nr) v[p] = __ldcv(a.inbox0 + p * a.stride4 + i);
        float4 s = v[0];
#pragma unroll
        for (int p = 1; p < (N > 0 ? N : GROUP_MAX); ++p)
            if (p < nr) { s.x += v[p].x; s.y += v[p].y; s.z += v[p].z; s.w += v[p].w; }
        s.x *= a.scale; s.y *= a.scale; s.z *= a.scale; s.w *= a.scale;
        a.image[i] = s;
    }
}

// fallbacks for cuStreamWriteValue32 / cuStreamWaitValue32 where the driver does not offer them
struct FlagList {
    uint32_t* addr[GROUP_MAX];
    uint32_t n;
};
__global__ void flag_write_kernel(const __grid_constant__ FlagList f, uint32_t value)
{
    if (threadIdx.x < f.n) {
        __threadfence_system();
        st_release_sys(f.addr[threadIdx.x], value);
    }
}
__global__ void flag_wait_kernel(const __grid_constant__ FlagList f, uint32_t value, unsigned long long timeout_ns, uint32_t* error)
{
    if (threadIdx.x < f.n && !wait_flag(f.addr[threadIdx.x], value, timeout_ns)) atomicExch(error, 3u);
}

__global__ void group_scale_kernel(float4* __restrict__ img, uint32_t n4, float scale)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += gridDim.x * blockDim.x) {
        float4 v = img[i];
        v.x *= scale; v.y *= scale; v.z *= scale; v.w *= scale;
        img[i] = v;
    }
}

// ---- NCCL through the library the process has loaded (no link-time dependency) ------------------------------------------
struct Id128 { char b[128]; };                     // ncclUniqueId, passed by value
struct NcclApi {
    void* lib = nullptr;
    int (*GetUniqueId)(void*) = nullptr;
    int (*CommInitRank)(void**, int, Id128, int) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};
NcclApi g_nccl;

bool nccl_load()
{
    if (g_nccl.lib) return true;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return false;
    g_nccl.GetUniqueId = reinterpret_cast<int (*)(void*)>(dlsym(h, "ncclGetUniqueId"));
    g_nccl.CommInitRank = reinterpret_cast<int (*)(void**, int, Id128, int)>(dlsym(h, "ncclCommInitRank"));
    g_nccl.AllReduce = reinterpret_cast<int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t)>(dlsym(h, "ncclAllReduce"));
    g_nccl.CommDestroy = reinterpret_cast<int (*)(void*)>(dlsym(h, "ncclCommDestroy"));
    g_nccl.GetErrorString = reinterpret_cast<const char* (*)(int)>(dlsym(h, "ncclGetErrorString"));
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllReduce || !g_nccl.CommDestroy) {
        dlclose(h);
        return false;
    }
    g_nccl.lib = h;
    return true;
}

}  // namespace

// cuStreamWriteValue32 / cuStreamWaitValue32 through the runtime's driver entry points (no link-time libcuda dependency)
typedef int (*StreamValue32Fn)(cudaStream_t, unsigned long long, uint32_t, unsigned int);
static StreamValue32Fn g_stream_write32 = nullptr, g_stream_wait32 = nullptr;
static bool g_memops_probed = false;
static void probe_stream_memops()
{
    if (g_memops_probed) return;
    g_memops_probed = true;
    void *w = nullptr, *q = nullptr;
    cudaDriverEntryPointQueryResult rw, rq;
    if (cudaGetDriverEntryPoint("cuStreamWriteValue32", &w, cudaEnableDefault, &rw) == cudaSuccess && rw == cudaDriverEntryPointSuccess &&
        cudaGetDriverEntryPoint("cuStreamWaitValue32", &q, cudaEnableDefault, &rq) == cudaSuccess && rq == cudaDriverEntryPointSuccess && w && q) {
        g_stream_write32 = reinterpret_cast<StreamValue32Fn>(w);
        g_stream_wait32 = reinterpret_cast<StreamValue32Fn>(q);
    }
    (void)cudaGetLastError();
}

struct crv_group {
    crv_ctx* ctx = nullptr;
    uint32_t rank = 0, nranks = 1, nbuf = 1;
    uint32_t width = 0, height = 0;
    size_t image_bytes = 0, block_bytes = 0;
    size_t staging_bytes = 0, staging_offset = 0;  // per-buffer size (256-aligned) and offset of staging buffer 0 in the block
    uint32_t stage_epoch[GROUP_MAX_BUF] = {};
    size_t inbox_offset = 0;                       // inbox of (buffer b, slot s, source member r) at inbox_offset + ((b * 2 + s) * nranks + r) * image_bytes
    uint32_t push_epoch[GROUP_MAX_BUF] = {};
    bool use_memops = false;
    unsigned char* block = nullptr;                // own allocation
    unsigned char* peer[GROUP_MAX] = {};           // mapped peers (own block at [rank])
    bool peer_ipc[GROUP_MAX] = {};                 // opened with cudaIpcOpenMemHandle (to be closed)
    bool connected = false;
    std::vector<crv_image*> images;
    uint32_t epoch = 0;
    unsigned long long timeout_ns = 10ull * 1000ull * 1000ull * 1000ull;
    void* nccl_comm = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    bool timed = false;
    cudaStream_t aux = nullptr;                    // crv_group_status uses it to release a wait whose peer never arrived
    bool broken = false;                           // a barrier timed out: the group's flags no longer mean anything
};

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }
#define CRV_TRY_G(expr)                \
    do {                               \
        const int _rc = (expr);        \
        if (_rc != CRV_OK) return _rc; \
    } while (0)

extern "C" {

static int flags_write(crv_group* g, uint32_t* const* addrs, uint32_t n, uint32_t value, cudaStream_t s, bool prefer_kernel = false);
static int flags_wait(crv_group* g, uint32_t* const* addrs, uint32_t n, uint32_t value, cudaStream_t s);

int crv_group_create(crv_ctx* ctx, uint32_t rank, uint32_t nranks, uint32_t width, uint32_t height, uint32_t nbuf, uint64_t staging_bytes,
                     crv_group** out)
{
    CRV_REQUIRE(ctx && out, "crv_group_create: null argument");
    *out = nullptr;
    CRV_REQUIRE(nranks >= 1 && nranks <= (uint32_t)GROUP_MAX && rank < nranks, "crv_group_create: rank / nranks out of range (at most 16 members)");
    CRV_REQUIRE(width > 0 && height > 0 && nbuf >= 1 && nbuf <= 8, "crv_group_create: bad image size or buffer count");
    CRV_CUDA(cudaSetDevice(ctx->device));
    crv_group* g = new (std::nothrow) crv_group();
    if (!g) return CRV_ERR_OOM;
    g->ctx = ctx;
    g->rank = rank;
    g->nranks = nranks;
    g->nbuf = nbuf;
    g->width = width;
    g->height = height;
    g->image_bytes = ((size_t)width * height * 16 + 255) / 256 * 256;
    g->staging_bytes = ((size_t)staging_bytes + 255) / 256 * 256;
    g->staging_offset = FLAG_PAGE + g->image_bytes * nbuf;
    g->inbox_offset = g->staging_offset + g->staging_bytes * nbuf;
    g->block_bytes = g->inbox_offset + g->image_bytes * 2 * nranks * nbuf;
    probe_stream_memops();
    g->use_memops = g_stream_write32 != nullptr;
    {
        // Load the kernels now.  CUDA loads a kernel at its first launch, and that may synchronise the device: behind a stream
        // that is waiting for a peer served later by the same host thread (several members in one process) it would never return.
        cudaFuncAttributes fa;
        (void)cudaFuncGetAttributes(&fa, group_reduce_kernel<0>);
        (void)cudaFuncGetAttributes(&fa, group_reduce_kernel<2>);
        (void)cudaFuncGetAttributes(&fa, group_reduce_kernel<4>);
        (void)cudaFuncGetAttributes(&fa, group_reduce_kernel<8>);
        (void)cudaFuncGetAttributes(&fa, group_scale_kernel);
        (void)cudaFuncGetAttributes(&fa, group_reduce_lists_kernel<0>);
        (void)cudaFuncGetAttributes(&fa, group_reduce_lists_kernel<2>);
        (void)cudaFuncGetAttributes(&fa, group_reduce_lists_kernel<4>);
        (void)cudaFuncGetAttributes(&fa, group_reduce_lists_kernel<8>);
        (void)cudaFuncGetAttributes(&fa, group_scale_lists_kernel);
        (void)cudaFuncGetAttributes(&fa, group_push_finish_kernel<0>);
        (void)cudaFuncGetAttributes(&fa, group_push_finish_kernel<1>);
        (void)cudaFuncGetAttributes(&fa, group_push_finish_kernel<2>);
        (void)cudaFuncGetAttributes(&fa, group_push_finish_kernel<4>);
        (void)cudaFuncGetAttributes(&fa, group_push_finish_kernel<8>);
        (void)cudaFuncGetAttributes(&fa, flag_write_kernel);
        (void)cudaFuncGetAttributes(&fa, flag_wait_kernel);
        (void)cudaGetLastError();
    }
    cudaError_t e = cudaMalloc(&g->block, g->block_bytes);
    if (e != cudaSuccess) {
        crv_set_error("crv_group_create: cudaMalloc of %zu bytes failed: %s", g->block_bytes, cudaGetErrorString(e));
        delete g;
        return CRV_ERR_OOM;
    }
    e = cudaMemset(g->block, 0, g->block_bytes);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&g->aux, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreate(&g->ev0);
    if (e == cudaSuccess) e = cudaEventCreate(&g->ev1);
    if (e != cudaSuccess) {
        crv_set_error("crv_group_create: %s", cudaGetErrorString(e));
        cudaFree(g->block);
        delete g;
        return CRV_ERR_CUDA;
    }
    g->peer[rank] = g->block;
    for (uint32_t k = 0; k < nbuf; ++k) {
        crv_image* img = nullptr;
        const int rc = crv_image_wrap(ctx, reinterpret_cast<float*>(g->block + FLAG_PAGE + k * g->image_bytes), width, height, &img);
        if (rc != CRV_OK) {
            for (crv_image* i : g->images) crv_image_destroy(i);
            cudaFree(g->block);
            delete g;
            return rc;
        }
        g->images.push_back(img);
    }
    g->connected = nranks == 1;
    *out = g;
    return CRV_OK;
}

// Release every wait this member may be sitting in: its own flags are set to the epochs it is waiting for.  Only after a
// timeout (the group is unusable afterwards) or at destruction.
static void group_break_waits(crv_group* g)
{
    GroupFlags h{};
    for (uint32_t p = 0; p < (uint32_t)GROUP_MAX; ++p) {
        h.start[p] = h.end[p] = g->epoch;
        for (uint32_t b = 0; b < (uint32_t)GROUP_MAX_BUF; ++b) {
            h.arrived[b][p] = h.released[b][p] = g->stage_epoch[b];
            h.pushed[b][p] = g->push_epoch[b];
        }
    }
    h.error = 4u;
    cudaMemcpyAsync(g->block, &h, sizeof(h), cudaMemcpyHostToDevice, g->aux);
    cudaStreamSynchronize(g->aux);
    g->broken = true;
}

int crv_group_destroy(crv_group* g)
{
    if (!g) return CRV_OK;
    cudaSetDevice(g->ctx->device);
    if (g->nranks > 1 && g->connected) group_break_waits(g);      // a stream still waiting for a peer must not block the teardown
    cudaDeviceSynchronize();
    if (g->nccl_comm && g_nccl.CommDestroy) g_nccl.CommDestroy(g->nccl_comm);
    for (uint32_t p = 0; p < g->nranks; ++p)
        if (p != g->rank && g->peer[p] && g->peer_ipc[p]) cudaIpcCloseMemHandle(g->peer[p]);
    for (crv_image* i : g->images) crv_image_destroy(i);
    if (g->ev0) cudaEventDestroy(g->ev0);
    if (g->ev1) cudaEventDestroy(g->ev1);
    if (g->aux) cudaStreamDestroy(g->aux);
    if (g->block) cudaFree(g->block);
    delete g;
    return CRV_OK;
}

int crv_group_export(crv_group* g, crv_group_handle* out)
{
    CRV_REQUIRE(g && out, "crv_group_export: null argument");
    CRV_CUDA(cudaSetDevice(g->ctx->device));
    std::memset(out, 0, sizeof(*out));
    out->pid = (uint64_t)getpid();
    out->device = g->ctx->device;
    out->rank = g->rank;
    out->base = (uint64_t)(uintptr_t)g->block;
    out->bytes = (uint64_t)g->block_bytes;
    static_assert(sizeof(cudaIpcMemHandle_t) <= sizeof(out->ipc), "ipc handle size");
    cudaIpcMemHandle_t h;
    CRV_CUDA(cudaIpcGetMemHandle(&h, g->block));
    std::memcpy(out->ipc, &h, sizeof(h));
    return CRV_OK;
}

int crv_group_connect(crv_group* g, const crv_group_handle* handles, uint32_t count)
{
    CRV_REQUIRE(g && handles, "crv_group_connect: null argument");
    CRV_REQUIRE(count == g->nranks, "crv_group_connect: one handle per member, in rank order");
    CRV_CUDA(cudaSetDevice(g->ctx->device));
    for (uint32_t p = 0; p < count; ++p) {
        const crv_group_handle& h = handles[p];
        CRV_REQUIRE(h.rank == p, "crv_group_connect: handles are not in rank order");
        CRV_REQUIRE(h.bytes == (uint64_t)g->block_bytes, "crv_group_connect: a member was created with another image size or buffer count");
        if (p == g->rank) continue;
        if (g->peer[p]) continue;
        if (h.pid == (uint64_t)getpid()) {
            // same process: the pointer is valid as it is once peer access is on (same device: nothing to do)
            if (h.device != g->ctx->device) {
                int can = 0;
                CRV_CUDA(cudaDeviceCanAccessPeer(&can, g->ctx->device, h.device));
                CRV_REQUIRE(can, "crv_group_connect: no peer access between two member devices");
                const cudaError_t e = cudaDeviceEnablePeerAccess(h.device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) CRV_CUDA(e);
                (void)cudaGetLastError();
            }
            g->peer[p] = reinterpret_cast<unsigned char*>((uintptr_t)h.base);
        } else {
            cudaIpcMemHandle_t ih;
            std::memcpy(&ih, h.ipc, sizeof(ih));
            void* ptr = nullptr;
            CRV_CUDA(cudaIpcOpenMemHandle(&ptr, ih, cudaIpcMemLazyEnablePeerAccess));
            g->peer[p] = static_cast<unsigned char*>(ptr);
            g->peer_ipc[p] = true;
        }
    }
    g->connected = true;
    return CRV_OK;
}

int crv_group_image(crv_group* g, uint32_t buffer, crv_image** out)
{
    CRV_REQUIRE(g && out, "crv_group_image: null argument");
    CRV_REQUIRE(buffer < g->nbuf, "crv_group_image: buffer index out of range");
    *out = g->images[buffer];
    return CRV_OK;
}

int crv_group_set_timeout_ms(crv_group* g, uint32_t ms)
{
    CRV_REQUIRE(g && ms > 0, "crv_group_set_timeout_ms: bad argument");
    g->timeout_ns = (unsigned long long)ms * 1000000ull;
    return CRV_OK;
}

static int group_all_reduce(crv_group* g, uint32_t buffer, float scale, const crv_pixel_lists* lists, float others_scale, void* stream)
{
    CRV_REQUIRE(g != nullptr, "crv_group_all_reduce_sum_scale: null group");
    CRV_REQUIRE(buffer < g->nbuf, "crv_group_all_reduce_sum_scale: buffer index out of range");
    CRV_REQUIRE(g->connected, "crv_group_all_reduce_sum_scale: crv_group_connect has not been called");
    CRV_REQUIRE(!g->broken, "crv_group_all_reduce_sum_scale: a barrier of this group timed out earlier");
    CRV_CUDA(cudaSetDevice(g->ctx->device));
    cudaStream_t s = as_stream(stream);
    const uint32_t n4 = g->width * g->height;
    CRV_CUDA(cudaEventRecord(g->ev0, s));
    if (g->nranks == 1) {
        float4* img = reinterpret_cast<float4*>(g->block + FLAG_PAGE + buffer * g->image_bytes);
        if (lists) group_scale_lists_kernel<<<g->ctx->sm_count * 2, 512, 0, s>>>(img, lists->tracked_xy, lists->other_pixels, lists->counts, lists->width, scale, others_scale);
        else group_scale_kernel<<<g->ctx->sm_count * 2, 512, 0, s>>>(img, n4, scale);
        g->ctx->launches++;
    } else {
        const uint32_t epoch = ++g->epoch;
        GroupFlags* mine = reinterpret_cast<GroupFlags*>(g->block);
        uint32_t* addrs[GROUP_MAX];
        uint32_t n = 0;
        // start barrier: "my image is complete" to everybody (behind the render in stream order, with a system-scope fence),
        // then wait for the same word from everybody -- stream operations, no SM is held while waiting
        for (uint32_t p = 0; p < g->nranks; ++p) if (p != g->rank) addrs[n++] = &reinterpret_cast<GroupFlags*>(g->peer[p])->start[g->rank];
        CRV_TRY_G(flags_write(g, addrs, n, epoch, s, true));
        n = 0;
        for (uint32_t p = 0; p < g->nranks; ++p) if (p != g->rank) addrs[n++] = &mine->start[p];
        CRV_TRY_G(flags_wait(g, addrs, n, epoch, s));
        ReduceArgs a{};
        for (uint32_t p = 0; p < g->nranks; ++p) a.peer[p] = g->peer[p];
        a.image_offset = FLAG_PAGE + buffer * g->image_bytes;
        a.n4 = n4;
        a.rank = g->rank;
        a.nranks = g->nranks;
        a.epoch = epoch;
        a.scale = scale;
        a.timeout_ns = g->timeout_ns;
        if (lists) {
            a.tracked_xy = lists->tracked_xy;
            a.other_pixels = lists->other_pixels;
            a.counts = lists->counts;
            a.width = lists->width;
            a.others_scale = others_scale;
        }
        const uint32_t band = lists ? n4 : (n4 + g->nranks - 1) / g->nranks;
        unsigned blocks = (band + RED_THREADS - 1) / RED_THREADS;
        const unsigned cap = (unsigned)g->ctx->sm_count * 2u;
        if (blocks > cap) blocks = cap;
        if (blocks < 1u) blocks = 1u;
        if (lists) {
            switch (g->nranks) {
                case 2: group_reduce_lists_kernel<2><<<blocks, RED_THREADS, 0, s>>>(a); break;
                case 4: group_reduce_lists_kernel<4><<<blocks, RED_THREADS, 0, s>>>(a); break;
                case 8: group_reduce_lists_kernel<8><<<blocks, RED_THREADS, 0, s>>>(a); break;
                default: group_reduce_lists_kernel<0><<<blocks, RED_THREADS, 0, s>>>(a); break;
            }
        } else {
            switch (g->nranks) {
                case 2: group_reduce_kernel<2><<<blocks, RED_THREADS, 0, s>>>(a); break;
                case 4: group_reduce_kernel<4><<<blocks, RED_THREADS, 0, s>>>(a); break;
                case 8: group_reduce_kernel<8><<<blocks, RED_THREADS, 0, s>>>(a); break;
                default: group_reduce_kernel<0><<<blocks, RED_THREADS, 0, s>>>(a); break;
            }
        }
        g->ctx->launches++;
        CRV_CUDA(cudaGetLastError());
        // end barrier: the kernel's last CTA has told everybody "my band is in your image"; wait for theirs: when the stream gets
        // past this point the member's image holds the final picture and no peer touches it before the next call's start barrier
        n = 0;
        for (uint32_t p = 0; p < g->nranks; ++p) if (p != g->rank) addrs[n++] = &mine->end[p];
        CRV_TRY_G(flags_wait(g, addrs, n, epoch, s));
    }
    CRV_CUDA(cudaGetLastError());
    CRV_CUDA(cudaEventRecord(g->ev1, s));
    g->timed = true;
    return CRV_OK;
}

int crv_group_all_reduce_sum_scale(crv_group* g, uint32_t buffer, float scale, void* stream)
{
    return group_all_reduce(g, buffer, scale, nullptr, 0.0f, stream);
}

int crv_group_all_reduce_lists_sum_scale(crv_group* g, uint32_t buffer, float scale, float others_scale, const crv_pixel_lists* lists, void* stream)
{
    CRV_REQUIRE(lists && lists->tracked_xy && lists->other_pixels && lists->counts && lists->width > 0,
                "crv_group_all_reduce_lists_sum_scale: incomplete pixel lists");
    CRV_REQUIRE(g == nullptr || lists->width == g->width, "crv_group_all_reduce_lists_sum_scale: the lists belong to another image width");
    return group_all_reduce(g, buffer, scale, lists, others_scale, stream);
}

int crv_group_push_begin(crv_group* g, uint32_t buffer, crv_push_targets* out)
{
    CRV_REQUIRE(g && out, "crv_group_push_begin: null argument");
    CRV_REQUIRE(buffer < g->nbuf, "crv_group_push_begin: buffer index out of range");
    CRV_REQUIRE(g->connected, "crv_group_push_begin: crv_group_connect has not been called");
    CRV_REQUIRE(!g->broken, "crv_group_push_begin: a barrier of this group timed out earlier");
    std::memset(out, 0, sizeof(*out));
    const uint32_t epoch = ++g->push_epoch[buffer];
    const size_t slot = g->inbox_offset + ((size_t)(buffer * 2u + (epoch & 1u)) * g->nranks + g->rank) * g->image_bytes;
    for (uint32_t p = 0; p < g->nranks; ++p) {
        out->inbox[p] = reinterpret_cast<float*>(g->peer[p] + slot);
        out->flag[p] = &reinterpret_cast<GroupFlags*>(g->peer[p])->pushed[buffer][g->rank];
    }
    out->done = &reinterpret_cast<GroupFlags*>(g->block)->push_done[buffer];
    out->nranks = g->nranks;
    out->rank = g->rank;
    out->epoch = epoch;
    return CRV_OK;
}

int crv_group_push_finish(crv_group* g, uint32_t buffer, float scale, const crv_pixel_lists* lists, void* stream)
{
    CRV_REQUIRE(g != nullptr, "crv_group_push_finish: null group");
    CRV_REQUIRE(buffer < g->nbuf, "crv_group_push_finish: buffer index out of range");
    CRV_REQUIRE(lists && lists->tracked_xy && lists->counts && lists->width == g->width, "crv_group_push_finish: pixel lists missing or of another image width");
    CRV_REQUIRE(g->push_epoch[buffer] > 0, "crv_group_push_finish: crv_group_push_begin has not been called for this buffer");
    CRV_REQUIRE(!g->broken, "crv_group_push_finish: a barrier of this group timed out earlier");
    CRV_CUDA(cudaSetDevice(g->ctx->device));
    cudaStream_t s = as_stream(stream);
    const uint32_t epoch = g->push_epoch[buffer];
    CRV_CUDA(cudaEventRecord(g->ev0, s));
    if (g->nranks > 1) {
        GroupFlags* mine = reinterpret_cast<GroupFlags*>(g->block);
        uint32_t* addrs[GROUP_MAX];
        uint32_t n = 0;
        for (uint32_t p = 0; p < g->nranks; ++p) if (p != g->rank) addrs[n++] = &mine->pushed[buffer][p];
        CRV_TRY_G(flags_wait(g, addrs, n, epoch, s));
    }
    FinishArgs a{};
    a.inbox0 = reinterpret_cast<const float4*>(g->block + g->inbox_offset + (size_t)(buffer * 2u + (epoch & 1u)) * g->nranks * g->image_bytes);
    a.stride4 = g->image_bytes / 16;
    a.image = reinterpret_cast<float4*>(g->block + FLAG_PAGE + buffer * g->image_bytes);
    a.tracked_xy = lists->tracked_xy;
    a.counts = lists->counts;
    a.width = g->width;
    a.nranks = g->nranks;
    a.scale = scale;
    const unsigned blocks = (unsigned)g->ctx->sm_count * 2u;
    switch (g->nranks) {
        case 1: group_push_finish_kernel<1><<<blocks, RED_THREADS, 0, s>>>(a); break;
        case 2: group_push_finish_kernel<2><<<blocks, RED_THREADS, 0, s>>>(a); break;
        case 4: group_push_finish_kernel<4><<<blocks, RED_THREADS, 0, s>>>(a); break;
        case 8: group_push_finish_kernel<8><<<blocks, RED_THREADS, 0, s>>>(a); break;
        default: group_push_finish_kernel<0><<<blocks, RED_THREADS, 0, s>>>(a); break;
    }
    g->ctx->launches++;
    CRV_CUDA(cudaGetLastError());
    CRV_CUDA(cudaEventRecord(g->ev1, s));
    g->timed = true;
    return CRV_OK;
}

int crv_group_status(crv_group* g, void* stream)
{
    CRV_REQUIRE(g != nullptr, "crv_group_status: null group");
    CRV_CUDA(cudaSetDevice(g->ctx->device));
    cudaStream_t s = as_stream(stream);
    // wait for the stream with the group's deadline: a wait for a peer that never arrives is a stream operation that would
    // block a plain synchronise for ever
    const auto t0 = std::chrono::steady_clock::now();
    bool timed_out = false;
    while (true) {
        const cudaError_t q = cudaStreamQuery(s);
        if (q == cudaSuccess) break;
        if (q != cudaErrorNotReady) CRV_CUDA(q);
        (void)cudaGetLastError();
        const auto waited = std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t0).count();
        if ((unsigned long long)waited > g->timeout_ns) { timed_out = true; break; }
        std::this_thread::sleep_for(std::chrono::microseconds(50));
    }
    if (timed_out) {
        group_break_waits(g);
        CRV_CUDA(cudaStreamSynchronize(s));
        crv_set_error("crv_group: member %u gave up waiting for a peer (timeout %llu ms); the group is unusable now", g->rank,
                      g->timeout_ns / 1000000ull);
        return CRV_ERR_CUDA;
    }
    uint32_t err = 0;
    CRV_CUDA(cudaMemcpyAsync(&err, g->block + offsetof(GroupFlags, error), sizeof(err), cudaMemcpyDeviceToHost, s));
    CRV_CUDA(cudaStreamSynchronize(s));
    if (err != 0 || g->broken) {
        crv_set_error("crv_group: member %u gave up waiting for a peer at the %s (timeout %llu ms)", g->rank,
                      err == 1 ? "start barrier" : err == 2 ? "end barrier" : err == 3 ? "staging exchange" : "barrier", g->timeout_ns / 1000000ull);
        return CRV_ERR_CUDA;
    }
    return CRV_OK;
}

int crv_group_last_reduce_ms(crv_group* g, float* ms)
{
    CRV_REQUIRE(g && ms, "crv_group_last_reduce_ms: null argument");
    CRV_REQUIRE(g->timed, "crv_group_last_reduce_ms: nothing reduced yet");
    CRV_CUDA(cudaSetDevice(g->ctx->device));
    CRV_CUDA(cudaEventSynchronize(g->ev1));
    CRV_CUDA(cudaEventElapsedTime(ms, g->ev0, g->ev1));
    return CRV_OK;
}

// ---- staging all-gather: a replicated input (the density grid) crosses PCIe once per job, 1/N on every member's link ----------
static int flags_write(crv_group* g, uint32_t* const* addrs, uint32_t n, uint32_t value, cudaStream_t s, bool prefer_kernel)
{
    // one-warp kernel: all peers' words in parallel, one launch (a few microseconds); stream memory operations: one driver call
    // and a few microseconds per word, but no SM -- the staging exchange, which must run beside a kernel that holds every SM
    if (g->use_memops && !prefer_kernel) {
        bool ok = true;
        for (uint32_t i = 0; i < n && ok; ++i) ok = g_stream_write32(s, (unsigned long long)(uintptr_t)addrs[i], value, 0u) == 0;
        if (ok) return CRV_OK;
        g->use_memops = false;                     // e.g. peer addresses not accepted: the kernels do the same (re-writing a flag is harmless)
    }
    FlagList f{};
    f.n = n;
    for (uint32_t i = 0; i < n; ++i) f.addr[i] = addrs[i];
    flag_write_kernel<<<1, 32, 0, s>>>(f, value);
    g->ctx->launches++;
    CRV_CUDA(cudaGetLastError());
    return CRV_OK;
}

static int flags_wait(crv_group* g, uint32_t* const* addrs, uint32_t n, uint32_t value, cudaStream_t s)
{
    if (g->use_memops) {
        bool ok = true;
        for (uint32_t i = 0; i < n && ok; ++i) ok = g_stream_wait32(s, (unsigned long long)(uintptr_t)addrs[i], value, /* CU_STREAM_WAIT_VALUE_GEQ */ 0u) == 0;
        if (ok) return CRV_OK;
        g->use_memops = false;
    }
    FlagList f{};
    f.n = n;
    for (uint32_t i = 0; i < n; ++i) f.addr[i] = addrs[i];
    flag_wait_kernel<<<1, 32, 0, s>>>(f, value, g->timeout_ns, &reinterpret_cast<GroupFlags*>(g->block)->error);
    g->ctx->launches++;
    CRV_CUDA(cudaGetLastError());
    return CRV_OK;
}

int crv_group_staging_ptr(crv_group* g, uint32_t buffer, void** out)
{
    CRV_REQUIRE(g && out, "crv_group_staging_ptr: null argument");
    CRV_REQUIRE(buffer < g->nbuf && g->staging_bytes > 0, "crv_group_staging_ptr: no such staging buffer");
    *out = g->block + g->staging_offset + buffer * g->staging_bytes;
    return CRV_OK;
}

int crv_group_staging_all_gather(crv_group* g, uint32_t buffer, const void* host_src, uint64_t bytes, void* stream)
{
    CRV_REQUIRE(g && host_src, "crv_group_staging_all_gather: null argument");
    CRV_REQUIRE(buffer < g->nbuf && bytes > 0 && bytes <= g->staging_bytes, "crv_group_staging_all_gather: buffer index or size out of range");
    CRV_REQUIRE(g->connected, "crv_group_staging_all_gather: crv_group_connect has not been called");
    CRV_REQUIRE(!g->broken, "crv_group_staging_all_gather: a barrier of this group timed out earlier");
    CRV_CUDA(cudaSetDevice(g->ctx->device));
    cudaStream_t s = as_stream(stream);
    const size_t off0 = g->staging_offset + buffer * g->staging_bytes;
    const uint32_t epoch = ++g->stage_epoch[buffer];
    if (g->nranks == 1) {
        CRV_CUDA(cudaMemcpyAsync(g->block + off0, host_src, bytes, cudaMemcpyHostToDevice, s));
        return CRV_OK;
    }
    const size_t slice = ((bytes + g->nranks - 1) / g->nranks + 255) / 256 * 256;
    const size_t lo = (size_t)g->rank * slice < bytes ? (size_t)g->rank * slice : bytes;
    const size_t len = lo + slice <= bytes ? slice : bytes - lo;
    GroupFlags* mine = reinterpret_cast<GroupFlags*>(g->block);
    uint32_t* addrs[GROUP_MAX];
    uint32_t n = 0;
    // the peers must have finished reading what the previous round left in their buffers
    for (uint32_t p = 0; p < g->nranks; ++p) if (p != g->rank) addrs[n++] = &mine->released[buffer][p];
    CRV_TRY_G(flags_wait(g, addrs, n, epoch - 1u, s));
    if (len > 0) {
        CRV_CUDA(cudaMemcpyAsync(g->block + off0 + lo, static_cast<const unsigned char*>(host_src) + lo, len, cudaMemcpyHostToDevice, s));
        // copy engines, NVLink: my slice into every peer's buffer.  Member r starts with r + 1, r + 2, ...: at any moment the
        // members write to different destinations (all starting with member 0 made its ingress the bottleneck: 80 GB/s per sender)
        for (uint32_t k = 1; k < g->nranks; ++k) {
            const uint32_t p = (g->rank + k) % g->nranks;
            CRV_CUDA(cudaMemcpyAsync(g->peer[p] + off0 + lo, g->block + off0 + lo, len, cudaMemcpyDefault, s));
        }
    }
    n = 0;
    for (uint32_t k = 1; k < g->nranks; ++k) addrs[n++] = &reinterpret_cast<GroupFlags*>(g->peer[(g->rank + k) % g->nranks])->arrived[buffer][g->rank];
    CRV_TRY_G(flags_write(g, addrs, n, epoch, s));
    n = 0;
    for (uint32_t p = 0; p < g->nranks; ++p) if (p != g->rank) addrs[n++] = &mine->arrived[buffer][p];
    CRV_TRY_G(flags_wait(g, addrs, n, epoch, s));
    return CRV_OK;
}

int crv_group_staging_release(crv_group* g, uint32_t buffer, void* stream)
{
    CRV_REQUIRE(g != nullptr, "crv_group_staging_release: null group");
    CRV_REQUIRE(buffer < g->nbuf && g->staging_bytes > 0, "crv_group_staging_release: no such staging buffer");
    if (g->nranks == 1) return CRV_OK;
    CRV_CUDA(cudaSetDevice(g->ctx->device));
    uint32_t* addrs[GROUP_MAX];
    uint32_t n = 0;
    for (uint32_t p = 0; p < g->nranks; ++p) if (p != g->rank) addrs[n++] = &reinterpret_cast<GroupFlags*>(g->peer[p])->released[buffer][g->rank];
    return flags_write(g, addrs, n, g->stage_epoch[buffer], as_stream(stream));
}

// ---- NCCL variant ----------------------------------------------------------------------------------------------------------
int crv_group_nccl_unique_id(unsigned char id_out[128])
{
    CRV_REQUIRE(id_out != nullptr, "crv_group_nccl_unique_id: null argument");
    CRV_REQUIRE(nccl_load(), "crv_group_nccl_unique_id: libnccl.so.2 is not loadable in this process");
    const int rc = g_nccl.GetUniqueId(id_out);
    if (rc != 0) {
        crv_set_error("ncclGetUniqueId failed: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "?");
        return CRV_ERR_CUDA;
    }
    return CRV_OK;
}

int crv_group_nccl_init(crv_group* g, const unsigned char id[128])
{
    CRV_REQUIRE(g && id, "crv_group_nccl_init: null argument");
    CRV_REQUIRE(nccl_load(), "crv_group_nccl_init: libnccl.so.2 is not loadable in this process");
    CRV_CUDA(cudaSetDevice(g->ctx->device));
    Id128 uid;
    std::memcpy(uid.b, id, 128);
    const int rc = g_nccl.CommInitRank(&g->nccl_comm, (int)g->nranks, uid, (int)g->rank);
    if (rc != 0) {
        crv_set_error("ncclCommInitRank failed: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "?");
        g->nccl_comm = nullptr;
        return CRV_ERR_CUDA;
    }
    return CRV_OK;
}

int crv_group_nccl_all_reduce_sum_scale(crv_group* g, uint32_t buffer, float scale, void* stream)
{
    CRV_REQUIRE(g != nullptr, "crv_group_nccl_all_reduce_sum_scale: null group");
    CRV_REQUIRE(buffer < g->nbuf, "crv_group_nccl_all_reduce_sum_scale: buffer index out of range");
    CRV_REQUIRE(g->nccl_comm != nullptr || g->nranks == 1, "crv_group_nccl_all_reduce_sum_scale: crv_group_nccl_init has not been called");
    CRV_CUDA(cudaSetDevice(g->ctx->device));
    cudaStream_t s = as_stream(stream);
    float* img = reinterpret_cast<float*>(g->block + FLAG_PAGE + buffer * g->image_bytes);
    const size_t n = (size_t)g->width * g->height * 4;
    CRV_CUDA(cudaEventRecord(g->ev0, s));
    if (g->nranks > 1) {
        const int rc = g_nccl.AllReduce(img, img, n, /* ncclFloat32 */ 7, /* ncclSum */ 0, g->nccl_comm, s);
        if (rc != 0) {
            crv_set_error("ncclAllReduce failed: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "?");
            return CRV_ERR_CUDA;
        }
    }
    group_scale_kernel<<<g->ctx->sm_count * 2, 512, 0, s>>>(reinterpret_cast<float4*>(img), (uint32_t)(n / 4), scale);
    g->ctx->launches++;
    CRV_CUDA(cudaGetLastError());
    CRV_CUDA(cudaEventRecord(g->ev1, s));
    g->timed = true;
    return CRV_OK;
}

}  // extern "C"
