/* crv_c_api.h -- C ABI of the B200-native PathTracer + ShadowVolume hot path.
 *
 * Drop-in boundary for the one hot path of jeanfilho/CloudRendering-Vulkan:
 *   shaders/PathTracer.comp   (delta tracking + HG phase + progressive accumulation)
 *   shaders/ShadowVolume.comp (sun-transmittance volume it samples)
 * behind the reference's RenderTechnique interface (RenderTechnique.h:15-66,
 * RenderTechniquePT.h:8-40, RenderTechniqueSV.h:7-39).  Vulkan handle types are replaced
 * by opaque device-resource handles:
 *   VkDescriptorBufferInfo  -> pointer to the host parameter struct (same bytes the UBO held)
 *   VkDescriptorImageInfo   -> crv_volume* / crv_image*
 *   VkCommandBuffer         -> cudaStream_t passed as void*
 * Everything is plain pointers and sizes: no C++ types, no torch types.
 *
 * Conventions
 *   - every function returns CRV_OK (0) or a negative crv_status; crv_last_error() gives a
 *     thread-local message.  Nothing aborts or throws across this boundary (the reference's
 *     ValidCheck abort()s, Validation.cpp:4-10).
 *   - parameter structs are borrowed: crv_*_queue_update_* records the pointer,
 *     crv_*_update_descriptor_sets snapshots the bytes (== VulkanBuffer::SetData +
 *     vkUpdateDescriptorSets, VulkanBuffer.cpp:33-39, RenderTechnique.cpp:14-18).
 *     PushConstants / CameraProperties(resolution) / ShadowVolumeProperties(voxelAxisCount)
 *     given to the constructors are read at record time, as the reference does
 *     (RenderTechniquePT.cpp:132,141; RenderTechniqueSV.cpp:110).
 *   - all launches are asynchronous on the caller's stream; crv_ctx_synchronize waits.
 *   - there is NO CPU fallback: every entry point that computes fails with
 *     CRV_ERR_NO_DEVICE if no sm_100-class CUDA device is usable.
 */
#ifndef CRV_C_API_H
#define CRV_C_API_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------------------------------
 * Parameter blocks: byte-for-byte the structs of UniformBuffers.h (sizes 24/68/60/28/152,
 * which are also the std140 offsets the shaders declare, PathTracer.comp:43-99).
 * ---------------------------------------------------------------------------------------- */

/* UniformBuffers.h:6-13; the shader reads the first 20 bytes (PathTracer.comp:93-99). */
typedef struct crv_push_constants {
    double   time;
    int32_t  seed;
    uint32_t frameCount;
    float    pmRadius;
    uint32_t currentBuffer;
} crv_push_constants;

/* UniformBuffers.h:15-69 / PathTracer.comp:43-55 */
typedef struct crv_camera_properties {
    float   position[3];
    int32_t halfWidth;
    float   forward[3];
    int32_t halfHeight;
    float   right[3];
    float   nearPlane;
    float   up[3];
    float   pixelSizeY;
    float   pixelSizeX;
} crv_camera_properties;

/* UniformBuffers.h:71-78 / PathTracer.comp:58-66 */
typedef struct crv_cloud_properties {
    float    bounds[2][4];
    uint32_t voxelCount[4];
    float    maxExtinction;
    float    baseScaling;
    float    densityScaling;
} crv_cloud_properties;

/* UniformBuffers.h:80-112 / PathTracer.comp:68-78.  `isotropic` is a 1-byte C++ bool followed
 * by 3 padding bytes; only the first byte is read by this library. */
typedef struct crv_parameters {
    uint32_t maxRayBounces;
    float    lightIntensity;
    float    phaseG;
    float    phaseOnePlusG2;
    float    phaseOneMinusG2;
    float    phaseOneOver2G;
    uint8_t  isotropic;
    uint8_t  _pad[3];
} crv_parameters;

/* UniformBuffers.h:114-167 / PathTracer.comp:81-91.  basisChange is column-major (glm::mat4). */
typedef struct crv_shadow_volume_properties {
    float    bounds[2][4];
    float    lightDirection[4];
    float    right[4];
    float    up[4];
    float    basisChange[4][4];
    uint32_t voxelAxisCount;
    float    voxelSize;
} crv_shadow_volume_properties;

/* Photon of the PPM technique (UniformBuffers.h:203-210, PPM_PT.comp:14-21): 48 bytes in the std430 storage buffer. */
typedef struct crv_photon {
    float position[4];
    float power[4];
    float phi;
    float theta;
    float _padding[2];
} crv_photon;

/* PhotonMapProperties (UniformBuffers.h:212-249; std140 block of PPM_PT.comp:57-67): 76 bytes.  The grid of cells covers a
 * cube: bounds[1] = bounds[0] + the cloud's largest extent (SetBounds), voxelSize = that extent / voxelCount.x. */
typedef struct crv_photon_map_properties {
    float    lightDirection[4];   /* normalize((1,-1,0,0)) by default; main.cpp:225 keeps it equal to the sun direction */
    float    bounds[2][4];
    int32_t  voxelCount[3];       /* 100^3 */
    float    voxelSize;
    uint32_t photonSize;          /* sizeof(Photon) = 48 */
    float    stepSize;            /* 10; used by the estimate pass only */
    float    absorption;          /* 0 */
} crv_photon_map_properties;

/* Extension block (NOT in the reference).  The reference path tracer has no absorption and no
 * bounce cap (PathTracer.comp:396-416).  absorption > 0 adds one RNG draw per real collision,
 * after the direct-light add and before scatterRay, and ends the path when xi <= absorption
 * (idiom of PPM_PT.comp:445-446).  With absorption == 0 no extra draw is made and the RNG
 * stream is exactly the reference's. */
typedef struct crv_pt_extensions {
    float    absorption;      /* 0 = reference semantics */
    uint32_t math_mode;       /* crv_pt_math_mode; 0 = CRV_PT_MATH_STRICT */
    uint32_t reserved[2];
} crv_pt_extensions;

/* How the tracker evaluates the one transcendental on its decision path, the free-flight update
 * `t += -log(zeta) / sigma` (PathTracer.comp:338).  GLSL leaves the precision of log(), of `/` and
 * the contraction of the add to the driver (no `precise` qualifier in the shader); the two modes
 * are the two ends of that latitude.  Nothing else differs: RNG stream, draw order, sampler,
 * phase function and accumulation are the same code.
 *   STRICT : three separately rounded operations -- logf-accurate log (<= 1 ulp), IEEE division,
 *            IEEE add -- the evaluation the CPU oracle performs with libm.
 *   DRIVER : log as lg2 on the special-function unit (absolute error < 2^-22, inside Vulkan's
 *            bound for Log), multiplication by the reciprocal, FMA-contracted add: the code a GPU
 *            Vulkan driver emits for this line.  t differs from STRICT in its last bits; a few
 *            paths per million take a different branch ("threshold-divergent", counted by the
 *            parity tests and by bench.py), all others agree within the 1e-4 radiance tolerance. */
typedef enum crv_pt_math_mode {
    CRV_PT_MATH_STRICT = 0,
    CRV_PT_MATH_DRIVER = 1
} crv_pt_math_mode;

/* ------------------------------------------------------------------------------------------ */

typedef enum crv_status {
    CRV_OK               = 0,
    CRV_ERR_INVALID      = -1,   /* bad argument / object state (e.g. nothing bound)          */
    CRV_ERR_NO_DEVICE    = -2,   /* no usable CUDA device: there is no CPU fallback           */
    CRV_ERR_CUDA         = -3,   /* a CUDA runtime call failed; see crv_last_error            */
    CRV_ERR_OOM          = -4,
    CRV_ERR_IO           = -5
} crv_status;

/* Storage layout of a 3D float volume in HBM (DESIGN.md section "data layout").  The logical
 * array is always x-fastest nx*ny*nz as Grid3D stores it (Grid3D.h:112-116). */
typedef enum crv_volume_layout {
    CRV_LAYOUT_LINEAR = 0,   /* x + nx*(y + ny*z), global memory                               */
    CRV_LAYOUT_BRICK  = 1,   /* 8x8x8 bricks (2 KiB), x-fastest inside and across bricks       */
    CRV_LAYOUT_TEX3D  = 2    /* cudaArray 3D, point-sampled texture object (no HW filtering)   */
} crv_volume_layout;

typedef enum crv_accum_mode {
    CRV_ACCUM_RUNNING_MEAN = 0,  /* out = (sample + old*frameCount)/(frameCount+1), PathTracer.comp:420-424 */
    CRV_ACCUM_SUM          = 1   /* out = old + sample (multi-GPU spp split; divide after the reduce)       */
} crv_accum_mode;

typedef enum crv_sv_mode {
    CRV_SV_MODE_AUTO        = 0,
    CRV_SV_MODE_GENERIC     = 1,   /* one thread per column, AABB test + sample per voxel: ShadowVolume.comp literally */
    CRV_SV_MODE_SINGLE_PASS = 2,   /* per-column z-interval by bisection + exact FMA division, still one kernel        */
    CRV_SV_MODE_TWO_PASS    = 3,   /* parallel gather of the in-AABB addends, then a streaming per-column scan         */
    CRV_SV_MODE_FUSED       = 4,   /* one CTA per 8x8 column bundle; brick output: chunks classified against the cloud's brick maxima,
                                      all-ones bricks stored whole, gather + serial sums (shared memory, two barriers) + exp where density is */
    CRV_SV_MODE_FUSED_V1    = 5,   /* the fused build without the chunk classification (per-voxel pre-test), kept for A/B       */
    CRV_SV_MODE_FUSED_ROWS  = 6    /* default for brick output: the same classification, every warp walks one row of the bundle on its
                                      own -- serial sums through shuffles, no CTA barrier, texel loads of two samples issued together */
} crv_sv_mode;

typedef struct crv_ctx    crv_ctx;      /* one per GPU  (~ VulkanDevice)                       */
typedef struct crv_volume crv_volume;   /* R32_SFLOAT 3D image + sampler (cloud or shadow)     */
typedef struct crv_image  crv_image;    /* R32G32B32A32_SFLOAT 2D storage image (result)       */
typedef struct crv_sv     crv_sv;       /* RenderTechniqueSV                                   */
typedef struct crv_pt     crv_pt;       /* RenderTechniquePT                                   */
typedef struct crv_ppm    crv_ppm;      /* RenderTechniquePPM, photon-tracing pass             */
typedef struct crv_group  crv_group;    /* one member of a multi-GPU sample split              */

/* what a group member publishes so that its peers can map its accumulation images (crv_group_export / crv_group_connect);
 * plain bytes, to be carried by whatever the host uses between ranks (torch.distributed, MPI, a pipe) */
/* Which pixels of the last render call went through the tracker (device pointers into the technique's own buffers, valid
 * until its next re-classification: a change of camera, cloud box, cloud contents or skip level).  tracked_xy: x | y << 16 of
 * the pixels whose paths were tracked, counts[0] entries; other_pixels: pixel index y * width + x of all others (their sample
 * is the same sampleBackground(dir) in every frame), counts[1] entries; counts lives on the device. */
typedef struct crv_pixel_lists {
    const uint32_t* tracked_xy;
    const uint32_t* other_pixels;
    const uint32_t* counts;
    uint32_t width;
    uint32_t reserved;
} crv_pixel_lists;

/* Where a group member's sums of one picture go (crv_group_push_begin fills it, crv_pt_render_batch_push consumes it): one
 * inbox image per member -- inbox[p] lies in member p's memory, mapped into this device's address space -- the word in every
 * member's flag page that announces this member's sums, and this member's CTA counter. */
typedef struct crv_push_targets {
    float*    inbox[16];
    uint32_t* flag[16];
    uint32_t* done;
    uint32_t  nranks, rank, epoch, reserved;
} crv_push_targets;

typedef struct crv_group_handle {
    uint64_t pid;            /* exporting process: members of one process use `base` directly            */
    int32_t  device;         /* CUDA device ordinal of the member                                        */
    uint32_t rank;
    uint64_t base;           /* device address of the member's block in the exporting process            */
    uint64_t bytes;          /* size of the block (flag page + images)                                   */
    unsigned char ipc[64];   /* cudaIpcMemHandle_t of the block                                          */
} crv_group_handle;

/* per-render statistics, filled by crv_pt_get_stats (all counts are exact u64 sums) */
typedef struct crv_pt_stats {
    uint64_t paths;            /* path samples traced (pixels x frames, on-image only)         */
    uint64_t tracking_steps;   /* tentative collisions (delta-tracking loop iterations)        */
    uint64_t collisions;       /* real collisions                                              */
    uint64_t density_fetches;  /* trilinear cloud fetches actually issued (rest proven empty)  */
    uint64_t kernel_launches;  /* kernels launched by this technique since creation            */
    uint64_t tracked_paths;    /* of `paths`: those that went through the tracker; the others are camera rays that miss
                                  the cloud box or (skip level 1) cross it without meeting an occupied brick, whose sample
                                  is sampleBackground(dir) whatever their random numbers are                              */
} crv_pt_stats;

/* ---- library / context ------------------------------------------------------------------ */
const char* crv_version(void);
const char* crv_last_error(void);
int  crv_device_count(void);
/* VulkanInstance/PhysicalDevice/Device (main.cpp:774-793) -> one CUDA context on `device`. */
int  crv_ctx_create(int device, crv_ctx** out);
int  crv_ctx_destroy(crv_ctx* ctx);
int  crv_ctx_synchronize(crv_ctx* ctx, void* stream);                 /* vkQueueWaitIdle        */
int  crv_ctx_launch_count(crv_ctx* ctx, uint64_t* out);               /* kernels launched so far */
/* a command stream of the context's device (a non-blocking cudaStream_t, usable wherever a `void* stream` is taken);
 * any cudaStream_t of the same device created by the caller works as well, NULL is the default stream */
int  crv_stream_create(crv_ctx* ctx, void** out);
int  crv_stream_destroy(crv_ctx* ctx, void* stream);
int  crv_ctx_sm_count(crv_ctx* ctx, int* out);

/* ---- volumes: VulkanImage(R32_SFLOAT, 3D) + VulkanSampler (main.cpp:277-285,802-804) ------ */
int  crv_volume_create(crv_ctx* ctx, uint32_t nx, uint32_t ny, uint32_t nz, int layout, crv_volume** out);
int  crv_volume_destroy(crv_volume* vol);
/* staging buffer -> image copy (main.cpp:288-295): x-fastest host floats; is_device!=0 means
 * `data` already is a device pointer (x-fastest linear). */
int  crv_volume_upload(crv_volume* vol, const float* data, int is_device, void* stream);
/* read back as x-fastest linear floats regardless of layout */
int  crv_volume_download(crv_volume* vol, float* host_out, void* stream);
int  crv_volume_dims(const crv_volume* vol, uint32_t* nx, uint32_t* ny, uint32_t* nz, int* layout);
/* Grid3D<float>::GetMajorant (Grid3D.h:181-192) as a device reduction */
int  crv_volume_majorant(crv_volume* vol, float* out, void* stream);
/* raw device pointer of a LINEAR volume (NULL otherwise); for zero-copy interop.  After WRITING through it call
 * crv_volume_mark_dirty: the majorant, the brick maxima and everything the techniques cache from them (skip tables, occupied
 * spans, pixel classification) belong to the old contents, and the skips are proofs -- stale tables give wrong pictures. */
int  crv_volume_device_ptr(crv_volume* vol, float** out);
int  crv_volume_mark_dirty(crv_volume* vol);

/* ---- result image: VulkanImage(R32G32B32A32_SFLOAT, 2D, STORAGE) (main.cpp:405-414) ------ */
int  crv_image_create(crv_ctx* ctx, uint32_t width, uint32_t height, crv_image** out);
/* wrap caller-owned device memory (width*height*4 floats), e.g. a torch tensor's data_ptr */
int  crv_image_wrap(crv_ctx* ctx, float* device_rgba, uint32_t width, uint32_t height, crv_image** out);
int  crv_image_destroy(crv_image* img);
int  crv_image_clear(crv_image* img, void* stream);
int  crv_image_download(crv_image* img, float* host_rgba, void* stream);
int  crv_image_upload(crv_image* img, const float* host_rgba, void* stream);
int  crv_image_device_ptr(crv_image* img, float** out);
/* Present step: what the reference does right after the dispatch -- vkCmdBlitImage of the RGBA32F result
 * image into the VK_FORMAT_B8G8R8A8_SRGB swapchain image (RenderTechniquePT.cpp:143-167,
 * VulkanSwapchain.cpp:68-79): clamp, linear->sRGB on R,G,B, round to UNORM8, bytes B,G,R,A.  The
 * conversion runs on the device; host_bgra8 receives width*height*4 bytes.  Synchronous. */
int  crv_image_blit_srgb8(crv_image* img, uint8_t* host_bgra8, void* stream);
/* the presented image as an 8-bit RGBA PNG / the linear result image as an RGB PFM (bottom-up rows) */
int  crv_image_save_png(crv_image* img, const char* path, void* stream);
int  crv_image_save_pfm(crv_image* img, const char* path, void* stream);
/* accumulation checkpoint / resume: result image + the frameCount the next dispatch pushes + last seed */
int  crv_image_save_checkpoint(crv_image* img, uint32_t frame_count, int32_t seed, const char* path, void* stream);
int  crv_image_load_checkpoint(crv_image* img, uint32_t* frame_count, int32_t* seed, const char* path, void* stream);
/* host-only helper behind crv_image_save_png: 8-bit RGBA PNG with stored (uncompressed) deflate blocks */
int  crv_write_png_rgba8(const char* path, uint32_t width, uint32_t height, const unsigned char* rgba);
/* out = in * scale (the divide-by-spp epilogue after the multi-GPU sum reduce) */
int  crv_image_scale(crv_image* img, float scale, void* stream);

/* ---- multi-GPU sample split (no counterpart in the single-GPU reference) -----------------------------------
 * north_star: the density grid and the shadow volume are replicated, the frames (samples per pixel) of one picture are
 * split across the GPUs of one node -- legal because a sample depends on (frame seed, pixel) only, PathTracer.comp:368,
 * and the mean is the only coupling, :420-424 -- every GPU accumulates a plain fp32 SUM of its frames (CRV_ACCUM_SUM)
 * and the buffers are combined and divided by the total sample count.
 *
 * A group has one member per GPU: one per process (torchrun, MPI) or several in one process (crv_headless --gpus N).
 * Each member owns `nbuf` RGBA32F accumulation images (render targets for crv_pt_set_frame_references) in memory its
 * peers can map over NVLink:
 *     crv_group_create(ctx, rank, nranks, w, h, nbuf, 0, &g); crv_group_export(g, &mine);
 *     ... exchange the handles, all members ...            crv_group_connect(g, all_handles_in_rank_order, nranks);
 *     crv_group_image(g, 0, &img); crv_pt_set_frame_references(pt, &img, 1);
 *     per picture, on every member:  crv_image_clear(img) ; crv_pt_render_batch(pt, my_seeds, n, 0, CRV_ACCUM_SUM, ...)
 *                                    crv_group_all_reduce_sum_scale(g, 0, 1.0f / spp_total, stream)
 * crv_group_all_reduce_sum_scale enqueues, per member: a start barrier with its peers, ONE kernel that sums its band of
 * rows from all members' images in rank order, scales it and writes it into all members' images, and an end barrier:
 * afterwards every member's image holds the same, final picture (bit-identical on all members, and independent of timing).
 * The barriers are stream memory operations on flags in mapped memory (no SM is held while waiting).  All members must call
 * it the same number of times with the same buffer index; each call is asynchronous on the caller's stream.
 * crv_group_status waits for the stream with the group's deadline (crv_group_set_timeout_ms, default 10 s): if a peer never
 * arrives it releases the wait, returns an error and the group is unusable from then on -- use it instead of a plain
 * synchronise wherever a member might be missing.
 * crv_group_nccl_*: the same operation as ncclAllReduce + a scale kernel through the libnccl.so.2 already loaded in the
 * process (rank 0 creates the id, everybody passes it to crv_group_nccl_init). */
int  crv_group_create(crv_ctx* ctx, uint32_t rank, uint32_t nranks, uint32_t width, uint32_t height, uint32_t nbuf, uint64_t staging_bytes,
                      crv_group** out);
int  crv_group_destroy(crv_group* g);
int  crv_group_export(crv_group* g, crv_group_handle* out);
int  crv_group_connect(crv_group* g, const crv_group_handle* handles, uint32_t count);
int  crv_group_image(crv_group* g, uint32_t buffer, crv_image** out);              /* owned by the group */
int  crv_group_all_reduce_sum_scale(crv_group* g, uint32_t buffer, float scale, void* stream);
/* The same for members that hold replicas of one scene (same camera, cloud and skip level, hence the same pixel lists):
 * only the tracked pixels are combined across the members (sum in rank order x scale, into all members' images); every other
 * pixel holds n_mine identical samples and is finished locally as value x others_scale (= scale x n_total / n_mine).  At 1080p
 * the c2 scene moves 2 MB instead of 33 MB per member.  `lists` = crv_pt_pixel_lists of this member's own technique. */
int  crv_group_all_reduce_lists_sum_scale(crv_group* g, uint32_t buffer, float scale, float others_scale, const crv_pixel_lists* lists, void* stream);
/* One-barrier combine for replicas of one scene (the fast path of the split; the calls above are the general ones).
 *     crv_group_push_begin(g, b, &targets)                                        next picture of buffer b: inbox slots + flags
 *     crv_pt_render_batch_push(pt, my_seeds, n, &targets, others_scale, stream)   tracker + fold: no clear, no private sum image
 *     crv_group_push_finish(g, b, scale, &lists, stream)                          wait for the members' flags, then one local kernel
 * The fold stores every tracked pixel's sum into the inbox each member keeps for this member (peer stores over NVLink, under
 * the kernel) and its last CTA raises this member's flag at every peer; every other pixel is finished by the fold itself in
 * the group image ((n * sampleBackground) * others_scale).  crv_group_push_finish waits for the peers' flags (stream memory
 * operations) and runs one kernel on local memory only: tracked pixel = (inbox[0] + inbox[1] + ...) * scale, rank order --
 * the same bits as crv_group_all_reduce_lists_sum_scale, on every member.  One barrier instead of two, no all-gather, and the
 * transfer is hidden behind the fold.  Inboxes are double-buffered per image buffer (a member may run one picture ahead of
 * its peers); all calls concerning one buffer must be ordered on one stream.  Memory: 2 * nranks images per buffer. */
int  crv_group_push_begin(crv_group* g, uint32_t buffer, crv_push_targets* out);
int  crv_group_push_finish(crv_group* g, uint32_t buffer, float scale, const crv_pixel_lists* lists, void* stream);
int  crv_group_set_timeout_ms(crv_group* g, uint32_t ms);
int  crv_group_status(crv_group* g, void* stream);                                  /* waits for `stream` (with the deadline); CRV_OK or the time-out */
int  crv_group_last_reduce_ms(crv_group* g, float* ms);                             /* device time of the last reduce call of this member */
/* Replicated input from the host (the density grid is the same on every member): with staging_bytes > 0 every member also
 * owns `nbuf` staging buffers.  crv_group_staging_all_gather: member r copies slice r (1/N) of the host array into its
 * buffer and pushes it into every peer's buffer with copy-engine peer copies; when the call has completed on `stream` the
 * member's staging buffer (crv_group_staging_ptr, e.g. for crv_volume_upload(..., is_device = 1)) holds the whole array.
 * So the array crosses PCIe once per job, on N links in parallel, and no kernel is involved (the waits and signals are
 * stream memory operations): the exchange overlaps a kernel that holds every SM.  crv_group_staging_release tells the
 * peers that this member has finished reading its buffer; the next all-gather into the same buffer index waits for it. */
int  crv_group_staging_ptr(crv_group* g, uint32_t buffer, void** device_ptr_out);
int  crv_group_staging_all_gather(crv_group* g, uint32_t buffer, const void* host_src, uint64_t bytes, void* stream);
int  crv_group_staging_release(crv_group* g, uint32_t buffer, void* stream);
int  crv_group_nccl_unique_id(unsigned char id_out[128]);
int  crv_group_nccl_init(crv_group* g, const unsigned char id[128]);
int  crv_group_nccl_all_reduce_sum_scale(crv_group* g, uint32_t buffer, float scale, void* stream);

/* ---- RenderTechniqueSV (RenderTechniqueSV.h:7-39, .cpp:4-113) ----------------------------- */
int  crv_sv_create(crv_ctx* ctx, const crv_shadow_volume_properties* props, const crv_push_constants* pc, crv_sv** out);
int  crv_sv_destroy(crv_sv* sv);
int  crv_sv_set_frame_references(crv_sv* sv, crv_volume* shadow_image);                     /* binding 0 */
int  crv_sv_clear_frame_references(crv_sv* sv);
int  crv_sv_queue_update_cloud_data(crv_sv* sv, const crv_cloud_properties* cloud);          /* binding 3 */
int  crv_sv_queue_update_cloud_data_sampler(crv_sv* sv, crv_volume* cloud_image);            /* binding 2 */
int  crv_sv_queue_update_camera_properties(crv_sv* sv, const crv_camera_properties* cam);    /* no-op, .cpp:79-81 */
int  crv_sv_queue_update_parameters(crv_sv* sv, const crv_parameters* params);               /* no-op, .cpp:83-85 */
int  crv_sv_queue_update_shadow_volume(crv_sv* sv, const crv_shadow_volume_properties* p);   /* binding 1 */
int  crv_sv_queue_update_shadow_volume_sampler(crv_sv* sv, crv_volume* shadow_image);        /* binding 0, .cpp:92-95 */
int  crv_sv_update_descriptor_sets(crv_sv* sv);
/* vkCmdDispatch(N/32+1, N/32+1, 1) of ShadowVolume.comp (RenderTechniqueSV.cpp:102-113) */
int  crv_sv_record_draw_commands(crv_sv* sv, void* stream, unsigned image_index);
/* device time of the last recorded build in ms (CUDA events on `stream`); syncs the stream */
int  crv_sv_last_build_ms(crv_sv* sv, float* ms);
/* Which kernels build the volume; every mode produces the same bits (tests/test_gpu_parity.py).  AUTO picks the
 * fastest mode whose preconditions hold (finite parameters, voxelSize > 0, N < 65536), else GENERIC. */
int  crv_sv_set_tuning(crv_sv* sv, int mode /* crv_sv_mode */);
int  crv_sv_last_build_mode(crv_sv* sv, int* out);

/* ---- RenderTechniquePT (RenderTechniquePT.h:8-40, .cpp:12-168) ---------------------------- */
int  crv_pt_create(crv_ctx* ctx, const crv_camera_properties* cam, const crv_push_constants* pc, crv_pt** out);
int  crv_pt_destroy(crv_pt* pt);
int  crv_pt_set_frame_references(crv_pt* pt, crv_image* const* result_images, unsigned count); /* binding 0 */
int  crv_pt_clear_frame_references(crv_pt* pt);
int  crv_pt_queue_update_cloud_data(crv_pt* pt, const crv_cloud_properties* cloud, unsigned image_index);          /* binding 3 */
int  crv_pt_queue_update_cloud_data_sampler(crv_pt* pt, crv_volume* cloud_image, unsigned image_index);            /* binding 2 */
int  crv_pt_queue_update_camera_properties(crv_pt* pt, const crv_camera_properties* cam, unsigned image_index);    /* binding 1 */
int  crv_pt_queue_update_parameters(crv_pt* pt, const crv_parameters* params, unsigned image_index);               /* binding 4 */
int  crv_pt_queue_update_shadow_volume(crv_pt* pt, const crv_shadow_volume_properties* p, unsigned image_index);   /* binding 6 */
int  crv_pt_queue_update_shadow_volume_sampler(crv_pt* pt, crv_volume* shadow_image, unsigned image_index);        /* binding 5 */
int  crv_pt_update_descriptor_sets(crv_pt* pt);
/* One progressive frame with the exact shader semantics: seed and frameCount are read from the
 * PushConstants given at creation (vkCmdPushConstants, RenderTechniquePT.cpp:132) and the result
 * image `image_index` is updated in place with the running mean (PathTracer.comp:420-424). */
int  crv_pt_record_draw_commands(crv_pt* pt, void* stream, unsigned image_index);

/* ---- extensions beyond the reference interface -------------------------------------------- */
int  crv_pt_set_extensions(crv_pt* pt, const crv_pt_extensions* ext);
/* n progressive frames in one call: identical, bit for bit, to n crv_pt_record_draw_commands
 * calls with seed = seeds[i] and frameCount = first_frame_count + i (CRV_ACCUM_RUNNING_MEAN),
 * or the fp32 sum of the n samples added to the image (CRV_ACCUM_SUM: the multi-GPU split's mode; deterministic, but the
 * order of the additions is the library's choice -- a pixel whose sample is the same in every frame gets n * sample).
 * This is the persistent-thread wavefront fast path.  PushConstants are not modified. */
int  crv_pt_render_batch(crv_pt* pt, const int32_t* seeds, unsigned n, uint32_t first_frame_count,
                         int accum_mode, void* stream, unsigned image_index);
/* enable per-path counters (steps/collisions/fetches); costs a few atomics per CTA */
int  crv_pt_get_stats(crv_pt* pt, crv_pt_stats* out, void* stream);
/* crv_pt_render_batch in CRV_ACCUM_SUM mode for a group member (see crv_group_push_begin): the image needs no clear; after
 * the call (in stream order) it holds the final value of every pixel that was not tracked, the tracked pixels' sums are on
 * their way to the members' inboxes. */
int  crv_pt_render_batch_push(crv_pt* pt, const int32_t* seeds, unsigned n, const crv_push_targets* targets, float others_scale,
                              void* stream, unsigned image_index);
/* the pixel lists of the last render call (see crv_pixel_lists) */
int  crv_pt_pixel_lists(crv_pt* pt, crv_pixel_lists* out);
int  crv_pt_reset_stats(crv_pt* pt, void* stream);
/* Device time (CUDA events on the render's stream) spent in the path-tracing kernel launches of the last
 * crv_pt_record_draw_commands / crv_pt_render_batch call, and how many launches that was; syncs on them. */
int  crv_pt_last_render_ms(crv_pt* pt, float* tracker_ms, unsigned* tracker_launches);
/* Debug/parity taps for ONE frame: per-pixel tracking steps and collisions (u32 each, W*H,
 * host pointers, may be NULL) of the last crv_pt_record_draw_commands. */
int  crv_pt_set_debug_counters(crv_pt* pt, int enable);
int  crv_pt_download_debug_counters(crv_pt* pt, uint32_t* steps, uint32_t* collisions, void* stream);
/* First `n_draws` uniform draws (raw u32 states after warm-up) for every pixel of a WxH image
 * with dispatch stride 32*(W/32+1) (PathTracer.comp:220-241,368): bit-exact RNG tap. */
int  crv_pt_rng_tap(crv_ctx* ctx, int32_t seed, uint32_t width, uint32_t height, uint32_t n_draws,
                    uint32_t* host_out, void* stream);
/* Self-test: the tracker's -log(zeta) (the math library's logf algorithm restricted to zeta in {0} U [2^-32, 1]) against
 * logf for EVERY float of that set; returns the number of inputs whose results differ in any bit (expected 0). */
int  crv_pt_selftest_log(crv_ctx* ctx, uint64_t* mismatches, void* stream);
/* tuning knobs (0 = library default): CTAs per SM, min tracking lanes before servicing, max
 * bytes of the per-round sample buffer, tracking steps between two scheduling decisions (lanes scheduler only: the
 * pool scheduler's round is unrolled at compile time) */
int  crv_pt_set_tuning(crv_pt* pt, int ctas_per_sm, int track_threshold, uint64_t sample_buffer_bytes, int steps_per_decision);
/* Proof-based skipping of ray sections that cannot collide (identical images, bit for bit; default 1):
 *   1  a brick of 8^3 voxels whose texels (and the ring of texels around it) are all zero cannot produce a collision: every
 *      tentative collision of PathTracer.comp:335-355 inside it is rejected whatever the random numbers are.  Camera rays
 *      that cross the cloud box through such bricks only become background pixels; a segment is tracked up to its last
 *      occupied brick only (the path's sample is final there); the stretch in front of the first occupied brick advances
 *      the RNG and t without table lookups.  crv_pt_stats.tracking_steps then counts the steps that were executed.
 *   0  every step the reference executes is executed (the stretch in front of the first occupied brick still without
 *      lookups): tracking_steps / collisions and the debug counters equal the reference's.  The debug counters force 0. */
int  crv_pt_set_skip_level(crv_pt* pt, int level);
/* The tracker's coarse threshold table holds one byte per cell of (8 << shift)^3 voxels; shift is the smallest that fits
 * its shared-memory budget (0 up to 304^3 voxels, 1 at 512^3, 2 at 1024^3).  A larger minimum can be requested (identical
 * results, bit for bit; -1 = automatic): the tests run the production table shapes on small grids with it.
 * crv_pt_get_coarse_shift: the shift of the tables built by the last render call. */
int  crv_pt_set_coarse_shift(crv_pt* pt, int min_shift);
int  crv_pt_get_coarse_shift(crv_pt* pt, int* shift);
/* In which order the (pixel, frame) jobs of a launch are handed to the tracker (identical samples, bit for bit):
 *   0  frame-major over the hit list as the classification left it (neighbouring lanes start on neighbouring pixels)
 *   1  pixel-major over the hit list sorted by descending expected work per path (span length and optical depth of the
 *      pixel's first segment): all frames of the longest pixels start first, so the paths still in flight when the job
 *      queue runs dry are short -- the tail of the launch shrinks, which is what a member of a sample split with few
 *      frames per GPU pays for
 *  -1  automatic (default): 1 for launches of at least 4 frames, else 0 */
int  crv_pt_set_job_order(crv_pt* pt, int order);
/* which persistent-thread scheduler runs the tracker (identical results, bit for bit):
 *   0 = pool  (default) every warp owns 128 paths in shared memory and runs one kind of work at a time on 32 of them
 *   1 = lanes every lane owns one path plus one in reserve (kept for A/B measurements) */
int  crv_pt_set_scheduler(crv_pt* pt, int scheduler);

/* ---- RenderTechniquePPM (RenderTechniquePPM.h / .cpp; shaders/PPM_PT.comp + PPM_PE.comp) ------------------------
 * Progressive photon mapping.  Every frame (RecordDrawCommands, RenderTechniquePPM.cpp:229-273):
 *   UpdateRadius  pmRadius shrinks with the frame number (:302-312), written into the caller's PushConstants
 *   tracing       64 * groups_x photons from the sun and the sky are delta-tracked through the cloud with the path
 *                 tracer's code and deposited into a uniform grid of cells (32 slots each, an atomic counter per cell)
 *   estimate      one thread per pixel marches the camera ray through the cloud and adds, per step, the sun term through
 *                 the shadow volume, a sky term and the photon-map density estimate; running mean into the result image
 * Tracing bindings: 0 photon map, 1 PhotonMapProperties, 2 cloud sampler, 3 CloudProperties, 4 Parameters, 5 collision
 * map (RenderTechniquePPM.cpp:22-38).  Estimate bindings: 0 result image, 1 CameraProperties, 2 photon map,
 * 3 CloudProperties, 4 Parameters, 5 PhotonMapProperties, 6 collision map, 7 cloud sampler, 8 shadow sampler,
 * 9 ShadowVolumeProperties (:59-81).  The swapchain blit is crv_image_blit_srgb8. */
int  crv_ppm_create(crv_ctx* ctx, const crv_photon_map_properties* props, crv_push_constants* pc, float initial_radius /* main.cpp:813: 10 */,
                    crv_ppm** out);
int  crv_ppm_destroy(crv_ppm* ppm);
/* AllocateResources / FreeResources (RenderTechniquePPM.cpp:113-160): photon map = cells*32 photons, collision map =
 * cells uvec4, sized from the PhotonMapProperties given at creation; both start zeroed */
int  crv_ppm_allocate_resources(crv_ppm* ppm);
int  crv_ppm_free_resources(crv_ppm* ppm);
int  crv_ppm_set_frame_references(crv_ppm* ppm, crv_image* const* result_images, unsigned count);
int  crv_ppm_queue_update_cloud_data(crv_ppm* ppm, const crv_cloud_properties* cloud);
int  crv_ppm_queue_update_cloud_data_sampler(crv_ppm* ppm, crv_volume* cloud_image);
int  crv_ppm_queue_update_parameters(crv_ppm* ppm, const crv_parameters* params);
int  crv_ppm_queue_update_camera_properties(crv_ppm* ppm, const crv_camera_properties* cam);
int  crv_ppm_queue_update_shadow_volume(crv_ppm* ppm, const crv_shadow_volume_properties* p);
int  crv_ppm_queue_update_shadow_volume_sampler(crv_ppm* ppm, crv_volume* shadow_image);
int  crv_ppm_update_descriptor_sets(crv_ppm* ppm);      /* also snapshots the PhotonMapProperties (SetData) */
/* RecordDrawCommands (:229-273): UpdateRadius(frameCount), tracing with the reference's 200 groups, estimate */
int  crv_ppm_record_draw_commands(crv_ppm* ppm, void* stream, unsigned image_index);
/* its parts: the radius update, the tracing dispatch (clear both buffers, groups_x x 1 x 1 groups of 64 photons;
 * groups_x = 0 means 200) and the estimate dispatch over the current photon map; seed / frameCount / pmRadius are read
 * from the PushConstants at the call */
int  crv_ppm_update_radius(crv_ppm* ppm, unsigned frame_number);
int  crv_ppm_record_photon_tracing(crv_ppm* ppm, void* stream, uint32_t groups_x);
int  crv_ppm_record_photon_estimate(crv_ppm* ppm, void* stream, unsigned image_index);
/* photons: cells*32 records, collisions: cells*4 u32 (x = deposits in the cell; may exceed 32); either may be NULL */
int  crv_ppm_download(crv_ppm* ppm, crv_photon* host_photons, uint32_t* host_collisions, void* stream);
/* replace the photon map by host data (tests / tooling): an estimate over a known map */
int  crv_ppm_upload(crv_ppm* ppm, const crv_photon* host_photons, const uint32_t* host_collisions, void* stream);
/* tracking iterations / deposits / deposits dropped because their cell was full, of the last tracing dispatch; device ms */
int  crv_ppm_last_trace_stats(crv_ppm* ppm, uint64_t* steps, uint64_t* deposits, uint64_t* dropped, float* ms);
/* ray-march steps and device ms of the last estimate dispatch */
int  crv_ppm_last_estimate_stats(crv_ppm* ppm, uint64_t* march_steps, float* ms);
void crv_photon_map_init(crv_photon_map_properties* p);                                             /* :212-222 */
void crv_photon_map_set_bounds(crv_photon_map_properties* p, const float bounds[2][4]);             /* :225-243 */
uint32_t crv_photon_map_total_size(const crv_photon_map_properties* p);                             /* :245-248 */

/* ---- host-side helpers with the reference's setter semantics (UniformBuffers.h) ------------ */
void crv_push_constants_init(crv_push_constants* p);                                   /* :6-13   */
void crv_camera_init(crv_camera_properties* c);                                        /* :15-27  */
void crv_camera_set_resolution(crv_camera_properties* c, int width, int height);       /* :46-50  */
void crv_camera_set_rotation(crv_camera_properties* c, float* rotation_xy);            /* :29-35  */
void crv_camera_set_fov(crv_camera_properties* c, float* fov);                         /* :52-58  */
void crv_cloud_init(crv_cloud_properties* c);                                          /* :71-78  */
/* SetCloudProperties (main.cpp:144-163): bounds, voxelCount, maxExtinction = max(majorant, .01) */
void crv_cloud_set_from_grid(crv_cloud_properties* c, const uint32_t counts[3], const double voxel_size[3], float majorant);
void crv_parameters_init(crv_parameters* p);                                           /* :80-91  */
void crv_parameters_set_phase_g(crv_parameters* p, float g);                           /* :93-106 */
void crv_shadow_init(crv_shadow_volume_properties* s);                                 /* :114-127*/
void crv_shadow_set_light_direction(crv_shadow_volume_properties* s, const float dir[3]);            /* :131-143 */
void crv_shadow_set_origin(crv_shadow_volume_properties* s, const float lower[3], const float upper[3]); /* :145-151 */
void crv_orthonormal_basis(const float n[3], float right[3], float up[3]);             /* Utilities.cpp:7-16 */
/* MSVC rand() sequence after srand(seed0) (main.cpp:722,949): frame seeds */
void crv_msvc_rand_sequence(uint32_t seed0, unsigned n, int32_t* out);

/* ---- Grid3D .xyz files (Grid3D.h:45-90,167-178; GridData.h) -------------------------------- */
/* Load: 3 x u32 counts, 3 x f64 voxel sizes, floats z-fastest -> x-fastest out.  Two-call
 * protocol: data == NULL returns the header only.  Missing file -> CRV_ERR_IO (Load returns nullptr). */
int  crv_grid_load(const char* path, uint32_t counts[3], double voxel_size[3], float* data, size_t capacity);
/* Save writes the raw x-fastest array after the header, exactly as Grid3D::Save does (:167-178). */
int  crv_grid_save(const char* path, const uint32_t counts[3], const double voxel_size[3], const float* data);
/* GridData::save convention (GridData.h:12-25): z-fastest on disk, what Load expects. */
int  crv_grid_save_zfastest(const char* path, const uint32_t counts[3], const double voxel_size[3], const float* data_xfastest);

/* ---- synthetic scene (bench/test scaffolding; SURVEY section 8d generator) ------------------ */
/* Fill a LINEAR/BRICK/TEX3D volume with the frozen fBm cloud of side n (seed 1234 by default). */
int  crv_fbm_generate(crv_volume* vol, uint32_t seed, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CRV_C_API_H */
