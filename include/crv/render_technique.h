// crv/render_technique.h -- the reference's RenderTechnique interface, headless, over the C ABI.
//
// Mirrors RenderTechnique.h:15-66 (abstract operator API), RenderTechniquePT.h:8-40 / .cpp:12-168 and
// RenderTechniqueSV.h:7-39 / .cpp:4-113 of jeanfilho/CloudRendering-Vulkan: same class names, same method set,
// same call order (ctor -> SetFrameReferences -> QueueUpdate* -> UpdateDescriptorSets -> RecordDrawCommands),
// same ownership (the technique borrows the parameter structs and images, the caller owns them).
// Vulkan handle types are replaced one for one:
//
//     VulkanDevice*              -> crv::Device*          (one CUDA device, crv_ctx)
//     VulkanImage* (R32F 3D)     -> crv::Volume*          (cloud density / shadow volume, crv_volume)
//     VulkanImage* (RGBA32F 2D)  -> crv::Image*           (progressive result image, crv_image)
//     VkDescriptorBufferInfo&    -> crv::DescriptorBufferInfo { pointer to the host parameter struct }
//     VkDescriptorImageInfo&     -> crv::DescriptorImageInfo  { crv::Volume* }
//     VkCommandBuffer            -> crv::CommandStream    (a cudaStream_t passed as void*, nullptr = default stream)
//
// Dropped, because they are Vulkan plumbing with no counterpart here: GetDescriptorSetLayout,
// GetDescriptorPoolSizes, the VulkanSwapchain* argument and the blit to the swapchain image
// (RenderTechniquePT.cpp:143-167).  GetRequiredSetCount is kept (it sizes the per-image binding tables).
//
// Errors: the C ABI returns status codes; this wrapper throws crv::Error (std::runtime_error) carrying the code
// and crv_last_error().  The reference abort()s on a failed VkResult (Validation.cpp:4-10).
// Header-only; link with -lcrv_b200.  Everything computes on the GPU: there is no CPU fallback.
#pragma once

#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "../crv_c_api.h"
#include "uniform_buffers.h"

namespace crv {

class Error : public std::runtime_error {
public:
    Error(int code, const std::string& what) : std::runtime_error(what), m_code(code) {}
    int code() const { return m_code; }

private:
    int m_code;
};

inline void Check(int rc)
{
    if (rc != CRV_OK) throw Error(rc, std::string(crv_last_error()));
}

using CommandStream = void*;

// VulkanInstance + VulkanPhysicalDevice + VulkanDevice (main.cpp:774-793)
class Device {
public:
    explicit Device(int index = 0) { Check(crv_ctx_create(index, &m_ctx)); }
    ~Device() { crv_ctx_destroy(m_ctx); }
    Device(const Device&) = delete;
    Device& operator=(const Device&) = delete;
    crv_ctx* Get() const { return m_ctx; }
    void WaitIdle(CommandStream stream = nullptr) { Check(crv_ctx_synchronize(m_ctx, stream)); }     // vkQueueWaitIdle
    uint64_t LaunchCount() const
    {
        uint64_t n = 0;
        Check(crv_ctx_launch_count(m_ctx, &n));
        return n;
    }

private:
    crv_ctx* m_ctx = nullptr;
};

// VulkanImage(R32_SFLOAT, 3D) + VulkanImageView + VulkanSampler(LINEAR, CLAMP_TO_EDGE) (main.cpp:277-285, 802-804)
class Volume {
public:
    Volume(Device* device, uint32_t nx, uint32_t ny, uint32_t nz, crv_volume_layout layout = CRV_LAYOUT_BRICK)
    {
        Check(crv_volume_create(device->Get(), nx, ny, nz, layout, &m_vol));
    }
    ~Volume() { crv_volume_destroy(m_vol); }
    Volume(const Volume&) = delete;
    Volume& operator=(const Volume&) = delete;
    crv_volume* Get() const { return m_vol; }
    // staging-buffer -> image copy (main.cpp:288-295): x-fastest host floats, as Grid3D::GetData() holds them
    void Upload(const float* host, CommandStream stream = nullptr) { Check(crv_volume_upload(m_vol, host, 0, stream)); }
    void Download(float* host, CommandStream stream = nullptr) { Check(crv_volume_download(m_vol, host, stream)); }
    // Grid3D::GetMajorant (Grid3D.h:181-192) as a device reduction
    float GetMajorant(CommandStream stream = nullptr)
    {
        float m = 0.0f;
        Check(crv_volume_majorant(m_vol, &m, stream));
        return m;
    }
    void GenerateFbm(uint32_t seed = 1234, CommandStream stream = nullptr) { Check(crv_fbm_generate(m_vol, seed, stream)); }

private:
    crv_volume* m_vol = nullptr;
};

// VulkanImage(R32G32B32A32_SFLOAT, 2D, STORAGE): one progressive result image (main.cpp:405-414)
class Image {
public:
    Image(Device* device, uint32_t width, uint32_t height) : m_width(width), m_height(height)
    {
        Check(crv_image_create(device->Get(), width, height, &m_img));
    }
    // an image owned by somebody else (a DeviceGroup's accumulation image): not destroyed with this object
    Image(crv_image* borrowed, uint32_t width, uint32_t height) : m_img(borrowed), m_width(width), m_height(height), m_owned(false) {}
    ~Image() { if (m_owned) crv_image_destroy(m_img); }
    Image(const Image&) = delete;
    Image& operator=(const Image&) = delete;
    crv_image* Get() const { return m_img; }
    uint32_t Width() const { return m_width; }
    uint32_t Height() const { return m_height; }
    void Clear(CommandStream stream = nullptr) { Check(crv_image_clear(m_img, stream)); }
    void Download(float* host_rgba, CommandStream stream = nullptr) { Check(crv_image_download(m_img, host_rgba, stream)); }
    void Scale(float s, CommandStream stream = nullptr) { Check(crv_image_scale(m_img, s, stream)); }
    // present step (RenderTechniquePT.cpp:143-167): the result image as the B8G8R8A8_SRGB swapchain image would hold it
    void BlitSrgb8(uint8_t* host_bgra8, CommandStream stream = nullptr) { Check(crv_image_blit_srgb8(m_img, host_bgra8, stream)); }
    void SavePng(const char* path, CommandStream stream = nullptr) { Check(crv_image_save_png(m_img, path, stream)); }
    void SavePfm(const char* path, CommandStream stream = nullptr) { Check(crv_image_save_pfm(m_img, path, stream)); }
    void SaveCheckpoint(const char* path, uint32_t frameCount, int32_t seed, CommandStream stream = nullptr)
    {
        Check(crv_image_save_checkpoint(m_img, frameCount, seed, path, stream));
    }
    void LoadCheckpoint(const char* path, uint32_t* frameCount, int32_t* seed, CommandStream stream = nullptr)
    {
        Check(crv_image_load_checkpoint(m_img, frameCount, seed, path, stream));
    }

private:
    crv_image* m_img = nullptr;
    uint32_t m_width, m_height;
    bool m_owned = true;
};

// a command stream of the device (VkQueue role): a non-blocking cudaStream_t
class Stream {
public:
    explicit Stream(Device* device) : m_device(device) { Check(crv_stream_create(device->Get(), &m_stream)); }
    ~Stream() { crv_stream_destroy(m_device->Get(), m_stream); }
    Stream(const Stream&) = delete;
    Stream& operator=(const Stream&) = delete;
    CommandStream Get() const { return m_stream; }

private:
    Device* m_device;
    void* m_stream = nullptr;
};

// One member of a multi-GPU sample split (crv_group_*; SURVEY 8e -- no counterpart in the single-GPU reference).  Owns the
// accumulation image(s) its RenderTechniquePT renders into with CRV_ACCUM_SUM, in memory the other members map over NVLink;
// AllReduceSumScale adds all members' images in rank order, scales and leaves the same picture in every member's image.
class DeviceGroup {
public:
    DeviceGroup(Device* device, uint32_t rank, uint32_t nranks, uint32_t width, uint32_t height, uint32_t nbuf = 1, uint64_t stagingBytes = 0)
        : m_width(width), m_height(height)
    {
        Check(crv_group_create(device->Get(), rank, nranks, width, height, nbuf, stagingBytes, &m_group));
    }
    // replicated host input: 1/N per member over PCIe, exchanged GPU to GPU by the copy engines (crv_group_staging_*)
    void* StagingAllGather(const void* hostSrc, uint64_t bytes, CommandStream stream = nullptr, uint32_t buffer = 0)
    {
        void* p = nullptr;
        Check(crv_group_staging_all_gather(m_group, buffer, hostSrc, bytes, stream));
        Check(crv_group_staging_ptr(m_group, buffer, &p));
        return p;
    }
    void StagingRelease(CommandStream stream = nullptr, uint32_t buffer = 0) { Check(crv_group_staging_release(m_group, buffer, stream)); }
    ~DeviceGroup() { crv_group_destroy(m_group); }
    DeviceGroup(const DeviceGroup&) = delete;
    DeviceGroup& operator=(const DeviceGroup&) = delete;
    crv_group_handle Export()
    {
        crv_group_handle h{};
        Check(crv_group_export(m_group, &h));
        return h;
    }
    void Connect(const std::vector<crv_group_handle>& handlesInRankOrder) { Check(crv_group_connect(m_group, handlesInRankOrder.data(), (uint32_t)handlesInRankOrder.size())); }
    // the member's accumulation image `buffer` as a (borrowed) crv::Image, e.g. for SetFrameReferences
    Image GetImage(uint32_t buffer = 0)
    {
        crv_image* img = nullptr;
        Check(crv_group_image(m_group, buffer, &img));
        return Image(img, m_width, m_height);
    }
    void AllReduceSumScale(float scale, CommandStream stream = nullptr, uint32_t buffer = 0) { Check(crv_group_all_reduce_sum_scale(m_group, buffer, scale, stream)); }
    // members that hold replicas of one scene: only the tracked pixels travel (lists = this member's RenderTechniquePT::GetPixelLists())
    void AllReduceListsSumScale(float scale, float othersScale, const crv_pixel_lists& lists, CommandStream stream = nullptr, uint32_t buffer = 0)
    {
        Check(crv_group_all_reduce_lists_sum_scale(m_group, buffer, scale, othersScale, &lists, stream));
    }
    // one-barrier combine: PushBegin() -> RenderTechniquePT::RenderBatchPush(seeds, n, targets, othersScale) -> PushFinish()
    crv_push_targets PushBegin(uint32_t buffer = 0)
    {
        crv_push_targets t{};
        Check(crv_group_push_begin(m_group, buffer, &t));
        return t;
    }
    void PushFinish(float scale, const crv_pixel_lists& lists, CommandStream stream = nullptr, uint32_t buffer = 0)
    {
        Check(crv_group_push_finish(m_group, buffer, scale, &lists, stream));
    }
    void CheckStatus(CommandStream stream = nullptr) { Check(crv_group_status(m_group, stream)); }
    float LastReduceMs()
    {
        float ms = 0.0f;
        Check(crv_group_last_reduce_ms(m_group, &ms));
        return ms;
    }
    void SetTimeoutMs(uint32_t ms) { Check(crv_group_set_timeout_ms(m_group, ms)); }

private:
    crv_group* m_group = nullptr;
    uint32_t m_width, m_height;
};

// VkDescriptorBufferInfo: which host parameter struct backs the uniform buffer (the bytes are snapshotted by
// UpdateDescriptorSets, like VulkanBuffer::SetData + vkUpdateDescriptorSets)
struct DescriptorBufferInfo {
    const void* buffer = nullptr;
    size_t range = 0;
};
// VkDescriptorImageInfo: the sampled / storage 3D image
struct DescriptorImageInfo {
    Volume* image = nullptr;
};
// element of the frameImages vector of SetFrameReferences: a 2D result image (PT) or the 3D shadow image (SV)
struct FrameImage {
    Image* image2D = nullptr;
    Volume* image3D = nullptr;
    FrameImage(Image* i) : image2D(i) {}
    FrameImage(Volume* v) : image3D(v) {}
};

// RenderTechnique.h:15-66
class RenderTechnique {
public:
    RenderTechnique(Device* device, PushConstants* pushConstants) : m_device(device), m_pushConstants(pushConstants) {}
    virtual ~RenderTechnique() = default;

    virtual void UpdateDescriptorSets() = 0;     // RenderTechnique.cpp:14-18: flush the queued writes
    virtual void SetFrameReferences(std::vector<FrameImage>& frameImages) = 0;
    virtual void ClearFrameReferences() = 0;
    virtual uint32_t GetRequiredSetCount() const = 0;

    virtual void QueueUpdateCloudData(DescriptorBufferInfo& cloudBufferInfo, unsigned int imageIdx) = 0;
    virtual void QueueUpdateCloudDataSampler(DescriptorImageInfo& cloudImageInfo, unsigned int imageIdx) = 0;
    virtual void QueueUpdateCameraProperties(DescriptorBufferInfo& cameraBufferInfo, unsigned int imageIdx) = 0;
    virtual void QueueUpdateParameters(DescriptorBufferInfo& parametersBufferInfo, unsigned int imageIdx) = 0;
    virtual void QueueUpdateShadowVolume(DescriptorBufferInfo& shadowVolumeBufferInfo, unsigned int imageIdx) = 0;
    virtual void QueueUpdateShadowVolumeSampler(DescriptorImageInfo& shadowVolumeImageInfo, unsigned int imageIdx) = 0;

    virtual void RecordDrawCommands(CommandStream commandBuffer, unsigned int imageIndex) = 0;

protected:
    template <typename T>
    static const T* As(const DescriptorBufferInfo& info, const char* what)
    {
        if (info.buffer == nullptr || info.range < sizeof(T)) throw Error(CRV_ERR_INVALID, std::string(what) + ": descriptor does not cover the struct");
        return static_cast<const T*>(info.buffer);
    }
    Device* m_device = nullptr;
    PushConstants* m_pushConstants = nullptr;
};

// RenderTechniqueSV.h:7-39
class RenderTechniqueSV : public RenderTechnique {
public:
    RenderTechniqueSV(Device* device, const ShadowVolumeProperties* shadowVolumeProperties, PushConstants* pushConstants)
        : RenderTechnique(device, pushConstants)
    {
        Check(crv_sv_create(device->Get(), shadowVolumeProperties, pushConstants, &m_sv));
    }
    ~RenderTechniqueSV() override { crv_sv_destroy(m_sv); }

    void UpdateDescriptorSets() override { Check(crv_sv_update_descriptor_sets(m_sv)); }
    void SetFrameReferences(std::vector<FrameImage>& frameImages) override
    {
        if (frameImages.empty() || frameImages[0].image3D == nullptr) throw Error(CRV_ERR_INVALID, "RenderTechniqueSV::SetFrameReferences: a 3D image is required");
        Check(crv_sv_set_frame_references(m_sv, frameImages[0].image3D->Get()));
    }
    void ClearFrameReferences() override { Check(crv_sv_clear_frame_references(m_sv)); }
    uint32_t GetRequiredSetCount() const override { return 1; }      // RenderTechniqueSV.cpp:97-100

    void QueueUpdateCloudData(DescriptorBufferInfo& info, unsigned int) override
    {
        Check(crv_sv_queue_update_cloud_data(m_sv, As<crv_cloud_properties>(info, "QueueUpdateCloudData")));
    }
    void QueueUpdateCloudDataSampler(DescriptorImageInfo& info, unsigned int) override
    {
        Check(crv_sv_queue_update_cloud_data_sampler(m_sv, info.image ? info.image->Get() : nullptr));
    }
    void QueueUpdateCameraProperties(DescriptorBufferInfo&, unsigned int) override {}      // no-op, RenderTechniqueSV.cpp:79-81
    void QueueUpdateParameters(DescriptorBufferInfo&, unsigned int) override {}            // no-op, RenderTechniqueSV.cpp:83-85
    void QueueUpdateShadowVolume(DescriptorBufferInfo& info, unsigned int) override
    {
        Check(crv_sv_queue_update_shadow_volume(m_sv, As<crv_shadow_volume_properties>(info, "QueueUpdateShadowVolume")));
    }
    void QueueUpdateShadowVolumeSampler(DescriptorImageInfo& info, unsigned int) override
    {
        Check(crv_sv_queue_update_shadow_volume_sampler(m_sv, info.image ? info.image->Get() : nullptr));
    }
    void RecordDrawCommands(CommandStream commandBuffer, unsigned int imageIndex) override
    {
        Check(crv_sv_record_draw_commands(m_sv, commandBuffer, imageIndex));
    }

    // extensions
    float LastBuildMs()
    {
        float ms = 0.0f;
        Check(crv_sv_last_build_ms(m_sv, &ms));
        return ms;
    }
    void SetTuning(crv_sv_mode mode) { Check(crv_sv_set_tuning(m_sv, mode)); }

private:
    crv_sv* m_sv = nullptr;
};

// RenderTechniquePT.h:8-40
class RenderTechniquePT : public RenderTechnique {
public:
    RenderTechniquePT(Device* device, const CameraProperties* cameraProperties, PushConstants* pushConstants)
        : RenderTechnique(device, pushConstants)
    {
        Check(crv_pt_create(device->Get(), cameraProperties, pushConstants, &m_pt));
    }
    ~RenderTechniquePT() override { crv_pt_destroy(m_pt); }

    void UpdateDescriptorSets() override { Check(crv_pt_update_descriptor_sets(m_pt)); }
    void SetFrameReferences(std::vector<FrameImage>& frameImages) override
    {
        std::vector<crv_image*> raw;
        for (const FrameImage& f : frameImages) {
            if (f.image2D == nullptr) throw Error(CRV_ERR_INVALID, "RenderTechniquePT::SetFrameReferences: 2D result images are required");
            raw.push_back(f.image2D->Get());
        }
        m_setCount = static_cast<uint32_t>(raw.size());
        Check(crv_pt_set_frame_references(m_pt, raw.data(), static_cast<unsigned>(raw.size())));
    }
    void ClearFrameReferences() override { Check(crv_pt_clear_frame_references(m_pt)); }
    uint32_t GetRequiredSetCount() const override { return m_setCount; }      // one set per swapchain image, RenderTechniquePT.cpp:124-127

    void QueueUpdateCloudData(DescriptorBufferInfo& info, unsigned int imageIdx) override
    {
        Check(crv_pt_queue_update_cloud_data(m_pt, As<crv_cloud_properties>(info, "QueueUpdateCloudData"), imageIdx));
    }
    void QueueUpdateCloudDataSampler(DescriptorImageInfo& info, unsigned int imageIdx) override
    {
        Check(crv_pt_queue_update_cloud_data_sampler(m_pt, info.image ? info.image->Get() : nullptr, imageIdx));
    }
    void QueueUpdateCameraProperties(DescriptorBufferInfo& info, unsigned int imageIdx) override
    {
        Check(crv_pt_queue_update_camera_properties(m_pt, As<crv_camera_properties>(info, "QueueUpdateCameraProperties"), imageIdx));
    }
    void QueueUpdateParameters(DescriptorBufferInfo& info, unsigned int imageIdx) override
    {
        Check(crv_pt_queue_update_parameters(m_pt, As<crv_parameters>(info, "QueueUpdateParameters"), imageIdx));
    }
    void QueueUpdateShadowVolume(DescriptorBufferInfo& info, unsigned int imageIdx) override
    {
        Check(crv_pt_queue_update_shadow_volume(m_pt, As<crv_shadow_volume_properties>(info, "QueueUpdateShadowVolume"), imageIdx));
    }
    void QueueUpdateShadowVolumeSampler(DescriptorImageInfo& info, unsigned int imageIdx) override
    {
        Check(crv_pt_queue_update_shadow_volume_sampler(m_pt, info.image ? info.image->Get() : nullptr, imageIdx));
    }
    // one progressive frame: seed / frameCount are read from the PushConstants now (vkCmdPushConstants, .cpp:132)
    void RecordDrawCommands(CommandStream commandBuffer, unsigned int imageIndex) override
    {
        Check(crv_pt_record_draw_commands(m_pt, commandBuffer, imageIndex));
    }

    // extensions beyond the reference interface
    void SetExtensions(const PathTracerExtensions& ext) { Check(crv_pt_set_extensions(m_pt, &ext)); }
    // n frames in one call, bit-identical to n RecordDrawCommands with seed = seeds[i], frameCount = first + i
    void RenderBatch(const int32_t* seeds, unsigned n, uint32_t firstFrameCount, crv_accum_mode mode, CommandStream stream,
                     unsigned int imageIndex)
    {
        Check(crv_pt_render_batch(m_pt, seeds, n, firstFrameCount, mode, stream, imageIndex));
    }
    // the same as a plain sum for a member of a DeviceGroup: the fold sends the tracked pixels' sums to every member's inbox
    // (targets = DeviceGroup::PushBegin()) and finishes the other pixels; the image needs no clear
    void RenderBatchPush(const int32_t* seeds, unsigned n, const crv_push_targets& targets, float othersScale, CommandStream stream,
                         unsigned int imageIndex)
    {
        Check(crv_pt_render_batch_push(m_pt, seeds, n, &targets, othersScale, stream, imageIndex));
    }
    crv_pt_stats GetStats(CommandStream stream = nullptr)
    {
        crv_pt_stats s{};
        Check(crv_pt_get_stats(m_pt, &s, stream));
        return s;
    }
    void ResetStats(CommandStream stream = nullptr) { Check(crv_pt_reset_stats(m_pt, stream)); }
    // 0 = pool scheduler (default), 1 = lane-owned scheduler; identical results
    void SetScheduler(int scheduler) { Check(crv_pt_set_scheduler(m_pt, scheduler)); }
    // 1 (default): ray sections that provably cannot collide are not tracked; 0: every step of the reference is executed
    void SetSkipLevel(int level) { Check(crv_pt_set_skip_level(m_pt, level)); }
    // which pixels of the last render call were tracked / are the same background sample in every frame (device-side lists)
    crv_pixel_lists GetPixelLists()
    {
        crv_pixel_lists l{};
        Check(crv_pt_pixel_lists(m_pt, &l));
        return l;
    }
    // minimum cell size (8 << shift voxels) of the tracker's coarse threshold table; -1 = automatic
    void SetCoarseShift(int minShift) { Check(crv_pt_set_coarse_shift(m_pt, minShift)); }
    // job order of a launch: 0 frame-major, 1 pixel-major over the hit list sorted by expected work, -1 automatic; identical results
    void SetJobOrder(int order) { Check(crv_pt_set_job_order(m_pt, order)); }
    float LastRenderMs(unsigned* launches = nullptr)
    {
        float ms = 0.0f;
        Check(crv_pt_last_render_ms(m_pt, &ms, launches));
        return ms;
    }

private:
    crv_pt* m_pt = nullptr;
    uint32_t m_setCount = 1;
};

// RenderTechniquePPM.h:8-70 / .cpp: progressive photon mapping -- per frame the radius update, the
// photon-tracing dispatch (shaders/PPM_PT.comp) and the photon-estimate dispatch (shaders/PPM_PE.comp)
class RenderTechniquePPM : public RenderTechnique {
public:
    RenderTechniquePPM(Device* device, const CameraProperties* /*cameraProperties: read from its descriptor*/, PhotonMapProperties* photonMapProperties,
                       PushConstants* pushConstants, float initialRadius)
        : RenderTechnique(device, pushConstants)
    {
        Check(crv_ppm_create(device->Get(), photonMapProperties, pushConstants, initialRadius, &m_ppm));
    }
    ~RenderTechniquePPM() override { crv_ppm_destroy(m_ppm); }

    void AllocateResources() { Check(crv_ppm_allocate_resources(m_ppm)); }          // RenderTechniquePPM.cpp:127-160
    void FreeResources() { Check(crv_ppm_free_resources(m_ppm)); }                  // :113-125
    void SetFrameReferences(std::vector<FrameImage>& frameImages) override
    {
        std::vector<crv_image*> imgs;
        for (auto& f : frameImages) imgs.push_back(f.image2D ? f.image2D->Get() : nullptr);
        Check(crv_ppm_set_frame_references(m_ppm, imgs.data(), (unsigned)imgs.size()));
    }
    void ClearFrameReferences() override { Check(crv_ppm_set_frame_references(m_ppm, nullptr, 0)); }
    uint32_t GetRequiredSetCount() const override { return 2; }                     // tracing and estimate sets (:191-194)
    void QueueUpdateCloudData(DescriptorBufferInfo& info, unsigned int) override
    {
        Check(crv_ppm_queue_update_cloud_data(m_ppm, static_cast<const crv_cloud_properties*>(info.buffer)));
    }
    void QueueUpdateCloudDataSampler(DescriptorImageInfo& info, unsigned int) override
    {
        Check(crv_ppm_queue_update_cloud_data_sampler(m_ppm, info.image ? info.image->Get() : nullptr));
    }
    void QueueUpdateCameraProperties(DescriptorBufferInfo& info, unsigned int) override
    {
        Check(crv_ppm_queue_update_camera_properties(m_ppm, static_cast<const crv_camera_properties*>(info.buffer)));
    }
    void QueueUpdateParameters(DescriptorBufferInfo& info, unsigned int) override
    {
        Check(crv_ppm_queue_update_parameters(m_ppm, static_cast<const crv_parameters*>(info.buffer)));
    }
    void QueueUpdateShadowVolume(DescriptorBufferInfo& info, unsigned int) override
    {
        Check(crv_ppm_queue_update_shadow_volume(m_ppm, static_cast<const crv_shadow_volume_properties*>(info.buffer)));
    }
    void QueueUpdateShadowVolumeSampler(DescriptorImageInfo& info, unsigned int) override
    {
        Check(crv_ppm_queue_update_shadow_volume_sampler(m_ppm, info.image ? info.image->Get() : nullptr));
    }
    void UpdateDescriptorSets() override { Check(crv_ppm_update_descriptor_sets(m_ppm)); }
    // one progressive frame: UpdateRadius(frameCount), photon tracing, photon estimate (:229-273)
    void RecordDrawCommands(CommandStream commandBuffer, unsigned int imageIndex) override
    {
        Check(crv_ppm_record_draw_commands(m_ppm, commandBuffer, imageIndex));
    }
    void RecordPhotonTracing(CommandStream commandBuffer, uint32_t groupsX = 0) { Check(crv_ppm_record_photon_tracing(m_ppm, commandBuffer, groupsX)); }
    void RecordPhotonEstimate(CommandStream commandBuffer, unsigned int imageIndex) { Check(crv_ppm_record_photon_estimate(m_ppm, commandBuffer, imageIndex)); }
    void Download(crv_photon* photons, uint32_t* collisions, CommandStream stream = nullptr) { Check(crv_ppm_download(m_ppm, photons, collisions, stream)); }
    crv_ppm* Get() const { return m_ppm; }

private:
    crv_ppm* m_ppm = nullptr;
};

}  // namespace crv
