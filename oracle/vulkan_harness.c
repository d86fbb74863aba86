/* oracle/vulkan_harness.c -- TEST INFRASTRUCTURE: the oracle north_star names.
 *
 * Dispatches the reference's OWN compute shaders (shaders/ShadowVolume.comp, shaders/PathTracer.comp, compiled to SPIR-V by
 * glslc exactly as shaders/Compile.bat does) from a minimal headless Vulkan program: instance without surface, one compute
 * queue, the descriptor set layouts of RenderTechniqueSV.cpp:13-22 (4 bindings) and RenderTechniquePT.cpp:20-35 (7 bindings),
 * the 24-byte push-constant range of RenderTechniquePT.cpp:39-42, the images and sampler of main.cpp:277-285,405-414,802-804 /
 * Initializers.cpp:408-418, the dispatch shapes of RenderTechniqueSV.cpp:110 and RenderTechniquePT.cpp:141, and reads the
 * shadow volume and the result image back.  On Mesa lavapipe this is the CPU oracle + baseline of north_star; on a GPU's
 * Vulkan driver it is the reference itself.
 *
 * The build image has no Vulkan SDK: no vulkan.h, no loader, no ICD, no glslc (oracle/vulkan_probe.py reports which are
 * missing).  So this file declares the few dozen Vulkan 1.0 types and prototypes it uses by hand, binds the loader with
 * dlopen("libvulkan.so.1") at run time, and exits with code 77 ("skipped") where anything is missing.  It compiles and links
 * everywhere (gcc, -ldl only) but could not be EXECUTED in this round's images: it carries no parity claim.
 *
 *   vulkan_harness <ShadowVolume.comp.spv> <PathTracer.comp.spv> <scene.bin> <out.bin>
 * scene.bin (written by oracle/vulkan_harness.py):
 *   u32 magic 'CRVS', u32 nx, ny, nz, u32 width, height, u32 n_sv, u32 n_frames,
 *   camera 68 B (padded to 80), cloud 60 B (padded to 64), parameters 28 B (padded to 32), shadow-volume properties 152 B (padded to 160),
 *   n_frames x { i32 seed, u32 frameCount }, nx*ny*nz floats (x fastest).
 * out.bin: n_sv^3 floats (shadow volume, x fastest), width*height*4 floats (result image after the last frame).
 */
#include <dlfcn.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ---- hand-declared subset of vulkan_core.h (Vulkan 1.0 ABI, 64-bit) --------------------------------------------------- */
typedef uint32_t VkFlags;
typedef uint32_t VkBool32;
typedef uint64_t VkDeviceSize;
typedef int32_t VkResult;
#define VK_HANDLE(n) typedef struct n##_T* n
VK_HANDLE(VkInstance); VK_HANDLE(VkPhysicalDevice); VK_HANDLE(VkDevice); VK_HANDLE(VkQueue); VK_HANDLE(VkCommandPool);
VK_HANDLE(VkCommandBuffer); VK_HANDLE(VkBuffer); VK_HANDLE(VkImage); VK_HANDLE(VkDeviceMemory); VK_HANDLE(VkImageView);
VK_HANDLE(VkSampler); VK_HANDLE(VkDescriptorSetLayout); VK_HANDLE(VkPipelineLayout); VK_HANDLE(VkShaderModule);
VK_HANDLE(VkPipeline); VK_HANDLE(VkDescriptorPool); VK_HANDLE(VkDescriptorSet); VK_HANDLE(VkPipelineCache); VK_HANDLE(VkFence);
typedef void (*PFN_vkVoidFunction)(void);

enum { ST_APPLICATION_INFO = 0, ST_INSTANCE_CREATE_INFO = 1, ST_DEVICE_QUEUE_CREATE_INFO = 2, ST_DEVICE_CREATE_INFO = 3, ST_SUBMIT_INFO = 4,
       ST_MEMORY_ALLOCATE_INFO = 5, ST_BUFFER_CREATE_INFO = 12, ST_IMAGE_CREATE_INFO = 14, ST_IMAGE_VIEW_CREATE_INFO = 15,
       ST_SHADER_MODULE_CREATE_INFO = 16, ST_PIPELINE_SHADER_STAGE_CREATE_INFO = 18, ST_COMPUTE_PIPELINE_CREATE_INFO = 29,
       ST_PIPELINE_LAYOUT_CREATE_INFO = 30, ST_SAMPLER_CREATE_INFO = 31, ST_DESCRIPTOR_SET_LAYOUT_CREATE_INFO = 32,
       ST_DESCRIPTOR_POOL_CREATE_INFO = 33, ST_DESCRIPTOR_SET_ALLOCATE_INFO = 34, ST_WRITE_DESCRIPTOR_SET = 35,
       ST_COMMAND_POOL_CREATE_INFO = 39, ST_COMMAND_BUFFER_ALLOCATE_INFO = 40, ST_COMMAND_BUFFER_BEGIN_INFO = 42,
       ST_IMAGE_MEMORY_BARRIER = 45, ST_MEMORY_BARRIER = 46 };
enum { FORMAT_R32_SFLOAT = 100, FORMAT_R32G32B32A32_SFLOAT = 109 };
enum { IMAGE_TYPE_2D = 1, IMAGE_TYPE_3D = 2, VIEW_TYPE_2D = 1, VIEW_TYPE_3D = 2 };
enum { LAYOUT_UNDEFINED = 0, LAYOUT_GENERAL = 1, LAYOUT_SHADER_READ_ONLY = 5, LAYOUT_TRANSFER_SRC = 6, LAYOUT_TRANSFER_DST = 7 };
enum { IMG_USAGE_TRANSFER_SRC = 1, IMG_USAGE_TRANSFER_DST = 2, IMG_USAGE_SAMPLED = 4, IMG_USAGE_STORAGE = 8 };
enum { BUF_USAGE_TRANSFER_SRC = 1, BUF_USAGE_TRANSFER_DST = 2, BUF_USAGE_UNIFORM = 0x10 };
enum { MEM_DEVICE_LOCAL = 1, MEM_HOST_VISIBLE = 2, MEM_HOST_COHERENT = 4 };
enum { DESC_COMBINED_IMAGE_SAMPLER = 1, DESC_STORAGE_IMAGE = 3, DESC_UNIFORM_BUFFER = 6 };
enum { STAGE_COMPUTE_BIT = 0x20, BIND_POINT_COMPUTE = 1, QUEUE_COMPUTE_BIT = 2, ASPECT_COLOR = 1 };
enum { FILTER_LINEAR = 1, ADDRESS_CLAMP_TO_EDGE = 2 };
enum { PSTAGE_TOP = 1, PSTAGE_COMPUTE = 0x800, PSTAGE_TRANSFER = 0x1000, PSTAGE_HOST = 0x4000 };
enum { ACCESS_SHADER_READ = 0x20, ACCESS_SHADER_WRITE = 0x40, ACCESS_TRANSFER_READ = 0x800, ACCESS_TRANSFER_WRITE = 0x1000, ACCESS_HOST_READ = 0x2000 };
#define QUEUE_FAMILY_IGNORED (~0u)
#define WHOLE_SIZE (~0ull)

typedef struct { uint32_t sType; const void* pNext; const char* pApplicationName; uint32_t applicationVersion; const char* pEngineName;
                 uint32_t engineVersion; uint32_t apiVersion; } VkApplicationInfo;
typedef struct { uint32_t sType; const void* pNext; VkFlags flags; const VkApplicationInfo* pApplicationInfo; uint32_t enabledLayerCount;
                 const char* const* ppEnabledLayerNames; uint32_t enabledExtensionCount; const char* const* ppEnabledExtensionNames; } VkInstanceCreateInfo;
typedef struct { uint32_t sType; const void* pNext; VkFlags flags; uint32_t queueFamilyIndex; uint32_t queueCount; const float* pQueuePriorities; } VkDeviceQueueCreateInfo;
typedef struct { uint32_t sType; const void* pNext; VkFlags flags; uint32_t queueCreateInfoCount; const VkDeviceQueueCreateInfo* pQueueCreateInfos;
                 uint32_t enabledLayerCount; const char* const* ppEnabledLayerNames; uint32_t enabledExtensionCount;
                 const char* const* ppEnabledExtensionNames; const void* pEnabledFeatures; } VkDeviceCreateInfo;
typedef struct { VkFlags queueFlags; uint32_t queueCount; uint32_t timestampValidBits; uint32_t granularity[3]; } VkQueueFamilyProperties;
typedef struct { VkFlags propertyFlags; uint32_t heapIndex; } VkMemoryType;
typedef struct { VkDeviceSize size; VkFlags flags; } VkMemoryHeap;
typedef struct { uint32_t memoryTypeCount; VkMemoryType memoryTypes[32]; uint32_t memoryHeapCount; VkMemoryHeap memoryHeaps[16]; } VkPhysicalDeviceMemoryProperties;
typedef struct { uint32_t sType; const void* pNext; VkFlags flags; uint32_t queueFamilyIndex; } VkCommandPoolCreateInfo;
typedef struct { uint32_t sType; const void* pNext; VkCommandPool commandPool; uint32_t level; uint32_t commandBufferCount; } VkCommandBufferAllocateInfo;
typedef struct { uint32_t sType; const void* pNext; VkFlags flags; VkDeviceSize size; VkFlags usage; uint32_t sharingMode; uint32_t queueFamilyIndexCount;
                 const uint32_t* pQueueFamilyIndices; } VkBufferCreateInfo;
typedef struct { VkDeviceSize size; VkDeviceSize alignment; uint32_t memoryTypeBits; } VkMemoryRequirements;
typedef struct { uint32_t sType; const void* pNext; VkDeviceSize allocationSize; uint32_t memoryTypeIndex; } VkMemoryAllocateInfo;
typedef struct { uint32_t width, height, depth; } VkExtent3D;
typedef struct { int32_t x, y, z; } VkOffset3D;
typedef struct { uint32_t sType; const void* pNext; VkFlags flags; uint32_t imageType; uint32_t format; VkExtent3D extent; uint32_t mipLevels; uint32_t arrayLayers;
                 uint32_t samples; uint32_t tiling; VkFlags usage; uint32_t sharingMode; uint32_t queueFamilyIndexCount; const uint32_t* pQueueFamilyIndices;
                 uint32_t initialLayout; } VkImageCreateInfo;
typedef struct { VkFlags aspectMask; uint32_t baseMipLevel, levelCount, baseArrayLayer, layerCount; } VkImageSubresourceRange;
typedef struct { uint32_t sType; const void* pNext; VkFlags flags; VkImage image; uint32_t viewType; uint32_t format; uint32_t components[4];
                 VkImageSubresourceRange subresourceRange; } VkImageViewCreateInfo;
typedef struct { uint32_t sType; const void* pNext; VkFlags flags; uint32_t magFilter, minFilter, mipmapMode, addressModeU, addressModeV, addressModeW; float mipLodBias;
                 VkBool32 anisotropyEnable; float maxAnisotropy; VkBool32 compareEnable; uint32_t compareOp; float minLod, maxLod; uint32_t borderColor;
                 VkBool32 unnormalizedCoordinates; } VkSamplerCreateInfo;
typedef struct { uint32_t binding; uint32_t descriptorType; uint32_t descriptorCount; VkFlags stageFlags; const VkSampler* pImmutableSamplers; } VkDescriptorSetLayoutBinding;
typedef struct { uint32_t sType; const void* pNext; VkFlags flags; uint32_t bindingCount; const VkDescriptorSetLayoutBinding* pBindings; } VkDescriptorSetLayoutCreateInfo;
typedef struct { VkFlags stageFlags; uint32_t offset; uint32_t size; } VkPushConstantRange;
typedef struct { uint32_t sType; const void* pNext; VkFlags flags; uint32_t setLayoutCount; const VkDescriptorSetLayout* pSetLayouts; uint32_t pushConstantRangeCount;
                 const VkPushConstantRange* pPushConstantRanges; } VkPipelineLayoutCreateInfo;
typedef struct { uint32_t sType; const void* pNext; VkFlags flags; size_t codeSize; const uint32_t* pCode; } VkShaderModuleCreateInfo;
typedef struct { uint32_t sType; const void* pNext; VkFlags flags; uint32_t stage; VkShaderModule module; const char* pName; const void* pSpecializationInfo; } VkPipelineShaderStageCreateInfo;
typedef struct { uint32_t sType; const void* pNext; VkFlags flags; VkPipelineShaderStageCreateInfo stage; VkPipelineLayout layout; VkPipeline basePipelineHandle;
                 int32_t basePipelineIndex; } VkComputePipelineCreateInfo;
typedef struct { uint32_t type; uint32_t descriptorCount; } VkDescriptorPoolSize;
typedef struct { uint32_t sType; const void* pNext; VkFlags flags; uint32_t maxSets; uint32_t poolSizeCount; const VkDescriptorPoolSize* pPoolSizes; } VkDescriptorPoolCreateInfo;
typedef struct { uint32_t sType; const void* pNext; VkDescriptorPool descriptorPool; uint32_t descriptorSetCount; const VkDescriptorSetLayout* pSetLayouts; } VkDescriptorSetAllocateInfo;
typedef struct { VkSampler sampler; VkImageView imageView; uint32_t imageLayout; } VkDescriptorImageInfo;
typedef struct { VkBuffer buffer; VkDeviceSize offset; VkDeviceSize range; } VkDescriptorBufferInfo;
typedef struct { uint32_t sType; const void* pNext; VkDescriptorSet dstSet; uint32_t dstBinding; uint32_t dstArrayElement; uint32_t descriptorCount; uint32_t descriptorType;
                 const VkDescriptorImageInfo* pImageInfo; const VkDescriptorBufferInfo* pBufferInfo; const void* pTexelBufferView; } VkWriteDescriptorSet;
typedef struct { uint32_t sType; const void* pNext; VkFlags flags; const void* pInheritanceInfo; } VkCommandBufferBeginInfo;
typedef struct { uint32_t sType; const void* pNext; VkFlags srcAccessMask, dstAccessMask; uint32_t oldLayout, newLayout; uint32_t srcQueueFamilyIndex, dstQueueFamilyIndex;
                 VkImage image; VkImageSubresourceRange subresourceRange; } VkImageMemoryBarrier;
typedef struct { uint32_t sType; const void* pNext; VkFlags srcAccessMask, dstAccessMask; } VkMemoryBarrier;
typedef struct { VkFlags aspectMask; uint32_t mipLevel, baseArrayLayer, layerCount; } VkImageSubresourceLayers;
typedef struct { VkDeviceSize bufferOffset; uint32_t bufferRowLength, bufferImageHeight; VkImageSubresourceLayers imageSubresource; VkOffset3D imageOffset;
                 VkExtent3D imageExtent; } VkBufferImageCopy;
typedef struct { uint32_t sType; const void* pNext; uint32_t waitSemaphoreCount; const void* pWaitSemaphores; const VkFlags* pWaitDstStageMask; uint32_t commandBufferCount;
                 const VkCommandBuffer* pCommandBuffers; uint32_t signalSemaphoreCount; const void* pSignalSemaphores; } VkSubmitInfo;
typedef union { float float32[4]; int32_t int32[4]; uint32_t uint32[4]; } VkClearColorValue;

#define SKIP 77
#define DIE(...) do { fprintf(stderr, "vulkan_harness: " __VA_ARGS__); fprintf(stderr, "\n"); exit(1); } while (0)
#define VK(call) do { VkResult r_ = (call); if (r_ < 0) DIE("%s failed: VkResult %d", #call, (int)r_); } while (0)

static PFN_vkVoidFunction (*p_vkGetInstanceProcAddr)(VkInstance, const char*);
static VkInstance g_instance;
#define LOAD(name) name = (__typeof__(name))p_vkGetInstanceProcAddr(g_instance, #name); if (!name) DIE("the loader has no %s", #name)

static VkResult (*vkCreateInstance)(const VkInstanceCreateInfo*, const void*, VkInstance*);
static VkResult (*vkEnumeratePhysicalDevices)(VkInstance, uint32_t*, VkPhysicalDevice*);
static void (*vkGetPhysicalDeviceProperties)(VkPhysicalDevice, void*);
static void (*vkGetPhysicalDeviceFeatures)(VkPhysicalDevice, void*);
static void (*vkGetPhysicalDeviceQueueFamilyProperties)(VkPhysicalDevice, uint32_t*, VkQueueFamilyProperties*);
static void (*vkGetPhysicalDeviceMemoryProperties)(VkPhysicalDevice, VkPhysicalDeviceMemoryProperties*);
static VkResult (*vkCreateDevice)(VkPhysicalDevice, const VkDeviceCreateInfo*, const void*, VkDevice*);
static void (*vkGetDeviceQueue)(VkDevice, uint32_t, uint32_t, VkQueue*);
static VkResult (*vkCreateCommandPool)(VkDevice, const VkCommandPoolCreateInfo*, const void*, VkCommandPool*);
static VkResult (*vkAllocateCommandBuffers)(VkDevice, const VkCommandBufferAllocateInfo*, VkCommandBuffer*);
static VkResult (*vkCreateBuffer)(VkDevice, const VkBufferCreateInfo*, const void*, VkBuffer*);
static void (*vkGetBufferMemoryRequirements)(VkDevice, VkBuffer, VkMemoryRequirements*);
static VkResult (*vkAllocateMemory)(VkDevice, const VkMemoryAllocateInfo*, const void*, VkDeviceMemory*);
static VkResult (*vkBindBufferMemory)(VkDevice, VkBuffer, VkDeviceMemory, VkDeviceSize);
static VkResult (*vkMapMemory)(VkDevice, VkDeviceMemory, VkDeviceSize, VkDeviceSize, VkFlags, void**);
static VkResult (*vkCreateImage)(VkDevice, const VkImageCreateInfo*, const void*, VkImage*);
static void (*vkGetImageMemoryRequirements)(VkDevice, VkImage, VkMemoryRequirements*);
static VkResult (*vkBindImageMemory)(VkDevice, VkImage, VkDeviceMemory, VkDeviceSize);
static VkResult (*vkCreateImageView)(VkDevice, const VkImageViewCreateInfo*, const void*, VkImageView*);
static VkResult (*vkCreateSampler)(VkDevice, const VkSamplerCreateInfo*, const void*, VkSampler*);
static VkResult (*vkCreateDescriptorSetLayout)(VkDevice, const VkDescriptorSetLayoutCreateInfo*, const void*, VkDescriptorSetLayout*);
static VkResult (*vkCreatePipelineLayout)(VkDevice, const VkPipelineLayoutCreateInfo*, const void*, VkPipelineLayout*);
static VkResult (*vkCreateShaderModule)(VkDevice, const VkShaderModuleCreateInfo*, const void*, VkShaderModule*);
static VkResult (*vkCreateComputePipelines)(VkDevice, VkPipelineCache, uint32_t, const VkComputePipelineCreateInfo*, const void*, VkPipeline*);
static VkResult (*vkCreateDescriptorPool)(VkDevice, const VkDescriptorPoolCreateInfo*, const void*, VkDescriptorPool*);
static VkResult (*vkAllocateDescriptorSets)(VkDevice, const VkDescriptorSetAllocateInfo*, VkDescriptorSet*);
static void (*vkUpdateDescriptorSets)(VkDevice, uint32_t, const VkWriteDescriptorSet*, uint32_t, const void*);
static VkResult (*vkBeginCommandBuffer)(VkCommandBuffer, const VkCommandBufferBeginInfo*);
static VkResult (*vkEndCommandBuffer)(VkCommandBuffer);
static VkResult (*vkResetCommandBuffer)(VkCommandBuffer, VkFlags);
static void (*vkCmdPipelineBarrier)(VkCommandBuffer, VkFlags, VkFlags, VkFlags, uint32_t, const VkMemoryBarrier*, uint32_t, const void*, uint32_t, const VkImageMemoryBarrier*);
static void (*vkCmdCopyBufferToImage)(VkCommandBuffer, VkBuffer, VkImage, uint32_t, uint32_t, const VkBufferImageCopy*);
static void (*vkCmdCopyImageToBuffer)(VkCommandBuffer, VkImage, uint32_t, VkBuffer, uint32_t, const VkBufferImageCopy*);
static void (*vkCmdClearColorImage)(VkCommandBuffer, VkImage, uint32_t, const VkClearColorValue*, uint32_t, const VkImageSubresourceRange*);
static void (*vkCmdBindPipeline)(VkCommandBuffer, uint32_t, VkPipeline);
static void (*vkCmdBindDescriptorSets)(VkCommandBuffer, uint32_t, VkPipelineLayout, uint32_t, uint32_t, const VkDescriptorSet*, uint32_t, const uint32_t*);
static void (*vkCmdPushConstants)(VkCommandBuffer, VkPipelineLayout, VkFlags, uint32_t, uint32_t, const void*);
static void (*vkCmdDispatch)(VkCommandBuffer, uint32_t, uint32_t, uint32_t);
static VkResult (*vkQueueSubmit)(VkQueue, uint32_t, const VkSubmitInfo*, VkFence);
static VkResult (*vkQueueWaitIdle)(VkQueue);

static VkDevice dev;
static VkPhysicalDeviceMemoryProperties memprops;

static uint32_t memory_type(uint32_t bits, VkFlags want)
{
    for (uint32_t i = 0; i < memprops.memoryTypeCount; ++i)
        if ((bits & (1u << i)) && (memprops.memoryTypes[i].propertyFlags & want) == want) return i;
    DIE("no memory type with properties 0x%x", want);
    return 0;
}

typedef struct { VkBuffer buf; VkDeviceMemory mem; void* map; VkDeviceSize size; } HostBuffer;
static HostBuffer host_buffer(VkDeviceSize size, VkFlags usage)
{
    HostBuffer b = {0};
    VkBufferCreateInfo ci = {ST_BUFFER_CREATE_INFO, NULL, 0, size, usage, 0, 0, NULL};
    VK(vkCreateBuffer(dev, &ci, NULL, &b.buf));
    VkMemoryRequirements req;
    vkGetBufferMemoryRequirements(dev, b.buf, &req);
    VkMemoryAllocateInfo ai = {ST_MEMORY_ALLOCATE_INFO, NULL, req.size, memory_type(req.memoryTypeBits, MEM_HOST_VISIBLE | MEM_HOST_COHERENT)};
    VK(vkAllocateMemory(dev, &ai, NULL, &b.mem));
    VK(vkBindBufferMemory(dev, b.buf, b.mem, 0));
    VK(vkMapMemory(dev, b.mem, 0, WHOLE_SIZE, 0, &b.map));
    b.size = size;
    return b;
}

typedef struct { VkImage img; VkDeviceMemory mem; VkImageView view; } DeviceImage;
static DeviceImage device_image(uint32_t type, uint32_t format, uint32_t w, uint32_t h, uint32_t d, VkFlags usage)
{
    DeviceImage im = {0};
    VkImageCreateInfo ci = {ST_IMAGE_CREATE_INFO, NULL, 0, type, format, {w, h, d}, 1, 1, /* VK_SAMPLE_COUNT_1_BIT */ 1, /* OPTIMAL */ 0, usage, 0, 0, NULL, LAYOUT_UNDEFINED};
    VK(vkCreateImage(dev, &ci, NULL, &im.img));
    VkMemoryRequirements req;
    vkGetImageMemoryRequirements(dev, im.img, &req);
    VkMemoryAllocateInfo ai = {ST_MEMORY_ALLOCATE_INFO, NULL, req.size, memory_type(req.memoryTypeBits, MEM_DEVICE_LOCAL)};
    VK(vkAllocateMemory(dev, &ai, NULL, &im.mem));
    VK(vkBindImageMemory(dev, im.img, im.mem, 0));
    VkImageViewCreateInfo vi = {ST_IMAGE_VIEW_CREATE_INFO, NULL, 0, im.img, type == IMAGE_TYPE_3D ? VIEW_TYPE_3D : VIEW_TYPE_2D, format, {0, 0, 0, 0}, {ASPECT_COLOR, 0, 1, 0, 1}};
    VK(vkCreateImageView(dev, &vi, NULL, &im.view));
    return im;
}

static void transition(VkCommandBuffer cb, VkImage img, uint32_t from, uint32_t to, VkFlags src_access, VkFlags dst_access, VkFlags src_stage, VkFlags dst_stage)
{
    VkImageMemoryBarrier b = {ST_IMAGE_MEMORY_BARRIER, NULL, src_access, dst_access, from, to, QUEUE_FAMILY_IGNORED, QUEUE_FAMILY_IGNORED, img, {ASPECT_COLOR, 0, 1, 0, 1}};
    vkCmdPipelineBarrier(cb, src_stage, dst_stage, 0, 0, NULL, 0, NULL, 1, &b);
}

static void* read_file(const char* path, size_t* n)
{
    FILE* f = fopen(path, "rb");
    if (!f) DIE("cannot open %s", path);
    fseek(f, 0, SEEK_END);
    *n = (size_t)ftell(f);
    fseek(f, 0, SEEK_SET);
    void* p = malloc(*n ? *n : 1);
    if (fread(p, 1, *n, f) != *n) DIE("short read on %s", path);
    fclose(f);
    return p;
}

static VkPipeline compute_pipeline(const char* spv_path, VkPipelineLayout layout)
{
    size_t n;
    uint32_t* code = (uint32_t*)read_file(spv_path, &n);
    if (n < 20 || code[0] != 0x07230203u) DIE("%s is not a SPIR-V module", spv_path);
    VkShaderModuleCreateInfo mi = {ST_SHADER_MODULE_CREATE_INFO, NULL, 0, n, code};
    VkShaderModule mod;
    VK(vkCreateShaderModule(dev, &mi, NULL, &mod));
    VkComputePipelineCreateInfo ci = {ST_COMPUTE_PIPELINE_CREATE_INFO, NULL, 0, {ST_PIPELINE_SHADER_STAGE_CREATE_INFO, NULL, 0, STAGE_COMPUTE_BIT, mod, "main", NULL}, layout, NULL, -1};
    VkPipeline p;
    VK(vkCreateComputePipelines(dev, NULL, 1, &ci, NULL, &p));
    free(code);
    return p;
}

int main(int argc, char** argv)
{
    if (argc != 5) { fprintf(stderr, "usage: vulkan_harness ShadowVolume.comp.spv PathTracer.comp.spv scene.bin out.bin\n"); return 2; }
    void* lib = dlopen("libvulkan.so.1", RTLD_NOW);
    if (!lib) lib = dlopen("libvulkan.so", RTLD_NOW);
    if (!lib) { fprintf(stderr, "vulkan_harness: skipped: no Vulkan loader (libvulkan.so.1)\n"); return SKIP; }
    p_vkGetInstanceProcAddr = (PFN_vkVoidFunction(*)(VkInstance, const char*))dlsym(lib, "vkGetInstanceProcAddr");
    if (!p_vkGetInstanceProcAddr) { fprintf(stderr, "vulkan_harness: skipped: loader without vkGetInstanceProcAddr\n"); return SKIP; }

    /* ---- scene ---- */
    size_t scene_n;
    unsigned char* scene = (unsigned char*)read_file(argv[3], &scene_n);
    const uint32_t* hdr = (const uint32_t*)scene;
    if (scene_n < 32 || hdr[0] != 0x53565243u) DIE("%s is not a scene file", argv[3]);
    const uint32_t nx = hdr[1], ny = hdr[2], nz = hdr[3], W = hdr[4], H = hdr[5], NSV = hdr[6], NF = hdr[7];
    const unsigned char* p_cam = scene + 32;
    const unsigned char* p_cloud = p_cam + 80;
    const unsigned char* p_par = p_cloud + 64;
    const unsigned char* p_sv = p_par + 32;
    const unsigned char* p_frames = p_sv + 160;
    const float* grid = (const float*)(p_frames + 8 * (size_t)NF);
    if ((const unsigned char*)(grid + (size_t)nx * ny * nz) > scene + scene_n) DIE("scene file truncated");

    /* ---- instance / device: VulkanInstance, VulkanPhysicalDevice, VulkanDevice (main.cpp:774-793), no surface, no swapchain ---- */
    vkCreateInstance = (__typeof__(vkCreateInstance))p_vkGetInstanceProcAddr(NULL, "vkCreateInstance");
    if (!vkCreateInstance) { fprintf(stderr, "vulkan_harness: skipped: no vkCreateInstance\n"); return SKIP; }
    VkApplicationInfo app = {ST_APPLICATION_INFO, NULL, "crv vulkan harness", 1, "none", 0, (1u << 22) | (2u << 12)};   /* apiVersion 1.2 (VulkanConfiguration.h:11) */
    VkInstanceCreateInfo ici = {ST_INSTANCE_CREATE_INFO, NULL, 0, &app, 0, NULL, 0, NULL};
    if (vkCreateInstance(&ici, NULL, &g_instance) < 0) { fprintf(stderr, "vulkan_harness: skipped: vkCreateInstance failed (no ICD?)\n"); return SKIP; }
    LOAD(vkEnumeratePhysicalDevices); LOAD(vkGetPhysicalDeviceProperties); LOAD(vkGetPhysicalDeviceFeatures); LOAD(vkGetPhysicalDeviceQueueFamilyProperties);
    LOAD(vkGetPhysicalDeviceMemoryProperties); LOAD(vkCreateDevice); LOAD(vkGetDeviceQueue); LOAD(vkCreateCommandPool); LOAD(vkAllocateCommandBuffers);
    LOAD(vkCreateBuffer); LOAD(vkGetBufferMemoryRequirements); LOAD(vkAllocateMemory); LOAD(vkBindBufferMemory); LOAD(vkMapMemory); LOAD(vkCreateImage);
    LOAD(vkGetImageMemoryRequirements); LOAD(vkBindImageMemory); LOAD(vkCreateImageView); LOAD(vkCreateSampler); LOAD(vkCreateDescriptorSetLayout);
    LOAD(vkCreatePipelineLayout); LOAD(vkCreateShaderModule); LOAD(vkCreateComputePipelines); LOAD(vkCreateDescriptorPool); LOAD(vkAllocateDescriptorSets);
    LOAD(vkUpdateDescriptorSets); LOAD(vkBeginCommandBuffer); LOAD(vkEndCommandBuffer); LOAD(vkResetCommandBuffer); LOAD(vkCmdPipelineBarrier);
    LOAD(vkCmdCopyBufferToImage); LOAD(vkCmdCopyImageToBuffer); LOAD(vkCmdClearColorImage); LOAD(vkCmdBindPipeline); LOAD(vkCmdBindDescriptorSets);
    LOAD(vkCmdPushConstants); LOAD(vkCmdDispatch); LOAD(vkQueueSubmit); LOAD(vkQueueWaitIdle);

    uint32_t npd = 0;
    VK(vkEnumeratePhysicalDevices(g_instance, &npd, NULL));
    if (npd == 0) { fprintf(stderr, "vulkan_harness: skipped: no Vulkan physical device\n"); return SKIP; }
    VkPhysicalDevice pds[16];
    if (npd > 16) npd = 16;
    VK(vkEnumeratePhysicalDevices(g_instance, &npd, pds));
    /* CRV_VK_DEVICE=<substring of deviceName> picks a device (e.g. "llvmpipe" for lavapipe, "NVIDIA" for the GPU driver) */
    const char* want = getenv("CRV_VK_DEVICE");
    VkPhysicalDevice pd = NULL;
    uint32_t qf = 0;
    char name[256] = "";
    for (uint32_t i = 0; i < npd && !pd; ++i) {
        uint64_t props[160];                                   /* VkPhysicalDeviceProperties is 824 bytes; deviceName at offset 20 */
        memset(props, 0, sizeof(props));
        vkGetPhysicalDeviceProperties(pds[i], props);
        const char* dn = (const char*)props + 20;
        if (want && !strstr(dn, want)) continue;
        uint32_t nq = 0;
        vkGetPhysicalDeviceQueueFamilyProperties(pds[i], &nq, NULL);
        VkQueueFamilyProperties qp[32];
        if (nq > 32) nq = 32;
        vkGetPhysicalDeviceQueueFamilyProperties(pds[i], &nq, qp);
        for (uint32_t q = 0; q < nq; ++q)
            if (qp[q].queueFlags & QUEUE_COMPUTE_BIT) { pd = pds[i]; qf = q; snprintf(name, sizeof(name), "%.200s", dn); break; }
    }
    if (!pd) { fprintf(stderr, "vulkan_harness: skipped: no device with a compute queue%s%s\n", want ? " matching " : "", want ? want : ""); return SKIP; }
    fprintf(stderr, "vulkan_harness: device \"%s\", queue family %u\n", name, qf);
    uint32_t features[64];                                     /* VkPhysicalDeviceFeatures: 55 VkBool32; everything supported is enabled */
    memset(features, 0, sizeof(features));                     /* (the push-constant block holds a double: shaderFloat64, PathTracer.comp:93-99) */
    vkGetPhysicalDeviceFeatures(pd, features);
    vkGetPhysicalDeviceMemoryProperties(pd, &memprops);
    const float prio = 1.0f;
    VkDeviceQueueCreateInfo qci = {ST_DEVICE_QUEUE_CREATE_INFO, NULL, 0, qf, 1, &prio};
    VkDeviceCreateInfo dci = {ST_DEVICE_CREATE_INFO, NULL, 0, 1, &qci, 0, NULL, 0, NULL, features};
    VK(vkCreateDevice(pd, &dci, NULL, &dev));
    VkQueue queue;
    vkGetDeviceQueue(dev, qf, 0, &queue);
    VkCommandPoolCreateInfo cpi = {ST_COMMAND_POOL_CREATE_INFO, NULL, /* RESET_COMMAND_BUFFER */ 2, qf};
    VkCommandPool pool;
    VK(vkCreateCommandPool(dev, &cpi, NULL, &pool));
    VkCommandBufferAllocateInfo cai = {ST_COMMAND_BUFFER_ALLOCATE_INFO, NULL, pool, 0, 1};
    VkCommandBuffer cb;
    VK(vkAllocateCommandBuffers(dev, &cai, &cb));

    /* ---- resources: cloud image + sampler (main.cpp:277-295), shadow-volume image (:802-804), result image (:405-414), UBOs (:796-801) ---- */
    const size_t grid_bytes = (size_t)nx * ny * nz * 4, sv_bytes = (size_t)NSV * NSV * NSV * 4, img_bytes = (size_t)W * H * 16;
    DeviceImage cloud = device_image(IMAGE_TYPE_3D, FORMAT_R32_SFLOAT, nx, ny, nz, IMG_USAGE_SAMPLED | IMG_USAGE_TRANSFER_DST);
    DeviceImage shadow = device_image(IMAGE_TYPE_3D, FORMAT_R32_SFLOAT, NSV, NSV, NSV, IMG_USAGE_STORAGE | IMG_USAGE_SAMPLED | IMG_USAGE_TRANSFER_SRC);
    DeviceImage result = device_image(IMAGE_TYPE_2D, FORMAT_R32G32B32A32_SFLOAT, W, H, 1, IMG_USAGE_STORAGE | IMG_USAGE_TRANSFER_SRC | IMG_USAGE_TRANSFER_DST);
    VkSamplerCreateInfo sci;
    memset(&sci, 0, sizeof(sci));
    sci.sType = ST_SAMPLER_CREATE_INFO;                        /* Initializers.cpp:408-418: LINEAR / CLAMP_TO_EDGE, everything else zero */
    sci.magFilter = sci.minFilter = FILTER_LINEAR;
    sci.addressModeU = sci.addressModeV = sci.addressModeW = ADDRESS_CLAMP_TO_EDGE;
    VkSampler sampler;
    VK(vkCreateSampler(dev, &sci, NULL, &sampler));
    HostBuffer staging = host_buffer(grid_bytes > sv_bytes ? (grid_bytes > img_bytes ? grid_bytes : img_bytes) : (sv_bytes > img_bytes ? sv_bytes : img_bytes),
                                     BUF_USAGE_TRANSFER_SRC | BUF_USAGE_TRANSFER_DST);
    HostBuffer ubo_cam = host_buffer(80, BUF_USAGE_UNIFORM), ubo_cloud = host_buffer(64, BUF_USAGE_UNIFORM), ubo_par = host_buffer(32, BUF_USAGE_UNIFORM),
               ubo_sv = host_buffer(160, BUF_USAGE_UNIFORM);
    memcpy(ubo_cam.map, p_cam, 80); memcpy(ubo_cloud.map, p_cloud, 64); memcpy(ubo_par.map, p_par, 32); memcpy(ubo_sv.map, p_sv, 160);

    /* ---- descriptor set layouts, pipeline layouts, pipelines ---- */
    const VkDescriptorSetLayoutBinding sv_bind[4] = {                                   /* RenderTechniqueSV.cpp:13-22 */
        {0, DESC_STORAGE_IMAGE, 1, STAGE_COMPUTE_BIT, NULL}, {1, DESC_UNIFORM_BUFFER, 1, STAGE_COMPUTE_BIT, NULL},
        {2, DESC_COMBINED_IMAGE_SAMPLER, 1, STAGE_COMPUTE_BIT, NULL}, {3, DESC_UNIFORM_BUFFER, 1, STAGE_COMPUTE_BIT, NULL}};
    const VkDescriptorSetLayoutBinding pt_bind[7] = {                                   /* RenderTechniquePT.cpp:20-35 */
        {0, DESC_STORAGE_IMAGE, 1, STAGE_COMPUTE_BIT, NULL}, {1, DESC_UNIFORM_BUFFER, 1, STAGE_COMPUTE_BIT, NULL},
        {2, DESC_COMBINED_IMAGE_SAMPLER, 1, STAGE_COMPUTE_BIT, NULL}, {3, DESC_UNIFORM_BUFFER, 1, STAGE_COMPUTE_BIT, NULL},
        {4, DESC_UNIFORM_BUFFER, 1, STAGE_COMPUTE_BIT, NULL}, {5, DESC_COMBINED_IMAGE_SAMPLER, 1, STAGE_COMPUTE_BIT, NULL},
        {6, DESC_UNIFORM_BUFFER, 1, STAGE_COMPUTE_BIT, NULL}};
    VkDescriptorSetLayout sv_dsl, pt_dsl;
    VkDescriptorSetLayoutCreateInfo li = {ST_DESCRIPTOR_SET_LAYOUT_CREATE_INFO, NULL, 0, 4, sv_bind};
    VK(vkCreateDescriptorSetLayout(dev, &li, NULL, &sv_dsl));
    li.bindingCount = 7; li.pBindings = pt_bind;
    VK(vkCreateDescriptorSetLayout(dev, &li, NULL, &pt_dsl));
    VkPipelineLayout sv_pl, pt_pl;
    VkPipelineLayoutCreateInfo pli = {ST_PIPELINE_LAYOUT_CREATE_INFO, NULL, 0, 1, &sv_dsl, 0, NULL};   /* no push constants (RenderTechniqueSV.cpp:26) */
    VK(vkCreatePipelineLayout(dev, &pli, NULL, &sv_pl));
    const VkPushConstantRange pcr = {STAGE_COMPUTE_BIT, 0, 24};                         /* sizeof(PushConstants) (RenderTechniquePT.cpp:39-42) */
    pli.pSetLayouts = &pt_dsl; pli.pushConstantRangeCount = 1; pli.pPushConstantRanges = &pcr;
    VK(vkCreatePipelineLayout(dev, &pli, NULL, &pt_pl));
    VkPipeline sv_pipe = compute_pipeline(argv[1], sv_pl), pt_pipe = compute_pipeline(argv[2], pt_pl);

    /* ---- descriptor pool + sets (VulkanDescriptorPool.cpp:24-47), updates (RenderTechnique.cpp:14-18) ---- */
    const VkDescriptorPoolSize sizes[3] = {{DESC_STORAGE_IMAGE, 2}, {DESC_UNIFORM_BUFFER, 6}, {DESC_COMBINED_IMAGE_SAMPLER, 3}};
    VkDescriptorPoolCreateInfo dpi = {ST_DESCRIPTOR_POOL_CREATE_INFO, NULL, 0, 2, 3, sizes};
    VkDescriptorPool dpool;
    VK(vkCreateDescriptorPool(dev, &dpi, NULL, &dpool));
    VkDescriptorSet sv_set, pt_set;
    VkDescriptorSetAllocateInfo dai = {ST_DESCRIPTOR_SET_ALLOCATE_INFO, NULL, dpool, 1, &sv_dsl};
    VK(vkAllocateDescriptorSets(dev, &dai, &sv_set));
    dai.pSetLayouts = &pt_dsl;
    VK(vkAllocateDescriptorSets(dev, &dai, &pt_set));
    const VkDescriptorImageInfo i_sv_store = {NULL, shadow.view, LAYOUT_GENERAL}, i_cloud = {sampler, cloud.view, LAYOUT_SHADER_READ_ONLY},
                                i_sv_sample = {sampler, shadow.view, LAYOUT_GENERAL}, i_result = {NULL, result.view, LAYOUT_GENERAL};
    const VkDescriptorBufferInfo b_cam = {ubo_cam.buf, 0, 80}, b_cloud = {ubo_cloud.buf, 0, 64}, b_par = {ubo_par.buf, 0, 32}, b_sv = {ubo_sv.buf, 0, 160};
    const VkWriteDescriptorSet writes[11] = {
        {ST_WRITE_DESCRIPTOR_SET, NULL, sv_set, 0, 0, 1, DESC_STORAGE_IMAGE, &i_sv_store, NULL, NULL},
        {ST_WRITE_DESCRIPTOR_SET, NULL, sv_set, 1, 0, 1, DESC_UNIFORM_BUFFER, NULL, &b_sv, NULL},
        {ST_WRITE_DESCRIPTOR_SET, NULL, sv_set, 2, 0, 1, DESC_COMBINED_IMAGE_SAMPLER, &i_cloud, NULL, NULL},
        {ST_WRITE_DESCRIPTOR_SET, NULL, sv_set, 3, 0, 1, DESC_UNIFORM_BUFFER, NULL, &b_cloud, NULL},
        {ST_WRITE_DESCRIPTOR_SET, NULL, pt_set, 0, 0, 1, DESC_STORAGE_IMAGE, &i_result, NULL, NULL},
        {ST_WRITE_DESCRIPTOR_SET, NULL, pt_set, 1, 0, 1, DESC_UNIFORM_BUFFER, NULL, &b_cam, NULL},
        {ST_WRITE_DESCRIPTOR_SET, NULL, pt_set, 2, 0, 1, DESC_COMBINED_IMAGE_SAMPLER, &i_cloud, NULL, NULL},
        {ST_WRITE_DESCRIPTOR_SET, NULL, pt_set, 3, 0, 1, DESC_UNIFORM_BUFFER, NULL, &b_cloud, NULL},
        {ST_WRITE_DESCRIPTOR_SET, NULL, pt_set, 4, 0, 1, DESC_UNIFORM_BUFFER, NULL, &b_par, NULL},
        {ST_WRITE_DESCRIPTOR_SET, NULL, pt_set, 5, 0, 1, DESC_COMBINED_IMAGE_SAMPLER, &i_sv_sample, NULL, NULL},
        {ST_WRITE_DESCRIPTOR_SET, NULL, pt_set, 6, 0, 1, DESC_UNIFORM_BUFFER, NULL, &b_sv, NULL}};
    vkUpdateDescriptorSets(dev, 11, writes, 0, NULL);

    const VkCommandBufferBeginInfo begin = {ST_COMMAND_BUFFER_BEGIN_INFO, NULL, /* ONE_TIME_SUBMIT */ 1, NULL};
    const VkSubmitInfo submit = {ST_SUBMIT_INFO, NULL, 0, NULL, NULL, 1, &cb, 0, NULL};
    FILE* out = fopen(argv[4], "wb");
    if (!out) DIE("cannot create %s", argv[4]);

    /* ---- UpdateCloudData (main.cpp:266-335): staging -> cloud image; UpdateShadowVolume (:223-264): one blocking dispatch ---- */
    memcpy(staging.map, grid, grid_bytes);
    VK(vkBeginCommandBuffer(cb, &begin));
    transition(cb, cloud.img, LAYOUT_UNDEFINED, LAYOUT_TRANSFER_DST, 0, ACCESS_TRANSFER_WRITE, PSTAGE_TOP, PSTAGE_TRANSFER);
    const VkBufferImageCopy up = {0, 0, 0, {ASPECT_COLOR, 0, 0, 1}, {0, 0, 0}, {nx, ny, nz}};
    vkCmdCopyBufferToImage(cb, staging.buf, cloud.img, LAYOUT_TRANSFER_DST, 1, &up);
    transition(cb, cloud.img, LAYOUT_TRANSFER_DST, LAYOUT_SHADER_READ_ONLY, ACCESS_TRANSFER_WRITE, ACCESS_SHADER_READ, PSTAGE_TRANSFER, PSTAGE_COMPUTE);
    transition(cb, shadow.img, LAYOUT_UNDEFINED, LAYOUT_GENERAL, 0, ACCESS_SHADER_WRITE, PSTAGE_TOP, PSTAGE_COMPUTE);      /* RenderTechniqueSV.cpp:104 */
    vkCmdBindPipeline(cb, BIND_POINT_COMPUTE, sv_pipe);
    vkCmdBindDescriptorSets(cb, BIND_POINT_COMPUTE, sv_pl, 0, 1, &sv_set, 0, NULL);
    vkCmdDispatch(cb, NSV / 32 + 1, NSV / 32 + 1, 1);                                                                       /* RenderTechniqueSV.cpp:110 */
    transition(cb, shadow.img, LAYOUT_GENERAL, LAYOUT_TRANSFER_SRC, ACCESS_SHADER_WRITE, ACCESS_TRANSFER_READ, PSTAGE_COMPUTE, PSTAGE_TRANSFER);
    const VkBufferImageCopy down_sv = {0, 0, 0, {ASPECT_COLOR, 0, 0, 1}, {0, 0, 0}, {NSV, NSV, NSV}};
    vkCmdCopyImageToBuffer(cb, shadow.img, LAYOUT_TRANSFER_SRC, staging.buf, 1, &down_sv);
    transition(cb, shadow.img, LAYOUT_TRANSFER_SRC, LAYOUT_GENERAL, ACCESS_TRANSFER_READ, ACCESS_SHADER_READ, PSTAGE_TRANSFER, PSTAGE_COMPUTE);
    /* headless: ONE zero-cleared result image, frameCount counts from the scene's value */
    transition(cb, result.img, LAYOUT_UNDEFINED, LAYOUT_TRANSFER_DST, 0, ACCESS_TRANSFER_WRITE, PSTAGE_TOP, PSTAGE_TRANSFER);
    const VkClearColorValue zero = {{0.0f, 0.0f, 0.0f, 0.0f}};
    const VkImageSubresourceRange whole = {ASPECT_COLOR, 0, 1, 0, 1};
    vkCmdClearColorImage(cb, result.img, LAYOUT_TRANSFER_DST, &zero, 1, &whole);
    transition(cb, result.img, LAYOUT_TRANSFER_DST, LAYOUT_GENERAL, ACCESS_TRANSFER_WRITE, ACCESS_SHADER_READ | ACCESS_SHADER_WRITE, PSTAGE_TRANSFER, PSTAGE_COMPUTE);
    VK(vkEndCommandBuffer(cb));
    VK(vkQueueSubmit(queue, 1, &submit, NULL));
    VK(vkQueueWaitIdle(queue));                                                                                              /* main.cpp:244-249 */
    if (fwrite(staging.map, 1, sv_bytes, out) != sv_bytes) DIE("short write");

    /* ---- RenderLoop / DrawFrame (main.cpp:714-736, 603-712): per frame push constants + one dispatch ---- */
    for (uint32_t f = 0; f < NF; ++f) {
        struct { double time; int32_t seed; uint32_t frameCount; float pmRadius; uint32_t currentBuffer; } pc = {0.0, 0, 0, 0.0f, 0};   /* UniformBuffers.h:6-13 */
        memcpy(&pc.seed, p_frames + 8 * (size_t)f, 4);
        memcpy(&pc.frameCount, p_frames + 8 * (size_t)f + 4, 4);
        VK(vkResetCommandBuffer(cb, 0));
        VK(vkBeginCommandBuffer(cb, &begin));
        vkCmdBindPipeline(cb, BIND_POINT_COMPUTE, pt_pipe);
        vkCmdPushConstants(cb, pt_pl, STAGE_COMPUTE_BIT, 0, 24, &pc);                                                       /* RenderTechniquePT.cpp:132 */
        vkCmdBindDescriptorSets(cb, BIND_POINT_COMPUTE, pt_pl, 0, 1, &pt_set, 0, NULL);
        vkCmdDispatch(cb, W / 32 + 1, H / 32 + 1, 1);                                                                       /* RenderTechniquePT.cpp:141 */
        const VkMemoryBarrier mb = {ST_MEMORY_BARRIER, NULL, ACCESS_SHADER_WRITE, ACCESS_SHADER_READ | ACCESS_SHADER_WRITE};
        vkCmdPipelineBarrier(cb, PSTAGE_COMPUTE, PSTAGE_COMPUTE, 0, 1, &mb, 0, NULL, 0, NULL);
        VK(vkEndCommandBuffer(cb));
        VK(vkQueueSubmit(queue, 1, &submit, NULL));
        VK(vkQueueWaitIdle(queue));
    }
    VK(vkResetCommandBuffer(cb, 0));
    VK(vkBeginCommandBuffer(cb, &begin));
    transition(cb, result.img, LAYOUT_GENERAL, LAYOUT_TRANSFER_SRC, ACCESS_SHADER_WRITE, ACCESS_TRANSFER_READ, PSTAGE_COMPUTE, PSTAGE_TRANSFER);
    const VkBufferImageCopy down_img = {0, 0, 0, {ASPECT_COLOR, 0, 0, 1}, {0, 0, 0}, {W, H, 1}};
    vkCmdCopyImageToBuffer(cb, result.img, LAYOUT_TRANSFER_SRC, staging.buf, 1, &down_img);
    VK(vkEndCommandBuffer(cb));
    VK(vkQueueSubmit(queue, 1, &submit, NULL));
    VK(vkQueueWaitIdle(queue));
    if (fwrite(staging.map, 1, img_bytes, out) != img_bytes) DIE("short write");
    fclose(out);
    fprintf(stderr, "vulkan_harness: %u^3 shadow volume and %ux%u image after %u frame(s) written to %s\n", NSV, W, H, NF, argv[4]);
    return 0;
}
