"""Headless mirror of the reference's RenderTechnique interface over the C ABI.

Reference: RenderTechnique.h:15-66 (abstract operator API), RenderTechniquePT.h:8-40 / .cpp:12-168,
RenderTechniqueSV.h:7-39 / .cpp:4-113.  Same method names and argument meaning; Vulkan handle types are
replaced by device-resource objects:

    VulkanDevice             -> Context       (one CUDA device)
    VulkanImage R32F 3D      -> Volume        (cloud density / shadow volume; LINEAR, BRICK or TEX3D storage)
    VulkanImage RGBA32F 2D   -> Image         (progressive result image)
    VkDescriptorBufferInfo   -> the host parameter struct itself (structs.py)
    VkDescriptorImageInfo    -> Volume
    VkCommandBuffer          -> a CUDA stream handle (int, 0 = default stream)

Everything computes on the GPU through libcrv_b200.so; nothing here falls back to the CPU.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np

from . import capi
from .capi import check, lib
from .structs import (CameraProperties, CloudProperties, Parameters, PathTracerExtensions, PhotonMapProperties, PushConstants,
                      ShadowVolumeProperties)


def _stream(s) -> C.c_void_p:
    return C.c_void_p(int(s) if s else 0)


class Context:
    """VulkanInstance + VulkanPhysicalDevice + VulkanDevice (main.cpp:774-793) -> one CUDA device."""

    def __init__(self, device: int = 0):
        self.handle = C.c_void_p()
        check(lib.crv_ctx_create(int(device), C.byref(self.handle)))
        self.device = device

    def synchronize(self, stream=0):
        check(lib.crv_ctx_synchronize(self.handle, _stream(stream)))

    @property
    def launch_count(self) -> int:
        n = C.c_uint64()
        check(lib.crv_ctx_launch_count(self.handle, C.byref(n)))
        return n.value

    @property
    def sm_count(self) -> int:
        n = C.c_int()
        check(lib.crv_ctx_sm_count(self.handle, C.byref(n)))
        return n.value

    def close(self):
        if self.handle:
            lib.crv_ctx_destroy(self.handle)
            self.handle = C.c_void_p()


class Stream:
    """A command stream of the context's device (VkQueue + VkCommandBuffer role): a non-blocking cudaStream_t."""

    def __init__(self, ctx: Context):
        self.ctx = ctx
        h = C.c_void_p()
        check(lib.crv_stream_create(ctx.handle, C.byref(h)))
        self.handle = h.value or 0

    def __int__(self):
        return self.handle

    def synchronize(self):
        self.ctx.synchronize(self.handle)

    def close(self):
        if self.handle:
            lib.crv_stream_destroy(self.ctx.handle, C.c_void_p(self.handle))
            self.handle = 0


class Volume:
    """R32_SFLOAT 3D image + LINEAR/CLAMP_TO_EDGE sampler (main.cpp:277-285, Initializers.cpp:408-418)."""

    def __init__(self, ctx: Context, nx: int, ny: int, nz: int, layout="linear"):
        self.ctx = ctx
        self.shape = (int(nx), int(ny), int(nz))
        self.layout = capi.LAYOUTS[layout] if isinstance(layout, str) else int(layout)
        self.handle = C.c_void_p()
        check(lib.crv_volume_create(ctx.handle, nx, ny, nz, self.layout, C.byref(self.handle)))

    def upload(self, data: np.ndarray, stream=0, wait: bool = True):
        """x-fastest host floats (what Grid3D::GetData holds; staging copy of main.cpp:288-295).
        wait=False: asynchronous on `stream`; the caller keeps `data` (contiguous float32, ideally pinned) alive."""
        a = np.ascontiguousarray(data, dtype=np.float32).reshape(-1)
        if a.size != self.shape[0] * self.shape[1] * self.shape[2]:
            raise ValueError("volume upload: size mismatch")
        if not wait and a.ctypes.data != np.asarray(data).ctypes.data:
            raise ValueError("volume upload: wait=False needs a contiguous float32 array (no temporary copy)")
        check(lib.crv_volume_upload(self.handle, a.ctypes.data_as(C.c_void_p), 0, _stream(stream)))
        if wait:
            self.ctx.synchronize(stream)      # `a` may be a temporary

    def upload_device(self, device_ptr: int, stream=0):
        check(lib.crv_volume_upload(self.handle, C.c_void_p(int(device_ptr)), 1, _stream(stream)))

    def download(self, stream=0) -> np.ndarray:
        """x-fastest flat array, whatever the storage layout."""
        out = np.empty(self.shape[0] * self.shape[1] * self.shape[2], dtype=np.float32)
        check(lib.crv_volume_download(self.handle, out.ctypes.data_as(C.c_void_p), _stream(stream)))
        return out

    def majorant(self, stream=0) -> float:
        m = C.c_float()
        check(lib.crv_volume_majorant(self.handle, C.byref(m), _stream(stream)))
        return m.value

    def fbm_generate(self, seed: int = 1234, stream=0):
        check(lib.crv_fbm_generate(self.handle, C.c_uint32(seed), _stream(stream)))

    @property
    def device_ptr(self) -> int:
        p = C.c_void_p()
        check(lib.crv_volume_device_ptr(self.handle, C.byref(p)))
        return p.value or 0

    def mark_dirty(self):
        """after writing through device_ptr: majorant, brick maxima and everything cached from them are stale"""
        check(lib.crv_volume_mark_dirty(self.handle))

    def close(self):
        if self.handle:
            lib.crv_volume_destroy(self.handle)
            self.handle = C.c_void_p()


class Image:
    """R32G32B32A32_SFLOAT 2D storage image (main.cpp:405-414), zero-cleared on creation."""

    def __init__(self, ctx: Context, width: int, height: int, device_ptr: Optional[int] = None, borrowed_handle=None):
        self.ctx = ctx
        self.width, self.height = int(width), int(height)
        self.handle = C.c_void_p()
        self.borrowed = borrowed_handle is not None
        if self.borrowed:                      # an image owned by another object (crv_group_image)
            self.handle = C.c_void_p(borrowed_handle)
        elif device_ptr is None:
            check(lib.crv_image_create(ctx.handle, width, height, C.byref(self.handle)))
        else:
            check(lib.crv_image_wrap(ctx.handle, C.c_void_p(int(device_ptr)), width, height, C.byref(self.handle)))

    def clear(self, stream=0):
        check(lib.crv_image_clear(self.handle, _stream(stream)))

    def download(self, stream=0, out: Optional[np.ndarray] = None) -> np.ndarray:
        if out is None:
            out = np.empty((self.height, self.width, 4), dtype=np.float32)
        check(lib.crv_image_download(self.handle, out.ctypes.data_as(C.c_void_p), _stream(stream)))
        return out

    def upload(self, rgba: np.ndarray, stream=0):
        a = np.ascontiguousarray(rgba, dtype=np.float32)
        if a.size != self.width * self.height * 4:
            raise ValueError("image upload: size mismatch")
        check(lib.crv_image_upload(self.handle, a.ctypes.data_as(C.c_void_p), _stream(stream)))
        self.ctx.synchronize(stream)

    def scale(self, factor: float, stream=0):
        check(lib.crv_image_scale(self.handle, C.c_float(factor), _stream(stream)))

    # ---- present step and files (RenderTechniquePT.cpp:143-167: blit to the B8G8R8A8_SRGB swapchain image) ----
    def blit_srgb8(self, stream=0) -> np.ndarray:
        """HxWx4 bytes in swapchain order B,G,R,A: clamp, linear->sRGB on colour, UNORM8"""
        out = np.empty((self.height, self.width, 4), dtype=np.uint8)
        check(lib.crv_image_blit_srgb8(self.handle, out.ctypes.data_as(C.c_void_p), _stream(stream)))
        return out

    def save_png(self, path: str, stream=0):
        check(lib.crv_image_save_png(self.handle, path.encode(), _stream(stream)))

    def save_pfm(self, path: str, stream=0):
        check(lib.crv_image_save_pfm(self.handle, path.encode(), _stream(stream)))

    def save_checkpoint(self, path: str, frame_count: int, seed: int = 0, stream=0):
        check(lib.crv_image_save_checkpoint(self.handle, int(frame_count), int(seed), path.encode(), _stream(stream)))

    def load_checkpoint(self, path: str, stream=0):
        """-> (frameCount the next dispatch pushes, seed of the last rendered frame)"""
        fc, seed = C.c_uint32(), C.c_int32()
        check(lib.crv_image_load_checkpoint(self.handle, C.byref(fc), C.byref(seed), path.encode(), _stream(stream)))
        return fc.value, seed.value

    @property
    def device_ptr(self) -> int:
        p = C.c_void_p()
        check(lib.crv_image_device_ptr(self.handle, C.byref(p)))
        return p.value or 0

    def close(self):
        if self.handle and not self.borrowed:
            lib.crv_image_destroy(self.handle)
        self.handle = C.c_void_p()


class Group:
    """One member of a multi-GPU sample split (crv_group_*): owns the accumulation image(s) its technique renders
    into, in memory the other members map over NVLink, and combines them -- sum over members x scale -- in one kernel.

        g = Group(ctx, rank, world, W, H); handles = <all members' g.export(), in rank order>; g.connect(handles)
        pt.SetFrameReferences([g.image(0)]) ... RenderBatch(my_seeds, 0, ACCUM_SUM) ... g.all_reduce_sum_scale(1 / spp_total)
    """

    def __init__(self, ctx: Context, rank: int, nranks: int, width: int, height: int, nbuf: int = 1, staging_bytes: int = 0):
        self.ctx = ctx
        self.rank, self.nranks, self.nbuf = int(rank), int(nranks), int(nbuf)
        self.width, self.height = int(width), int(height)
        self.handle = C.c_void_p()
        check(lib.crv_group_create(ctx.handle, rank, nranks, width, height, nbuf, C.c_uint64(staging_bytes), C.byref(self.handle)))
        self._images = {}

    def export(self) -> bytes:
        h = capi.GroupHandle()
        check(lib.crv_group_export(self.handle, C.byref(h)))
        return bytes(h)

    def connect(self, handles: Sequence[bytes]):
        if len(handles) != self.nranks:
            raise ValueError("Group.connect: one handle per member, in rank order")
        arr = (capi.GroupHandle * self.nranks)()
        for i, b in enumerate(handles):
            if len(b) != C.sizeof(capi.GroupHandle):
                raise ValueError("Group.connect: not a group handle")
            C.memmove(C.byref(arr[i]), bytes(b), C.sizeof(capi.GroupHandle))
        check(lib.crv_group_connect(self.handle, arr, self.nranks))

    def image(self, buffer: int = 0) -> "Image":
        if buffer not in self._images:
            h = C.c_void_p()
            check(lib.crv_group_image(self.handle, buffer, C.byref(h)))
            self._images[buffer] = Image(self.ctx, self.width, self.height, borrowed_handle=h.value)
        return self._images[buffer]

    def all_reduce_sum_scale(self, scale: float, stream=0, buffer: int = 0):
        check(lib.crv_group_all_reduce_sum_scale(self.handle, buffer, C.c_float(scale), _stream(stream)))

    # ---- replicated host input: 1/N per member over PCIe, exchanged GPU to GPU by the copy engines ----
    def staging_ptr(self, buffer: int = 0) -> int:
        p = C.c_void_p()
        check(lib.crv_group_staging_ptr(self.handle, buffer, C.byref(p)))
        return p.value or 0

    def staging_all_gather(self, buffer: int, host_ptr: int, nbytes: int, stream=0):
        """host_ptr: address of the whole array in (pinned) host memory; asynchronous on `stream`"""
        check(lib.crv_group_staging_all_gather(self.handle, buffer, C.c_void_p(int(host_ptr)), C.c_uint64(nbytes), _stream(stream)))

    def staging_release(self, buffer: int = 0, stream=0):
        check(lib.crv_group_staging_release(self.handle, buffer, _stream(stream)))

    def all_reduce_lists_sum_scale(self, scale: float, others_scale: float, lists, stream=0, buffer: int = 0):
        """lists: RenderTechniquePT.PixelLists() of this member's technique; only the tracked pixels cross NVLink"""
        check(lib.crv_group_all_reduce_lists_sum_scale(self.handle, buffer, C.c_float(scale), C.c_float(others_scale), C.byref(lists), _stream(stream)))

    # ---- one-barrier combine: the fold pushes the tracked sums into the members' inboxes ----
    def push_begin(self, buffer: int = 0) -> "capi.PushTargets":
        """next picture of `buffer`: where this member's sums go (for RenderTechniquePT.RenderBatchPush)"""
        t = capi.PushTargets()
        check(lib.crv_group_push_begin(self.handle, buffer, C.byref(t)))
        return t

    def push_finish(self, scale: float, lists, stream=0, buffer: int = 0):
        """waits for the members' sums of the picture begun with push_begin and finishes the tracked pixels locally"""
        check(lib.crv_group_push_finish(self.handle, buffer, C.c_float(scale), C.byref(lists), _stream(stream)))

    def set_timeout_ms(self, ms: int):
        check(lib.crv_group_set_timeout_ms(self.handle, int(ms)))

    def status(self, stream=0):
        """raises CrvError if a barrier of an earlier reduce timed out (synchronises the stream)"""
        check(lib.crv_group_status(self.handle, _stream(stream)))

    def last_reduce_ms(self) -> float:
        ms = C.c_float()
        check(lib.crv_group_last_reduce_ms(self.handle, C.byref(ms)))
        return ms.value

    # ---- the NCCL route (north_star's wording), for comparison ----
    @staticmethod
    def nccl_unique_id() -> bytes:
        buf = (C.c_ubyte * 128)()
        check(lib.crv_group_nccl_unique_id(buf))
        return bytes(buf)

    def nccl_init(self, unique_id: bytes):
        buf = (C.c_ubyte * 128).from_buffer_copy(unique_id)
        check(lib.crv_group_nccl_init(self.handle, buf))

    def nccl_all_reduce_sum_scale(self, scale: float, stream=0, buffer: int = 0):
        check(lib.crv_group_nccl_all_reduce_sum_scale(self.handle, buffer, C.c_float(scale), _stream(stream)))

    def close(self):
        if self.handle:
            lib.crv_group_destroy(self.handle)
            self.handle = C.c_void_p()
            self._images = {}


class RenderTechnique:
    """Abstract operator interface (RenderTechnique.h:15-66).

    QueueUpdate* record a pending descriptor write; UpdateDescriptorSets flushes them (RenderTechnique.cpp:14-18)
    -- the parameter-struct bytes are snapshotted at that moment (VulkanBuffer::SetData + vkUpdateDescriptorSets).
    """

    def __init__(self, device: Context, pushConstants: PushConstants):
        self.m_device = device
        self.m_pushConstants = pushConstants

    def GetRequiredSetCount(self) -> int:      # RenderTechniquePT.cpp:124-127 / RenderTechniqueSV.cpp:97-100
        return 1

    def UpdateDescriptorSets(self):
        raise NotImplementedError

    def SetFrameReferences(self, frameImages, frameImageViews=None, swapchain=None):
        raise NotImplementedError

    def ClearFrameReferences(self):
        raise NotImplementedError

    def QueueUpdateCloudData(self, cloudBufferInfo, imageIdx=0):
        raise NotImplementedError

    def QueueUpdateCloudDataSampler(self, cloudImageInfo, imageIdx=0):
        raise NotImplementedError

    def QueueUpdateCameraProperties(self, cameraBufferInfo, imageIdx=0):
        raise NotImplementedError

    def QueueUpdateParameters(self, parametersBufferInfo, imageIdx=0):
        raise NotImplementedError

    def QueueUpdateShadowVolume(self, shadowVolumeBufferInfo, imageIdx=0):
        raise NotImplementedError

    def QueueUpdateShadowVolumeSampler(self, shadowVolumeImageInfo, imageIdx=0):
        raise NotImplementedError

    def RecordDrawCommands(self, commandBuffer=0, imageIndex=0):
        raise NotImplementedError


class RenderTechniqueSV(RenderTechnique):
    """Shadow-volume precompute (RenderTechniqueSV.h:7-39): bindings 0 SV storage image, 1 SV UBO, 2 cloud
    sampler, 3 cloud UBO; dispatch (N/32+1)^2 groups of ShadowVolume.comp (RenderTechniqueSV.cpp:102-113)."""

    def __init__(self, device: Context, shadowVolumeProperties: ShadowVolumeProperties, pushConstants: PushConstants):
        super().__init__(device, pushConstants)
        self.m_shadowVolumeProperties = shadowVolumeProperties
        self._keep = []
        self.handle = C.c_void_p()
        check(lib.crv_sv_create(device.handle, C.byref(shadowVolumeProperties), C.byref(pushConstants), C.byref(self.handle)))

    def SetFrameReferences(self, frameImages: Sequence[Volume], frameImageViews=None, swapchain=None):
        self._keep.append(frameImages[0])
        check(lib.crv_sv_set_frame_references(self.handle, frameImages[0].handle))

    def ClearFrameReferences(self):
        check(lib.crv_sv_clear_frame_references(self.handle))

    def QueueUpdateCloudData(self, cloudBufferInfo: CloudProperties, imageIdx=0):
        self._keep.append(cloudBufferInfo)
        check(lib.crv_sv_queue_update_cloud_data(self.handle, C.byref(cloudBufferInfo)))

    def QueueUpdateCloudDataSampler(self, cloudImageInfo: Volume, imageIdx=0):
        self._keep.append(cloudImageInfo)
        check(lib.crv_sv_queue_update_cloud_data_sampler(self.handle, cloudImageInfo.handle))

    def QueueUpdateCameraProperties(self, cameraBufferInfo, imageIdx=0):      # no-op (.cpp:79-81)
        check(lib.crv_sv_queue_update_camera_properties(self.handle, None))

    def QueueUpdateParameters(self, parametersBufferInfo, imageIdx=0):        # no-op (.cpp:83-85)
        check(lib.crv_sv_queue_update_parameters(self.handle, None))

    def QueueUpdateShadowVolume(self, shadowVolumeBufferInfo: ShadowVolumeProperties, imageIdx=0):
        self._keep.append(shadowVolumeBufferInfo)
        check(lib.crv_sv_queue_update_shadow_volume(self.handle, C.byref(shadowVolumeBufferInfo)))

    def QueueUpdateShadowVolumeSampler(self, shadowVolumeImageInfo: Volume, imageIdx=0):
        self._keep.append(shadowVolumeImageInfo)
        check(lib.crv_sv_queue_update_shadow_volume_sampler(self.handle, shadowVolumeImageInfo.handle))

    def UpdateDescriptorSets(self):
        check(lib.crv_sv_update_descriptor_sets(self.handle))
        self._keep = self._keep[-8:]

    def RecordDrawCommands(self, commandBuffer=0, imageIndex=0):
        check(lib.crv_sv_record_draw_commands(self.handle, _stream(commandBuffer), int(imageIndex)))

    def last_build_ms(self) -> float:
        ms = C.c_float()
        check(lib.crv_sv_last_build_ms(self.handle, C.byref(ms)))
        return ms.value

    SV_MODES = {"auto": 0, "generic": 1, "single_pass": 2, "two_pass": 3, "fused": 4, "fused_v1": 5, "fused_rows": 6}

    def SetTuning(self, mode="auto"):
        """which kernels build the volume (crv_sv_mode); every mode produces the same bits"""
        check(lib.crv_sv_set_tuning(self.handle, self.SV_MODES[mode] if isinstance(mode, str) else int(mode)))

    def last_build_mode(self) -> str:
        v = C.c_int()
        check(lib.crv_sv_last_build_mode(self.handle, C.byref(v)))
        return {v: k for k, v in self.SV_MODES.items()}[v.value]

    def close(self):
        if self.handle:
            lib.crv_sv_destroy(self.handle)
            self.handle = C.c_void_p()


class RenderTechniquePT(RenderTechnique):
    """Path tracing (RenderTechniquePT.h:8-40): bindings 0 result image, 1 camera UBO, 2 cloud sampler, 3 cloud UBO,
    4 parameters UBO, 5 SV sampler, 6 SV UBO (RenderTechniquePT.cpp:20-35); push constants seed/frameCount; one
    RecordDrawCommands = one dispatch of PathTracer.comp = one progressive frame (.cpp:129-141; the swapchain blit
    of .cpp:143-167 is dropped: headless)."""

    def __init__(self, device: Context, swapchain, cameraProperties: CameraProperties, pushConstants: PushConstants):
        super().__init__(device, pushConstants)
        self.m_cameraProperties = cameraProperties
        self.m_swapchain = swapchain            # unused: headless
        self._keep = []
        self._images = []
        self.handle = C.c_void_p()
        check(lib.crv_pt_create(device.handle, C.byref(cameraProperties), C.byref(pushConstants), C.byref(self.handle)))

    def SetFrameReferences(self, frameImages: Sequence[Image], frameImageViews=None, swapchain=None):
        self._images = list(frameImages)
        arr = (C.c_void_p * len(frameImages))(*[im.handle for im in frameImages])
        check(lib.crv_pt_set_frame_references(self.handle, arr, len(frameImages)))

    def ClearFrameReferences(self):
        self._images = []
        check(lib.crv_pt_clear_frame_references(self.handle))

    def QueueUpdateCloudData(self, cloudBufferInfo: CloudProperties, imageIdx=0):
        self._keep.append(cloudBufferInfo)
        check(lib.crv_pt_queue_update_cloud_data(self.handle, C.byref(cloudBufferInfo), int(imageIdx)))

    def QueueUpdateCloudDataSampler(self, cloudImageInfo: Volume, imageIdx=0):
        self._keep.append(cloudImageInfo)
        check(lib.crv_pt_queue_update_cloud_data_sampler(self.handle, cloudImageInfo.handle, int(imageIdx)))

    def QueueUpdateCameraProperties(self, cameraBufferInfo: CameraProperties, imageIdx=0):
        self._keep.append(cameraBufferInfo)
        check(lib.crv_pt_queue_update_camera_properties(self.handle, C.byref(cameraBufferInfo), int(imageIdx)))

    def QueueUpdateParameters(self, parametersBufferInfo: Parameters, imageIdx=0):
        self._keep.append(parametersBufferInfo)
        check(lib.crv_pt_queue_update_parameters(self.handle, C.byref(parametersBufferInfo), int(imageIdx)))

    def QueueUpdateShadowVolume(self, shadowVolumeBufferInfo: ShadowVolumeProperties, imageIdx=0):
        self._keep.append(shadowVolumeBufferInfo)
        check(lib.crv_pt_queue_update_shadow_volume(self.handle, C.byref(shadowVolumeBufferInfo), int(imageIdx)))

    def QueueUpdateShadowVolumeSampler(self, shadowVolumeImageInfo: Volume, imageIdx=0):
        self._keep.append(shadowVolumeImageInfo)
        check(lib.crv_pt_queue_update_shadow_volume_sampler(self.handle, shadowVolumeImageInfo.handle, int(imageIdx)))

    def UpdateDescriptorSets(self):
        check(lib.crv_pt_update_descriptor_sets(self.handle))
        self._keep = self._keep[-16:]

    def RecordDrawCommands(self, commandBuffer=0, imageIndex=0):
        check(lib.crv_pt_record_draw_commands(self.handle, _stream(commandBuffer), int(imageIndex)))

    # ---- extensions beyond the reference interface -----------------------------------------------------
    def SetExtensions(self, ext: PathTracerExtensions):
        check(lib.crv_pt_set_extensions(self.handle, C.byref(ext)))

    def RenderBatch(self, seeds, first_frame_count: int = 0, accum_mode: int = capi.ACCUM_RUNNING_MEAN,
                    commandBuffer=0, imageIndex=0):
        """len(seeds) progressive frames in one call (bit-identical to as many RecordDrawCommands)."""
        s = np.ascontiguousarray(seeds, dtype=np.int32)
        check(lib.crv_pt_render_batch(self.handle, s.ctypes.data_as(C.POINTER(C.c_int32)), s.size,
                                      C.c_uint32(first_frame_count), int(accum_mode), _stream(commandBuffer), int(imageIndex)))

    def RenderBatchPush(self, seeds, targets, others_scale: float, commandBuffer=0, imageIndex=0):
        """CRV_ACCUM_SUM render of a group member whose fold sends the tracked sums to the members' inboxes
        (targets = Group.push_begin(); the image needs no clear; follow with Group.push_finish)"""
        s = np.ascontiguousarray(seeds, dtype=np.int32)
        check(lib.crv_pt_render_batch_push(self.handle, s.ctypes.data_as(C.POINTER(C.c_int32)), s.size, C.byref(targets),
                                           C.c_float(others_scale), _stream(commandBuffer), int(imageIndex)))

    def GetStats(self, stream=0) -> capi.PTStats:
        st = capi.PTStats()
        check(lib.crv_pt_get_stats(self.handle, C.byref(st), _stream(stream)))
        return st

    def ResetStats(self, stream=0):
        check(lib.crv_pt_reset_stats(self.handle, _stream(stream)))

    def LastRenderMs(self):
        """(device ms spent in the tracker kernel launches of the last render call, number of launches)"""
        ms, n = C.c_float(), C.c_uint()
        check(lib.crv_pt_last_render_ms(self.handle, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def SetDebugCounters(self, enable: bool):
        check(lib.crv_pt_set_debug_counters(self.handle, 1 if enable else 0))

    def DownloadDebugCounters(self, stream=0):
        w, h = self.m_cameraProperties.GetWidth(), self.m_cameraProperties.GetHeight()
        steps = np.empty((h, w), dtype=np.uint32)
        coll = np.empty((h, w), dtype=np.uint32)
        check(lib.crv_pt_download_debug_counters(self.handle, steps.ctypes.data_as(C.c_void_p),
                                                 coll.ctypes.data_as(C.c_void_p), _stream(stream)))
        return steps, coll

    def SetTuning(self, ctas_per_sm: int = 0, track_threshold: int = 0, sample_buffer_bytes: int = 0, steps_per_decision: int = 0):
        check(lib.crv_pt_set_tuning(self.handle, int(ctas_per_sm), int(track_threshold), C.c_uint64(sample_buffer_bytes),
                                    int(steps_per_decision)))

    SCHEDULERS = {"pool": 0, "lanes": 1}

    def SetScheduler(self, scheduler="pool"):
        """which persistent-thread scheduler runs the tracker; both produce the same bits"""
        check(lib.crv_pt_set_scheduler(self.handle, self.SCHEDULERS[scheduler] if isinstance(scheduler, str) else int(scheduler)))

    def PixelLists(self) -> capi.PixelLists:
        """device-side lists of the tracked / other pixels of the last render call (crv_pt_pixel_lists)"""
        out = capi.PixelLists()
        check(lib.crv_pt_pixel_lists(self.handle, C.byref(out)))
        return out

    def SetCoarseShift(self, min_shift: int = -1):
        """at least this cell size (8 << shift voxels) for the tracker's coarse threshold table; -1 = automatic"""
        check(lib.crv_pt_set_coarse_shift(self.handle, int(min_shift)))

    def SetJobOrder(self, order: int = -1):
        """0 frame-major, 1 pixel-major over the hit list sorted by expected work per path, -1 automatic; identical samples"""
        check(lib.crv_pt_set_job_order(self.handle, int(order)))

    def GetCoarseShift(self) -> int:
        v = C.c_int()
        check(lib.crv_pt_get_coarse_shift(self.handle, C.byref(v)))
        return v.value

    def SetSkipLevel(self, level: int = 1):
        """1 (default): ray sections that provably cannot collide are not tracked (same image bits);
        0: every step of the reference is executed, so the statistics count the reference's work"""
        check(lib.crv_pt_set_skip_level(self.handle, int(level)))

    def close(self):
        if self.handle:
            lib.crv_pt_destroy(self.handle)
            self.handle = C.c_void_p()


class RenderTechniquePPM:
    """RenderTechniquePPM (RenderTechniquePPM.h/.cpp): progressive photon mapping -- per frame the radius update,
    the photon-tracing dispatch (shaders/PPM_PT.comp) and the photon-estimate dispatch (shaders/PPM_PE.comp)."""

    def __init__(self, device: Context, photonMapProperties, pushConstants: PushConstants, initialRadius: float = 10.0):
        self.ctx = device
        self.photonMapProperties = photonMapProperties       # borrowed, like the reference's pointer member
        self.pushConstants = pushConstants                   # UpdateRadius writes pmRadius into it
        self.handle = C.c_void_p()
        check(lib.crv_ppm_create(device.handle, C.byref(photonMapProperties), C.byref(pushConstants), C.c_float(initialRadius),
                                 C.byref(self.handle)))
        self._keep = []
        self._images = []

    def SetFrameReferences(self, images):
        self._images = list(images)
        arr = (C.c_void_p * len(self._images))(*[im.handle for im in self._images])
        check(lib.crv_ppm_set_frame_references(self.handle, arr, len(self._images)))

    def QueueUpdateCameraProperties(self, cameraProperties, imageIdx=0):
        self._keep.append(cameraProperties)
        check(lib.crv_ppm_queue_update_camera_properties(self.handle, C.byref(cameraProperties)))

    def QueueUpdateShadowVolume(self, shadowVolumeProperties, imageIdx=0):
        self._keep.append(shadowVolumeProperties)
        check(lib.crv_ppm_queue_update_shadow_volume(self.handle, C.byref(shadowVolumeProperties)))

    def QueueUpdateShadowVolumeSampler(self, shadowImage: Volume, imageIdx=0):
        self._keep.append(shadowImage)
        check(lib.crv_ppm_queue_update_shadow_volume_sampler(self.handle, shadowImage.handle))

    def RecordDrawCommands(self, stream=0, imageIndex: int = 0):
        """RenderTechniquePPM.cpp:229-273: UpdateRadius(frameCount), tracing (200 groups), estimate"""
        check(lib.crv_ppm_record_draw_commands(self.handle, _stream(stream), int(imageIndex)))

    def UpdateRadius(self, frameNumber: int):
        check(lib.crv_ppm_update_radius(self.handle, int(frameNumber)))

    def RecordPhotonEstimate(self, stream=0, imageIndex: int = 0):
        check(lib.crv_ppm_record_photon_estimate(self.handle, _stream(stream), int(imageIndex)))

    def Upload(self, photons: np.ndarray, collisions: np.ndarray, stream=0):
        ph = np.ascontiguousarray(photons, np.float32)
        co = np.ascontiguousarray(collisions, np.uint32)
        check(lib.crv_ppm_upload(self.handle, ph.ctypes.data_as(C.c_void_p), co.ctypes.data_as(C.c_void_p), _stream(stream)))

    def LastEstimateStats(self):
        a, ms = C.c_uint64(), C.c_float()
        check(lib.crv_ppm_last_estimate_stats(self.handle, C.byref(a), C.byref(ms)))
        return {"march_steps": a.value, "ms": ms.value}

    def AllocateResources(self):
        check(lib.crv_ppm_allocate_resources(self.handle))

    def FreeResources(self):
        check(lib.crv_ppm_free_resources(self.handle))

    def QueueUpdateCloudData(self, cloudProperties, imageIdx=0):
        self._keep.append(cloudProperties)
        check(lib.crv_ppm_queue_update_cloud_data(self.handle, C.byref(cloudProperties)))

    def QueueUpdateCloudDataSampler(self, cloudImage: Volume, imageIdx=0):
        self._keep.append(cloudImage)
        check(lib.crv_ppm_queue_update_cloud_data_sampler(self.handle, cloudImage.handle))

    def QueueUpdateParameters(self, parameters, imageIdx=0):
        self._keep.append(parameters)
        check(lib.crv_ppm_queue_update_parameters(self.handle, C.byref(parameters)))

    def UpdateDescriptorSets(self):
        check(lib.crv_ppm_update_descriptor_sets(self.handle))
        self._keep = self._keep[-8:]

    def RecordPhotonTracing(self, stream=0, groups_x: int = 0):
        """the tracing half of RecordDrawCommands (RenderTechniquePPM.cpp:229-254); groups_x = 0 -> the reference's 200"""
        check(lib.crv_ppm_record_photon_tracing(self.handle, _stream(stream), int(groups_x)))

    def Download(self, stream=0):
        """-> (photons as a (cells, 32, 12) float32 array, collisions as (cells, 4) uint32)"""
        cells = self.photonMapProperties.GetTotalSize()
        ph = np.empty((cells, 32, 12), dtype=np.float32)
        co = np.empty((cells, 4), dtype=np.uint32)
        check(lib.crv_ppm_download(self.handle, ph.ctypes.data_as(C.c_void_p), co.ctypes.data_as(C.c_void_p), _stream(stream)))
        return ph, co

    def LastTraceStats(self):
        a, b, c, ms = C.c_uint64(), C.c_uint64(), C.c_uint64(), C.c_float()
        check(lib.crv_ppm_last_trace_stats(self.handle, C.byref(a), C.byref(b), C.byref(c), C.byref(ms)))
        return {"tracking_steps": a.value, "deposits": b.value, "dropped": c.value, "ms": ms.value}

    def close(self):
        if self.handle:
            lib.crv_ppm_destroy(self.handle)
            self.handle = C.c_void_p()


def rng_tap(ctx: Context, seed: int, width: int, height: int, n_draws: int, stream=0) -> np.ndarray:
    """Raw xorshift states of the first n_draws draws after warm-up for every pixel (bit-exact RNG tap)."""
    out = np.empty((height, width, n_draws), dtype=np.uint32)
    check(lib.crv_pt_rng_tap(ctx.handle, C.c_int32(seed), width, height, n_draws, out.ctypes.data_as(C.c_void_p), _stream(stream)))
    return out
