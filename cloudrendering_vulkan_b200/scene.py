"""Headless driver: the reference's call order around the two techniques, without window/swapchain/UI.

Restates what main.cpp does around the hot path -- StartSimulation (main.cpp:927-944), SetCloudProperties
(:144-163), UpdateCloudData (:266-335), UpdateShadowVolume (:223-264) and the per-frame part of RenderLoop /
DrawFrame (:714-736, :603-712: seed = std::rand(), RecordDrawCommands, frameCount++) -- on top of
RenderTechniquePT / RenderTechniqueSV.  Differences from the interactive app, on purpose:
ONE zero-cleared result image and frameCount starting at 0, so the running mean is the plain mean of the frames.

Also holds the BASELINE.json workload table (CONFIGS) and the frozen synthetic-scene parameters.
"""
from __future__ import annotations

from typing import Optional

import numpy as np

from . import capi
from .structs import (CameraProperties, CloudProperties, Parameters, PathTracerExtensions, PushConstants,
                      ShadowVolumeProperties, msvc_rand_sequence)
from .technique import Context, Image, RenderTechniquePT, RenderTechniqueSV, Volume


from .workloads import CONFIGS, Workload  # noqa: E402,F401  (re-exported)


class HeadlessRenderer:
    """g_* globals of main.cpp (:31-97) that the hot path touches, as one object."""

    def __init__(self, ctx: Context, width: int, height: int, n_sv: int = 500, cloud_layout="linear",
                 sv_layout=None, fov: float = 90.0):
        self.ctx = ctx
        self.cloud_layout = cloud_layout
        self.sv_layout = sv_layout or cloud_layout
        # --- shader resources (main.cpp:31-46) ---
        self.shadowVolumeProperties = ShadowVolumeProperties()
        self.shadowVolumeProperties.voxelAxisCount = n_sv
        self.cameraProperties = CameraProperties()
        self.cloudProperties = CloudProperties()
        self.parameters = Parameters()
        self.pushConstants = PushConstants()
        self.extensions = PathTracerExtensions(0.0)
        self.pushConstants.frameCount = 0          # headless: zero-cleared image, plain mean
        self.lightDirection = (1.0, -1.0, 0.0)     # g_UILightDirection default (UniformBuffers.h:120)
        # SetResolution before SetFOV (main.cpp:383-384)
        self.cameraProperties.SetResolution(width, height)
        self.cameraProperties.SetFOV(fov)
        # --- images ---
        self.cloudImage: Optional[Volume] = None
        self.shadowVolumeImage = Volume(ctx, n_sv, n_sv, n_sv, self.sv_layout)       # main.cpp:802-804
        self.resultImage = Image(ctx, width, height)                                  # main.cpp:405-414
        # --- techniques (main.cpp:811-812) ---
        self.shadowVolumeTechnique = RenderTechniqueSV(ctx, self.shadowVolumeProperties, self.pushConstants)
        self.pathTracingTechnique = RenderTechniquePT(ctx, None, self.cameraProperties, self.pushConstants)
        self.shadowVolumeTechnique.SetFrameReferences([self.shadowVolumeImage])      # main.cpp:850-853
        self.pathTracingTechnique.SetFrameReferences([self.resultImage])             # main.cpp:842-848
        self.pathTracingTechnique.QueueUpdateParameters(self.parameters, 0)          # main.cpp:859-883
        self.pathTracingTechnique.QueueUpdateCameraProperties(self.cameraProperties, 0)
        self.pathTracingTechnique.QueueUpdateShadowVolumeSampler(self.shadowVolumeImage, 0)
        self.pathTracingTechnique.UpdateDescriptorSets()
        self._seed_state = 0                        # srand(0) (main.cpp:949)
        self._frames_rendered = 0

    # ---- cloud ------------------------------------------------------------------------------------------
    def SetCloudProperties(self, counts, voxel_size, majorant: float):
        """main.cpp:144-163"""
        self.cloudProperties.SetFromGrid(counts, voxel_size, majorant)

    def set_cloud_host(self, data: np.ndarray, counts, voxel_size, stream=0):
        """LoadCloudFile + UpdateCloudData for an x-fastest host array (main.cpp:166-190, 266-335)."""
        self._new_cloud_image(counts)
        self.cloudImage.upload(data, stream)
        self.SetCloudProperties(counts, voxel_size, self.cloudImage.majorant(stream))
        self.UpdateCloudData(stream)

    def set_cloud_synthetic(self, n: int, seed: int = 1234, stream=0):
        """The frozen fBm scene generated on the device; box = 1000^3 world units for every n."""
        self._new_cloud_image((n, n, n))
        self.cloudImage.fbm_generate(seed, stream)
        self.SetCloudProperties((n, n, n), (1.0 / n,) * 3, self.cloudImage.majorant(stream))
        self.UpdateCloudData(stream)

    def _new_cloud_image(self, counts):
        if self.cloudImage is not None and self.cloudImage.shape != tuple(int(c) for c in counts):
            self.cloudImage.close()
            self.cloudImage = None
        if self.cloudImage is None:
            self.cloudImage = Volume(self.ctx, counts[0], counts[1], counts[2], self.cloud_layout)

    def UpdateCloudData(self, stream=0):
        """Descriptor updates of main.cpp:299-322, then UpdateShadowVolume (:334)."""
        pt, sv = self.pathTracingTechnique, self.shadowVolumeTechnique
        pt.QueueUpdateCloudDataSampler(self.cloudImage, 0)
        pt.QueueUpdateCloudData(self.cloudProperties, 0)
        pt.UpdateDescriptorSets()
        sv.QueueUpdateCloudData(self.cloudProperties, 0)
        sv.QueueUpdateCloudDataSampler(self.cloudImage, 0)
        sv.UpdateDescriptorSets()
        self.UpdateShadowVolume(stream)

    # ---- shadow volume ----------------------------------------------------------------------------------
    def UpdateShadowVolume(self, stream=0):
        """main.cpp:223-264: SetLightDirection, SetOrigin, rebuild, rebind on PT, restart accumulation."""
        p = self.shadowVolumeProperties
        p.SetLightDirection(self.lightDirection)
        b0, b1 = self.cloudProperties.bounds[0], self.cloudProperties.bounds[1]
        p.SetOrigin((b0[0], b0[1], b0[2]), (b1[0], b1[1], b1[2]))
        self.shadowVolumeTechnique.QueueUpdateShadowVolume(p, 0)
        self.shadowVolumeTechnique.UpdateDescriptorSets()
        self.shadowVolumeTechnique.RecordDrawCommands(stream, 0)
        self.pathTracingTechnique.QueueUpdateShadowVolume(p, 0)
        self.pathTracingTechnique.UpdateDescriptorSets()
        self.reset_accumulation(stream)

    def reset_accumulation(self, stream=0):
        self.resultImage.clear(stream)
        self.pushConstants.frameCount = 0
        self._frames_rendered = 0

    # ---- parameters -------------------------------------------------------------------------------------
    def apply_parameters(self):
        """UI "Apply" path (main.cpp:557-595) for Parameters / CloudProperties / camera."""
        pt = self.pathTracingTechnique
        pt.QueueUpdateParameters(self.parameters, 0)
        pt.QueueUpdateCloudData(self.cloudProperties, 0)
        pt.QueueUpdateCameraProperties(self.cameraProperties, 0)
        pt.UpdateDescriptorSets()
        pt.SetExtensions(self.extensions)
        self.shadowVolumeTechnique.QueueUpdateCloudData(self.cloudProperties, 0)
        self.shadowVolumeTechnique.UpdateDescriptorSets()

    # ---- frames -----------------------------------------------------------------------------------------
    def next_seeds(self, n: int) -> np.ndarray:
        """The next n values of std::rand() (MSVC LCG) continuing from srand(0)."""
        total = msvc_rand_sequence(self._frames_rendered + n, 0)
        return total[self._frames_rendered:]

    def DrawFrame(self, stream=0):
        """One progressive frame exactly as RenderLoop does it (main.cpp:719-728)."""
        self.pushConstants.seed = int(self.next_seeds(1)[0])
        self.pathTracingTechnique.RecordDrawCommands(stream, 0)
        self.pushConstants.frameCount += 1
        self._frames_rendered += 1

    def render(self, spp: int, stream=0, accum_mode: int = capi.ACCUM_RUNNING_MEAN, seeds: Optional[np.ndarray] = None):
        """spp progressive frames through the wavefront fast path (== spp DrawFrame calls)."""
        if seeds is None:
            seeds = self.next_seeds(spp)
        self.pathTracingTechnique.RenderBatch(seeds, self.pushConstants.frameCount, accum_mode, stream, 0)
        self.pushConstants.frameCount += len(seeds)
        self._frames_rendered += len(seeds)

    def close(self):
        self.pathTracingTechnique.close()
        self.shadowVolumeTechnique.close()
        self.resultImage.close()
        self.shadowVolumeImage.close()
        if self.cloudImage is not None:
            self.cloudImage.close()


def make_renderer(ctx: Context, w: Workload, cloud_layout="linear", sv_layout=None, n_cloud: Optional[int] = None,
                  width: Optional[int] = None, height: Optional[int] = None, n_sv: Optional[int] = None,
                  math_mode: int = capi.PT_MATH_STRICT) -> HeadlessRenderer:
    """Renderer configured for a BASELINE workload (optionally shrunk for tests); the cloud is NOT set yet."""
    r = HeadlessRenderer(ctx, width or w.width, height or w.height, n_sv or w.n_sv, cloud_layout, sv_layout, w.fov)
    r.cloudProperties.densityScaling = w.density_scaling
    if w.phase_g is not None:
        r.parameters.SetPhaseG(w.phase_g)
    r.extensions.absorption = w.absorption
    r.extensions.math_mode = math_mode
    r.lightDirection = w.light_dir
    r.pathTracingTechnique.QueueUpdateParameters(r.parameters, 0)
    r.pathTracingTechnique.UpdateDescriptorSets()
    r.pathTracingTechnique.SetExtensions(r.extensions)
    return r
