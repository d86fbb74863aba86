"""The C++ host side (include/crv/render_technique.h + crv_headless, the reference's main.cpp role) over the C ABI."""
import json
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "cloudrendering_vulkan_b200", "bin", "crv_headless")


def _run(args):
    return subprocess.run([BIN] + [str(a) for a in args], capture_output=True, text=True, timeout=300)


def _fnv1a(buf: bytes) -> str:
    h = 1469598103934665603
    for b in buf:
        h = ((h ^ b) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return f"{h:016x}"


def test_cpp_driver_is_built_and_fails_loudly_without_a_gpu(crv):
    assert os.path.exists(BIN), "crv_headless was not built (make -C cloudrendering_vulkan_b200/csrc)"
    if crv.capi.lib.crv_device_count() > 0:
        pytest.skip("a CUDA device is present")
    r = _run(["--fbm", 16, "--width", 32, "--height", 32, "--nsv", 16])
    assert r.returncode == 3 and "no CPU fallback" in r.stderr


def test_cpp_driver_rejects_bad_arguments():
    assert _run(["--no-such-flag"]).returncode == 2


@pytest.mark.gpu
def test_cpp_driver_matches_the_python_host_bitwise(crv, ctx):
    """same scene through the C++ RenderTechnique classes (batch and frame-by-frame) and through the Python mirror:
    identical image bits -- both are thin hosts over the same C ABI"""
    n, w, h, nsv, spp = 40, 96, 64, 56, 5
    args = ["--fbm", n, "--width", w, "--height", h, "--nsv", nsv, "--spp", spp, "--g", 0.85, "--absorption", 0.1]
    out = {}
    for mode in ([], ["--per-frame"]):
        r = _run(args + mode)
        assert r.returncode == 0, r.stderr
        out[bool(mode)] = json.loads(r.stdout)
    assert out[False]["image_fnv1a"] == out[True]["image_fnv1a"]
    assert out[False]["paths"] == w * h * spp and out[False]["collisions"] > 0
    wl = crv.Workload("t", n, w, h, spp, nsv, 50.0, 0.1, 0.85)
    r = crv.make_renderer(ctx, wl, cloud_layout="brick")
    r.set_cloud_synthetic(n, 1234)
    seeds = crv.msvc_rand_sequence(spp)
    r.pathTracingTechnique.RenderBatch(seeds, 0, crv.ACCUM_RUNNING_MEAN, 0, 0)
    img = np.ascontiguousarray(r.resultImage.download(), np.float32)
    assert _fnv1a(img.tobytes()) == out[False]["image_fnv1a"]
    r.close()


@pytest.mark.gpu
@pytest.mark.parametrize("gpus,same_device,spp", [(2, True, 7), (3, True, 7), (2, True, 8), (2, False, 7), (4, False, 8)])
def test_cpp_driver_sample_split(crv, gpus, same_device, spp):
    """crv_headless --gpus N: crv::DeviceGroup (crv_group_*) from C++ -- one replica per member, frames f mod N, one combine
    kernel per member.  Every member must hold the same picture, equal to the 1-GPU picture within summation order (1e-6)."""
    if not same_device and crv.capi.lib.crv_device_count() < gpus:
        pytest.skip(f"needs {gpus} GPUs")
    n, w, h, nsv = 40, 96, 64, 56          # spp 7: ragged split, whole images combined; spp 8: even split, tracked pixels only
    args = ["--fbm", n, "--width", w, "--height", h, "--nsv", nsv, "--spp", spp, "--g", 0.85, "--absorption", 0.1]
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        one, many = os.path.join(d, "one.pfm"), os.path.join(d, "many.pfm")
        r1 = _run(args + ["--pfm", one])
        assert r1.returncode == 0, r1.stderr
        rn = _run(args + ["--pfm", many, "--gpus", gpus] + (["--same-device"] if same_device else []))
        assert rn.returncode == 0, rn.stderr
        o1, on = json.loads(r1.stdout), json.loads(rn.stdout)
        assert on["gpus"] == gpus and on["members_identical"] is True and on["mode"] == "split"
        assert on["paths"] == o1["paths"] == w * h * spp and on["collisions"] == o1["collisions"]

        def pfm(path):
            with open(path, "rb") as f:
                assert f.readline().strip() == b"PF"
                ww, hh = map(int, f.readline().split())
                f.readline()
                return np.frombuffer(f.read(), np.float32).reshape(hh, ww, 3)
        a, b = pfm(one), pfm(many)
        assert a.max() > 0.1
        assert (np.abs(a - b) / np.maximum(np.abs(a), 1e-3)).max() <= 1e-5      # running mean (1 GPU) vs sum / spp (split)


@pytest.mark.gpu
def test_cpp_driver_photon_mapping_frames_match_the_python_host(crv, ctx):
    """--ppm N: crv::RenderTechniquePPM (include/crv/render_technique.h) driven like main.cpp:204-207/813 over the scene
    crv_headless already holds.  Photon slots are claimed with atomics, so two runs are not bit-identical (PPM_PT.comp's own
    property); the frame mean, the radius schedule (RenderTechniquePPM.cpp:302-312) and finiteness are what is compared
    against the Python mirror of the same class."""
    n, w, h, nsv, frames = 40, 96, 64, 56, 4
    r = _run(["--fbm", n, "--width", w, "--height", h, "--nsv", nsv, "--spp", 1, "--g", 0.85, "--absorption", 0.1, "--ppm", frames])
    assert r.returncode == 0, r.stderr
    out = json.loads(r.stdout)
    assert out["ppm_frames"] == frames and out["ppm_frame_ms"] > 0 and np.isfinite(out["ppm_image_mean"]) and out["ppm_image_mean"] > 0
    wl = crv.Workload("t", n, w, h, 1, nsv, 50.0, 0.1, 0.85)
    rr = crv.make_renderer(ctx, wl, cloud_layout="brick")
    rr.set_cloud_synthetic(n, 1234)
    pm = crv.PhotonMapProperties()
    pm.SetBounds(rr.cloudProperties.bounds)
    for i in range(4):
        pm.lightDirection[i] = rr.shadowVolumeProperties.lightDirection[i]
    pc = crv.PushConstants()
    img = crv.Image(ctx, w, h)
    ppm = crv.RenderTechniquePPM(ctx, pm, pc, 10.0)
    ppm.AllocateResources()
    ppm.SetFrameReferences([img])
    ppm.QueueUpdateCloudData(rr.cloudProperties)
    ppm.QueueUpdateCloudDataSampler(rr.cloudImage)
    ppm.QueueUpdateParameters(rr.parameters)
    ppm.QueueUpdateCameraProperties(rr.cameraProperties)
    ppm.QueueUpdateShadowVolume(rr.shadowVolumeProperties)
    ppm.QueueUpdateShadowVolumeSampler(rr.shadowVolumeImage)
    ppm.UpdateDescriptorSets()
    seeds = crv.msvc_rand_sequence(frames)
    for f in range(frames):
        pc.seed, pc.frameCount = int(seeds[f]), f
        ppm.RecordDrawCommands(0, 0)
    mean = float(img.download().astype(np.float64).mean())
    assert abs(pc.pmRadius - out["ppm_radius"]) <= 1e-6 * pc.pmRadius
    assert abs(mean - out["ppm_image_mean"]) <= 2e-3 * mean, (mean, out["ppm_image_mean"])
    ppm.close()
    img.close()
    rr.close()
