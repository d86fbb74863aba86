#!/usr/bin/env python
"""Run the BASELINE.json configurations other than the bench headline on ONE GPU and print one JSON line each.

  python tools/run_configs.py [--configs c1_iso,c1_hg,c3,c4,c5] [--c3-spp 64] [--c5-frames 8]

  c1_*  128^3 / 512x512 / 16 spp, reference-pure semantics: GPU time + full-frame parity against the CPU oracle
  c3    1024^3 / 3840x2160 (spp reduced on one GPU; the 1024-spp split is what tools/scaling.sh --config c3 runs)
  c4    512^3 dense (densityScaling 200, albedo 0.99) / 1080p / 64 spp: long-path stress, steps and collisions per path
  c5    1024^3, sun animated per frame: shadow-volume rebuild + 64 spp 1080p per animated frame
Timing: CUDA events on the default stream, one warm-up pass.  The oracle is used as the checker for c1 only.
"""
import argparse
import json
import math
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cloudrendering_vulkan_b200 as crv  # noqa: E402


MATH_MODE = crv.PT_MATH_DRIVER


def timed(fn):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1)


def run_plain(ctx, name, spp=None, parity=False, row_parity=0):
    import dataclasses
    wl = crv.CONFIGS[name]
    if spp:
        wl = dataclasses.replace(wl, spp=spp)
    stream = torch.cuda.current_stream().cuda_stream
    r = crv.make_renderer(ctx, wl, cloud_layout="brick", math_mode=MATH_MODE)
    r.set_cloud_synthetic(wl.n_cloud, wl.fbm_seed, stream)
    seeds = crv.msvc_rand_sequence(wl.spp)
    pt = r.pathTracingTechnique

    def render():
        r.resultImage.clear(stream)
        pt.RenderBatch(seeds, 0, crv.ACCUM_RUNNING_MEAN, stream, 0)

    render()
    pt.ResetStats(stream)
    ms = timed(render)
    st = pt.GetStats(stream)
    sv = []
    for i in range(6):
        r.UpdateShadowVolume(stream)
        if i >= 2:
            sv.append(r.shadowVolumeTechnique.last_build_ms())
    out = {"config": name, "n_cloud": wl.n_cloud, "resolution": [wl.width, wl.height], "spp": wl.spp, "n_sv": wl.n_sv,
           "density_scaling": wl.density_scaling, "absorption": wl.absorption, "g": wl.phase_g,
           "render_ms": ms, "mpath_samples_per_s": wl.width * wl.height * wl.spp / ms / 1e3,
           "steps_per_path": st.tracking_steps / st.paths, "collisions_per_path": st.collisions / st.paths,
           "fetch_fraction": st.density_fetches / max(1, st.tracking_steps),
           "algorithmic_gbs": (32.0 * (st.tracking_steps + st.collisions) + 32.0 * st.paths) / ms / 1e6,
           "sv_rebuild_ms": float(np.median(sv)), "sv_algorithmic_gbs": (4.0 * wl.n_sv ** 3 + 4.0 * wl.n_cloud ** 3) / float(np.median(sv)) / 1e6}
    if parity:
        from tests import oracle_binding as orc
        cloud = r.cloudImage.download(stream)
        sv_gpu = r.shadowVolumeImage.download(stream)
        t0 = time.perf_counter()
        sv_cpu = orc.shadow_volume(cloud, r.cloudProperties, r.shadowVolumeProperties)
        sv_cpu_s = time.perf_counter() - t0
        big = sv_cpu >= 1e-30
        out["sv_max_rel_err"] = float((np.abs(sv_gpu[big] - sv_cpu[big]) / sv_cpu[big]).max())
        out["sv_cpu_s"] = sv_cpu_s
        # every frame of the 16-spp render, one at a time, against the oracle: counters + radiance
        pt.SetDebugCounters(True)
        div = tot = 0
        max_err = 0.0
        cpu_s = 0.0
        mean_gpu = np.zeros((wl.height, wl.width, 4), np.float64)
        mean_cpu = np.zeros_like(mean_gpu)
        for f in range(wl.spp):
            r.reset_accumulation(stream)
            r.pushConstants.seed = int(seeds[f])
            pt.RecordDrawCommands(stream, 0)
            g = r.resultImage.download(stream)
            gs, gc = pt.DownloadDebugCounters(stream)
            pc = crv.PushConstants()
            pc.seed, pc.frameCount = int(seeds[f]), 0
            c_img = np.zeros((wl.height, wl.width, 4), np.float32)
            t0 = time.perf_counter()
            _, cs, cc = orc.path_trace_frame(cloud, sv_cpu, r.cameraProperties, r.cloudProperties, r.parameters,
                                             r.shadowVolumeProperties, pc, wl.absorption, c_img)
            cpu_s += time.perf_counter() - t0
            same = (gs == cs) & (gc == cc)
            err = np.abs(g - c_img) / np.maximum(np.abs(c_img), 1e-3)
            div += int((~same).sum())
            tot += same.size
            max_err = max(max_err, float(err[same].max()))
            mean_gpu += g
            mean_cpu += c_img
        out.update({"parity_paths": tot, "threshold_divergent": div, "max_rel_err_matching": max_err,
                    "converged_max_abs_diff": float(np.abs(mean_gpu - mean_cpu).max() / wl.spp),
                    "cpu_oracle_mpath_samples_per_s": tot / cpu_s / 1e6, "cpu_threads": orc.max_threads()})
    elif row_parity:
        # full-size configs: a strided sample of image rows of one frame against the oracle (the shadow volume is the
        # GPU's -- its own parity is the c1 / test-suite check -- so this isolates the tracker at full size)
        from tests import oracle_binding as orc
        cloud = r.cloudImage.download(stream)
        sv_gpu = r.shadowVolumeImage.download(stream)
        pt.SetDebugCounters(True)
        r.reset_accumulation(stream)
        r.pushConstants.seed = int(seeds[1])
        pt.RecordDrawCommands(stream, 0)
        g = r.resultImage.download(stream)
        gs, gc = pt.DownloadDebugCounters(stream)
        pc = crv.PushConstants()
        pc.seed, pc.frameCount = int(seeds[1]), 0
        rows = list(range(3, wl.height, max(1, wl.height // row_parity)))
        div = tot = 0
        max_err = 0.0
        for y in rows:
            c_img = np.zeros((wl.height, wl.width, 4), np.float32)
            _, cs, cc = orc.path_trace_frame(cloud, sv_gpu, r.cameraProperties, r.cloudProperties, r.parameters,
                                             r.shadowVolumeProperties, pc, wl.absorption, c_img, rect=(0, y, wl.width, y + 1))
            same = (gs[y] == cs[y]) & (gc[y] == cc[y])
            err = np.abs(g[y] - c_img[y]) / np.maximum(np.abs(c_img[y]), 1e-3)
            div += int((~same).sum())
            tot += wl.width
            if same.any():
                max_err = max(max_err, float(err[same].max()))
        out.update({"parity_paths": tot, "parity_rows": len(rows), "threshold_divergent": div, "max_rel_err_matching": max_err})
    r.close()
    return out


def run_c5(ctx, frames):
    wl = crv.CONFIGS["c5"]
    stream = torch.cuda.current_stream().cuda_stream
    r = crv.make_renderer(ctx, wl, cloud_layout="brick", math_mode=MATH_MODE)
    r.set_cloud_synthetic(wl.n_cloud, wl.fbm_seed, stream)
    seeds = crv.msvc_rand_sequence(wl.spp)
    pt = r.pathTracingTechnique
    rows = []
    for k in range(-1, frames):            # k = -1: warm-up
        th = 2.0 * math.pi * max(k, 0) / 64.0
        r.lightDirection = (math.cos(th), -1.0, math.sin(th))

        def frame():
            r.UpdateShadowVolume(stream)          # SetLightDirection + SetOrigin + rebuild + rebind + clear
            pt.RenderBatch(seeds, 0, crv.ACCUM_RUNNING_MEAN, stream, 0)

        ms = timed(frame)
        if k >= 0:
            rows.append((ms, r.shadowVolumeTechnique.last_build_ms(), pt.LastRenderMs()[0]))
    a = np.array(rows)
    out = {"config": "c5", "n_cloud": wl.n_cloud, "n_sv": wl.n_sv, "resolution": [wl.width, wl.height], "spp_per_frame": wl.spp,
           "animated_frames": frames, "frame_ms_mean": float(a[:, 0].mean()), "frame_ms_max": float(a[:, 0].max()),
           "sv_rebuild_ms_mean": float(a[:, 1].mean()), "sv_rebuild_ms_max": float(a[:, 1].max()),
           "sv_algorithmic_gbs": (4.0 * wl.n_sv ** 3 + 4.0 * wl.n_cloud ** 3) / float(a[:, 1].mean()) / 1e6,
           "tracker_ms_mean": float(a[:, 2].mean()),
           "mpath_samples_per_s_end_to_end": wl.width * wl.height * wl.spp / float(a[:, 0].mean()) / 1e3}
    r.close()
    return out


def run_ppm(ctx, groups_list=(200, 2000, 20000)):
    """the photon-tracing dispatch of RenderTechniquePPM on the c2 scene (512^3 cloud, default 100^3 photon map,
    densityScaling 200 -- the reference default): the reference's 200 groups of 64 photons, and larger dispatches."""
    n = 512
    vol = crv.Volume(ctx, n, n, n, layout="brick")
    vol.fbm_generate(1234)
    cp = crv.CloudProperties()
    cp.SetFromGrid((n, n, n), (1.0 / n,) * 3, vol.majorant())
    par = crv.Parameters()
    par.SetPhaseG(0.85)
    pm = crv.PhotonMapProperties()
    pm.SetBounds(cp.bounds)
    pc = crv.PushConstants()
    pc.seed = 38
    ppm = crv.RenderTechniquePPM(ctx, pm, pc)
    ppm.AllocateResources()
    ppm.QueueUpdateCloudData(cp)
    ppm.QueueUpdateCloudDataSampler(vol)
    ppm.QueueUpdateParameters(par)
    ppm.UpdateDescriptorSets()
    rows = []
    for g in groups_list:
        ms = []
        for i in range(4):
            ppm.RecordPhotonTracing(0, g)
            st = ppm.LastTraceStats()
            if i:
                ms.append(st["ms"])
        rows.append({"groups_x": g, "photons": 64 * g, "ms_median": float(np.median(ms)), "mphotons_per_s": 64 * g / float(np.median(ms)) / 1e3,
                     "tracking_steps": st["tracking_steps"], "deposits": st["deposits"], "dropped_cell_full": st["dropped"]})
    # the whole technique at 1080p: RecordDrawCommands = UpdateRadius + tracing (200 groups) + estimate, a few progressive frames
    wl = crv.CONFIGS["c2"]
    cam = crv.CameraProperties()
    cam.SetResolution(wl.width, wl.height)
    cam.SetFOV(wl.fov)
    sp = crv.ShadowVolumeProperties()
    sp.voxelAxisCount = wl.n_sv
    sp.SetLightDirection(wl.light_dir)
    sp.SetOrigin(tuple(cp.bounds[0])[:3], tuple(cp.bounds[1])[:3])
    svol = crv.Volume(ctx, wl.n_sv, wl.n_sv, wl.n_sv, layout="brick")
    svt = crv.RenderTechniqueSV(ctx, sp, pc)
    svt.SetFrameReferences([svol])
    svt.QueueUpdateCloudData(cp, 0)
    svt.QueueUpdateCloudDataSampler(vol, 0)
    svt.QueueUpdateShadowVolume(sp, 0)
    svt.UpdateDescriptorSets()
    svt.RecordDrawCommands(0, 0)
    img = crv.Image(ctx, wl.width, wl.height)
    ppm.SetFrameReferences([img])
    ppm.QueueUpdateCameraProperties(cam)
    ppm.QueueUpdateShadowVolume(sp)
    ppm.QueueUpdateShadowVolumeSampler(svol)
    ppm.UpdateDescriptorSets()
    seeds = crv.msvc_rand_sequence(4)
    frame_ms, est_ms, march = [], [], 0
    for f in range(4):
        pc.seed, pc.frameCount = int(seeds[f]), f
        frame_ms.append(timed(lambda: ppm.RecordDrawCommands(0, 0)))
        e = ppm.LastEstimateStats()
        est_ms.append(e["ms"])
        march = e["march_steps"]
    frame = {"resolution": [wl.width, wl.height], "frame_ms": float(np.median(frame_ms[1:])), "estimate_ms": float(np.median(est_ms[1:])),
             "march_steps_per_frame": march, "gmarch_steps_per_s": march / float(np.median(est_ms[1:])) / 1e6,
             "finite": bool(np.isfinite(img.download()).all())}
    svt.close()
    svol.close()
    img.close()
    ppm.close()
    vol.close()
    return {"config": "ppm_photon_tracing", "full_frame_1080p": frame, "n_cloud": n, "photon_map_cells": pm.GetTotalSize(), "density_scaling": cp.densityScaling,
            "g": 0.85, "dispatches": rows,
            "note": "includes the per-dispatch clear of the previously touched cells (the reference zero-fills 1.5 GB + 16 MB) and the "
                    "q-table refresh; one thread per photon like PPM_PT.comp, latency of the longest photon path dominates small dispatches"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="c1_iso,c1_hg,c4,c3,c5")
    ap.add_argument("--c3-spp", type=int, default=64)
    ap.add_argument("--c5-frames", type=int, default=8)
    ap.add_argument("--math", default="driver", choices=["strict", "driver"])
    args = ap.parse_args()
    global MATH_MODE
    MATH_MODE = crv.PT_MATH_DRIVER if args.math == "driver" else crv.PT_MATH_STRICT
    torch.cuda.set_device(0)
    ctx = crv.Context(0)
    for name in args.configs.split(","):
        if name == "ppm":
            out = run_ppm(ctx)
        elif name == "c5":
            out = run_c5(ctx, args.c5_frames)
        elif name == "c3":
            out = run_plain(ctx, "c3", spp=args.c3_spp, row_parity=24)
        else:
            out = run_plain(ctx, name, parity=name.startswith("c1"), row_parity=0 if name.startswith("c1") else 48)
        out["math"] = args.math
        print(json.dumps(out), flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
