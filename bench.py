#!/usr/bin/env python
"""bench.py -- the contract benchmark of the PathTracer + ShadowVolume hot path.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config c1_iso|c1_hg|c2|c3|c4|c5] [--scaling strong|weak]
  torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W        (N > 1)

One "step" = one full progressive picture of the workload (all its samples per pixel) from inputs already resident in HBM:
clear, path-trace this GPU's frames into its fp32 SUM image, combine the GPUs' images and divide by the sample count
(crv_group_all_reduce_sum_scale: one kernel per GPU over NVLink peer memory; a plain scale at N = 1).  Default workload is
BASELINE.json's headline configuration `c2` (512^3 fBm cloud, 1920x1080, 256 spp, albedo 0.9, HG g = 0.85).
metric = Mpath-samples/s = W*H*spp_total / step time, whole job.

Scaling (N > 1): the density grid and the shadow volume are replicated, frame f of the picture's seed list goes to GPU
f mod N.  `--scaling strong` (default; north_star: "samples per pixel are split across GPUs"): the picture keeps its spp,
every GPU renders spp/N frames.  `--scaling weak`: every GPU renders the workload's spp (spp_total = N * spp).  The other
mode's figure is measured in the same run and reported under `scaling_other`.  c5 is the animated-sun loop: per step
ShadowVolume rebuild + 64-spp picture, `sv` reports the rebuild on its own.

Extra keys: `e2e` (same metric through the public API from HOST buffers: H2D grid upload + majorant / skip tables +
shadow-volume rebuild + render + combine + D2H of the picture inside the timed region), `roofline` (tracker kernel: the
HBM block SURVEY 8d defines, the measured-DRAM and instruction-issue fractions that actually bind, from an ncu pass stamped
with the kernel sources' hash), `sv` (shadow-volume rebuild ms, algorithmic and measured GB/s), `cpu_baseline` (the CPU side
on a bounded sample of the same workload), `parity`, `clocks`, `gpu_launches`.

--impl reference times the reference's CPU implementation of the same workload on all host cores (no product code in the
process): the reference's own shaders compiled by g++ (oracle/_ref) when the workload is expressible in them, else the
restatement (oracle/crv_oracle.cpp) -- the reference's SPIR-V/lavapipe path cannot exist in this image.
"""
from __future__ import annotations

import argparse
import dataclasses
import glob
import hashlib
import importlib.util
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "Mpath-samples/s"
PKG = os.path.join(ROOT, "cloudrendering_vulkan_b200")


def load_workloads():
    """cloudrendering_vulkan_b200/workloads.py by path: plain data, does not load libcrv_b200.so"""
    spec = importlib.util.spec_from_file_location("crv_workloads", os.path.join(PKG, "workloads.py"))
    m = importlib.util.module_from_spec(spec)
    sys.modules["crv_workloads"] = m
    spec.loader.exec_module(m)
    return m


def source_sha() -> str:
    """hash of the kernel sources: an ncu summary under profiles/ belongs to the build whose sources hash the same"""
    h = hashlib.sha256()
    for f in sorted(glob.glob(os.path.join(PKG, "csrc", "*.cu")) + glob.glob(os.path.join(PKG, "csrc", "*.cuh")) +
                    glob.glob(os.path.join(PKG, "csrc", "*.h")) + [os.path.join(PKG, "csrc", "Makefile")]):
        h.update(os.path.basename(f).encode())
        h.update(open(f, "rb").read())
    return h.hexdigest()[:16]


def _peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json, burst copy)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def _ncu_profile():
    """profiles/ncu_live.json: what the last ncu pass measured per kernel (tools/gpu_profile.sh), with the source hash"""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "ncu_live.json")))
    except Exception:
        return None


class ClockSampler(threading.Thread):
    """samples SM clock / throttle reasons through NVML while the timed region runs"""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index = index
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksThrottleReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake",
        }
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop_evt.wait(0.01)

    def stop(self):
        self._stop_evt.set()
        if self.is_alive():
            self.join(timeout=2)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def _workload_text(wl, spp_total, world):
    per = f"{spp_total} spp" if world == 1 else f"{spp_total} spp split over {world} GPUs"
    sun = ", sun direction animated per step (ShadowVolume rebuild inside the step)" if wl.name == "c5" else ""
    return (f"{wl.name}: {wl.n_cloud}^3 fBm cloud, {wl.width}x{wl.height}, {per}, densityScaling {wl.density_scaling:g}, "
            f"absorption {wl.absorption:g}, g {wl.phase_g if wl.phase_g is not None else 'isotropic'}, N_sv {wl.n_sv}{sun}")


def _config_dict(wl, args, world, spp_total, layout, math):
    return {"workload": _workload_text(wl, spp_total, world), "spp_total": spp_total,
            "parallelism": f"frames f mod {world} per GPU, grid + shadow volume replicated, the fold pushes the tracked sums into every member over NVLink, then one wait + one local kernel per GPU"
                           if world > 1 else "one GPU",
            "layout": layout, "math": math,
            "l2": "inputs larger than L2 (grid + shadow volume >= 1 GB vs 126 MB)" if wl.n_cloud >= 512 else
                  "grid + shadow volume 0.5 GB (N_sv 500) vs 126 MB of L2"}


# =============================================================================================================
# reference arm / cpu_baseline: the CPU side (the ONLY code here allowed to touch oracle/)
# =============================================================================================================
def _oracle_scene(wl, log):
    """Host-side scene for the oracle with NO product code: grid from the oracle's generator, parameter structs from
    the reference's own setters (oracle/_ref shim) + the restated SetCloudProperties, shadow volume from the oracle."""
    from tests import oracle_binding as orc
    t0 = time.perf_counter()
    cloud = orc.fbm_generate(wl.n_cloud, wl.fbm_seed)
    log(f"[reference] fBm {wl.n_cloud}^3 generated on {orc.max_threads()} threads in {time.perf_counter() - t0:.1f}s")
    cam, cp, par, sp = orc.reference_scene_structs(wl.n_cloud, wl.width, wl.height, wl.fov, wl.n_sv, wl.light_dir,
                                                   wl.density_scaling, wl.phase_g, float(cloud.max()))
    return orc, cloud, cam, cp, par, sp


def _cpu_kind(orc, wl):
    """"reference": the reference's OWN shaders compiled by g++ (oracle/_ref/libcrv_refshaders.so) -- possible when that
    library exists and the workload is expressible in the reference (no absorption: PathTracer.comp has none);
    "port": the hand restatement (oracle/crv_oracle.cpp), which carries the absorption extension, otherwise."""
    return "reference" if (orc.ref_shaders_lib() is not None and wl.absorption == 0.0) else "port"


def _vulkan_probe():
    """oracle/vulkan_probe.py: is the north star's SPIR-V / lavapipe oracle available here?  (never, in this image)"""
    try:
        spec = importlib.util.spec_from_file_location("vulkan_probe", os.path.join(ROOT, "oracle", "vulkan_probe.py"))
        m = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(m)
        r = m.probe()
        return "available" if r["available"] else r["reason"]
    except Exception as e:       # the probe is informational
        return f"probe failed: {e}"


def _cpu_note(kind, wl):
    if kind == "reference":
        return ("the reference's own PathTracer.comp compiled by g++ through oracle/glsl_compat.h (OpenMP over image rows); "
                "SPIR-V on lavapipe: " + _vulkan_probe())
    note = "CPU restatement of the shaders (oracle/crv_oracle.cpp, bit-identical to the compiled shaders on the golden vectors)"
    if wl.absorption > 0:
        note += (f"; the reference shader has no absorption term, so the workload's absorption {wl.absorption:g} (an extension of "
                 f"this repo) can only run in the restatement; `reference_shader_value` is the reference's own shader on the "
                 f"same scene at albedo 1")
    return note + "; SPIR-V on lavapipe: " + _vulkan_probe()


def _oracle_rows_step(orc, cloud, sv, cam, cp, par, sp, wl, seed, rows, kind="port"):
    """every `rows`-th image row of one frame; returns (paths, seconds)"""
    pc = orc.push_constants_bytes(seed, 0)
    if kind == "reference":
        w, h = orc.image_size(cam)
        scratch = np.zeros((h, w, 4), np.float32)
        t0 = time.perf_counter()
        orc.refsh_path_trace(cloud, sv, cam, cp, par, sp, pc, scratch, row_step=rows, want_counters=False)
        return w * len(range(0, h, int(rows))), time.perf_counter() - t0
    t0 = time.perf_counter()
    paths, _, _ = orc.path_trace_rows(cloud, sv, cam, cp, par, sp, pc, wl.absorption, rows)
    return paths, time.perf_counter() - t0


def _oracle_plan(orc, cloud, sv, cam, cp, par, sp, wl, seed, target_s, kind="port"):
    """size a bounded sample of the workload worth ~target_s seconds of CPU time: (frames, row_step)"""
    p, s = _oracle_rows_step(orc, cloud, sv, cam, cp, par, sp, wl, seed, 16, kind)
    frame_s = s * (wl.width * wl.height) / p
    if frame_s >= target_s:
        return 1, max(1, int(round(frame_s / target_s)))
    return max(1, int(round(target_s / frame_s))), 1


def _oracle_sample(orc, cloud, sv, cam, cp, par, sp, wl, seeds, frames, rows, kind="port"):
    """`frames` frames (seeds[0..frames)), every `rows`-th row of each: (paths, seconds)"""
    paths, secs = 0, 0.0
    for f in range(frames):
        p, s = _oracle_rows_step(orc, cloud, sv, cam, cp, par, sp, wl, seeds[f % len(seeds)], rows, kind)
        paths += p
        secs += s
    return paths, secs


def _sample_text(wl, frames, rows, paths, secs):
    what = "all rows" if rows == 1 else f"every {rows}th image row"
    return f"{what} of {frames} {wl.width}x{wl.height} frame(s) of this workload ({paths} path samples, {secs:.1f} s)"


def _use_all_host_threads(log=None):
    """torchrun exports OMP_NUM_THREADS=1 to every rank; the CPU arm is meant to use all the host threads it can, so the
    OpenMP runtime the oracle libraries share is told so explicitly"""
    try:
        import ctypes
        n = len(os.sched_getaffinity(0))
        ctypes.CDLL("libgomp.so.1").omp_set_num_threads(n)
        if log:
            log(f"[reference] OpenMP threads: {n} (OMP_NUM_THREADS was {os.environ.get('OMP_NUM_THREADS', 'unset')})")
    except Exception as e:      # keep whatever the environment gave
        if log:
            log(f"[reference] could not raise the OpenMP thread count: {e}")


def run_reference(args, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    log = lambda s: print(s, file=sys.stderr, flush=True)
    _use_all_host_threads(log)
    orc, cloud, cam, cp, par, sp = _oracle_scene(wl, log)
    kind = _cpu_kind(orc, wl)
    have_shaders = orc.ref_shaders_lib() is not None
    t0 = time.perf_counter()
    sv = orc.refsh_shadow_volume(cloud, cp, sp) if have_shaders else orc.shadow_volume(cloud, cp, sp)
    sv_s = time.perf_counter() - t0
    log(f"[reference] shadow volume {wl.n_sv}^3 on the CPU ({'reference shader' if have_shaders else 'port'}): {sv_s:.1f}s")
    seeds = orc.msvc_rand(256)
    # each step is a bounded sample of ~3 s of CPU work on all host cores
    frames, rows = _oracle_plan(orc, cloud, sv, cam, cp, par, sp, wl, seeds[0], 3.0, kind)
    for i in range(args.warmup):
        _oracle_sample(orc, cloud, sv, cam, cp, par, sp, wl, seeds, 1, max(rows, 4), kind)
    paths, secs = 0, 0.0
    for i in range(args.steps):
        p, s = _oracle_sample(orc, cloud, sv, cam, cp, par, sp, wl, np.roll(seeds, -i * frames), frames, rows, kind)
        paths += p
        secs += s
    value = paths / secs / 1e6
    sample = "per step: " + _sample_text(wl, frames, rows, paths // args.steps, secs / args.steps)
    shader_value = None
    if kind == "port" and have_shaders:     # alongside: the reference's own shader on one step's sample (albedo 1: it has no absorption)
        p, s = _oracle_sample(orc, cloud, sv, cam, cp, par, sp, wl, seeds, frames, rows, "reference")
        shader_value = p / s / 1e6
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "Mpath-samples/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * secs / args.steps, "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": _config_dict(wl, args, max(1, args.gpus), wl.spp * (max(1, args.gpus) if args.scaling == "weak" else 1), args.layout, args.math),
        "reference_arm": {"runs_on": "host cores", "math": "libm, no FMA contraction", "note": "`config` is the product arm's, key for key: the same workload, picture and sample budget"},
        "cpu_baseline": {"value": value, "unit": "Mpath-samples/s", "cores": orc.max_threads(), "kind": kind,
                         "sample": sample, "sv_build_s": sv_s, "reference_shader_value": shader_value, "note": _cpu_note(kind, wl)},
        "e2e": {"value": value, "unit": "Mpath-samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    _emit(line)
    return 0


# =============================================================================================================
# our arm
# =============================================================================================================
def run_ours(args, wl):
    import torch

    import cloudrendering_vulkan_b200 as crv
    from cloudrendering_vulkan_b200.multigpu import SampleSplit, frames_for_rank, init_group_nccl

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    stream_obj = torch.cuda.Stream(device=dev)          # the stream everything is launched and timed on
    stream = stream_obj.cuda_stream
    log = (lambda s: print(s, file=sys.stderr, flush=True)) if rank == 0 else (lambda s: None)

    ctx = crv.Context(local)
    W, H = wl.width, wl.height
    animated = wl.name == "c5"
    math_mode = crv.PT_MATH_DRIVER if args.math == "driver" else crv.PT_MATH_STRICT
    r = crv.make_renderer(ctx, wl, cloud_layout=args.layout, math_mode=math_mode)
    own_image = r.resultImage
    stage = 4 * wl.n_cloud ** 3 if (world > 1 and not args.no_e2e) else 0      # e2e at N > 1: the grid is gathered through the group
    split = SampleSplit(r, rank, world, nbuf=2, staging_bytes=stage)           # collective at N > 1: exchanges the members' memory handles
    own_image.close()
    img = split.image
    pt = r.pathTracingTechnique
    pt.SetJobOrder(args.job_order)
    if args.track_threshold or args.ctas_per_sm or args.steps_per_decision or args.sample_buffer_mib:
        pt.SetTuning(ctas_per_sm=args.ctas_per_sm, track_threshold=args.track_threshold,
                     steps_per_decision=args.steps_per_decision, sample_buffer_bytes=args.sample_buffer_mib << 20)
    pt.SetScheduler(args.scheduler)
    pt.SetSkipLevel(args.skip_level)
    r.set_cloud_synthetic(wl.n_cloud, wl.fbm_seed, stream)
    ctx.synchronize(stream)
    # shadow-volume rebuild: 3 warm rebuilds, then the median of 7 (CUDA events on the launching stream, around the kernel)
    sv_times = []
    for i in range(10):
        r.UpdateShadowVolume(stream)
        if i >= 3:
            sv_times.append(r.shadowVolumeTechnique.last_build_ms())
    sv_ms = float(np.median(sv_times))

    def seeds_for(scaling):
        spp_total = wl.spp * (world if scaling == "weak" else 1)
        return crv.msvc_rand_sequence(spp_total), spp_total

    def light_of(k):                                      # c5: L_k = normalize(cos th_k, -1, sin th_k), th_k = 2 pi k / 64
        th = 2.0 * np.pi * (k % 64) / 64.0
        return (float(np.cos(th)), -1.0, float(np.sin(th)))

    step_no = [0]

    def make_step(all_seeds, backend=None):
        backend = backend or args.combine
        def step():
            if animated:
                r.lightDirection = light_of(step_no[0])
                step_no[0] += 1
                r.UpdateShadowVolume(stream)              # SetLightDirection + SetOrigin + SV rebuild + rebind (main.cpp:223-264)
            split.render_picture(all_seeds, stream, backend)
        return step

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(step, steps, warmup):
        for _ in range(warmup):
            step()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record(stream_obj)
        for _ in range(steps):
            step()
        e1.record(stream_obj)
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms / steps

    # ---- the timed region: K steps, max over ranks, CUDA events on the launching stream ----
    all_seeds, spp_total = seeds_for(args.scaling)
    step = make_step(all_seeds)
    for _ in range(args.warmup):
        step()
    barrier()
    pt.ResetStats(stream)
    launches0 = ctx.launch_count
    sampler = ClockSampler(local)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record(stream_obj)
    for _ in range(args.steps):
        step()
    ev1.record(stream_obj)
    barrier()
    dev_ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop()
    split.group.status(stream)
    launches = ctx.launch_count - launches0
    st = pt.GetStats(stream)
    k_ms, k_n = pt.LastRenderMs()          # tracker launches of the LAST step (events on the launching stream)
    reduce_ms = split.group.last_reduce_ms()
    if world > 1:
        t = torch.tensor([dev_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms = float(t.item())
        cnt = torch.tensor([st.paths, st.tracking_steps, st.collisions, st.density_fetches, st.tracked_paths, launches],
                           dtype=torch.float64, device=dev)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
        tot_paths, tot_steps, tot_coll, tot_fetch, tot_tracked, launches = [float(x) for x in cnt.tolist()]
    else:
        tot_paths, tot_steps, tot_coll, tot_fetch, tot_tracked = (float(st.paths), float(st.tracking_steps), float(st.collisions),
                                                                  float(st.density_fetches), float(st.tracked_paths))
    ms_per_step = dev_ms / args.steps
    paths_per_step = W * H * spp_total
    value = paths_per_step / (ms_per_step * 1e-3) / 1e6
    assert abs(tot_paths - paths_per_step * args.steps) < 0.5, "kernel path counter disagrees with the workload size"
    final_picture = img.download(stream).copy() if rank == 0 else None

    # ---- N > 1: the combined picture against the same seed list rendered by rank 0 alone (<= 1e-6) ----
    multi_gpu_parity = None
    if world > 1 and not animated:
        if rank == 0:
            solo = crv.Image(ctx, W, H)
            pt.SetFrameReferences([solo])
            pt.RenderBatch(all_seeds, 0, crv.ACCUM_SUM, stream, 0)
            solo.scale(1.0 / spp_total, stream)
            ref = solo.download(stream)
            pt.SetFrameReferences([img])
            solo.close()
            err = np.abs(final_picture - ref) / np.maximum(np.abs(ref), 1e-3)
            multi_gpu_parity = {"max_rel_err_vs_one_gpu": float(err.max()), "tolerance": 1e-6, "pixels": int(err.shape[0] * err.shape[1]),
                                "what": f"picture combined over {world} GPUs vs the same {spp_total} frames summed on rank 0 alone"}
            assert err.max() <= 1e-6, f"combined picture differs from the one-GPU picture: {err.max():.3e}"
        barrier()

    # ---- the other scaling mode, the other math mode, the drop-in accumulation mode: short extra measurements ----
    scaling_other = None
    if world > 1 and not args.quick:
        other = "weak" if args.scaling == "strong" else "strong"
        o_seeds, o_total = seeds_for(other)
        o_ms = timed(make_step(o_seeds), max(3, args.steps), 2)
        scaling_other = {"scaling": other, "spp_total": o_total, "ms_per_step": o_ms, "value": W * H * o_total / (o_ms * 1e-3) / 1e6,
                         "unit": "Mpath-samples/s"}
    reduce_cmp = {"timed_path": args.combine, "push_ms" if args.combine == "push" else "timed_ms": reduce_ms,
                  "note": "device time of the combine on rank 0, waiting for the slowest member included.  push (the timed path): "
                          "crv_group_push_finish -- the fold has already stored the tracked sums into every member's inbox; one wait, one "
                          "local kernel; peer_kernel: crv_group_all_reduce_lists_sum_scale, two barriers around a band reduce over peer "
                          "memory (tracked pixels only); peer_dense: whole images through the same kernel; nccl: ncclAllReduce of the "
                          "whole image + scale kernel"}
    if world > 1 and not args.quick:
        reduce_cmp["peer_ms_per_step"] = timed(make_step(all_seeds, "peer"), 3, 2)
        reduce_cmp["peer_kernel_ms"] = split.group.last_reduce_ms()
        timed(make_step(all_seeds, "peer_dense"), 3, 2)
        reduce_cmp["peer_dense_kernel_ms"] = split.group.last_reduce_ms()
        try:
            init_group_nccl(split.group)
            timed(make_step(all_seeds, "nccl"), 3, 2)
            reduce_cmp["nccl_allreduce_plus_scale_ms"] = split.group.last_reduce_ms()
        except Exception as e:          # the comparison is informational
            reduce_cmp["nccl_allreduce_plus_scale_ms"] = None
            reduce_cmp["nccl_note"] = str(e)[:200]
    math_modes = {args.math: value}
    running_mean = None
    if world == 1 and not args.quick:
        other = "strict" if args.math == "driver" else "driver"
        r.extensions.math_mode = crv.PT_MATH_STRICT if other == "strict" else crv.PT_MATH_DRIVER
        pt.SetExtensions(r.extensions)
        math_modes[other] = paths_per_step / (timed(step, args.steps, 2) * 1e-3) / 1e6
        r.extensions.math_mode = math_mode
        pt.SetExtensions(r.extensions)
        if not animated:
            # the reference's own accumulation, frame by frame in frame order (PathTracer.comp:420-424), as one batch:
            # bit-identical to spp RecordDrawCommands calls
            def mean_step():
                img.clear(stream)
                pt.RenderBatch(all_seeds, 0, crv.ACCUM_RUNNING_MEAN, stream, 0)
            rm_ms = timed(mean_step, args.steps, 2)
            rm = img.download(stream)
            err = np.abs(rm - final_picture) / np.maximum(np.abs(final_picture), 1e-3)
            running_mean = {"value": paths_per_step / (rm_ms * 1e-3) / 1e6, "ms_per_step": rm_ms, "unit": "Mpath-samples/s",
                            "max_rel_diff_vs_sum_over_spp": float(err.max()),
                            "note": "CRV_ACCUM_RUNNING_MEAN: the shader's (sample + old*fc)/(fc+1) per frame, bit-identical to "
                                    "frame-by-frame dispatches; `value` uses CRV_ACCUM_SUM + one scale, the multi-GPU split's mode"}

    # ---- census: the same picture with every reference step executed (skip level 0): the algorithmic work, and a
    #      bit-for-bit check that the proof-based skipping did not change the picture ----
    census = None
    if not animated and not args.quick:
        pt.SetSkipLevel(0)
        pt.ResetStats(stream)
        step()
        ctx.synchronize(stream)
        c_ms, _ = pt.LastRenderMs()
        cs = pt.GetStats(stream)
        full_picture = img.download(stream)
        pt.SetSkipLevel(args.skip_level)
        if world > 1:
            cnt = torch.tensor([cs.paths, cs.tracking_steps, cs.collisions], dtype=torch.float64, device=dev)
            dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
            c_paths, c_steps, c_coll = [float(x) for x in cnt.tolist()]
        else:
            c_paths, c_steps, c_coll = float(cs.paths), float(cs.tracking_steps), float(cs.collisions)
        # (CRV_ACCUM_SUM adds a background pixel's n identical samples as n * sample, a tracked pixel's one by one: the two
        #  levels agree to summation order here; bit-identity of the levels is asserted on the reference accumulation, `parity`)
        c_err = float((np.abs(full_picture - final_picture) / np.maximum(np.abs(final_picture), 1e-3)).max()) if rank == 0 else 0.0
        census = {"steps_per_path": c_steps / c_paths, "collisions_per_path": c_coll / c_paths, "tracker_ms_at_skip_level_0": c_ms,
                  "picture_max_rel_diff_vs_skip_level_1": c_err}
        assert c_err <= 2e-6, f"skip level 1 changed the picture: {c_err:.3e}"
        step()                                # leave the tracker's cached classification at the benchmark's level
        ctx.synchronize(stream)

    # ---- roofline of the dominant kernel (the tracker), this rank, last step ----
    peak, peak_src = _peaks()
    sha = source_sha()
    prof = _ncu_profile()
    ref_steps = census["steps_per_path"] if census else tot_steps / tot_paths
    ref_coll = census["collisions_per_path"] if census else tot_coll / tot_paths
    paths_rank = float(st.paths) / args.steps
    alg_bytes_rank = (32.0 * (ref_steps + ref_coll) + 32.0) * paths_rank
    achieved = alg_bytes_rank / (k_ms * 1e-3) / 1e9 if k_ms > 0 else 0.0
    traffic, dram, issue = None, None, None
    if prof and "pt_wavefront_kernel" in prof.get("kernels", {}):
        pk = prof["kernels"]["pt_wavefront_kernel"]
        stale = prof.get("source_sha") != sha or prof.get("config") != wl.name
        traffic = pk.get("dram_bytes")
        if traffic and pk.get("duration_ms"):
            g = traffic / (pk["duration_ms"] * 1e-3) / 1e9
            dram = {"bytes_per_launch": traffic, "achieved": g, "unit": "GB/s", "frac": g / peak,
                    "what": "measured DRAM bytes of one tracker launch / its duration under ncu / HBM peak"}
        if pk.get("warp_inst_per_cycle_per_sm") is not None:
            issue = {"primary": True, "achieved": pk["warp_inst_per_cycle_per_sm"], "peak": 4.0, "unit": "warp instructions / cycle / SM",
                     "frac": pk["warp_inst_per_cycle_per_sm"] / 4.0, "issue_active_pct": pk.get("issue_active_pct"),
                     "active_lanes_per_warp_inst": pk.get("active_lanes_per_warp_inst"),
                     "thread_inst_frac": (pk["warp_inst_per_cycle_per_sm"] / 4.0) * (pk.get("active_lanes_per_warp_inst") or 32.0) / 32.0,
                     "what": "the limiter that binds: 4 schedulers x 1 warp instruction per cycle per SM; thread_inst_frac also counts idle lanes"}
        prof_info = {"file": "profiles/ncu_live.json", "source_sha": prof.get("source_sha"), "bench_source_sha": sha, "stale": stale,
                     "config": prof.get("config"), "kernel_instance": pk.get("kernel"), "launch": pk.get("launch")}
    else:
        prof_info = {"file": None, "bench_source_sha": sha, "stale": True}
    roofline = {"kernel": "pt_wavefront_kernel", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic,
                # which limiter binds, and how close the kernel runs to it (the HBM block above is SURVEY 8d's definition and is
                # void as a bound for this kernel: see `note`)
                "binding": "issue" if issue else None, "binding_frac": issue["frac"] if issue else None,
                "dram": dram, "issue": issue, "ncu": prof_info, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_bytes_rank / max(1, k_n), "launches_per_step": k_n,
                "kernel_ms_per_step": k_ms, "kernel_share_of_step": k_ms / ms_per_step if ms_per_step > 0 else None,
                "reference_steps_per_path": ref_steps, "reference_collisions_per_path": ref_coll,
                "executed_steps_per_path": tot_steps / tot_paths, "tracked_path_fraction": tot_tracked / tot_paths,
                "fetch_fraction_of_executed_steps": tot_fetch / max(1.0, tot_steps),
                "note": "`achieved`/`frac` follow SURVEY 8d: ALGORITHMIC bytes = the reference's texel reads (32 B per tracking step and "
                        "per collision of the reference, counted by a census render at skip level 0, + 32 B of accumulator per path) / "
                        "tracker time.  The tracker proves most of them unnecessary (empty bricks, byte thresholds) and never issues them, "
                        "so this fraction can exceed 1 and does not bound the kernel; `dram` is the measured traffic and `issue` the "
                        "roofline that binds (ncu pass of tools/gpu_profile.sh, `ncu.stale` says whether it belongs to these sources)"}
    sv_bytes = 4.0 * wl.n_sv ** 3 + 4.0 * wl.n_cloud ** 3
    sv = {"mode": r.shadowVolumeTechnique.last_build_mode(), "rebuild_ms": sv_ms, "rebuild_ms_min": float(min(sv_times)),
          "n_sv": wl.n_sv, "n_cloud": wl.n_cloud, "algorithmic_bytes": sv_bytes, "achieved_gbs": sv_bytes / sv_ms / 1e6,
          "frac_of_peak": sv_bytes / sv_ms / 1e6 / peak, "peak_source": peak_src}
    if prof and "sv_fused" in prof.get("kernels", {}):
        sk = prof["kernels"]["sv_fused"]              # the last sv_fused* launch of the ncu pass (tools/ncu_summary.py live)
        sv["ncu_kernel"] = sk.get("kernel")
        if sk.get("dram_bytes") and sk.get("duration_ms"):
            sv["dram_bytes_measured"] = sk["dram_bytes"]
            sv["dram_frac"] = sk["dram_bytes"] / (sk["duration_ms"] * 1e-3) / 1e9 / peak
            sv["dram_frac_at_live_time"] = sk["dram_bytes"] / (sv_ms * 1e-3) / 1e9 / peak
            sv["ncu_stale"] = prof.get("source_sha") != sha or prof.get("config") != wl.name

    e2e = None
    parity = None
    cpu_baseline = None
    # ---- e2e: host buffers in, host picture out ----
    # Every step uploads the density grid from pinned host memory, rebuilds the majorant / skip tables / shadow volume,
    # renders, combines and downloads the picture.  Two step contexts (own volumes, group buffers, streams) alternate, so the
    # upload of step i+1 (copy engine) overlaps the render of step i; `serial_ms_per_step` is one step alone.
    # N > 1: every rank uploads 1/N of the grid and the slices are exchanged GPU to GPU by the copy engines
    # (crv_group_staging_all_gather), so the PCIe links work in parallel and nothing is broadcast by a kernel.
    if not args.no_e2e:
        n_vox = wl.n_cloud ** 3
        host_grid = torch.empty(n_vox, dtype=torch.float32).pin_memory()
        host_grid.numpy()[:] = r.cloudImage.download(stream)      # filled once, outside the timed region
        hg = host_grid.numpy()

        class Pipe:
            pass

        def make_pipe(k):
            p = Pipe()
            p.stream_obj = torch.cuda.Stream(device=dev)
            p.s = p.stream_obj.cuda_stream
            p.k = k
            if k == 0:
                p.r = r
            else:
                p.r = crv.make_renderer(ctx, wl, cloud_layout=args.layout, math_mode=math_mode)
                p.r.resultImage.close()
                p.r.pathTracingTechnique.SetScheduler(args.scheduler)
                p.r.pathTracingTechnique.SetSkipLevel(args.skip_level)
                p.r.pathTracingTechnique.SetJobOrder(args.job_order)
            p.img = split.group.image(k)
            p.r.resultImage = p.img
            p.r.pathTracingTechnique.SetFrameReferences([p.img])
            if k != 0:
                p.r.set_cloud_synthetic(wl.n_cloud, wl.fbm_seed, stream)
            p.host_img = torch.empty((H, W, 4), dtype=torch.float32).pin_memory()
            p.hi = p.host_img.numpy()
            return p

        my = frames_for_rank(all_seeds, rank, world)

        def e2e_upload(p):
            """first half of a step: the density grid leaves host memory (copy engine; at N > 1 the slices then travel GPU to GPU)"""
            if world > 1:
                split.group.staging_all_gather(p.k, host_grid.data_ptr(), 4 * n_vox, p.s)       # H2D of my slice + peer exchange
                p.r.cloudImage.upload_device(split.group.staging_ptr(p.k), p.s)                # retile into the brick layout
                split.group.staging_release(p.k, p.s)
            else:
                p.r.cloudImage.upload(hg, p.s, wait=False)                                     # H2D: the density grid

        def e2e_render(p):
            """second half: majorant (read back: SetCloudProperties needs it on the host), tables, shadow volume, picture, combine"""
            p.r.SetCloudProperties((wl.n_cloud,) * 3, (1.0 / wl.n_cloud,) * 3, p.r.cloudImage.majorant(p.s))
            p.r.UpdateCloudData(p.s)                                                           # descriptor updates + shadow-volume rebuild
            if len(my) * world == spp_total and args.combine == "push":      # even split: only the tracked pixels cross between the GPUs, pushed by the fold
                targets = split.group.push_begin(p.k)
                p.r.pathTracingTechnique.RenderBatchPush(my, targets, 1.0 / len(my), p.s, 0)
                split.group.push_finish(1.0 / spp_total, p.r.pathTracingTechnique.PixelLists(), p.s, p.k)
            else:
                p.img.clear(p.s)
                p.r.pathTracingTechnique.RenderBatch(my, 0, crv.ACCUM_SUM, p.s, 0)
                if len(my) * world == spp_total:
                    split.group.all_reduce_lists_sum_scale(1.0 / spp_total, 1.0 / len(my), p.r.pathTracingTechnique.PixelLists(), p.s, p.k)
                else:
                    split.group.all_reduce_sum_scale(1.0 / spp_total, p.s, p.k)

        def e2e_enqueue(p):
            e2e_upload(p)
            e2e_render(p)

        def e2e_finish(p):
            p.img.download(p.s, out=p.hi)                                                      # D2H: the picture (synchronises p.s)

        pipes = [make_pipe(0), make_pipe(1)]
        ctx.synchronize(stream)
        for p in pipes:                                                         # warm both contexts, one at a time
            e2e_enqueue(p)
            e2e_finish(p)
        barrier()
        t0 = time.perf_counter()
        e2e_enqueue(pipes[0])
        e2e_finish(pipes[0])
        barrier()
        serial_s = time.perf_counter() - t0
        n_e2e = max(16, args.steps)                                             # pipelined steps in the timed region (the fill is inside)
        barrier()
        t0 = time.perf_counter()
        # (enqueueing step i + 1's upload before step i's majorant read-back -- so that the copy engine never idles -- was
        # measured and serialised the steps instead: 16.9 ms per step at N = 1, profiles/r2t_e2e_reorder.txt)
        e2e_enqueue(pipes[0])
        for i in range(1, n_e2e):
            e2e_enqueue(pipes[i % 2])
            e2e_finish(pipes[(i - 1) % 2])
        e2e_finish(pipes[(n_e2e - 1) % 2])
        barrier()
        e2e_s = (time.perf_counter() - t0) / n_e2e
        split.group.status(pipes[0].s)
        if world > 1:
            t = torch.tensor([e2e_s, serial_s], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_s, serial_s = float(t[0].item()), float(t[1].item())
        e2e_err = None
        if rank == 0 and not animated:
            e2e_err = float((np.abs(pipes[0].hi - final_picture) / np.maximum(np.abs(final_picture), 1e-3)).max())
        e2e = {"value": paths_per_step / e2e_s / 1e6, "unit": "Mpath-samples/s",
               "h2d_bytes_per_step": int(4 * n_vox + 68 + 60 + 28 + 152 + 24),
               "d2h_bytes_per_step": int(W * H * 16) * world, "ms_per_step": e2e_s * 1e3, "steps": n_e2e,
               "serial_ms_per_step": serial_s * 1e3, "serial_value": paths_per_step / serial_s / 1e6,
               "max_rel_diff_vs_device_resident_picture": e2e_err,
               "pcie_floor_ms": 4.0 * n_vox / world / 53e9 * 1e3,
               "includes": "every step: pinned-host grid upload, majorant + skip tables, shadow-volume rebuild, render, combine, picture "
                           "download; throughput of two step contexts on two streams (step i+1's upload overlaps step i's render, "
                           "pipeline fill included); serial_* = one step alone; N > 1: every rank uploads 1/N of the grid and the slices "
                           "travel GPU to GPU on the copy engines, so h2d_bytes_per_step is per job; every rank downloads the picture; "
                           "pcie_floor_ms = the upload alone at the 53 GB/s this box's PCIe link sustains"}
        for p in pipes[1:]:
            p.r.resultImage = crv.Image(ctx, 4, 4)
            p.r.close()
        r.resultImage = img
        pt.SetFrameReferences([img])

    # ---- rank 0, N = 1: bounded CPU baseline + parity of a sample of the same workload against the oracle ----
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from tests import oracle_binding as orc
        if animated:
            r.lightDirection = wl.light_dir
            r.UpdateShadowVolume(stream)
        cloud_h = r.cloudImage.download(stream)
        sv_h = r.shadowVolumeImage.download(stream)
        seeds = crv.msvc_rand_sequence(256)
        oargs = (orc, cloud_h, sv_h, r.cameraProperties, r.cloudProperties, r.parameters, r.shadowVolumeProperties, wl)
        kind = _cpu_kind(orc, wl)
        frames, rows = _oracle_plan(*oargs, seeds[0], 12.0, kind)
        p, s = _oracle_sample(*oargs, seeds, frames, rows, kind)
        cpu_baseline = {"value": p / s / 1e6, "unit": "Mpath-samples/s", "cores": orc.max_threads(), "kind": kind,
                        "sample": _sample_text(wl, frames, rows, p, s), "note": _cpu_note(kind, wl)}
        if kind == "port" and orc.ref_shaders_lib() is not None:
            p2, s2 = _oracle_sample(*oargs, seeds, max(1, frames // 4), rows, "reference")
            cpu_baseline["reference_shader_value"] = p2 / s2 / 1e6
        # parity of one frame's rows against the GPU (same seed, single dispatch, reference accumulation)
        r.reset_accumulation(stream)
        pt.SetDebugCounters(True)
        r.pushConstants.seed = int(seeds[1])
        pt.RecordDrawCommands(stream, 0)
        g_img = img.download(stream)
        gs, gc = pt.DownloadDebugCounters(stream)
        pt.SetDebugCounters(False)
        r.reset_accumulation(stream)
        pt.RecordDrawCommands(stream, 0)          # the same frame on the production path (skip level 1): must be the same bits
        fast_img = img.download(stream)
        pc = crv.PushConstants(); pc.seed = int(seeds[1]); pc.frameCount = 0
        ys = list(range(0, H, max(rows, 8)))
        n_same = n_tot = 0
        max_err = 0.0
        for y in ys:
            c_img = np.zeros((H, W, 4), np.float32)
            _, cs, cc = orc.path_trace_frame(cloud_h, sv_h, r.cameraProperties, r.cloudProperties, r.parameters,
                                             r.shadowVolumeProperties, pc, wl.absorption, c_img, rect=(0, y, W, y + 1))
            same = (gs[y] == cs[y]) & (gc[y] == cc[y])
            err = np.abs(g_img[y] - c_img[y]) / np.maximum(np.abs(c_img[y]), 1e-3)
            n_same += int(same.sum()); n_tot += W
            if same.any():
                max_err = max(max_err, float(err[same].max()))
        parity = {"checked_paths": n_tot, "threshold_divergent": n_tot - n_same, "max_rel_err_matching": max_err,
                  "tolerance": 1e-4, "production_path_bit_identical": bool(np.array_equal(g_img.view(np.uint32), fast_img.view(np.uint32)))}
        assert max_err <= 1e-4, f"per-sample radiance differs from the oracle: {max_err:.3e}"
        assert parity["production_path_bit_identical"], "skip level 1 changed a frame"

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "Mpath-samples/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": _config_dict(wl, args, world, spp_total, args.layout, args.math),
            "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "sv": sv,
            "tracked_paths": {"value": tot_tracked / args.steps / (ms_per_step * 1e-3) / 1e6, "unit": "Mpath-samples/s",
                              "fraction_of_paths": tot_tracked / tot_paths,
                              "note": "paths that go through the tracker; the rest are camera rays that miss the cloud box or cross it "
                                      "through empty bricks only (sample = sampleBackground(dir), proven per pixel once per scene)"},
            "combine": reduce_cmp, "scaling_other": scaling_other, "multi_gpu_parity": multi_gpu_parity,
            "math_modes": {"unit": "Mpath-samples/s", **math_modes,
                           "note": "free-flight update t += -log(zeta)/sigma: 'strict' (library default) = logf-accurate log, IEEE divide, "
                                   "IEEE add (the CPU oracle's evaluation); 'driver' = lg2 on the SFU and one FMA, inside the precision "
                                   "Vulkan grants a driver for the same GLSL; `value` and `parity` are the --math mode's"},
            "running_mean": running_mean, "census": census, "cpu_baseline": cpu_baseline, "parity": parity,
        }
        _emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def _claim_stdout():
    """Route everything C libraries print on fd 1 (NCCL's version banner) to stderr; the JSON line keeps the real stdout."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)


_REAL_STDOUT = None


def _emit(line: dict):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c2")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="N > 1: 'strong' splits the workload's spp over the GPUs (the picture is fixed), 'weak' gives every GPU the workload's spp")
    ap.add_argument("--layout", default="brick", choices=["linear", "brick", "tex3d"])
    ap.add_argument("--spp", type=int, default=0, help="override the workload's spp")
    ap.add_argument("--track-threshold", type=int, default=0)
    ap.add_argument("--ctas-per-sm", type=int, default=0)
    ap.add_argument("--steps-per-decision", type=int, default=0)
    ap.add_argument("--sample-buffer-mib", type=int, default=0, help="sample-record buffer of the tracker (0 = library default)")
    ap.add_argument("--math", default="strict", choices=["strict", "driver"],
                    help="free-flight update t += -log(zeta)/sigma: 'strict' (the library default) = logf-accurate log, IEEE divide, IEEE "
                         "add; 'driver' = lg2 on the SFU * (-ln2/sigma) in one FMA, the latitude GLSL gives a Vulkan driver "
                         "(crv_c_api.h: crv_pt_math_mode); the other mode's figure is reported under math_modes")
    ap.add_argument("--scheduler", default="pool", choices=["pool", "lanes"])
    ap.add_argument("--skip-level", type=int, default=1, choices=[0, 1])
    ap.add_argument("--combine", default="push", choices=["push", "peer", "peer_dense"],
                    help="how the members' sums are combined in the timed region (A/B; see `combine` in the output line)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--job-order", type=int, default=-1, choices=[-1, 0, 1],
                    help="crv_pt_set_job_order: 0 frame-major, 1 pixel-major over the sorted hit list, -1 automatic (identical pictures)")
    ap.add_argument("--quick", action="store_true", help="only the timed region (no census / other modes): for profiler passes")
    ap.add_argument("--print-source-sha", action="store_true")
    args = ap.parse_args()
    if args.print_source_sha:
        print(source_sha())
        return 0
    if args.warmup < 3:
        args.warmup = 3
    wl = load_workloads().CONFIGS[args.config]
    if args.spp:
        wl = dataclasses.replace(wl, spp=args.spp)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "ours" and world != args.gpus and args.gpus > 1:
        print(f"bench.py: --gpus {args.gpus} needs torchrun with {args.gpus} ranks (WORLD_SIZE={world})", file=sys.stderr)
        return 2
    if args.impl == "ours" and args.scaling == "strong" and wl.spp % max(1, world) != 0:
        print(f"bench.py: spp {wl.spp} does not divide over {world} GPUs", file=sys.stderr)
        return 2
    _claim_stdout()
    if args.impl == "reference":
        return run_reference(args, wl)
    return run_ours(args, wl)


if __name__ == "__main__":
    sys.exit(main())
