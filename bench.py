#!/usr/bin/env python
"""bench.py -- the contract benchmark of the PathTracer + ShadowVolume hot path.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config c2]
  torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W        (N > 1)

One "step" = one full progressive render of the workload (all spp) of BASELINE.json's headline configuration
(`c2`: 512^3 fBm cloud, 1920x1080, 256 spp, albedo 0.9, HG g = 0.85) from parameters already resident in HBM.
metric = Mpath-samples/s = W*H*spp_total / step time (PT launches + accumulation + cross-GPU reduce), whole job.
N > 1: the density grid and shadow volume are replicated, every rank renders `spp` frames of its own
(frame f of the global list goes to rank f mod N; weak scaling: spp_total = N * spp), partial fp32 SUMs are
combined with one NCCL all-reduce over NVLink and divided by spp_total.

Extra keys: `e2e` (same metric through the public API from HOST buffers: H2D grid upload + majorant/brick-max +
shadow-volume rebuild + render + D2H of the image inside the timed region), `roofline` (the tracker kernel against
the measured HBM peak, algorithmic bytes = 32 B x (tracking steps + collisions) + 32 B per path, counted by the
kernel), `sv` (shadow-volume rebuild ms and GB/s), `cpu_baseline` (the CPU oracle on a bounded sample of the same
workload), `clocks`, `gpu_launches`, `parity`.

--impl reference times the reference's CPU restatement (oracle/, OpenMP on all host cores) on the same config:
the reference's own SPIR-V/lavapipe path cannot exist in this image (no Vulkan loader, glslc or .spv).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "Mpath-samples/s"


def _peak():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json, burst copy)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


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
            self._stop_evt.wait(0.02)

    def stop(self):
        self._stop_evt.set()
        if self.is_alive():
            self.join(timeout=2)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# =============================================================================================================
# reference arm / cpu_baseline: the CPU oracle (the ONLY code here allowed to touch oracle/)
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


def _cpu_kind(orc):
    """"reference": the reference's OWN shaders compiled by g++ (oracle/_ref/libcrv_refshaders.so, see
    oracle/ref_shaders_host.cpp) when that library exists; "port": the hand restatement (crv_oracle.cpp) otherwise."""
    return "reference" if orc.ref_shaders_lib() is not None else "port"


def _vulkan_probe():
    """oracle/vulkan_probe.py: is the north star's SPIR-V / lavapipe oracle available here?  (never, in this round's images)"""
    try:
        import importlib.util
        spec = importlib.util.spec_from_file_location("vulkan_probe", os.path.join(ROOT, "oracle", "vulkan_probe.py"))
        m = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(m)
        r = m.probe()
        return "available" if r["available"] else r["reason"]
    except Exception as e:       # the probe is informational
        return f"probe failed: {e}"


def _cpu_note(kind, wl):
    if kind == "reference":
        note = ("the reference's own PathTracer.comp compiled by g++ through oracle/glsl_compat.h (OpenMP over image rows); "
                "SPIR-V on lavapipe: " + _vulkan_probe())
        if wl.absorption > 0:
            note += (f"; the reference has no absorption term, so its paths run at albedo 1 on the same scene "
                     f"(the workload's absorption {wl.absorption:g} is an extension of this repo; `port_value` is the "
                     f"restatement timed WITH it)")
        return note
    return "CPU restatement of the shaders (lavapipe / SPIR-V unavailable in this image)"


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
    kind = _cpu_kind(orc)
    t0 = time.perf_counter()
    sv = orc.refsh_shadow_volume(cloud, cp, sp) if kind == "reference" else orc.shadow_volume(cloud, cp, sp)
    sv_s = time.perf_counter() - t0
    log(f"[reference] shadow volume {wl.n_sv}^3 on the CPU ({kind}): {sv_s:.1f}s")
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
    port_value = None
    if kind == "reference":      # the restatement on one step's sample, with the workload's absorption: apples to apples
        p, s = _oracle_sample(orc, cloud, sv, cam, cp, par, sp, wl, seeds, frames, rows, "port")
        port_value = p / s / 1e6
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "Mpath-samples/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * secs / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": _config_dict(wl, args, 1, "cpu"),
        "cpu_baseline": {"value": value, "unit": "Mpath-samples/s", "cores": orc.max_threads(), "kind": kind,
                         "sample": sample, "sv_build_s": sv_s, "port_value": port_value, "note": _cpu_note(kind, wl)},
        "e2e": {"value": value, "unit": "Mpath-samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    _emit(line)
    return 0


def _config_dict(wl, args, world, layout):
    return {"workload": f"{wl.name}: {wl.n_cloud}^3 fBm cloud, {wl.width}x{wl.height}, {wl.spp} spp/GPU, "
                        f"densityScaling {wl.density_scaling:g}, absorption {wl.absorption:g}, "
                        f"g {wl.phase_g if wl.phase_g is not None else 'isotropic'}, N_sv {wl.n_sv}",
            "spp_total": wl.spp * world, "parallelism": f"spp-split x{world} (grid replicated, fp32 sum all-reduce)",
            "layout": layout, "math": "libm" if layout == "cpu" else getattr(args, "math", "strict"),
            "l2": "inputs larger than L2 (grid + shadow volume >= 1 GB vs 126 MB)"}


# =============================================================================================================
# our arm
# =============================================================================================================
def run_ours(args, wl):
    import torch

    import cloudrendering_vulkan_b200 as crv

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
    stream = torch.cuda.current_stream().cuda_stream
    log = (lambda s: print(s, file=sys.stderr, flush=True)) if rank == 0 else (lambda s: None)

    ctx = crv.Context(local)
    W, H, spp = wl.width, wl.height, wl.spp
    math_mode = crv.PT_MATH_DRIVER if args.math == "driver" else crv.PT_MATH_STRICT
    r = crv.make_renderer(ctx, wl, cloud_layout=args.layout, math_mode=math_mode)
    # accumulation buffer owned by torch so that NCCL can reduce it in place
    accum = torch.zeros((H, W, 4), dtype=torch.float32, device=dev)
    img = crv.Image(ctx, W, H, device_ptr=accum.data_ptr())
    r.resultImage.close()
    r.resultImage = img
    r.pathTracingTechnique.SetFrameReferences([img])
    if args.track_threshold or args.ctas_per_sm or args.steps_per_decision or args.sample_buffer_mib:
        r.pathTracingTechnique.SetTuning(ctas_per_sm=args.ctas_per_sm, track_threshold=args.track_threshold,
                                         steps_per_decision=args.steps_per_decision,
                                         sample_buffer_bytes=args.sample_buffer_mib << 20)
    r.pathTracingTechnique.SetScheduler(args.scheduler)
    r.set_cloud_synthetic(wl.n_cloud, wl.fbm_seed, stream)
    ctx.synchronize(stream)
    # shadow-volume rebuild: 3 warm rebuilds, then the median of 7 (CUDA events on the launching stream, around the kernel)
    sv_times = []
    for i in range(10):
        r.UpdateShadowVolume(stream)
        if i >= 3:
            sv_times.append(r.shadowVolumeTechnique.last_build_ms())
    sv_ms = float(np.median(sv_times))

    all_seeds = crv.msvc_rand_sequence(spp * world)
    my_seeds = np.ascontiguousarray(all_seeds[rank::world])          # frame f -> rank f mod world
    mode = crv.ACCUM_RUNNING_MEAN if world == 1 else crv.ACCUM_SUM
    pt = r.pathTracingTechnique

    def step():
        img.clear(stream)
        pt.RenderBatch(my_seeds, 0, mode, stream, 0)
        if world > 1:
            dist.all_reduce(accum, op=dist.ReduceOp.SUM)
            img.scale(1.0 / (spp * world), stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    pt.ResetStats(stream)
    launches0 = ctx.launch_count
    sampler = ClockSampler(local)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    barrier()
    dev_ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop()
    launches = ctx.launch_count - launches0
    st = pt.GetStats(stream)
    k_ms, k_n = pt.LastRenderMs()          # tracker launches of the LAST step (events on the launching stream)
    if world > 1:
        t = torch.tensor([dev_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms = float(t.item())
        cnt = torch.tensor([st.paths, st.tracking_steps, st.collisions, st.density_fetches, launches], dtype=torch.float64, device=dev)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
        tot_paths, tot_steps, tot_coll, tot_fetch, launches = [float(x) for x in cnt.tolist()]
    else:
        tot_paths, tot_steps, tot_coll, tot_fetch = float(st.paths), float(st.tracking_steps), float(st.collisions), float(st.density_fetches)
    ms_per_step = dev_ms / args.steps
    paths_per_step = W * H * spp * world
    value = paths_per_step / (ms_per_step * 1e-3) / 1e6
    assert abs(tot_paths - paths_per_step * args.steps) < 0.5, "kernel path counter disagrees with the workload size"

    # ---- the same steps in the other math mode (N = 1): both ends of the latitude GLSL leaves the driver are on record ----
    math_modes = {args.math: value}
    if world == 1:
        other = "strict" if args.math == "driver" else "driver"
        r.extensions.math_mode = crv.PT_MATH_STRICT if other == "strict" else crv.PT_MATH_DRIVER
        pt.SetExtensions(r.extensions)
        for _ in range(2):
            step()
        torch.cuda.synchronize()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for _ in range(args.steps):
            step()
        a1.record()
        torch.cuda.synchronize()
        math_modes[other] = paths_per_step / (a0.elapsed_time(a1) / args.steps * 1e-3) / 1e6
        r.extensions.math_mode = math_mode
        pt.SetExtensions(r.extensions)

    # ---- roofline of the dominant kernel (the tracker), this rank, last step ----
    peak, peak_src = _peak()
    alg_bytes_rank = (32.0 * (st.tracking_steps + st.collisions) + 32.0 * st.paths) / args.steps
    achieved = alg_bytes_rank / (k_ms * 1e-3) / 1e9
    traffic, issue = None, None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        traffic = tj.get("pt_wavefront_kernel_bytes_per_launch")
        if "warp_inst_per_cycle_per_sm" in tj:      # the limiter that actually binds: issue slots (from the same ncu pass)
            issue = {"warp_inst_per_cycle_per_sm": tj["warp_inst_per_cycle_per_sm"], "peak": 4.0,
                     "frac": tj["warp_inst_per_cycle_per_sm"] / 4.0, "issue_active_pct": tj.get("issue_active_pct"),
                     "active_lanes_per_warp_inst": tj.get("active_lanes_per_warp_inst"), "source": "profiles/traffic.json (ncu, not live)"}
    except Exception:
        pass
    roofline = {"kernel": "pt_wavefront_kernel", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "issue": issue, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_bytes_rank / max(1, k_n), "launches_per_step": k_n,
                "kernel_ms_per_step": k_ms, "kernel_share_of_step": k_ms / ms_per_step,
                "steps_per_path": tot_steps / tot_paths, "collisions_per_path": tot_coll / tot_paths,
                "fetch_fraction": tot_fetch / max(1.0, tot_steps),
                "note": "algorithmic bytes = the reference's texel reads (32 B per tracking step and per collision + 32 B of "
                        "accumulator per path, SURVEY 8d); the proof-based skip tables avoid ~98 % of the step reads, so "
                        "`achieved` can exceed what DRAM could deliver -- `traffic` is the measured DRAM bytes per launch and "
                        "the kernel's real limiter is instruction issue (ncu: profiles/*_pt_ncu.txt)"}
    sv_bytes = 4.0 * wl.n_sv ** 3 + 4.0 * wl.n_cloud ** 3

    line = None
    e2e = None
    parity = None
    cpu_baseline = None
    # ---- e2e: host buffers in, host image out (rank-local inputs; all ranks take part in the reduce) ----
    # Every step uploads the density grid from pinned host memory, rebuilds the majorant / brick tables / shadow volume,
    # renders, reduces and downloads the image.  Two step contexts (own volumes, images, streams) alternate, so the upload
    # of step i+1 (copy engine, ~10 ms at PCIe rate) overlaps the render of step i; `serial_ms_per_step` is one step alone.
    if not args.no_e2e:
        host_grid = torch.empty(wl.n_cloud ** 3, dtype=torch.float32).pin_memory()
        host_grid.numpy()[:] = r.cloudImage.download(stream)      # filled once, outside the timed region
        hg = host_grid.numpy()

        class Pipe:
            pass

        def make_pipe(first):
            p = Pipe()
            p.stream_obj = torch.cuda.Stream(device=dev)
            p.s = p.stream_obj.cuda_stream
            if first:
                p.r, p.accum, p.img = r, accum, img
            else:
                p.r = crv.make_renderer(ctx, wl, cloud_layout=args.layout, math_mode=math_mode)
                p.accum = torch.zeros((H, W, 4), dtype=torch.float32, device=dev)
                p.img = crv.Image(ctx, W, H, device_ptr=p.accum.data_ptr())
                p.r.resultImage.close()
                p.r.resultImage = p.img
                p.r.pathTracingTechnique.SetFrameReferences([p.img])
                p.r.pathTracingTechnique.SetScheduler(args.scheduler)
                p.r.set_cloud_synthetic(wl.n_cloud, wl.fbm_seed, stream)
            p.host_img = torch.empty((H, W, 4), dtype=torch.float32).pin_memory()
            p.hi = p.host_img.numpy()
            # N > 1: the (replicated) grid crosses PCIe once, on rank 0, and reaches the other GPUs by an NCCL broadcast over
            # NVLink -- eight concurrent 537 MB host reads contend for the host's memory system (measured: 25 ms instead of 10)
            p.grid_dev = torch.empty(wl.n_cloud ** 3, dtype=torch.float32, device=dev) if world > 1 else None
            return p

        def e2e_enqueue(p):
            with torch.cuda.stream(p.stream_obj):
                if world > 1:
                    if rank == 0:
                        p.grid_dev.copy_(host_grid, non_blocking=True)          # H2D: the density grid, once per job
                    dist.broadcast(p.grid_dev, src=0)                           # NVLink
                    p.r.cloudImage.upload_device(p.grid_dev.data_ptr(), p.s)    # retile into the brick layout
                else:
                    p.r.cloudImage.upload(hg, p.s)                              # H2D: the density grid
                p.r.SetCloudProperties((wl.n_cloud,) * 3, (1.0 / wl.n_cloud,) * 3, p.r.cloudImage.majorant(p.s))
                p.r.UpdateCloudData(p.s)                                        # descriptor updates + shadow-volume rebuild
                p.r.pathTracingTechnique.RenderBatch(my_seeds, 0, mode, p.s, 0)
                if world > 1:
                    dist.all_reduce(p.accum, op=dist.ReduceOp.SUM)
                    p.img.scale(1.0 / (spp * world), p.s)

        def e2e_finish(p):
            p.img.download(p.s, out=p.hi)                                       # D2H: the result image (synchronises p.s)

        pipes = [make_pipe(True), make_pipe(False)]
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
        n_e2e = max(6, args.steps)                                              # pipelined steps in the timed region (the fill is inside)
        barrier()
        t0 = time.perf_counter()
        e2e_enqueue(pipes[0])
        for i in range(1, n_e2e):
            e2e_enqueue(pipes[i % 2])
            e2e_finish(pipes[(i - 1) % 2])
        e2e_finish(pipes[(n_e2e - 1) % 2])
        barrier()
        e2e_s = (time.perf_counter() - t0) / n_e2e
        if world > 1:
            t = torch.tensor([e2e_s, serial_s], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_s, serial_s = float(t[0].item()), float(t[1].item())
        e2e = {"value": paths_per_step / e2e_s / 1e6, "unit": "Mpath-samples/s", "h2d_bytes_per_step": int(4 * wl.n_cloud ** 3 + 68 + 60 + 28 + 152 + 24),
               "d2h_bytes_per_step": int(W * H * 16), "ms_per_step": e2e_s * 1e3, "steps": n_e2e,
               "serial_ms_per_step": serial_s * 1e3, "serial_value": paths_per_step / serial_s / 1e6,
               "includes": "every step: pinned-host grid upload, majorant + brick tables, shadow-volume rebuild, render, reduce, image "
                           "download; two step contexts on two streams alternate so step i+1's upload overlaps step i's render "
                           "(pipeline fill included); serial_* = one step alone; N > 1: rank 0 uploads, NCCL broadcast to the "
                           "replicas, so h2d_bytes_per_step is per job, not per GPU"}
        pipes[1].r.close()

    # ---- rank 0, N = 1: bounded CPU-oracle baseline + parity of a sample of the same workload ----
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from tests import oracle_binding as orc
        cloud_h = r.cloudImage.download(stream)
        sv_h = r.shadowVolumeImage.download(stream)
        seeds = crv.msvc_rand_sequence(256)
        oargs = (orc, cloud_h, sv_h, r.cameraProperties, r.cloudProperties, r.parameters, r.shadowVolumeProperties, wl)
        kind = _cpu_kind(orc)
        frames, rows = _oracle_plan(*oargs, seeds[0], 12.0, kind)
        p, s = _oracle_sample(*oargs, seeds, frames, rows, kind)
        cpu_baseline = {"value": p / s / 1e6, "unit": "Mpath-samples/s", "cores": orc.max_threads(), "kind": kind,
                        "sample": _sample_text(wl, frames, rows, p, s), "note": _cpu_note(kind, wl)}
        if kind == "reference":
            p2, s2 = _oracle_sample(*oargs, seeds, max(1, frames // 4), rows, "port")
            cpu_baseline["port_value"] = p2 / s2 / 1e6
        # parity of one frame's rows against the GPU (same seed, single dispatch)
        r.reset_accumulation(stream)
        pt.SetDebugCounters(True)
        r.pushConstants.seed = int(seeds[1])
        pt.RecordDrawCommands(stream, 0)
        g_img = img.download(stream)
        gs, gc = pt.DownloadDebugCounters(stream)
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
                  "tolerance": 1e-4}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "Mpath-samples/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": _config_dict(wl, args, world, args.layout),
            "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline,
            "sv": {"mode": r.shadowVolumeTechnique.last_build_mode(),
                   "rebuild_ms": sv_ms, "rebuild_ms_min": float(min(sv_times)), "n_sv": wl.n_sv, "n_cloud": wl.n_cloud,
                   "algorithmic_bytes": sv_bytes, "achieved_gbs": sv_bytes / sv_ms / 1e6,
                   "frac_of_peak": sv_bytes / sv_ms / 1e6 / peak, "peak_source": peak_src},
            "math_modes": {"unit": "Mpath-samples/s", **math_modes,
                           "note": "free-flight update t += -log(zeta)/sigma: 'strict' = logf-accurate log, IEEE divide, IEEE add "
                                   "(the CPU oracle's evaluation); 'driver' = lg2 on the SFU and one FMA, inside the precision "
                                   "Vulkan grants a driver for the same GLSL; `value` and `parity` are the --math mode's"},
            "cpu_baseline": cpu_baseline, "parity": parity,
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
    ap.add_argument("--layout", default="brick", choices=["linear", "brick", "tex3d"])
    ap.add_argument("--spp", type=int, default=0, help="override spp per GPU (the default is the config's)")
    ap.add_argument("--track-threshold", type=int, default=0)
    ap.add_argument("--ctas-per-sm", type=int, default=0)
    ap.add_argument("--steps-per-decision", type=int, default=0)
    ap.add_argument("--sample-buffer-mib", type=int, default=0, help="sample-record buffer of the tracker (0 = library default)")
    ap.add_argument("--math", default="driver", choices=["strict", "driver"],
                    help="free-flight update t += -log(zeta)/sigma: 'strict' = logf-accurate log, IEEE divide, IEEE add; "
                         "'driver' = lg2 on the SFU * (-ln2/sigma) in one FMA, the latitude GLSL gives a Vulkan driver "
                         "(crv_c_api.h: crv_pt_math_mode)")
    ap.add_argument("--scheduler", default="pool", choices=["pool", "lanes"])
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    from cloudrendering_vulkan_b200.scene import CONFIGS
    import dataclasses
    wl = CONFIGS[args.config]
    if args.spp:
        wl = dataclasses.replace(wl, spp=args.spp)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "ours" and world != args.gpus and args.gpus > 1:
        print(f"bench.py: --gpus {args.gpus} needs torchrun with {args.gpus} ranks (WORLD_SIZE={world})", file=sys.stderr)
        return 2
    _claim_stdout()
    if args.impl == "reference":
        return run_reference(args, wl)
    return run_ours(args, wl)


if __name__ == "__main__":
    sys.exit(main())
