#!/usr/bin/env python
"""Summarise ncu output brought back in gpurun_out/ into small text files under profiles/ (tracked).

  tools/ncu_summary.py launches gpurun_out/<tag>_launches.csv  > profiles/<tag>_launches.txt
  tools/ncu_summary.py full     gpurun_out/<tag>_pt.ncu-rep    > profiles/<tag>_pt_ncu.txt
  tools/ncu_summary.py source   gpurun_out/<tag>_pt.ncu-rep [N] > profiles/<tag>_pt_source.txt   (top-N lines by samples)
"""
import collections
import csv
import io
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed.avg.per_cycle_elapsed", "smsp__inst_executed.sum", "smsp__thread_inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__cycles_active.avg",
    "sm__inst_executed_pipe_xu.sum", "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_alu.sum",
    "sm__inst_executed_pipe_lsu.sum",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
]


def launches(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, data = rows[hi], rows[hi + 1:]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in data:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        v = v / 1e3 if r[ui] == "ns" else v * 1e3 if r[ui] == "ms" else v
        a = agg.setdefault(r[ki], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print(f"# {path}: {sum(a[0] for a in agg.values())} launches, {tot / 1e3:.2f} ms of kernel time (ncu: cold-cache, serialised)")
    print(f"{'n':>5} {'total_us':>12} {'avg_us':>10} {'share':>7}  kernel")
    for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"{a[0]:5d} {a[1]:12.1f} {a[1] / a[0]:10.1f} {100 * a[1] / tot:6.2f}%  {k}")


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    h, u = rows[0], rows[1]
    for v in rows[2:]:
        print(f"# {path}: {v[h.index('Kernel Name')]}  (ncu --set full --clock-control none)")
        for w in WANT:
            if w in h:
                i = h.index(w)
                print(f"{w:82s} {v[i]:>18s} {u[i]}")
        if "dram__bytes_read.sum" in h:
            def b(name):
                i = h.index(name)
                x = float(v[i].replace(",", ""))
                return x * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[u[i]]
            print(f"{'traffic = dram__bytes_read.sum + dram__bytes_write.sum':82s} {b('dram__bytes_read.sum') + b('dram__bytes_write.sum'):18.0f} byte")


def source(path, top=40):
    """per source line: share of warp-stall samples and of executed warp instructions (file:line, both sorted lists)"""
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "cuda,sass"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    fname, h, lines = "?", None, []
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            fname = r[1].split("/")[-1]
        elif r[0] == "Line No":
            h = r
        elif h is not None and len(r) == len(h) and r[0].isdigit():
            def num(name):
                try:
                    return float(r[h.index(name)].replace(",", ""))
                except Exception:
                    return 0.0
            lines.append((fname, int(r[0]), r[1].strip(), num("# Samples"), num("Instructions Executed"), num("Avg. Threads Executed"),
                          num("stall_long_sb"), num("stall_lg"), num("stall_wait"), num("stall_math")))
    ts = sum(l[3] for l in lines) or 1.0
    ti = sum(l[4] for l in lines) or 1.0
    print(f"# {path}: {int(ts)} warp-stall samples, {int(ti)} warp instructions executed")
    print(f"## top {top} source lines by stall samples   (share, long_sb/lg/wait/math samples, file:line, source)")
    for l in sorted(lines, key=lambda l: -l[3])[:top]:
        print(f"{100 * l[3] / ts:6.2f}%  {int(l[6]):6d} {int(l[7]):6d} {int(l[8]):6d} {int(l[9]):6d}  {l[0]}:{l[1]:<4d} {l[2][:110]}")
    print(f"## top {top} source lines by executed warp instructions   (share, avg active threads, file:line, source)")
    for l in sorted(lines, key=lambda l: -l[4])[:top]:
        print(f"{100 * l[4] / ti:6.2f}%  {l[5]:5.1f}  {l[0]}:{l[1]:<4d} {l[2][:110]}")


def functions(path):
    """warp instructions and average active lanes per source function (inlined device functions keep their own lines)"""
    import re
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "cuda,sass"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    fpath, h, lines = "?", None, []
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            fpath = r[1]
        elif r[0] == "Line No":
            h = r
        elif h is not None and len(r) == len(h) and r[0].isdigit():
            g = lambda n: float(r[h.index(n)].replace(",", "") or 0)
            lines.append((fpath, int(r[0]), g("Instructions Executed"), g("Thread Instructions Executed"), g("# Samples")))
    cache = {}

    def region(f, l):
        if f not in cache:
            try:
                cache[f] = open(f).read().split("\n")
            except Exception:
                cache[f] = None
        src = cache[f]
        if src is None:
            return f.split("/")[-1]
        for i in range(min(l, len(src)) - 1, -1, -1):
            if "__device__" in src[i] or "__global__" in src[i]:
                j = i
                while j < len(src) and "(" not in src[j].split("__launch_bounds__")[-1].split("__forceinline__")[-1]:
                    j += 1
                m = re.search(r"(\w+)\s*\(", src[j].split("__forceinline__")[-1].split("__launch_bounds__(PT_THREADS)")[-1])
                return f.split("/")[-1] + ":" + (m.group(1) if m else "?")
        return f.split("/")[-1]

    agg = {}
    for f, l, wi, thi, smp in lines:
        a = agg.setdefault(region(f, l), [0.0, 0.0, 0.0])
        a[0] += wi
        a[1] += thi
        a[2] += smp
    ti = sum(a[0] for a in agg.values()) or 1.0
    tt = sum(a[1] for a in agg.values())
    ts = sum(a[2] for a in agg.values()) or 1.0
    print(f"# {path}: {int(ti)} warp instructions, {tt / ti:.2f} of 32 lanes active on average")
    print(f"{'function':44s} {'warp-inst':>9s} {'lanes':>6s} {'stall samples':>13s}")
    for k, a in sorted(agg.items(), key=lambda x: -x[1][0]):
        print(f"{k:44s} {100 * a[0] / ti:8.2f}% {a[1] / max(a[0], 1):6.1f} {100 * a[2] / ts:12.2f}%")


def traffic(path):
    """ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --csv of one launch -> the JSON bench.py reads (profiles/traffic.json)"""
    import json
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr = rows[hi]
    ni, vi, ui, ki = hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit"), hdr.index("Kernel Name")
    mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    tot, kernel, dur, extra = 0.0, "?", None, {}
    for r in rows[hi + 1:]:
        if len(r) <= vi:
            continue
        kernel = r[ki]
        if r[ni].startswith("dram__bytes"):
            tot += float(r[vi].replace(",", "")) * mult[r[ui]]
        elif r[ni].startswith("gpu__time_duration"):
            dur = r[vi] + " " + r[ui]
        elif r[ni] == "sm__inst_executed.avg.per_cycle_elapsed":
            extra["warp_inst_per_cycle_per_sm"] = float(r[vi].replace(",", ""))      # of 4 issue slots per cycle
        elif r[ni] == "smsp__thread_inst_executed_per_inst_executed.ratio":
            extra["active_lanes_per_warp_inst"] = float(r[vi].replace(",", ""))      # of 32
        elif r[ni] == "smsp__issue_active.avg.pct_of_peak_sustained_active":
            extra["issue_active_pct"] = float(r[vi].replace(",", ""))
    print(json.dumps({"pt_wavefront_kernel_bytes_per_launch": tot, **extra, "kernel": kernel, "launch_duration_under_ncu": dur,
                      "source": path, "how": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none, one 256-frame "
                                             "launch of the default bench.py step"}))


def live(path, source_sha, config):
    """ncu --metrics <dram bytes, duration, issue counters> --csv over a short bench.py run -> profiles/ncu_live.json: per kernel
    family the LAST profiled launch (the one inside bench.py's timed region), stamped with the hash of the kernel sources"""
    import json
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr = rows[hi]
    ii, ni, vi, ui, ki = hdr.index("ID"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit"), hdr.index("Kernel Name")
    mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    tms = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}
    by_id = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= vi:
            continue
        d = by_id.setdefault(int(r[ii]), {"kernel": r[ki], "dram_bytes": 0.0})
        v = float(r[vi].replace(",", ""))
        if r[ni].startswith("dram__bytes"):
            d["dram_bytes"] += v * mult[r[ui]]
        elif r[ni].startswith("gpu__time_duration"):
            d["duration_ms"] = v * tms[r[ui]]
        elif r[ni] == "sm__inst_executed.avg.per_cycle_elapsed":
            d["warp_inst_per_cycle_per_sm"] = v          # of 4 issue slots per cycle and SM
        elif r[ni] == "smsp__thread_inst_executed_per_inst_executed.ratio":
            d["active_lanes_per_warp_inst"] = v          # of 32
        elif r[ni] == "smsp__issue_active.avg.pct_of_peak_sustained_active":
            d["issue_active_pct"] = v
        elif r[ni] == "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed":
            d["dram_throughput_pct"] = v
    kernels = {}
    for fam in ("pt_wavefront_kernel", "sv_fused", "fold_kernel", "group_reduce_kernel", "group_scale_kernel", "pt_classify_kernel"):
        ids = [i for i, d in by_id.items() if fam in d["kernel"]]
        if ids:
            d = dict(by_id[ids[-1]])
            d["launch"] = f"profiled launch {len(ids)} of {len(ids)} of this kernel in the run (ncu id {ids[-1]})"
            kernels[fam] = d
    print(json.dumps({"source_sha": source_sha, "config": config, "file": path,
                      "how": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__inst_executed.avg.per_cycle_elapsed,"
                             "smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none "
                             "python bench.py --quick --steps 1 --warmup 3 --no-e2e --no-cpu-baseline (tools/gpu_profile.sh)",
                      "kernels": kernels}, indent=1))


if __name__ == "__main__":
    mode, path = sys.argv[1], sys.argv[2]
    if mode == "traffic":
        traffic(path)
    elif mode == "live":
        live(path, sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else "c2")
    elif mode == "launches":
        launches(path)
    elif mode == "full":
        full(path)
    elif mode == "functions":
        functions(path)
    else:
        source(path, int(sys.argv[3]) if len(sys.argv) > 3 else 40)
