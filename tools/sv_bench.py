#!/usr/bin/env python
"""Shadow-volume rebuild microbenchmark (GPU): ms and achieved algorithmic GB/s for the BASELINE shapes.

  python tools/sv_bench.py [--cases 512:512,1024:1024,500:128] [--layout linear] [--mode two_pass]

algorithmic bytes per rebuild = 4*N_sv^3 (every voxel stored once) + 4*N_cloud^3 (cloud read once).
"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cloudrendering_vulkan_b200 as crv  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", default="512:512,1024:1024,500:128")
    ap.add_argument("--layout", default="linear")
    ap.add_argument("--mode", default="auto", choices=["auto", "generic", "single_pass", "two_pass", "fused", "fused_v1", "fused_rows"])
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--light", default="1,-1,0")
    ap.add_argument("--zero", action="store_true", help="an empty cloud: the build degenerates to storing ones (its floor)")
    args = ap.parse_args()
    try:
        peak = float(json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"])
    except Exception:
        peak = 6650.0
    light = tuple(float(v) for v in args.light.split(","))
    ctx = crv.Context(0)
    for case in args.cases.split(","):
        nsv, nc = (int(v) for v in case.split(":"))
        wl = crv.Workload("svbench", nc, 64, 64, 1, nsv, 50.0, 0.0, 0.85, light)
        r = crv.make_renderer(ctx, wl, cloud_layout=args.layout)
        r.shadowVolumeTechnique.SetTuning(args.mode)
        if args.zero:
            r.set_cloud_host(np.zeros(nc ** 3, np.float32), (nc,) * 3, (1.0 / nc,) * 3)
        else:
            r.set_cloud_synthetic(nc, 1234, 0)
        ts = []
        for i in range(args.reps + 3):
            r.UpdateShadowVolume(0)
            if i >= 3:
                ts.append(r.shadowVolumeTechnique.last_build_ms())
        ms = float(np.median(ts))
        b = 4.0 * nsv ** 3 + 4.0 * nc ** 3
        print(json.dumps({"n_sv": nsv, "n_cloud": nc, "layout": args.layout, "mode": r.shadowVolumeTechnique.last_build_mode(),
                          "ms_median": ms, "ms_min": float(min(ts)), "algorithmic_gb": b / 1e9, "achieved_gbs": b / ms / 1e6,
                          "frac_of_peak": b / ms / 1e6 / peak, "peak_gbs": peak}), flush=True)
        r.close()
    ctx.close()


if __name__ == "__main__":
    main()
