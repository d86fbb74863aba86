"""Exploration harness (not the contract bench): times the SV build and the path tracer at BASELINE shapes for
every storage layout / tuning knob and prints one JSON line per measurement.  Used under gpurun to choose
defaults; the numbers that count come from bench.py.

  python tools/explore.py --config c2 --spp 16 --layouts linear,brick,tex3d --thresholds 8,12,16,24
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

import cloudrendering_vulkan_b200 as crv  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c2")
    ap.add_argument("--spp", type=int, default=16)
    ap.add_argument("--layouts", default="linear,brick,tex3d")
    ap.add_argument("--sv-layouts", default="")
    ap.add_argument("--thresholds", default="12")
    ap.add_argument("--ctas", default="0")
    ap.add_argument("--n-cloud", type=int, default=0)
    ap.add_argument("--n-sv", type=int, default=0)
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--sv-reps", type=int, default=3)
    args = ap.parse_args()

    wl = crv.CONFIGS[args.config]
    n_cloud = args.n_cloud or wl.n_cloud
    n_sv = args.n_sv or wl.n_sv
    ctx = crv.Context(0)
    peak = 6451.5
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    sv_layouts = [s for s in args.sv_layouts.split(",") if s]
    for layout in args.layouts.split(","):
        for svl in (sv_layouts or [layout]):
            r = crv.make_renderer(ctx, wl, cloud_layout=layout, sv_layout=svl, n_cloud=n_cloud, n_sv=n_sv)
            t0 = time.perf_counter()
            r.set_cloud_synthetic(n_cloud, wl.fbm_seed)
            ctx.synchronize()
            setup_s = time.perf_counter() - t0
            # ---- shadow-volume rebuild ----
            sv_ms = []
            for _ in range(args.sv_reps):
                r.shadowVolumeTechnique.RecordDrawCommands(0, 0)
                sv_ms.append(r.shadowVolumeTechnique.last_build_ms())
            sv_bytes = 4.0 * n_sv ** 3 + 4.0 * n_cloud ** 3
            best = min(sv_ms)
            print(json.dumps({"what": "sv_build", "cloud_layout": layout, "sv_layout": svl, "n_cloud": n_cloud, "n_sv": n_sv,
                              "ms": [round(x, 4) for x in sv_ms], "gbs": round(sv_bytes / best / 1e6, 1),
                              "frac_of_measured_hbm": round(sv_bytes / best / 1e6 / peak, 4), "setup_s": round(setup_s, 2)}), flush=True)
            # ---- path tracing ----
            for ctas in [int(x) for x in args.ctas.split(",")]:
                for thr in [int(x) for x in args.thresholds.split(",")]:
                    r.pathTracingTechnique.SetTuning(ctas_per_sm=ctas, track_threshold=thr)
                    seeds = crv.msvc_rand_sequence(args.spp)
                    times, kms = [], []
                    for rep in range(args.reps):
                        r.reset_accumulation()
                        r.pathTracingTechnique.ResetStats()
                        ctx.synchronize()
                        t0 = time.perf_counter()
                        r.render(args.spp, seeds=seeds)
                        ctx.synchronize()
                        times.append(time.perf_counter() - t0)
                        kms.append(r.pathTracingTechnique.LastRenderMs()[0])
                    st = r.pathTracingTechnique.GetStats()
                    paths = wl.width * wl.height * args.spp
                    tbest = min(times)
                    alg_bytes = 32.0 * (st.tracking_steps + st.collisions) + 32.0 * st.paths
                    img = r.resultImage.download()
                    print(json.dumps({
                        "what": "pt", "cloud_layout": layout, "sv_layout": svl, "ctas_per_sm": ctas, "track_threshold": thr,
                        "spp": args.spp, "wall_s": [round(t, 4) for t in times], "tracker_ms": [round(k, 2) for k in kms],
                        "mpaths_s": round(paths / tbest / 1e6, 1), "steps_per_path": round(st.tracking_steps / st.paths, 2),
                        "collisions_per_path": round(st.collisions / st.paths, 3),
                        "fetch_fraction": round(st.density_fetches / max(1, st.tracking_steps), 4),
                        "gsteps_s": round(st.tracking_steps / tbest / 1e9, 2),
                        "alg_gbs": round(alg_bytes / tbest / 1e9, 1), "frac_of_measured_hbm": round(alg_bytes / tbest / 1e9 / peak, 4),
                        "img_mean": [round(float(x), 5) for x in img.reshape(-1, 4).mean(0)],
                        "finite": bool(np.isfinite(img).all())}), flush=True)
            r.close()
    ctx.close()


if __name__ == "__main__":
    main()
