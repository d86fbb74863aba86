#!/usr/bin/env python
"""Where the end-to-end step of bench.py spends its time (run on the GPU box): each stage of the e2e step is
bracketed by a stream synchronize + perf_counter, 5 repetitions, medians printed as one JSON line."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cloudrendering_vulkan_b200 as crv  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "c2"
    wl = crv.CONFIGS[name]
    ctx = crv.Context(0)
    stream = torch.cuda.current_stream().cuda_stream
    r = crv.make_renderer(ctx, wl, cloud_layout="brick")
    r.set_cloud_synthetic(wl.n_cloud, wl.fbm_seed, stream)
    ctx.synchronize(stream)
    W, H = wl.width, wl.height
    host_grid = torch.empty(wl.n_cloud ** 3, dtype=torch.float32).pin_memory()
    host_img = torch.empty((H, W, 4), dtype=torch.float32).pin_memory()
    host_grid.numpy()[:] = r.cloudImage.download(stream)
    hg, hi = host_grid.numpy(), host_img.numpy()
    seeds = crv.msvc_rand_sequence(wl.spp)
    pt = r.pathTracingTechnique
    stages = {}

    def timed(label, fn):
        ctx.synchronize(stream)
        t0 = time.perf_counter()
        fn()
        ctx.synchronize(stream)
        stages.setdefault(label, []).append((time.perf_counter() - t0) * 1e3)

    for _ in range(6):
        timed("h2d_upload+retile", lambda: r.cloudImage.upload(hg, stream))
        maj = []
        timed("majorant+brick_tables", lambda: maj.append(r.cloudImage.majorant(stream)))
        timed("set_props+sv_rebuild", lambda: (r.SetCloudProperties((wl.n_cloud,) * 3, (1.0 / wl.n_cloud,) * 3, maj[0]),
                                               r.UpdateCloudData(stream)))
        timed("render", lambda: pt.RenderBatch(seeds, 0, crv.ACCUM_RUNNING_MEAN, stream, 0))
        timed("d2h_image", lambda: r.resultImage.download(stream, out=hi))
    out = {k: float(np.median(v[1:])) for k, v in stages.items()}
    out["total_ms"] = float(sum(out.values()))
    out["h2d_GBps"] = 4.0 * wl.n_cloud ** 3 / (out["h2d_upload+retile"] * 1e-3) / 1e9
    print(json.dumps({"workload": name, **out}))


if __name__ == "__main__":
    main()
