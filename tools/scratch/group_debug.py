import sys, os, faulthandler
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
faulthandler.dump_traceback_later(40, exit=True)
import numpy as np
import cloudrendering_vulkan_b200 as crv
from cloudrendering_vulkan_b200.multigpu import SampleSplit
P = lambda *a: print(*a, flush=True)
n = 2
wl = crv.Workload("t", 40, 96, 64, 1, 56, 50.0, 0.1, 0.85)
ms = []
for rank in range(n):
    ctx = crv.Context(0)
    r = crv.make_renderer(ctx, wl, cloud_layout="brick")
    r.set_cloud_synthetic(40, 1234)
    ctx.synchronize()
    own = r.resultImage
    split = SampleSplit(r, rank, n, connect=False)
    own.close()
    ms.append((ctx, r, split, crv.Stream(ctx)))
P("members created")
hs = [m[2].group.export() for m in ms]
for m in ms:
    m[2].group.connect(hs)
P("connected")
seeds = crv.msvc_rand_sequence(13)
for i, (ctx, r, split, stream) in enumerate(ms):
    split.image.clear(int(stream)); P(i, "cleared")
    my = split.my_seeds(seeds)
    r.pathTracingTechnique.RenderBatch(my, 0, crv.ACCUM_SUM, int(stream), 0); P(i, "render enqueued")
    split.reduce(len(seeds), int(stream)); P(i, "reduce enqueued")
for i, (ctx, r, split, stream) in enumerate(ms):
    split.group.status(int(stream)); P(i, "status ok")
    img = split.image.download(int(stream)); P(i, "downloaded", float(img.max()))
P("done")
