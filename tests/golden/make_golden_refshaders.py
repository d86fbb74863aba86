"""Golden vectors produced by the REFERENCE'S OWN SHADERS (tests/golden/refshader_*.npz|json).

Source of every number in these fixtures: /root/reference/shaders/PathTracer.comp and ShadowVolume.comp compiled by
g++ where they lie (oracle/glsl_to_cpp.py + oracle/glsl_compat.h -> oracle/_ref/libcrv_refshaders.so), driven with
parameter blocks made by the reference's own setters (UniformBuffers.h through oracle/_ref/libcrv_refstructs.so).
No product code and no hand restatement (crv_oracle.cpp) takes part; the only non-reference input is the synthetic
density grid (the frozen fBm generator, or numpy noise for the ragged case).

    python tests/golden/make_golden_refshaders.py        (build container only: needs /root/reference)
"""
from __future__ import annotations

import json
import os
import struct
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from tests import oracle_binding as orc  # noqa: E402
from tests.refshader_cases import CASES, build_case  # noqa: E402


def main():
    if orc.ref_shaders_lib() is None or orc.ref_structs_lib() is None:
        raise SystemExit("oracle/_ref is not built (needs /root/reference): run `make -C oracle`")
    arrays, meta = {}, {}
    for case in CASES:
        cloud, cam, cp, par, sp = build_case(orc, case)
        sv = orc.refsh_shadow_volume(cloud, cp, sp)
        w, h = orc.image_size(cam)
        img = np.zeros((h, w, 4), np.float32)
        seeds = orc.msvc_rand(case["frames"])
        for f in range(case["frames"]):
            steps, coll = orc.refsh_path_trace(cloud, sv, cam, cp, par, sp, orc.push_constants_bytes(seeds[f], f), img)
        name = case["name"]
        arrays[f"{name}_sv"] = sv
        arrays[f"{name}_image"] = img
        arrays[f"{name}_steps"] = steps
        arrays[f"{name}_coll"] = coll
        meta[name] = {"camera": bytes(cam).hex(), "cloud": bytes(cp).hex(), "parameters": bytes(par).hex(),
                      "shadow": bytes(sp).hex(), "seeds": [int(s) for s in seeds],
                      "steps_total": int(steps.sum()), "collisions_total": int(coll.sum())}
        print(f"{name}: {w}x{h}, sv {sv.size} voxels, last frame {int(steps.sum())} steps / {int(coll.sum())} collisions")
    np.savez_compressed(os.path.join(HERE, "refshader_small.npz"), **arrays)
    json.dump(meta, open(os.path.join(HERE, "refshader_small.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
