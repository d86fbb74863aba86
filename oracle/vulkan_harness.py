#!/usr/bin/env python
"""oracle/vulkan_harness.py -- TEST INFRASTRUCTURE: drives oracle/vulkan_harness (vulkan_harness.c).

The oracle north_star names: the reference's own PathTracer.comp / ShadowVolume.comp, compiled to SPIR-V with glslc the way
shaders/Compile.bat does, dispatched by a headless Vulkan program on Mesa lavapipe (or on the GPU's Vulkan driver).  This
module
  * probes for the three things that needs (oracle/vulkan_probe.py: GLSL->SPIR-V compiler, loader, ICD),
  * compiles the shaders from the reference checkout WHERE THEY LIE into oracle/_ref/*.spv (git-ignored),
  * writes the scene file the harness reads (raw parameter-struct bytes, frame seeds, the density grid),
  * runs the harness and returns the shadow volume and the result image -- or the reason it was skipped.
In this round's images all three prerequisites are missing, so `run` returns {"skipped": "..."}; tests/test_ref_shaders.py
checks that the harness builds, that it skips cleanly, and -- on a machine that has Vulkan -- compares it with the CPU oracle.
"""
import importlib.util
import os
import struct
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
BIN = os.path.join(HERE, "vulkan_harness")
REF_SHADERS = os.environ.get("CRV_REFERENCE", "/root/reference") + "/shaders"


def probe():
    spec = importlib.util.spec_from_file_location("vulkan_probe", os.path.join(HERE, "vulkan_probe.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m.probe()


def build_harness():
    if not os.path.exists(BIN):
        subprocess.run(["make", "vulkan_harness"], cwd=HERE, check=True, stdout=subprocess.DEVNULL)
    return BIN


def compile_shaders(compiler):
    """shaders/Compile.bat: glslc <file> -o <file>.spv, for the two shaders of the path; outputs under oracle/_ref/"""
    out_dir = os.path.join(HERE, "_ref")
    os.makedirs(out_dir, exist_ok=True)
    outs = []
    for name in ("ShadowVolume.comp", "PathTracer.comp"):
        src, dst = os.path.join(REF_SHADERS, name), os.path.join(out_dir, name + ".spv")
        if not os.path.exists(dst) or (os.path.exists(src) and os.path.getmtime(src) > os.path.getmtime(dst)):
            if not os.path.exists(src):
                raise FileNotFoundError(f"{src}: the reference checkout is not present and no prebuilt {dst} exists")
            if os.path.basename(compiler).startswith("glslang"):
                subprocess.run([compiler, "-V", "-S", "comp", src, "-o", dst], check=True, stdout=subprocess.DEVNULL)
            else:
                subprocess.run([compiler, "-fshader-stage=compute", src, "-o", dst], check=True)
        outs.append(dst)
    return outs


def _padded(raw: bytes, size: int) -> bytes:
    raw = bytes(raw)
    if len(raw) > size:
        raise ValueError("parameter block larger than its std140 size")
    return raw + b"\0" * (size - len(raw))


def write_scene(path, cloud, counts, camera68, cloud60, params28, shadow152, width, height, n_sv, frames):
    """frames: [(seed, frameCount), ...]; the parameter blocks are the host structs' bytes (UniformBuffers.h), padded to the
    std140 block sizes the shaders declare (80 / 64 / 32 / 160)"""
    nx, ny, nz = (int(c) for c in counts)
    grid = np.ascontiguousarray(cloud, np.float32).reshape(-1)
    assert grid.size == nx * ny * nz
    with open(path, "wb") as f:
        f.write(struct.pack("<8I", 0x53565243, nx, ny, nz, int(width), int(height), int(n_sv), len(frames)))
        f.write(_padded(camera68, 80) + _padded(cloud60, 64) + _padded(params28, 32) + _padded(shadow152, 160))
        for seed, fc in frames:
            f.write(struct.pack("<iI", int(seed), int(fc)))
        f.write(grid.tobytes())


def run(cloud, counts, camera68, cloud60, params28, shadow152, width, height, n_sv, frames, device=None, timeout=3600):
    """returns {"shadow_volume": ndarray n_sv^3, "image": ndarray (H, W, 4), "device": str} or {"skipped": reason}"""
    p = probe()
    if not p["available"]:
        return {"skipped": p["reason"]}
    sv_spv, pt_spv = compile_shaders(p["glsl_compiler"])
    exe = build_harness()
    with tempfile.TemporaryDirectory() as d:
        scene, out = os.path.join(d, "scene.bin"), os.path.join(d, "out.bin")
        write_scene(scene, cloud, counts, camera68, cloud60, params28, shadow152, width, height, n_sv, frames)
        env = dict(os.environ)
        if device:
            env["CRV_VK_DEVICE"] = device
        r = subprocess.run([exe, sv_spv, pt_spv, scene, out], capture_output=True, text=True, timeout=timeout, env=env)
        if r.returncode == 77:
            return {"skipped": r.stderr.strip().splitlines()[-1] if r.stderr.strip() else "skipped"}
        if r.returncode != 0:
            raise RuntimeError(f"vulkan_harness failed ({r.returncode}): {r.stderr[-2000:]}")
        raw = np.fromfile(out, np.float32)
        nsv3 = int(n_sv) ** 3
        return {"shadow_volume": raw[:nsv3].copy(), "image": raw[nsv3:nsv3 + width * height * 4].reshape(height, width, 4).copy(),
                "device": r.stderr.strip().splitlines()[0] if r.stderr.strip() else "?"}


if __name__ == "__main__":
    import json
    print(json.dumps({"probe": probe(), "harness": build_harness()}))
