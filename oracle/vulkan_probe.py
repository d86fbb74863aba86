#!/usr/bin/env python
"""oracle/vulkan_probe.py -- TEST INFRASTRUCTURE.  Is the oracle the north star names available on this machine?

The north star's oracle is the reference's own compute shaders compiled to SPIR-V and dispatched from a headless Vulkan harness
on Mesa lavapipe.  That needs (1) a GLSL -> SPIR-V compiler (the reference ships no .spv: shaders/Compile.bat
calls %VULKAN_SDK%/Bin/glslc.exe and *.spv is git-ignored), (2) the Vulkan loader and (3) an ICD (lavapipe, or the GPU's).
This probe checks the three and prints one JSON line; it never raises.  In the build image and on the GPU boxes of this
round all three are absent, which is why oracle/_ref compiles the same shader SOURCES with g++ instead (glsl_to_cpp.py +
glsl_compat.h) -- same text, same evaluation order, driver-defined precision fixed as crv_oracle.h states.

    python oracle/vulkan_probe.py   ->   {"available": false, "glsl_compiler": null, "loader": null, "icds": [], "reason": "..."}
"""
import ctypes
import glob
import json
import os
import shutil


def probe():
    compiler = shutil.which("glslc") or shutil.which("glslangValidator")
    loader = None
    for name in ("libvulkan.so.1", "libvulkan.so"):
        try:
            ctypes.CDLL(name)
            loader = name
            break
        except OSError:
            continue
    icds = []
    for d in ("/usr/share/vulkan/icd.d", "/etc/vulkan/icd.d", os.path.expanduser("~/.local/share/vulkan/icd.d")):
        icds += sorted(glob.glob(os.path.join(d, "*.json")))
    if os.environ.get("VK_ICD_FILENAMES"):
        icds += [p for p in os.environ["VK_ICD_FILENAMES"].split(":") if os.path.exists(p)]
    missing = [n for n, v in (("GLSL->SPIR-V compiler (glslc / glslangValidator)", compiler), ("Vulkan loader (libvulkan.so.1)", loader),
                              ("Vulkan ICD (lavapipe or the GPU driver's)", icds)) if not v]
    return {"available": not missing, "glsl_compiler": compiler, "loader": loader, "icds": icds,
            "reason": None if not missing else "skipped: no " + "; no ".join(missing)}


if __name__ == "__main__":
    print(json.dumps(probe()))
