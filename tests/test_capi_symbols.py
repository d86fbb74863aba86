"""The C-ABI library loads without a GPU and exports every symbol include/crv_c_api.h declares (no compute calls)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "crv_c_api.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(crv_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound(crv):
    names = declared_symbols()
    assert len(names) > 60
    lib = ctypes.CDLL(crv.capi.LIB_PATH)
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f"declared in crv_c_api.h but not exported: {missing}"
    unbound = [n for n in names if n not in crv.capi.SIGNATURES]
    assert not unbound, f"declared in crv_c_api.h but not bound in capi.py: {unbound}"
    extra = [n for n in crv.capi.SIGNATURES if n not in names]
    assert not extra, f"bound in capi.py but not declared in the header: {extra}"


def test_library_is_sm100a_only(crv):
    """the product carries sm_100a SASS and nothing else (no PTX fallback for other targets)"""
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "-lelf", crv.capi.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_no_device_fails_loudly(crv):
    """without a usable B200 the context constructor raises CRV_ERR_NO_DEVICE: there is no CPU fallback"""
    if crv.capi.lib.crv_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(crv.CrvError) as e:
        crv.Context(0)
    assert e.value.code == crv.capi.CRV_ERR_NO_DEVICE
    assert "no CPU fallback" in e.value.message


def test_null_handles_return_status_not_crash(crv):
    lib = crv.capi.lib
    assert lib.crv_ctx_synchronize(None, None) == crv.capi.CRV_ERR_INVALID
    assert lib.crv_pt_record_draw_commands(None, None, 0) == crv.capi.CRV_ERR_INVALID
    assert lib.crv_sv_record_draw_commands(None, None, 0) == crv.capi.CRV_ERR_INVALID
    assert lib.crv_volume_upload(None, None, 0, None) == crv.capi.CRV_ERR_INVALID
    assert b"null" in lib.crv_last_error()


def test_product_never_touches_the_oracle():
    """only tests/, __graft_entry__.smoke() and bench.py may use oracle/: the package must not mention it"""
    pkg = os.path.join(ROOT, "cloudrendering_vulkan_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h", ".cuh")) or f == "Makefile":
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "crv_oracle" not in text and "oracle_binding" not in text and "libcrv_oracle" not in text, f
