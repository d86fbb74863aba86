"""bench.py keeps the driver's contract: one JSON line on stdout with the required keys, for both arms."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
             "dtype", "data", "config", "e2e", "cpu_baseline"}


def _run(args, timeout):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, timeout=timeout, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, f"stdout must hold exactly one JSON line, got {len(lines)}"
    return json.loads(lines[0])


def test_reference_arm_contract():
    """--impl reference: the reference's CPU implementation (oracle/_ref when present, else the port) on the host cores"""
    d = _run(["--impl", "reference", "--config", "c1_hg", "--steps", "1", "--warmup", "3"], 600)
    assert BASE_KEYS <= set(d) and d["impl"] == "reference" and d["metric"] == "Mpath-samples/s"
    assert d["value"] > 0 and d["higher_is_better"] is True and d["vs_baseline"] is None and d["dtype"] == "f32"
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


@pytest.mark.gpu
def test_our_arm_contract():
    """a shrunk run of the product arm (c1 scene, 4 spp): every key the contract names, with sane values"""
    d = _run(["--config", "c1_hg", "--spp", "4", "--steps", "2", "--warmup", "3"], 900)
    assert BASE_KEYS | {"gpu_launches", "clocks", "roofline", "sv", "parity", "math_modes"} <= set(d)
    assert d["n_gpus"] == 1 and d["value"] > 0 and d["gpu_launches"] > 0 and d["data"] == "synthetic"
    e = d["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] >= 4 * 128 ** 3 and e["d2h_bytes_per_step"] == 512 * 512 * 16
    rf = d["roofline"]
    assert rf["bound"] == "hbm" and rf["unit"] == "GB/s" and rf["peak"] > 1000 and abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-9
    assert d["clocks"]["sm_max_mhz"] and isinstance(d["clocks"]["reasons"], list)
    cb = d["cpu_baseline"]
    assert cb["value"] > 0 and cb["cores"] >= 1 and cb["kind"] in ("reference", "port")
    p = d["parity"]
    assert p["checked_paths"] > 0 and p["threshold_divergent"] <= 0.01 * p["checked_paths"] and p["max_rel_err_matching"] <= 1e-4
    assert d["sv"]["rebuild_ms"] > 0 and set(d["math_modes"]) >= {"strict", "driver"}
