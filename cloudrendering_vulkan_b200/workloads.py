"""The BASELINE.json workload table (SURVEY.md section 8d "Per-config parameters") and the frozen synthetic-scene parameters.

Plain data, no native code: bench.py's reference arm loads this file by path so that timing the CPU side never maps
libcrv_b200.so into the process.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional


@dataclass
class Workload:
    """One BASELINE.json configuration (SURVEY.md section 8d "Per-config parameters")."""
    name: str
    n_cloud: int
    width: int
    height: int
    spp: int
    n_sv: int
    density_scaling: float
    absorption: float
    phase_g: Optional[float]          # None = reference default (isotropic)
    light_dir: tuple = (1.0, -1.0, 0.0)
    fov: float = 90.0
    fbm_seed: int = 1234


CONFIGS = {
    # reference-pure semantics: CPU-runnable parity case
    "c1_iso": Workload("c1_iso", 128, 512, 512, 16, 500, 50.0, 0.0, None),
    "c1_hg": Workload("c1_hg", 128, 512, 512, 16, 500, 50.0, 0.0, 0.85),
    # the headline: 512^3 cloud, 1080p, 256 spp, albedo 0.9, HG g = 0.85, one B200
    "c2": Workload("c2", 512, 1920, 1080, 256, 512, 50.0, 0.1, 0.85),
    # 1024^3 cloud, 4K, 1024 spp split across GPUs
    "c3": Workload("c3", 1024, 3840, 2160, 1024, 1024, 50.0, 0.1, 0.85),
    # long-path stress: dense cloud, albedo 0.99
    "c4": Workload("c4", 512, 1920, 1080, 64, 512, 200.0, 0.01, 0.85),
    # animated sun: SV rebuild + 64 spp per frame
    "c5": Workload("c5", 1024, 1920, 1080, 64, 1024, 50.0, 0.1, 0.85),
}
