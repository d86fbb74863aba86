"""pytest configuration: the `gpu` marker and shared fixtures.

`-m "not gpu"` : oracle vs golden vectors, host logic, C-ABI symbol check (no compute calls) -- runs anywhere.
`-m gpu`       : parity tests proper, through the C ABI on a real B200.
"""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def crv():
    import cloudrendering_vulkan_b200 as m
    return m


@pytest.fixture(scope="session")
def orc():
    from tests import oracle_binding
    return oracle_binding


@pytest.fixture(scope="session")
def ctx(crv):
    c = crv.Context(0)          # raises CrvError(-2) without a B200: GPU tests fail loudly, never fall back
    yield c
    c.close()
