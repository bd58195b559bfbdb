import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _cuda_devices() -> int:
    """Devices as the product library sees them (oz_device_count); 0 when the library is missing or CUDA is absent."""
    try:
        import ctypes
        from othellozero_b200 import _lib
        n = ctypes.c_int32(0)
        return n.value if _lib.load().oz_device_count(ctypes.byref(n)) == 0 else 0
    except Exception:
        return 0


def pytest_collection_modifyitems(config, items):
    """A plain `pytest` on a CUDA-less box reports the GPU tests as skipped, not failed (so CPU regressions stay
    visible).  `-m gpu` on the GPU box is unaffected: there the device exists."""
    if not any("gpu" in it.keywords for it in items) or _cuda_devices() > 0:
        return
    skip = pytest.mark.skip(reason="no CUDA device: the product path has no CPU fallback")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


def load_golden(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def golden_rules():
    return load_golden("rules.json")


@pytest.fixture(scope="session")
def golden_playouts():
    return load_golden("playouts.json")


@pytest.fixture(scope="session")
def golden_episodes():
    return load_golden("episodes.json")


@pytest.fixture(scope="session")
def golden_roots():
    return load_golden("roots.json")


@pytest.fixture(scope="session")
def golden_rng_episodes():
    return load_golden("episodes_rng.json")


@pytest.fixture(scope="session")
def golden_examples():
    return load_golden("examples.json")
