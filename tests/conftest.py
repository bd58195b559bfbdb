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
