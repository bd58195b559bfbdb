"""CPU-side checks of the boundary: the library loads without a GPU, exports every symbol that
include/oz_b200.h declares, and refuses to compute without a device (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def L():
    from othellozero_b200 import build, _lib
    build.build()
    return _lib.load()


def test_exports_every_declared_symbol(L):
    hdr = open(os.path.join(ROOT, "include", "oz_b200.h")).read()
    names = set(re.findall(r"\b(oz_[a-z0-9_]+)\s*\(", hdr))
    from othellozero_b200 import _lib
    assert names == set(_lib.SIGNATURES), names ^ set(_lib.SIGNATURES)
    for n in names:
        assert hasattr(L, n), n
    assert L.oz_abi_version() == 1


def test_no_cpu_fallback(L):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from othellozero_b200 import _lib, engine
    n = C.c_int32(-1)
    assert L.oz_device_count(C.byref(n)) == _lib.OZ_ERR_CUDA
    with pytest.raises(_lib.OzError):
        engine.Engine(8, 1, 64)
    with pytest.raises(_lib.OzError):
        engine.legal_moves([1], [2], 8)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "othellozero_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "oz_oracle" not in src, f
