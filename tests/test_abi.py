"""The C-ABI library loads and exports every symbol include/lsb200.h declares (no GPU needed)."""
import os
import re

import pytest

from loopstructural_b200 import _native as nat
from loopstructural_b200 import build as lsbuild

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    lsbuild.build()
    return nat.load()


def _declared():
    src = open(os.path.join(ROOT, "include", "lsb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(lsb200_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported_and_bound(lib):
    names = _declared()
    assert len(names) >= 25
    for name in names:
        assert hasattr(lib, name), f"{name} declared in lsb200.h but not exported"
        assert name in nat.SIGNATURES, f"{name} has no ctypes signature"
    assert sorted(nat.SIGNATURES) == names


def test_abi_version_and_error_string(lib):
    assert lib.lsb200_abi_version() == 1
    assert isinstance(lib.lsb200_last_error(), bytes)


def test_compute_fails_loudly_without_gpu(lib):
    """No CPU fallback: creating a handle without a CUDA device is an error, not a slow path."""
    import numpy as np

    if lib.lsb200_device_count() > 0:
        pytest.skip("a CUDA device is present")
    from loopstructural_b200 import Engine

    with pytest.raises(nat.Lsb200Error) as ei:
        Engine([4, 4, 4], np.zeros(3), np.ones(3))
    assert "no CUDA device" in str(ei.value)
