"""The C-ABI library: loads, exports every entry point include/pvtrace_b200.h declares, and fails LOUDLY (no CPU
fallback) when no CUDA device is usable."""
import ctypes
import os
import re

import numpy as np
import pytest

import pvtrace_b200 as pv
from pvtrace_b200.engine import _cuda
from tests import scenes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_entry_points():
    text = open(os.path.join(ROOT, "include", "pvtrace_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pvt_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(_cuda.LIB_PATH)
    names = declared_entry_points()
    assert len(names) >= 20
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/pvtrace_b200.h but not exported"
    assert set(names) == set(_cuda.EXPORTED_SYMBOLS)


def test_version_and_struct_layout():
    lib = _cuda.load_library()
    assert lib.pvt_version() == 200
    # the ctypes mirrors must have the compiled structs' sizes (load_library() enforces it as well)
    sizes = (ctypes.c_int32 * 4)()
    lib.pvt_struct_sizes(sizes)
    assert list(sizes) == [ctypes.sizeof(_cuda.PvtScene), ctypes.sizeof(_cuda.PvtEmit), ctypes.sizeof(_cuda.PvtParams),
                           ctypes.sizeof(_cuda.PvtOut)]
    assert ctypes.sizeof(_cuda.PvtOut) == 18 * 8


@pytest.mark.skipif(_cuda.device_count() > 0, reason="a CUDA device is present")
def test_compute_entries_fail_loudly_without_a_gpu():
    assert not pv.engine.is_available()
    scene = scenes.fresnel()
    with pytest.raises(_cuda.LibraryError, match="no CUDA device|CUDA"):
        pv.engine.simulate(scene, 10, seed=1)
    out = np.zeros(4)
    status = _cuda.load_library().pvt_test_fresnel_reflectivity(4, _cuda._vp(np.zeros(4)), _cuda._vp(np.ones(4)),
                                                                _cuda._vp(np.ones(4)), _cuda._vp(out), 0)
    assert status != 0 and b"CUDA" in _cuda.load_library().pvt_last_error()


def test_product_never_imports_the_oracle():
    """The shipped package must not reach into oracle/ (test infrastructure)."""
    for folder, _, files in os.walk(os.path.join(ROOT, "pvtrace_b200")):
        for name in files:
            if name.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(folder, name)).read()
                assert "pvt_oracle" not in text and "from oracle" not in text and "import oracle" not in text, name
