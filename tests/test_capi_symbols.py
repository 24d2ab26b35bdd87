"""The C-ABI library must load and export every symbol include/chiml_gpu.h declares (no compute
calls here: this runs without a GPU), and the product path must fail loudly without a device."""
import ctypes
import os
import re

import pytest

from chiml_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "chiml_gpu.h")).read()
    return sorted(set(re.findall(r"\b(chiml_gpu_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    L = capi.lib()._dll          # the raw CDLL: the wrapper tolerates missing symbols of older builds
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/chiml_gpu.h but not exported"
    assert sorted(capi.EXPORTED_SYMBOLS) == names


def test_record_layouts_match_reference_pods():
    from chiml_b200 import plan as P
    assert P.RUN_DTYPE.itemsize == 56      # pair<array<int,6>, array<double,4>>
    assert P.PSI_DTYPE.itemsize == 32      # updatePsiParams
    assert P.GRIDP_DTYPE.itemsize == 32    # updateGridParams
    assert ctypes.sizeof(capi.GridDesc) == 72


def test_no_cpu_fallback_without_device():
    if capi.device_count() > 0:
        pytest.skip("a GPU is visible")
    import util
    with pytest.raises(capi.ChimlError, match="NO_DEVICE"):
        capi.GpuSim(util.load_plan("te_vacuum"))
