"""CPU-side checks of the drop-in boundary: libsage2gpu.so loads without a GPU and exports every
symbol include/sage2gpu.h declares; without a device the product fails loudly (no CPU fallback)."""
import ctypes
import os
import re

import pytest

from sage2_b200 import api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "sage2gpu.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sage2gpu_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    if not os.path.exists(api.LIB_PATH):
        api.build_library()
    lib = ctypes.CDLL(api.LIB_PATH)
    names = _declared()
    assert len(names) >= 16
    for n in names:
        assert hasattr(lib, n), n
    assert sorted(api.EXPORTS) == names


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(api.Sage2GpuError):
        api.Sage2Gpu(0)


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under sage2_b200/ may reference it."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "sage2_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", "Makefile")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "liboracle" not in text and "sgo_" not in text and "from oracle" not in text, f
