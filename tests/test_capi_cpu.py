"""The C-ABI library: it must build, load and export exactly what include/pydsb.h declares.
No compute call is made here (there is no GPU on the CPU box and no CPU fallback)."""
import ctypes
import os
import re

import pytest

from pyds_b200 import capi, models

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "pydsb.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pydsb_[a-z_]+)\s*\(", text)))


def test_library_builds_and_exports_every_declared_symbol():
    lib = capi.load()
    names = _declared_symbols()
    assert len(names) >= 11
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/pydsb.h but not exported"
    assert sorted(capi.EXPORTS) == names
    assert lib.pydsb_version() == 1


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU box: covered by the gpu tests")
    lm = models.kinetics("twostage").lower()
    with pytest.raises(capi.PydsbError, match="no CUDA device"):
        capi.Engine(lm.blob, 0)
    with pytest.raises(capi.PydsbError):
        capi.dfma_peak(0, 1)


def test_bad_blob_rejected_before_touching_the_device():
    lib = capi.load()
    h = ctypes.c_void_p()
    rc = lib.pydsb_model_create(b"\0" * 8, 8, 0, ctypes.byref(h))
    assert rc == -1 and b"blob" in lib.pydsb_last_error()


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "pyds_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f"{f} imports oracle/"
