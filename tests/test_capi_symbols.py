"""The C-ABI library builds, loads and exports every symbol include/dmc_b200.h declares.
No compute call is made here (there is no GPU on the build host)."""
import ctypes
import os
import re

import pytest
import torch

from helpers import D, ROOT

HEADER = os.path.join(ROOT, "include", "dmc_b200.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dmci?_[a-z0-9_]+)\s*\(", src)))


def test_header_functions_are_exported():
    lib = D._capi.load()
    names = declared_functions()
    assert len(names) >= 17
    for n in names:
        assert hasattr(lib, n), f"{n} declared in dmc_b200.h but not exported"
        assert n in D._capi.SIGNATURES, f"{n} has no ctypes signature in _capi.py"
    assert set(D._capi.SIGNATURES) == set(names)


def test_library_is_in_tree_and_sm100a():
    path = D._capi.library_path()
    assert os.path.exists(path) and path.startswith(ROOT)
    assert b"sm_100a" in D._capi.load().dmc_version()


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    lib = D._capi.load()
    h = ctypes.c_void_p()
    rc = lib.dmc_create(0, 1, 64, 64, 0, ctypes.byref(h))
    assert rc != 0 and not h.value
    assert b"CUDA" in lib.dmc_last_error(None)
    m = D.DMC_old().eval()
    with torch.no_grad(), pytest.raises(RuntimeError, match="no CPU path"):
        m(torch.rand(1, 3, 64, 64), 32, {"frame": torch.rand(1, 3, 64, 64), "feature": None})


def test_create_rejects_bad_geometry():
    lib = D._capi.load()
    h = ctypes.c_void_p()
    assert lib.dmc_create(0, 1, 100, 64, 0, ctypes.byref(h)) != 0
    assert lib.dmc_create(9, 1, 64, 64, 0, ctypes.byref(h)) != 0
    assert lib.dmc_create(0, 0, 64, 64, 0, ctypes.byref(h)) != 0
