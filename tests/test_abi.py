"""CPU tests (-m "not gpu"): the C-ABI library loads, exports every symbol include/proxtv_b200.h declares, and refuses to
compute without a device (no silent CPU fallback).  No compute calls are made here."""
import os
import re

import numpy as np
import pytest

import proxtv_b200
from proxtv_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "proxtv_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return re.findall(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\([^;{}]*\)\s*;", src)


def test_library_exports_every_declared_symbol():
    lib = proxtv_b200.load()
    names = header_functions()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), "missing export: " + n
    assert set(names) == set(_lib.SIGNATURES), set(names) ^ set(_lib.SIGNATURES)


def test_reference_hot_path_symbols_present():
    """The subset of the reference's cffi cdef (prox_tv/prox_tv_build.py:8-77) that is the hot path."""
    lib = proxtv_b200.load()
    for n in ["hybridTautString_TV1", "hybridTautString_TV1_custom", "classicTautString_TV1", "linearizedTautString_TV1",
              "TV1D_denoise", "tautString_TV1_Weighted", "TV", "DR2_TV", "DR2L1W_TV", "PD2_TV", "PD_TV"]:
        assert hasattr(lib, n)


def test_version_and_error_strings():
    lib = proxtv_b200.load()
    assert b"sm_100a" in lib.proxtv_version()
    assert isinstance(_lib.last_error(), str)


def _no_gpu():
    return proxtv_b200.load().proxtv_device_count() == 0


@pytest.mark.skipif(not _no_gpu(), reason="a CUDA device is present")
def test_fails_loudly_without_device():
    x = np.random.default_rng(0).normal(size=(8, 8))
    with pytest.raises(proxtv_b200.ProxTVError):
        proxtv_b200.tv1_2d(x, 0.1)
    with pytest.raises(proxtv_b200.ProxTVError):
        proxtv_b200.tv1_1d(x[0], 0.1)
    # raw C ABI: output untouched, info[RC] = RC_ERROR (3), DR2_TV still returns 0 like the reference
    lib = proxtv_b200.load()
    Y = np.asfortranarray(x); out = np.full_like(Y, -7.0); info = np.zeros(3)
    rc = lib.DR2_TV(8, 8, Y.ctypes.data, 0.1, 0.1, 1.0, 1.0, out.ctypes.data, 1, 0, info.ctypes.data)
    assert rc == 0 and info[2] == 3 and np.all(out == -7.0)
    assert "no usable CUDA device" in _lib.last_error()


def test_python_surface_asserts_like_reference():
    with pytest.raises(AssertionError):
        proxtv_b200.tv1_1d(np.zeros(4), -1.0)
    with pytest.raises(AssertionError):
        proxtv_b200.tv1w_1d(np.zeros(4), np.ones(4))
    with pytest.raises(AssertionError):
        proxtv_b200.tv1_2d(np.zeros((4, 4)), 0.1, method="nope")
    with pytest.raises(AssertionError):
        proxtv_b200.tvgen(np.zeros((4, 4)), [1.0], [1, 2], [1])


def test_header_is_plain_c_and_links(tmp_path):
    """include/proxtv_b200.h is the boundary a C / cgo / cffi binding would consume: it must compile as C99 and link."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if not gcc:
        pytest.skip("gcc not available")
    src = tmp_path / "hdr_check.c"
    src.write_text('#include "proxtv_b200.h"\nint main(void) { return proxtv_device_count() < 0; }\n')
    exe = tmp_path / "hdr_check"
    libdir = os.path.join(ROOT, "proxtv_b200")
    subprocess.check_call([gcc, "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                           "-L", libdir, "-lproxtv_b200", "-Wl,-rpath," + libdir])
    assert subprocess.call([str(exe)]) == 0
