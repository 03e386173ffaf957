"""Executable drop-in proof: the reference's own, UNMODIFIED Python wrapper (prox_tv/__init__.py) bound to libproxtv_b200.so.

The reference does `from _prox_tv import ffi, lib` (prox_tv/__init__.py:64) and calls `lib.<symbol>` through `_call` (:70-77).  Here a
`_prox_tv` module is made at test time from the reference's own cdef (the string handed to `ffi.cdef` in prox_tv/prox_tv_build.py:8-77,
read from /root/reference when it exists -- nothing of it is copied into this repository) with cffi in ABI mode,
`lib = ffi.dlopen(libproxtv_b200.so)`, and the reference's __init__.py is imported on top of it.  INTEGRATION.md describes the same
binding.

* without a GPU (this container): every hot-path symbol the wrapper uses must resolve, the wrapper must import, and a call must
  reach this library (which then fails loudly: it has no CPU fallback) -- the output stays the wrapper's `np.zeros`.
* with a GPU *and* the reference tree (neither the build container nor the GPU box has both; kept for maintainers): the hot-path
  subset of the reference's own tests, prox_tv/prox_tv_test.py:7-63,129-216, runs against this library.
"""
import importlib.util
import os
import re
import sys
import types

import numpy as np
import pytest

REF = "/root/reference/prox_tv"
HOT = ["TV1D_denoise", "TV1D_denoise_tautstring", "dp", "PN_TV1", "linearizedTautString_TV1", "classicTautString_TV1",
       "hybridTautString_TV1", "hybridTautString_TV1_custom", "SolveTVConvexQuadratic_a1_nw", "PN_TV1_Weighted",
       "tautString_TV1_Weighted", "SolveTVConvexQuadratic_a1", "PD2_TV", "PD_TV", "DR2_TV", "DR2L1W_TV"]


@pytest.fixture(scope="module")
def ref_wrapper():
    if not os.path.isdir(REF):
        pytest.skip("reference tree not present")
    cffi = pytest.importorskip("cffi")
    import proxtv_b200._lib as L
    if not os.path.exists(L.LIB_PATH):
        pytest.skip("libproxtv_b200.so not built")
    text = open(os.path.join(REF, "prox_tv_build.py")).read()
    cdef = re.search(r'ffi\.cdef\("""(.*?)"""\)', text, re.S).group(1)
    cdef = re.sub(r"typedef struct \{\s*\.\.\.;\s*\} Workspace;", "typedef struct Workspace Workspace;", cdef)     # opaque in ABI mode
    ffi = cffi.FFI()
    ffi.cdef(cdef)
    lib = ffi.dlopen(L.LIB_PATH)
    shim = types.ModuleType("_prox_tv"); shim.ffi = ffi; shim.lib = lib
    saved = sys.modules.get("_prox_tv")
    sys.modules["_prox_tv"] = shim
    try:
        spec = importlib.util.spec_from_file_location("prox_tv_reference_wrapper", os.path.join(REF, "__init__.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        if saved is None:
            sys.modules.pop("_prox_tv", None)
        else:
            sys.modules["_prox_tv"] = saved
    mod._shim_lib = lib
    return mod


def test_reference_wrapper_binds_and_calls_reach_this_library(ref_wrapper):
    import proxtv_b200 as ptv
    for name in HOT:
        assert getattr(ref_wrapper._shim_lib, name) is not None, name        # AttributeError if the symbol did not resolve
    lib = ptv.load()
    x = np.random.default_rng(0).normal(size=50)
    if lib.proxtv_device_count() > 0:
        from oracle import oracle as O
        assert np.abs(ref_wrapper.tv1_1d(x, 0.5) - O.Port().tv1_hybrid(x, 0.5)).max() <= 1e-9
        return
    out = ref_wrapper.tv1_1d(x, 0.5)                      # no device: the library refuses loudly and leaves the wrapper's zeros
    assert out.shape == x.shape and not out.any()
    assert b"no usable CUDA device" in lib.proxtv_last_error()
    Y = np.random.default_rng(1).normal(size=(12, 9))
    out2 = ref_wrapper.tv1_2d(Y, 0.3)                     # DR2_TV through the reference's Fortran-order coercion (:402-404)
    assert out2.shape == Y.shape and out2.flags.f_contiguous and not out2.any()
    assert b"DR2_TV" in lib.proxtv_last_error()


@pytest.mark.gpu
def test_reference_hot_path_tests_against_this_library(ref_wrapper):
    """prox_tv/prox_tv_test.py's hot-path tests, run as written with `prox_tv` = the reference wrapper over this library.
    test_tv1_2d is restricted to the methods of the hot path ('pd', 'dr'); tv2 / tvp tests are out of scope."""
    src = open(os.path.join(REF, "prox_tv_test.py")).read()
    saved = sys.modules.get("prox_tv")
    fake = types.ModuleType("prox_tv")
    for n in ("tv1_1d", "tv1w_1d", "tv2_1d", "tv1_2d", "tvp_1d", "tv1w_2d", "tvp_2d", "tvgen"):
        setattr(fake, n, getattr(ref_wrapper, n))
    sys.modules["prox_tv"] = fake
    try:
        ns = {}
        exec(compile(src, "prox_tv_test.py", "exec"), ns)
        np.random.seed(0)
        for t in ("test_tv1w_1d", "test_tv1w_1d_uniform_weights_small_input", "test_tv1_1d", "test_tv1_1d_int", "test_tv1_tv1w_2d",
                  "test_tv1w_2d_uniform_weights", "test_tv1w_2d_emengd", "test_tvgen_1d", "test_tvgen_2d", "test_tvgen_nd",
                  "test_tvgen_multireg"):
            ns[t]()
        for _ in range(10):                               # the 'pd' / 'dr' half of test_tv1_2d (:106-116)
            x = ns["_generate2d"](); w = 20 * np.random.rand()
            a = ref_wrapper.tv1_2d(x, w, method="dr", max_iters=5000); b = ref_wrapper.tv1_2d(x, w, method="pd", max_iters=5000)
            assert np.allclose(a, b, atol=1e-3)
    finally:
        if saved is None:
            sys.modules.pop("prox_tv", None)
        else:
            sys.modules["prox_tv"] = saved
