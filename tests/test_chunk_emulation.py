"""CPU test (-m "not gpu") of the product's chunk-stitching logic: tests/emu/emu_chunked.cu runs the per-lane phases of
proxtv_b200/csrc/chunk_core.cuh (the code the CUDA kernel executes) lane by lane on the host; the result must be
BIT-IDENTICAL to the oracle's sequential linearized / weighted taut-string for every input, including the adversarial ones
for speculation: flat fibers (no breaks at all), ramps, huge lambda (one segment spanning all chunks), integer ties."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
EMU_DIR = os.path.join(HERE, "emu")
dp = C.POINTER(C.c_double)


@pytest.fixture(scope="module")
def emu():
    so = os.path.join(EMU_DIR, "libptv_emu.so")
    src = os.path.join(EMU_DIR, "emu_chunked.cu")
    core = [os.path.join(HERE, "..", "proxtv_b200", "csrc", f) for f in ("chunk_core.cuh", "taut_scan.cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(f) > os.path.getmtime(so) for f in [src] + core):
        nvcc = "/usr/local/cuda/bin/nvcc"
        if not os.path.exists(nvcc):
            pytest.skip("nvcc not available")
        subprocess.check_call([nvcc, "-O2", "-std=c++17", "-arch=sm_100a", "-fmad=false", "--expt-relaxed-constexpr",
                               "-Xcompiler", "-fPIC,-ffp-contract=off", "-shared", "-x", "cu", src, "-o", so])
    lib = C.CDLL(so)
    lib.emu_chunked_f64.argtypes = [dp, C.c_int, C.c_double, dp, dp, C.c_int, C.POINTER(C.c_int)]
    lib.emu_check_table_division.argtypes = [C.c_long, C.c_ulonglong]; lib.emu_check_table_division.restype = C.c_long

    lib.emu_chunked_f64_ch16.argtypes = lib.emu_chunked_f64.argtypes

    def run(y, lam, w=None, out_op=0, ch16=False):
        y = np.ascontiguousarray(y, dtype=np.float64); x = np.empty_like(y); r = C.c_int(0)
        wp = None
        if w is not None:
            w = np.ascontiguousarray(w, dtype=np.float64); wp = w.ctypes.data_as(dp)
        fn = lib.emu_chunked_f64_ch16 if ch16 else lib.emu_chunked_f64
        rc = fn(y.ctypes.data_as(dp), y.size, float(lam), wp, x.ctypes.data_as(dp), out_op, C.byref(r))
        assert rc == 0, "emulated CTA broke the barrier discipline (rc=%d)" % rc
        return x, r.value
    run.lib = lib
    return run


def cases():
    rng = np.random.default_rng(42)
    out = []
    for n in [1, 2, 3, 31, 32, 33, 63, 64, 65, 100, 257, 1000, 4096]:
        for lam in [0.0, 0.05, 0.2, 1.0, 5.0, 50.0, 1e4]:
            out.append((np.repeat(rng.normal(0, 1, n // 16 + 1), 16)[:n] + rng.normal(0, 0.3, n), lam))
    out.append((np.full(777, 2.5), 0.3))                                   # flat: no break anywhere
    out.append((np.zeros(100), 0.0))
    out.append((np.linspace(-3, 3, 1500), 0.4))                            # ramp
    out.append((np.sin(np.linspace(0, 20, 3000)) * 3, 0.3))                # smooth
    out.append((np.round(rng.normal(0, 3, 2000)), 2.0))                    # integer data: exact ties
    out.append((np.concatenate([np.zeros(700), rng.normal(0, 1, 300), np.ones(900)]), 0.2))   # flat / noisy / flat
    out.append((np.concatenate([rng.normal(0, 1, 500), np.full(1500, 0.3), rng.normal(0, 1, 500)]), 0.5))
    out.append((np.where(np.arange(3000) % 2 == 0, 1.0, -1.0), 0.6))       # alternating
    out.append((1e6 + rng.normal(0, 1, 2000), 0.3))                        # large offset
    return out


def test_emulated_cta_is_bit_identical_to_sequential_scan(emu, port):
    worst_rounds = 0
    for y, lam in cases():
        x, r = emu(y, lam)
        assert np.array_equal(x, port.tv1_linearized(y, lam)), (y.size, lam)
        worst_rounds = max(worst_rounds, r)
    assert worst_rounds >= 3        # the adversarial cases really exercised the multi-round path


def test_emulated_sparse_result_expansion(emu, port):
    """The sparse result (segment values at their starts + per-window start mask and entering value) expanded the way the fused
    scatter does it equals the dense result, for 32- and 16-sample chunks."""
    rng = np.random.default_rng(9)
    for y, lam in cases():
        x, _ = emu(y, lam, out_op=-1)
        assert np.array_equal(x, port.tv1_linearized(y, lam)), (y.size, lam)
        if y.size >= 2:
            w = rng.uniform(0, 2 * max(lam, 0.1), y.size - 1)
            x, _ = emu(y, 0.0, w, out_op=-1, ch16=True)
            assert np.array_equal(x, port.tv1_weighted(y, w)), (y.size, lam, "ch16")


def test_emulated_weighted_and_output_ops(emu, port):
    rng = np.random.default_rng(7)
    for y, lam in cases():
        if y.size < 2:
            continue
        w = rng.uniform(0, 2 * max(lam, 0.1), y.size - 1)
        if y.size > 50:
            w[10:40] = 0.0                                                 # zero weights: free jumps
        x, _ = emu(y, 0.0, w)
        assert np.array_equal(x, port.tv1_weighted(y, w)), (y.size, lam)
        x, _ = emu(y, 0.0, w, ch16=True)                                  # 16-sample chunks (weighted float64 kernel)
        assert np.array_equal(x, port.tv1_weighted(y, w)), (y.size, lam, "ch16")
        x, _ = emu(y, lam, ch16=True)
        assert np.array_equal(x, port.tv1_linearized(y, lam)), (y.size, lam, "ch16 unweighted")
    y, lam = cases()[40]
    x = port.tv1_linearized(y, lam)
    assert np.array_equal(emu(y, lam, out_op=1)[0], 2 * (y - x) - y)       # fused DR reflection
    assert np.array_equal(emu(y, lam, out_op=2)[0], y - x)                 # fused prox difference


def test_typical_data_needs_one_round(emu):
    from oracle import oracle as O
    y = O.gen_cfg2(4096, 4, seed=1)[:, 0]
    _, r = emu(y, 0.2)
    assert r <= 3


def test_table_division_is_exact(emu):
    """q = a*r ; q' = fma(fma(-q, d, a), r, q) with r = RN(1/d) must equal IEEE a/d for every table divisor (f64 and f32)."""
    assert emu.lib.emu_check_table_division(200000, 12345) == 0
