"""Generates tests/golden/golden_v1.npz from the UNMODIFIED reference (oracle/_ref/libproxtv_ref.so, compiled from
/root/reference/src by oracle/build_ref.sh).  The reference ships no golden vectors, so these
fixtures -- inputs (seeded) and the reference's outputs -- are the portable pin for the oracle port and the CUDA path.

Run from the repo root in a container that has /root/reference:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402


def main():
    R = O.Ref()
    out = {}
    rng = np.random.default_rng(20260924)
    # ---- 1D: (name, y, lam) ----
    cases = []
    for n, lam, scale in [(1, 0.5, 1), (2, 0.3, 1), (3, 10.0, 1), (17, 0.0, 1), (64, 0.2, 0.3), (200, 1.0, 1.0),
                          (257, 5.0, 1.0), (1000, 0.5, 0.5), (1000, 50.0, 0.5), (4096, 0.2, 0.3)]:
        y = np.repeat(rng.normal(0, 2, n // 16 + 1), 16)[:n] + rng.normal(0, scale, n)
        cases.append((y, lam))
    cases.append((np.round(rng.normal(0, 3, 300)), 2.0))          # integer data: exact ties
    cases.append((np.full(100, 1.25), 0.7))                       # constant
    cases.append((np.linspace(-3, 3, 500), 0.4))                  # ramp
    cases.append((np.sin(np.linspace(0, 20, 2000)) * 3, 0.3))     # smooth: long hull
    for k, (y, lam) in enumerate(cases):
        out["d1_%02d_y" % k] = y
        out["d1_%02d_lam" % k] = np.float64(lam)
        out["d1_%02d_hybrid" % k] = R.tv1_hybrid(y, lam)
        out["d1_%02d_linearized" % k] = R.tv1_linearized(y, lam)
        out["d1_%02d_classic" % k] = R.tv1_classic(y, lam)
        out["d1_%02d_condat" % k] = R.tv1_condat(y, lam)
        if y.size >= 2:
            w = rng.uniform(0.0, 2 * max(lam, 0.1), y.size - 1)
            out["d1_%02d_w" % k] = w
            out["d1_%02d_weighted" % k] = R.tv1_weighted(y, w)
    out["d1_count"] = np.int64(len(cases))
    # cfg-1 style long signal, stored as a digest only (sum, sumsq, jump count, 64 samples) to keep the file small
    y = O.gen_cfg1(200_000, seed=0)
    x = R.tv1_hybrid(y, 0.5)
    out["cfg1_200k_digest"] = np.array([x.sum(), (x * x).sum(), np.count_nonzero(np.diff(x))])
    out["cfg1_200k_samples"] = x[:: 200_000 // 64][:64]
    # ---- 2D DR / PD2 ----
    for k, (M, N, lam, it) in enumerate([(2, 2, 0.3, 0), (3, 5, 0.5, 0), (37, 53, 0.2, 0), (64, 48, 0.2, 7), (96, 128, 1.0, 0)]):
        Y = O.gen_cfg2(M, N, seed=100 + k, block=8)
        o, info = R.dr2_tv(Y, lam, maxit=it)
        out["dr_%d_Y" % k] = Y; out["dr_%d_lam" % k] = np.float64(lam); out["dr_%d_it" % k] = np.int64(it)
        out["dr_%d_out" % k] = o; out["dr_%d_info" % k] = info
        o, info = R.pd2_tv(Y, [lam, 1.5 * lam], [1, 2], maxit=it)
        out["pd2_%d_out" % k] = o; out["pd2_%d_info" % k] = info
    out["dr_count"] = np.int64(5)
    # ---- ND PD ----
    V = O.gen_cfg4((24, 20, 12), seed=2, block=4)
    o, info = R.pd_tv(V, [0.2, 0.2, 0.2], [1, 2, 3])
    out["pd3_V"] = V; out["pd3_out"] = o; out["pd3_info"] = info
    V4 = np.asfortranarray(rng.normal(0, 1, (6, 5, 4, 3)))
    o, info = R.pd_tv(V4, [0.3, 0.1, 0.2, 0.4, 0.05], [1, 2, 3, 4, 2])
    out["pd4_V"] = V4; out["pd4_out"] = o; out["pd4_info"] = info
    o, info = R.pd_tv(cases[5][0], [0.7], [1])                     # 1 term on a 1D signal (tvgen_1d)
    out["pd1_out"] = o; out["pd1_info"] = info
    # ---- weighted 2D DR (DR2L1W_TV); its own generator so that everything above keeps its draws ----
    rw = np.random.default_rng(20260925)
    shapes = [(2, 2, 0), (3, 4, 0), (5, 2, 4), (37, 53, 0), (64, 48, 7), (96, 128, 0), (130, 100, 0)]
    for k, (M, N, it) in enumerate(shapes):
        Y = O.gen_cfg2(M, N, seed=300 + k, block=8) if min(M, N) >= 8 else np.asfortranarray(rw.normal(0, 1, (M, N)))
        W1 = np.asfortranarray(rw.uniform(0.02, 0.6, (M - 1, N))); W2 = np.asfortranarray(rw.uniform(0.02, 0.6, (M, N - 1)))
        if k == 4: W1[::3] = 0.0                                   # some zero weights (no coupling across those edges)
        o, info = R.dr2l1w_tv(Y, W1, W2, maxit=it)
        out["drw_%d_Y" % k] = Y; out["drw_%d_W1" % k] = W1; out["drw_%d_W2" % k] = W2; out["drw_%d_it" % k] = np.int64(it)
        out["drw_%d_out" % k] = o; out["drw_%d_info" % k] = info
    out["drw_count"] = np.int64(len(shapes))
    path = os.path.join(ROOT, "tests", "golden", "golden_v1.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main()
