"""CPU probe for the round-2 warm-start idea: how much of the segmentation of each DR pass survives from one iteration to the
next, at the granularity the kernel would re-scan (32-sample chunks)?  Runs the DR2_TV iteration with the oracle's 1D solver.
usage: python tools/warmstart_probe.py [M=1024] [lam=0.2]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle as O
M = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
lam = float(sys.argv[2]) if len(sys.argv) > 2 else 0.2
P = O.Port()
Y = np.ascontiguousarray(O.gen_cfg2(M, M, seed=0))
def prox_rows(A):      # prox of every row of a C-ordered array
    return np.stack([P.tv1_linearized(A[i], lam) for i in range(A.shape[0])])
def breaks(X):         # boolean (rows, n-1): jump between i and i+1
    return X[:, 1:] != X[:, :-1]
def chunk_dirty(b0, b1):
    d = b0 != b1; n = d.shape[1] + 1
    pad = np.zeros((d.shape[0], (-n) % 32 + 1), bool)
    d = np.concatenate([d, pad], axis=1).reshape(d.shape[0], -1, 32).any(axis=2)
    return d
t = np.full_like(Y, 2 * Y.mean()); prev = {}
print("iter  pass   breaks/sample  changed-breaks%  dirty-chunks%  mean-dirty-run  fibers-with-no-dirty-chunk%")
for it in range(int(sys.argv[3]) if len(sys.argv) > 3 else 12):
    xc = prox_rows(t.T.copy()).T                        # columns
    s = 2 * (t - xc) - t
    inp = Y - s
    xr = prox_rows(inp)                                 # rows
    tb = Y - (inp - xr); tb = 2 * tb - s; t = 0.5 * (t + tb)
    for name, X in (("cols", xc.T), ("rows", xr)):
        b = breaks(np.ascontiguousarray(X))
        if name in prev:
            d = chunk_dirty(prev[name], b)
            runs = []
            for row in d:
                k = 0
                for v in row:
                    if v: k += 1
                    elif k: runs.append(k); k = 0
                if k: runs.append(k)
            print(f"{it:3d}   {name}   {b.mean():.3f}          {100*(prev[name]!=b).sum()/max(b.sum(),1):6.2f}          {100*d.mean():6.2f}        {np.mean(runs) if runs else 0:5.2f}          {100*(~d.any(axis=1)).mean():5.1f}")
        prev[name] = b
