import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import proxtv_b200 as ptv
from oracle import oracle as O
y = O.gen_cfg1(1_000_000, seed=0)
ptv.tv1_1d(y, 0.5)
t0 = time.perf_counter(); x = ptv.tv1_1d(y, 0.5); dt = time.perf_counter() - t0
P = O.Port(); t0 = time.perf_counter(); r = P.tv1_hybrid(y, 0.5); dc = time.perf_counter() - t0
print("cfg1 tv1_1d n=1e6 f64: GPU end-to-end (host in/out) %.2f ms = %.0f Msamples/s ; oracle port on 1 CPU core %.1f ms ; bit-exact vs linearized: %s"
      % (dt * 1e3, 1.0 / dt, dc * 1e3, bool(np.array_equal(x, P.tv1_linearized(y, 0.5)))))
