"""DR2_TV 4096 x 4096 f64 for growing lambda (longer segments -> more lanes outgrow the window and go through the repair path):
lane engine vs chunked engine, device-resident solve time and repaired fibers per solve."""
import ctypes as C, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import proxtv_b200 as ptv
import synth_inputs as S
lib = ptv.require_device(); lib.proxtv_lane_stats.restype = C.c_ulonglong
Y = S.gen_cfg2(4096, 4096, seed=0)
y = torch.tensor(np.ascontiguousarray(Y.T), device="cuda").t()
for lam in (0.2, 1.0, 5.0, 25.0):
    row = []
    for eng in ("lane", "chunked"):
        ptv.set_engine(eng)
        out = ptv.tv1_2d(y, lam); torch.cuda.synchronize(); lib.proxtv_lane_stats(1)
        t0 = time.perf_counter(); out2 = ptv.tv1_2d(y, lam); torch.cuda.synchronize(); dt = time.perf_counter() - t0
        row.append((eng, dt * 1e3, int(lib.proxtv_lane_stats(1)), out2))
    err = ((row[0][3] - row[1][3]).abs().max() / row[1][3].abs().max()).item()
    print("lambda %5.1f: lane %.2f ms (%d repaired fibers)   chunked %.2f ms   rel diff %.1e" % (lam, row[0][1], row[0][2], row[1][1], err), flush=True)
