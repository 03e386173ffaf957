"""Run one DR2_TV solve per engine setting and print the lane engine's repair counter (how many fibers needed the sequential
repair path) -- a health check of the speculation on a given input.  usage: lane_stats.py [size] [lam]"""
import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import proxtv_b200 as ptv
import synth_inputs as O
lib = ptv.require_device(); vp = C.c_void_p
lib.proxtv_lane_stats.argtypes = [C.c_int]; lib.proxtv_lane_stats.restype = C.c_ulonglong
M = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
lam = float(sys.argv[2]) if len(sys.argv) > 2 else 0.2
Y = O.gen_cfg2(M, M, seed=0)
Yd = torch.from_numpy(np.ascontiguousarray(Y.T)).cuda(); out = torch.empty_like(Yd); info = np.zeros(3)
st = vp(torch.cuda.current_stream().cuda_stream)
for it in (1, 2, 5, 35):
    lib.proxtv_lane_stats(1)
    lib.proxtv_DR2_TV_dev_f64(M, M, 1, 0, vp(Yd.data_ptr()), lam, lam, vp(out.data_ptr()), it, vp(info.ctypes.data), st)
    torch.cuda.synchronize()
    print("maxit", it, "repairs", lib.proxtv_lane_stats(1))
