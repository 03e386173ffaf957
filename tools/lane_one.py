"""Launch the lane engine a few times on the config-2 strided pass (for ncu).  usage: lane_one.py [variant] [clen] [op]"""
import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import proxtv_b200 as ptv
import synth_inputs as O
lib = ptv.require_device(); vp = C.c_void_p
lib.proxtv_lane_prox_dev_f64.argtypes = [C.c_int, vp, vp, vp, vp, C.c_longlong, C.c_int, C.c_longlong, C.c_double, vp]
lib.proxtv_lane_tuning.argtypes = [C.c_int, C.c_int, C.c_int]
variant = int(sys.argv[1]) if len(sys.argv) > 1 else 0
clen = int(sys.argv[2]) if len(sys.argv) > 2 else 256
op = int(sys.argv[3]) if len(sys.argv) > 3 else 0
M = N = 4096
Y = O.gen_cfg2(M, N, seed=0)
x = torch.tensor(np.ascontiguousarray(Y.T), device="cuda"); out = torch.empty_like(x)
t = torch.tensor(np.random.default_rng(1).normal(0, 1, (N, M)), device="cuda"); xa = t * 0.5
lib.proxtv_lane_tuning(clen, 32, variant)
st = vp(torch.cuda.current_stream().cuda_stream)
for _ in range(3):
    lib.proxtv_lane_prox_dev_f64(op, vp(x.data_ptr()), vp(xa.data_ptr()) if op else None, vp(t.data_ptr()) if op else None, vp(out.data_ptr()), M, N, M, 0.2, st)
torch.cuda.synchronize()
print("done")
