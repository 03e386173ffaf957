"""Times the default (pipelined, graph-replayed) DR2_TV 4096x4096 f64 solve (device-resident, CUDA events) and prints a checksum."""
import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import proxtv_b200 as ptv
import synth_inputs as O
lib = ptv.require_device(); M = 4096
Yd = torch.from_numpy(np.ascontiguousarray(O.gen_cfg2(M, M, seed=0).T)).cuda(); out = torch.empty_like(Yd); info = np.zeros(3)
f = lambda: lib.proxtv_DR2_TV_dev_f64(M, M, 1, 0, C.c_void_p(Yd.data_ptr()), 0.2, 0.2, C.c_void_p(out.data_ptr()), 0, C.c_void_p(info.ctypes.data), None)
for _ in range(3): f()
torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True); e0.record()
for _ in range(5): f()
e1.record(); torch.cuda.synchronize()
print("DR2_TV %dx%d f64: %.2f ms  checksum %.12e" % (M, M, e0.elapsed_time(e1) / 5, float(out.sum().item())))
