"""Times cfg3-like (weighted batch) and cfg5-like (f32 image batch) workloads on a given library file (raw ctypes, so that
libraries of older revisions work).  usage: python tools/bisect_bench.py lib1.so lib2.so ..."""
import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import synth_inputs as O

def ev(fn, reps=5):
    for _ in range(2): fn()
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / reps

B, L = 16384, 4096
xs, ws = O.gen_cfg3(4096, L, seed=0)
X = torch.from_numpy(np.tile(xs, (4, 1))).cuda(); W = torch.from_numpy(np.tile(ws, (4, 1))).cuda(); out = torch.empty_like(X)
H = 2048; Bn = 8
imgs = torch.stack([torch.from_numpy(np.ascontiguousarray(O.gen_cfg2(H, H, seed=s).astype(np.float32))) for s in range(Bn)]).cuda()
iout = torch.empty_like(imgs); info = np.zeros(3)
Y64 = torch.from_numpy(np.ascontiguousarray(O.gen_cfg2(4096, 4096, seed=0).T)).cuda(); o64 = torch.empty_like(Y64)
vp = C.c_void_p
for path in sys.argv[1:]:
    lib = C.CDLL(os.path.abspath(path))
    f = lib.proxtv_prox_fibers_dev_f64; f.argtypes = [vp, vp, C.c_longlong, C.c_int, C.c_longlong, C.c_double, vp, vp]
    t3 = ev(lambda: f(X.data_ptr(), out.data_ptr(), B, L, 1, 0.0, W.data_ptr(), None))
    tu = ev(lambda: f(X.data_ptr(), out.data_ptr(), B, L, 1, 0.5, None, None))
    g = lib.proxtv_DR2_TV_dev_f32; g.argtypes = [C.c_size_t, C.c_size_t, C.c_int, C.c_int, vp, C.c_float, C.c_float, vp, C.c_int, vp, vp]
    t5 = ev(lambda: g(H, H, Bn, 1, imgs.data_ptr(), 0.2, 0.2, iout.data_ptr(), 0, info.ctypes.data, None), reps=3)
    h = lib.proxtv_DR2_TV_dev_f64; h.argtypes = [C.c_size_t, C.c_size_t, C.c_int, C.c_int, vp, C.c_double, C.c_double, vp, C.c_int, vp, vp]
    t2 = ev(lambda: h(4096, 4096, 1, 0, Y64.data_ptr(), 0.2, 0.2, o64.data_ptr(), 0, info.ctypes.data, None), reps=3)
    print(f"{os.path.basename(path):18s} weighted {B}x{L} f64: {t3:7.2f} ms | unweighted same: {tu:6.2f} ms | DR2 f32 {Bn}x{H}^2 row-major: {t5:6.1f} ms | DR2 f64 4096^2: {t2:6.2f} ms", flush=True)
