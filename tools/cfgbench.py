"""Times the other BASELINE configs on one GPU (device-resident, CUDA events) and checks a sample against the oracle.
usage: python tools/cfgbench.py [cfg3] [cfg4] [cfg5] [small]"""
import ctypes as C, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import proxtv_b200 as ptv
from oracle import oracle as O

def ev_time(fn, reps=5, warm=4):      # the GPU idles (clocks drop) while the host generates data: warm up properly
    for _ in range(warm): fn()
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / reps

which = sys.argv[1:] or ["cfg3", "cfg4", "cfg5"]
small = "small" in which
P = O.Port(); lib = ptv.require_device()
if "cfg3" in which:
    B, L = (4096, 4096) if small else (65536, 4096)
    rng = np.random.default_rng(0)
    # generate on the GPU-side in slabs to keep host memory modest
    X = torch.empty((B, L), dtype=torch.float64, device="cuda"); W = torch.empty((B, L - 1), dtype=torch.float64, device="cuda")
    for b0 in range(0, B, 4096):
        xs, ws = O.gen_cfg3(min(4096, B - b0), L, seed=b0)
        X[b0:b0 + xs.shape[0]] = torch.from_numpy(xs).cuda(); W[b0:b0 + xs.shape[0]] = torch.from_numpy(ws).cuda()
    out = None
    def run():
        global out
        out = ptv.tv1w_1d_batched(X, W)
    ms = ev_time(run, reps=10, warm=5)          # the GPU idles (clocks drop) while the host generates the data: warm up properly
    xs, ws = O.gen_cfg3(min(4096, B), L, seed=0)
    ok = all(np.array_equal(out[b].cpu().numpy(), P.tv1_weighted(xs[b], ws[b])) for b in (0, 1, 4095))
    U = torch.cumsum(X - out, dim=1); kkt = bool((U[:, :-1].abs() <= W + 1e-9).all().item()) and float(U[:, -1].abs().max()) < 1e-7
    print(f"cfg3 tv1w_1d batch {B}x{L} f64: {ms:.2f} ms  {B*L/ms/1e3:.1f} Msamples/s  {24*B*L/ms/1e6:.1f} GB/s(24B/sample)  bit-exact rows: {ok}  KKT all rows: {kkt}", flush=True)
    del X, W, out, U; torch.cuda.empty_cache()
if "cfg4" in which:
    shp = (128, 128, 64) if small else (512, 512, 256)
    V = O.gen_cfg4(shp, seed=0)
    Vf = np.asfortranarray(V.astype(np.float32))
    t0 = time.perf_counter(); g = ptv.tvgen(Vf, [0.2, 0.2, 0.2], [1, 2, 3], [1, 1, 1]); t1 = time.perf_counter() - t0
    t0 = time.perf_counter(); g = ptv.tvgen(Vf, [0.2, 0.2, 0.2], [1, 2, 3], [1, 1, 1]); t1 = time.perf_counter() - t0
    info = ptv.tvgen.last_info.copy()
    print(f"cfg4 tvgen PD_TV {shp} f32 (host in/out, e2e wall): {t1*1e3:.1f} ms  {V.size/t1/1e6:.1f} Mvox/s  iters={info[0]} stop={info[1]:.3e} rc={info[2]}", flush=True)
    # device-resident timing through the C ABI
    Vd = torch.from_numpy(np.ascontiguousarray(Vf.transpose(2, 1, 0))).cuda(); outd = torch.empty_like(Vd)   # memory order == F-order
    ns = np.array(shp, dtype=np.int32); dims = np.array([1.0, 2.0, 3.0]); inf = np.zeros(3)
    def runpd():
        lam = np.array([0.2, 0.2, 0.2])
        lib.proxtv_PD_TV_dev_f32(C.c_void_p(Vd.data_ptr()), C.c_void_p(lam.ctypes.data), C.c_void_p(dims.ctypes.data), C.c_void_p(outd.data_ptr()),
                                 C.c_void_p(inf.ctypes.data), C.c_void_p(ns.ctypes.data), 3, 3, 0, None)
    ms = ev_time(runpd, reps=2)
    print(f"cfg4 PD_TV {shp} f32 device-resident: {ms:.1f} ms  {V.size/ms/1e3:.1f} Mvox/s iters={inf[0]}", flush=True)
    if small:
        want, winfo = P.pd_tv(Vf.astype(np.float64), [0.2, 0.2, 0.2], [1, 2, 3])
        print("   vs f64 oracle on f32-rounded input: rel err %.2e  iters oracle %d" % (np.abs(g - want).max() / np.abs(want).max(), winfo[0]))
if "w2d" in which:
    M = 1024 if small else 4096
    rng = np.random.default_rng(0)
    Y = O.gen_cfg2(M, M, seed=0)
    Yd = torch.from_numpy(np.ascontiguousarray(Y.T)).cuda()                     # column-major image
    W1 = torch.from_numpy(rng.uniform(0.1, 0.3, (M, M - 1))).cuda()              # (M-1) x N column-major == (N, M-1) row-major
    W2 = torch.from_numpy(rng.uniform(0.1, 0.3, (M - 1, M))).cuda()              # M x (N-1) column-major == (N-1, M) row-major
    outd = torch.empty_like(Yd); inf = np.zeros(3)
    def runw():
        lib.proxtv_DR2L1W_TV_dev_f64(M, M, C.c_void_p(Yd.data_ptr()), C.c_void_p(W1.data_ptr()), C.c_void_p(W2.data_ptr()),
                                     C.c_void_p(outd.data_ptr()), 0, C.c_void_p(inf.ctypes.data), None)
    ms = ev_time(runw, reps=3)
    line = f"tv1w_2d DR2L1W_TV {M}x{M} f64 device-resident: {ms:.2f} ms  {M*M/ms/1e3:.1f} Mpix/s"
    if small:
        want, _ = P.dr2l1w_tv(Y, W1.cpu().numpy().T, W2.cpu().numpy().T)
        line += "  rel err vs oracle %.2e" % (np.abs(outd.cpu().numpy().T - want).max() / np.abs(want).max())
    print(line, flush=True)
if "cfg5" in which:
    Bn, H = (16, 2048) if not small else (8, 512)
    imgs = torch.stack([torch.from_numpy(np.ascontiguousarray(O.gen_cfg2(H, H, seed=s).astype(np.float32))) for s in range(Bn)]).cuda()
    out = None
    def run5():
        global out
        out = ptv.tv1_2d_batched(imgs, 0.2)
    ms = ev_time(run5, reps=2)
    want = P.dr2_tv(imgs[1].cpu().numpy().astype(np.float64), 0.2)[0] if H <= 512 else None
    err = "" if want is None else " rel err vs f64 oracle %.2e" % (np.abs(out[1].cpu().numpy() - want).max() / np.abs(want).max())
    print(f"cfg5 tv1_2d batch {Bn}x{H}x{H} f32: {ms:.1f} ms  {Bn*H*H/ms/1e3:.1f} Mpix/s per GPU{err}", flush=True)
