"""Kernel micro-benchmark: times proxtv_prox_fibers_dev_* on DR-like data (CUDA events on torch's stream).
usage: python tools/kbench.py [n=4096] [nf=4096] [dtype=f64] [engine=auto]"""
import ctypes as C, sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import proxtv_b200 as ptv
import synth_inputs as O

def timeit(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    nf = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
    dt = sys.argv[3] if len(sys.argv) > 3 else "f64"
    engs = sys.argv[4].split(",") if len(sys.argv) > 4 else ["auto"]
    lib = ptv.require_device()
    tdt = torch.float64 if dt == "f64" else torch.float32
    Y = O.gen_cfg2(nf, n, seed=0)                      # rows = fibers of length n, noise-like
    datasets = {"noisy(cfg2 rows)": np.ascontiguousarray(Y), "constant": np.full((nf, n), 0.37),
                "smooth(lam=1.0)": np.ascontiguousarray(Y)}
    lams = {"noisy(cfg2 rows)": 0.2, "constant": 0.2, "smooth(lam=1.0)": 1.0}
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    fn_c = lib.proxtv_prox_fibers_dev_f64 if dt == "f64" else lib.proxtv_prox_fibers_dev_f32
    for eng in engs:
        ptv.set_engine(eng)
        for name, arr in datasets.items():
            x = torch.tensor(arr, dtype=tdt, device="cuda"); out = torch.empty_like(x)
            lam = lams[name]
            for inc, label in ((1, "contiguous"), (nf, "strided")):
                # strided: treat x as a column-major (nf x n) matrix whose ROW fibers (length n, stride nf) we process
                if inc == 1:
                    f = lambda: fn_c(C.c_void_p(x.data_ptr()), C.c_void_p(out.data_ptr()), nf, n, 1, lam, None, st)
                else:
                    xt = x.t().contiguous()              # (n, nf) row-major == (nf x n) column-major
                    ot = torch.empty_like(xt)
                    f = lambda: fn_c(C.c_void_p(xt.data_ptr()), C.c_void_p(ot.data_ptr()), nf, n, nf, lam, None, st)
                ms = timeit(f, reps=5 if eng == "seq" else 20)
                if hasattr(lib, "proxtv_debug_phase_read"):       # debug build: make EXTRA=-DPTV_PHASE_TIMING
                    ph = (C.c_ulonglong * 8)(); lib.proxtv_debug_phase_read(ph, 1)
                    if ph[7]:
                        tot = sum(ph[k] for k in range(5))
                        print("    phases (CTA-time share) stage/round0/rounds/cval/fill: " + " ".join(f"{100*ph[k]/tot:.1f}%" for k in range(5))
                              + f"  mean CTA life {tot/ph[7]/1e3:.1f} us, rounds {ph[6]/ph[7]:.2f}")
                gbs = 2 * nf * n * x.element_size() / ms / 1e6
                print(f"{eng:8s} {name:18s} {label:10s} n={n} nf={nf} {dt}: {ms*1e3:9.1f} us  {gbs:8.1f} GB/s (1R+1W)", flush=True)

main()
