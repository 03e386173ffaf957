"""Per-kernel-class event timing of one DR2_TV solve under a given engine (profiling forces plain launches)."""
import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import proxtv_b200 as ptv
import synth_inputs as O
lib = ptv.require_device()
M = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
engs = sys.argv[2].split(",") if len(sys.argv) > 2 else ["auto"]
Yh = O.gen_cfg2(M, M, seed=0); Yd = torch.from_numpy(np.ascontiguousarray(Yh.T)).cuda(); out = torch.empty_like(Yd); info = np.zeros(3)
for eng in engs:
    ptv.set_engine(eng)
    def solve():
        lib.proxtv_DR2_TV_dev_f64(M, M, 1, 0, C.c_void_p(Yd.data_ptr()), 0.2, 0.2, C.c_void_p(out.data_ptr()), 0, C.c_void_p(info.ctypes.data), None)
    solve(); torch.cuda.synchronize()
    lib.proxtv_profile_reset(); lib.proxtv_profile_enable(1)
    solve(); torch.cuda.synchronize()
    lib.proxtv_profile_enable(0)
    kms = (C.c_double * 3)(); kl = (C.c_longlong * 3)(); ks = (C.c_longlong * 3)()
    lib.proxtv_profile_read(kms, kl, ks)
    print(eng, "class ms:", [round(kms[i], 2) for i in range(3)], "spans:", [ks[i] for i in range(3)], "avg us:", [round(1e3 * kms[i] / max(ks[i], 1), 1) for i in range(3)])
    if hasattr(lib, "proxtv_debug_phase_read"):               # debug build: make EXTRA=-DPTV_PHASE_TIMING
        ph = (C.c_ulonglong * 8)(); lib.proxtv_debug_phase_read(ph, 1)
        tot = sum(ph[k] for k in range(5))
        print("    contig-kernel phases (CTA-time share) stage/round0/rounds/cval/fill: " + " ".join(f"{100*ph[k]/tot:.1f}%" for k in range(5))
              + f"  mean CTA life {tot/ph[7]/1e3:.1f} us, rounds {ph[6]/ph[7]:.2f}")
