"""Per-kernel-class event timing of PD_TV (config 4 shape by default)."""
import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import proxtv_b200 as ptv
import synth_inputs as O
lib = ptv.require_device()
shp = tuple(int(v) for v in sys.argv[1].split("x")) if len(sys.argv) > 1 else (512, 512, 256)
V = O.gen_cfg4(shp, seed=0).astype(np.float32)
Vd = torch.from_numpy(np.ascontiguousarray(np.asfortranarray(V).transpose(2, 1, 0))).cuda(); outd = torch.empty_like(Vd)
ns = np.array(shp, dtype=np.int32); dims = np.array([1.0, 2.0, 3.0]); inf = np.zeros(3)
def run():
    lam = np.array([0.2, 0.2, 0.2])
    lib.proxtv_PD_TV_dev_f32(C.c_void_p(Vd.data_ptr()), C.c_void_p(lam.ctypes.data), C.c_void_p(dims.ctypes.data), C.c_void_p(outd.data_ptr()),
                             C.c_void_p(inf.ctypes.data), C.c_void_p(ns.ctypes.data), 3, 3, 0, None)
run(); torch.cuda.synchronize()
lib.proxtv_profile_reset(); lib.proxtv_profile_enable(1)
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True); e0.record(); run(); e1.record(); torch.cuda.synchronize()
lib.proxtv_profile_enable(0)
kms = (C.c_double * 3)(); kl = (C.c_longlong * 3)(); ks = (C.c_longlong * 3)()
lib.proxtv_profile_read(kms, kl, ks)
print("PD_TV", shp, "f32 total %.1f ms iters %d" % (e0.elapsed_time(e1), inf[0]), "class ms:", [round(kms[i], 1) for i in range(3)], "spans:", [ks[i] for i in range(3)],
      "avg us:", [round(1e3 * kms[i] / max(ks[i], 1), 1) for i in range(3)])
if hasattr(lib, "proxtv_debug_phase_read"):               # debug build: make EXTRA=-DPTV_PHASE_TIMING
    ph = (C.c_ulonglong * 8)(); lib.proxtv_debug_phase_read(ph, 1)
    tot = sum(ph[k] for k in range(5))
    print("    contig-kernel phases (CTA-time share) stage/round0/rounds/cval/fill: " + " ".join(f"{100*ph[k]/tot:.1f}%" for k in range(5))
          + f"  mean CTA life {tot/ph[7]/1e3:.1f} us, rounds {ph[6]/ph[7]:.2f}")
