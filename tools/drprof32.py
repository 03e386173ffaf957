"""Per-kernel-class event timing of a float32 batched, row-major DR2_TV solve (config 5 shape)."""
import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import proxtv_b200 as ptv
import synth_inputs as O
lib = ptv.require_device()
H = int(sys.argv[1]) if len(sys.argv) > 1 else 2048; Bn = int(sys.argv[2]) if len(sys.argv) > 2 else 8
imgs = torch.stack([torch.from_numpy(np.ascontiguousarray(O.gen_cfg2(H, H, seed=s).astype(np.float32))) for s in range(Bn)]).cuda()
out = torch.empty_like(imgs); info = np.zeros(3)
def solve(): lib.proxtv_DR2_TV_dev_f32(H, H, Bn, 1, C.c_void_p(imgs.data_ptr()), C.c_float(0.2), C.c_float(0.2), C.c_void_p(out.data_ptr()), 0, C.c_void_p(info.ctypes.data), None)
solve(); torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True); e0.record(); solve(); e1.record(); torch.cuda.synchronize()
if hasattr(lib, "proxtv_debug_phase_read"):
    ph = (C.c_ulonglong * 8)(); lib.proxtv_debug_phase_read(ph, 1)
lib.proxtv_profile_reset(); lib.proxtv_profile_enable(1); solve(); torch.cuda.synchronize(); lib.proxtv_profile_enable(0)
kms = (C.c_double * 3)(); kl = (C.c_longlong * 3)(); ks = (C.c_longlong * 3)(); lib.proxtv_profile_read(kms, kl, ks)
print("DR2 f32 %dx%d^2 row-major: %.1f ms; class ms:" % (Bn, H, e0.elapsed_time(e1)), [round(kms[i], 2) for i in range(3)], "spans:", [ks[i] for i in range(3)],
      "avg us:", [round(1e3 * kms[i] / max(ks[i], 1), 1) for i in range(3)])
if hasattr(lib, "proxtv_debug_phase_read"):               # debug build: phases of ALL contig-kernel launches of the profiled solve
    ph = (C.c_ulonglong * 8)(); lib.proxtv_debug_phase_read(ph, 1)
    tot = sum(ph[k] for k in range(5))
    print("    contig-kernel phases (CTA-time share) stage/round0/rounds/cval/fill: " + " ".join(f"{100*ph[k]/tot:.1f}%" for k in range(5))
          + f"  mean CTA life {tot/ph[7]/1e3:.1f} us, rounds {ph[6]/ph[7]:.2f}")
