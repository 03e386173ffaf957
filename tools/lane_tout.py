"""Strided plain pass writing its result in place (op 0) vs transposed, fiber-major (op 6).  Measured: +3.7 us per 134 MB array with the
warp-cooperative in-place exchange, +10 us with naive 16-byte scattered stores (the first version of this experiment)."""
import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import proxtv_b200 as ptv
import synth_inputs as O
lib = ptv.require_device(); vp = C.c_void_p
M = N = 4096
Y = O.gen_cfg2(M, N, seed=0)
x = torch.tensor(np.ascontiguousarray(Y.T), device="cuda")      # (N, M) row-major == column-major M x N; strided fibers = rows of the image
st = vp(torch.cuda.current_stream().cuda_stream)


def run(variant, out):
    lib.proxtv_lane_tuning(0, 32, variant & 0xffff)
    op = 6 if variant & 0x10000 else 0            # 6 = plain pass, result transposed (in-place warp-cooperative exchange, store8_transposed)
    assert lib.proxtv_lane_prox2_dev_f64(op, vp(x.data_ptr()), None, None, vp(out.data_ptr()), None, M, N, M, 0.2, st), lib.proxtv_last_error()


def timeit(variant, out, reps=20):
    for _ in range(3): run(variant, out)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): run(variant, out)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


a = torch.empty_like(x); b = torch.zeros_like(x)
for v in (0, 1):
    ta = timeit(v, a); tb = timeit(v | 0x10000, b)
    # a[j, i] = result for fiber i (image row i), sample j ; transposed output: b.view(M, N)[i, j]
    same = torch.equal(a.t().contiguous(), b.view(M, N))
    print("variant %d: in place %.1f us   transposed %.1f us   identical %s" % (v, ta, tb, same))
