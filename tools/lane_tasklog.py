"""Per-task timeline of one lane-engine launch (config-2 passes): when each warp task starts / ends, on which SM, how many warps are
resident over time -- tells how much of a pass is a tail.  usage: lane_tasklog.py [variant] [layout: s|c] [op] [clen]"""
import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import proxtv_b200 as ptv
import synth_inputs as O
lib = ptv.require_device(); vp = C.c_void_p
variant = int(sys.argv[1]) if len(sys.argv) > 1 else 0
lay = sys.argv[2] if len(sys.argv) > 2 else "s"
op = int(sys.argv[3]) if len(sys.argv) > 3 else 0
clen = int(sys.argv[4]) if len(sys.argv) > 4 else 0
M = N = 4096
Y = O.gen_cfg2(M, N, seed=0)
x = torch.tensor(np.ascontiguousarray(Y.T), device="cuda"); out = torch.empty_like(x)
t = torch.tensor(np.random.default_rng(1).normal(0, 1, (N, M)), device="cuda"); xa = t * 0.5
CAP = 1 << 16
log = torch.zeros(CAP * 4, dtype=torch.int64, device="cuda")
lib.proxtv_lane_tuning(clen, 32, variant)
st = vp(torch.cuda.current_stream().cuda_stream)


out2 = torch.empty_like(x)


def run():
    if lay == "s":
        ok = lib.proxtv_lane_prox2_dev_f64(op, vp(x.data_ptr()), vp(xa.data_ptr()) if op else None, vp(t.data_ptr()) if op else None, vp(out.data_ptr()),
                                          vp(out2.data_ptr()), M, N, M, 0.2, st)
    else:
        ok = lib.proxtv_lane_prox_dev_f64(0, vp(x.data_ptr()), None, None, vp(out.data_ptr()), N, M, 1, 0.2, st)
    assert ok, lib.proxtv_last_error()


for _ in range(3):
    run()
torch.cuda.synchronize()
lib.proxtv_lane_tasklog(vp(log.data_ptr()), CAP)
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record(); run(); e1.record()
torch.cuda.synchronize()
lib.proxtv_lane_tasklog(None, 0)
L = log.cpu().numpy().reshape(-1, 4)
L = L[L[:, 0] > 0]
nt = len(L)
t0 = L[:, 0].min()
start = (L[:, 0] - t0) / 1e3; scan_end = (L[:, 1] - t0) / 1e3; end = (L[:, 2] - t0) / 1e3; sm = L[:, 3]
print("variant %d layout %s op %d: %d tasks, event time %.1f us, log span %.1f us" % (variant, lay, op, nt, e0.elapsed_time(e1) * 1e3, end.max()))
print("start  : min %.1f  median %.1f  p99 %.1f  max %.1f us" % (start.min(), np.median(start), np.percentile(start, 99), start.max()))
dur = scan_end - start
print("scan   : min %.1f  p10 %.1f  median %.1f  p90 %.1f  max %.1f us" % (dur.min(), np.percentile(dur, 10), np.median(dur), np.percentile(dur, 90), dur.max()))
print("scanend: min %.1f  p10 %.1f  median %.1f  p90 %.1f  max %.1f us" % (scan_end.min(), np.percentile(scan_end, 10), np.median(scan_end), np.percentile(scan_end, 90), scan_end.max()))
post = end - scan_end
print("post   : median %.2f  p99 %.1f  max %.1f us   (records, verification, repairs)" % (np.median(post), np.percentile(post, 99), post.max()))
per_sm = np.bincount(sm.astype(int), minlength=148)
print("tasks per SM: min %d max %d  histogram %s" % (per_sm.min(), per_sm.max(), dict(zip(*np.unique(per_sm, return_counts=True)))))
for k in np.unique(per_sm):
    sel = np.isin(sm, np.nonzero(per_sm == k)[0])
    print("   SMs with %2d tasks: scan median %.1f us, last scan end %.1f us" % (k, np.median(dur[sel]), scan_end[sel].max()))
# resident warps over time
T = end.max(); grid = np.linspace(0, T, 201)
act = [(np.sum((start <= g) & (end > g))) for g in grid]
print("resident warps at 10%% steps of the span: %s" % [int(act[i]) for i in range(0, 201, 20)])
area = np.sum(end - start)
print("mean residency %.1f%% of %d tasks x span (tail loss %.1f%%)" % (100 * area / (nt * T), nt, 100 - 100 * area / (nt * T)))
# does duration depend on chunk index / on data?
