"""Micro-benchmark + correctness check of the lane-per-fiber engine (kernels_lane.cu) against the bit-faithful chunked engine.
usage: python tools/lanebench.py [quick]"""
import ctypes as C, json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import proxtv_b200 as ptv
import synth_inputs as O

lib = ptv.require_device()
vp = C.c_void_p
lib.proxtv_lane_prox_dev_f64.argtypes = [C.c_int, vp, vp, vp, vp, C.c_longlong, C.c_int, C.c_longlong, C.c_double, vp]
lib.proxtv_lane_prox_dev_f32.argtypes = [C.c_int, vp, vp, vp, vp, C.c_longlong, C.c_int, C.c_longlong, C.c_float, vp]
lib.proxtv_lane_prox2_dev_f64.argtypes = [C.c_int, vp, vp, vp, vp, vp, C.c_longlong, C.c_int, C.c_longlong, C.c_double, vp]
lib.proxtv_lane_tuning.argtypes = [C.c_int, C.c_int, C.c_int]
lib.proxtv_lane_stats.argtypes = [C.c_int]; lib.proxtv_lane_stats.restype = C.c_ulonglong
OUT = {}


def timeit(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3     # us


def stream():
    return vp(torch.cuda.current_stream().cuda_stream)


def old_prox(x, nf, ln, inc, lam):
    ptv.set_engine("chunked")                       # the bit-faithful chunked engine (engine 'auto' would dispatch to the lane engine)
    out = torch.empty_like(x)
    fn = lib.proxtv_prox_fibers_dev_f64 if x.dtype == torch.float64 else lib.proxtv_prox_fibers_dev_f32
    assert fn(vp(x.data_ptr()), vp(out.data_ptr()), nf, ln, inc, lam, None, stream()) == 1
    return out


def lane(op, A, B, Cc, nf, ln, inc, lam, out=None):
    out = torch.empty_like(A) if out is None else out
    fn = lib.proxtv_lane_prox_dev_f64 if A.dtype == torch.float64 else lib.proxtv_lane_prox_dev_f32
    p = lambda t: vp(t.data_ptr()) if t is not None else None
    rc = fn(op, p(A), p(B), p(Cc), p(out), nf, ln, inc, lam, stream())
    if rc != 1: return None          # shape not supported by the lane engine (the library falls back to the chunked engine)
    return out


def check_small():
    rng = np.random.default_rng(3)
    ok = True
    for (M, N, lam, clen, halo) in [(200, 300, 0.2, 0, 32), (64, 1000, 0.2, 128, 32), (96, 257, 1.0, 64, 16), (32, 64, 0.05, 0, 8),
                                    (130, 515, 0.2, 256, 32), (256, 4096, 0.2, 256, 32), (40, 700, 5.0, 128, 32)]:
        Y = O.gen_cfg2(M, N, seed=int(rng.integers(100)), block=16)          # F-order M x N
        x = torch.tensor(np.ascontiguousarray(Y.T), device="cuda")           # (N, M) row-major == column-major M x N
        want = old_prox(x, M, N, M, lam)
        for variant in (0, 1):
            lib.proxtv_lane_tuning(clen, halo, variant)
            got = lane(0, x, None, None, M, N, M, lam)
            err = (got - want).abs().max().item()
            rep = lib.proxtv_lane_stats(1)
            print(f"small M={M} N={N} lam={lam} clen={clen} halo={halo} v={variant}: max|diff| {err:.2e} repairs {rep}", flush=True)
            ok &= err < 1e-9
        # fused Douglas-Rachford forms
        t = torch.tensor(rng.normal(0, 1, (N, M)), device="cuda"); xa = torch.tensor(rng.normal(0, 1, (N, M)), device="cuda")
        for op in (1, 2):
            d = t - xa
            inn = x - (2 * d - t) if op == 1 else x - d
            px = old_prox(inn.contiguous(), M, N, M, lam)
            want2 = d + px if op == 1 else px
            lib.proxtv_lane_tuning(clen, halo, 0)
            got2 = lane(op, x, xa, t, M, N, M, lam)
            err = (got2 - want2).abs().max().item()
            print(f"   op={op}: max|diff| {err:.2e} repairs {lib.proxtv_lane_stats(1)}", flush=True)
            ok &= err < 1e-9
        # drain-side forms with transposed results (the lane-t schedule): fibers = rows of the M x N column-major image x
        if N % 2 == 0:
            px = old_prox(x, M, N, M, lam)                          # prox of every row, (N, M) layout
            lib.proxtv_lane_tuning(clen, halo, 0)
            T_ = lambda a: a.t().contiguous()                       # fiber-major = the transposed image
            u = torch.full_like(x, float("nan")).view(M, N); d2 = torch.full_like(x, float("nan")).view(M, N)
            assert lib.proxtv_lane_prox2_dev_f64(3, vp(x.data_ptr()), vp(xa.data_ptr()), vp(x.data_ptr()), vp(u.data_ptr()), vp(d2.data_ptr()), M, N, M, lam, stream())
            dd = x - px
            e1 = (d2 - T_(dd)).abs().max().item(); e2 = (u - T_(xa - (2 * dd - x))).abs().max().item()
            g4 = torch.full_like(x, float("nan")).view(M, N)
            assert lib.proxtv_lane_prox2_dev_f64(4, vp(x.data_ptr()), vp(xa.data_ptr()), vp(x.data_ptr()), vp(g4.data_ptr()), None, M, N, M, lam, stream())
            e3 = (g4 - T_(xa - (x - px))).abs().max().item()
            g5 = torch.full_like(x, float("nan")).view(M, N)
            assert lib.proxtv_lane_prox2_dev_f64(5, vp(x.data_ptr()), vp(xa.data_ptr()), None, vp(g5.data_ptr()), None, M, N, M, lam, stream())
            e4 = (g5 - T_(xa + px)).abs().max().item()
            g6 = torch.full_like(x, float("nan")).view(M, N)
            assert lib.proxtv_lane_prox2_dev_f64(6, vp(x.data_ptr()), None, None, vp(g6.data_ptr()), None, M, N, M, lam, stream())
            e5 = (g6 - T_(px)).abs().max().item()
            print(f"   ops 3..6 (transposed results): d {e1:.1e} u {e2:.1e} final-u {e3:.1e} t' {e4:.1e} plain {e5:.1e} repairs {lib.proxtv_lane_stats(1)}", flush=True)
            ok &= max(e1, e2, e3, e4, e5) < 1e-9
    # contiguous fibers (CONTIG layout: TMA box + in-place transpose in, swizzled staging + TMA store out)
    for (nf, ln, lam, clen) in [(64, 256, 0.2, 0), (100, 300, 0.2, 128), (33, 1000, 1.0, 256), (256, 4096, 0.2, 0), (7, 64, 0.1, 0), (40, 130, 0.2, 64)]:
        x = torch.tensor(np.ascontiguousarray(O.gen_cfg2(nf, ln, seed=5, block=16)), device="cuda")     # (nf, ln) row-major: fibers contiguous
        want = old_prox(x, nf, ln, 1, lam)
        for variant in (0, 1):
            lib.proxtv_lane_tuning(clen, 32, variant)
            got = lane(0, x, None, None, nf, ln, 1, lam)
            err = (got - want).abs().max().item()
            print(f"contig nf={nf} len={ln} lam={lam} clen={clen} v={variant}: max|diff| {err:.2e} repairs {lib.proxtv_lane_stats(1)}", flush=True)
            ok &= err < 1e-9
        x32 = x.float(); want32 = old_prox(x32, nf, ln, 1, lam)
        lib.proxtv_lane_tuning(clen, 32, 0)
        got32 = lane(0, x32, None, None, nf, ln, 1, lam)
        if got32 is None: print("contig f32: unsupported shape (pitch)"); continue
        err = (got32 - want32).abs().max().item()
        print(f"contig f32 nf={nf} len={ln}: max|diff| {err:.2e}", flush=True)
        ok &= err < 2e-4
    # adversarial: constant / smooth data, large lambda (long segments -> retire + repair path)
    for name, arr, lam in [("const", np.full((300, 64), 0.37), 0.2), ("smooth", np.sin(np.linspace(0, 20, 700))[:, None] * np.ones((1, 96)), 0.3),
                           ("biglam", rng.normal(0, 1, (500, 128)), 50.0)]:
        x = torch.tensor(np.ascontiguousarray(arr), device="cuda"); N, M = arr.shape
        want = old_prox(x, M, N, M, lam)
        for clen in (0, 128):
            lib.proxtv_lane_tuning(clen, 32, 0)
            got = lane(0, x, None, None, M, N, M, lam)
            err = (got - want).abs().max().item()
            print(f"adversarial {name} clen={clen}: max|diff| {err:.2e} repairs {lib.proxtv_lane_stats(1)}", flush=True)
            ok &= err < 1e-9
    OUT["small_ok"] = bool(ok)
    return ok


def bench_big(quick):
    M = N = 4096; lam = 0.2
    Y = O.gen_cfg2(M, N, seed=0)
    x = torch.tensor(np.ascontiguousarray(Y.T), device="cuda")
    ptv.set_engine("auto")
    want = old_prox(x, M, N, M, lam)
    t_old = timeit(lambda: old_prox(x, M, N, M, lam), 10)
    print(f"old engine strided pass (gather+scan+scatter): {t_old:.1f} us", flush=True)
    OUT["old_strided_us"] = t_old
    out = torch.empty_like(x)
    res = []
    configs = [(0, 32)]
    variants = [0, 1]        # 0: 4-warp CTAs (default), 1: single-warp CTAs
    if quick: configs = configs[:2]; variants = [0, 1]
    for variant in variants:
        for clen, halo in configs:
            lib.proxtv_lane_tuning(clen, halo, variant)
            got = lane(0, x, None, None, M, N, M, lam, out)
            err = (got - want).abs().max().item()
            rep = lib.proxtv_lane_stats(1)
            us = timeit(lambda: lane(0, x, None, None, M, N, M, lam, out))
            print(f"lane v={variant} clen={clen} halo={halo}: {us:8.1f} us  {2*M*N*8/us/1e3:7.1f} GB/s  max|diff| {err:.1e} repairs {rep}", flush=True)
            res.append(dict(variant=variant, clen=clen, halo=halo, us=us, err=err, repairs=int(rep)))
    OUT["plain_f64_4096"] = res
    # fused DR second half
    t = torch.tensor(np.random.default_rng(1).normal(0, 1, (N, M)), device="cuda"); xa = t * 0.5
    for variant in (0, 1):
        for clen in (0,):
            lib.proxtv_lane_tuning(clen, 32, variant)
            us = timeit(lambda: lane(1, x, xa, t, M, N, M, lam, out))
            print(f"lane DR_B v={variant} clen={clen}: {us:8.1f} us  ({4*M*N*8/us/1e3:7.1f} GB/s over 3R+1W)", flush=True)
            res.append(dict(op=1, variant=variant, clen=clen, us=us))
    o2 = torch.empty_like(x)
    for op, name, nb in ((3, "DRA (2R+2W, both results transposed)", 4), (4, "DRA final (2R+1W)", 3), (5, "DRB (2R+1W)", 3), (6, "plain, transposed result", 2)):
        lib.proxtv_lane_tuning(0, 32, 0)
        f = lambda: lib.proxtv_lane_prox2_dev_f64(op, vp(x.data_ptr()), vp(xa.data_ptr()), vp(x.data_ptr()), vp(out.data_ptr()), vp(o2.data_ptr()), M, N, M, lam, stream())
        us = timeit(f)
        print(f"lane {name}: {us:8.1f} us  ({nb*M*N*8/us/1e3:7.1f} GB/s)", flush=True)
        res.append(dict(op=op, us=us))
    # contiguous pass (columns of the image)
    wantc = old_prox(x, N, M, 1, lam)
    t_oldc = timeit(lambda: old_prox(x, N, M, 1, lam), 10)
    print(f"old engine contiguous pass: {t_oldc:.1f} us", flush=True)
    OUT["old_contig_us"] = t_oldc
    for variant in (0, 1):
        lib.proxtv_lane_tuning(0, 32, variant)
        got = lane(0, x, None, None, N, M, 1, lam, out)
        err = (got - wantc).abs().max().item()
        rep = lib.proxtv_lane_stats(1)
        us = timeit(lambda: lane(0, x, None, None, N, M, 1, lam, out))
        print(f"lane CONTIG v={variant}: {us:8.1f} us  {2*M*N*8/us/1e3:7.1f} GB/s  max|diff| {err:.1e} repairs {rep}", flush=True)
        res.append(dict(layout="contig", variant=variant, us=us, err=err))
    # whole-fiber mode on a batch-like shape: 32768 fibers of 512 (same data, viewed as 8 slabs)
    lib.proxtv_lane_tuning(0, 32, 0)
    xb = x.reshape(8, 512, 4096).contiguous()
    us = timeit(lambda: lane(0, xb, None, None, 8 * 4096, 512, 4096, lam, out))
    print(f"lane whole fibers 32768 x 512: {us:8.1f} us  {2*M*N*8/us/1e3:7.1f} GB/s", flush=True)
    OUT["whole_512_us"] = us
    # f32
    x32 = x.float(); o32 = torch.empty_like(x32)
    want32 = old_prox(x32, M, N, M, lam)
    for clen in (0, 256):
        lib.proxtv_lane_tuning(clen, 32, 0)
        got = lane(0, x32, None, None, M, N, M, lam, o32)
        err = (got - want32).abs().max().item()
        us = timeit(lambda: lane(0, x32, None, None, M, N, M, lam, o32))
        print(f"lane f32 clen={clen}: {us:8.1f} us  {2*M*N*4/us/1e3:7.1f} GB/s  max|diff| {err:.1e}", flush=True)
        OUT["f32_%d_us" % clen] = us


if __name__ == "__main__":
    quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
    ok = check_small()
    print("small checks:", "OK" if ok else "FAILED", flush=True)
    bench_big(quick)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(OUT, open("gpurun_out/lanebench.json", "w"), indent=1)
