"""ctypes loaders for the two CPU checkers (test infrastructure only, see oracle/tv_oracle.c):

* ``port``  -- oracle/liborc.so, this repo's C restatement of the reference algorithms (oracle/tv_oracle.c);
* ``ref``   -- oracle/_ref/libproxtv_ref.so, the UNMODIFIED reference compiled by oracle/build_ref.sh from
               /root/reference/src (absent if that tree never existed on this machine and no prebuilt copy travelled).

Both expose the same numpy-level helpers so tests can run every check against either.
All 2D/ND arrays are column-major (Fortran order) float64, like the reference's Python wrapper
(prox_tv/__init__.py:402,575).
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


def _p(a):
    return a.ctypes.data_as(_dp)


def build(ref=True):
    """Compile liborc.so (always) and oracle/_ref (when the reference tree is present)."""
    subprocess.check_call(["make", "-s", "-C", _HERE, "liborc.so"])
    if ref:
        subprocess.check_call([os.path.join(_HERE, "build_ref.sh")], stdout=subprocess.DEVNULL)


def _f64(a, order="C"):
    return np.require(np.asarray(a, dtype=np.float64), requirements=["A", "W", order == "F" and "F" or "C"])


class Port:
    """oracle/liborc.so (kind = "port")."""
    kind = "port"

    def __init__(self):
        path = os.path.join(_HERE, "liborc.so")
        if not os.path.exists(path):
            build(ref=False)
        self.lib = L = C.CDLL(path)
        L.orc_tv1_hybrid.argtypes = [_dp, C.c_int, C.c_double, _dp, C.c_double, C.POINTER(C.c_long)]
        L.orc_tv1_hybrid.restype = None
        L.orc_tv1_linearized.argtypes = [_dp, C.c_int, C.c_double, _dp]
        L.orc_tv1_classic.argtypes = [_dp, C.c_int, C.c_double, _dp]
        L.orc_tv1_condat.argtypes = [_dp, _dp, C.c_int, C.c_double]
        L.orc_tv1_weighted.argtypes = [_dp, _dp, _dp, C.c_int]
        L.orc_dr2_tv.argtypes = [C.c_size_t, C.c_size_t, _dp, C.c_double, C.c_double, _dp, C.c_int, _dp]
        L.orc_dr2l1w_tv.argtypes = [C.c_size_t, C.c_size_t, _dp, _dp, _dp, _dp, C.c_int, _dp]
        L.orc_pd2_tv.argtypes = [_dp, _dp, _dp, _dp, _dp, _ip, C.c_int, C.c_int, C.c_int]
        L.orc_pd_tv.argtypes = [_dp, _dp, _dp, _dp, _dp, _ip, C.c_int, C.c_int, C.c_int]
        L.orc_pdr_tv.argtypes = [_dp, _dp, _dp, _dp, _dp, _ip, C.c_int, C.c_int, C.c_int]
        L.orc_jump_set.argtypes = [_dp, C.c_long, C.c_long, _ip]
        L.orc_jump_set.restype = C.c_long

    # ---- 1D ----
    def tv1_hybrid(self, y, lam, exp=1.05, return_steps=False):
        y = _f64(y).ravel(); x = np.empty_like(y); st = C.c_long(0)
        self.lib.orc_tv1_hybrid(_p(y), y.size, float(lam), _p(x), float(exp), C.byref(st))
        return (x, st.value) if return_steps else x

    def tv1_linearized(self, y, lam):
        y = _f64(y).ravel(); x = np.empty_like(y)
        self.lib.orc_tv1_linearized(_p(y), y.size, float(lam), _p(x)); return x

    def tv1_classic(self, y, lam):
        y = _f64(y).ravel(); x = np.empty_like(y)
        self.lib.orc_tv1_classic(_p(y), y.size, float(lam), _p(x)); return x

    def tv1_condat(self, y, lam):
        y = _f64(y).ravel(); x = np.empty_like(y)
        self.lib.orc_tv1_condat(_p(y), _p(x), y.size, float(lam)); return x

    def tv1_weighted(self, y, w):
        y = _f64(y).ravel(); w = _f64(w).ravel(); x = np.empty_like(y)
        wpad = np.concatenate([w, [0.0]])   # the reference reads lam[0] even when n == 1
        self.lib.orc_tv1_weighted(_p(y), _p(wpad), _p(x), y.size); return x

    # ---- 2D / ND ----
    def dr2_tv(self, Y, w1, w2=None, maxit=0, n_threads=1):
        Y = _f64(Y, "F"); out = np.zeros(Y.shape, order="F"); info = np.zeros(3)
        self.lib.orc_dr2_tv(Y.shape[0], Y.shape[1], _p(Y), float(w1), float(w1 if w2 is None else w2), _p(out),
                            int(maxit), _p(info))
        return out, info

    def dr2l1w_tv(self, Y, W1, W2, maxit=0, n_threads=1):
        """W1: (M-1, N) weights of the edges along columns, W2: (M, N-1) along rows (prox_tv.tv1w_2d's w_col, w_row)."""
        Y = _f64(Y, "F"); W1 = _f64(W1, "F"); W2 = _f64(W2, "F"); out = np.zeros(Y.shape, order="F"); info = np.zeros(3)
        assert W1.shape == (Y.shape[0] - 1, Y.shape[1]) and W2.shape == (Y.shape[0], Y.shape[1] - 1)
        self.lib.orc_dr2l1w_tv(Y.shape[0], Y.shape[1], _p(Y), _p(np.append(W1.ravel("F"), 0.0)), _p(np.append(W2.ravel("F"), 0.0)),
                               _p(out), int(maxit), _p(info))
        return out, info

    def _pd(self, fn, Y, ws, ds, maxit):
        Y = _f64(Y, "F"); out = np.zeros(Y.shape, order="F"); info = np.zeros(3)
        ws = np.array(ws, dtype=np.float64); ds = np.array(ds, dtype=np.float64)
        ns = np.array(Y.shape, dtype=np.int32)
        fn(_p(Y), _p(ws), _p(ds), _p(out), _p(info), ns.ctypes.data_as(_ip), Y.ndim, len(ws), int(maxit))
        return out, info, ws

    def pd2_tv(self, Y, ws, ds, maxit=0, n_threads=1):
        return self._pd(self.lib.orc_pd2_tv, Y, ws, ds, maxit)[:2]

    def pd_tv(self, Y, ws, ds, maxit=0, n_threads=1):
        return self._pd(self.lib.orc_pd_tv, Y, ws, ds, maxit)[:2]

    def pdr_tv(self, Y, ws, ds, maxit=0, n_threads=1):
        return self._pd(self.lib.orc_pdr_tv, Y, ws, ds, maxit)[:2]

    def jump_set(self, x):
        x = _f64(x).ravel(); idx = np.empty(max(x.size, 1), dtype=np.int32)
        c = self.lib.orc_jump_set(_p(x), x.size, 1, idx.ctypes.data_as(_ip))
        return idx[:c].copy()


class Ref:
    """oracle/_ref/libproxtv_ref.so -- the reference's own code (kind = "reference"); symbols per src/TVopt.h:88-141."""
    kind = "reference"

    def __init__(self):
        path = os.path.join(_HERE, "_ref", "libproxtv_ref.so")
        if not os.path.exists(path):
            build(ref=True)
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.lib = L = C.CDLL(path)
        L.hybridTautString_TV1.argtypes = [_dp, C.c_int, C.c_double, _dp]; L.hybridTautString_TV1.restype = None
        L.hybridTautString_TV1_custom.argtypes = [_dp, C.c_int, C.c_double, _dp, C.c_double]
        L.hybridTautString_TV1_custom.restype = None
        L.linearizedTautString_TV1.argtypes = [_dp, C.c_double, _dp, C.c_int]
        L.classicTautString_TV1.argtypes = [_dp, C.c_int, C.c_double, _dp]
        L.TV1D_denoise.argtypes = [_dp, _dp, C.c_int, C.c_double]; L.TV1D_denoise.restype = None
        L.tautString_TV1_Weighted.argtypes = [_dp, _dp, _dp, C.c_int]
        L.DR2_TV.argtypes = [C.c_size_t, C.c_size_t, _dp, C.c_double, C.c_double, C.c_double, C.c_double, _dp,
                             C.c_int, C.c_int, _dp]
        L.DR2L1W_TV.argtypes = [C.c_size_t, C.c_size_t, _dp, _dp, _dp, _dp, C.c_int, C.c_int, _dp]
        for f in (L.PD2_TV, L.PD_TV, L.PDR_TV):
            f.argtypes = [_dp, _dp, _dp, _dp, _dp, _dp, _ip, C.c_int, C.c_int, C.c_int, C.c_int]

    def tv1_hybrid(self, y, lam, exp=None):
        y = _f64(y).ravel(); x = np.empty_like(y)
        if exp is None:
            self.lib.hybridTautString_TV1(_p(y), y.size, float(lam), _p(x))
        else:
            self.lib.hybridTautString_TV1_custom(_p(y), y.size, float(lam), _p(x), float(exp))
        return x

    def tv1_linearized(self, y, lam):
        y = _f64(y).ravel(); x = np.empty_like(y)
        self.lib.linearizedTautString_TV1(_p(y), float(lam), _p(x), y.size); return x

    def tv1_classic(self, y, lam):
        y = _f64(y).ravel(); x = np.empty_like(y)
        self.lib.classicTautString_TV1(_p(y), y.size, float(lam), _p(x)); return x

    def tv1_condat(self, y, lam):
        y = _f64(y).ravel(); x = np.empty_like(y)
        self.lib.TV1D_denoise(_p(y), _p(x), y.size, float(lam)); return x

    def tv1_weighted(self, y, w):
        y = _f64(y).ravel(); w = _f64(w).ravel(); x = np.empty_like(y)
        wpad = np.concatenate([w, [0.0]])
        self.lib.tautString_TV1_Weighted(_p(y), _p(wpad), _p(x), y.size); return x

    def dr2_tv(self, Y, w1, w2=None, maxit=0, n_threads=1):
        Y = _f64(Y, "F"); out = np.zeros(Y.shape, order="F"); info = np.zeros(3)
        self.lib.DR2_TV(Y.shape[0], Y.shape[1], _p(Y), float(w1), float(w1 if w2 is None else w2), 1.0, 1.0, _p(out),
                        int(n_threads), int(maxit), _p(info))
        return out, info

    def dr2l1w_tv(self, Y, W1, W2, maxit=0, n_threads=1):
        Y = _f64(Y, "F"); W1 = _f64(W1, "F"); W2 = _f64(W2, "F"); out = np.zeros(Y.shape, order="F"); info = np.zeros(3)
        assert W1.shape == (Y.shape[0] - 1, Y.shape[1]) and W2.shape == (Y.shape[0], Y.shape[1] - 1)
        self.lib.DR2L1W_TV(Y.shape[0], Y.shape[1], _p(Y), _p(np.append(W1.ravel("F"), 0.0)), _p(np.append(W2.ravel("F"), 0.0)),
                           _p(out), int(n_threads), int(maxit), _p(info))
        return out, info

    def _pd(self, fn, Y, ws, ds, maxit, n_threads):
        Y = _f64(Y, "F"); out = np.zeros(Y.shape, order="F"); info = np.zeros(3)
        ws = np.array(ws, dtype=np.float64); ds = np.array(ds, dtype=np.float64); ps = np.ones(len(ws))
        ns = np.array(Y.shape, dtype=np.int32)
        fn(_p(Y), _p(ws), _p(ps), _p(ds), _p(out), _p(info), ns.ctypes.data_as(_ip), Y.ndim, len(ws),
           int(n_threads), int(maxit))
        return out, info

    def pd2_tv(self, Y, ws, ds, maxit=0, n_threads=1):
        return self._pd(self.lib.PD2_TV, Y, ws, ds, maxit, n_threads)

    def pd_tv(self, Y, ws, ds, maxit=0, n_threads=1):
        return self._pd(self.lib.PD_TV, Y, ws, ds, maxit, n_threads)

    def pdr_tv(self, Y, ws, ds, maxit=0, n_threads=1):
        return self._pd(self.lib.PDR_TV, Y, ws, ds, maxit, n_threads)


def have_ref():
    return os.path.exists(os.path.join(_HERE, "_ref", "libproxtv_ref.so")) or os.path.isdir("/root/reference/src")


# ---- synthetic inputs of SURVEY.md section 8d: defined in synth_inputs.py at the repository root (bench.py's measured arm and the
# tools use them without touching anything under oracle/); re-exported here so that tests keep writing O.gen_cfg2(...) ----
import sys as _sys
_ROOT = os.path.dirname(_HERE)
if _ROOT not in _sys.path:
    _sys.path.insert(0, _ROOT)
from synth_inputs import gen_cfg1, gen_cfg2, gen_cfg3, gen_cfg4  # noqa: E402,F401
