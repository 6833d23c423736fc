"""ctypes wrapper of the CPU oracle (oracle/libggp_oracle.so) and of the reference build (oracle/_ref).

TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline / reference arm may
import this module; the product package (gfp_gaussian_process_b200/) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(_HERE, "libggp_oracle.so")
REF_SO = os.path.join(_HERE, "_ref", "libggp_ref.so")
ORACLE_REF_SO = os.path.join(_HERE, "_ref", "libggp_oracle_ref.so")
REF_WRAPPERS_SO = os.path.join(_HERE, "_ref", "libggp_ref_wrappers.so")
REF_BRIDGE_SO = os.path.join(_HERE, "_ref", "libggp_ref_bridge.so")

dp = C.POINTER(C.c_double)
lp = C.POINTER(C.c_long)
ip = C.POINTER(C.c_int)


class OracleForest(C.Structure):
    _fields_ = [("n_cells", C.c_long), ("n_ctp", C.c_long), ("cell_offset", lp), ("parent", ip), ("daughter1", ip),
                ("daughter2", ip), ("time", dp), ("log_length", dp), ("fp", dp), ("segment", ip),
                ("noise_model", C.c_int), ("division_model", C.c_int), ("fp_auto", C.c_double),
                ("init_f", C.c_double * 4), ("init_r", C.c_double * 4)]


def build(force=False):
    """compile the oracle (and, where /root/reference is mounted, the reference shim)."""
    if force or not os.path.exists(ORACLE_SO):
        subprocess.check_call(["make", "-C", _HERE, "libggp_oracle.so"] + (["-B"] if force else []))
    if os.path.exists("/root/reference/src/mean_cov_model.h") and (force or not os.path.exists(REF_SO) or not os.path.exists(ORACLE_REF_SO)):
        subprocess.check_call(["make", "-C", _HERE, "_ref/libggp_ref.so", "_ref/libggp_oracle_ref.so"])
    if os.path.exists("/root/reference/src/likelihood.h") and (force or not os.path.exists(REF_WRAPPERS_SO)):
        subprocess.check_call(["make", "-C", _HERE, "_ref/libggp_ref_wrappers.so"])
    product = os.path.join(os.path.dirname(_HERE), "gfp_gaussian_process_b200", "libggp_b200.so")
    if os.path.exists("/root/reference/src/likelihood.h") and os.path.exists(product):
        subprocess.check_call(["make", "-s", "-C", _HERE, "_ref/libggp_ref_bridge.so"])


_oracle = None
_ref = None
_use_ref_math = False


def use_reference_math(flag=True):
    """route the oracle's arithmetic core through the reference's own mean_cov_model.h + Faddeeva.cc
    (oracle/_ref/libggp_oracle_ref.so).  Returns False if that build is not present."""
    global _oracle, _use_ref_math
    if flag and not os.path.exists(ORACLE_REF_SO):
        try:
            build()
        except Exception:
            pass
        if not os.path.exists(ORACLE_REF_SO):
            return False
    if flag != _use_ref_math:
        _oracle = None
        _use_ref_math = flag
    return True


ORACLE_ULP_SO = os.path.join(_HERE, "_ref", "libggp_oracle_ulp.so")


def _configure(L):
    L.ggp_oracle_dawson.restype = C.c_double
    L.ggp_oracle_dawson.argtypes = [C.c_double]
    L.ggp_oracle_tauint.restype = C.c_double
    L.ggp_oracle_tauint.argtypes = [C.c_int] + [C.c_double] * 5
    L.ggp_oracle_total_loglik.restype = C.c_double
    L.ggp_oracle_total_loglik.argtypes = [C.POINTER(OracleForest), dp, dp, dp, dp, lp, lp]
    if hasattr(L, 'ggp_oracle_joints'):
        L.ggp_oracle_joints.restype = C.c_long
    return L


def oracle_ulp():
    """the oracle's loop around the reference core with every exp / pow / Dawson result moved +-1 ulp
    (oracle/_ref/libggp_oracle_ulp.so; ggp_ref_set_ulp_seed(seed), seed 0 = untouched); None if it cannot be built"""
    if not os.path.exists(ORACLE_ULP_SO):
        if not os.path.exists("/root/reference/src/mean_cov_model.h"):
            return None
        subprocess.check_call(["make", "-s", "-C", _HERE, "_ref/libggp_oracle_ulp.so"])
    L = _configure(C.CDLL(ORACLE_ULP_SO))
    L.ggp_ref_set_ulp_seed.argtypes = [C.c_ulonglong]
    return L


def oracle():
    global _oracle
    if _oracle is None:
        build()
        L = C.CDLL(ORACLE_REF_SO if _use_ref_math else ORACLE_SO)
        L.ggp_oracle_dawson.restype = C.c_double
        L.ggp_oracle_dawson.argtypes = [C.c_double]
        L.ggp_oracle_tauint.restype = C.c_double
        L.ggp_oracle_tauint.argtypes = [C.c_int] + [C.c_double] * 5
        L.ggp_oracle_total_loglik.restype = C.c_double
        L.ggp_oracle_total_loglik.argtypes = [C.POINTER(OracleForest), dp, dp, dp, dp, lp, lp]
        if hasattr(L, 'ggp_oracle_joints'):
            L.ggp_oracle_joints.restype = C.c_long
        _oracle = L
    return _oracle


def ref():
    """the reference's own mean_cov_model.h + Faddeeva.cc (None if it was never built)."""
    global _ref
    if _ref is None:
        if not os.path.exists(REF_SO):
            try:
                build()
            except Exception:
                pass
        if not os.path.exists(REF_SO):
            return None
        L = C.CDLL(REF_SO)
        for n in ("ggp_ref_dawson", "ggp_ref_exp", "ggp_ref_log"):
            getattr(L, n).restype = C.c_double
            getattr(L, n).argtypes = [C.c_double]
        L.ggp_ref_pow.restype = C.c_double
        L.ggp_ref_pow.argtypes = [C.c_double, C.c_double]
        L.ggp_ref_tauint.restype = C.c_double
        L.ggp_ref_tauint.argtypes = [C.c_int] + [C.c_double] * 5
        _ref = L
    return _ref


def _p(a, t=dp):
    return a.ctypes.data_as(t)


class Oracle:
    """the oracle bound to one LineageData-like object (any object with the same attribute names)."""

    def __init__(self, data, lib=None):
        self.L = lib if lib is not None else oracle()
        self.data = data
        self._keep = dict(
            off=np.ascontiguousarray(data.cell_offset, dtype=np.int64), parent=np.ascontiguousarray(data.parent, dtype=np.int32),
            d1=np.ascontiguousarray(data.daughter1, dtype=np.int32), d2=np.ascontiguousarray(data.daughter2, dtype=np.int32),
            time=np.ascontiguousarray(data.time), x=np.ascontiguousarray(data.log_length), g=np.ascontiguousarray(data.fp),
            seg=np.ascontiguousarray(data.segment, dtype=np.int32))
        k = self._keep
        f = OracleForest()
        f.n_cells, f.n_ctp = len(k["parent"]), len(k["time"])
        f.cell_offset = _p(k["off"], lp)
        f.parent, f.daughter1, f.daughter2 = _p(k["parent"], ip), _p(k["d1"], ip), _p(k["d2"], ip)
        f.time, f.log_length, f.fp, f.segment = _p(k["time"]), _p(k["x"]), _p(k["g"]), _p(k["seg"], ip)
        f.noise_model = {"const": 0, "scaled": 1}[data.noise_model]
        f.division_model = {"gauss": 0, "binomial": 1}[data.division_model]
        f.fp_auto = data.fp_auto
        self.f = f
        if getattr(data, "init_f", None) is not None:
            f.init_f = (C.c_double * 4)(*[float(v) for v in data.init_f])
            f.init_r = (C.c_double * 4)(*[float(v) for v in data.init_r])
        else:
            self.L.ggp_oracle_init_stats(C.byref(f))
        self.n_cells, self.n_ctp = f.n_cells, f.n_ctp
        self.cell_mean = np.zeros((f.n_cells, 4))
        self.cell_cov = np.zeros((f.n_cells, 16))

    def reset(self):
        """fresh MOMAdata::mean/cov, i.e. the state right after get_segment (moma_input.h:587)"""
        self.cell_mean[:] = 0
        self.cell_cov[:] = 0

    def init_stats(self):
        return np.array(self.f.init_f), np.array(self.f.init_r)

    def total_loglik(self, params, fresh=True, per_cell=False):
        if fresh:
            self.reset()
        p = np.ascontiguousarray(params, dtype=np.float64)
        pc = np.zeros(self.n_cells) if per_cell else None
        nc, nt = C.c_long(-1), C.c_long(-1)
        ll = self.L.ggp_oracle_total_loglik(C.byref(self.f), _p(p), _p(self.cell_mean), _p(self.cell_cov),
                                            _p(pc) if per_cell else None, C.byref(nc), C.byref(nt))
        self.nan = (nc.value, nt.value)
        return (ll, pc) if per_cell else ll

    def predictions(self, params_vecs, fresh=True):
        """forward, backward, combined: each (mean [n_ctp][4], cov [n_ctp][4][4])"""
        if fresh:
            self.reset()
        p = np.ascontiguousarray(params_vecs, dtype=np.float64).reshape(-1, 11)
        M = self.n_ctp
        mf, cf, mb, cb, mp, cp = (np.zeros((M, 4)), np.zeros((M, 16)), np.zeros((M, 4)), np.zeros((M, 16)),
                                  np.zeros((M, 4)), np.zeros((M, 16)))
        self.L.ggp_oracle_prediction_forward(C.byref(self.f), _p(p), p.shape[0], _p(self.cell_mean), _p(self.cell_cov), _p(mf), _p(cf))
        self.L.ggp_oracle_prediction_backward(C.byref(self.f), _p(p), p.shape[0], _p(self.cell_mean), _p(self.cell_cov), _p(mb), _p(cb))
        self.L.ggp_oracle_combine_predictions(C.byref(self.f), _p(p), p.shape[0], _p(mf), _p(cf), _p(mb), _p(cb), _p(mp), _p(cp))
        sh = (M, 4, 4)
        self._pred = (p, mf, cf, mb, cb)
        return {"forward": (mf, cf.reshape(sh)), "backward": (mb, cb.reshape(sh)), "prediction": (mp, cp.reshape(sh))}

    def joints(self, tol, cap):
        """collect_joint_distributions as a sparse list; call predictions() first"""
        p, mf, cf, mb, cb = self._pred
        row = np.zeros(cap, dtype=np.int64)
        col = np.zeros(cap, dtype=np.int64)
        rec = np.zeros((cap, 44))
        self.L.ggp_oracle_joints.argtypes = [C.POINTER(OracleForest), dp, C.c_int, C.c_double, dp, dp, dp, dp, dp, C.c_long, lp, lp, dp]
        n = self.L.ggp_oracle_joints(C.byref(self.f), _p(p), p.shape[0], tol, _p(self.cell_mean), _p(mf), _p(cf), _p(mb), _p(cb),
                                     cap, _p(row, lp), _p(col, lp), _p(rec))
        return n, row[:min(n, cap)], col[:min(n, cap)], rec[:min(n, cap)]


def mean_cov_model(state_mean, state_cov16, t, p7, which="oracle"):
    L = oracle() if which == "oracle" else ref()
    fn = L.ggp_oracle_mean_cov_model if which == "oracle" else L.ggp_ref_mean_cov_model
    m = np.ascontiguousarray(state_mean, dtype=np.float64)
    c = np.ascontiguousarray(state_cov16, dtype=np.float64).reshape(16)
    p = np.ascontiguousarray(p7, dtype=np.float64)
    mo, co = np.zeros(4), np.zeros(16)
    fn(_p(m), _p(c), C.c_double(t), _p(p), _p(mo), _p(co))
    return mo, co


def cross_cov_model(state_mean, state_cov16, t, p7, which="oracle"):
    L = oracle() if which == "oracle" else ref()
    fn = L.ggp_oracle_cross_cov_model if which == "oracle" else L.ggp_ref_cross_cov_model
    m = np.ascontiguousarray(state_mean, dtype=np.float64)
    c = np.ascontiguousarray(state_cov16, dtype=np.float64).reshape(16)
    p = np.ascontiguousarray(p7, dtype=np.float64)
    out = np.zeros(16)
    fn(_p(m), _p(c), C.c_double(t), _p(p), _p(out))
    return out


_refw = None
_refb = None


def ref_wrappers(bridge=False):
    """bridge=True: the same build plus the reference-side binding include/ggp_bridge.h, linked to libggp_b200.so.
    the reference's own likelihood.h / predictions.h / correlation_tree.h / Gaussians.h / moma_input.h compiled unmodified
    over oracle/eigen_shim (oracle/ref_wrappers.cpp); None if that build is not present"""
    global _refw, _refb
    so = REF_BRIDGE_SO if bridge else REF_WRAPPERS_SO
    if (_refb if bridge else _refw) is None:
        if not os.path.exists(so):
            try:
                build()
            except Exception:
                pass
        if not os.path.exists(so):
            return None
        L = C.CDLL(so)
        L.ggp_refw_create.restype = C.c_void_p
        L.ggp_refw_create.argtypes = [C.c_long, lp, ip, dp, dp, dp, ip, C.c_int, C.c_int, C.c_double]
        L.ggp_refw_destroy.argtypes = [C.c_void_p]
        L.ggp_refw_get_daughters.argtypes = [C.c_void_p, ip, ip]
        L.ggp_refw_get_init.argtypes = [C.c_void_p, dp, dp]
        L.ggp_refw_get_state.argtypes = [C.c_void_p, dp, dp]
        L.ggp_refw_set_state.argtypes = [C.c_void_p, dp, dp]
        L.ggp_refw_total_loglik.restype = C.c_double
        L.ggp_refw_total_loglik.argtypes = [C.c_void_p, dp]
        L.ggp_refw_per_cell_loglik.restype = C.c_int
        L.ggp_refw_per_cell_loglik.argtypes = [C.c_void_p, dp, dp]
        L.ggp_refw_predict.argtypes = [C.c_void_p, dp, C.c_int] + [dp] * 6
        L.ggp_refw_joints.restype = C.c_long
        L.ggp_refw_joints.argtypes = [C.c_void_p, dp, C.c_int, C.c_double, C.c_long, lp, lp, dp]
        if bridge:
            L.ggp_refb_open.restype = C.c_void_p
            L.ggp_refb_open.argtypes = [C.c_void_p, C.c_int]
            L.ggp_refb_close.argtypes = [C.c_void_p]
            L.ggp_refb_loglik_chain.argtypes = [C.c_void_p, dp, C.c_int, dp]
            L.ggp_refb_loglik_batch.argtypes = [C.c_void_p, dp, C.c_int, dp]
            L.ggp_refb_predict.argtypes = [C.c_void_p, C.c_void_p, dp, C.c_int] + [dp] * 6
            L.ggp_refb_joints_text.restype = C.c_long
            L.ggp_refb_joints_text.argtypes = [C.c_void_p, C.c_void_p, dp, C.c_int, C.c_double, C.c_int, C.c_char_p, C.c_long]
            _refb = L
        else:
            _refw = L
    return _refb if bridge else _refw


class RefWrappers:
    """the reference's own std::vector<MOMAdata> passes (same method names and state handling as Oracle)"""

    def __init__(self, data, bridge=False):
        self.L = ref_wrappers(bridge)
        self.b = None
        if self.L is None:
            raise RuntimeError("oracle/_ref/libggp_ref_wrappers.so / libggp_ref_bridge.so is not built (needs /root/reference)")
        off = np.ascontiguousarray(data.cell_offset, dtype=np.int64)
        par = np.ascontiguousarray(data.parent, dtype=np.int32)
        seg = np.ascontiguousarray(data.segment, dtype=np.int32)
        t, x, g = (np.ascontiguousarray(a, dtype=np.float64) for a in (data.time, data.log_length, data.fp))
        self.n_cells, self.n_ctp = len(par), len(t)
        self.h = self.L.ggp_refw_create(self.n_cells, _p(off, lp), _p(par, ip), _p(t), _p(x), _p(g), _p(seg, ip),
                                        {"const": 0, "scaled": 1}[data.noise_model], {"gauss": 0, "binomial": 1}[data.division_model],
                                        float(data.fp_auto))

    def close(self):
        if self.b:
            self.L.ggp_refb_close(self.b)
            self.b = None
        if self.h:
            self.L.ggp_refw_destroy(self.h)
            self.h = None

    # ---- the reference-side binding (include/ggp_bridge.h) on the GPU, on this object's std::vector<MOMAdata> ----
    def open_bridge(self, device=0):
        self.b = self.L.ggp_refb_open(self.h, device)
        if not self.b:
            raise RuntimeError("GgpBridge could not be created (no GPU?)")

    def bridge_loglik(self, params_rows, batch=False):
        p = np.ascontiguousarray(params_rows, dtype=np.float64).reshape(-1, 11)
        out = np.zeros(p.shape[0])
        rc = (self.L.ggp_refb_loglik_batch if batch else self.L.ggp_refb_loglik_chain)(self.b, _p(p), p.shape[0], _p(out))
        if rc:
            raise RuntimeError("binding threw")
        return out

    def bridge_predictions(self, params_vecs):
        p = np.ascontiguousarray(params_vecs, dtype=np.float64).reshape(-1, 11)
        M = self.n_ctp
        o = [np.zeros((M, 4)), np.zeros((M, 16)), np.zeros((M, 4)), np.zeros((M, 16)), np.zeros((M, 4)), np.zeros((M, 16))]
        if self.L.ggp_refb_predict(self.h, self.b, _p(p), p.shape[0], *[_p(a) for a in o]):
            raise RuntimeError("binding threw")
        self._p = p
        sh = (M, 4, 4)
        return {"forward": (o[0], o[1].reshape(sh)), "backward": (o[2], o[3].reshape(sh)), "prediction": (o[4], o[5].reshape(sh))}

    def joints_text(self, tol, gpu, precision=6, cap=1 << 28):
        buf = C.create_string_buffer(cap)
        n = self.L.ggp_refb_joints_text(self.h, self.b if gpu else None, _p(self._p), self._p.shape[0], tol, precision, buf, cap)
        if n < 0 or n > cap:
            raise RuntimeError("joints text failed or too large")
        return buf.raw[:n]

    def __del__(self):
        self.close()

    def daughters(self):
        d1, d2 = np.zeros(self.n_cells, dtype=np.int32), np.zeros(self.n_cells, dtype=np.int32)
        self.L.ggp_refw_get_daughters(self.h, _p(d1, ip), _p(d2, ip))
        return d1, d2

    def init_stats(self):
        f, r = np.zeros(4), np.zeros(4)
        self.L.ggp_refw_get_init(self.h, _p(f), _p(r))
        return f, r

    def state(self):
        m, c = np.zeros((self.n_cells, 4)), np.zeros((self.n_cells, 16))
        self.L.ggp_refw_get_state(self.h, _p(m), _p(c))
        return m, c

    def reset(self):
        self.L.ggp_refw_set_state(self.h, _p(np.zeros((self.n_cells, 4))), _p(np.zeros((self.n_cells, 16))))

    def total_loglik(self, params, fresh=True, per_cell=False):
        if fresh:
            self.reset()
        p = np.ascontiguousarray(params, dtype=np.float64)
        if per_cell:
            pc = np.zeros(self.n_cells)
            self.L.ggp_refw_per_cell_loglik(self.h, _p(p), _p(pc))
            return pc
        return self.L.ggp_refw_total_loglik(self.h, _p(p))

    def predictions(self, params_vecs, fresh=True):
        if fresh:
            self.reset()
        p = np.ascontiguousarray(params_vecs, dtype=np.float64).reshape(-1, 11)
        M = self.n_ctp
        o = [np.zeros((M, 4)), np.zeros((M, 16)), np.zeros((M, 4)), np.zeros((M, 16)), np.zeros((M, 4)), np.zeros((M, 16))]
        self.L.ggp_refw_predict(self.h, _p(p), p.shape[0], *[_p(a) for a in o])
        self._p = p
        sh = (M, 4, 4)
        return {"forward": (o[0], o[1].reshape(sh)), "backward": (o[2], o[3].reshape(sh)), "prediction": (o[4], o[5].reshape(sh))}

    def joints(self, tol, cap):
        row, col, rec = np.zeros(cap, dtype=np.int64), np.zeros(cap, dtype=np.int64), np.zeros((cap, 44))
        n = self.L.ggp_refw_joints(self.h, _p(self._p), self._p.shape[0], tol, cap, _p(row, lp), _p(col, lp), _p(rec))
        return n, row[:min(n, cap)], col[:min(n, cap)], rec[:min(n, cap)]
