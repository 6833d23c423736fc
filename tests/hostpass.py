"""Runs the product's per-cell bodies on the HOST (tests/hostcheck) through the same ForestDesc the C ABI takes.
Test infrastructure: lets CPU-only tests compare whole passes with the oracle bit for bit."""
import ctypes as C
import os

import numpy as np

from gfp_gaussian_process_b200 import _lib
from gfp_gaussian_process_b200.forest import NOISE_MODELS, DIVISION_MODELS

_HC = None


def hc():
    global _HC
    if _HC is None:
        _HC = C.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)), "hostcheck", "libhostcheck.so"))
    return _HC


def make_desc(data, compute_init=True):
    d = _lib.ForestDesc()
    d.n_cells, d.n_ctp = data.n_cells, data.n_ctp
    d.cell_offset = data.cell_offset.ctypes.data_as(_lib.c_int64_p)
    d.parent = data.parent.ctypes.data_as(_lib.c_int32_p)
    d.daughter1 = data.daughter1.ctypes.data_as(_lib.c_int32_p)
    d.daughter2 = data.daughter2.ctypes.data_as(_lib.c_int32_p)
    d.time = data.time.ctypes.data_as(_lib.c_double_p)
    d.log_length = data.log_length.ctypes.data_as(_lib.c_double_p)
    d.fp = data.fp.ctypes.data_as(_lib.c_double_p)
    d.segment = data.segment.ctypes.data_as(_lib.c_int32_p)
    d.noise_model = NOISE_MODELS[data.noise_model]
    d.division_model = DIVISION_MODELS[data.division_model]
    d.fp_auto = data.fp_auto
    if data.init_f is not None:
        d.init_f = (C.c_double * 4)(*[float(v) for v in data.init_f])
        d.init_r = (C.c_double * 4)(*[float(v) for v in data.init_r])
        d.compute_init = 0
    else:
        d.compute_init = 1
    return d


def host_loglik(data, params, carry=None):
    p = np.ascontiguousarray(params, dtype=np.float64).reshape(-1, 11)
    n_vec = p.shape[0]
    cell_ll = np.zeros((n_vec, data.n_cells))
    nan = np.zeros(n_vec, dtype=np.int64)
    d = make_desc(data)
    rc = hc().hc_loglik(C.byref(d), p.ctypes.data_as(_lib.c_double_p), n_vec,
                        carry.ctypes.data_as(_lib.c_double_p) if carry is not None else None,
                        cell_ll.ctypes.data_as(_lib.c_double_p), nan.ctypes.data_as(C.POINTER(C.c_longlong)))
    assert rc == 0
    return cell_ll, nan


def host_loglik_coop(data, params):
    """the cooperative (four-role) likelihood step of ggp_coop.cuh run on the host; fresh mode"""
    p = np.ascontiguousarray(params, dtype=np.float64).reshape(-1, 11)
    n_vec = p.shape[0]
    cell_ll = np.zeros((n_vec, data.n_cells))
    nan = np.zeros(n_vec, dtype=np.int64)
    d = make_desc(data)
    rc = hc().hc_loglik_coop(C.byref(d), p.ctypes.data_as(_lib.c_double_p), n_vec, cell_ll.ctypes.data_as(_lib.c_double_p),
                             nan.ctypes.data_as(C.POINTER(C.c_longlong)))
    assert rc == 0
    return cell_ll, nan


def host_predict(data, params):
    p = np.ascontiguousarray(params, dtype=np.float64).reshape(-1, 11)
    M = data.n_ctp
    fwd, bwd, comb = np.zeros((M, 20)), np.zeros((M, 20)), np.zeros((M, 20))
    bstate = np.zeros((data.n_cells, 20))
    d = make_desc(data)
    rc = hc().hc_predict(C.byref(d), p.ctypes.data_as(_lib.c_double_p), p.shape[0], fwd.ctypes.data_as(_lib.c_double_p),
                         bwd.ctypes.data_as(_lib.c_double_p), comb.ctypes.data_as(_lib.c_double_p),
                         bstate.ctypes.data_as(_lib.c_double_p))
    assert rc == 0
    sp = lambda a: (a[:, :4], a[:, 4:].reshape(-1, 4, 4))
    return {"forward": sp(fwd), "backward": sp(bwd), "prediction": sp(comb), "bstate": sp(bstate)}


def host_predict_coop(data, params):
    """forward / backward passes with the cooperative step (what the GPU kernels run), on the host"""
    p = np.ascontiguousarray(params, dtype=np.float64).reshape(-1, 11)
    M = data.n_ctp
    fwd, bwd = np.zeros((M, 20)), np.zeros((M, 20))
    bstate = np.zeros((data.n_cells, 20))
    d = make_desc(data)
    rc = hc().hc_predict_coop(C.byref(d), p.ctypes.data_as(_lib.c_double_p), p.shape[0], fwd.ctypes.data_as(_lib.c_double_p),
                              bwd.ctypes.data_as(_lib.c_double_p), bstate.ctypes.data_as(_lib.c_double_p))
    assert rc == 0
    sp = lambda a: (a[:, :4], a[:, 4:].reshape(-1, 4, 4))
    return {"forward": sp(fwd), "backward": sp(bwd), "bstate": sp(bstate)}


def host_joints(data, params, tol, cap):
    """sparse joints from the product's host-compiled code, sorted by (row, col)"""
    p = np.ascontiguousarray(params, dtype=np.float64).reshape(-1, 11)
    row = np.zeros(cap, dtype=np.int64)
    col = np.zeros(cap, dtype=np.int64)
    rec = np.zeros((cap, 44))
    d = make_desc(data)
    hc().hc_joints.restype = C.c_longlong
    n = hc().hc_joints(C.byref(d), p.ctypes.data_as(_lib.c_double_p), p.shape[0], C.c_double(tol), C.c_longlong(cap),
                       row.ctypes.data_as(C.POINTER(C.c_longlong)), col.ctypes.data_as(C.POINTER(C.c_longlong)),
                       rec.ctypes.data_as(_lib.c_double_p))
    assert n >= 0
    k = min(n, cap)
    order = np.lexsort((col[:k], row[:k]))
    return n, row[:k][order], col[:k][order], rec[:k][order]
