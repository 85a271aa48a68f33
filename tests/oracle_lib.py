"""Loader for the CPU oracle (oracle/liboracle.so).  TEST INFRASTRUCTURE: imported only by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg."""
import ctypes as C
import os
import subprocess

import numpy as np

from trajopt_b200 import capi

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_LIB = None
_dbl_p = C.POINTER(C.c_double)
_i32_p = C.POINTER(C.c_int32)


def build():
    subprocess.run(["make", "-C", os.path.join(_ROOT, "oracle"), "-s"], check=True)


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_ROOT, "oracle", "liboracle.so")
        if not os.path.exists(path):
            build()
        _LIB = C.CDLL(path)
        _LIB.oracle_last_error.restype = C.c_char_p
        _LIB.oracle_quad_value.restype = C.c_double
    return _LIB


def _dp(a):
    return a.ctypes.data_as(_dbl_p) if a is not None else None


def _ip(a):
    return a.ctypes.data_as(_i32_p) if a is not None else None


def layout(desc):
    L = capi.Layout()
    rc = lib().oracle_layout(C.byref(desc.c), C.byref(L))
    assert rc == 0, lib().oracle_last_error()
    return L


def solve_batch(desc, b0=0, b1=None, n_threads=0, trace_b=-1):
    b1 = desc.B if b1 is None else b1
    L = layout(desc)
    buf, res = capi.alloc_results(desc.B, desc.T, desc.D, L.n_costs, L.n_cnts)
    secs = C.c_double(0)
    trace = np.zeros((4096, 14))
    tlen = C.c_int(0)
    rc = lib().oracle_solve_batch(C.byref(desc.c), b0, b1, n_threads, C.byref(res), C.byref(secs), trace_b,
                                  _dp(trace), 4096, C.byref(tlen))
    assert rc == 0, lib().oracle_last_error()
    buf["seconds"] = secs.value
    buf["trace"] = trace[:tlen.value]
    return buf


def convexify_batch(desc, x, b0=0, b1=None):
    b1 = desc.B if b1 is None else b1
    L = layout(desc)
    x = np.ascontiguousarray(x, dtype=np.float64)
    out = dict(cart_err=np.zeros((desc.B, max(L.n_cart_rows, 1))),
               cart_jac=np.zeros((desc.B, max(L.n_cart_rows, 1), max(L.cart_jac_stride, 1))),
               coll_rows=np.zeros((desc.B, max(L.n_coll_cand, 1), L.coll_row_stride)),
               cost_vals=np.zeros((desc.B, max(L.n_costs, 1))), cnt_viols=np.zeros((desc.B, max(L.n_cnts, 1))))
    co = capi.ConvexifyOut(*[_dp(out[k]) for k in ("cart_err", "cart_jac", "coll_rows", "cost_vals", "cnt_viols")])
    rc = lib().oracle_convexify_batch(C.byref(desc.c), b0, b1, _dp(x), C.byref(co))
    assert rc == 0, lib().oracle_last_error()
    out["cart_err"] = out["cart_err"][:, :L.n_cart_rows]
    out["cart_jac"] = out["cart_jac"][:, :L.n_cart_rows]
    out["coll_rows"] = out["coll_rows"][:, :L.n_coll_cand]
    out["cost_vals"] = out["cost_vals"][:, :L.n_costs]
    out["cnt_viols"] = out["cnt_viols"][:, :L.n_cnts]
    return out


def qp_solve_batch(desc, x, trust, merit_coeffs, b0=0, b1=None):
    b1 = desc.B if b1 is None else b1
    L = layout(desc)
    x = np.ascontiguousarray(x, dtype=np.float64)
    trust = np.ascontiguousarray(np.broadcast_to(trust, (desc.B,)), dtype=np.float64)
    mc = np.ascontiguousarray(np.broadcast_to(merit_coeffs, (desc.B, max(L.n_cnts, 1))), dtype=np.float64)
    out = dict(new_x=np.zeros((desc.B, desc.T, desc.D)), qp_status=np.zeros(desc.B, np.int32),
               model_cost_vals=np.zeros((desc.B, max(L.n_costs, 1))), model_cnt_viols=np.zeros((desc.B, max(L.n_cnts, 1))),
               admm_iters=np.zeros(desc.B, np.int32), kkt=np.zeros((desc.B, 3)), polish=np.zeros(desc.B, np.int32))
    rc = lib().oracle_qp_solve_batch(C.byref(desc.c), b0, b1, _dp(x), _dp(trust), _dp(mc), _dp(out["new_x"]),
                                     _ip(out["qp_status"]), _dp(out["model_cost_vals"]), _dp(out["model_cnt_viols"]),
                                     _ip(out["admm_iters"]), _dp(out["kkt"]), _ip(out["polish"]))
    assert rc == 0, lib().oracle_last_error()
    out["model_cost_vals"] = out["model_cost_vals"][:, :L.n_costs]
    out["model_cnt_viols"] = out["model_cnt_viols"][:, :L.n_cnts]
    return out
