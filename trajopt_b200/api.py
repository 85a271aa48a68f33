"""Python host layer over the C ABI (include/trajopt_b200.h).  CUDA only: `Problem` raises when the
library is missing or no device is visible — there is no CPU fallback (the CPU oracle lives under oracle/
and is test infrastructure)."""
import ctypes as C

import numpy as np

from . import capi

_dbl_p = C.POINTER(C.c_double)
_i32_p = C.POINTER(C.c_int32)


def _dp(a):
    return a.ctypes.data_as(_dbl_p)


def _ip(a):
    return a.ctypes.data_as(_i32_p)


class Problem:
    """Owns a tb200_problem handle (device buffers for one batched description)."""

    def __init__(self, desc, device=0):
        self.lib = capi.load_library()
        self.desc = desc
        self.handle = C.c_void_p()
        rc = self.lib.tb200_problem_create(C.byref(desc.c), device, C.byref(self.handle))
        if rc != 0:
            raise RuntimeError(f"tb200_problem_create failed ({rc}): {self.lib.tb200_last_error().decode()}")
        self.layout = capi.Layout()
        self._check(self.lib.tb200_problem_layout(self.handle, C.byref(self.layout)))

    def _check(self, rc):
        if rc != 0:
            raise RuntimeError(f"trajopt_b200 error {rc}: {self.lib.tb200_last_error().decode()}")

    def close(self):
        if self.handle:
            self.lib.tb200_problem_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_inputs(self, init_traj=None, cart_targets=None, obstacles=None):
        def ptr(a):
            return None if a is None else _dp(np.ascontiguousarray(a, dtype=np.float64))
        keep = [np.ascontiguousarray(a, dtype=np.float64) if a is not None else None for a in (init_traj, cart_targets, obstacles)]
        self._check(self.lib.tb200_problem_set_inputs(self.handle, *[None if a is None else _dp(a) for a in keep]))

    def _results(self):
        L, d = self.layout, self.desc
        return capi.alloc_results(d.B, d.T, d.D, L.n_costs, L.n_cnts)

    def solve(self):
        """BasicTrustRegionSQP::optimize() for the whole batch; host buffers in and out."""
        buf, res = self._results()
        self._check(self.lib.tb200_solve_batch(self.handle, C.byref(res)))
        buf["timing"] = self.timing()
        return buf

    def solve_resident(self):
        self._check(self.lib.tb200_solve_batch_resident(self.handle))

    def fetch(self):
        buf, res = self._results()
        self._check(self.lib.tb200_fetch_results(self.handle, C.byref(res)))
        buf["timing"] = self.timing()
        return buf

    def timing(self):
        t = capi.Timing()
        self._check(self.lib.tb200_last_timing(self.handle, C.byref(t)))
        return {k: getattr(t, k) for k, _ in capi.Timing._fields_}

    def convexify(self, x):
        L, d = self.layout, self.desc
        x = np.ascontiguousarray(x, dtype=np.float64)
        out = dict(cart_err=np.zeros((d.B, max(L.n_cart_rows, 1))),
                   cart_jac=np.zeros((d.B, max(L.n_cart_rows, 1), max(L.cart_jac_stride, 1))),
                   coll_rows=np.zeros((d.B, max(L.n_coll_cand, 1), L.coll_row_stride)),
                   cost_vals=np.zeros((d.B, max(L.n_costs, 1))), cnt_viols=np.zeros((d.B, max(L.n_cnts, 1))))
        co = capi.ConvexifyOut(*[_dp(out[k]) if n else None for k, n in
                                 (("cart_err", L.n_cart_rows), ("cart_jac", L.n_cart_rows), ("coll_rows", L.n_coll_cand),
                                  ("cost_vals", L.n_costs), ("cnt_viols", L.n_cnts))])
        self._check(self.lib.tb200_convexify_batch(self.handle, _dp(x), C.byref(co)))
        out["cart_err"] = out["cart_err"][:, :L.n_cart_rows]
        out["cart_jac"] = out["cart_jac"][:, :L.n_cart_rows]
        out["coll_rows"] = out["coll_rows"][:, :L.n_coll_cand]
        out["cost_vals"] = out["cost_vals"][:, :L.n_costs]
        out["cnt_viols"] = out["cnt_viols"][:, :L.n_cnts]
        return out

    def convexify_timed(self, x):
        """One full-batch launch of the convexify kernel at `x` without fetching its rows; returns the device
        time (CUDA events on the launching stream) and the algorithmic bytes of the launch."""
        x = np.ascontiguousarray(x, dtype=np.float64)
        co = capi.ConvexifyOut(None, None, None, None, None)
        self._check(self.lib.tb200_convexify_batch(self.handle, _dp(x), C.byref(co)))
        return self.timing()

    def qp_solve(self, x, trust, merit_coeffs):
        L, d = self.layout, self.desc
        x = np.ascontiguousarray(x, dtype=np.float64)
        trust = np.ascontiguousarray(np.broadcast_to(trust, (d.B,)), dtype=np.float64)
        mc = np.ascontiguousarray(np.broadcast_to(merit_coeffs, (d.B, max(L.n_cnts, 1))), dtype=np.float64)
        out = dict(new_x=np.zeros((d.B, d.T, d.D)), qp_status=np.zeros(d.B, np.int32),
                   model_cost_vals=np.zeros((d.B, max(L.n_costs, 1))), model_cnt_viols=np.zeros((d.B, max(L.n_cnts, 1))),
                   admm_iters=np.zeros(d.B, np.int32))
        self._check(self.lib.tb200_qp_solve_batch(self.handle, _dp(x), _dp(trust), _dp(mc), _dp(out["new_x"]),
                                                  _ip(out["qp_status"]), _dp(out["model_cost_vals"]) if L.n_costs else None,
                                                  _dp(out["model_cnt_viols"]) if L.n_cnts else None, _ip(out["admm_iters"])))
        out["model_cost_vals"] = out["model_cost_vals"][:, :L.n_costs]
        out["model_cnt_viols"] = out["model_cnt_viols"][:, :L.n_cnts]
        out["polish"] = np.zeros(d.B, np.int32)
        self._check(self.lib.tb200_last_qp_polish(self.handle, _ip(out["polish"])))
        return out


def solve(desc, device=0):
    """One-shot: create, solve, destroy."""
    p = Problem(desc, device)
    try:
        return p.solve()
    finally:
        p.close()


def qp_solve_general(P, q, A, l, u, settings=None, device=0):
    """One sco::Model::optimize() worth of QP on the GPU: min 1/2 x'Px + q'x s.t. l <= Ax <= u (dense inputs; a leading
    batch dimension is allowed).  Returns dict(x, y, status, iters, polish) with OSQP's status values."""
    lib = capi.load_library()
    P = np.ascontiguousarray(P, dtype=np.float64)
    batched = P.ndim == 3
    if not batched:
        P, q, A, l, u = (np.asarray(a, dtype=np.float64)[None] for a in (P, q, A, l, u))
    P, q, A, l, u = (np.ascontiguousarray(a, dtype=np.float64) for a in (P, q, A, l, u))
    B, n = q.shape
    m = l.shape[1]
    g = capi.QpGeneral(n, m, B, 0, _dp(P), _dp(q), _dp(A) if m else None, _dp(l) if m else None, _dp(u) if m else None)
    x, y = np.zeros((B, n)), np.zeros((B, max(m, 1)))
    status, iters, polish = np.zeros(B, np.int32), np.zeros(B, np.int32), np.zeros(B, np.int32)
    rc = lib.tb200_qp_solve_general(C.byref(g), C.byref(settings) if settings is not None else None, device, _dp(x), _dp(y),
                                    _ip(status), _ip(iters), _ip(polish))
    if rc != 0:
        lib.tb200_qp_general_last_error.restype = C.c_char_p
        raise RuntimeError(f"tb200_qp_solve_general failed ({rc}): {lib.tb200_qp_general_last_error().decode()}")
    out = dict(x=x, y=y[:, :m], status=status, iters=iters, polish=polish)
    return out if batched else {k: v[0] for k, v in out.items()}
