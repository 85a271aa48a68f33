import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import oracle_lib as O
from trajopt_b200 import api, problems
np.set_printoptions(linewidth=200, precision=6)
d = problems.config3(B=16, T=12, via_every=4)
ref = O.solve_batch(d)
p = api.Problem(d)
for name, x in (("init", d.init_traj), ("ref solution", ref["x"])):
    got = p.convexify(x); r = O.convexify_batch(d, x)
    for k in ("cart_err", "cart_jac", "coll_rows", "cost_vals", "cnt_viols"):
        diff = np.abs(got[k] - r[k]); 
        print(name, k, "max abs diff %.3e" % diff.max(), "at", np.unravel_index(diff.argmax(), diff.shape))
    act_g = got["coll_rows"][..., -1] != 0; act_r = r["coll_rows"][..., -1] != 0
    print(name, "active rows gpu/ref", act_g.sum(), act_r.sum(), "mismatch", (act_g != act_r).sum())
cap = 600
p.lib.tb200_debug_enable_trace(p.handle, cap)
got = p.solve()
tr = np.zeros((d.B, cap, 14)); tl = np.zeros(d.B, np.int32)
p.lib.tb200_debug_fetch_trace(p.handle, tr.ctypes.data_as(C.POINTER(C.c_double)), tl.ctypes.data_as(C.POINTER(C.c_int32)))
for b in range(d.B):
    dc = abs(got["total_cost"][b] - ref["total_cost"][b])
    if dc < 1e-6: continue
    rb = O.solve_batch(d, b0=b, b1=b + 1, trace_b=b)["trace"]; gt = tr[b, :tl[b]]
    n = min(len(rb), len(gt))
    print(f"traj {b}: dcost {dc:.2e} status {got['status'][b]}/{ref['status'][b]} nqp {got['n_qp_solves'][b]}/{ref['n_qp_solves'][b]}")
    shown = 0
    for i in range(n):
        rel = lambda a, c: abs(a - c) / max(abs(c), 1e-12)
        bad = rel(gt[i, 4], rb[i, 4]) > 1e-8 or rel(gt[i, 5], rb[i, 5]) > 1e-8 or gt[i, 8] != rb[i, 8] or gt[i, 12] != rb[i, 12]
        if bad and shown < 6:
            shown += 1
            print(f"  entry {i}: it {gt[i,7]:.0f}/{rb[i,7]:.0f} pol {gt[i,12]:.0f}/{rb[i,12]:.0f} act {gt[i,8]:.0f}/{rb[i,8]:.0f} old {gt[i,3]:.10g}/{rb[i,3]:.10g} model {gt[i,4]:.10g}/{rb[i,4]:.10g} new {gt[i,5]:.10g}/{rb[i,5]:.10g} qpst {gt[i,6]:.0f}/{rb[i,6]:.0f}")
