import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import oracle_lib as O
from trajopt_b200 import api, problems
np.set_printoptions(linewidth=220, precision=6)
name = sys.argv[1]
B = int(sys.argv[2]) if len(sys.argv) > 2 else 16
T = int(sys.argv[3]) if len(sys.argv) > 3 else 12
d = {"cfg1": problems.config1, "cfg2": problems.config2,
     "cfg3": lambda B, T: problems.config3(B=B, T=T, via_every=4 if T < 20 else 10)}[name](B=B, T=T)
p = api.Problem(d)
cap = 600
p.lib.tb200_debug_enable_trace(p.handle, cap)
got = p.solve()
tr = np.zeros((d.B, cap, 14)); tl = np.zeros(d.B, np.int32)
p.lib.tb200_debug_fetch_trace(p.handle, tr.ctypes.data_as(C.POINTER(C.c_double)), tl.ctypes.data_as(C.POINTER(C.c_int32)))
only = int(sys.argv[4]) if len(sys.argv) > 4 else -1
for b in range(d.B):
    if only >= 0 and b != only: continue
    ref = O.solve_batch(d, b0=b, b1=b + 1, trace_b=b)
    rt = ref["trace"]
    gt = tr[b, :tl[b]]
    n = min(len(rt), len(gt))
    # compare decisions and merits
    bad = None
    for i in range(n):
        if rt[i, 13] != gt[i, 13] or rt[i, 12] != gt[i, 12] or rt[i, 8] != gt[i, 8] or rt[i, 7] != gt[i, 7] or abs(rt[i, 4] - gt[i, 4]) > 1e-9 * max(1, abs(rt[i, 4])) or abs(rt[i, 5] - gt[i, 5]) > 1e-9 * max(1, abs(rt[i, 5])):
            bad = i
            break
    dc = abs(got["total_cost"][b] - ref["total_cost"][b])
    print(f"traj {b}: status {got['status'][b]}/{ref['status'][b]} nqp {got['n_qp_solves'][b]}/{ref['n_qp_solves'][b]} len {len(gt)}/{len(rt)} dcost {dc:.2e} first_div {bad}")
    if bad is not None:
        for i in range(max(0, bad - 1), min(n, bad + 2)):
            print("   ref", rt[i])
            print("   gpu", gt[i])
    if only >= 0:
        for i in range(n):
            rel = lambda a, c: abs(a - c) / max(abs(c), 1e-300)
            print(i, "it %d/%d pol %d/%d warm %d/%d act %d/%d" % (gt[i,7], rt[i,7], gt[i,12], rt[i,12], gt[i,13], rt[i,13], gt[i,8], rt[i,8]),
                  "rel pri %.1e dua %.1e rho %.1e model %.1e new %.1e" % (rel(gt[i,9], rt[i,9]), rel(gt[i,10], rt[i,10]), rel(gt[i,11], rt[i,11]), rel(gt[i,4], rt[i,4]), rel(gt[i,5], rt[i,5])))
