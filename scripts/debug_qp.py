import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import oracle_lib as O
from trajopt_b200 import api, problems
np.set_printoptions(linewidth=220, precision=4)
name, trust = sys.argv[1], float(sys.argv[2])
d = {"cfg1": problems.config1, "cfg2": problems.config2}[name](B=16, T=12)
x = d.init_traj.copy()
p = api.Problem(d)
got = p.qp_solve(x, trust, 10.0)
ref = O.qp_solve_batch(d, x, trust, 10.0)
conv = O.convexify_batch(d, x)
nact = (conv["coll_rows"][..., -1] != 0).sum(axis=1) if conv["coll_rows"].size else np.zeros(d.B)
import ctypes as C
dbg = np.zeros((d.B, 16))
p.lib.tb200_debug_last_qp(p.handle, dbg.ctypes.data_as(C.POINTER(C.c_double)))
for b in range(d.B):
    if got["polish"][b] != ref["polish"][b]:
        print("   DBG", b, dict(zip("status iters polish rho pri dua pol_pri pol_dua c pol_ok rho_upd nr naux nnzA warm".split(), dbg[b])))
for b in range(d.B):
    dx = np.abs(got["new_x"][b] - ref["new_x"][b]).max()
    print(b, "st", got["qp_status"][b], ref["qp_status"][b], "it", got["admm_iters"][b], ref["admm_iters"][b], "pol", got["polish"][b], ref["polish"][b],
          "nact", nact[b], "dx %.2e" % dx, "kkt", ref["kkt"][b], "dmv %.2e" % np.abs(got["model_cnt_viols"][b]-ref["model_cnt_viols"][b]).max())
