import sys, os, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from trajopt_b200 import capi
capi.library_path = lambda: os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "trajopt_b200", "csrc", "libtrajopt_b200_prof.so")
from trajopt_b200 import api, problems
B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
d = problems.config1(B=B, T=30)
p = api.Problem(d)
p.lib.tb200_debug_prof(None, 1)
t0 = time.time(); got = p.solve(); dt = time.time() - t0
prof = (C.c_ulonglong * 16)()
p.lib.tb200_debug_prof(prof, 0)
names = ["rows1", "scatter", "solve", "rows2", "xloop", "info+check", "factor", "scale", "qp_solve_warp total", "launch-trajs", "iters-sum"]
tot_iters = got["n_admm_iters"].sum()
print(f"B={B} wall {dt:.2f}s, total ADMM iters {tot_iters}, qp solves {got['n_qp_solves'].sum()}, launches*trajs {prof[9]}")
for i, n in enumerate(names[:9]):
    print(f"  {n:22s} {prof[i]/1e6:10.1f} Mcycles  per-iter {prof[i]/max(tot_iters,1):9.0f} cycles")
