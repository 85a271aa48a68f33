import sys, os, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
# needs a build of solve_kernels.cu with -DTB200_PROFILE, selected with TB200_LIB=<path to that .so>
from trajopt_b200 import api, problems
B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
cfg = sys.argv[2] if len(sys.argv) > 2 else "cfg2"
d = problems.CONFIGS[cfg](B=B, T={"cfg3": 50, "cfg4": 40}.get(cfg, 30))
p = api.Problem(d)
p.lib.tb200_debug_prof(None, 1)
t0 = time.time(); got = p.solve(); dt = time.time() - t0
prof = (C.c_ulonglong * 16)()
p.lib.tb200_debug_prof(prof, 0)
names = ["rows_coef", "scatter", "solve", "rows+vars", "-", "info+check", "factor(initial)", "scale", "qp_solve total", "launch-trajs", "iters-sum"]
tot_iters = got["n_admm_iters"].sum()
print(f"B={B} wall {dt:.2f}s, total ADMM iters {tot_iters}, qp solves {got['n_qp_solves'].sum()}, launches*trajs {prof[9]}")
for i, n in enumerate(names[:9]):
    print(f"  {n:22s} {prof[i]/1e6:10.1f} Mcycles  per-iter {prof[i]/max(tot_iters,1):9.0f} cycles")
nq = got['n_qp_solves'].sum()
for slot, n in ((4, "solve: wait for the rhs barrier"), (14, "solve: level 0 down"), (15, "solve: upper levels down"), (9, "solve: upper levels up"),):
    print(f"  {n:32s} {prof[slot]/1e6:10.1f} Mcycles  per-iter {prof[slot]/max(tot_iters,1):9.0f} cycles")
print(f"  assemble (all calls) {prof[10]/1e6:10.1f} Mcycles, bcr_factor (all calls) {prof[11]/1e6:10.1f} Mcycles, polish passes {prof[12]/1e6:10.1f} Mcycles in {prof[13]} polishes ({prof[13]/nq:.2f} per QP)")
