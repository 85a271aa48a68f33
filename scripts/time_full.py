import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from trajopt_b200 import api, problems
name = sys.argv[1]; B = int(sys.argv[2]); nref = int(sys.argv[3]) if len(sys.argv) > 3 else 0
t0 = time.time()
d = {"cfg1": problems.config1, "cfg2": problems.config2}[name](B=B, T=30)
print("gen %.1fs" % (time.time() - t0))
p = api.Problem(d)
for rep in range(2):
    t0 = time.time(); got = p.solve(); dt = time.time() - t0
    tm = got["timing"]
    print(f"{name} B={B}: wall {dt*1e3:.1f} ms  gpu {tm['total_ms']:.1f} ms steps {tm['outer_steps']} qp {tm['qp_ms']:.1f} ms ({tm['qp_launches']}) eval {tm['convexify_ms']:.1f} ms ({tm['convexify_launches']})"
          f" converged {(got['status']==0).sum()} hist {np.bincount(got['status'], minlength=5)} qp/traj {got['n_qp_solves'].mean():.1f} admm/qp {got['n_admm_iters'].sum()/got['n_qp_solves'].sum():.0f} traj/s {(got['status']==0).sum()/dt:.0f}")
if nref:
    import oracle_lib as O
    t0 = time.time(); ref = O.solve_batch(d, b0=0, b1=nref); dt = time.time() - t0
    print(f"oracle {nref} trajs {dt:.2f}s on {O.lib().oracle_num_threads()} threads -> {nref/dt:.1f} traj/s")
    sl = slice(0, nref)
    print("status match", (got["status"][sl] == ref["status"][sl]).mean(), "nqp match", (got["n_qp_solves"][sl] == ref["n_qp_solves"][sl]).mean(),
          "max dcost %.2e" % np.abs(got["total_cost"][sl] - ref["total_cost"][sl]).max(), "max dx %.2e" % np.abs(got["x"][sl] - ref["x"][sl]).max())
    bad = np.where(np.abs(got["total_cost"][sl] - ref["total_cost"][sl]) > 1e-6)[0]
    print("trajectories off by >1e-6 in cost:", bad.tolist())
import ctypes as C
dbg = np.zeros((B, 16))
p.lib.tb200_debug_last_qp(p.handle, dbg.ctypes.data_as(C.POINTER(C.c_double)))
nr = dbg[:, 11]
print("rows of the last QP: mean %.0f p50 %.0f p90 %.0f max %.0f | n_aux mean %.0f | admm iters total per traj: mean %.0f p90 %.0f max %.0f" % (
    nr.mean(), np.percentile(nr, 50), np.percentile(nr, 90), nr.max(), dbg[:, 12].mean(),
    got["n_admm_iters"].mean(), np.percentile(got["n_admm_iters"], 90), got["n_admm_iters"].max()))
