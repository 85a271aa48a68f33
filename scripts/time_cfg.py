"""Times one batched solve of a named config on the GPU: python scripts/time_cfg.py cfg3|cfg4|cfg2 B T"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from trajopt_b200 import api, problems

name, B, T = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
d = problems.CONFIGS[name](B=B, T=T)
p = api.Problem(d)
for rep in range(2):
    t0 = time.time()
    got = p.solve()
    dt = time.time() - t0
    tm = got["timing"]
    it = got["n_admm_iters"]
    print(f"{name} B={B} T={T}: wall {dt*1e3:.1f} ms gpu {tm['total_ms']:.1f} ms qp {tm['qp_ms']:.1f} eval {tm['convexify_ms']:.1f} converged "
          f"{(got['status']==0).sum()} hist {np.bincount(got['status'], minlength=5)} qp/traj {got['n_qp_solves'].mean():.1f} "
          f"admm iters mean {it.mean():.0f} max {it.max()} -> {tm['qp_ms']*1e3/max(it.max(),1):.2f} us per iteration of the longest trajectory (upper bound), "
          f"traj/s {(got['status']==0).sum()/dt:.1f}")
