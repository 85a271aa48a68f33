"""Per-iteration latency probe: tiny batches, so every CTA has an SM to itself."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from trajopt_b200 import api, problems
name = sys.argv[1]; B = int(sys.argv[2])
d = {"cfg1": problems.config1, "cfg2": problems.config2}[name](B=B, T=30)
p = api.Problem(d)
for rep in range(2):
    got = p.solve(); tm = got["timing"]
it = got["n_admm_iters"]; nq = got["n_qp_solves"]
print(f"{name} B={B} slice={os.environ.get('TB200_SLICE','100')}: gpu {tm['total_ms']:.1f} ms steps {tm['outer_steps']} qp {tm['qp_ms']:.1f} ms eval {tm['convexify_ms']:.1f} ms"
      f" | max iters {it.max()} (qps {nq[it.argmax()]}) -> {1e3*tm['qp_ms']/it.max():.2f} us per iteration of the longest trajectory; per step {1e3*tm['qp_ms']/tm['outer_steps']:.0f} us")
