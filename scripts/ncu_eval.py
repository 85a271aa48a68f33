"""Full-batch launches of the convexify kernel only (for ncu): the initial trajectories of a cfg2 batch."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from trajopt_b200 import api, problems
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
cfg = sys.argv[2] if len(sys.argv) > 2 else "cfg2"
d = {"cfg1": problems.config1, "cfg2": problems.config2}[cfg](B=B, T=30)
p = api.Problem(d)
rng = np.random.default_rng(0)
for k in range(4):
    x = d.init_traj + 0.05 * rng.standard_normal(d.init_traj.shape)
    tm = p.convexify_timed(x)
    print(f"launch {k}: {tm['convexify_ms']*1e3:.1f} us, {tm['convexify_bytes']/1e6:.1f} MB algorithmic -> {tm['convexify_bytes']/tm['convexify_ms']/1e6:.0f} GB/s")
