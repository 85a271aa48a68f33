import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from trajopt_b200 import api, problems
B = int(sys.argv[1]) if len(sys.argv) > 1 else 592
cfg = sys.argv[2] if len(sys.argv) > 2 else "cfg1"
d = {"cfg1": problems.config1, "cfg2": problems.config2}[cfg](B=B, T=30)
p = api.Problem(d)
got = p.solve()
print("done", (got["status"] == 0).sum())
