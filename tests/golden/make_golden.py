"""Regenerates tests/golden/*.npz.

The reference (C++: Eigen + tesseract + OSQP) cannot be built or imported in this image, so these vectors are
outputs of the CPU oracle (oracle/), which is itself pinned against the reference's own known-answer tests in
tests/test_oracle_golden.py.  They freeze today's oracle so that (a) an accidental change of the oracle shows up in
the CPU suite and (b) the GPU suite can check the CUDA path without building the oracle.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]

from trajopt_b200 import problems  # noqa: E402
import oracle_lib  # noqa: E402

CASES = {"cfg1_B4_T12": lambda: problems.config1(B=4, T=12), "cfg2_B4_T12": lambda: problems.config2(B=4, T=12),
         "cfg2_B2_T30": lambda: problems.config2(B=2, T=30),
         "cfg3_B4_T12": lambda: problems.config3(B=4, T=12, via_every=4)}


def perturbed(desc):
    rng = np.random.default_rng(11)
    return desc.init_traj + 0.05 * rng.standard_normal(desc.init_traj.shape)


def main():
    oracle_lib.build()
    for name, make in CASES.items():
        d = make()
        x = perturbed(d)
        cv = oracle_lib.convexify_batch(d, x)
        r = oracle_lib.solve_batch(d, n_threads=1)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), x_eval=x, cart_err=cv["cart_err"], cart_jac=cv["cart_jac"],
                            coll_rows=cv["coll_rows"], cost_vals_at_x=cv["cost_vals"], cnt_viols_at_x=cv["cnt_viols"],
                            x=r["x"], status=r["status"], total_cost=r["total_cost"], cost_vals=r["cost_vals"],
                            cnt_viols=r["cnt_viols"], n_qp_solves=r["n_qp_solves"])
        print(name, "status", r["status"].tolist(), "cost", np.round(r["total_cost"], 6).tolist())


if __name__ == "__main__":
    main()
