"""Committed golden vectors (tests/golden/*.npz, made by tests/golden/make_golden.py from the CPU oracle).
CPU: the oracle still reproduces them.  GPU (-m gpu): the CUDA path reproduces them through the C ABI."""
import os

import numpy as np
import pytest

from golden.make_golden import CASES, perturbed

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _load(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_reproduces_golden(oracle, name):
    g, d = _load(name), CASES[name]()
    np.testing.assert_array_equal(perturbed(d), g["x_eval"])  # the synthetic inputs are part of the contract
    cv = oracle.convexify_batch(d, g["x_eval"])
    np.testing.assert_allclose(cv["coll_rows"], g["coll_rows"], rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(cv["cart_err"], g["cart_err"], rtol=1e-12, atol=1e-14)
    r = oracle.solve_batch(d, n_threads=1)
    assert (r["status"] == g["status"]).all() and (r["n_qp_solves"] == g["n_qp_solves"]).all()
    np.testing.assert_allclose(r["total_cost"], g["total_cost"], rtol=0, atol=1e-9)
    np.testing.assert_allclose(r["x"], g["x"], rtol=0, atol=1e-7)


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
def test_cuda_reproduces_golden(name):
    from trajopt_b200 import api
    g, d = _load(name), CASES[name]()
    p = api.Problem(d)
    cv = p.convexify(g["x_eval"])
    np.testing.assert_allclose(cv["cart_err"], g["cart_err"], rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(cv["cart_jac"], g["cart_jac"], rtol=1e-6, atol=2e-9)  # forward-difference quotient
    np.testing.assert_allclose(cv["cost_vals"], g["cost_vals_at_x"], rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(cv["cnt_viols"], g["cnt_viols_at_x"], rtol=1e-9, atol=1e-12)
    if g["coll_rows"].size:
        np.testing.assert_allclose(cv["coll_rows"], g["coll_rows"], rtol=1e-10, atol=1e-12)
    r = p.solve()
    p.close()
    # configs[3]: a trajectory cut off by an iteration limit amplifies back-end rounding (tests/test_gpu_parity.py)
    sel = (g["status"] == 0) if name.startswith("cfg3") else np.ones(len(g["status"]), bool)
    assert (r["status"][sel] == g["status"][sel]).all() and (r["n_qp_solves"][sel] == g["n_qp_solves"][sel]).all()
    np.testing.assert_allclose(r["total_cost"][sel], g["total_cost"][sel], rtol=0, atol=1e-6)  # north_star: final cost within 1e-6
    np.testing.assert_allclose(r["x"][sel], g["x"][sel], rtol=0, atol=1e-5)
