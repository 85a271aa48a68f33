"""The algebra behind the partition-inverse form of the QP step's linear system and its Gauss-Jordan elimination
(trajopt_b200/csrc/qp_cta_kernel.cuh: PinvPlan, pinv_factor, gj_rows, admm_block_pinv), restated in numpy
(scripts/probes/pinv_proto.py, with the index conventions of the assembled blocks) and checked against dense solves.
CPU only; the device code itself is covered by the parity tests."""
import importlib.util
import os

import numpy as np
import pytest

_SPEC = importlib.util.spec_from_file_location(
    "pinv_proto", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scripts", "probes", "pinv_proto.py"))
proto = importlib.util.module_from_spec(_SPEC)
_SPEC.loader.exec_module(proto)


@pytest.mark.parametrize("M,NB", [(15, 14), (6, 14), (5, 14), (4, 14), (1, 14), (3, 6), (8, 12), (16, 4), (7, 14)])
def test_two_step_solve_equals_a_dense_solve(M, NB):
    rng = np.random.default_rng(100 * M + NB)
    K, SA, SLM = proto.build(M, NB, rng)
    parts, seps, PI, W, Z = proto.pinv_factor(SA, SLM, M, NB)
    assert seps == [b for b in range(M) if b % 4 == 3] and all(len(pb) <= 3 for pb in parts)
    b = rng.standard_normal(M * NB)
    x = proto.pinv_solve(parts, seps, PI, W, Z, b, M, NB)
    np.testing.assert_allclose(x, np.linalg.solve(K, b), rtol=0, atol=1e-12)
    # Z really is the separator rows of the inverse of the whole matrix
    Kinv = np.linalg.inv(K)
    for s, blk in enumerate(seps):
        np.testing.assert_allclose(Z[s * NB:(s + 1) * NB], Kinv[blk * NB:(blk + 1) * NB], atol=1e-12)


@pytest.mark.parametrize("n,scale", [(42, 1.0), (42, 1e4), (14, 1e6), (42, 1e8)])
def test_elimination_with_deferred_scales_is_accurate(n, scale):
    """gj_rows: columns stay put, finished columns / pivot rows carry scales, diagonals in their own registers."""
    rng = np.random.default_rng(n)
    B = rng.standard_normal((n, n))
    D = np.diag(np.exp(rng.uniform(0, np.log(scale), n))) if scale > 1 else np.eye(n)
    A = D @ (B @ B.T + n * np.eye(n)) @ D
    X, Xt = proto.gj_rows(A), np.linalg.inv(A)
    assert np.abs(X - Xt).max() / np.abs(Xt).max() < 1e-13
