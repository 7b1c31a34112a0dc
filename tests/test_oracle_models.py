"""Oracle model functions vs the reference's own CasADi graphs (golden: tests/golden/model_functions.npz).

The golden values were produced by the unmodified /root/reference/highway_branch_dyn.py and
quadruped_branch_dyn.py (see tests/golden/make_golden.py).  Tolerances: 1e-12 for function values
and analytic derivatives (float64 round-off only); dp is finite-differenced in the oracle (it never
enters a QP), so 1e-7 there.
"""
import os

import numpy as np
import pytest

from oracle import params


@pytest.fixture(scope="module")
def g(golden_dir):
    return np.load(os.path.join(golden_dir, "model_functions.npz"))


@pytest.mark.parametrize("kind", ["hw", "qd"])
def test_model_functions_match_reference(g, kind):
    if kind == "hw":
        model = params.highway_model(lc_target=g["hw_lc_target"])
    else:
        model = params.quadruped_model()
    X, Z, U = g[kind + "_X"], g[kind + "_Z"], g[kind + "_U"]
    for k in range(len(X)):
        A, B, C, xp = model.dyn_linearization(X[k], U[k])
        np.testing.assert_allclose(A, g[kind + "_A"][k], atol=1e-12)
        np.testing.assert_allclose(B, g[kind + "_B"][k], atol=1e-12)
        np.testing.assert_allclose(C, g[kind + "_C"][k], atol=1e-12)
        np.testing.assert_allclose(xp, g[kind + "_xp"][k], atol=1e-12)
        np.testing.assert_allclose(model.zpred_eval(Z[k]), g[kind + "_zpred"][k], atol=1e-12)
        p, dp = model.branch_eval(X[k], Z[k], with_dp=True)
        np.testing.assert_allclose(p, g[kind + "_p"][k], atol=1e-12)
        np.testing.assert_allclose(dp, g[kind + "_dp"][k], atol=1e-7)
        assert abs(p.sum() - 1.0) < 1e-12
        h, dh = model.col_eval(X[k], Z[k])
        np.testing.assert_allclose(h, g[kind + "_hlin"][k], atol=1e-11)
        np.testing.assert_allclose(dh, g[kind + "_dh"][k], atol=1e-12)


def test_linearisation_is_first_order_exact():
    """C = xp - A x - B u, and A is the Jacobian of the Euler step (finite-difference check)."""
    rng = np.random.default_rng(3)
    for model in (params.highway_model(), params.quadruped_model()):
        for _ in range(5):
            x = rng.normal(size=model.n) * np.array([10, 3, 5, 0.2][: model.n]) + np.array([0, 5, 20, 0][: model.n])
            u = rng.normal(size=model.d) * 0.2
            A, B, C, xp = model.dyn_linearization(x, u)
            np.testing.assert_allclose(A @ x + B @ u + C, xp, atol=1e-12)
            for k in range(model.n):
                e = np.zeros(model.n)
                e[k] = 1e-6
                fd = (model.step(x + e, u) - model.step(x - e, u)) / 2e-6
                np.testing.assert_allclose(A[:, k], fd, atol=1e-7)


def test_softmin_shift_invariance_large_arguments():
    """The max-shifted forms agree with the reference's unshifted ones where those do not overflow."""
    from oracle.models import softmin, softmax
    v = np.array([0.3, -1.2, 2.0])
    for gam in (1.0, 5.0):
        assert abs(softmin(v, gam) - np.sum(np.exp(-gam * v) * v) / np.sum(np.exp(-gam * v))) < 1e-13
        assert abs(softmax(v, gam) - np.sum(np.exp(gam * v) * v) / np.sum(np.exp(gam * v))) < 1e-13
    assert np.isfinite(softmin(np.array([500.0, 900.0]), 5.0))
