"""The belief-state MPC oracle (oracle/belief_mpc.py) against fixtures recorded from the UNMODIFIED reference classes
(PredictiveControllers.MPC, HMM_backup_dyn.PredictiveModel, Init_MPC.initMPCParams; tests/golden/make_golden.py `belief`):
the model's linearisation at every stage, every matrix of the assembled QP, and the closed-loop optimum."""
import numpy as np
import pytest
import scipy.sparse as sp

from tests.helpers import BELIEF_FIXTURES, load_fixture
from oracle.belief_mpc import BeliefModelOracle, BeliefMPCOracle


@pytest.mark.parametrize("name", BELIEF_FIXTURES)
def test_belief_oracle_matches_reference(name):
    g = load_fixture(name)
    N, M, m = int(g["meta_N"]), int(g["meta_M"]), int(g["meta_m"])
    mpc = BeliefMPCOracle(BeliefModelOracle(M, m, float(g["meta_dt"])), N, float(g["meta_ydes"]), float(g["meta_vdes"]))
    for k in range(int(g["meta_steps"])):
        pre = "s%d_" % k
        u = mpc.solve(g[pre + "x0"], g[pre + "b0"], g[pre + "xbackup"], g[pre + "xref"][:4])
        np.testing.assert_allclose(np.array(mpc.A), g[pre + "A"], atol=1e-12)
        np.testing.assert_allclose(np.array(mpc.B), g[pre + "B"], atol=1e-12)
        np.testing.assert_allclose(np.array(mpc.C), g[pre + "C"], atol=1e-11)
        np.testing.assert_allclose(np.array(mpc.h0), g[pre + "h0"], atol=1e-11)
        np.testing.assert_allclose(np.array(mpc.Jh), g[pre + "Jh"], atol=1e-12)
        P, q, A, lo, hi = mpc.qp
        Pg = sp.coo_matrix((g[pre + "P_v"], (g[pre + "P_r"], g[pre + "P_c"])), shape=tuple(g[pre + "P_shape"])).tocsc()
        Ag = sp.coo_matrix((g[pre + "A_v"], (g[pre + "A_r"], g[pre + "A_c"])), shape=tuple(g[pre + "A_shape"])).tocsc()
        assert A.shape == Ag.shape and P.shape == Pg.shape
        assert abs(sp.triu(P) - Pg).max() < 1e-11 and abs(A - Ag).max() < 1e-11
        np.testing.assert_allclose(q, g[pre + "q"], atol=1e-11)
        np.testing.assert_allclose(hi, g[pre + "u"], atol=1e-11)
        fin = np.isfinite(lo)
        assert np.array_equal(fin, np.isfinite(g[pre + "l"]))
        np.testing.assert_allclose(lo[fin], g[pre + "l"][fin], atol=1e-11)
        assert mpc.feasible == 1
        np.testing.assert_allclose(mpc.uPred, g[pre + "uPred"], atol=1e-7)
        np.testing.assert_allclose(mpc.xPred, g[pre + "xPred"], atol=1e-7)
        assert abs(mpc.objective - float(g[pre + "objective"])) <= 1e-9 * abs(float(g[pre + "objective"]))
