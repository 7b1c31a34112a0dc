"""Oracle of the belief-state model functions (rows H1, H2 of SURVEY.md 8a) vs values produced by the unmodified
reference module (tests/golden/hmm_functions.npz)."""
import numpy as np

from tests.helpers import load_fixture
from oracle import hmm

CONS = dict(L=4.0, W=2.5, ylb=0.0, yub=7.2, col_alpha=5.0)


def test_backup_rollout_matches_reference():
    g = load_fixture("hmm_functions")
    for k in range(g["X0"].shape[0]):
        xb = hmm.backup_rollout(g["X0"][k], [hmm.MAINTAIN, hmm.BRAKE], int(g["N"]), float(g["dt"]), 0.1)
        np.testing.assert_allclose(xb, g["XB"][k], atol=1e-12)


def test_sensitivity_rollout_matches_reference():
    g = load_fixture("hmm_functions")
    idx = 0
    for k in range(g["X0"].shape[0]):
        for kind in (hmm.MAINTAIN, hmm.BRAKE):
            xx, QQ, Qt = hmm.rollout_sensitivity(g["X0"][k, 0], kind, int(g["sens_steps"]), float(g["sens_ts"]), g["sens_f0"], 0.1)
            np.testing.assert_allclose(xx, g["sens_x"][idx], atol=1e-12)
            np.testing.assert_allclose(QQ, g["sens_Q"][idx], atol=1e-9)
            np.testing.assert_allclose(Qt, g["sens_Qt"][idx], atol=1e-12)
            idx += 1


def test_belief_transition_matches_reference():
    g = load_fixture("hmm_functions")
    K, M, m = g["B"].shape
    N, t = int(g["N"]), int(g["t_index"])
    for k in range(K):
        for i in range(M):
            h = np.array([hmm.safety(g["EGO"][k], g["XB"][k, m * i + j].reshape(N, 4, order="F")[t], **CONS) for j in range(m)])
            np.testing.assert_allclose(h, g["h"][k, i], atol=1e-12)
            H = hmm.backup_trans(h, 2.0, 0.3)
            np.testing.assert_allclose(H, g["H"][k, i], atol=1e-12)
            np.testing.assert_allclose(H.sum(axis=1), 1.0, atol=1e-12)
            np.testing.assert_allclose(hmm.belief_step(g["B"][k, i], H), g["b_next"][k, i], atol=1e-12)
            np.testing.assert_allclose(hmm.belief_step(g["B"][k, i], H, g["CBF"][k, i]), g["b_next_env"][k, i], atol=1e-12)
