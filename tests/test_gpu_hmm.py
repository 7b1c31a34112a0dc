"""Belief-state model kernels (rows H1/H2) vs values produced by the unmodified reference module."""
import numpy as np
import pytest

from tests.helpers import load_fixture

pytestmark = pytest.mark.gpu
PAR = dict(Kpsi=0.1, L=4.0, W=2.5, ylb=0.0, yub=7.2, col_alpha=5.0, s1=2.0, s2=3.0, c2=0.5, tran_diag=0.3)


def test_backup_rollout_kernel():
    from _bmpc import abi, hmm
    g = load_fixture("hmm_functions")
    xb = hmm.backup_rollout(g["X0"], [abi.HMM_MAINTAIN, abi.HMM_BRAKE], int(g["N"]), float(g["dt"]), 0.1)
    np.testing.assert_allclose(xb, g["XB"], atol=1e-11)


def test_sensitivity_rollout_kernel():
    from _bmpc import abi, hmm
    g = load_fixture("hmm_functions")
    xx, QQ, Qt = hmm.rollout_sensitivity(g["X0"][:, 0], [abi.HMM_MAINTAIN, abi.HMM_BRAKE], int(g["sens_steps"]),
                                         float(g["sens_ts"]), g["sens_f0"], 0.1)
    K = g["X0"].shape[0]
    np.testing.assert_allclose(xx.reshape(2 * K, -1, 4), g["sens_x"], atol=1e-11)
    np.testing.assert_allclose(QQ.reshape(2 * K, -1, 4, 4), g["sens_Q"], atol=1e-8)
    np.testing.assert_allclose(Qt.reshape(2 * K, -1, 4), g["sens_Qt"], atol=1e-11)


def test_belief_update_kernel():
    from _bmpc import hmm
    g = load_fixture("hmm_functions")
    K, M, m = g["B"].shape
    N, t = int(g["N"]), int(g["t_index"])
    xb = g["XB"].reshape(K, M, m, 4, N)[..., t]          # component-major rows -> state at step t
    h, H, bn = hmm.belief_update(g["EGO"], xb, g["B"], PAR)
    np.testing.assert_allclose(h, g["h"], atol=1e-11)
    np.testing.assert_allclose(H, g["H"], atol=1e-11)
    np.testing.assert_allclose(bn, g["b_next"], atol=1e-11)
    _, _, be = hmm.belief_update(g["EGO"], xb, g["B"], PAR, cbf=g["CBF"])
    np.testing.assert_allclose(be, g["b_next_env"], atol=1e-11)
    assert np.abs(be.sum(axis=-1) - 1.0).max() < 1e-12
