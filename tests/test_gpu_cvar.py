"""BranchMPC_CVaR on the B200 through the C ABI against the fixtures recorded from the UNMODIFIED reference controller
(tests/golden/make_golden.py `cvar`): first input within 1e-3, objective J within 1e-4 relative, for ralpha = 0.9 (what
main_branch.py:48 uses) and 0.1 (sim_merge, :92), on 3-4 closed-loop steps each."""
import numpy as np
import pytest

from tests.helpers import CVAR_FIXTURES, TOL_OBJ, TOL_U0, check_cvar_fixture, cvar_fixture_config, load_fixture
from _bmpc import batch

pytestmark = pytest.mark.gpu


def _restore(mpc):
    def set_state(uLin, pbest, old_input):
        st = mpc.get_state(1)
        st["uLin"][0, :len(uLin)] = uLin
        st["pbest"][0] = pbest
        st["old_input"][0] = old_input
        st["started"][0] = 1
        mpc.set_state(st)
    return set_state


@pytest.mark.parametrize("name", CVAR_FIXTURES)
def test_cvar_fixture_closed_loop(name):
    g = load_fixture(name)
    mpc = batch.BatchedBranchMPC(cvar_fixture_config(g))
    check_cvar_fixture(lambda x, z, r: mpc.solve_host(x, z, r), _restore(mpc), g)
    mpc.close()


def test_cvar_batch_properties_and_dropin():
    """A batch of random scenes through the drop-in class: every episode solved, inputs inside their box, J at least the
    risk-neutral tree QP's objective structure (J >= root cost), and the same episode solved alone gives the same answer."""
    import Init_MPC
    import MPC_branch
    from highway_branch_dyn import PredictiveModel, backup_brake, backup_lc, backup_maintain
    from utils import Branch_constants
    from _bmpc import scenarios
    cons = Branch_constants(s1=2, s2=3, c2=0.5, tran_diag=0.3, alpha=1, R=1.2, am=6.0, rm=0.3, J_c=20, s_c=1, ylb=0., yub=7.2,
                            L=4, W=2.5, col_alpha=5, Kpsi=0.1)
    xRef = np.array([0.5, 1.8, 15, 0])
    backupcons = [lambda x: backup_maintain(x, cons), lambda x: backup_brake(x, cons), lambda x: backup_lc(x, xRef)]
    model = PredictiveModel(4, 2, 8, backupcons, 0.1, cons)
    par = Init_MPC.initBranchMPC(4, 2, 8, 2, xRef, 6.0, 0.3, 4, cons.W)
    mpc = MPC_branch.BranchMPC_CVaR(par, model, ralpha=0.9)
    B = 256
    x0, z0, xref, _ = scenarios.highway_batch(B, seed=5)
    mpc.solve(x0, z0, xref)
    assert mpc.uPred.shape == (B, 97, 2) and mpc.xPred.shape == (B, 106, 4)
    assert (np.asarray(mpc.feasible) == 1).all(), np.bincount(mpc.status)
    assert (np.abs(mpc.uPred[:, :, 0]) <= 6.0 + 1e-9).all() and (np.abs(mpc.uPred[:, :, 1]) <= 0.3 + 1e-9).all()
    J = mpc.objective.copy()
    assert np.isfinite(J).all() and (J > 0).all()
    one = MPC_branch.BranchMPC_CVaR(par, model, ralpha=0.9)
    for i in (0, 17, 101):
        one.reset()
        one.solve(x0[i], z0[i], xref[i])
        assert np.abs(one.uPred[0] - mpc.uPred[i, 0]).max() < TOL_U0
        assert abs(one.objective - J[i]) <= TOL_OBJ * abs(J[i])
    # risk aversion: a smaller ralpha can only raise the nested-CVaR objective of the same scene
    averse = MPC_branch.BranchMPC_CVaR(par, model, ralpha=0.1)
    averse.solve(x0[:32], z0[:32], xref[:32])
    assert (averse.objective >= J[:32] * (1 - 1e-6)).all()
