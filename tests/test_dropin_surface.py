"""The drop-in modules expose the reference's names and signatures (checked without a GPU; compute needs one)."""
import inspect

import numpy as np
import pytest

from tests.helpers import PKG  # noqa: F401  (puts the package directory on sys.path, as a user would)


def _highway():
    import highway_branch_dyn as hw
    from utils import Branch_constants
    cons = Branch_constants(s1=2, s2=3, c2=0.5, tran_diag=0.3, alpha=1, R=1.2, am=6.0, rm=0.3, J_c=20, s_c=1, ylb=0.,
                            yub=7.2, L=4, W=2.5, col_alpha=5, Kpsi=0.1)
    xRef = np.array([0.5, 1.8, 15, 0])
    backupcons = [lambda x: hw.backup_maintain(x, cons), lambda x: hw.backup_brake(x, cons), lambda x: hw.backup_lc(x, xRef)]
    return hw, cons, xRef, backupcons


def test_reference_module_names_and_signatures():
    import Init_MPC
    import MPC_branch
    import highway_branch_dyn as hw
    import quadruped_branch_dyn as qd
    import utils
    for name in ("BranchMPC", "BranchMPCProx", "robustMPC", "BranchMPC_CVaR", "BranchMPCParams", "BranchTree"):
        assert hasattr(MPC_branch, name)
    assert list(inspect.signature(MPC_branch.BranchMPC.solve).parameters)[:4] == ["self", "x", "z", "xRef"]
    assert list(inspect.signature(hw.PredictiveModel.__init__).parameters) == ["self", "n", "d", "N", "backupcons", "dt",
                                                                               "cons", "N_lane"]
    assert list(inspect.signature(qd.PredictiveModel.__init__).parameters) == ["self", "n", "d", "N", "backupcons", "dt", "cons"]
    assert list(inspect.signature(Init_MPC.initBranchMPC).parameters) == ["n", "d", "N", "NB", "xRef", "am", "rm", "N_lane", "W"]
    assert list(inspect.signature(Init_MPC.initquadBranchMPC).parameters) == ["n", "d", "N", "NB", "xRef", "vxm", "vym", "rm"]
    for name in ("dyn_linearization", "branch_eval", "zpred_eval", "xpred_eval", "col_eval", "update_backup"):
        assert hasattr(hw.PredictiveModel, name) and hasattr(qd.PredictiveModel, name)
    for name in ("Branch_constants", "Quad_constants", "MPCParams", "PythonMsg"):
        assert hasattr(utils, name)


def test_policy_closures_become_a_policy_table():
    from _bmpc import abi
    hw, cons, xRef, backupcons = _highway()
    model = hw.PredictiveModel(4, 2, 8, backupcons, 0.1, cons)
    assert model.m == 3 and model.LB == [1.25, 3 * 3.6 - 1.25]
    assert [d.kind for d in model.descriptors] == [abi.POLICY_MAINTAIN, abi.POLICY_BRAKE, abi.POLICY_LC]
    assert model.policy_params()[2].tolist() == [0.5, 1.8, 15.0, 0.0]
    # update_backup with a new lane-change target is a parameter change (Highway_env_branch.py:117-118)
    tgt = np.array([0., 5.4, 17., 0.])
    model.update_backup(backupcons[:2] + [lambda x: hw.backup_lc(x, tgt)])
    assert model.policy_params()[2].tolist() == [0.0, 5.4, 17.0, 0.0]
    with pytest.raises(TypeError):
        hw.PredictiveModel(4, 2, 8, [lambda x: np.array([0., 0.])], 0.1, cons)      # not a library policy


def test_numeric_policy_branches_match_the_reference_formulas():
    hw, cons, xRef, backupcons = _highway()
    x = np.array([3.0, 2.0, 18.0, 0.05])
    assert np.allclose(backupcons[0](x), [0.0, -0.005])
    a = (-5 * np.exp(-15) + -18 * np.exp(-54)) / (np.exp(-15) + np.exp(-54))          # softmax([-5,-v],3), numeric branch
    assert np.allclose(backupcons[1](x), [a, -0.005])
    assert np.allclose(backupcons[2](x), [-0.8558 * 3.0, -0.3162 * 0.2 - 3.9889 * 0.05])
    assert hw.veh_col(np.array([0., 0.]), np.array([30., 0.]), [5, 2.7]) == pytest.approx(
        (5 * np.exp(5) - 2.7 * np.exp(-2.7)) / (np.exp(5) + np.exp(-2.7)))                # +-5 clip of the numeric branch
    assert hw.lane_bdry_h(np.array([0., 3.6, 0., 0.]), 0, 7.2) == pytest.approx(3.6)


def test_parameter_objects_keep_the_reference_quirks():
    import Init_MPC
    from MPC_branch import BranchMPCParams
    p = Init_MPC.initBranchMPC(4, 2, 8, 2, np.array([0.5, 1.8, 15, 0]), 6.0, 0.3, 4, 2.5)
    assert isinstance(p.bx, tuple) and p.bx[0].shape == (4, 1)          # 1-tuple (Init_MPC.py:48-51)
    assert p.Qf is p.Q and np.array_equal(p.dR, np.zeros(2))
    with pytest.raises(TypeError):
        p.not_a_field = 1
    q = Init_MPC.initquadBranchMPC(3, 3, 25, 2, np.array([5., 5., 0.]), 0.2, 0.1, 0.5)
    assert q.Fx.shape == (0, 3) and q.dR.tolist() == [0.9, 5, 1]
    assert isinstance(BranchMPCParams(n=2, d=1, Q=np.eye(2)).xRef, np.ndarray)


def test_environment_modules_keep_the_reference_surface():
    """Highway_env_branch / quadruped_env drop-ins: names and signatures of the reference (no GPU needed to import)."""
    import inspect
    import Highway_env_branch as henv
    import quadruped_env as qenv
    assert henv.v0 == 20 and henv.lane_width == 3.6
    assert list(inspect.signature(henv.Highway_env.__init__).parameters)[:4] == ["self", "NV", "mpc", "N_lane"]
    assert list(inspect.signature(henv.Highway_env.step).parameters) == ["self", "t_"]
    assert list(inspect.signature(henv.Highway_sim).parameters) == ["env", "T"]
    assert list(inspect.signature(henv.sim_overtake).parameters)[:2] == ["mpc", "N_lane"]
    v = henv.vehicle([1, 2, 3, 0.1], dt=0.1)
    assert v.v_length == 4 and v.v_width == 2.4 and v.laneidx == 0 and v.backupidx == 0
    assert list(inspect.signature(qenv.Quad_env.__init__).parameters)[:4] == ["self", "NR", "mpc", "x_des"]
    assert list(inspect.signature(qenv.Robot_sim).parameters) == ["env", "T"]
