"""The reference-facing Python interface on the GPU: the setup lines of main_branch.py (:24-46) run against the drop-in
modules, then the closed loop of the golden fixture is repeated through BranchMPC.solve()."""
import numpy as np
import pytest

from tests.helpers import TOL_OBJ, TOL_U0, load_fixture

pytestmark = pytest.mark.gpu


def _main_branch_setup(lc_target=(0.5, 1.8, 15, 0)):
    # --- main_branch.py:24-46, with BranchMPC instead of BranchMPC_CVaR (:46 vs :48) ---
    from Init_MPC import initBranchMPC
    from MPC_branch import BranchMPC
    from highway_branch_dyn import PredictiveModel, backup_brake, backup_lc, backup_maintain
    from utils import Branch_constants
    N = 8
    n = 4; d = 2
    am = 6.0
    rm = 0.3
    dt = 0.1
    NB = 2
    N_lane = 4
    xRef = np.array(lc_target, dtype=float)
    cons = Branch_constants(s1=2, s2=3, c2=0.5, tran_diag=0.3, alpha=1, R=1.2, am=am, rm=rm, J_c=20, s_c=1, ylb=0., yub=7.2,
                            L=4, W=2.5, col_alpha=5, Kpsi=0.1)
    backupcons = [lambda x: backup_maintain(x, cons), lambda x: backup_brake(x, cons), lambda x: backup_lc(x, xRef)]
    model = PredictiveModel(n, d, N, backupcons, dt, cons)
    mpcParam = initBranchMPC(n, d, N, NB, xRef, am, rm, N_lane, cons.W)
    mpc = BranchMPC(mpcParam, model)
    return mpc, model, cons, backupcons


@pytest.mark.parametrize("name", ["highway_branch_default", "highway_branch_close"])
def test_solve_interface_reproduces_reference_closed_loop(name):
    g = load_fixture(name)
    mpc, model, cons, _ = _main_branch_setup(tuple(g["meta_lc_target"]))
    for k in range(int(g["meta_steps"])):
        pre = "s%d_" % k
        mpc.solve(g[pre + "x0"], g[pre + "z0"], g[pre + "xref"])
        assert mpc.feasible == 1 and mpc.timeStep == k + 1
        assert mpc.uPred.shape == (97, 2) and mpc.xPred.shape == (106, 4) and mpc.uLin.shape == (98, 2)
        assert np.abs(mpc.uPred[0] - g[pre + "uPred"][0]).max() < TOL_U0
        np.testing.assert_allclose(mpc.uPred, g[pre + "uPred"], atol=1e-6)
        np.testing.assert_allclose(mpc.xPred, g[pre + "xPred"], atol=1e-6)
        assert abs(mpc.objective - float(g[pre + "objective"])) <= TOL_OBJ * abs(float(g[pre + "objective"]))
        assert np.array_equal(mpc.OldInput, mpc.uPred[0]) and np.array_equal(mpc.uLin[-1], mpc.uPred[-1])
        xs, zs, us, ws = mpc.BT2array()
        assert len(xs) == 12 and xs[0].shape == (9, 4) and zs[0].shape == (9, 4) and us[0].shape == (9, 2)
        np.testing.assert_allclose(ws, g[pre + "w"][1:], atol=1e-6)
        np.testing.assert_allclose(zs[0][1:], g[pre + "zbar"][1:9], atol=1e-9)
    assert mpc.ndx[5] == 34 and mpc.ndu[5] == 33 and mpc.totalx == 106 and mpc.totalu == 97


def test_model_methods_on_device_match_reference():
    g = load_fixture("model_functions")
    mpc, model, cons, backupcons = _main_branch_setup()
    from highway_branch_dyn import backup_lc
    tgt = g["hw_lc_target"]
    model.update_backup(backupcons[:2] + [lambda x: backup_lc(x, tgt)])
    for k in range(6):
        A, B, C, xp = model.dyn_linearization(g["hw_X"][k], g["hw_U"][k])
        np.testing.assert_allclose(A, g["hw_A"][k], atol=1e-12)
        np.testing.assert_allclose(B, g["hw_B"][k], atol=1e-12)
        np.testing.assert_allclose(C, g["hw_C"][k], atol=1e-12)
        np.testing.assert_allclose(xp, g["hw_xp"][k], atol=1e-12)
        np.testing.assert_allclose(model.zpred_eval(g["hw_Z"][k]), g["hw_zpred"][k], atol=1e-11)
        p, dp = model.branch_eval(g["hw_X"][k], g["hw_Z"][k])
        np.testing.assert_allclose(p, g["hw_p"][k], atol=1e-11)
        np.testing.assert_allclose(dp, g["hw_dp"][k], atol=1e-6)
        h, dh = model.col_eval(g["hw_X"][k], g["hw_Z"][k])
        np.testing.assert_allclose(h, g["hw_hlin"][k], atol=1e-10)
        np.testing.assert_allclose(dh, g["hw_dh"][k], atol=1e-11)


def test_batched_solve_through_the_same_interface():
    from _bmpc import scenarios
    mpc, model, cons, _ = _main_branch_setup()
    B = 256
    x0, z0, xref, pp = scenarios.highway_batch(B, seed=11)
    u = mpc.solve(x0, z0, xref)
    assert u.shape == (B, 2) and mpc.uPred.shape == (B, 97, 2) and mpc.xPred.shape == (B, 106, 4)
    assert mpc.feasible.shape == (B,) and mpc.feasible.all()
    one, _, _, _ = _main_branch_setup()
    for i in (0, 17, 255):
        ui = one.solve(x0[i], z0[i], xref[i])
        assert np.array_equal(ui, u[i])
        one.reset()


def test_highway_env_dropin_reproduces_reference_sim():
    """main_branch.sim_overtake's loop (Highway_env_branch.Highway_sim) through the drop-in environment: the recorded
    trace of the reference's own environment is reproduced state by state."""
    import Highway_env_branch as henv
    g = load_fixture("highway_env_default")
    mpc, model, cons, _ = _main_branch_setup()
    env = henv.Highway_env(NV=2, mpc=mpc, N_lane=4)
    steps = 12
    state_rec, input_rec, _, choice_rec, xPred_rec, zPred_rec, w_rec, collision = henv.Highway_sim(env, steps * env.dt)
    assert state_rec.shape == (2, steps, 4) and input_rec.shape == (2, steps, 2)
    np.testing.assert_allclose(state_rec[0], g["x"][:steps], atol=1e-5)
    np.testing.assert_allclose(state_rec[1], g["z"][:steps], atol=1e-8)
    np.testing.assert_allclose(input_rec[1], g["u_obs"][:steps], atol=1e-9)
    assert [int(c) for c in choice_rec[1]] == [int(c) for c in g["backupidx"][:steps]]
    assert collision == bool(g["collision"][steps - 1])
    assert len(xPred_rec[0]) == 12 and xPred_rec[0][0].shape == (9, 4) and len(w_rec[0]) == 12
    assert mpc.timeStep == steps and mpc.uPred.shape == (97, 2)


def test_highway_env_dropin_batched():
    import Highway_env_branch as henv
    from _bmpc import scenarios
    mpc, model, cons, _ = _main_branch_setup()
    B = 256
    x0, z0, _, _ = scenarios.highway_batch(B, seed=5)
    env = henv.Highway_env(NV=2, mpc=mpc, N_lane=4, x0=np.stack([x0, z0], axis=1))
    state_rec, input_rec, *_rest, collision = henv.Highway_sim(env, 0.5)
    assert state_rec.shape == (B, 2, 5, 4) and collision.shape == (B,)
    assert np.isfinite(state_rec).all() and (np.abs(input_rec[:, 0, :, 0]) <= 6 + 1e-9).all()
    assert mpc.uPred.shape == (B, 97, 2) and (np.asarray(mpc.feasible) == 1).all()


def test_quadruped_env_dropin_runs():
    from Init_MPC import initquadBranchMPC
    from MPC_branch import BranchMPCProx
    from quadruped_branch_dyn import PredictiveModel, backup_forward, backup_stop
    from utils import Quad_constants
    import quadruped_env as qenv
    cons = Quad_constants(s1=2, s2=3, c2=0.5, alpha=1, R=1.2, vxm=0.2, vym=0.1, rm=0.5, L1=0.5, W1=0.3, L2=1, W2=0.6,
                          col_tol=0.2, col_alpha=5)
    v0 = 0.2
    model = PredictiveModel(3, 3, 25, [lambda x: backup_forward(x, v0), lambda x: backup_stop(x)], 0.2, cons)
    x_des = np.array([5., -3., 0.])
    mpc = BranchMPCProx(initquadBranchMPC(3, 3, 25, 2, x_des, 0.2, 0.1, 0.5), model)
    env = qenv.Quad_env(2, mpc, x_des)
    state_rec, input_rec, xPred_rec, zPred_rec = qenv.Robot_sim(env, 0.6)      # main_quadruped.py crashes here in the reference
    assert state_rec.shape == (2, 3, 3) and np.isfinite(state_rec).all()
    assert (input_rec[0, :, 0] <= 0.2 + 1e-9).all() and (input_rec[0, :, 0] >= -1e-9).all()
    assert mpc.uPred.shape == (151, 3) and mpc.feasible == 1
