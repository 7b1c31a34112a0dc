"""Merge scenario (row f4: PredictiveModel_merge, Highway_env_merge, BranchMPC_CVaR.solve(x, z, xRef, S, Fx=None, bx)) on
the B200 through the C ABI, against what the UNMODIFIED reference produced (tests/golden/make_golden.py `merge`,
`merge_models`): the model's point functions incl. the lookup-table ramp policies, all 60 controller calls of `sim_merge()`
(first input within 1e-3, objective within 1e-4 relative), and the drop-in classes driving the drop-in environment."""
import numpy as np
import pytest

from tests.helpers import (TOL_OBJ, TOL_U0, check_merge_fixture, check_merge_model_functions, load_fixture,
                           merge_fixture_config)
from _bmpc import batch, scenarios

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


def test_merge_model_functions_match_reference():
    check_merge_model_functions(batch.BatchedBranchMPC)


def test_merge_fixture_replay():
    g = load_fixture("highway_merge_default")
    mpc = batch.BatchedBranchMPC(merge_fixture_config(g))
    assert (mpc.totalu, mpc.totalx, mpc.nbranch) == (81, 83, 3)
    check_merge_fixture(lambda x, z, r, S, bd: mpc.solve_transformed_host_views(x, z, r, S, bd), _restore(mpc), g)
    mpc.close()


def _sim_merge_objects():
    """The statement sequence of main_branch.sim_merge (:53-87) on the drop-in modules."""
    import Init_MPC
    import MPC_branch
    import Highway_env_branch as henv
    from highway_branch_dyn import PredictiveModel_merge, backup_brake, backup_maintain_trackV, interpolant
    from utils import Branch_constants
    N, NB, N_lane, am, rm = 40, 1, 2, 7.0, 0.3
    cons = Branch_constants(s1=2, s2=3, c2=0.5, tran_diag=0.3, alpha=1, R=1.2, am=am, rm=rm, J_c=20, s_c=1, ylb=0., yub=7.2, L=4,
                            W=2.5, col_alpha=5, Kpsi=0.1)
    X1, X2, Y1, Y2, P1, P2 = henv.merge_geometry(N_lane, 1, 50, 300, 0)
    refY = interpolant('refY', 'linear', [np.append(X1, X2)], np.append(Y1, Y2))
    refpsi = interpolant('refpsi', 'linear', [np.append(X1, X2)], np.append(P1, P2))
    v0 = henv.v0
    ramp = [lambda x: backup_maintain_trackV(x, cons, v0, refpsi), lambda x: backup_brake(x, cons, refpsi)]
    normal = [lambda x: backup_maintain_trackV(x, cons, v0), lambda x: backup_brake(x, cons)]
    pred_model = [PredictiveModel_merge(4, 2, N, normal, 0.1, cons, (refY, refpsi), laneID=0, N_lane1=N_lane, N_lane2=1),
                  PredictiveModel_merge(4, 2, N, ramp, 0.1, cons, (refY, refpsi), laneID=1, N_lane1=N_lane, N_lane2=1)]
    par = Init_MPC.initBranchMPC(4, 2, N, NB, np.array([0.5, 1.8, 15, 0]), am, rm, N_lane, cons.W)
    mpc = MPC_branch.BranchMPC_CVaR(par, pred_model[0], ralpha=0.1)
    return henv, mpc, pred_model, N_lane


def test_dropin_sim_merge_closed_loop():
    """60 control periods of Highway_env_merge around the device controller: the geometry tables and the first call equal the
    reference's, the ego leaves the ramp without a collision and settles in the highway's lanes."""
    g = load_fixture("highway_merge_default")
    henv, mpc, pred_model, N_lane = _sim_merge_objects()
    env = henv.Highway_env_merge(2, N_lane, mpc, pred_model, 1, 50, 300, 0, pred_model[0].dt)
    np.testing.assert_allclose(env.merge_lane_ref_X, g["table_X"], atol=1e-12)
    np.testing.assert_allclose(env.merge_lane_ref_Y, g["table_Y"], atol=1e-12)
    np.testing.assert_allclose(env.merge_lane_ref_psi, g["table_psi"], atol=1e-12)
    state_rec, input_rec, _, choice, xPred_rec, zPred_rec, w_rec, collision = henv.Highway_sim(env, 6)
    assert state_rec.shape == (2, 60, 4) and np.isfinite(state_rec).all()
    np.testing.assert_allclose(input_rec[0, 0], g["input_rec"][0, 0], atol=TOL_U0)
    np.testing.assert_allclose(state_rec[:, 0], g["state_rec"][:, 0], atol=1e-4)
    np.testing.assert_allclose(input_rec[1], g["input_rec"][1], atol=1e-9)       # the obstacle's policy does not depend on the ego
    assert mpc.timeStep == 60 and mpc.feasible == 1 and not collision
    assert env.laneID == [0, 0]
    assert (np.abs(input_rec[0, :, 0]) <= 7.0 + 1e-9).all() and (np.abs(input_rec[0, :, 1]) <= 0.3 + 1e-9).all()
    assert 0.75 < state_rec[0, -1, 1] < 7.2 and abs(state_rec[0, -1, 3]) < 0.3
    # the ramp model's obstacle prediction (lookup-table policies) as the reference's environment recorded it at t = 0
    zp = pred_model[1].zpred_eval(g["x_init"][0])
    np.testing.assert_allclose(zp, g["backup_rec_ego"][0], atol=1e-9)


def test_merge_batch_matches_single_solves():
    """Per-episode S and bounds in one batched call: 48 perturbed copies of recorded ramp / highway scenes give what the same
    scenes give one at a time, all solved."""
    g = load_fixture("highway_merge_default")
    rng = np.random.default_rng(3)
    picks = [0, 5, 12, 17, 19, 30, 44, 59]
    X, Z, R, S, BD = [], [], [], [], []
    for k in picks:
        pre = "s%d_" % k
        for _ in range(6):
            X.append(g[pre + "x0"] + rng.normal(0, [0.5, 0.1, 0.5, 0.01]))
            Z.append(g[pre + "z0"] + rng.normal(0, [1.0, 0.1, 0.5, 0.0]))
            R.append(g[pre + "xref"]); S.append(g[pre + "S"]); BD.append(scenarios.bounds_from_bx(g[pre + "bx"])[0])
    X, Z, R, S, BD = (np.array(a) for a in (X, Z, R, S, BD))
    B = len(X)
    mpc = batch.BatchedBranchMPC(merge_fixture_config(g, batch_capacity=B))
    r = {k: np.array(v) for k, v in mpc.solve_transformed_host_views(X, Z, R, S, BD).items()}
    assert (r["status"] <= 1).all(), np.bincount(r["status"])
    one = batch.BatchedBranchMPC(merge_fixture_config(g))
    for i in range(0, B, 5):
        one.reset()
        q = one.solve_transformed_host_views(X[i], Z[i], R[i], S[i][None], BD[i][None])
        assert np.abs(q["u0"][0] - r["u0"][i]).max() < TOL_U0, i
        assert abs(q["objective"][0] - r["objective"][i]) <= TOL_OBJ * abs(r["objective"][i]), i
    # identity transform and the handle's own bounds through NULL pointers = the explicit ones
    hw = [i for i in range(B) if np.array_equal(S[i], np.eye(4))]
    sub = batch.BatchedBranchMPC(merge_fixture_config(g, batch_capacity=len(hw)))
    q = sub.solve_transformed_host_views(X[hw], Z[hw], R[hw], None, None)
    assert np.abs(q["u0"] - r["u0"][hw]).max() < TOL_U0
    for m_ in (mpc, one, sub):
        m_.close()


def test_device_merge_environment_matches_the_host_stepped_one():
    """bmpc_env_step_merge (lane id, ramp coordinates from the lookup tables, solve, plants on the device) against the
    drop-in Highway_env_merge, which does the same bookkeeping on the host around the same controller: 30 control periods
    from the reference's initial state give the same closed loop; a perturbed batch runs next to it."""
    from _bmpc import env as benv
    henv, mpc, pred_model, N_lane = _sim_merge_objects()
    host = henv.Highway_env_merge(2, N_lane, mpc, pred_model, 1, 50, 300, 0, pred_model[0].dt)
    x0 = np.array([v.state for v in host.veh_set])
    rng = np.random.default_rng(5)
    B = 64
    X = np.tile(x0[0], (B, 1)); Z = np.tile(x0[1], (B, 1))
    X[2:] += rng.normal(0, [1.0, 0.1, 0.5, 0.01], (B - 2, 4))
    Z[2:] += rng.normal(0, [2.0, 0.05, 0.5, 0.0], (B - 2, 4))
    dev_mpc = batch.BatchedBranchMPC(scenarios.merge_config(batch_capacity=B))
    dev = benv.BatchedMergeEnv(dev_mpc, X, Z, host.merge_lane_ref_X, host.merge_lane_ref_Y, host.merge_lane_ref_psi,
                               N_lane=N_lane, merge_lane=1, merge_s=50.0, v0=float(henv.v0))
    for t in range(30):
        u_set, x_set, *_ = host.step(t)
        out = dev.step(outputs=("u0", "status"))
        h = dev.host()
        assert (out["status"].cpu().numpy() <= 1).all(), t
        # the cutting-plane loop determines the first input to ~1e-4; the two loops see inputs that differ in the
        # last bits (numpy vs device tan / table lookup), so they agree within the parity bar, not bit for bit
        np.testing.assert_allclose(out["u0"][0].cpu().numpy(), u_set[0], atol=2 * TOL_U0, err_msg="step %d" % t)
        np.testing.assert_allclose(h["x"][0], x_set[0], atol=2e-3)
        np.testing.assert_allclose(h["z"][0], x_set[1], atol=1e-9)
        np.testing.assert_allclose(h["x"][1], h["x"][0], atol=0)          # identical episodes stay identical
        assert h["lane_id"][0] == host.laneID[0]
    assert host.laneID[0] == 0 and not h["collided"][:2].any()
    assert np.isfinite(h["x"]).all() and (h["lane_id"] == 0).sum() > B // 2
    dev_mpc.close()
