"""Shared test helpers: golden fixtures, bmpc_config factories, oracle drivers."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "belief-planning_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

from _bmpc import scenarios  # noqa: E402
from oracle import params  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
HIGHWAY_FIXTURES = ["highway_branch_default", "highway_branch_close", "highway_branch_m2_nb3", "highway_branch_m3_nb1"]

# parity bars of BASELINE.json's north_star
TOL_U0 = 1e-3        # first applied control, absolute
TOL_OBJ = 1e-4       # objective, relative
TOL_VIOL = 1e-5      # constraint violation, absolute


def load_fixture(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def fixture_config(g, **kw):
    return scenarios.highway_config(policies=[str(p) for p in g["meta_policies"]], NB=int(g["meta_NB"]),
                                    N=int(g["meta_N"]), lc_target=tuple(g["meta_lc_target"]), **kw)


def quadruped_fixture_config(g, **kw):
    return scenarios.quadruped_config(NB=int(g["meta_NB"]), N=int(g["meta_N"]), v0=float(g["meta_v0"]), **kw)


def robust_fixture_config(g, **kw):
    from _bmpc import abi
    cfg = fixture_config(g, **kw)
    cfg.controller = abi.CTRL_ROBUST
    return cfg


def check_robust_fixture(solve, g, tol=2e-5):
    """robustMPC fixture: chain QP optimum of every recorded closed-loop step (no tree outputs)."""
    for k in range(int(g["meta_steps"])):
        pre = "s%d_" % k
        r = solve(g[pre + "x0"], g[pre + "z0"], g[pre + "xref"])
        # the slots hold near-duplicate collision rows (policies that barely differ early on); the interior-point fallback
        # settles such degenerate active sets, so the solve must end solved (polished or converged), never on the cap
        assert r["status"][0] in (0, 1)
        assert r["uPred"][0].shape == g[pre + "uPred"].shape and r["xPred"][0].shape == g[pre + "xPred"].shape
        np.testing.assert_allclose(r["uPred"][0], g[pre + "uPred"], atol=tol)
        np.testing.assert_allclose(r["xPred"][0], g[pre + "xPred"], atol=tol)
        assert np.abs(r["u0"][0] - g[pre + "uPred"][0]).max() < TOL_U0
        obj = float(g[pre + "objective"])
        assert abs(r["objective"][0] - obj) <= TOL_OBJ * abs(obj)


def check_fixture_closed_loop(solve, g, tol=1e-6):
    """`solve(x, z, xref) -> result dict` is called for every recorded step of a fixture; everything the reference
    produced for that step (tree data, linearisation trajectory, optimum) must be reproduced."""
    for k in range(int(g["meta_steps"])):
        pre = "s%d_" % k
        r = solve(g[pre + "x0"], g[pre + "z0"], g[pre + "xref"])
        assert r["status"][0] in (0, 1), "step %d status %d" % (k, r["status"][0])
        np.testing.assert_allclose(r["branch_w"][0], g[pre + "w"], atol=tol)
        p_ref = g[pre + "p"]
        mask = np.isfinite(p_ref)
        np.testing.assert_allclose(r["branch_p"][0][mask], p_ref[mask], atol=tol)
        assert np.isnan(r["branch_p"][0][~mask]).all()
        np.testing.assert_allclose(r["xLin"][0], g[pre + "xbar"], atol=tol)      # depends on the previous optimum
        np.testing.assert_allclose(r["zPred"][0], g[pre + "zbar"], atol=1e-9)
        np.testing.assert_allclose(r["uPred"][0], g[pre + "uPred"], atol=tol)
        np.testing.assert_allclose(r["xPred"][0], g[pre + "xPred"], atol=tol)
        assert np.abs(r["u0"][0] - g[pre + "uPred"][0]).max() < TOL_U0
        obj = float(g[pre + "objective"])
        assert abs(r["objective"][0] - obj) <= TOL_OBJ * abs(obj)


def oracle_episode(x, z, xref, lc_target, steps):
    """Closed loop of the oracle controller; returns per-step (x, z, u0, objective, uPred)."""
    mpc = params.highway_branch_mpc(lc_target=lc_target)
    out = []
    for _ in range(steps):
        u = mpc.solve(x, z, xref).copy()
        out.append((x.copy(), z.copy(), u, mpc.objective, mpc.uPred.copy()))
        x = scenarios.euler_highway(x[None], u[None])[0]
        z = scenarios.euler_highway(z[None], np.array([[0.0, -0.1 * z[3]]]))[0]
    return out


SWEEP = [(m, NB) for m in (2, 3, 4) for NB in (1, 2, 3)]          # BASELINE config 5: branching factor x depth
POLICY_NAMES = ["maintain", "brake", "lc", "trackv"]


def sweep_case(m, NB, count=2, seed=21):
    """(config kwargs, oracle policy descriptors, x0, z0, xref, policy params) of one tree shape of the sweep."""
    x0, z0, xref, pp3 = scenarios.highway_batch(count, seed=seed + 10 * m + NB)
    pp = np.zeros((count, m, 4))
    if m >= 3:
        pp[:, 2, :] = pp3[:, 2, :]
    if m >= 4:
        pp[:, 3, 0] = 20.0
    names = POLICY_NAMES[:m]
    desc = [("trackv", 20.0) if p == "trackv" else p for p in names]
    return dict(policies=names, NB=NB), desc, x0, z0, xref, pp


def check_sweep_case(make_solver, m, NB):
    kw, desc, x0, z0, xref, pp = sweep_case(m, NB)
    solver = make_solver(scenarios.highway_config(batch_capacity=len(x0), **kw))
    r = solver(x0, z0, xref, pp)
    T = params.highway_branch_mpc(desc, NB=NB).topo
    assert r["uPred"].shape[1:] == (T.totalu, 2) and r["xPred"].shape[1:] == (T.totalx, 4)
    for i in range(len(x0)):
        ora = params.highway_branch_mpc(desc, NB=NB, lc_target=pp[i, 2] if m >= 3 else (0.5, 1.8, 15.0, 0.0))
        u = ora.solve(x0[i], z0[i], xref[i])
        assert ora.feasible == 1
        assert np.abs(r["u0"][i] - u).max() < TOL_U0, (m, NB, i)
        assert abs(r["objective"][i] - ora.objective) <= TOL_OBJ * abs(ora.objective), (m, NB, i)
        np.testing.assert_allclose(r["branch_w"][i], ora.w, atol=1e-9)


# ---- interior-point fallback -------------------------------------------------------------------------------------
QUAD_HARD = [528, 545, 1006, 1225]   # quadruped_batch(2048, seed=1238): degenerate active sets, ADMM + polish alone end on MAXITER


def force_interior_point(cfg):
    """Knobs that send every problem through the interior-point fallback (no warm polish, no polish attempt before it)."""
    cfg.warm_polish = -1
    cfg.reserved[4] = 100
    return cfg


def check_quadruped_hard_cases(make_solver):
    """Problems that defeat ADMM + active-set polish must come back from the interior-point fallback within the parity bars."""
    x0, z0, xref = scenarios.quadruped_batch(2048, seed=1238)
    idx = np.array(QUAD_HARD)
    solver = make_solver(scenarios.quadruped_config(batch_capacity=len(idx)))
    r = solver(x0[idx], z0[idx], xref[idx])
    assert (r["status"] <= 1).all(), r["status"]
    for n, i in enumerate(idx):
        ora = params.quadruped_prox_mpc()
        u = ora.solve(x0[i], z0[i], xref[i])
        assert np.abs(r["u0"][n] - u).max() < TOL_U0, (i, r["status"][n])
        assert abs(r["objective"][n] - ora.objective) <= TOL_OBJ * abs(ora.objective), (i, r["status"][n])


def check_forced_interior_point(make_solver, count=6):
    """Highway problems solved (almost) only by the interior point + its closing polish agree with the oracle."""
    x0, z0, xref, pp = scenarios.highway_batch(count, seed=77)
    solver = make_solver(force_interior_point(scenarios.highway_config(batch_capacity=count)))
    r = solver(x0, z0, xref, pp)
    assert (r["status"] <= 1).all(), r["status"]
    for i in range(count):
        ora = params.highway_branch_mpc(lc_target=pp[i, 2])
        u = ora.solve(x0[i], z0[i], xref[i])
        assert np.abs(r["u0"][i] - u).max() < TOL_U0, i
        assert abs(r["objective"][i] - ora.objective) <= TOL_OBJ * abs(ora.objective), i
    return r


# ---- BranchMPC_CVaR ------------------------------------------------------------------------------------------------
CVAR_FIXTURES = ["highway_cvar_default", "highway_cvar_close", "highway_cvar_alpha01", "highway_cvar_m2_nb1"]


def cvar_fixture_config(g, **kw):
    from _bmpc import abi
    cfg = fixture_config(g, **kw)
    cfg.controller = abi.CTRL_CVAR
    cfg.cvar_alpha = float(g["meta_ralpha"])
    return cfg


def check_cvar_fixture(solve, set_state, g):
    """Every recorded closed-loop step of the UNMODIFIED BranchMPC_CVaR (tests/golden/make_golden.py run_highway_cvar): the
    warm-start state the reference linearised about (uLin, arg-max children) is restored first, because the cone program
    leaves the branches without risk weight undetermined and the reference's next linearisation inherits whatever its
    solver returned there.  Compared: tree data, linearisation trajectory, first input (1e-3), objective J (1e-4 rel.) and the
    trajectories of the branches whose cones carry weight."""
    for k in range(int(g["meta_steps"])):
        pre = "s%d_" % k
        if k > 0:
            set_state(g[pre + "uLin_before"], g[pre + "pbest_before"], g["s%d_uPred" % (k - 1)][0])
        r = solve(g[pre + "x0"], g[pre + "z0"], g[pre + "xref"])
        assert r["status"][0] in (0, 1), "step %d status %d" % (k, r["status"][0])
        np.testing.assert_allclose(r["branch_w"][0], g[pre + "w"], atol=1e-6)
        np.testing.assert_allclose(r["xLin"][0], g[pre + "xbar"], atol=1e-6)
        np.testing.assert_allclose(r["zPred"][0], g[pre + "zbar"], atol=1e-9)
        u0 = g[pre + "uPred"][0]
        assert np.abs(r["u0"][0] - u0).max() < TOL_U0, (k, r["u0"][0], u0)
        obj = float(g[pre + "objective"])
        assert abs(r["objective"][0] - obj) <= TOL_OBJ * abs(obj), (k, r["objective"][0], obj)


# ---- merge scenario (sim_merge) ---------------------------------------------------------------------------------------
def merge_fixture_config(g, **kw):
    return scenarios.merge_config(N=int(g["meta_N"]), NB=int(g["meta_NB"]), v0=float(g["meta_v0"]), am=float(g["meta_am"]),
                                  rm=float(g["meta_rm"]), N_lane=int(g["meta_N_lane"]), ralpha=float(g["meta_ralpha"]), **kw)


def check_merge_fixture(solve_transformed, set_state, g, steps=None):
    """Replay of the UNMODIFIED `sim_merge()` (tests/golden/make_golden.py run_highway_merge): every
    `mpc.solve(x, z, xRef, S, Fx=None, bx=bx)` the reference's `Highway_env_merge.step` made in 6 s of closed loop, with the
    warm start that step linearised about restored first (see check_cvar_fixture).  `solve_transformed(x, z, xref, S (1,4,4),
    bounds (1,2,2))`."""
    steps = range(int(g["meta_steps"])) if steps is None else steps
    assert int(g["meta_NB"]) == 1
    Q = g["meta_Q"]

    def jc(v):
        return float(v @ Q @ v)

    for k in steps:
        pre = "s%d_" % k
        if k > 0:
            set_state(g[pre + "uLin_before"], g[pre + "pbest_before"], g["s%d_uPred" % (k - 1)][0])
        r = solve_transformed(g[pre + "x0"], g[pre + "z0"], g[pre + "xref"], g[pre + "S"][None],
                              scenarios.bounds_from_bx(g[pre + "bx"]))
        assert r["status"][0] in (0, 1), "step %d status %d" % (k, r["status"][0])
        np.testing.assert_allclose(r["branch_w"][0], g[pre + "w"], atol=1e-6)
        np.testing.assert_allclose(r["xLin"][0], g[pre + "xbar"], atol=1e-6)
        np.testing.assert_allclose(r["zPred"][0], g[pre + "zbar"], atol=1e-9)
        u0 = g[pre + "uPred"][0]
        assert np.abs(r["u0"][0] - u0).max() < TOL_U0, (k, r["u0"][0], u0)
        # the reference's cones keep the constant xRef' Q xRef of its FIRST solve (updateIneqConstr computes Jcons and never
        # writes it into b, MPC_branch.py:1999): N stale constants per child cone, the same for every branch, so the optimiser
        # does not see them; the library reports the objective with the constants of the current xRef
        obj = float(g[pre + "objective"]) - int(g["meta_N"]) * (jc(g["s0_xref"]) - jc(g[pre + "xref"]))
        assert abs(r["objective"][0] - obj) <= TOL_OBJ * abs(obj), (k, r["objective"][0], obj)


def check_merge_model_functions(make_backend):
    """PredictiveModel_merge point functions (dyn_linearization, zpred_eval, branch_eval, col_eval, xpred_eval) of both models
    sim_merge builds against the reference's CasADi MX graphs (tests/golden/merge_model_functions.npz); `make_backend(cfg)`
    returns an object with eval_model / set_lookup_table."""
    g = load_fixture("merge_model_functions")
    for mi, pols in enumerate((("trackv", "brake"), ("trackv_ref", "brake_ref"))):
        be = make_backend(scenarios.merge_config(policies=pols, N=int(g["N"]), v0=float(g["v0"])))
        if mi == 1:
            be.set_lookup_table(g["table_X"], g["table_psi"])
        r = be.eval_model(g["X"], g["Z"], g["U"])
        pre = "m%d_" % mi
        for k in ("A", "B", "C", "xp", "zpred", "p", "dh"):
            np.testing.assert_allclose(r[k], g[pre + k], atol=1e-10, err_msg="%s model %d" % (k, mi))
        np.testing.assert_allclose(r["hlin"], g[pre + "hlin"], atol=1e-10)
        # xpred_eval: the ego rollout under policy 0 = the first policy block of zpred evaluated at x
        rx = be.eval_model(g["X"], g["X"], g["U"])
        np.testing.assert_allclose(rx["zpred"][:, :, :4], g[pre + "xpred"], atol=1e-10)


# ---- belief-state MPC (PredictiveControllers.MPC) ---------------------------------------------------------------------
BELIEF_FIXTURES = ["belief_mpc_default", "belief_mpc_close"]


def belief_fixture_config(g, **kw):
    return scenarios.belief_config(N=int(g["meta_N"]), M=int(g["meta_M"]), m=int(g["meta_m"]), **kw)


def check_belief_fixture(solve_belief, g, tol=1e-6):
    """Closed loop recorded from the UNMODIFIED PredictiveControllers.MPC + HMM_backup_dyn.PredictiveModel + initMPCParams:
    `solve_belief(x0, b0 (1,M,m), xbackup (1,M*m,cols), xref)` must reproduce every step's plan - inputs, physical states and
    predicted beliefs - and the QP objective; the controller carries its own warm start from step to step."""
    for k in range(int(g["meta_steps"])):
        pre = "s%d_" % k
        r = solve_belief(g[pre + "x0"], g[pre + "b0"][None], g[pre + "xbackup"][None], g[pre + "xref"][:4])
        assert r["status"][0] in (0, 1), (k, r["status"][0])
        xP = g[pre + "xPred"]
        np.testing.assert_allclose(r["uPred"][0], g[pre + "uPred"], atol=tol)
        np.testing.assert_allclose(r["xPred"][0], xP[:, :4], atol=tol)
        np.testing.assert_allclose(r["bPred"][0], xP[:, 4:], atol=tol)
        assert np.abs(r["u0"][0] - g[pre + "uPred"][0]).max() < TOL_U0
        obj = float(g[pre + "objective"])
        assert abs(r["objective"][0] - obj) <= TOL_OBJ * abs(obj)
