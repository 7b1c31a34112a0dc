"""Parity at batch scale.

(a) An INDEPENDENT feasibility check of whole BASELINE-size batches: the linearised dynamics are re-evaluated point-wise with
    bmpc_eval_model at the returned linearisation states and every transition of the returned plan is checked against
    them (x+ = A x + B u + C within 1e-5, the north_star bar for constraint violation), the objective the kernel reports is
    recomputed in numpy from the returned plan (soft-row penalties included, i.e. the slacks are consistent), inputs sit
    in their box.
(b) Oracle comparisons on hundreds of random problems (first input 1e-3, objective 1e-4 relative): >= 256 highway,
    >= 64 quadruped, >= 64 robustMPC, oracle solves spread over the host cores.
"""
import multiprocessing as mp
import os

import numpy as np
import pytest

from tests.helpers import TOL_OBJ, TOL_U0, TOL_VIOL
from _bmpc import abi, batch, scenarios

pytestmark = pytest.mark.gpu


def _transitions(topo, N):
    """(u-node k, x-node of k, successor x-nodes) for every input node, from the BFS tables (id, depth, ndx, ndu, parent)."""
    nb = len(topo)
    ndx, ndu, depth, parent = topo[:, 2], topo[:, 3], topo[:, 1], topo[:, 4]
    NB = depth.max()
    children = {b: [c for c in range(nb) if parent[c] == b] for b in range(nb)}
    out = []
    for b in range(nb):
        l = 1 if b == 0 else N
        for t in range(l):
            k, xk = ndu[b] + t, ndx[b] + t
            if t < l - 1:
                succ = [xk + 1]
            elif depth[b] < NB:
                succ = [ndx[c] for c in children[b]]
            else:
                succ = [xk + 1]
            out.append((k, xk, succ))
    return out


def _highway_objective(r, topo, xref, N, lam=300.0, bx=(4 * 3.6 - 1.25, -1.25, 0.25, 0.25)):
    """QP objective (1/2 z'Pz + q'z with the slacks eliminated) of the effective BranchMPC, recomputed from the outputs:
    MPC_branch.py:1064-1112 in stage form."""
    Q = np.diag([0., 3., 3., 10.])
    R = np.diag([1., 100.])
    nb = len(topo)
    ndx, ndu, depth = topo[:, 2], topo[:, 3], topo[:, 1]
    NB = depth.max()
    x, u, xbar, w = r["xPred"], r["uPred"], r["xLin"], r["branch_w"]
    z = r["zPred"]
    J = np.zeros(x.shape[0])
    for b in range(nb):
        l = 1 if b == 0 else N
        wb = w[:, b]
        for t in range(l):
            k, xk = ndu[b] + t, ndx[b] + t
            xs, us, xb = x[:, xk], u[:, k], xbar[:, k]
            Ql = Q                                     # Qf = Q on the highway (MPC_branch.py:52)
            # w [x'(Q + dQ)x - 2 (xRef'Q + xbar'dQ) x + u'R u],  dQ = Q/2 (:1070-1099)
            J += wb * (np.einsum("bi,ij,bj->b", xs, 1.5 * Q, xs) - 2.0 * np.einsum("bi,bi->b", xref @ Ql + 0.5 * (xb @ Q), xs)
                       + np.einsum("bi,ij,bj->b", us, R, us))
            # soft rows: collision row linearised at (xbar, zbar) + the four state rows
            dxv = np.abs(xb[:, 0] - z[:, k, 0]) - 5.0
            dyv = np.abs(xb[:, 1] - z[:, k, 1]) - 2.7
            mx = np.maximum(dxv, dyv)
            ex, ey = np.exp(dxv - mx), np.exp(dyv - mx)
            h = (dxv * ex + dyv * ey) / (ex + ey)
            wx, wy = ex / (ex + ey), ey / (ex + ey)
            dhx = np.sign(xb[:, 0] - z[:, k, 0]) * wx * (1 + dxv - h)
            dhy = np.sign(xb[:, 1] - z[:, k, 1]) * wy * (1 + dyv - h)
            viol = np.maximum(-(dhx * xs[:, 0] + dhy * xs[:, 1]) - (h - dhx * xb[:, 0] - dhy * xb[:, 1]), 0.0)
            viol += np.maximum(xs[:, 1] - bx[0], 0) + np.maximum(-xs[:, 1] - bx[1], 0)
            viol += np.maximum(xs[:, 3] - bx[2], 0) + np.maximum(-xs[:, 3] - bx[3], 0)
            J += lam * wb * viol
        if depth[b] == NB:
            xT = x[:, ndx[b] + N]
            J += wb * np.einsum("bi,ij,bj->b", xT, Q, xT)      # terminal node: w Qf, no linear term (:1094)
    return J


def _check_plan(mpc, r, x0, topo, N, n, d, pp):
    B = x0.shape[0]
    assert np.array_equal(r["xPred"][:, 0], x0)
    tr = _transitions(topo, N)
    # re-linearise at the returned linearisation states (A, C of the highway model do not depend on the input)
    ks = np.array([k for k, _, _ in tr])
    pts = r["xLin"][:, ks].reshape(-1, n)
    ppb = None if pp is None else np.repeat(pp, len(ks), axis=0)
    ev = mpc.eval_model(pts, pts, np.zeros((pts.shape[0], d)), ppb)
    A = ev["A"].reshape(B, len(ks), n, n)
    Bm = ev["B"].reshape(B, len(ks), n, d)
    C = ev["C"].reshape(B, len(ks), n)
    worst = 0.0
    for j, (k, xk, succ) in enumerate(tr):
        pred = np.einsum("bij,bj->bi", A[:, j], r["xPred"][:, xk]) + np.einsum("bij,bj->bi", Bm[:, j], r["uPred"][:, k]) + C[:, j]
        for s in succ:
            worst = max(worst, float(np.abs(r["xPred"][:, s] - pred).max()))
    return worst


def test_highway_16384_plan_is_feasible_and_consistent():
    B = 16384
    mpc = batch.BatchedBranchMPC(scenarios.highway_config(batch_capacity=B))
    x0, z0, xref, pp = scenarios.highway_batch(B, seed=31)
    topo = mpc.topology()
    for step in range(2):                                # cold, then warm
        r = mpc.solve_host(x0, z0, xref, pp)
        assert (r["status"] <= abi.STATUS_CONVERGED).all(), np.bincount(r["status"])
        assert (np.abs(r["uPred"][:, :, 0]) <= 6.0 + 1e-12).all() and (np.abs(r["uPred"][:, :, 1]) <= 0.3 + 1e-12).all()
        worst = _check_plan(mpc, r, x0, topo, 8, 4, 2, pp)
        assert worst <= TOL_VIOL, worst
        J = _highway_objective(r, topo, xref, 8)
        rel = np.abs(J - r["objective"]) / np.maximum(1.0, np.abs(r["objective"]))
        assert rel.max() < 1e-9, rel.max()
        x0 = scenarios.euler_highway(x0, r["u0"])
        z0 = scenarios.euler_highway(z0, np.column_stack([np.zeros(B), -0.1 * z0[:, 3]]))
    mpc.close()


def test_robust_4096_plan_is_feasible():
    B = 4096
    cfg = scenarios.highway_config(batch_capacity=B)
    cfg.controller = abi.CTRL_ROBUST
    mpc = batch.BatchedBranchMPC(cfg)
    ev = batch.BatchedBranchMPC(scenarios.highway_config(batch_capacity=1))     # point-wise model functions
    x0, z0, xref, pp = scenarios.highway_batch(B, seed=32)
    r = mpc.solve_host(x0, z0, xref, pp, outputs=("u0", "uPred", "xPred", "objective", "status"))
    assert (r["status"] <= abi.STATUS_CONVERGED).all(), np.bincount(r["status"])
    r2 = mpc.solve_host(scenarios.euler_highway(x0, r["u0"]), scenarios.euler_highway(z0, np.column_stack([np.zeros(B), -0.1 * z0[:, 3]])),
                        xref, pp, outputs=("u0", "uPred", "xPred", "objective", "status"))
    assert (r2["status"] <= abi.STATUS_CONVERGED).all(), np.bincount(r2["status"])
    # second step linearises every node about the first plan shifted by one step (MPC_branch.py:1429-1431)
    Nx = 18
    xl = np.concatenate([r["xPred"][:, 1:], r["xPred"][:, -1:]], axis=1)
    pts = xl[:, :Nx - 1].reshape(-1, 4)
    e = ev.eval_model(pts, pts, np.zeros((pts.shape[0], 2)))
    A = e["A"].reshape(B, Nx - 1, 4, 4)
    C = e["C"].reshape(B, Nx - 1, 4)
    Bm = e["B"].reshape(B, Nx - 1, 4, 2)
    pred = np.einsum("bkij,bkj->bki", A, r2["xPred"][:, :-1]) + np.einsum("bkij,bkj->bki", Bm, r2["uPred"]) + C
    assert np.abs(pred - r2["xPred"][:, 1:]).max() <= TOL_VIOL
    assert (np.abs(r2["uPred"][:, :, 0]) <= 6.0 + 1e-12).all() and (np.abs(r2["uPred"][:, :, 1]) <= 0.3 + 1e-12).all()
    mpc.close()
    ev.close()


# ---- oracle comparisons on hundreds of problems ------------------------------------------------------------------------
def _ora_highway(a):
    from oracle import params
    x, z, r, lc = a
    o = params.highway_branch_mpc(lc_target=lc)
    u = o.solve(x, z, r).copy()
    return u, o.objective, o.feasible


def _ora_quadruped(a):
    from oracle import params
    x, z, r = a
    o = params.quadruped_prox_mpc()
    u = o.solve(x, z, r).copy()
    return u, o.objective, o.feasible


def _ora_robust(a):
    from oracle import params
    x, z, r, lc = a
    o = params.highway_robust_mpc(lc_target=lc)
    u = o.solve(x, z, r).copy()
    return u, o.objective, o.feasible


def _pool_map(fn, items):
    with mp.get_context("spawn").Pool(min(len(items), os.cpu_count() or 1)) as pool:
        return pool.map(fn, items)


def _compare(r, ref, min_certified):
    """Only problems whose ORACLE solve certified its optimum (KKT residuals + verified active set) are a reference; on the
    few it cannot certify (degenerate robustMPC rows) the device result must at least not be worse than the oracle's iterate."""
    ok = [i for i in range(len(ref)) if ref[i][2] == 1]
    assert len(ok) >= min_certified, len(ok)
    du = max(float(np.abs(r["u0"][i] - ref[i][0]).max()) for i in ok)
    dj = max(abs(r["objective"][i] - ref[i][1]) / abs(ref[i][1]) for i in ok)
    assert du < TOL_U0 and dj < TOL_OBJ, (du, dj)
    for i in range(len(ref)):
        if ref[i][2] != 1:
            assert r["objective"][i] <= ref[i][1] + TOL_OBJ * abs(ref[i][1]), (i, r["objective"][i], ref[i][1])


def test_256_highway_problems_against_oracle():
    B = 256
    x0, z0, xref, pp = scenarios.highway_batch(B, seed=4041)
    mpc = batch.BatchedBranchMPC(scenarios.highway_config(batch_capacity=B))
    r = mpc.solve_host(x0, z0, xref, pp, outputs=("u0", "objective", "status"))
    assert (r["status"] <= abi.STATUS_CONVERGED).all()
    _compare(r, _pool_map(_ora_highway, [(x0[i], z0[i], xref[i], pp[i, 2]) for i in range(B)]), 256)
    mpc.close()


def test_64_quadruped_problems_against_oracle():
    B = 64
    x0, z0, xref = scenarios.quadruped_batch(B, seed=4042, goal=(5.0, 5.0, 0.0))
    mpc = batch.BatchedBranchMPC(scenarios.quadruped_config(batch_capacity=B))
    r = mpc.solve_host(x0, z0, xref, outputs=("u0", "objective", "status"))
    assert (r["status"] <= abi.STATUS_CONVERGED).all()
    _compare(r, _pool_map(_ora_quadruped, [(x0[i], z0[i], xref[i]) for i in range(B)]), 62)
    mpc.close()


def test_64_robust_problems_against_oracle():
    B = 64
    x0, z0, xref, pp = scenarios.highway_batch(B, seed=4043)
    cfg = scenarios.highway_config(batch_capacity=B)
    cfg.controller = abi.CTRL_ROBUST
    mpc = batch.BatchedBranchMPC(cfg)
    r = mpc.solve_host(x0, z0, xref, pp, outputs=("u0", "objective", "status"))
    assert (r["status"] <= abi.STATUS_CONVERGED).all()
    _compare(r, _pool_map(_ora_robust, [(x0[i], z0[i], xref[i], pp[i, 2]) for i in range(B)]), 60)
    mpc.close()
