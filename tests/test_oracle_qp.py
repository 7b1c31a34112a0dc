"""Oracle controller vs the QPs the unmodified reference assembles (golden fixtures).

Bit-level expectations: tree tables identical; every matrix/vector of the QP within 1e-11 of the
reference's (float64 round-off of differently ordered sums); exact optimum within 1e-8.
"""
import os

import numpy as np
import pytest
import scipy.sparse as sp

from oracle import params, qp_exact
from oracle.branch_mpc import TreeTopology


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name + ".npz"))


def _make(name, g):
    if name.startswith("highway_branch"):
        return params.highway_branch_mpc(list(g["meta_policies"]), int(g["meta_NB"]), int(g["meta_N"]),
                                         g["meta_lc_target"])
    return params.quadruped_prox_mpc(int(g["meta_NB"]), int(g["meta_N"]))


CASES = ["highway_branch_default", "highway_branch_close", "highway_branch_m2_nb3", "highway_branch_m3_nb1",
         "quadruped_prox_default"]


@pytest.mark.parametrize("name", CASES)
def test_assembled_qp_matches_reference(golden_dir, name):
    g = _load(golden_dir, name)
    mpc = _make(name, g)
    for k in range(int(g["meta_steps"])):
        pre = "s%d_" % k
        mpc.solve(g[pre + "x0"], g[pre + "z0"], g[pre + "xref"])
        P, q, A, l, u = mpc.qp
        Pg = sp.coo_matrix((g[pre + "P_v"], (g[pre + "P_r"], g[pre + "P_c"])), shape=tuple(g[pre + "P_shape"])).tocsc()
        Ag = sp.coo_matrix((g[pre + "A_v"], (g[pre + "A_r"], g[pre + "A_c"])), shape=tuple(g[pre + "A_shape"])).tocsc()
        assert abs(sp.triu(P) - Pg).max() < 1e-11
        assert abs(A - Ag).max() < 1e-11
        fin = np.isfinite(g[pre + "l"])
        assert (np.isfinite(l) == fin).all()
        np.testing.assert_allclose(l[fin], g[pre + "l"][fin], atol=1e-11)
        np.testing.assert_allclose(u, g[pre + "u"], atol=1e-11)
        np.testing.assert_allclose(q, g[pre + "q"], atol=1e-11)
        # topology: bit exact
        assert (mpc.topo.table() == g[pre + "tree"]).all()
        assert [mpc.topo.totalx, mpc.topo.totalu] == list(g[pre + "totals"])
        np.testing.assert_allclose(mpc.w, g[pre + "w"], atol=1e-13)
        np.testing.assert_allclose(np.vstack(mpc.xbar), g[pre + "xbar"], atol=1e-11)
        np.testing.assert_allclose(np.vstack(mpc.zbar), g[pre + "zbar"], atol=1e-11)
        np.testing.assert_allclose(np.vstack(mpc.ubar), g[pre + "ubar"], atol=1e-11)
        # optimum
        assert mpc.feasible == 1
        np.testing.assert_allclose(mpc.uPred, g[pre + "uPred"], atol=1e-8)
        np.testing.assert_allclose(mpc.xPred, g[pre + "xPred"], atol=1e-8)
        assert abs(mpc.objective - float(g[pre + "objective"])) < 1e-7 * max(1.0, abs(mpc.objective))
        # the stored solution is KKT-certified and HiGHS agreed wherever HiGHS reported success
        assert g[pre + "kkt"][0] < 1e-9 and g[pre + "kkt"][1] < 1e-9


def test_known_answer_default_first_solve(golden_dir):
    """Known answer of the reference workload's first solve (reproduced by the reference run)."""
    g = _load(golden_dir, "highway_branch_default")
    assert abs(float(g["s0_objective"]) + 42505.667089) < 1e-5
    np.testing.assert_allclose(g["s0_uPred"][0], [2.20795, -0.197606], atol=2e-5)
    np.testing.assert_allclose(g["s0_p"][0], [0.418793, 0.428505, 0.152702], atol=1e-6)


def test_topology_literal_table():
    """Highway default tree: (id, depth, ndx, ndu, parent)."""
    want = [(0, 0, 0, 0, -1), (1, 1, 1, 1, 0), (2, 1, 9, 9, 0), (3, 1, 17, 17, 0), (4, 2, 25, 25, 1), (5, 2, 34, 33, 1),
            (6, 2, 43, 41, 1), (7, 2, 52, 49, 2), (8, 2, 61, 57, 2), (9, 2, 70, 65, 2), (10, 2, 79, 73, 3),
            (11, 2, 88, 81, 3), (12, 2, 97, 89, 3)]
    T = TreeTopology(3, 2, 8)
    assert [tuple(r) for r in T.table()] == want
    assert (T.totalx, T.totalu) == (106, 97)


@pytest.mark.parametrize("m,NB,tx,tu,nb", [(2, 1, 19, 17, 3), (3, 1, 28, 25, 4), (4, 1, 37, 33, 5), (2, 2, 53, 49, 7),
                                            (3, 2, 106, 97, 13), (4, 2, 177, 161, 21), (2, 3, 121, 113, 15),
                                            (3, 3, 340, 313, 40), (4, 3, 737, 673, 85)])
def test_topology_closed_form_sizes(m, NB, tx, tu, nb):
    T = TreeTopology(m, NB, 8)
    assert (T.totalx, T.totalu, T.nbranch) == (tx, tu, nb)
    assert T.totalu == 1 + 8 * sum(m ** k for k in range(1, NB + 1))
    assert T.totalx == T.totalu + m ** NB


def test_two_solvers_agree(golden_dir):
    """Interior point + polish vs HiGHS on a stored reference QP (only where HiGHS reports Optimal)."""
    g = _load(golden_dir, "highway_branch_m3_nb1")
    pre = "s0_"
    P = sp.coo_matrix((g[pre + "P_v"], (g[pre + "P_r"], g[pre + "P_c"])), shape=tuple(g[pre + "P_shape"])).tocsc()
    A = sp.coo_matrix((g[pre + "A_v"], (g[pre + "A_r"], g[pre + "A_c"])), shape=tuple(g[pre + "A_shape"])).tocsc()
    z, y, info = qp_exact.solve_qp(P, g[pre + "q"], A, g[pre + "l"], g[pre + "u"])
    assert info["polished"]
    cert = qp_exact.kkt_residuals(P, g[pre + "q"], A, g[pre + "l"], g[pre + "u"], z, y)
    assert cert["primal"] < 1e-9 and cert["dual"] < 1e-9
    zh, _, status = qp_exact.solve_qp_highs(P, g[pre + "q"], A, g[pre + "l"], g[pre + "u"])
    if status == "Optimal":
        nxu = 28 * 4 + 25 * 2
        assert np.abs(zh[:nxu] - z[:nxu]).max() < 1e-4


def test_robust_mpc_qp_matches_reference(golden_dir):
    """robustMPC (MPC_branch.py:1275): assembled chain QP and optimum vs the unmodified reference, 3 closed-loop steps."""
    g = _load(golden_dir, "highway_robust_default")
    mpc = params.highway_robust_mpc(list(g["meta_policies"]), int(g["meta_NB"]), int(g["meta_N"]), g["meta_lc_target"])
    for k in range(int(g["meta_steps"])):
        pre = "s%d_" % k
        mpc.solve(g[pre + "x0"], g[pre + "z0"], g[pre + "xref"])
        P, q, A, l, u = mpc.qp
        Pg = sp.coo_matrix((g[pre + "P_v"], (g[pre + "P_r"], g[pre + "P_c"])), shape=tuple(g[pre + "P_shape"])).tocsc()
        Ag = sp.coo_matrix((g[pre + "A_v"], (g[pre + "A_r"], g[pre + "A_c"])), shape=tuple(g[pre + "A_shape"])).tocsc()
        assert abs(sp.triu(P) - Pg).max() < 1e-11 and abs(A - Ag).max() < 1e-11
        np.testing.assert_allclose(q, g[pre + "q"], atol=1e-11)
        np.testing.assert_allclose(u, g[pre + "u"], atol=1e-11)
        assert mpc.feasible == 1
        np.testing.assert_allclose(mpc.uPred, g[pre + "uPred"], atol=1e-8)
        np.testing.assert_allclose(mpc.xPred, g[pre + "xPred"], atol=1e-8)
