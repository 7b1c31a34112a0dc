"""Generate golden fixtures by running the UNMODIFIED reference in this container.

Run from the repo root (only where `/root/reference` exists; the GPU box never runs this):

    python tests/golden/make_golden.py

What runs: the reference's own `highway_branch_dyn.PredictiveModel`,
`quadruped_branch_dyn.PredictiveModel`, `Init_MPC.init*`, `MPC_branch.BranchMPC`,
`BranchMPCProx` and `robustMPC`, imported from `/root/reference` on top of the shims in
`tests/golden/shims/` (casadi -> a small expression graph with forward-mode AD; osqp -> exact
QP solve + HiGHS cross-check).  So every matrix stored here (H, q, F, b, G, E x + L), every tree
table and every model-function value was produced by the reference's code; only the third-party
numerical back-ends are substituted.

Fixtures (tests/golden/*.npz, a few tens of KB each) store the QP in sparse triplets.
"""
import os
import sys

import numpy as np
import scipy.sparse as sp

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "shims"))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import refenv  # noqa: E402

with refenv.reference_imports():
    import highway_branch_dyn as hw  # noqa: E402
    import quadruped_branch_dyn as qd  # noqa: E402
    import utils as rutils  # noqa: E402
    import MPC_branch  # noqa: E402
    import Init_MPC  # noqa: E402
    import osqp  # noqa: E402
    import ecos  # noqa: E402


def highway_cons():
    # main_branch.py:37
    return rutils.Branch_constants(s1=2, s2=3, c2=0.5, tran_diag=0.3, alpha=1, R=1.2, am=6.0, rm=0.3,
                                   J_c=20, s_c=1, ylb=0., yub=7.2, L=4, W=2.5, col_alpha=5, Kpsi=0.1)


def highway_policies(names, cons, lc_target, v0=20.0):
    table = {
        "maintain": lambda x: hw.backup_maintain(x, cons),
        "brake": lambda x: hw.backup_brake(x, cons),
        "lc": lambda x: hw.backup_lc(x, lc_target),
        "trackv": None,   # MX-only symbolic branch in the reference (highway_branch_dyn.py:80-88): not usable with SX models
    }
    return [table[n] for n in names]


def quad_cons():
    # main_quadruped.py:31
    return rutils.Quad_constants(s1=2, s2=3, c2=0.5, alpha=1, R=1.2, vxm=0.2, vym=0.1, rm=0.5,
                                 L1=0.5, W1=0.3, L2=1, W2=0.6, col_tol=0.2, col_alpha=5)


def tree_tables(mpc):
    """BFS tables of the reference tree: id, depth, ndx, ndu, parent, w, and stacked trajectories."""
    ids = {b: k for k, b in enumerate(mpc.ndx)}
    parent = {mpc.BT: -1}
    for b in mpc.ndx:
        for c in b.children:
            parent[c] = ids[b]
    rows = []
    for b in mpc.ndx:
        rows.append((ids[b], b.depth, mpc.ndx[b], mpc.ndu[b], parent[b]))
    w = np.array([b.w for b in mpc.ndx], dtype=float)
    p = np.array([b.p if b.p is not None else np.full(mpc.m, np.nan) for b in mpc.ndx], dtype=float)
    xbar = np.vstack([b.xtraj for b in mpc.ndx])
    zbar = np.vstack([b.ztraj for b in mpc.ndx])
    ubar = np.vstack([b.utraj for b in mpc.ndx])
    return np.array(rows, dtype=np.int64), w, p, xbar, zbar, ubar


def coo(mat):
    m = sp.coo_matrix(mat)
    return m.row.astype(np.int32), m.col.astype(np.int32), m.data.astype(float), np.array(m.shape, dtype=np.int64)


def record_step(store, k, mpc, x, z, xref, tree=True):
    lp = osqp.last_problem
    assert lp["ok"], "oracle QP solve did not certify"
    pre = "s%d_" % k
    store[pre + "x0"] = np.array(x, dtype=float)
    store[pre + "z0"] = np.array(z, dtype=float)
    store[pre + "xref"] = np.array(xref, dtype=float)
    for name, mat in (("P", sp.triu(lp["P"])), ("A", lp["A"])):
        r, c, v, shp = coo(mat)
        store[pre + name + "_r"] = r
        store[pre + name + "_c"] = c
        store[pre + name + "_v"] = v
        store[pre + name + "_shape"] = shp
    store[pre + "q"] = lp["q"]
    store[pre + "l"] = lp["l"]
    store[pre + "u"] = lp["u"]
    store[pre + "sol"] = lp["x"]
    store[pre + "objective"] = np.array(lp["cert"]["objective"])
    store[pre + "kkt"] = np.array([lp["cert"]["primal"], lp["cert"]["dual"]])
    store[pre + "highs_objective"] = np.array(lp["highs_objective"])
    store[pre + "highs_xu_maxdiff"] = np.array(
        np.abs(lp["highs_x"][: mpc_nxu(mpc)] - lp["x"][: mpc_nxu(mpc)]).max())
    store[pre + "xPred"] = np.array(mpc.xPred)
    store[pre + "uPred"] = np.array(mpc.uPred)
    if tree:
        tab, w, p, xbar, zbar, ubar = tree_tables(mpc)
        store[pre + "tree"] = tab
        store[pre + "w"] = w
        store[pre + "p"] = p
        store[pre + "xbar"] = xbar
        store[pre + "zbar"] = zbar
        store[pre + "ubar"] = ubar
        store[pre + "totals"] = np.array([mpc.totalx, mpc.totalu])
    print("   step %d: obj %.6f  u0 %s  kkt %.1e/%.1e  highs(%s) xu-diff %.1e" % (
        k, lp["cert"]["objective"], np.array2string(mpc.uPred[0], precision=6),
        lp["cert"]["primal"], lp["cert"]["dual"], lp["highs_status"], store[pre + "highs_xu_maxdiff"]))


def mpc_nxu(mpc):
    if hasattr(mpc, "totalx") and mpc.totalx:
        return mpc.totalx * mpc.n + mpc.totalu * mpc.d
    return mpc.Nx * mpc.n + mpc.Nu * mpc.d


def euler_highway(x, u, dt):
    return x + dt * np.array([x[2] * np.cos(x[3]), x[2] * np.sin(x[3]), u[0], u[1]])


def run_highway(name, ctrl, policies, NB, x, z, xref, steps, N=8, lc_target=(0.5, 1.8, 15., 0.), n_lane_mpc=4):
    print("highway fixture", name)
    cons = highway_cons()
    lc_target = np.array(lc_target, dtype=float)
    model = hw.PredictiveModel(4, 2, N, highway_policies(policies, cons, lc_target), 0.1, cons)
    par = Init_MPC.initBranchMPC(4, 2, N, NB, lc_target, 6.0, 0.3, n_lane_mpc, cons.W)
    mpc = getattr(MPC_branch, ctrl)(par, model)
    store = {"meta_policies": np.array(policies), "meta_ctrl": np.array(ctrl), "meta_NB": np.array(NB),
             "meta_N": np.array(N), "meta_lc_target": lc_target, "meta_steps": np.array(steps),
             "meta_n_lane_mpc": np.array(n_lane_mpc)}
    x = np.array(x, dtype=float)
    z = np.array(z, dtype=float)
    for k in range(steps):
        mpc.solve(x, z, np.array(xref, dtype=float))
        record_step(store, k, mpc, x, z, xref, tree=(ctrl != "robustMPC"))
        x = euler_highway(x, mpc.uPred[0], 0.1)
        z = euler_highway(z, np.array([0., -cons.Kpsi * z[3]]), 0.1)     # obstacle keeps 'maintain'
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **store)


def run_highway_cvar(name, policies, NB, x, z, xref, steps, ralpha, N=8, lc_target=(0.5, 1.8, 15., 0.), n_lane_mpc=4):
    """The controller main_branch.py:48 instantiates: the UNMODIFIED `BranchMPC_CVaR` (MPC_branch.py:1598-2152) in closed
    loop; its cone program (c, G, h, dims, A, b as handed to ecos.solve, :2136) is recorded at every step together with the
    optimum the oracle's interior point returns for it, the risk multipliers and the tree the program was built on."""
    print("highway CVaR fixture", name)
    cons = highway_cons()
    lc_target = np.array(lc_target, dtype=float)
    model = hw.PredictiveModel(4, 2, N, highway_policies(policies, cons, lc_target), 0.1, cons)
    par = Init_MPC.initBranchMPC(4, 2, N, NB, lc_target, 6.0, 0.3, n_lane_mpc, cons.W)
    mpc = MPC_branch.BranchMPC_CVaR(par, model, ralpha=ralpha)
    store = {"meta_policies": np.array(policies), "meta_ctrl": np.array("BranchMPC_CVaR"), "meta_NB": np.array(NB),
             "meta_N": np.array(N), "meta_lc_target": lc_target, "meta_steps": np.array(steps),
             "meta_n_lane_mpc": np.array(n_lane_mpc), "meta_ralpha": np.array(ralpha)}
    x = np.array(x, dtype=float)
    z = np.array(z, dtype=float)
    for k in range(steps):
        ulin_before = None if mpc.uLin is None or mpc.BT is None else np.array(mpc.uLin)
        pbest_before = None if mpc.BT is None else np.array([int(np.argmax(b.p)) if b.p is not None else 0 for b in mpc.ndx])
        mpc.solve(x, z, np.array(xref, dtype=float))
        lp = ecos.last_problem
        assert mpc.feasible and lp["info"]["exitFlag"] == 0, "oracle cone solve did not converge"
        pre = "s%d_" % k
        store[pre + "x0"], store[pre + "z0"], store[pre + "xref"] = x.copy(), z.copy(), np.array(xref, dtype=float)
        for nm, mat in (("G", lp["G"]), ("A", lp["A"])):
            r, c, v, shp = coo(mat)
            store[pre + nm + "_r"], store[pre + nm + "_c"], store[pre + nm + "_v"], store[pre + nm + "_shape"] = r, c, v, shp
        store[pre + "c"], store[pre + "h"], store[pre + "b"] = lp["c"], lp["h"], lp["b"]
        store[pre + "dims_l"] = np.array(lp["dims"]["l"])
        store[pre + "dims_q"] = np.array(lp["dims"]["q"])
        store[pre + "sol"] = lp["x"]
        store[pre + "objective"] = np.array(lp["info"]["pcost"])
        store[pre + "cert"] = np.array([lp["info"]["gap"], lp["info"]["pres"], lp["info"]["dres"]])
        store[pre + "xPred"], store[pre + "uPred"] = np.array(mpc.xPred), np.array(mpc.uPred)
        if ulin_before is not None:
            store[pre + "uLin_before"] = ulin_before        # warm-start state the step was linearised about
            store[pre + "pbest_before"] = pbest_before
        tab, w, p, xbar, zbar, ubar = tree_tables(mpc)
        store[pre + "tree"], store[pre + "w"], store[pre + "p"] = tab, w, p
        store[pre + "xbar"], store[pre + "zbar"], store[pre + "ubar"] = xbar, zbar, ubar
        store[pre + "totals"] = np.array([mpc.totalx, mpc.totalu])
        store[pre + "branchidx"] = np.array([list(mpc.ndx).index(b) for b in mpc.branchidx])
        print("   step %d: J %.8f  u0 %s  gap %.1e pres %.1e dres %.1e (%d it)" % (
            k, lp["info"]["pcost"], np.array2string(mpc.uPred[0], precision=6), lp["info"]["gap"], lp["info"]["pres"],
            lp["info"]["dres"], lp["info"]["iter"]))
        x = euler_highway(x, mpc.uPred[0], 0.1)
        z = euler_highway(z, np.array([0., -cons.Kpsi * z[3]]), 0.1)     # obstacle keeps 'maintain'
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **store)


def run_quadruped(name, x, z, xref, steps, N=25, NB=2):
    print("quadruped fixture", name)
    cons = quad_cons()
    v0 = 0.2
    policies = [lambda s: qd.backup_forward(s, v0), lambda s: qd.backup_stop(s)]
    model = qd.PredictiveModel(3, 3, N, policies, 0.2, cons)
    par = Init_MPC.initquadBranchMPC(3, 3, N, NB, np.array(xref, dtype=float), 0.2, 0.1, 0.5)
    mpc = MPC_branch.BranchMPCProx(par, model)
    store = {"meta_ctrl": np.array("BranchMPCProx"), "meta_NB": np.array(NB), "meta_N": np.array(N),
             "meta_steps": np.array(steps), "meta_v0": np.array(v0)}
    x = np.array(x, dtype=float)
    z = np.array(z, dtype=float)
    for k in range(steps):
        mpc.solve(x, z, np.array(xref, dtype=float))
        record_step(store, k, mpc, x, z, xref)
        u = mpc.uPred[0]
        x = x + 0.2 * np.array([u[0] * np.cos(x[2]) - u[1] * np.sin(x[2]),
                                u[0] * np.sin(x[2]) + u[1] * np.cos(x[2]), u[2]])
        z = z + 0.2 * np.array([v0 * np.cos(z[2]), v0 * np.sin(z[2]), 0.])  # obstacle keeps 'forward'
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **store)


def model_function_vectors():
    """Point-wise values of the reference's model Functions at seeded random inputs."""
    print("model-function fixture")
    rng = np.random.default_rng(20240607)
    cons = highway_cons()
    lc_target = np.array([0.5, 5.4, 18., 0.])
    model = hw.PredictiveModel(4, 2, 8, highway_policies(["maintain", "brake", "lc"], cons, lc_target), 0.1, cons)
    K = 24
    X = np.column_stack([rng.uniform(-5, 60, K), rng.uniform(0.5, 14, K), rng.uniform(5, 30, K), rng.normal(0, 0.1, K)])
    Z = np.column_stack([X[:, 0] + rng.uniform(-15, 25, K), rng.uniform(0.5, 14, K), rng.uniform(5, 30, K),
                         rng.normal(0, 0.1, K)])
    U = np.column_stack([rng.uniform(-6, 6, K), rng.uniform(-0.3, 0.3, K)])
    out = {"hw_X": X, "hw_Z": Z, "hw_U": U, "hw_lc_target": lc_target}
    A, B, C, XP, ZP, P, DP, H, DH = [], [], [], [], [], [], [], [], []
    for k in range(K):
        a, b, c, xp = model.dyn_linearization(X[k], U[k])
        A.append(a); B.append(b); C.append(c); XP.append(xp)
        ZP.append(model.zpred_eval(Z[k]))
        p, dp = model.branch_eval(X[k], Z[k])
        P.append(p); DP.append(dp)
        h, dh = model.col_eval(X[k], Z[k])
        H.append(h); DH.append(dh)
    out.update(hw_A=np.array(A), hw_B=np.array(B), hw_C=np.array(C), hw_xp=np.array(XP), hw_zpred=np.array(ZP),
               hw_p=np.array(P), hw_dp=np.array(DP), hw_hlin=np.array(H), hw_dh=np.array(DH))

    qc = quad_cons()
    v0 = 0.2
    qmodel = qd.PredictiveModel(3, 3, 25, [lambda s: qd.backup_forward(s, v0), lambda s: qd.backup_stop(s)], 0.2, qc)
    X = np.column_stack([rng.uniform(-2, 6, K), rng.uniform(-4, 4, K), rng.uniform(-np.pi, np.pi, K)])
    Z = np.column_stack([X[:, 0] + rng.uniform(-4, 4, K), X[:, 1] + rng.uniform(-4, 4, K), rng.uniform(-np.pi, np.pi, K)])
    U = np.column_stack([rng.uniform(0, 0.2, K), rng.uniform(-0.1, 0.1, K), rng.uniform(-0.5, 0.5, K)])
    out.update(qd_X=X, qd_Z=Z, qd_U=U)
    A, B, C, XP, ZP, P, DP, H, DH = [], [], [], [], [], [], [], [], []
    for k in range(K):
        a, b, c, xp = qmodel.dyn_linearization(X[k], U[k])
        A.append(a); B.append(b); C.append(c); XP.append(xp)
        ZP.append(qmodel.zpred_eval(Z[k]))
        p, dp = qmodel.branch_eval(X[k], Z[k])
        P.append(p); DP.append(dp)
        h, dh = qmodel.col_eval(X[k], Z[k])
        H.append(h); DH.append(dh)
    out.update(qd_A=np.array(A), qd_B=np.array(B), qd_C=np.array(C), qd_xp=np.array(XP), qd_zpred=np.array(ZP),
               qd_p=np.array(P), qd_dp=np.array(DP), qd_hlin=np.array(H), qd_dh=np.array(DH))
    np.savez_compressed(os.path.join(HERE, "model_functions.npz"), **out)


def hmm_vectors():
    """Numeric functions of the reference's HMM_backup_dyn.py (belief-state model): backup rollouts with and without
    the sensitivity matrix, belief transition, normalised collision function.  The module imports a name its `utils.py`
    does not define (`HMM_constants`, HMM_backup_dyn.py:5); a dataclass with the fields the module reads is injected
    before the import - the module's code is untouched."""
    import dataclasses
    print("hmm fixture")

    @dataclasses.dataclass
    class HMM_constants:
        s1: float = None; s2: float = None; c2: float = None; tran_diag: float = None; alpha: float = None
        R: float = None; am: float = None; rm: float = None; J_c: float = None; s_c: float = None
        ylb: float = None; yub: float = None; W: float = None; L: float = None; col_alpha: float = None
        Kpsi: float = None

    with refenv.reference_imports():
        rutils.HMM_constants = HMM_constants
        import HMM_backup_dyn as hmm
    cons = HMM_constants(s1=2, s2=3, c2=0.5, tran_diag=0.3, alpha=1, R=1.2, am=6.0, rm=0.3, J_c=20, s_c=1, ylb=0., yub=7.2,
                         L=4, W=2.5, col_alpha=5, Kpsi=0.1)
    rng = np.random.default_rng(777)
    K, M, m, N, dt = 12, 2, 2, 10, 0.1
    backupcons = [lambda s: hmm.backup_maintain(s, cons), lambda s: hmm.backup_brake(s, cons)]
    X0 = np.stack([np.column_stack([rng.uniform(-10, 30, M), rng.uniform(0.5, 6.5, M), rng.uniform(8, 25, M),
                                    rng.normal(0, 0.05, M)]) for _ in range(K)])
    # the constructor builds symbolic graphs the casadi stand-in does not cover; the rollout METHOD is called unbound on a
    # plain attribute holder (same code, HMM_backup_dyn.py:204-214)
    import types
    holder = types.SimpleNamespace(M=M, m=m, n=4, dt=dt, backupcons=backupcons)
    XB = np.stack([np.array(hmm.PredictiveModel.generate_backup_traj(holder, X0[k], N)) for k in range(K)])   # (K, M*m, N*n)
    # module-level rollout with sensitivity (HMM_backup_dyn.py:54-85): fixed number of steps
    steps, ts = 6, 0.05
    f0 = np.array([1.0, 0.0, 0.0, 0.0])
    XX, QQ, QT = [], [], []
    for k in range(K):
        for j in range(m):
            tt, xx, uu, Q, Qt = hmm.generate_backup_traj(X0[k, 0], backupcons[j], lambda s, t: t >= steps * ts - 1e-12, f0, ts=ts)
            XX.append(np.array(xx)); QQ.append(np.array(Q)); QT.append(np.array(Qt))
    # belief transition + input probability (HMM_backup_dyn.py:96-104) and the h used by the model (:255)
    EGO = np.column_stack([rng.uniform(-5, 5, K), rng.uniform(0.5, 6.5, K), rng.uniform(10, 25, K), rng.normal(0, 0.05, K)])
    H_all, h_all, bnext, bnext_env = [], [], [], []
    Bm = rng.dirichlet(np.ones(m), size=(K, M))
    CBF = rng.uniform(-1, 2, size=(K, M, m))
    for k in range(K):
        for i in range(M):
            h = np.zeros(m)
            for j in range(m):
                xb = XB[k, m * i + j].reshape(N, 4, order="F")[3]                     # backup state after 4 steps
                hc = float(hmm.veh_col(EGO[k], xb, [cons.L + 1, cons.W + 0.2]))
                # symbolic (un-clipped) branch value equals the numeric one inside the +-5 clip range
                h[j] = float(hmm.softmin(hc, float(hmm.lane_bdry_h(xb, cons.ylb, cons.yub)), cons.col_alpha))
            H = np.array(hmm.backup_trans(h, cons))
            bi = Bm[k, i] @ H
            be = bi * np.array([hmm.backup_input_prob(CBF[k, i, j], cons) for j in range(m)])
            H_all.append(H); h_all.append(h); bnext.append(bi); bnext_env.append(be / be.sum())
    np.savez_compressed(os.path.join(HERE, "hmm_functions.npz"), X0=X0, XB=XB, N=np.array(N), dt=np.array(dt),
                        sens_x=np.array(XX), sens_Q=np.array(QQ), sens_Qt=np.array(QT), sens_f0=f0, sens_ts=np.array(ts),
                        sens_steps=np.array(steps), EGO=EGO, B=Bm, CBF=CBF, H=np.array(H_all).reshape(K, M, m, m),
                        h=np.array(h_all).reshape(K, M, m), b_next=np.array(bnext).reshape(K, M, m),
                        b_next_env=np.array(bnext_env).reshape(K, M, m), t_index=np.array(3))


def _hmm_module():
    """The belief-state model module of the reference.  It imports a name its `utils.py` does not define (`HMM_constants`,
    HMM_backup_dyn.py:5); a dataclass with the fields the module reads is injected before the import - nothing else."""
    import dataclasses

    @dataclasses.dataclass
    class HMM_constants:
        s1: float = None; s2: float = None; c2: float = None; tran_diag: float = None; alpha: float = None
        R: float = None; am: float = None; rm: float = None; J_c: float = None; s_c: float = None
        ylb: float = None; yub: float = None; W: float = None; L: float = None; col_alpha: float = None
        Kpsi: float = None

    with refenv.reference_imports():
        rutils.HMM_constants = HMM_constants
        import HMM_backup_dyn as hmm
        import PredictiveControllers as pc
    cons = HMM_constants(s1=2, s2=3, c2=0.5, tran_diag=0.3, alpha=1, R=1.2, am=6.0, rm=0.3, J_c=20, s_c=1, ylb=0., yub=7.2,
                         L=4, W=2.5, col_alpha=5, Kpsi=0.1)
    return hmm, pc, cons


class _legacy_numpy_reshape:
    """`np.reshape(b0, -1, 1)` (PredictiveControllers.py:121) passes the integer 1 as `order`: the numpy of the reference's
    era converted it through NPY_ORDER (0 = C, 1 = Fortran), numpy 2 raises.  The old conversion is restored for the
    duration of the reference call; nothing in the reference is touched."""

    def __enter__(self):
        self._orig = np.reshape

        def reshape(a, *args, **kw):
            if len(args) == 2 and isinstance(args[1], (int, np.integer)):
                return self._orig(a, args[0], order={0: "C", 1: "F"}[int(args[1])])
            return self._orig(a, *args, **kw)

        np.reshape = reshape
        return self

    def __exit__(self, *exc):
        np.reshape = self._orig


def run_belief_mpc(name, x0, others, b0, ydes, vdes, steps, N=10, M=2):
    """Belief-state MPC: the UNMODIFIED `PredictiveControllers.MPC` (:56-340) on the UNMODIFIED
    `HMM_backup_dyn.PredictiveModel` (:177-276) with `Init_MPC.initMPCParams` (:7-34), closed loop: the ego applies uPred[0],
    the other agents keep 'maintain', the belief follows the model's own transition.  Recorded per step: inputs, xbackup, the
    assembled QP (through the osqp stand-in), its optimum, xLin/uLin, and the model's linearisation at every stage."""
    print("belief-state MPC fixture", name)
    hmm, pc, cons = _hmm_module()
    m, dt = 2, 0.1
    backupcons = [lambda s: hmm.backup_maintain(s, cons), lambda s: hmm.backup_brake(s, cons)]
    model = hmm.PredictiveModel(4, 2, M, backupcons, dt, cons)
    par = Init_MPC.initMPCParams(4, 2, N, M, m, ydes, vdes, 6.0, 0.3, 2, cons.W)
    mpc = pc.MPC(par, model)
    store = {"meta_N": np.array(N), "meta_M": np.array(M), "meta_m": np.array(m), "meta_steps": np.array(steps),
             "meta_ydes": np.array(ydes), "meta_vdes": np.array(vdes), "meta_dt": np.array(dt)}
    x = np.array(x0, dtype=float)
    Z = np.array(others, dtype=float)
    b = np.array(b0, dtype=float)
    xref = np.array([0, ydes, vdes, 0.])
    f0 = np.array([20., 0, 0, 0])
    for k in range(steps):
        # xbackup exactly as the reference's own caller builds it (Highway_env.py:135-142): per other agent and policy the
        # module-level rollout from the CURRENT state, first N+1 states, flattened time-major
        xbackup = np.empty([0, (N + 1) * 4])
        for zz in Z:
            for j in range(m):
                tt, xx, uu, QQ, Qt = hmm.generate_backup_traj(zz, backupcons[j], lambda s, t: t > dt * N + 2, f0, dt, True)
                xbackup = np.vstack((xbackup, np.reshape(np.array(xx[0:N + 1]), [1, -1])))
        ulin_before = None if mpc.uLin is None else np.array(mpc.uLin)
        with _legacy_numpy_reshape():
            mpc.solve(x, b, xbackup, xref)
        lp = osqp.last_problem
        assert lp["ok"], "oracle QP solve did not certify"
        pre = "s%d_" % k
        store[pre + "x0"], store[pre + "b0"], store[pre + "xbackup"], store[pre + "xref"] = x.copy(), b.copy(), xbackup, xref
        for nm, mat in (("P", sp.triu(lp["P"])), ("A", lp["A"])):
            r, c, v, shp = coo(mat)
            store[pre + nm + "_r"], store[pre + nm + "_c"], store[pre + nm + "_v"], store[pre + nm + "_shape"] = r, c, v, shp
        store[pre + "q"], store[pre + "l"], store[pre + "u"] = lp["q"], lp["l"], lp["u"]
        store[pre + "sol"] = lp["x"]
        store[pre + "objective"] = np.array(lp["cert"]["objective"])
        store[pre + "xPred"], store[pre + "uPred"] = np.array(mpc.xPred), np.array(mpc.uPred)
        store[pre + "A"], store[pre + "B"], store[pre + "C"] = np.array(mpc.A), np.array(mpc.B), np.array(mpc.C)
        store[pre + "h0"] = np.array([[np.ravel(hh) for hh in h0i] for h0i in mpc.h0])          # (N, M, m)
        store[pre + "Jh"] = np.array([[np.array(jj) for jj in jhi] for jhi in mpc.Jh])           # (N, M, m, n)
        if ulin_before is not None:
            store[pre + "uLin_before"] = ulin_before
        print("   step %d: obj %.6f  u0 %s  rows %d  kkt %.1e/%.1e" % (k, lp["cert"]["objective"],
              np.array2string(mpc.uPred[0], precision=6), lp["A"].shape[0], lp["cert"]["primal"], lp["cert"]["dual"]))
        # plants: ego under its first input, the others under 'maintain'; belief through the model's own transition
        xb = np.append(x, np.reshape(b, [-1, 1]))
        xbp = np.array(model.xbpsym(xb, mpc.uPred[0], xbackup[:, 0:4])).ravel()
        x = xbp[:4]
        b = np.reshape(xbp[4:], b.shape)
        Z = np.array([zz + dt * np.array([zz[2] * np.cos(zz[3]), zz[2] * np.sin(zz[3]), 0., -cons.Kpsi * zz[3]]) for zz in Z])
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **store)


if __name__ == "__main__":
    which = sys.argv[1:] or ["models", "hw_default", "hw_close", "hw_sweep", "robust", "cvar", "belief", "quad", "hmm"]
    if "models" in which:
        model_function_vectors()
    if "hw_default" in which:
        # main_branch.py:24-48 constants + Highway_env_branch.py:67 initial states; xRef as in the survey probe
        run_highway("highway_branch_default", "BranchMPC", ["maintain", "brake", "lc"], 2,
                    [0, 1.8, 20, 0], [5, 5.4, 20, 0], [0, 1.8, 26.5, 0], steps=4)
    if "hw_close" in which:
        # obstacle just ahead in the same lane: collision rows and lane rows become active
        run_highway("highway_branch_close", "BranchMPC", ["maintain", "brake", "lc"], 2,
                    [0, 1.9, 22, 0.02], [9, 1.8, 17, 0], [0, 1.8, 25, 0], steps=3, lc_target=(0.5, 5.4, 17., 0.))
    if "hw_sweep" in which:
        run_highway("highway_branch_m2_nb3", "BranchMPC", ["maintain", "brake"], 3,
                    [0, 5.4, 18, -0.01], [12, 5.6, 15, 0], [0, 5.4, 20, 0], steps=2)
        run_highway("highway_branch_m3_nb1", "BranchMPC", ["maintain", "brake", "lc"], 1,
                    [0, 9.0, 24, 0.0], [-6, 5.4, 25, 0], [0, 9.0, 22, 0], steps=2, lc_target=(0.5, 9.0, 25., 0.))
    if "robust" in which:
        run_highway("highway_robust_default", "robustMPC", ["maintain", "brake", "lc"], 2,
                    [0, 1.8, 20, 0], [5, 5.4, 20, 0], [0, 1.8, 26.5, 0], steps=3)
    if "cvar" in which:
        # main_branch.py:48 (ralpha = 0.9) on the default scene, a close-obstacle scene, and ralpha = 0.1 (the value of sim_merge, :92)
        run_highway_cvar("highway_cvar_default", ["maintain", "brake", "lc"], 2, [0, 1.8, 20, 0], [5, 5.4, 20, 0],
                         [0, 1.8, 26.5, 0], steps=4, ralpha=0.9)
        run_highway_cvar("highway_cvar_close", ["maintain", "brake", "lc"], 2, [0, 1.9, 22, 0.02], [9, 1.8, 17, 0],
                         [0, 1.8, 25, 0], steps=3, ralpha=0.9, lc_target=(0.5, 5.4, 17., 0.))
        run_highway_cvar("highway_cvar_alpha01", ["maintain", "brake", "lc"], 2, [0, 5.4, 18, -0.01], [12, 5.6, 15, 0],
                         [0, 5.4, 20, 0], steps=3, ralpha=0.1)
        run_highway_cvar("highway_cvar_m2_nb1", ["maintain", "brake"], 1, [0, 1.8, 20, 0], [8, 1.9, 16, 0],
                         [0, 1.8, 24, 0], steps=3, ralpha=0.1)
    if "belief" in which:
        run_belief_mpc("belief_mpc_default", [0, 1.8, 20, 0], [[12, 1.8, 16, 0], [-8, 5.4, 22, 0]], [[0.5, 0.5], [0.5, 0.5]],
                       1.8, 22.0, steps=3)
        run_belief_mpc("belief_mpc_close", [0, 1.9, 21, 0.01], [[9, 1.8, 14, 0], [6, 5.2, 20, -0.02]], [[0.7, 0.3], [0.2, 0.8]],
                       1.8, 24.0, steps=3)
    if "quad" in which:
        run_quadruped("quadruped_prox_default", [0, 0, 0], [2, 0.3, np.pi], [5., 5., 0.], steps=3)
    if "hmm" in which:
        hmm_vectors()


def run_highway_env(name, steps, x_ego=None, x_obs=None, N_lane=4):
    """Closed loop of the reference's OWN environment (Highway_env_branch.Highway_env.step + the collision check of
    Highway_sim, :83-184, :393-445) around the reference BranchMPC: lane bookkeeping, lane-change target updates, the
    obstacle's arg-max policy (numeric veh_col / lane_bdry_h branches), the xRef rule and both Euler plants."""
    print("highway env fixture", name)
    with refenv.reference_imports():
        import Highway_env_branch as henv
    np.random.seed(11)           # the coin flips only touch desired_x, which nothing reads (Highway_env_branch.py:121-133)
    cons = highway_cons()
    lc_target = np.array([0.5, 1.8, 15., 0.])      # main_branch.py:35
    model = hw.PredictiveModel(4, 2, 8, highway_policies(["maintain", "brake", "lc"], cons, lc_target), 0.1, cons)
    par = Init_MPC.initBranchMPC(4, 2, 8, 2, lc_target, 6.0, 0.3, N_lane, cons.W)
    mpc = MPC_branch.BranchMPC(par, model)
    env = henv.Highway_env(2, mpc, N_lane)
    if x_ego is not None:
        env.veh_set[0].state = np.array(x_ego, dtype=float)
    if x_obs is not None:
        env.veh_set[1].state = np.array(x_obs, dtype=float)
    store = {"meta_steps": np.array(steps), "meta_N_lane": np.array(N_lane), "x_init": env.veh_set[0].state.copy(),
             "z_init": env.veh_set[1].state.copy()}
    rec = {k: [] for k in ("x", "z", "lane", "backupidx", "xref", "u_ego", "u_obs", "lc_target", "collision")}
    solve = mpc.solve
    seen = {}

    def spy(x, z, xRef=None):
        seen["xref"] = np.array(xRef, dtype=float)
        return solve(x, z, xRef)

    mpc.solve = spy
    collision = False
    for t in range(steps):
        a, b = env.veh_set
        dis = max(abs(a.state[0] - b.state[0]) - 0.5 * (a.v_length + b.v_length),
                  abs(a.state[1] - b.state[1]) - 0.5 * (a.v_width + b.v_width))     # Highway_sim :421-429
        collision = collision or dis < 0
        u_set, x_set, *_ = env.step(t)
        # the lane-change target the MODEL's compiled graphs now use: first Euler step of the lc rollout from the zero
        # state is dt*[0, 0, 0.8558 v_t, 0.3162 y_t] (psi_t = 0 in every target the env builds).  The lambda itself cannot
        # be probed: it closes over the local `xRef` of step(), which is rebound to the MPC reference afterwards (:168).
        first = np.array(model.zpred_eval(np.zeros(4)))[0, 8:12]
        rec["lc_target"].append(np.array([0.0, first[3] / (0.1 * 0.3162), first[2] / (0.1 * 0.8558), 0.0]))
        rec["x"].append(np.array(x_set[0])); rec["z"].append(np.array(x_set[1]))
        rec["lane"].append([env.veh_set[0].laneidx, env.veh_set[1].laneidx])
        rec["backupidx"].append(env.veh_set[1].backupidx)
        rec["xref"].append(seen["xref"]); rec["u_ego"].append(np.array(u_set[0])); rec["u_obs"].append(np.array(u_set[1]))
        rec["collision"].append(collision)
        print("   t %d  x %s  z %s  obs policy %d  xref %s" % (t, np.round(x_set[0], 4), np.round(x_set[1], 4),
                                                             env.veh_set[1].backupidx, np.round(seen["xref"], 3)))
    for k, v in rec.items():
        store[k] = np.array(v)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **store)


def run_highway_merge(name="highway_merge_default", T=6.0):
    """`sim_merge()` of the UNMODIFIED main_branch.py (:53-88): N = 40, NB = 1, two `PredictiveModel_merge` (lookup-table
    policies for the ramp), `BranchMPC_CVaR(ralpha=0.1)` called as `solve(x, z, xRef, S, Fx=None, bx=bx)` by
    `Highway_env_merge.step` (Highway_env_branch.py:324-380) inside `Highway_sim` - all of it the reference's own code; only
    the plotting call is replaced by a no-op and the duration can be shortened.  Every MPC call is recorded (inputs, state
    transform, bounds, warm start the step linearised about, plan, objective, tree) together with the closed-loop records
    `Highway_sim` returns."""
    print("highway merge fixture", name)
    with refenv.reference_imports():
        import Highway_env_branch as henv
        import main_branch
    calls = []
    orig_solve = MPC_branch.BranchMPC_CVaR.solve

    def spy(self, x, z, xRef=None, S=None, Fx=None, bx=None):
        before = None if self.uLin is None or self.BT is None else np.array(self.uLin)
        pbest = None if self.BT is None else np.array([int(np.argmax(b.p)) if b.p is not None else 0 for b in self.ndx])
        assert Fx is None
        out = orig_solve(self, x, z, xRef, S, Fx, bx)
        lp = ecos.last_problem
        assert self.feasible and lp["info"]["exitFlag"] == 0, "oracle cone solve did not converge"
        tab, w, p, xbar, zbar, ubar = tree_tables(self)
        calls.append(dict(x0=np.array(x, dtype=float), z0=np.array(z, dtype=float), xref=np.array(xRef, dtype=float),
                          S=np.array(S, dtype=float), bx=np.asarray(bx, dtype=float).reshape(-1), uLin_before=before,
                          pbest_before=pbest, xPred=np.array(self.xPred), uPred=np.array(self.uPred),
                          objective=np.array(lp["info"]["pcost"]),
                          cert=np.array([lp["info"]["gap"], lp["info"]["pres"], lp["info"]["dres"]]),
                          w=w, p=p, xbar=xbar, zbar=zbar, ubar=ubar, tree=tab, mpc=self))
        print("   step %d: J %.8f  u0 %s  gap %.1e pres %.1e dres %.1e (%d it)" % (
            len(calls) - 1, lp["info"]["pcost"], np.array2string(self.uPred[0], precision=6), lp["info"]["gap"],
            lp["info"]["pres"], lp["info"]["dres"], lp["info"]["iter"]))
        return out

    sim_out = {}
    orig_sim = henv.Highway_sim

    def sim(env, T_):
        sim_out["env"] = env
        sim_out["x_init"] = np.array([v.state for v in env.veh_set])
        sim_out["rec"] = orig_sim(env, T)
        return sim_out["rec"]

    MPC_branch.BranchMPC_CVaR.solve = spy
    henv.Highway_sim = sim
    animate = henv.animate_scenario
    henv.animate_scenario = lambda *a, **k: None
    try:
        main_branch.sim_merge()
    finally:
        MPC_branch.BranchMPC_CVaR.solve = orig_solve
        henv.Highway_sim = orig_sim
        henv.animate_scenario = animate
    env = sim_out["env"]
    mpc = calls[0]["mpc"]
    state_rec, input_rec, backup_rec, backup_choice_rec, xPred_rec, zPred_rec, branch_w_rec, collision = sim_out["rec"]
    store = {"meta_ctrl": np.array("BranchMPC_CVaR"), "meta_NB": np.array(mpc.NB), "meta_N": np.array(mpc.N),
             "meta_steps": np.array(len(calls)), "meta_ralpha": np.array(mpc.ralpha), "meta_dt": np.array(env.dt),
             "meta_v0": np.array(henv.v0), "meta_N_lane": np.array(env.N_lane), "meta_merge_lane": np.array(env.merge_lane),
             "meta_merge_s": np.array(env.merge_s), "meta_merge_R": np.array(env.merge_R),
             "meta_merge_side": np.array(env.merge_side), "meta_am": np.array(env.cons.am), "meta_rm": np.array(env.cons.rm),
             "meta_bx": np.asarray(mpc.param.bx, dtype=float).reshape(-1), "meta_Q": np.array(mpc.Q), "meta_R": np.array(mpc.R),
             "meta_dR": np.array(mpc.dR), "meta_Qslack": np.array(mpc.Qslack),
             "table_X": np.array(env.merge_lane_ref_X), "table_Y": np.array(env.merge_lane_ref_Y),
             "table_psi": np.array(env.merge_lane_ref_psi), "x_init": sim_out["x_init"],
             "state_rec": np.array(state_rec), "input_rec": np.array(input_rec), "collision": np.array(bool(collision)),
             "backup_choice_rec": np.array(backup_choice_rec, dtype=np.int64),
             "backup_rec_ego": np.array([np.array(b) for b in backup_rec[0]]),
             "backup_rec_obs": np.array([np.array(b) for b in backup_rec[1]])}
    for k, c in enumerate(calls):
        pre = "s%d_" % k
        for key, val in c.items():
            if key != "mpc" and val is not None:
                store[pre + key] = val
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **store)


def merge_model_vectors():
    """Point-wise values of the reference's PredictiveModel_merge (highway_branch_dyn.py:400-502) at seeded random inputs, for
    both models sim_merge builds (main_branch.py:82-85): plain policies, and the ramp policies that steer along the heading
    lookup table."""
    print("merge model-function fixture")
    with refenv.reference_imports():
        import Highway_env_branch as henv
        from casadi import interpolant
    rng = np.random.default_rng(20241019)
    cons = rutils.Branch_constants(s1=2, s2=3, c2=0.5, tran_diag=0.3, alpha=1, R=1.2, am=7.0, rm=0.3, J_c=20, s_c=1, ylb=0.,
                                   yub=7.2, L=4, W=2.5, col_alpha=5, Kpsi=0.1)
    X1, X2, Y1, Y2, P1, P2 = henv.merge_geometry(2, 1, 50, 300, 0)
    gx, gy, gpsi = np.append(X1, X2), np.append(Y1, Y2), np.append(P1, P2)
    refY = interpolant('refY', 'linear', [gx], gy)
    refpsi = interpolant('refpsi', 'linear', [gx], gpsi)
    v0, N = float(henv.v0), 40
    normal = [lambda x: hw.backup_maintain_trackV(x, cons, v0), lambda x: hw.backup_brake(x, cons)]
    ramp = [lambda x: hw.backup_maintain_trackV(x, cons, v0, refpsi), lambda x: hw.backup_brake(x, cons, refpsi)]
    models = [hw.PredictiveModel_merge(4, 2, N, normal, 0.1, cons, (refY, refpsi), laneID=0, N_lane1=2, N_lane2=1),
              hw.PredictiveModel_merge(4, 2, N, ramp, 0.1, cons, (refY, refpsi), laneID=1, N_lane1=2, N_lane2=1)]
    K = 10
    X = np.column_stack([rng.uniform(5, 80, K), rng.uniform(1, 14, K), rng.uniform(8, 25, K), rng.normal(-0.1, 0.1, K)])
    Z = np.column_stack([X[:, 0] + rng.uniform(-12, 12, K), rng.uniform(1, 14, K), rng.uniform(8, 25, K), rng.normal(-0.05, 0.1, K)])
    U = np.column_stack([rng.uniform(-7, 7, K), rng.uniform(-0.3, 0.3, K)])
    out = {"X": X, "Z": Z, "U": U, "table_X": gx, "table_Y": gy, "table_psi": gpsi, "v0": np.array(v0), "N": np.array(N)}
    for mi, model in enumerate(models):
        A, B, C, XP, ZP, P, H, DH, X1P, U0 = [], [], [], [], [], [], [], [], [], []
        for k in range(K):
            a, b, c, xp = model.dyn_linearization(X[k], U[k])
            A.append(a); B.append(b); C.append(c); XP.append(xp)
            ZP.append(model.zpred_eval(Z[k]))
            p, _ = model.branch_eval(X[k], Z[k])
            P.append(p)
            h, dh = model.col_eval(X[k], Z[k])
            H.append(h); DH.append(dh)
            x1, u0 = model.xpred_eval(X[k])
            X1P.append(np.array(x1)); U0.append(np.array(u0).reshape(-1))
        pre = "m%d_" % mi
        out.update({pre + "A": np.array(A), pre + "B": np.array(B), pre + "C": np.array(C), pre + "xp": np.array(XP),
                    pre + "zpred": np.array(ZP), pre + "p": np.array(P), pre + "hlin": np.array(H), pre + "dh": np.array(DH),
                    pre + "xpred": np.array(X1P), pre + "u0": np.array(U0)})
    np.savez_compressed(os.path.join(HERE, "merge_model_functions.npz"), **out)


if __name__ == "__main__" and "merge_models" in sys.argv[1:]:
    merge_model_vectors()

if __name__ == "__main__" and "merge" in sys.argv[1:]:
    run_highway_merge()

if __name__ == "__main__" and ("env" in sys.argv[1:] or len(sys.argv) == 1):
    run_highway_env("highway_env_default", steps=30)
    run_highway_env("highway_env_overtake", steps=24, x_ego=[2.0, 5.5, 24.0, 0.0], x_obs=[14.0, 5.4, 17.0, 0.0])
