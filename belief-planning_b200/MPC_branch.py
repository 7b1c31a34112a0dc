"""Drop-in for the reference's `MPC_branch.py` controllers, solved in batches on the B200.

`BranchMPC(mpcParameters, predictiveModel)` and `solve(x, z, xRef=None)` keep the reference's interface
(MPC_branch.py:883-927, :1171-1229): after a solve the object carries `uPred (totalu, d)`, `xPred (totalx, n)`,
`xLin`, `uLin`, `feasible`, `solverTime`, `OldInput`, `timeStep`, and `BT2array()` returns the tree for plotting.  The
same call accepts a batch: x, z (and xRef) of shape (B, n) solve B independent episodes in one kernel launch, and the
result attributes gain a leading batch axis.  There is no CPU path: the constructor loads libbranchmpc.so.

What is replaced underneath (per solve): inittree/updatetree, buildCost, buildEqConstr, buildIneqConstr,
updateIneqConstr, the dense->CSC conversions, OSQP setup+solve+polish and unpackSolution, all inside one persistent
CUDA kernel (belief-planning_b200/csrc).
"""
import datetime
from dataclasses import dataclass, field

import numpy as np

from _bmpc import abi, batch, config
from utils import PythonMsg

__all__ = ["BranchMPCParams", "BranchTree", "BranchMPC", "BranchMPCProx", "robustMPC", "BranchMPC_CVaR"]


def _f():
    return field(default=None)


@dataclass
class BranchMPCParams(PythonMsg):
    """MPC_branch.py:27-54.  Note Qslack = [quadratic, linear] as the code uses it (:1105-1106)."""
    n: int = _f()
    d: int = _f()
    NB: int = _f()
    N: int = _f()
    A: np.ndarray = _f()
    B: np.ndarray = _f()
    Q: np.ndarray = _f()
    R: np.ndarray = _f()
    Qf: np.ndarray = _f()
    dR: np.ndarray = _f()
    Qslack: float = _f()
    Fx: np.ndarray = _f()
    bx: np.ndarray = _f()
    Fu: np.ndarray = _f()
    bu: np.ndarray = _f()
    xRef: np.ndarray = _f()
    slacks: bool = field(default=True)
    timeVarying: bool = field(default=False)

    def __post_init__(self):
        if self.Qf is None:
            self.Qf = self.Q
        if self.dR is None:
            self.dR = np.zeros(self.d)
        if self.xRef is None:
            self.xRef = np.zeros(self.n)


class BranchTree:
    """One branch of the scenario/trajectory tree as the reference exposes it (MPC_branch.py:65-78); built on demand
    from the flat device results for plotting and debugging."""

    def __init__(self, xtraj, ztraj, utraj, w, depth=0):
        self.xtraj, self.ztraj, self.utraj = xtraj, ztraj, utraj
        self.dynmatr = [None] * xtraj.shape[0]
        self.w = w
        self.children = []
        self.depth = depth
        self.p = None
        self.dp = None
        self.J = 0

    def addchild(self, BT):
        self.children.append(BT)


class _BatchedController:
    controller_kind = abi.CTRL_BRANCH

    def __init__(self, mpcParameters, predictiveModel, **solver_knobs):
        p = mpcParameters
        self.N, self.NB, self.n, self.d = p.N, p.NB, p.n, p.d
        self.Q, self.Qf, self.R, self.dR, self.Qslack = p.Q, p.Qf, p.R, p.dR, p.Qslack
        self.Fx, self.Fu, self.bx, self.bu = p.Fx, p.Fu, p.bx, p.bu
        self.xRef = p.xRef
        self.m = predictiveModel.m
        self.slacks, self.timeVarying = p.slacks, p.timeVarying
        self.predictiveModel = predictiveModel
        if not p.slacks:
            raise NotImplementedError("the reference only ever runs with slacks=True (Init_MPC.py:71,:93)")
        self._knobs = solver_knobs
        self._solver = None
        self._capacity = 0
        self.BT = None
        self.xPred = self.uPred = self.xLin = None
        self.zPred = self.branch_w = self.branch_p = None
        self.OldInput = np.zeros(self.d)
        self.feasible = 0
        self.status = None
        self.solverTime = datetime.timedelta(0)
        self.linearizationTime = datetime.timedelta(0)
        self.timeStep = 0
        abi.load_library()          # fail now, loudly, if the CUDA library is missing

    # -- handle management ----------------------------------------------------------------------------------
    def _make_solver(self, capacity):
        model = self.predictiveModel
        bx = np.squeeze(np.asarray(self.bx, dtype=float)).reshape(-1)      # the reference stores a 1-tuple (Init_MPC.py:48)
        bu = np.squeeze(np.asarray(self.bu, dtype=float)).reshape(-1)
        cfg = config.make_config(model.spec(), self.n, self.d, self.N, self.NB, self.Q, self.R, self.Fx, bx, self.Fu, bu,
                                 self.Qslack, controller=self.controller_kind, Qf=self.Qf, dR=self.dR,
                                 batch_capacity=capacity, **self._knobs)
        self._kinds = [dd.kind for dd in model.descriptors]
        solver = batch.BatchedBranchMPC(cfg)
        table = getattr(model, "lookup_table", None)
        if table is not None:
            solver.set_lookup_table(table.xs, table.ys)
        self.totalx, self.totalu = solver.totalx, solver.totalu
        topo = solver.topology()
        self.ndx = {int(r[0]): int(r[2]) for r in topo}
        self.ndu = {int(r[0]): int(r[3]) for r in topo}
        self._topo = topo
        return solver

    def _ensure_solver(self, B):
        """The batched handle for (at least) B episodes with the model's current policy kinds."""
        model = self.predictiveModel
        if self._solver is not None and (B > self._capacity or [dd.kind for dd in model.descriptors] != self._kinds):
            self._solver.close()
            self._solver = None
        if self._solver is None:
            self._solver = self._make_solver(B)
            self._capacity = B
        return self._solver

    def _absorb(self, r, single):
        """Result arrays of a batched solve -> the attributes the reference leaves behind (MPC_branch.py:1204-1229).
        Only solved problems are adopted (feasible = 1 for the solver's 'solved' alone, :1269-1272); an episode whose
        solve failed keeps its previous plan (:1224), per episode in a batch.
        The arrays are views into the library's pinned result blocks (no copy: a 16384-episode step returns 190 MB): like
        the reference's attributes they describe the LAST solve; they stay intact during the next call to solve() and are
        recycled by the one after it - copy what has to live longer."""
        ok = r["status"] <= abi.STATUS_CONVERGED
        self.status = r["status"][0] if single else r["status"]
        self.feasible = int(ok[0]) if single else ok.astype(int)
        self.iterations = r["iters"][0] if single else r["iters"]
        B = len(ok)
        names = ("xPred", "uPred", "xLin", "zPred", "branch_w", "branch_p", "objective")
        prev = getattr(self, "_res", None)
        if ok.all() or prev is None or prev["uPred"].shape[0] != B:
            if not ok.all() and single:
                raise RuntimeError("the first solve failed (status %d): there is no previous plan to keep" % r["status"][0])
            self._res = {k: r[k] for k in names}
        else:
            merged = {k: np.array(prev[k]) for k in names}       # failed episodes keep their previous rows
            for k in names:
                merged[k][ok] = r[k][ok]
            self._res = merged
        res = self._res
        pick = (lambda a: a[0]) if single else (lambda a: a)
        self.xPred, self.uPred = pick(res["xPred"]), pick(res["uPred"])
        self.xLin = self.xPred
        self._xbar, self.zPred = pick(res["xLin"]), pick(res["zPred"])
        self.branch_w, self.branch_p = pick(res["branch_w"]), pick(res["branch_p"])
        self.objective = pick(res["objective"])
        self.OldInput = self.uPred[0, :] if single else self.uPred[:, 0, :]
        self.timeStep += 1
        self.BT = True                              # "a tree exists": later solves are updatetree solves
        return self.OldInput

    @property
    def uLin(self):
        """uPred with its last row repeated (unpackSolution, MPC_branch.py:1228-1229); built on access."""
        if self.uPred is None:
            return None
        return np.concatenate([self.uPred, self.uPred[..., -1:, :]], axis=-2)

    @uLin.setter
    def uLin(self, value):
        pass                                          # the warm start lives on the device (bmpc_set_state)

    def solve(self, x, z, xRef=None):
        """Computes the control action(s).  x, z: (n,) or (B, n); xRef: (n,) or (B, n) or None (keep the previous)."""
        if xRef is not None:
            self.xRef = xRef
        x = np.asarray(x, dtype=float)
        single = x.ndim == 1
        X = np.atleast_2d(x)
        Z = np.atleast_2d(np.asarray(z, dtype=float))
        B = X.shape[0]
        R = np.broadcast_to(np.atleast_2d(np.asarray(self.xRef, dtype=float)), (B, self.n))
        model = self.predictiveModel
        solver = self._ensure_solver(B)
        pp = np.broadcast_to(model.policy_params(), (B, self.m, 4))        # update_backup() -> new per-episode parameters
        t0 = datetime.datetime.now()
        r = solver.solve_host_views(X, Z, R, pp)
        self.solverTime = datetime.datetime.now() - t0
        return self._absorb(r, single)

    def reset(self, episode_ids=None):
        """Forget the warm-start state of the given episodes (all by default): their next solve is an inittree solve."""
        if self._solver is not None:
            self._solver.reset(episode_ids)

    # -- plotting helper --------------------------------------------------------------------------------------
    def BT2array(self, episode=0):
        """(xtraj, ztraj, utraj, branch_w): per non-root branch, BFS order, the parent's last node stacked on the
        branch's linearisation trajectory (MPC_branch.py:1231-1246)."""
        if self.uPred is None:
            raise RuntimeError("solve() has not been called")
        if self.controller_kind == abi.CTRL_ROBUST:
            # robustMPC.BT2array (:1385-1395): one ego trajectory; the obstacle tree is not returned by the kernel
            pick = (lambda a: a[episode]) if np.ndim(self.uPred) == 3 else (lambda a: a)
            return [pick(self.xPred)], [], [pick(self.uPred)], []
        batched = np.ndim(self.uPred) == 3
        sel = (lambda a: a[episode]) if batched else (lambda a: a)
        xbar, zbar, w = sel(self._xbar), sel(self.zPred), sel(self.branch_w)
        # linearisation inputs = previous uLin shifted; the reference returns utraj of the tree it linearised about.
        ubar = sel(self.uPred)
        xs, zs, us, ws = [], [], [], []
        topo = self._topo
        length = lambda b: 1 if b == 0 else self.N
        for row in topo[1:]:
            b, par = int(row[0]), int(row[4])
            last = self.ndu[par] + length(par) - 1
            sl = slice(self.ndu[b], self.ndu[b] + self.N)
            xs.append(np.vstack((xbar[last], xbar[sl])))
            zs.append(np.vstack((zbar[last], zbar[sl])))
            us.append(np.vstack((ubar[last], ubar[sl])))
            ws.append(float(w[b]))
        return xs, zs, us, ws


class BranchMPC(_BatchedController):
    """The effective `BranchMPC` of the reference (its second definition, MPC_branch.py:881-1274)."""
    controller_kind = abi.CTRL_BRANCH


class BranchMPCProx(_BatchedController):
    """MPC_branch.BranchMPCProx (:82-487): prox weight dQ = 3Q and input-rate costs."""
    controller_kind = abi.CTRL_PROX


class robustMPC(_BatchedController):
    """MPC_branch.robustMPC (:1275-1595): one trajectory avoiding every obstacle node of the scenario tree."""
    controller_kind = abi.CTRL_ROBUST


class BranchMPC_CVaR(_BatchedController):
    """MPC_branch.BranchMPC_CVaR (:1598-2152), the controller main_branch.py:48 builds: nested-CVaR objective over the
    scenario tree (the reference hands a second-order-cone program to ECOS).  Same constructor and `solve` signature as the
    reference (`ralpha` :1601, `solve(x, z, xRef, S, Fx, bx)` :2043); on the device the cone program is solved as a
    cutting-plane loop over the risk multipliers whose inner problems are branch-weighted tree QPs (csrc/bmpc_solver.h).
    The state transformation `S` and the per-call bounds `bx` of the merge scene (:2054-2059) need a model built as
    `PredictiveModel_merge` (BMPC_MODEL_MERGE on the device); a different `Fx` per call is not built."""
    controller_kind = abi.CTRL_CVAR

    def __init__(self, mpcParameters, predictiveModel, ralpha=0.1, S=None, **solver_knobs):
        super().__init__(mpcParameters, predictiveModel, cvar_alpha=float(ralpha), **solver_knobs)
        self.ralpha = ralpha
        self.S = S
        self.param = mpcParameters
        self.psimax = np.squeeze(np.asarray(mpcParameters.bx, dtype=float)).reshape(-1)[2]     # bx[0][2][0], :1622

    def _transformed(self):
        return self.predictiveModel.spec().kind == abi.MODEL_MERGE

    def solve(self, x, z, xRef=None, S=None, Fx=None, bx=None):
        """solve(x, z, xRef, S, Fx, bx) (:2043-2059).  Batched: x, z (B, n); S (n, n) or (B, n, n); bx (4,) or (B, 4)."""
        if Fx is not None and not np.array_equal(np.squeeze(np.asarray(Fx, dtype=float)),
                                                 np.squeeze(np.asarray(self.Fx, dtype=float))):
            raise NotImplementedError("a different Fx per call is not built (the reference's callers pass Fx=None)")
        if not self._transformed():
            if S is not None:
                raise NotImplementedError("a state transformation S needs a model built as PredictiveModel_merge")
            if bx is not None and not np.array_equal(np.squeeze(np.asarray(bx, dtype=float)),
                                                     np.squeeze(np.asarray(self.bx, dtype=float))):
                raise NotImplementedError("per-call state bounds need a model built as PredictiveModel_merge")
            return super().solve(x, z, xRef)
        if xRef is not None:
            self.xRef = xRef
        self.S = S                                    # :2054 (None = no transform, also for the collision-gradient rule)
        if bx is not None:
            self.bx = bx
        x = np.asarray(x, dtype=float)
        single = x.ndim == 1
        X = np.atleast_2d(x)
        Z = np.atleast_2d(np.asarray(z, dtype=float))
        B = X.shape[0]
        R = np.broadcast_to(np.atleast_2d(np.asarray(self.xRef, dtype=float)), (B, self.n))
        Sb = None if S is None else np.broadcast_to(np.asarray(S, dtype=float).reshape(-1, self.n, self.n), (B, self.n, self.n))
        bxb = np.broadcast_to(np.asarray(self.bx, dtype=float).reshape(-1, 4), (B, 4))
        bounds = np.stack([np.stack([-bxb[:, 1], bxb[:, 0]], axis=1), np.stack([-bxb[:, 3], bxb[:, 2]], axis=1)], axis=1)
        model = self.predictiveModel
        solver = self._ensure_solver(B)
        pp = np.broadcast_to(model.policy_params(), (B, self.m, 4))
        t0 = datetime.datetime.now()
        r = solver.solve_transformed_host_views(X, Z, R, Sb, bounds, pp)
        self.solverTime = datetime.datetime.now() - t0
        return self._absorb(r, single)
