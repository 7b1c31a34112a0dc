"""CPU restatement (float64) of the reference's Branch-MPC controllers.

TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this.  Pinned against the matrices the unmodified reference
assembles (tests/golden/*.npz, produced by tests/golden/make_golden.py).

What it follows (all in /root/reference/MPC_branch.py):
  tree numbering        : inittree      :928-981   (BFS, FIFO, children in policy order)
  time shift / rollout  : updatetree    :1024-1061
  cost                  : buildCost     :1064-1112 (effective BranchMPC)  /  :265-325 (BranchMPCProx)
  dynamics equalities   : buildEqConstr :984-1022
  inequalities + slacks : buildIneqConstr / updateIneqConstr :1114-1168
  solver call           : osqp_solve_qp :1248-1274  (OSQP is third-party and absent: the oracle
                          returns the exact optimum of the same QP, see oracle/qp_exact.py)
  unpack                : unpackSolution :1222-1229, solve tail :1207-1209

The data layout is deliberately not the reference's (flat BFS tables instead of an object tree).
"""
import numpy as np
import scipy.sparse as sp

from . import qp_exact


class TreeTopology:
    """BFS numbering of the scenario tree: branch 0 is the root (one node); every other branch has N nodes."""

    def __init__(self, m, NB, N):
        self.m, self.NB, self.N = int(m), int(NB), int(N)
        depth = [0]
        parent = [-1]
        policy = [-1]
        first_child = []
        queue = [0]
        ndx = [0]
        ndu = [0]
        cx = cu = 1
        while queue:
            b = queue.pop(0)
            if depth[b] < self.NB:
                first_child.append(len(depth))
                for i in range(self.m):
                    c = len(depth)
                    depth.append(depth[b] + 1)
                    parent.append(b)
                    policy.append(i)
                    ndx.append(cx)
                    ndu.append(cu)
                    cx += self.N + (1 if depth[c] == self.NB else 0)
                    cu += self.N
                    queue.append(c)
            else:
                first_child.append(-1)
        self.depth = np.array(depth)
        self.parent = np.array(parent)
        self.policy = np.array(policy)
        self.first_child = np.array(first_child)
        self.ndx = np.array(ndx)
        self.ndu = np.array(ndu)
        self.nbranch = len(depth)
        self.length = np.where(self.depth == 0, 1, self.N)
        self.is_leaf = self.depth == self.NB
        self.totalx = cx
        self.totalu = cu

    def children(self, b):
        fc = self.first_child[b]
        return [] if fc < 0 else list(range(fc, fc + self.m))

    def table(self):
        """(id, depth, ndx, ndu, parent) rows, the layout stored in the golden fixtures."""
        return np.column_stack([np.arange(self.nbranch), self.depth, self.ndx, self.ndu, self.parent]).astype(np.int64)


class BranchMPCOracle:
    """variant='branch' -> effective BranchMPC; variant='prox' -> BranchMPCProx."""

    def __init__(self, model, NB, Q, R, Fx, bx, Fu, bu, Qslack, xRef, dR=None, Qf=None, variant="branch"):
        self.model = model
        self.n, self.d, self.N, self.m = model.n, model.d, model.N, model.m
        self.NB = int(NB)
        self.topo = TreeTopology(self.m, self.NB, self.N)
        self.Q = np.asarray(Q, dtype=float)
        self.R = np.asarray(R, dtype=float)
        self.Qf = self.Q if Qf is None else np.asarray(Qf, dtype=float)
        self.dR = np.zeros(self.d) if dR is None else np.asarray(dR, dtype=float)
        self.Fx = np.asarray(Fx, dtype=float).reshape(-1, self.n)
        self.bx = np.asarray(bx, dtype=float).reshape(-1)
        self.Fu = np.asarray(Fu, dtype=float).reshape(-1, self.d)
        self.bu = np.asarray(bu, dtype=float).reshape(-1)
        self.Qslack = np.asarray(Qslack, dtype=float)
        self.xRef = np.asarray(xRef, dtype=float)
        self.variant = variant
        self.dQ_scale = 0.5 if variant == "branch" else 3.0
        self.OldInput = np.zeros(self.d)
        self.uLin = None
        self.p = None          # per-branch child probabilities from the last tree expansion
        self.timeStep = 0
        self.feasible = 0
        self.xPred = self.uPred = None

    # ------------------------------------------------------------------------------------------
    def expand_tree(self, x, z):
        """Linearisation trajectories, obstacle predictions, weights (inittree / updatetree)."""
        T = self.topo
        n, d, N, m = self.n, self.d, self.N, self.m
        first = self.uLin is None
        ubar = [np.zeros((T.length[b], d)) for b in range(T.nbranch)]
        if not first:
            for b in range(T.nbranch):
                l = T.length[b]
                ubar[b][: l - 1] = self.uLin[T.ndu[b] + 1: T.ndu[b] + l]
                if not T.is_leaf[b]:
                    best = T.first_child[b] + int(np.argmax(self.p[b]))
                    ubar[b][-1] = self.uLin[T.ndu[best]]
                else:
                    ubar[b][-1] = ubar[b][-2]
        xbar = [None] * T.nbranch
        zbar = [None] * T.nbranch
        lin = [None] * T.nbranch
        w = np.zeros(T.nbranch)
        p = np.full((T.nbranch, m), np.nan)
        xbar[0] = np.asarray(x, dtype=float).reshape(1, n).copy()
        zbar[0] = np.asarray(z, dtype=float).reshape(1, n).copy()
        w[0] = 1.0
        lin[0] = [self.model.dyn_linearization(xbar[0][0], ubar[0][0])[:3]]
        for b in range(T.nbranch):          # BFS order == index order
            if T.is_leaf[b]:
                continue
            zpred = self.model.zpred_eval(zbar[b][-1])
            p[b], _ = self.model.branch_eval(xbar[b][-1], zbar[b][-1])
            x_next = self.model.dyn_linearization(xbar[b][-1], ubar[b][-1])[3]
            for i, c in enumerate(T.children(b)):
                w[c] = w[b] * p[b, i]
                zbar[c] = zpred[:, i * n:(i + 1) * n]
                xs = np.zeros((N, n))
                xs[0] = x_next
                mats = []
                for t in range(N):
                    A, B, C, xp = self.model.dyn_linearization(xs[t], ubar[c][t])
                    mats.append((A, B, C))
                    if t < N - 1:
                        xs[t + 1] = xp
                xbar[c] = xs
                lin[c] = mats
        self.ubar, self.xbar, self.zbar, self.lin, self.w, self.p = ubar, xbar, zbar, lin, w, p

    # ------------------------------------------------------------------------------------------
    def assemble(self, x0):
        """The QP in OSQP form: min 1/2 v'Pv + q'v,  l <= A v <= u,  v = [x | u | slack]."""
        T = self.topo
        n, d = self.n, self.d
        Nc = self.Fx.shape[0] + 1
        nxv, nuv, nsv = T.totalx * n, T.totalu * d, T.totalx * Nc
        nvar = nxv + nuv + nsv
        H = sp.lil_matrix((nvar, nvar))
        q = np.zeros(nvar)
        dQ = self.Q * self.dQ_scale
        dRm = np.diag(self.dR)
        prox = self.variant == "prox"
        slackw = np.zeros(nsv)
        for b in range(T.nbranch):
            l, ix, iu, wb = T.length[b], T.ndx[b], T.ndu[b], self.w[b]
            for i in range(l):
                sx = slice((ix + i) * n, (ix + i + 1) * n)
                su = slice(nxv + (iu + i) * d, nxv + (iu + i + 1) * d)
                H[sx, sx] = (self.Q + dQ) * wb
                last = i == l - 1
                Qlin = self.Qf if (last and T.is_leaf[b] and not prox) else self.Q
                q[sx] = -2.0 * wb * (self.xRef @ Qlin + self.xbar[b][i] @ dQ)
                if not prox:
                    H[su, su] = wb * self.R
                else:
                    if not last:
                        su2 = slice(nxv + (iu + i + 1) * d, nxv + (iu + i + 2) * d)
                        H[su, su] = H[su, su] + wb * (self.R + dRm)
                        H[su, su2] = H[su, su2] - wb * dRm
                        H[su2, su] = H[su2, su] - wb * dRm
                        H[su2, su2] = H[su2, su2] + wb * dRm
                    elif not T.is_leaf[b]:
                        H[su, su] = H[su, su] + wb * (self.R + dRm)
                        for c in T.children(b):
                            sc = slice(nxv + T.ndu[c] * d, nxv + (T.ndu[c] + 1) * d)
                            H[su, sc] = H[su, sc] - self.w[c] * dRm
                            H[sc, su] = H[sc, su] - self.w[c] * dRm
                            H[sc, sc] = H[sc, sc] + self.w[c] * dRm
                    else:
                        H[su, su] = wb * self.R          # '=' in the reference (:303): rate term dropped
                slackw[(ix + i) * Nc:(ix + i + 1) * Nc] = wb
            if T.is_leaf[b]:
                sx = slice((ix + l) * n, (ix + l + 1) * n)
                H[sx, sx] = self.Qf * wb
                if prox:
                    q[sx] = -2.0 * wb * (self.xRef @ self.Qf)
        # root input: rate cost w.r.t. the previously applied input (scalar / broadcast quirks kept)
        q[nxv:nxv + d] = -2.0 * (self.OldInput @ self.dR)
        if prox:
            root = H[nxv:nxv + d, nxv:nxv + d].toarray() + self.dR          # row-broadcast (:312)
            H[nxv:nxv + d, nxv:nxv + d] = root
        if nsv:
            H[nxv + nuv:, nxv + nuv:] = self.Qslack[0] * sp.eye(nsv)
            q[nxv + nuv:] = self.Qslack[1] * slackw
        P = (2.0 * H).tocsc()

        # inequalities
        rows_Fx = sp.lil_matrix((nsv, nxv))
        bxt = np.zeros(nsv)
        for b in range(T.nbranch):
            for i in range(T.length[b]):
                h, dh = self.model.col_eval(self.xbar[b][i], self.zbar[b][i])
                k = T.ndx[b] + i
                rows_Fx[k * Nc, k * n:(k + 1) * n] = -dh
                if Nc > 1:
                    rows_Fx[k * Nc + 1:(k + 1) * Nc, k * n:(k + 1) * n] = self.Fx
                bxt[k * Nc] = h
                bxt[k * Nc + 1:(k + 1) * Nc] = self.bx
        Fu_all = sp.kron(sp.eye(T.totalu), sp.csr_matrix(self.Fu))
        ncu = Fu_all.shape[0]
        F = sp.bmat([[rows_Fx, None, -sp.eye(nsv)],
                     [None, Fu_all, sp.csr_matrix((ncu, nsv))],
                     [sp.csr_matrix((nsv, nxv)), sp.csr_matrix((nsv, nuv)), -sp.eye(nsv)]], format="csr")
        bF = np.concatenate([bxt, np.tile(self.bu, T.totalu), np.zeros(nsv)])

        # dynamics equalities  G v = E x0 + L
        G = sp.lil_matrix((nxv, nvar))
        G[:, :nxv] = sp.eye(nxv)
        beq = np.zeros(nxv)
        beq[:n] = x0
        for b in range(T.nbranch):
            l, ix, iu = T.length[b], T.ndx[b], T.ndu[b]
            for t in range(l):
                A, B, C = self.lin[b][t]
                if t < l - 1 or T.is_leaf[b]:
                    targets = [ix + t + 1]
                else:
                    targets = [T.ndx[c] for c in T.children(b)]
                for k in targets:
                    G[k * n:(k + 1) * n, (ix + t) * n:(ix + t + 1) * n] = -A
                    G[k * n:(k + 1) * n, nxv + (iu + t) * d: nxv + (iu + t + 1) * d] = -B
                    beq[k * n:(k + 1) * n] = C
        Aall = sp.vstack([F, G.tocsr()]).tocsc()
        lo = np.concatenate([np.full(F.shape[0], -np.inf), beq])
        hi = np.concatenate([bF, beq])
        return P, q, Aall, lo, hi

    # ------------------------------------------------------------------------------------------
    def solve(self, x, z, xRef=None, qp_solver=None):
        if xRef is not None:
            self.xRef = np.asarray(xRef, dtype=float)
        x = np.asarray(x, dtype=float)
        self.expand_tree(x, z)
        self.qp = self.assemble(x)
        if qp_solver is None:
            sol, ydual, info = qp_exact.solve_qp(*self.qp)
            ok = info["polished"]
        else:
            sol, ok = qp_solver(*self.qp)
        self.feasible = int(bool(ok))
        T = self.topo
        if self.feasible:
            nxv, nuv = T.totalx * self.n, T.totalu * self.d
            self.Solution = sol
            self.xPred = sol[:nxv].reshape(-1, self.n)
            self.uPred = sol[nxv:nxv + nuv].reshape(-1, self.d)
            self.uLin = np.vstack([self.uPred, self.uPred[-1]])
            self.objective = qp_exact.kkt_residuals(*self.qp, sol)["objective"]
        self.OldInput = self.uPred[0].copy()
        self.timeStep += 1
        return self.uPred[0]

    # stage-form objective of an (x,u) pair with the slacks eliminated (exact penalty) -----------
    def objective_xu(self, xPred, uPred):
        P, q, A, lo, hi = self.qp
        T = self.topo
        Nc = self.Fx.shape[0] + 1
        nxv, nuv, nsv = T.totalx * self.n, T.totalu * self.d, T.totalx * Nc
        v = np.zeros(nxv + nuv + nsv)
        v[:nxv] = np.asarray(xPred).reshape(-1)
        v[nxv:nxv + nuv] = np.asarray(uPred).reshape(-1)
        rows = A[:nsv, :nxv] @ v[:nxv]            # soft state rows
        v[nxv + nuv:] = np.maximum(rows - hi[:nsv], 0.0)
        return qp_exact.kkt_residuals(P, q, A, lo, hi, v)["objective"]
