"""CPU restatement (float64) of the reference's robustMPC: one ego trajectory that avoids EVERY obstacle node of the
scenario tree.  TEST INFRASTRUCTURE, pinned against the QPs the unmodified reference assembles
(tests/golden/highway_robust_default.npz).

Follows /root/reference/MPC_branch.py: robustMPC.__init__ :1279-1322, get_xLin :1326-1335, inittree/updatetree
:1336-1383 (obstacle nodes collected per time slot t = (depth-1) N + i + 1), solve :1397-1435 (LTV shift of the previous
QP solution :1429-1431), computeLTVdynamics :1438-1443, buildIneqConstr :1467-1510, buildEqConstr :1512-1538,
buildCost :1540-1569.
"""
import numpy as np
import scipy.sparse as sp

from . import qp_exact


class RobustMPCOracle:
    def __init__(self, model, NB, Q, R, Fx, bx, Fu, bu, Qslack, xRef, dR=None, Qf=None):
        self.model = model
        self.n, self.d, self.N, self.m = model.n, model.d, model.N, model.m
        self.NB = int(NB)
        self.Nx, self.Nu = self.N * self.NB + 2, self.N * self.NB + 1
        self.Q = np.asarray(Q, dtype=float)
        self.R = np.asarray(R, dtype=float)
        self.Qf = self.Q if Qf is None else np.asarray(Qf, dtype=float)
        self.dR = np.zeros(self.d) if dR is None else np.asarray(dR, dtype=float)
        if np.any(self.dR != 0):
            raise NotImplementedError("rate costs of robustMPC (dR != 0) are not restated; the highway scenario has dR = 0")
        self.Fx = np.asarray(Fx, dtype=float).reshape(-1, self.n)
        self.bx = np.asarray(bx, dtype=float).reshape(-1)
        self.Fu = np.asarray(Fu, dtype=float).reshape(-1, self.d)
        self.bu = np.asarray(bu, dtype=float).reshape(-1)
        self.Qslack = np.asarray(Qslack, dtype=float)
        self.xRef = np.asarray(xRef, dtype=float)
        self.uLin = self.xLin = None
        self.xPred = self.uPred = None
        self.timeStep = 0

    def obstacle_slots(self, z):
        """zPred[t]: obstacle states of time slot t, in BFS order of the scenario tree (:1336-1383)."""
        slots = [[] for _ in range(self.N * self.NB + 1)]
        slots[0].append(np.asarray(z, dtype=float))
        level = [np.asarray(z, dtype=float)]           # last obstacle state of every branch of the current depth
        for depth in range(1, self.NB + 1):
            nxt = []
            for zl in level:
                zp = self.model.zpred_eval(zl)
                for i in range(self.m):
                    traj = zp[:, i * self.n:(i + 1) * self.n]
                    for t in range(self.N):
                        slots[(depth - 1) * self.N + t + 1].append(traj[t])
                    nxt.append(traj[-1])
            level = nxt
        return slots

    def linearisation_trajectory(self, x):
        if self.uLin is None:                          # first solve: zero inputs, nonlinear rollout (get_xLin)
            self.uLin = np.zeros((self.Nu + 1, self.d))
            self.xLin = np.zeros((self.Nx, self.n))
            self.xLin[0] = x
            for i in range(self.Nx - 1):
                self.xLin[i + 1] = self.model.dyn_linearization(self.xLin[i], self.uLin[i])[3]
        # later solves: xLin/uLin were shifted from the previous QP solution at the end of solve()

    def assemble(self, x0, slots):
        n, d, Nx, Nu = self.n, self.d, self.Nx, self.Nu
        nfx = self.Fx.shape[0]
        zcount = sum(len(s) for s in slots)
        nslack = Nx * nfx + zcount
        nxv, nuv = Nx * n, Nu * d
        nvar = nxv + nuv + nslack
        H = sp.lil_matrix((nvar, nvar))
        q = np.zeros(nvar)
        for k in range(Nx):
            Qk = self.Q if k < Nx - 1 else self.Qf
            H[k * n:(k + 1) * n, k * n:(k + 1) * n] = Qk
            q[k * n:(k + 1) * n] = -2.0 * self.xRef @ Qk
        for k in range(Nu):
            H[nxv + k * d:nxv + (k + 1) * d, nxv + k * d:nxv + (k + 1) * d] = self.R
        H[nxv + nuv:, nxv + nuv:] = self.Qslack[0] * sp.eye(nslack)
        q[nxv + nuv:] = self.Qslack[1]
        P = (2.0 * H).tocsc()
        # inequalities: [Fx rows of every node; collision rows] - s <= b ;  Fu u <= bu ;  -s <= 0
        rows = sp.lil_matrix((nslack, nxv))
        b = np.zeros(nslack)
        for k in range(Nx):
            rows[k * nfx:(k + 1) * nfx, k * n:(k + 1) * n] = self.Fx
            b[k * nfx:(k + 1) * nfx] = self.bx
        c = Nx * nfx
        for t, slot in enumerate(slots):
            for zt in slot:
                h, dh = self.model.col_eval(self.xLin[t], zt)
                rows[c, t * n:(t + 1) * n] = -dh
                b[c] = h
                c += 1
        Fu_all = sp.kron(sp.eye(Nu), sp.csr_matrix(self.Fu))
        ncu = Fu_all.shape[0]
        F = sp.bmat([[rows, None, -sp.eye(nslack)],
                     [None, Fu_all, sp.csr_matrix((ncu, nslack))],
                     [sp.csr_matrix((nslack, nxv)), sp.csr_matrix((nslack, nuv)), -sp.eye(nslack)]], format="csr")
        bF = np.concatenate([b, np.tile(self.bu, Nu), np.zeros(nslack)])
        G = sp.lil_matrix((nxv, nvar))
        G[:, :nxv] = sp.eye(nxv)
        beq = np.zeros(nxv)
        beq[:n] = x0
        for i in range(Nu):
            A, B, C, _ = self.model.dyn_linearization(self.xLin[i], self.uLin[i])
            G[(i + 1) * n:(i + 2) * n, i * n:(i + 1) * n] = -A
            G[(i + 1) * n:(i + 2) * n, nxv + i * d:nxv + (i + 1) * d] = -B
            beq[(i + 1) * n:(i + 2) * n] = C
        Aall = sp.vstack([F, G.tocsr()]).tocsc()
        lo = np.concatenate([np.full(F.shape[0], -np.inf), beq])
        hi = np.concatenate([bF, beq])
        return P, q, Aall, lo, hi

    def solve(self, x, z, xRef=None):
        if xRef is not None:
            self.xRef = np.asarray(xRef, dtype=float)
        x = np.asarray(x, dtype=float)
        slots = self.obstacle_slots(z)
        self.linearisation_trajectory(x)
        self.qp = self.assemble(x, slots)
        sol, _, info = qp_exact.solve_qp(*self.qp)
        self.feasible = int(bool(info["polished"]))
        nxv, nuv = self.Nx * self.n, self.Nu * self.d
        self.Solution = sol
        self.xPred = sol[:nxv].reshape(self.Nx, self.n)
        self.uPred = sol[nxv:nxv + nuv].reshape(self.Nu, self.d)
        self.objective = qp_exact.kkt_residuals(*self.qp, sol)["objective"]
        # LTV shift (:1429-1431)
        self.xLin = np.vstack([self.xPred[1:], self.xPred[-1]])
        self.uLin = np.vstack([self.uPred[1:], self.uPred[-1], self.uPred[-1]])
        self.timeStep += 1
        return self.uPred[0]
