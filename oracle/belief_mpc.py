"""TEST INFRASTRUCTURE (oracle): float64 numpy restatement of the reference's belief-state MPC.

  model      /root/reference/HMM_backup_dyn.py:177-276  PredictiveModel (calc_xp_expr :238-267, regressionAndLinearization :216-229),
             with the CasADi graphs written out in closed form (dubin :30-41, veh_col symbolic branch :136-143, lane_bdry_h :134,
             softmin :111, softsat :94, backup_trans :96-101)
  controller /root/reference/PredictiveControllers.py:56-340  MPC (get_xLin :115-127, solve :130-160, computeLTVdynamics :162-166,
             buildIneqConstr :195-242, buildEqConstr :244-268, buildCost :270-296)
  parameters /root/reference/Init_MPC.py:7-34  initMPCParams

Pinned by tests/golden/belief_mpc_*.npz: the UNMODIFIED reference classes in closed loop (tests/golden/make_golden.py `belief`).

Layout quirks that define parity (all reproduced):
  * the augmented state is xb = [x (4); b flattened]; the model reads the flat belief column-major (casadi.reshape, :244: entry
    k M + i is agent i, policy k) while the controller gates its rows with a row-major view (np.reshape, PredictiveControllers.py:212);
  * stage i of the horizon is linearised about xLin[i+1] with the backup states of step i (:164);
  * the row of (agent j, policy k) on node i+1 uses h0/Jh of linearisation i+1 and is present only if the row-major belief entry
    [j, k] of xLin[i+1] exceeds 0.1 (:211-216); the last node carries no rows; every state row is soft (weight Qslack[1] = 1000).
"""
import numpy as np
import scipy.sparse as sp

from . import qp_exact


def softmin2(x, y, g):
    mn = min(x, y)
    ex, ey = np.exp(-g * (x - mn)), np.exp(-g * (y - mn))
    return (ex * x + ey * y) / (ex + ey), ex / (ex + ey)


class BeliefModelOracle:
    """HMM_backup_dyn.PredictiveModel(n=4, d=2, M, backupcons, dt, cons): value and Jacobians of [x; b] -> [x+; b H(x)]."""

    def __init__(self, M, m, dt, L=4.0, W=2.5, ylb=0.0, yub=7.2, col_alpha=5.0, s1=2.0, tran_diag=0.3, alpha=1.0):
        self.M, self.m, self.dt = M, m, dt
        self.size = (L + 1.0, W + 0.2)
        self.ylb, self.yub, self.col_alpha, self.s1, self.tau, self.alpha = ylb, yub, col_alpha, s1, tran_diag, alpha

    def safety(self, x, xb):
        """h = softmin(veh_col(x, xb), lane_bdry_h(xb), col_alpha) (:255) and its gradient in (x, y)."""
        s0, s1 = self.size
        ex, ey = x[0] - xb[0], x[1] - xb[1]
        dx, dy = (abs(ex) - s0) / s0, (abs(ey) - s1) / s1
        mx = max(dx, dy)
        wx = np.exp(dx - mx) / (np.exp(dx - mx) + np.exp(dy - mx))
        v = wx * dx + (1 - wx) * dy
        gvx = wx * (1 + dx - v) * np.sign(ex) / s0
        gvy = (1 - wx) * (1 + dy - v) * np.sign(ey) / s1
        lane, _ = softmin2(xb[1] - self.ylb, self.yub - xb[1], 5.0)
        h, wv = softmin2(v, lane, self.col_alpha)
        dh_dv = wv * (1.0 - self.col_alpha * (v - h))
        return h, np.array([dh_dv * gvx, dh_dv * gvy])

    def linearize(self, xb, xbackup, u):
        """regressionAndLinearization (:216-229): A, B, C, h0 [M][m], Jh [M][m][n]."""
        M, m, dt = self.M, self.m, self.dt
        n = 4 + M * m
        x, b = xb[:4], xb[4:]
        c, s = np.cos(x[3]), np.sin(x[3])
        xp = x + dt * np.array([x[2] * c, x[2] * s, u[0], u[1]])
        A = np.zeros((n, n))
        A[:4, :4] = np.eye(4)
        A[0, 2], A[0, 3], A[1, 2], A[1, 3] = dt * c, -dt * x[2] * s, dt * s, dt * x[2] * c
        B = np.zeros((n, 2))
        B[2, 0] = B[3, 1] = dt
        bp = np.zeros(M * m)
        h0 = np.zeros((M, m))
        Jh = np.zeros((M, m, n))
        for i in range(M):
            hv = np.zeros(m)
            gh = np.zeros((m, 2))
            for j in range(m):
                hv[j], gh[j] = self.safety(x, xbackup[m * i + j])
            mh = 1.0 / (1.0 + np.exp(-self.s1 * hv))            # softsat(h, s1) = sigmoid(s1 h)
            pi = mh / mh.sum()
            idx = [k * M + i for k in range(m)]                 # column-major positions of b[i, :]
            bi = b[idx]
            bsum = bi.sum()
            bp[idx] = (1 - self.tau) * bsum * pi + self.tau * bi
            H = (1 - self.tau) * np.outer(np.ones(m), pi) + self.tau * np.eye(m)
            for k in range(m):
                for r in range(m):
                    A[4 + idx[k], 4 + idx[r]] = H[r, k]
            dm = self.s1 * mh * (1 - mh)                         # d softsat / dh
            dpi = (np.eye(m) - pi[:, None]) / mh.sum()           # d pi_k / d mh_l
            dbp_dh = (1 - self.tau) * bsum * dpi * dm[None, :]   # [k, l]
            for k in range(m):
                A[4 + idx[k], 0:2] = dbp_dh[k] @ gh
            for j in range(m):
                Jh[i, j, 0:2] = gh[j]
                h0[i, j] = hv[j] - gh[j] @ x[:2]
        xbp = np.concatenate([xp, bp])
        C = xbp - A @ xb - B @ u
        return A, B, C, h0, Jh, xbp


class BeliefMPCOracle:
    """PredictiveControllers.MPC with Init_MPC.initMPCParams constants."""

    def __init__(self, model, N, ydes, vdes, am=6.0, rm=0.3, N_lane=2, W=2.5, thres=0.1):
        self.model, self.N = model, N
        M, m = model.M, model.m
        self.n, self.d, self.nx = 4 + M * m, 2, 4
        Fx = np.array([[0., 1., 0., 0.], [0., -1., 0., 0.], [0., 0., 0., 1.], [0., 0., 0., -1.]])
        self.Fx = np.hstack([Fx, np.zeros((4, M * m))])
        self.bx = np.array([N_lane * 3.6 - W / 2, -W / 2, 0.25, 0.25])
        self.Fu = np.kron(np.eye(2), np.array([1., -1.])).T
        self.bu = np.array([am, 0.5 * am, rm, rm])
        self.Q = np.zeros((self.n, self.n))
        self.Q[:4, :4] = np.diag([0., 0.5, 0.2, 5.])
        self.R = np.diag([30., 100.])
        self.Qf = np.zeros((self.n, self.n))
        self.dR = np.zeros(2)
        self.Qslack = np.array([0., 1000.])
        self.xRef = np.concatenate([[0, ydes, vdes, 0.], np.zeros(M * m)])
        self.thres = thres
        self.uLin = None
        self.OldInput = np.zeros(2)
        self.xPred = self.uPred = None
        self.feasible = 0

    def solve(self, x0, b0, xbackup, xRef=None):
        N, n, d, nx = self.N, self.n, self.d, self.nx
        M, m = self.model.M, self.model.m
        if xRef is not None:
            self.xRef = np.concatenate([np.asarray(xRef, float), np.zeros(M * m)])
        b0 = np.asarray(b0, dtype=float)
        # get_xLin (:115-127): nonlinear rollout under the shifted inputs; the belief enters column-major (np.reshape(b0, -1, 1):
        # the legacy integer `order` 1 = Fortran)
        if self.uLin is None:
            self.uLin = np.zeros((N, d))
        self.uLin = np.vstack([self.uLin, self.uLin[-1]])
        xLin = np.zeros((N + 1, n))
        xb = np.concatenate([np.asarray(x0, float), b0.reshape(-1, order="F")])
        xLin[0] = xb
        for i in range(N):
            xb = self.model.linearize(xb, xbackup[:, i * nx:(i + 1) * nx], self.uLin[i])[5]
            xLin[i + 1] = xb
        self.xLin = xLin
        # computeLTVdynamics (:162-166)
        lin = [self.model.linearize(xLin[i + 1], xbackup[:, i * nx:(i + 1) * nx], self.uLin[i + 1]) for i in range(N)]
        self.A, self.B, self.C = [l[0] for l in lin], [l[1] for l in lin], [l[2] for l in lin]
        self.h0, self.Jh = [l[3] for l in lin], [l[4] for l in lin]
        # buildIneqConstr (:195-242)
        rows, rhs = [], []
        for i in range(N):
            for r in range(4):
                f = np.zeros(n * (N + 1))
                f[i * n:(i + 1) * n] = self.Fx[r]
                rows.append(f)
                rhs.append(self.bx[r])
        for i in range(N - 1):
            bq = xLin[i + 1][nx:].reshape(M, m)                   # row-major view (:212)
            for j in range(M):
                for k in range(m):
                    if bq[j, k] > self.thres:
                        f = np.zeros(n * (N + 1))
                        f[(i + 1) * n:(i + 2) * n] = -self.Jh[i + 1][j][k]
                        rows.append(f)
                        rhs.append(self.h0[i + 1][j][k])
        Fx_tot, bx_tot = np.array(rows), np.array(rhs)
        nc = Fx_tot.shape[0]
        Fu_tot = np.kron(np.eye(N), self.Fu)
        bu_tot = np.tile(self.bu, N)
        nxu = n * (N + 1) + d * N
        F = np.zeros((nc + 4 * N + nc, nxu + nc))
        F[:nc, :n * (N + 1)] = Fx_tot
        F[:nc, nxu:] = -np.eye(nc)
        F[nc:nc + 4 * N, n * (N + 1):nxu] = Fu_tot
        F[nc + 4 * N:, nxu:] = -np.eye(nc)
        bF = np.concatenate([bx_tot, bu_tot, np.zeros(nc)])
        # buildCost (:270-296), dR = 0
        H = np.zeros((nxu + nc, nxu + nc))
        for i in range(N):
            H[i * n:(i + 1) * n, i * n:(i + 1) * n] = self.Q
        H[N * n:(N + 1) * n, N * n:(N + 1) * n] = self.Qf
        for i in range(N):
            o = n * (N + 1) + i * d
            H[o:o + d, o:o + d] = self.R
        q = np.zeros(nxu + nc)
        q[:nxu] = -2 * (np.concatenate([np.tile(self.xRef, N + 1), np.zeros(d * N)]) @ H[:nxu, :nxu])
        q[n * (N + 1):n * (N + 1) + d] = -2 * self.OldInput @ np.diag(self.dR)
        q[nxu:] = self.Qslack[1]
        H = 2 * H
        # buildEqConstr (:244-268)
        G = np.zeros((n * (N + 1), nxu + nc))
        G[:, :n * (N + 1)] = np.eye(n * (N + 1))
        Lv = np.zeros(n * (N + 1))
        for i in range(N):
            G[n + i * n:2 * n + i * n, i * n:(i + 1) * n] = -self.A[i]
            G[n + i * n:2 * n + i * n, n * (N + 1) + i * d:n * (N + 1) + (i + 1) * d] = -self.B[i]
            Lv[n + i * n:2 * n + i * n] = self.C[i]
        xb0 = np.concatenate([np.asarray(x0, float), b0.reshape(-1)])           # row-major here (:147)
        beq = Lv.copy()
        beq[:n] += xb0
        Aall = sp.vstack([sp.csc_matrix(F), sp.csc_matrix(G)]).tocsc()
        lo = np.concatenate([np.full(F.shape[0], -np.inf), beq])
        hi = np.concatenate([bF, beq])
        self.qp = (sp.csc_matrix(H), q, Aall, lo, hi)
        sol, ydual, info = qp_exact.solve_qp(*self.qp)
        self.feasible = int(bool(info["polished"]))
        self.Solution = sol
        self.xPred = sol[:n * (N + 1)].reshape(N + 1, n)
        self.uPred = sol[n * (N + 1):nxu].reshape(N, d)
        self.objective = qp_exact.kkt_residuals(*self.qp, sol)["objective"]
        # timeVarying (:156-158)
        self.uLin = np.vstack([self.uPred[1:], self.uPred[-1]])
        self.OldInput = self.uPred[0].copy()
        return self.uPred[0]
