"""CPU model (float64 numpy) of the DEVICE algorithm: ADMM whose KKT step is a Riccati sweep over the branch tree.

TEST INFRASTRUCTURE.  This is not the reference's algorithm (the reference calls OSQP on a dense
QP); it is a readable model of what the CUDA kernel in belief-planning_b200/csrc does, used to
develop the method and to debug the kernel.  It consumes the stage data of a BranchMPCOracle
(effective BranchMPC variant) and solves the slack-free exact-penalty form of the same QP
(SURVEY Appendix A):

  min  sum_k [ 1/2 x'Qk x + qk'x + 1/2 u'Rk u + rk'u + sum_j lam_kj dist(f_kj'x, [lo_kj, hi_kj]) ] + terminal
  s.t. tree dynamics,  ulo <= u <= uhi
"""
import numpy as np


class StageProblem:
    """Flat per-node arrays extracted from an expanded BranchMPCOracle ('branch' variant)."""

    def __init__(self, mpc, x0):
        T = mpc.topo
        n, d = mpc.n, mpc.d
        self.n, self.d = n, d
        self.nu = T.totalu
        self.nx = T.totalx
        nu = self.nu
        self.x0 = np.asarray(x0, dtype=float)
        self.A = np.zeros((nu, n, n))
        self.B = np.zeros((nu, n, d))
        self.C = np.zeros((nu, n))
        self.Q = np.zeros((nu, n, n))
        self.q = np.zeros((nu, n))
        self.R = np.zeros((nu, d, d))
        self.r = np.zeros((nu, d))
        self.w = np.zeros(nu)
        self.xnode = np.zeros(nu, dtype=int)        # x-node index of each u-node
        self.succ_x = [None] * nu                   # x-node indices the dynamics row writes (children share a value)
        self.succ_u = [None] * nu                   # successor u-nodes ([] -> terminal)
        self.term_Q = {}                            # u-node -> terminal (Q, q) of a leaf's last node
        self.order = []                             # u-nodes in a topological (root first) order
        rows = merge_range_rows(mpc.Fx, mpc.bx)
        self.c = 1 + len(rows)
        c = self.c
        self.f = np.zeros((nu, c, n))
        self.lo = np.full((nu, c), -np.inf)
        self.hi = np.zeros((nu, c))
        self.lam = np.zeros((nu, c))
        self.ulo, self.uhi = box_from_Fu(mpc.Fu, mpc.bu, d)
        dQ = mpc.Q * mpc.dQ_scale
        for b in range(T.nbranch):
            l = T.length[b]
            for i in range(l):
                k = T.ndu[b] + i
                self.order.append(k)
                self.xnode[k] = T.ndx[b] + i
                A, B, C = mpc.lin[b][i]
                self.A[k], self.B[k], self.C[k] = A, B, C
                wb = mpc.w[b]
                self.w[k] = wb
                self.Q[k] = 2.0 * wb * (mpc.Q + dQ)
                last = i == l - 1
                Qlin = mpc.Qf if (last and T.is_leaf[b]) else mpc.Q
                self.q[k] = -2.0 * wb * (mpc.xRef @ Qlin + mpc.xbar[b][i] @ dQ)
                self.R[k] = 2.0 * wb * mpc.R
                hlin, dh = mpc.model.col_eval(mpc.xbar[b][i], mpc.zbar[b][i])
                self.f[k, 0] = -dh
                self.hi[k, 0] = hlin
                for j, (fv, lo, hi) in enumerate(rows):
                    self.f[k, 1 + j] = fv
                    self.lo[k, 1 + j] = lo
                    self.hi[k, 1 + j] = hi
                self.lam[k, :] = mpc.Qslack[1] * wb
                if not last:
                    self.succ_u[k] = [k + 1]
                    self.succ_x[k] = [self.xnode[k] + 1]
                elif not T.is_leaf[b]:
                    self.succ_u[k] = [T.ndu[cb] for cb in T.children(b)]
                    self.succ_x[k] = [T.ndx[cb] for cb in T.children(b)]
                else:
                    self.succ_u[k] = []
                    self.succ_x[k] = [self.xnode[k] + 1]
                    self.term_Q[k] = (2.0 * wb * mpc.Qf, np.zeros(n))
        self.r[0] = -2.0 * (mpc.OldInput @ mpc.dR)

    def rollout(self, u):
        x = np.zeros((self.nx, self.n))
        x[0] = self.x0
        for k in self.order:
            xn = self.A[k] @ x[self.xnode[k]] + self.B[k] @ u[k] + self.C[k]
            for j in self.succ_x[k]:
                x[j] = xn
        return x

    def objective(self, x, u):
        J = 0.0
        for k in self.order:
            xk = x[self.xnode[k]]
            J += 0.5 * xk @ self.Q[k] @ xk + self.q[k] @ xk + 0.5 * u[k] @ self.R[k] @ u[k] + self.r[k] @ u[k]
            v = self.f[k] @ xk
            J += (self.lam[k] * (np.maximum(v - self.hi[k], 0.0) + np.maximum(self.lo[k] - v, 0.0))).sum()
            if k in self.term_Q:
                xt = x[self.succ_x[k][0]]
                Qt, qt = self.term_Q[k]
                J += 0.5 * xt @ Qt @ xt + qt @ xt
        return J


def merge_range_rows(Fx, bx, tol=1e-12):
    """Pair opposite rows of Fx x <= bx into two-sided rows lo <= f'x <= hi."""
    Fx = np.asarray(Fx, dtype=float)
    bx = np.asarray(bx, dtype=float).reshape(-1)
    used = [False] * len(bx)
    rows = []
    for i in range(len(bx)):
        if used[i]:
            continue
        used[i] = True
        lo = -np.inf
        for j in range(i + 1, len(bx)):
            if not used[j] and np.abs(Fx[i] + Fx[j]).max() < tol:
                used[j] = True
                lo = -bx[j]
                break
        rows.append((Fx[i].copy(), lo, bx[i]))
    return rows


def box_from_Fu(Fu, bu, d):
    Fu = np.asarray(Fu, dtype=float)
    bu = np.asarray(bu, dtype=float).reshape(-1)
    lo = np.full(d, -np.inf)
    hi = np.full(d, np.inf)
    for row, b in zip(Fu, bu):
        k = int(np.argmax(np.abs(row)))
        assert np.count_nonzero(row) == 1, "input constraints must be a box"
        if row[k] > 0:
            hi[k] = min(hi[k], b / row[k])
        else:
            lo[k] = max(lo[k], b / row[k])
    return lo, hi


class TreeRiccati:
    """Factorisation of the tree LQ problem with stage Hessians Q + F'diag(rho)F and R + diag(rho_u)."""

    def __init__(self, sp, rho, rho_u):
        self.sp = sp
        n, d, nu = sp.n, sp.d, sp.nu
        self.K = np.zeros((nu, d, n))
        self.Sinv = np.zeros((nu, d, d))
        self.P = np.zeros((nu, n, n))
        self.Pn = np.zeros((nu, n, n))             # summed successor value Hessian
        for k in reversed(sp.order):
            if sp.succ_u[k]:
                Pn = sum(self.P[j] for j in sp.succ_u[k])
            else:
                Pn = sp.term_Q[k][0]
            self.Pn[k] = Pn
            A, B = sp.A[k], sp.B[k]
            Qt = sp.Q[k] + sp.f[k].T @ (rho[k][:, None] * sp.f[k])
            Rt = sp.R[k] + np.diag(rho_u[k])
            S = Rt + B.T @ Pn @ B
            self.Sinv[k] = np.linalg.inv(S)
            self.K[k] = self.Sinv[k] @ (B.T @ Pn @ A)
            Pk = Qt + A.T @ Pn @ (A - B @ self.K[k])
            self.P[k] = 0.5 * (Pk + Pk.T)

    def solve(self, qx, qu):
        """min sum 1/2 x'Q~x + qx'x + 1/2 u'R~u + qu'u (+ terminal) s.t. dynamics -> (x, u)."""
        sp = self.sp
        n, d, nu = sp.n, sp.d, sp.nu
        p = np.zeros((nu, n))
        kff = np.zeros((nu, d))
        for k in reversed(sp.order):
            if sp.succ_u[k]:
                pn = sum(p[j] for j in sp.succ_u[k])
            else:
                pn = sp.term_Q[k][1]
            g = pn + self.Pn[k] @ sp.C[k]
            rr = qu[k] + sp.B[k].T @ g
            kff[k] = -self.Sinv[k] @ rr
            p[k] = qx[k] + sp.A[k].T @ g - self.K[k].T @ rr
        x = np.zeros((sp.nx, n))
        u = np.zeros((nu, d))
        x[0] = sp.x0
        for k in sp.order:
            xk = x[sp.xnode[k]]
            u[k] = -self.K[k] @ xk + kff[k]
            xn = sp.A[k] @ xk + sp.B[k] @ u[k] + sp.C[k]
            for j in sp.succ_x[k]:
                x[j] = xn
        return x, u


def prox_range(s, lo, hi, t):
    """prox of lam*dist(., [lo,hi]) with threshold t = lam/rho."""
    return np.where(s > hi + t, s - t, np.where(s > hi, hi, np.where(s >= lo, s, np.where(s >= lo - t, lo, s + t))))


def admm_tree(sp, rho0=1.0, rho_u0=1.0, alpha=1.6, max_iter=2000, check_every=25, polish_every=25,
              eps_abs=1e-4, verbose=False, adapt=True, warm=None, polish_tol=1e-7):
    """ADMM with per-row rho = rho0*w (soft rows) / rho_u0*w (input box), periodic polish attempts.

    Returns dict(x, u, iters, polished, factorizations).
    """
    nu, c, d = sp.nu, sp.c, sp.d
    rho = rho0 * sp.w[:, None] * np.ones((nu, c))
    rho_u = rho_u0 * sp.w[:, None] * np.ones((nu, d))
    fac = TreeRiccati(sp, rho, rho_u)
    nfac = 1
    if warm is None:
        s = np.zeros((nu, c))          # s = v^ + y/rho  (Moreau variable) for soft rows
        su = np.zeros((nu, d))
        # start from "no constraint active": v = clip(0) ; y = 0
        v = np.clip(s, np.where(np.isfinite(sp.lo), sp.lo, -1e30), sp.hi)
        s = v.copy()
    else:
        s, su = warm
    info = {"polished": False}
    for it in range(1, max_iter + 1):
        v = prox_range(s, sp.lo, sp.hi, sp.lam / rho)
        y = rho * (s - v)
        tu = np.clip(su, sp.ulo, sp.uhi)
        yu = rho_u * (su - tu)
        qx = sp.q - np.einsum("kcn,kc->kn", sp.f, rho * v - y)
        qu = sp.r - (rho_u * tu - yu)
        x, u = fac.solve(qx, qu)
        Fx = np.einsum("kcn,kn->kc", sp.f, x[sp.xnode])
        s_new = alpha * Fx + (1 - alpha) * v + y / rho
        su_new = alpha * u + (1 - alpha) * tu + yu / rho_u
        v_new = prox_range(s_new, sp.lo, sp.hi, sp.lam / rho)
        tu_new = np.clip(su_new, sp.ulo, sp.uhi)
        r_prim = max(np.abs(Fx - v_new).max(), np.abs(u - tu_new).max())
        r_dual = max(np.abs(rho * (v_new - v)).max(), np.abs(rho_u * (tu_new - tu)).max())
        s, su = s_new, su_new
        if verbose and it % check_every == 0:
            print("it %4d  r_prim %.2e  r_dual %.2e  J %.6f" % (it, r_prim, r_dual, sp.objective(sp.rollout(np.clip(u, sp.ulo, sp.uhi)), np.clip(u, sp.ulo, sp.uhi))))
        if it % polish_every == 0:
            res = polish_tree(sp, s, su, rho, rho_u, tol=polish_tol)
            nfac += 1
            if res is not None:
                info.update(x=res[0], u=res[1], iters=it, polished=True, factorizations=nfac, pol_iters=res[2])
                return info
    tu = np.clip(u, sp.ulo, sp.uhi)
    info.update(x=sp.rollout(tu), u=tu, iters=max_iter, factorizations=nfac)
    return info


def polish_tree(sp, s, su, rho, rho_u, big=1e4, iters=6, tol=1e-7):
    """Guess the active set from the ADMM state, solve the equality-constrained problem by a few
    augmented-Lagrangian steps with a stiff penalty on the guessed-active rows, and verify KKT."""
    nu, c, d = sp.nu, sp.c, sp.d
    t = sp.lam / rho
    up_lin = s > sp.hi + t
    up_kink = (s > sp.hi) & ~up_lin
    lo_lin = s < sp.lo - t
    lo_kink = (s < sp.lo) & ~lo_lin
    kink = up_kink | lo_kink
    target = np.where(up_kink, sp.hi, np.where(lo_kink, sp.lo, 0.0))
    lin_grad = np.where(up_lin, sp.lam, np.where(lo_lin, -sp.lam, 0.0))
    at_hi = su > sp.uhi
    at_lo = su < sp.ulo
    fixed = at_hi | at_lo
    utarget = np.where(at_hi, sp.uhi, np.where(at_lo, sp.ulo, 0.0))
    scale = sp.w[:, None]
    rk = np.where(kink, big * scale, 0.0)
    ru = np.where(fixed, big * scale, 0.0)
    fac = TreeRiccati(sp, rk, ru)
    y = np.where(kink, np.clip(rho * (s - np.where(up_kink, sp.hi, sp.lo)), -sp.lam, sp.lam), 0.0)
    yu = np.where(fixed, rho_u * (su - utarget), 0.0)
    for it in range(iters):
        qx = sp.q + np.einsum("kcn,kc->kn", sp.f, lin_grad + y - rk * target)
        qu = sp.r + yu - ru * utarget
        x, u = fac.solve(qx, qu)
        Fx = np.einsum("kcn,kn->kc", sp.f, x[sp.xnode])
        res_k = np.where(kink, Fx - target, 0.0)
        res_u = np.where(fixed, u - utarget, 0.0)
        y = y + rk * res_k
        yu = yu + ru * res_u
        if max(np.abs(res_k).max(), np.abs(res_u).max()) < tol * 0.1:
            break
    # verification of the guessed sets
    ok = True
    ok &= np.all(np.where(kink, (np.abs(y) <= sp.lam * (1 + 1e-9) + 1e-12) & (np.where(up_kink, y, -y) >= -1e-9 * sp.lam), True))
    inactive = ~(kink | up_lin | lo_lin)
    ok &= np.all(np.where(inactive, (Fx <= sp.hi + tol) & (Fx >= sp.lo - tol), True))
    ok &= np.all(np.where(up_lin, Fx >= sp.hi - tol, True)) and np.all(np.where(lo_lin, Fx <= sp.lo + tol, True))
    ok &= np.all(np.where(~fixed, (u <= sp.uhi + tol) & (u >= sp.ulo - tol), True))
    ok &= np.all(np.where(at_hi, yu >= -1e-9, True)) and np.all(np.where(at_lo, yu <= 1e-9, True))
    ok &= max(np.abs(res_k).max(), np.abs(res_u).max()) < tol
    if not ok:
        return None
    u = np.where(fixed, utarget, u)
    return sp.rollout(u), u, it + 1
