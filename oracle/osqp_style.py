"""TEST INFRASTRUCTURE (oracle / CPU baseline): the solver call of the reference path restated.

The reference hands its QP to OSQP with default settings and `polish=True`, a fresh `setup` per solve and no warm start
(/root/reference/MPC_branch.py:1248-1274).  OSQP (un-vendored third-party C, version unpinned) is not installable here; this
is its published algorithm (Stellato et al., "OSQP: an operator splitting solver for quadratic programs") with the default
parameters: Ruiz equilibration (10 passes), rho = 0.1 (x 1e3 on equality rows), sigma = 1e-6, alpha = 1.6, eps_abs =
eps_rel = 1e-3, max_iter = 4000, termination checked every 25 iterations, adaptive rho, polish with delta = 1e-6 and 3
refinement steps.  The quasi-definite KKT matrix is factorised with scipy's SuperLU (OSQP uses QDLDL).

Used by bench.py's `cpu_baseline` / `--impl reference` legs as "the reference path on the host cores" (kind: "port") next
to the exact-optimum oracle.  It stops where OSQP's default tolerance stops, i.e. ~1e-2 from the optimum in the first input
 - parity is NOT defined against it; it exists to time the reference's work faithfully.
"""
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla


def ruiz(P, q, A, iters=10):
    n, m = P.shape[0], A.shape[0]
    D, E = np.ones(n), np.ones(m)
    c = 1.0
    Ps, As, qs = P.copy().tocsc(), A.copy().tocsc(), q.copy()
    for _ in range(iters):
        colP = np.abs(Ps).max(axis=0).toarray().ravel()
        colA = np.abs(As).max(axis=0).toarray().ravel() if m else np.zeros(n)
        dn = np.maximum(colP, colA)
        dn = np.where(dn < 1e-4, 1.0, dn)
        d = 1.0 / np.sqrt(np.clip(dn, 1e-4, 1e4))
        rowA = np.abs(As).max(axis=1).toarray().ravel() if m else np.zeros(0)
        rowA = np.where(rowA < 1e-4, 1.0, rowA)
        e = 1.0 / np.sqrt(np.clip(rowA, 1e-4, 1e4))
        Dm, Em = sp.diags(d), sp.diags(e)
        Ps = (Dm @ Ps @ Dm).tocsc()
        As = (Em @ As @ Dm).tocsc()
        qs = d * qs
        D, E = D * d, E * e
        colmean = np.abs(Ps).max(axis=0).toarray().ravel().mean()
        g = 1.0 / max(colmean, np.abs(qs).max(), 1e-4)
        g = float(np.clip(g, 1e-4, 1e4))
        Ps, qs, c = Ps * g, qs * g, c * g
    return Ps, qs, As, D, E, c


def solve(P, q, A, l, u, eps_abs=1e-3, eps_rel=1e-3, max_iter=4000, rho0=0.1, sigma=1e-6, alpha=1.6, check=25,
          polish=True, densify=False):
    """Returns (x, info).  densify=True first walks the dense->CSC conversions the reference does (MPC_branch.py:1191-1195)."""
    P = sp.csc_matrix(P)
    A = sp.csc_matrix(A)
    if densify:
        P = sp.csc_matrix(np.asarray(P.todense()))
        A = sp.csc_matrix(np.asarray(A.todense()))
    P = sp.triu(P).tocsc()
    P = (P + sp.triu(P, 1).T).tocsc()          # the Python wrapper keeps triu(P)
    q, l, u = np.asarray(q, float), np.asarray(l, float), np.asarray(u, float)
    n, m = P.shape[0], A.shape[0]
    Ps, qs, As, D, E, c = ruiz(P, q, A)
    ls, us = E * l, E * u
    eq = np.abs(us - ls) < 1e-4
    rho = rho0
    x, z, y = np.zeros(n), np.zeros(m), np.zeros(m)

    def factor(rho):
        rv = np.where(eq, 1e3 * rho, rho)
        K = sp.bmat([[Ps + sigma * sp.eye(n), As.T], [As, -sp.diags(1.0 / rv)]], format="csc")
        return spla.splu(K), rv

    lu, rv = factor(rho)
    it, status = 0, "max_iter"
    nfact = 1
    for it in range(1, max_iter + 1):
        sol = lu.solve(np.concatenate([sigma * x - qs, z - y / rv]))
        xt, nu = sol[:n], sol[n:]
        zt = z + (nu - y) / rv
        xn = alpha * xt + (1 - alpha) * x
        zr = alpha * zt + (1 - alpha) * z
        zn = np.clip(zr + y / rv, ls, us)
        y = y + rv * (zr - zn)
        x, z = xn, zn
        if it % check == 0:
            Ax, Px, Aty = As @ x, Ps @ x, As.T @ y
            # residuals in the unscaled space
            rp = np.abs((Ax - z) / E).max()
            rd = np.abs((Px + qs + Aty) / D).max() / c
            ep = eps_abs + eps_rel * max(np.abs(Ax / E).max(), np.abs(z / E).max())
            ed = eps_abs + eps_rel * max(np.abs(Px / D).max(), np.abs(Aty / D).max(), np.abs(qs / D).max()) / c
            if rp <= ep and rd <= ed:
                status = "solved"
                break
            nrm_p = rp / max(np.abs(Ax / E).max(), np.abs(z / E).max(), 1e-10)
            nrm_d = rd / max(np.abs(Px / D).max() / c, np.abs(Aty / D).max() / c, np.abs(qs / D).max() / c, 1e-10)
            new = float(np.clip(rho * np.sqrt(nrm_p / max(nrm_d, 1e-10)), 1e-6, 1e6))
            if new > 5 * rho or new < 0.2 * rho:
                rho = new
                lu, rv = factor(rho)
                nfact += 1
    xo, yo = D * x, E * y / c
    polished = False
    if polish and status == "solved":
        # active constraints from the duals, equality-constrained solve with delta-regularisation and 3 refinements
        low = yo < -1e-12
        upp = yo > 1e-12
        act = low | upp | eq
        if act.any():
            Aa = A[np.flatnonzero(act)]
            ba = np.where(low, l, u)[act]
            delta = 1e-6
            K = sp.bmat([[P + delta * sp.eye(n), Aa.T], [Aa, -delta * sp.eye(Aa.shape[0])]], format="csc")
            Kx = sp.bmat([[P, Aa.T], [Aa, None]], format="csc")
            lup = spla.splu(K)
            rhs = np.concatenate([-q, ba])
            t = lup.solve(rhs)
            for _ in range(3):
                t = t + lup.solve(rhs - Kx @ t)
            xp = t[:n]
            Axp = A @ xp
            if np.all(Axp >= l - 1e-6) and np.all(Axp <= u + 1e-6):
                xo = xp
                polished = True
            nfact += 1
    return xo, {"status": status, "iters": it, "factorizations": nfact, "polished": polished, "rho": rho}


last_info = {}


def qp_solver(P, q, A, l, u):
    """Adapter for BranchMPCOracle.solve(qp_solver=...): (solution, feasible) as osqp_solve_qp returns (:1269-1274)."""
    x, info = solve(P, q, A, l, u, densify=True)
    last_info.clear()
    last_info.update(info)
    return x, info["status"] == "solved"
