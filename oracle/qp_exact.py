"""Exact QP solves for the oracle.  TEST INFRASTRUCTURE — never imported by the product path.

The reference hands its assembled QP to OSQP (`/root/reference/MPC_branch.py:1248-1274`).
OSQP is an un-vendored third-party dependency (absent here, version unpinned), so the
oracle defines "the reference's answer" as the exact optimum of the QP the reference
assembles (unique in (x,u): R>0 and the dynamics are equalities).  Two independent
solvers give that optimum:

* `solve_qp_highs`  - HiGHS' active-set QP solver through scipy's vendored bindings;
* `solve_qp_ipm`    - a plain dense/sparse primal-dual interior-point method written here.

`kkt_residuals` certifies either answer independently of how it was obtained.
"""
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla


def solve_qp_highs(P, q, A, l, u, time_limit=60.0):
    """min 1/2 z'Pz + q'z  s.t.  l <= A z <= u   (OSQP's problem form).

    Returns (z, row_dual, status_string).  P is symmetrised from its upper triangle,
    which is what the osqp Python wrapper does before calling the C library.
    """
    from scipy.optimize._highspy import _core as hc

    P = sp.csc_matrix(P)
    Pu = sp.triu(P, format="csc")
    Psym = (Pu + sp.triu(Pu, 1).T).tocsc()
    Plow = sp.tril(Psym, format="csc")
    A = sp.csc_matrix(A)
    nvar = P.shape[0]
    nrow = A.shape[0]
    inf = hc.kHighsInf

    lp = hc.HighsLp()
    lp.num_col_ = nvar
    lp.num_row_ = nrow
    lp.col_cost_ = np.asarray(q, dtype=float)
    lp.col_lower_ = np.full(nvar, -inf)
    lp.col_upper_ = np.full(nvar, inf)
    lp.row_lower_ = np.where(np.isfinite(l), l, -inf).astype(float)
    lp.row_upper_ = np.where(np.isfinite(u), u, inf).astype(float)
    lp.a_matrix_.format_ = hc.MatrixFormat.kColwise
    lp.a_matrix_.num_col_ = nvar
    lp.a_matrix_.num_row_ = nrow
    lp.a_matrix_.start_ = A.indptr.astype(np.int32)
    lp.a_matrix_.index_ = A.indices.astype(np.int32)
    lp.a_matrix_.value_ = A.data.astype(float)

    hess = hc.HighsHessian()
    hess.dim_ = nvar
    hess.format_ = hc.HessianFormat.kTriangular
    hess.start_ = Plow.indptr.astype(np.int32)
    hess.index_ = Plow.indices.astype(np.int32)
    hess.value_ = Plow.data.astype(float)

    h = hc._Highs()
    h.setOptionValue("output_flag", False)
    h.setOptionValue("time_limit", float(time_limit))
    h.setOptionValue("primal_feasibility_tolerance", 1e-9)
    h.setOptionValue("dual_feasibility_tolerance", 1e-9)
    h.passModel(lp)
    h.passHessian(hess)
    h.run()
    status = h.modelStatusToString(h.getModelStatus())
    sol = h.getSolution()
    return np.array(sol.col_value), np.array(sol.row_dual), status


def solve_qp_ipm(P, q, A, l, u, tol=1e-10, max_iter=200, prox=1e-7):
    """Mehrotra predictor-corrector IPM for  min 1/2 z'Pz + q'z, l <= Az <= u.

    Rows with l == u are equalities; rows with l = -inf are one-sided.  Written
    independently of HiGHS so the two can cross-check each other.
    """
    P = sp.csc_matrix(P)
    Pu = sp.triu(P, format="csc")
    # `prox`: a tiny Tikhonov term so zero-cost, one-side-bounded variables (the slacks of the
    # reference's all-zero terminal rows, MPC_branch.py:1125) cannot drift; `polish` removes its bias.
    P = (Pu + sp.triu(Pu, 1).T).tocsc() + prox * sp.eye(P.shape[0], format="csc")
    A = sp.csr_matrix(A)
    l = np.asarray(l, dtype=float)
    u = np.asarray(u, dtype=float)
    eq = np.isfinite(l) & np.isfinite(u) & (l == u)
    up = np.isfinite(u) & ~eq
    lo = np.isfinite(l) & ~eq
    Ae = A[eq]
    be = u[eq]
    G = sp.vstack([A[up], -A[lo]]).tocsr()
    h = np.concatenate([u[up], -l[lo]])
    n = P.shape[0]
    me = Ae.shape[0]
    mi = G.shape[0]

    z = np.zeros(n)
    y = np.zeros(me)
    s = np.ones(mi)
    lam = np.ones(mi)
    # start: make slacks consistent and positive
    r = h - G @ z
    s = np.maximum(r, 1.0)

    def solve_newton(rd, rp_e, rp_i, rc):
        # [P + G' W G   Ae'] [dz]   [ -rd - G' (lam/s * rp_i ... ) ]
        w = lam / s
        K = sp.bmat([[P + G.T @ sp.diags(w) @ G + 1e-12 * sp.eye(n), Ae.T],
                     [Ae, -1e-12 * sp.eye(me)]], format="csc")
        rhs1 = -rd - G.T @ (w * rp_i - rc / s)
        rhs = np.concatenate([rhs1, -rp_e])
        sol = spla.splu(K).solve(rhs)
        dz = sol[:n]
        dy = sol[n:]
        ds = -rp_i - G @ dz
        dlam = -(rc + lam * ds) / s
        return dz, dy, ds, dlam

    for it in range(max_iter):
        rd = P @ z + q + Ae.T @ y + G.T @ lam
        rp_e = Ae @ z - be
        rp_i = G @ z + s - h
        mu = (s @ lam) / max(mi, 1)
        scale = 1.0 + max(np.abs(q).max(), 1.0)
        if max(np.abs(rd).max() / scale, np.abs(rp_e).max() if me else 0.0,
               np.abs(rp_i).max() if mi else 0.0) < tol and mu < tol:
            break
        # Near the optimum the barrier system becomes numerically singular; the last good iterate is then handed to
        # `polish`, which computes the exact optimum on the identified active set.
        try:
            dz, dy, ds, dl = solve_newton(rd, rp_e, rp_i, s * lam)
        except RuntimeError:
            break
        if not (np.isfinite(dz).all() and np.isfinite(dl).all()):
            break

        def step(v, dv):
            neg = dv < 0
            return min(1.0, (0.995 * (-v[neg] / dv[neg])).min()) if neg.any() else 1.0
        a_aff = min(step(s, ds), step(lam, dl))
        mu_aff = ((s + a_aff * ds) @ (lam + a_aff * dl)) / max(mi, 1)
        sigma = (mu_aff / mu) ** 3 if mu > 0 else 0.0
        rc = s * lam + ds * dl - sigma * mu
        try:
            dz, dy, ds, dl = solve_newton(rd, rp_e, rp_i, rc)
        except RuntimeError:
            break
        if not (np.isfinite(dz).all() and np.isfinite(dl).all()):
            break
        a = min(step(s, ds), step(lam, dl))
        z = z + a * dz
        y = y + a * dy
        s = s + a * ds
        lam = lam + a * dl
    # duals in OSQP's row convention: y>0 on an active upper bound, y<0 on an active lower bound
    ydual = np.zeros(A.shape[0])
    ydual[eq] = y
    nu_ = int(up.sum())
    yu = np.zeros(A.shape[0])
    yu[up] = lam[:nu_]
    yl = np.zeros(A.shape[0])
    yl[lo] = lam[nu_:]
    ydual += yu - yl
    return z, ydual, it


def polish(P, q, A, l, u, z, ydual, delta=1e-9, refine=8, act_tol=1e-6, max_passes=12):
    """Active-set polish of an interior-point answer (same idea as OSQP's polish step).

    Rows whose dual is clearly non-zero (and equalities) are imposed as equalities; the
    regularised KKT system is solved with iterative refinement.  Rows that come back with a
    wrong-sign multiplier are released and violated rows are added (a few active-set passes
    starting from the interior-point guess).  Returns (z, y, ok); ok means primal feasible to
    1e-9 with correctly signed multipliers, i.e. the point is the exact optimum.
    """
    P = sp.csc_matrix(P)
    Pu = sp.triu(P, format="csc")
    P = (Pu + sp.triu(Pu, 1).T).tocsc()
    A = sp.csr_matrix(A)
    l = np.asarray(l, dtype=float)
    u = np.asarray(u, dtype=float)
    n = P.shape[0]
    eq = np.isfinite(l) & np.isfinite(u) & (l == u)
    scale = max(1.0, np.abs(ydual).max())
    act_u = (~eq) & (ydual > act_tol * scale)
    act_l = (~eq) & (ydual < -act_tol * scale)
    zp, yp, ok = z, ydual, False
    for _ in range(max_passes):
        rows = np.where(eq | act_u | act_l)[0]
        Aa = A[rows]
        ba = np.where(act_l[rows], l[rows], u[rows])
        ma = Aa.shape[0]
        K = sp.bmat([[P + delta * sp.eye(n), Aa.T], [Aa, -delta * sp.eye(ma)]], format="csc")
        lu = spla.splu(K)
        Kexact = sp.bmat([[P, Aa.T], [Aa, None]], format="csr")
        rhs = np.concatenate([-q, ba])
        sol = lu.solve(rhs)
        for _r in range(refine):
            sol = sol + lu.solve(rhs - Kexact @ sol)
        zp = sol[:n]
        yp = np.zeros(A.shape[0])
        yp[rows] = sol[n:]
        Az = A @ zp
        viol_u = (~eq) & (~act_u) & (Az - u > 1e-9)
        viol_l = (~eq) & (~act_l) & (l - Az > 1e-9)
        bad_u = act_u & (yp < 0.0)
        bad_l = act_l & (yp > 0.0)
        if not (viol_u.any() or viol_l.any() or bad_u.any() or bad_l.any()):
            ok = True
            break
        act_u = (act_u & ~bad_u) | viol_u
        act_l = (act_l & ~bad_l) | viol_l
    return zp, yp, ok


def solve_qp(P, q, A, l, u):
    """Oracle entry: interior point + polish; returns (z, y, info dict)."""
    z, y, it = solve_qp_ipm(P, q, A, l, u, tol=1e-11)
    zp, yp, ok = polish(P, q, A, l, u, z, y)
    info = {"ipm_iters": it, "polished": ok}
    if ok:
        return zp, yp, info
    return z, y, info


def kkt_residuals(P, q, A, l, u, z, ydual=None):
    """Solver-independent certificate: primal infeasibility and (if duals given) stationarity."""
    P = sp.csc_matrix(P)
    Pu = sp.triu(P, format="csc")
    P = (Pu + sp.triu(Pu, 1).T).tocsc()
    Az = sp.csr_matrix(A) @ z
    prim = max(np.max(np.maximum(Az - u, 0.0)), np.max(np.maximum(l - Az, 0.0)))
    out = {"primal": float(prim), "objective": float(0.5 * z @ (P @ z) + q @ z)}
    if ydual is not None:
        out["dual"] = float(np.abs(P @ z + q + sp.csr_matrix(A).T @ ydual).max())
    return out
