"""TEST INFRASTRUCTURE (oracle): primal-dual interior-point solver for the cone programs `BranchMPC_CVaR` hands to ECOS.

    minimise  c'x   subject to   G x + s = h,  s in K,   A x = b            (the argument order of ecos.solve,
                                                                              /root/reference/MPC_branch.py:2136)
    K = R+^l  x  Q^{q_1} x ... x Q^{q_N}   (`dims = {'l': l, 'q': [q_1, ...]}`, second-order cones (t, v): t >= |v|)

ECOS itself (un-vendored third-party C, version unpinned by the reference) is not installable here; it solves these
problems to tolerances of 1e-8, so "the reference's answer" is the optimum of the cone program the reference assembles,
which is what this solver returns (Mehrotra predictor-corrector with Nesterov-Todd scaling, the algorithm of the ECOS paper
and of CVXOPT's coneqp, restated from the published descriptions).  The objective and every variable the objective
determines (first input, trajectories of the branches that carry risk weight) are unique; variables that the program leaves
free (slacks of zero rows, branches with zero risk weight) are not, and parity tests do not compare them.

Presolve: a variable with zero cost that only occurs in inequality rows of the form  -x_j <= 0  (the slacks of the terminal
state nodes, whose constraint rows are all-zero: MPC_branch.py:1869-1890 loops over the input nodes only) has an unbounded
optimal set and no central path; it is fixed at 0 and its rows are dropped.
"""
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla


class _Cone:
    def __init__(self, l, q):
        self.l = int(l)
        self.q = [int(v) for v in q]
        self.m = self.l + sum(self.q)
        self.starts = np.cumsum([self.l] + self.q[:-1]) if self.q else np.array([], dtype=int)

    def soc(self):
        for s0, n in zip(self.starts, self.q):
            yield int(s0), int(n)

    def degree(self):
        return self.l + len(self.q)

    def identity(self):
        e = np.zeros(self.m)
        e[: self.l] = 1.0
        for s0, _ in self.soc():
            e[s0] = 1.0
        return e

    def prod(self, u, v):
        """Jordan product u o v."""
        w = np.empty(self.m)
        w[: self.l] = u[: self.l] * v[: self.l]
        for s0, n in self.soc():
            w[s0] = u[s0:s0 + n] @ v[s0:s0 + n]
            w[s0 + 1:s0 + n] = u[s0] * v[s0 + 1:s0 + n] + v[s0] * u[s0 + 1:s0 + n]
        return w

    def div(self, lam, d):
        """x with lam o x = d."""
        x = np.empty(self.m)
        x[: self.l] = d[: self.l] / lam[: self.l]
        for s0, n in self.soc():
            l0, l1 = lam[s0], lam[s0 + 1:s0 + n]
            d0, d1 = d[s0], d[s0 + 1:s0 + n]
            det = l0 * l0 - l1 @ l1
            x0 = (l0 * d0 - l1 @ d1) / det
            x[s0] = x0
            x[s0 + 1:s0 + n] = (d1 - x0 * l1) / l0
        return x

    def max_step(self, s, ds):
        """largest a in (0, inf] with s + a ds in K."""
        a = np.inf
        neg = ds[: self.l] < 0
        if neg.any():
            a = min(a, float(np.min(-s[: self.l][neg] / ds[: self.l][neg])))
        for s0, n in self.soc():
            u0, u1 = s[s0], s[s0 + 1:s0 + n]
            v0, v1 = ds[s0], ds[s0 + 1:s0 + n]
            # (u0 + a v0)^2 - |u1 + a v1|^2 >= 0 and u0 + a v0 >= 0
            qa = v0 * v0 - v1 @ v1
            qb = 2.0 * (u0 * v0 - u1 @ v1)
            qc = u0 * u0 - u1 @ u1
            roots = []
            if abs(qa) > 1e-300:
                disc = qb * qb - 4 * qa * qc
                if disc >= 0:
                    sq = np.sqrt(disc)
                    qq = -0.5 * (qb + np.copysign(sq, qb))
                    roots = [qq / qa]
                    if qq != 0:
                        roots.append(qc / qq)
            elif abs(qb) > 1e-300:
                roots = [-qc / qb]
            pos = [r for r in roots if r > 0]
            if v0 < 0:
                pos.append(-u0 / v0)
            if pos:
                a = min(a, min(pos))
        return a

    def interior_shift(self, s):
        """smallest t >= 0 (plus margin) such that s + t e is strictly inside K."""
        t = 0.0
        if self.l:
            t = max(t, float(-np.min(s[: self.l])))
        for s0, n in self.soc():
            t = max(t, float(np.linalg.norm(s[s0 + 1:s0 + n]) - s[s0]))
        return t

    def scaling(self, s, z):
        """Nesterov-Todd scaling: returns (W^2 as sparse block matrix, apply_W, apply_Winv, lam = W z = W^{-1} s)."""
        d = np.sqrt(s[: self.l] / z[: self.l])
        blocks = [sp.diags(d * d)] if self.l else []
        soc = []
        for s0, n in self.soc():
            sk, zk = s[s0:s0 + n], z[s0:s0 + n]
            sres = np.sqrt(sk[0] ** 2 - sk[1:] @ sk[1:])
            zres = np.sqrt(zk[0] ** 2 - zk[1:] @ zk[1:])
            sb, zb = sk / sres, zk / zres
            gamma = np.sqrt(0.5 * (1.0 + sb @ zb))
            wbar = np.empty(n)
            wbar[0] = (sb[0] + zb[0]) / (2 * gamma)
            wbar[1:] = (sb[1:] - zb[1:]) / (2 * gamma)
            eta = np.sqrt(sres / zres)
            Wm = np.empty((n, n))
            Wm[0, 0] = wbar[0]
            Wm[0, 1:] = wbar[1:]
            Wm[1:, 0] = wbar[1:]
            Wm[1:, 1:] = np.eye(n - 1) + np.outer(wbar[1:], wbar[1:]) / (1.0 + wbar[0])
            Wm *= eta
            Wi = np.empty((n, n))
            Wi[0, 0] = wbar[0]
            Wi[0, 1:] = -wbar[1:]
            Wi[1:, 0] = -wbar[1:]
            Wi[1:, 1:] = np.eye(n - 1) + np.outer(wbar[1:], wbar[1:]) / (1.0 + wbar[0])
            Wi /= eta
            soc.append((s0, n, Wm, Wi))
            blocks.append(sp.csc_matrix(Wm @ Wm))
        W2 = sp.block_diag(blocks, format="csc") if blocks else sp.csc_matrix((0, 0))

        def apply(v, inv=False):
            o = np.empty(self.m)
            o[: self.l] = v[: self.l] / d if inv else v[: self.l] * d
            for s0, n, Wm, Wi in soc:
                o[s0:s0 + n] = (Wi if inv else Wm) @ v[s0:s0 + n]
            return o

        lam = apply(z)
        return W2, apply, lam


def _presolve_free_slacks(c, G, h, A, l):
    """columns with zero cost, absent from A, whose only entries sit in LP rows `-x_j <= 0` -> fixed at 0."""
    Gc = sp.csc_matrix(G)
    Gr = sp.csr_matrix(G)
    row_nnz = np.diff(Gr.indptr)
    in_A = np.zeros(G.shape[1], dtype=bool)
    if A is not None and A.shape[0]:
        in_A[np.diff(sp.csc_matrix(A).indptr) > 0] = True
    drop_cols, drop_rows = [], []
    for j in np.flatnonzero((c == 0) & ~in_A):
        rows = Gc.indices[Gc.indptr[j]:Gc.indptr[j + 1]]
        vals = Gc.data[Gc.indptr[j]:Gc.indptr[j + 1]]
        if len(rows) and (rows < l).all() and (vals < 0).all() and (row_nnz[rows] == 1).all() and (h[rows] == 0).all():
            drop_cols.append(j)
            drop_rows.extend(rows.tolist())
    return np.array(drop_cols, dtype=int), np.array(sorted(set(drop_rows)), dtype=int)


def rotated_blocks(G, h, dims):
    """BranchMPC_CVaR writes every cone as  |(2 W y, 1 + c(x))| <= 1 - c(x)  (MPC_branch.py:1936-1990: rows F1, F2, F3 = -F1,
    right-hand sides 1 - J, 0, 1 + J), which is the convex quadratic constraint  |W y|^2 + c(x) <= 0  in disguise.  Returns,
    per cone, (g0, Gv, k) with the constraint  x' Gv' Gv x + 4 g0 x + k <= 0,  or None if some cone is not of that form."""
    G = sp.csr_matrix(G)
    out = []
    row = int(dims.get("l", 0))
    for n in dims.get("q", []):
        blk = G[row:row + n]
        hb = h[row:row + n]
        g0, gl, Gv = blk[0], blk[n - 1], blk[1:n - 1]
        if abs(hb[0] + hb[n - 1] - 2.0) > 1e-12 or np.abs(hb[1:n - 1]).max(initial=0.0) > 0 or abs(g0 + gl).max() > 0:
            return None
        # s0^2 - sl^2 = (s0 - sl)(s0 + sl) = 2 (h0 - hl - 2 g0 x)  >=  |Gv x|^2
        out.append((g0.tocsr(), Gv.tocsr(), float(-2.0 * (hb[0] - hb[n - 1]))))
        row += n
    return out


def solve_qcqp(c, Gl, hl, quads, A, b, tol=1e-9, max_iter=300, verbose=False, prox=1e-8):
    """min c'x  s.t.  A x = b,  Gl x <= hl,  x' Gv_i' Gv_i x + 4 g0_i x + k_i <= 0.
    Primal-dual interior point (Mehrotra predictor-corrector) with one explicit slack per inequality; the slack of a quadratic
    constraint is carried as a variable of its own, so an active constraint costs no cancellation (the second-order-cone
    form of the same constraint stores it as the difference of two numbers of size ~|c(x)|)."""
    n = len(c)
    nl, nq, p = Gl.shape[0], len(quads), A.shape[0]
    Gl = sp.csr_matrix(Gl)
    Pq = [(2.0 * (Gv.T @ Gv)).tocsc() for _, Gv, _ in quads]
    lq = [4.0 * np.asarray(g0.todense()).ravel() for g0, _, _ in quads]
    kq = np.array([k for _, _, k in quads])

    def gvals(x):
        gq = np.array([0.5 * x @ (Pq[i] @ x) + lq[i] @ x + kq[i] for i in range(nq)])
        return np.concatenate([Gl @ x - hl, gq])

    def jac(x):
        rows = [sp.csr_matrix((Pq[i] @ x + lq[i])[None, :]) for i in range(nq)]
        return sp.vstack([Gl] + rows, format="csr") if nq else Gl

    # start: minimum-norm point of the equalities, slacks that make every inequality hold with room
    x = spla.lsqr(A, b, atol=1e-14, btol=1e-14)[0] if p else np.zeros(n)
    g = gvals(x)
    s = np.maximum(-g, 1.0)
    lam = np.ones(nl + nq)
    y = np.zeros(p)
    m = nl + nq
    nrm_c = max(1.0, np.linalg.norm(c))
    info, flag = {}, -1
    for it in range(max_iter + 1):
        g = gvals(x)
        J = jac(x)
        r_d = c + A.T @ y + J.T @ lam
        r_p = A @ x - b
        r_s = g + s
        mu = float(s @ lam) / m
        pcost = float(c @ x)
        # slack rows are measured relative to their slack: rows on an unbounded optimal face (branches without risk weight
        # and the risk variables that absorb their cost) drift with slacks of 1e4 and more without touching the objective
        pres = max(np.abs(r_p).max(initial=0.0), (np.abs(r_s) / (1.0 + s)).max(initial=0.0))
        dres = np.abs(r_d).max() / nrm_c
        if verbose:
            print("%3d pcost % .10e mu %.2e pres %.1e dres %.1e" % (it, pcost, mu, pres, dres))
        info = {"pcost": pcost, "dcost": pcost - m * mu, "gap": m * mu, "pres": float(pres), "dres": float(dres), "iter": it}
        if pres <= 100 * tol and dres <= 100 * tol and m * mu <= tol * (1.0 + abs(pcost)):
            flag = 0
            break
        if it == max_iter:
            break
        H = sp.csc_matrix((n, n))
        for i in range(nq):
            H = H + lam[nl + i] * Pq[i]
        D = sp.diags(lam / s)
        # proximal regularisation about the current iterate: leaves the optimum alone, damps the drift along unbounded optimal faces
        K = sp.bmat([[H + J.T @ D @ J + prox * sp.eye(n), A.T], [A, -1e-12 * sp.eye(p)]], format="csc")
        lu = spla.splu(K)

        def newton(rc):
            rhs = np.concatenate([-r_d - J.T @ ((rc + lam * r_s) / s), -r_p])
            sol = lu.solve(rhs)
            sol = sol + lu.solve(rhs - K @ sol)
            dx, dy = sol[:n], sol[n:]
            ds = -r_s - J @ dx
            dl = (rc - lam * ds) / s
            return dx, dy, ds, dl

        def max_step(v, dv):
            neg = dv < 0
            return min(1.0, float(np.min(-v[neg] / dv[neg]))) if neg.any() else 1.0

        dxa, dya, dsa, dla = newton(-s * lam)
        aa = min(max_step(s, dsa), max_step(lam, dla))
        mu_aff = float((s + aa * dsa) @ (lam + aa * dla)) / m
        sigma = (mu_aff / mu) ** 3
        # the barrier parameter is not driven below what the gap test asks for: past that point the steps only remove the
        # residuals (a long horizon's KKT matrix loses the late iterations to conditioning otherwise)
        mu_floor = 0.1 * tol * (1.0 + abs(pcost)) / m
        if sigma * mu < mu_floor:
            dx, dy, ds, dl = newton(max(mu_floor, 0.0) - s * lam)
        else:
            dx, dy, ds, dl = newton(sigma * mu - s * lam - dsa * dla)
        a_p = min(1.0, 0.995 * max_step(s, ds) if (ds < 0).any() else 1.0)
        a_d = min(1.0, 0.995 * max_step(lam, dl) if (dl < 0).any() else 1.0)
        a = min(a_p, a_d)
        x, y, s, lam = x + a * dx, y + a * dy, s + a * ds, lam + a * dl
        # the quadratic constraints are not linear in the step: keep their slacks consistent once the iterate is feasible
        gn = gvals(x)
        ok = (-gn > 0) & (np.arange(m) >= nl)
        s = np.where(ok & (np.abs(gn + s) > 0), np.maximum(-gn, 1e-300), s)
    info["exitFlag"] = flag
    return {"x": x, "y": y, "lam": lam, "s": s, "info": info}


def solve(c, G, h, dims, A=None, b=None, tol=1e-9, max_iter=80, verbose=False):
    """Returns {'x', 'y', 'z', 's', 'info': {'exitFlag', 'pcost', 'dcost', 'gap', 'pres', 'dres', 'iter'}}.
    Cone programs whose cones are all rotated quadratic constraints (everything BranchMPC_CVaR builds) go through the
    quadratic-constraint interior point, the rest through the Nesterov-Todd one."""
    c = np.asarray(c, dtype=float).ravel()
    h = np.asarray(h, dtype=float).ravel()
    quads = rotated_blocks(sp.csr_matrix(G), h, dims) if dims.get("q") else None
    if quads is not None:
        l0 = int(dims.get("l", 0))
        Gs = sp.csc_matrix(G)
        n_full = Gs.shape[1]
        As = sp.csc_matrix((0, n_full)) if A is None else sp.csc_matrix(A)
        bs = np.zeros(0) if b is None else np.asarray(b, dtype=float).ravel()
        dcols, drows = _presolve_free_slacks(c, Gs, h, As, l0)
        keep_c = np.setdiff1d(np.arange(n_full), dcols)
        keep_r = np.setdiff1d(np.arange(l0), drows)
        Gl = sp.csr_matrix(Gs)[keep_r][:, keep_c]
        q2 = [(g0[:, keep_c], Gv[:, keep_c], k) for g0, Gv, k in quads]
        r = solve_qcqp(c[keep_c], Gl, h[keep_r], q2, As[:, keep_c], bs, tol=tol, max_iter=max(max_iter, 300), verbose=verbose)
        xf = np.zeros(n_full)
        xf[keep_c] = r["x"]
        return {"x": xf, "y": r["y"], "z": None, "s": None, "lam_quad": r["lam"][len(keep_r):], "info": r["info"]}
    return solve_nt(c, G, h, dims, A, b, tol=tol, max_iter=max_iter, verbose=verbose)


def solve_nt(c, G, h, dims, A=None, b=None, tol=1e-9, max_iter=80, verbose=False):
    """Nesterov-Todd interior point for general second-order-cone programs (see the module docstring)."""
    c = np.asarray(c, dtype=float).ravel()
    h = np.asarray(h, dtype=float).ravel()
    G = sp.csc_matrix(G)
    n_full = G.shape[1]
    A = sp.csc_matrix((0, n_full)) if A is None else sp.csc_matrix(A)
    b = np.zeros(0) if b is None else np.asarray(b, dtype=float).ravel()
    l0 = int(dims.get("l", 0))
    q = list(dims.get("q", []))

    dcols, drows = _presolve_free_slacks(c, G, h, A, l0)
    keep_c = np.setdiff1d(np.arange(n_full), dcols)
    keep_r = np.setdiff1d(np.arange(G.shape[0]), drows)
    G = G[keep_r][:, keep_c].tocsc()
    A = A[:, keep_c].tocsc()
    h = h[keep_r]
    c = c[keep_c]
    cone = _Cone(l0 - len(drows), q)
    n, p, m = G.shape[1], A.shape[0], G.shape[0]
    assert cone.m == m

    def kkt_solve(W2, rx, ry, rz):
        # [0 A' G'; A 0 0; G 0 -W2] [x;y;z] = [rx; ry; rz], with a small static regularisation and iterative refinement
        eps = 1e-11
        K = sp.bmat([[eps * sp.eye(n), A.T, G.T], [A, -eps * sp.eye(p), None], [G, None, -W2]], format="csc")
        Kx = sp.bmat([[None, A.T, G.T], [A, None, None], [G, None, -W2]], format="csc") if p else \
            sp.bmat([[None, G.T], [G, -W2]], format="csc")
        if not p:
            K = sp.bmat([[eps * sp.eye(n), G.T], [G, -W2]], format="csc")
        lu = spla.splu(K)
        rhs = np.concatenate([rx, ry, rz])
        sol = lu.solve(rhs)
        for _ in range(3):
            r = rhs - Kx @ sol
            if np.abs(r).max() <= 1e-13 * (1.0 + np.abs(rhs).max()):
                break
            sol = sol + lu.solve(r)
        return sol[:n], sol[n:n + p], sol[n + p:]

    # starting point (CVXOPT coneqp): least-squares solution, shifted into the cone
    e = cone.identity()
    x, y, z = kkt_solve(sp.eye(m, format="csc"), -c, b, h)
    s = -z
    ts = cone.interior_shift(s)
    s = s + (1.0 + ts) * e if ts >= -1e-8 * max(1.0, np.linalg.norm(s)) else s
    tz = cone.interior_shift(z)
    z = z + (1.0 + tz) * e if tz >= -1e-8 * max(1.0, np.linalg.norm(z)) else z

    nrm_c, nrm_b, nrm_h = max(1.0, np.linalg.norm(c)), max(1.0, np.linalg.norm(b)), max(1.0, np.linalg.norm(h))
    info = {}
    flag = -1
    for it in range(max_iter + 1):
        rx = -(A.T @ y + G.T @ z + c)
        ry = b - A @ x
        rz = h - G @ x - s
        gap = float(s @ z)
        pcost = float(c @ x)
        dcost = float(-(b @ y) - (h @ z))
        pres = max(np.linalg.norm(ry) / nrm_b, np.linalg.norm(rz) / nrm_h)
        dres = np.linalg.norm(rx) / nrm_c
        relgap = gap / max(1e-300, min(abs(pcost), abs(dcost))) if min(abs(pcost), abs(dcost)) > 0 else np.inf
        if verbose:
            print("%3d pcost % .8e dcost % .8e gap %.1e pres %.1e dres %.1e" % (it, pcost, dcost, gap, pres, dres))
        info = {"pcost": pcost, "dcost": dcost, "gap": gap, "pres": float(pres), "dres": float(dres), "iter": it}
        if pres <= tol and dres <= tol and (gap <= tol or relgap <= tol):
            flag = 0
            break
        if it == max_iter:
            break
        W2, Wap, lam = cone.scaling(s, z)
        mu = gap / cone.degree()

        def direction(ds_rhs):
            # W^{-T} ds + W dz = lam \ ds_rhs ;  G dx + ds = rz ;  A dx = ry ;  A' dy + G' dz = rx
            t = Wap(cone.div(lam, ds_rhs))          # W (lam \ ds_rhs) = ds + W^2 dz
            dx, dy, dz = kkt_solve(W2, rx, ry, rz - t)
            ds = t - W2 @ dz
            return dx, dy, dz, ds

        dxa, dya, dza, dsa = direction(-cone.prod(lam, lam))
        a = min(1.0, cone.max_step(s, dsa), cone.max_step(z, dza))
        sigma = (1.0 - a) ** 3
        corr = cone.prod(Wap(dsa, inv=True), Wap(dza))
        dx, dy, dz, ds = direction(-cone.prod(lam, lam) - corr + sigma * mu * e)
        a = min(1.0, 0.99 * min(cone.max_step(s, ds), cone.max_step(z, dz)))
        x, y, z, s = x + a * dx, y + a * dy, z + a * dz, s + a * ds

    xf = np.zeros(n_full)
    xf[keep_c] = x
    zf = np.zeros(len(keep_r) + len(drows))
    sf = np.zeros_like(zf)
    zf[keep_r] = z
    sf[keep_r] = s
    info["exitFlag"] = flag
    return {"x": xf, "y": y, "z": zf, "s": sf, "info": info}
