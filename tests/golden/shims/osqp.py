"""Stand-in for the `osqp` wheel (absent here).  TEST INFRASTRUCTURE ONLY.

`OSQP().setup(P, q, A, l, u, ...)` / `.solve()` keep the call shape the reference uses
(`/root/reference/MPC_branch.py:1263-1274`) but the QP is solved to its exact optimum by
the oracle's interior-point + polish solver (oracle/qp_exact.py, KKT-certified) and, as an
independent cross-check, by HiGHS.  The captured problem data (`last_problem`) is what the golden
fixtures store: the matrices the UNMODIFIED reference assembled.
"""
import os
import sys
import numpy as np
import scipy.sparse as sp

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", ".."))
from oracle.qp_exact import solve_qp, solve_qp_highs, kkt_residuals  # noqa: E402

last_problem = {}


class _Info:
    status_val = 1
    status = "solved"


class _Res:
    pass


class OSQP:
    def setup(self, P=None, q=None, A=None, l=None, u=None, **kw):
        self.P = sp.csc_matrix(P)
        self.q = np.asarray(q, dtype=float)
        self.A = sp.csc_matrix(A)
        self.l = np.asarray(l, dtype=float)
        self.u = np.asarray(u, dtype=float)
        self.settings = kw
        last_problem.update(P=self.P, q=self.q, A=self.A, l=self.l, u=self.u, settings=kw)

    def warm_start(self, x=None, y=None):
        pass

    def solve(self):
        z, ydual, info = solve_qp(self.P, self.q, self.A, self.l, self.u)
        cert = kkt_residuals(self.P, self.q, self.A, self.l, self.u, z, ydual)
        ok = info["polished"] and cert["primal"] < 1e-8 and cert["dual"] < 1e-7
        zh, _, hstatus = solve_qp_highs(self.P, self.q, self.A, self.l, self.u)
        hcert = kkt_residuals(self.P, self.q, self.A, self.l, self.u, zh)
        res = _Res()
        res.x = z
        res.y = ydual
        res.info = _Info()
        res.info.status_val = 1 if ok else -3
        res.info.status = "solved" if ok else "unsolved"
        last_problem.update(x=z, y=ydual, cert=cert, ok=ok, highs_x=zh, highs_status=hstatus,
                            highs_objective=hcert["objective"])
        return res
