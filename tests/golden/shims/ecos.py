"""Stand-in for the `ecos` wheel (absent here).  TEST INFRASTRUCTURE ONLY.

`ecos.solve(c, G, h, dims, A, b, verbose=False)` keeps the call shape the reference uses
(`/root/reference/MPC_branch.py:2136`); the cone program is solved by the oracle's primal-dual interior point
(oracle/socp.py, tolerance 1e-9 on residuals and gap).  The captured problem data (`last_problem`) is what the golden
fixtures store: the matrices the UNMODIFIED reference `BranchMPC_CVaR` assembled.
"""
import os
import sys

import numpy as np
import scipy.sparse as sp

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", ".."))
from oracle import socp  # noqa: E402

last_problem = {}


def solve(c, G, h, dims, A=None, b=None, verbose=False, **kw):
    c = np.asarray(c, dtype=float)
    G = sp.csc_matrix(G)
    h = np.asarray(h, dtype=float)
    A = None if A is None else sp.csc_matrix(A)
    b = None if b is None else np.asarray(b, dtype=float)
    dims = {"l": int(dims["l"]), "q": [int(v) for v in dims["q"]]}
    sol = socp.solve(c, G, h, dims, A, b)
    last_problem.clear()
    last_problem.update(c=c, G=G, h=h, dims=dims, A=A, b=b, x=sol["x"], info=dict(sol["info"]))
    info = dict(sol["info"])
    info["exitFlag"] = 0 if info["exitFlag"] == 0 else -1
    return {"x": sol["x"], "y": sol["y"], "z": sol["z"], "s": sol["s"], "info": info}
