"""TEST INFRASTRUCTURE: a stand-in for `_bmpc.batch.BatchedBranchMPC` backed by the single-lane host build of the solver
text (tests/hostsim), so that the drop-in Python modules can be driven end to end on a CPU-only box - in particular by the
reference's UNMODIFIED entry scripts and environments (tests/test_reference_entry_scripts.py).  The product package
never imports this; without the patch the drop-in classes load libbranchmpc.so and need a GPU."""
import ctypes as C

import numpy as np

from _bmpc import abi
from tests.hostsim.driver import HostSim, lib, _ptr


class HostBackend:
    def __init__(self, cfg):
        self.cfg = cfg
        self.capacity = cfg.batch_capacity
        self._hs = HostSim(cfg, cfg.batch_capacity)
        self.nbranch, self.totalx, self.totalu = self._hs.nbranch, self._hs.totalx, self._hs.totalu

    def close(self):
        pass

    def topology(self):
        arr = [np.zeros(self.nbranch, dtype=np.int32) for _ in range(4)]
        rc = lib().hostsim_topology(C.byref(self.cfg), *[_ptr(a) for a in arr])
        assert rc == 0
        ndx, ndu, depth, parent = arr
        return np.column_stack([np.arange(self.nbranch), depth, ndx, ndu, parent]).astype(np.int64)

    def reset(self, episode_ids=None):
        ids = range(self.capacity) if episode_ids is None else episode_ids
        for e in ids:
            self._hs.uLin[e] = 0
            self._hs.pbest[e] = 0
            self._hs.oldin[e] = 0
            self._hs.started[e] = 0
            self._hs.cache_state[e] = -1

    def solve_host(self, x0, z0, xref, policy_params=None, outputs=tuple(k for k in abi.OUTPUT_NAMES if k != "bPred")):
        r = self._hs.solve(x0, z0, xref, policy_params)
        return {k: r[k] for k in outputs}

    solve_host_views = solve_host

    def set_lookup_table(self, xs, ys):
        xs = np.ascontiguousarray(xs, np.float64).reshape(-1)
        ys = np.ascontiguousarray(ys, np.float64).reshape(-1)
        assert lib().hostsim_set_lookup(_ptr(xs), _ptr(ys), C.c_int32(len(xs))) == 0

    def solve_transformed_host_views(self, x0, z0, xref, S=None, state_bounds=None, policy_params=None,
                                     outputs=tuple(k for k in abi.OUTPUT_NAMES if k != "bPred")):
        r = self._hs.solve_transformed(x0, z0, xref, S, state_bounds, policy_params)
        return {k: r[k] for k in outputs}

    def eval_model(self, x, z, u, policy_params=None):
        n, d, m, N = self.cfg.n, self.cfg.d, self.cfg.m, self.cfg.N
        x = np.ascontiguousarray(x, np.float64)
        z = np.ascontiguousarray(z, np.float64)
        u = np.ascontiguousarray(u, np.float64)
        K = x.shape[0]
        pp = None if policy_params is None else np.ascontiguousarray(policy_params, np.float64).reshape(K, m, 4)
        o = {"A": (K, n, n), "B": (K, n, d), "C": (K, n), "xp": (K, n), "zpred": (K, N, m * n), "p": (K, m), "hlin": (K,),
             "dh": (K, n)}
        t = {k: np.zeros(s) for k, s in o.items()}
        rc = lib().hostsim_eval_model(C.byref(self.cfg), _ptr(x), _ptr(z), _ptr(u), _ptr(pp), C.c_int64(K), _ptr(t["A"]),
                                      _ptr(t["B"]), _ptr(t["C"]), _ptr(t["xp"]), _ptr(t["zpred"]), _ptr(t["p"]),
                                      _ptr(t["hlin"]), _ptr(t["dh"]))
        assert rc == 0
        return t
