"""TEST INFRASTRUCTURE: ctypes driver of the single-lane host build of the solver text (tests/hostsim/hostsim.cpp).

Used by the CPU tests to check the device algorithm against the oracle without a GPU; the product package never
imports this.
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
PKG = os.path.join(ROOT, "belief-planning_b200")
if PKG not in sys.path:
    sys.path.insert(0, PKG)

from _bmpc import abi  # noqa: E402

SO = os.path.join(HERE, "libbmpc_hostsim.so")


def build(force=False):
    srcs = [os.path.join(HERE, "hostsim.cpp")] + [os.path.join(PKG, "csrc", f) for f in os.listdir(os.path.join(PKG, "csrc"))
                                                  if f.endswith(".h")] + [os.path.join(ROOT, "include", "branchmpc.h")]
    if not force and os.path.exists(SO) and all(os.path.getmtime(SO) >= os.path.getmtime(s) for s in srcs):
        return SO
    cmd = ["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-Wno-unknown-pragmas", "-I", os.path.join(ROOT, "include"),
           "-I", os.path.join(PKG, "csrc"), srcs[0], "-o", SO]
    subprocess.check_call(cmd)
    return SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.hostsim_last_error.restype = C.c_char_p
    return _lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class HostSim:
    """Batched solver with persistent warm-start state, host arrays only."""

    def __init__(self, cfg, capacity):
        self.cfg = cfg
        nb, tx, tu = C.c_int32(), C.c_int32(), C.c_int32()
        rc = lib().hostsim_sizes(C.byref(cfg), C.byref(nb), C.byref(tx), C.byref(tu))
        if rc != 0:
            raise RuntimeError("hostsim_sizes: %d %s" % (rc, lib().hostsim_last_error().decode()))
        self.nbranch, self.totalx, self.totalu = nb.value, tx.value, tu.value
        self.cap = capacity
        internal_u = self.totalu + (1 if cfg.controller in (abi.CTRL_ROBUST, abi.CTRL_BELIEF) else 0)   # robust chains carry a dummy stage
        self.uLin = np.zeros((capacity, internal_u + 1, cfg.d))
        self.xprev = np.zeros((capacity, self.totalx, cfg.n))
        self.pbest = np.zeros((capacity, self.nbranch), dtype=np.int32)
        self.oldin = np.zeros((capacity, cfg.d))
        self.started = np.zeros(capacity, dtype=np.int32)
        self.rho_cache = np.zeros(capacity * internal_u * 16)          # flat: the kernel strides by its own row width
        self.code_cache = np.zeros(capacity * (internal_u + 2), dtype=np.int64)   # per-episode stride padded to 16 bytes
        self.cache_state = np.full((capacity, 2), -1, dtype=np.int32)

    def set_state(self, uLin, pbest, old_input):
        """restore a warm-start state (single episode): what bmpc_set_state does on the device"""
        self.uLin[0, :len(uLin)] = uLin
        self.pbest[0] = pbest
        self.oldin[0] = old_input
        self.started[0] = 1
        self.cache_state[0] = -1

    def solve_belief(self, x0, b0, xbackup, xref):
        """PredictiveControllers.MPC.solve(x0, b0, xbackup, xRef) for a batch (bmpc_solve_belief on the device)."""
        b0 = np.ascontiguousarray(b0, dtype=float)
        xbackup = np.ascontiguousarray(xbackup, dtype=float)
        lib().hostsim_set_belief(_ptr(b0), _ptr(xbackup), C.c_int32(xbackup.shape[-1]))
        try:
            return self.solve(x0, np.atleast_2d(x0), xref, belief=True)
        finally:
            lib().hostsim_set_belief(None, None, C.c_int32(0))

    def solve_transformed(self, x0, z0, xref, S=None, state_bounds=None, policy_params=None):
        """BranchMPC_CVaR.solve(x, z, xRef, S, Fx=None, bx) for a batch (bmpc_solve_transformed on the device)."""
        S = None if S is None else np.ascontiguousarray(S, dtype=float)
        bd = None if state_bounds is None else np.ascontiguousarray(state_bounds, dtype=float)
        lib().hostsim_set_transform(_ptr(S), _ptr(bd))
        try:
            return self.solve(x0, z0, xref, policy_params)
        finally:
            lib().hostsim_set_transform(None, None)

    def solve(self, x0, z0, xref, policy_params=None, belief=False):
        cfg = self.cfg
        x0 = np.ascontiguousarray(np.atleast_2d(x0), dtype=float)
        z0 = np.ascontiguousarray(np.atleast_2d(z0), dtype=float)
        xref = np.ascontiguousarray(np.atleast_2d(xref), dtype=float)
        B = x0.shape[0]
        assert B <= self.cap
        pp = None if policy_params is None else np.ascontiguousarray(policy_params, dtype=float).reshape(B, cfg.m, 4)
        res = {
            "u0": np.zeros((B, cfg.d)), "uPred": np.zeros((B, self.totalu, cfg.d)),
            "xPred": np.zeros((B, self.totalx, cfg.n)), "xLin": np.zeros((B, self.totalu, cfg.n)),
            "zPred": np.zeros((B, self.totalu, cfg.n)), "branch_w": np.zeros((B, self.nbranch)),
            "branch_p": np.full((B, self.nbranch, cfg.m), np.nan), "objective": np.zeros(B),
            "status": np.full(B, -1, dtype=np.int32), "iters": np.zeros(B, dtype=np.int32),
            "nfact": np.zeros(B, dtype=np.int32), "nsolve": np.zeros(B, dtype=np.int32), "cycles": np.zeros(B, dtype=np.int64),
        }
        if belief:
            res["bPred"] = np.zeros((B, self.totalx, cfg.hmm_M * cfg.m))
        out = abi.Outputs(**{k: _ptr(v) for k, v in res.items()})
        rc = lib().hostsim_solve(C.byref(cfg), _ptr(x0), _ptr(z0), _ptr(xref), _ptr(pp), C.c_int64(B), _ptr(self.uLin),
                                 _ptr(self.pbest), _ptr(self.oldin), _ptr(self.started), _ptr(self.rho_cache),
                                 _ptr(self.code_cache), _ptr(self.cache_state), _ptr(self.xprev), C.byref(out))
        if rc != 0:
            raise RuntimeError("hostsim_solve: %d %s" % (rc, lib().hostsim_last_error().decode()))
        return res
