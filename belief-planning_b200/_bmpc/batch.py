"""Batched Branch-MPC solver handle: thousands of independent scenario-tree problems per call on one B200.

PyTorch is used only to own device memory and to name the CUDA stream; all arithmetic happens inside
libbranchmpc.so (csrc/), reached through the C ABI of include/branchmpc.h.
"""
import ctypes as C

import numpy as np

from . import abi

_OUT_SHAPES = {
    "u0": lambda s: (s.cfg.d,), "uPred": lambda s: (s.totalu, s.cfg.d), "xPred": lambda s: (s.totalx, s.cfg.n),
    "xLin": lambda s: (s.totalu, s.cfg.n), "zPred": lambda s: (s.totalu, s.cfg.n), "branch_w": lambda s: (s.nbranch,),
    "branch_p": lambda s: (s.nbranch, s.cfg.m), "objective": lambda s: (), "status": lambda s: (),
    "iters": lambda s: (), "nfact": lambda s: (), "nsolve": lambda s: (), "cycles": lambda s: (),
    "bPred": lambda s: (s.totalx, s.cfg.hmm_M * s.cfg.m),
}
# everything the tree controllers produce (bPred belongs to the belief-state MPC, bmpc_solve_belief)
ALL_OUTPUTS = tuple(k for k in abi.OUTPUT_NAMES if k != "bPred")
_INT_OUTPUTS = ("status", "iters", "nfact", "nsolve")
LIGHT_OUTPUTS = ("u0", "objective", "status", "iters", "nfact", "nsolve")


class BmpcError(RuntimeError):
    pass


class BatchedBranchMPC:
    """One libbranchmpc handle (one device, one configuration, `batch_capacity` persistent episode slots)."""

    def __init__(self, cfg):
        self.lib = abi.load_library()
        self.cfg = cfg
        h = C.c_void_p()
        rc = self.lib.bmpc_create(C.byref(cfg), C.byref(h))
        if rc != abi.OK:
            raise BmpcError("bmpc_create failed (%d): %s" % (rc, self.lib.bmpc_last_error(None).decode()))
        self.h = h
        self.nbranch = self.lib.bmpc_num_branches(h)
        self.totalx = self.lib.bmpc_total_x(h)
        self.totalu = self.lib.bmpc_total_u(h)
        self.ulin_rows = self.lib.bmpc_ulin_rows(h)
        self.capacity = cfg.batch_capacity
        self._dev_out = {}

    def close(self):
        if getattr(self, "h", None):
            self.lib.bmpc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc != abi.OK:
            raise BmpcError("%s failed (%d): %s" % (what, rc, self.lib.bmpc_last_error(self.h).decode()))

    # -- topology ------------------------------------------------------------------------------------------
    def topology(self):
        """(id, depth, ndx, ndu, parent) rows in BFS order, the numbering of MPC_branch.py:928-981."""
        arr = [np.zeros(self.nbranch, dtype=np.int32) for _ in range(4)]
        ptr = [a.ctypes.data_as(C.POINTER(C.c_int32)) for a in arr]
        self._check(self.lib.bmpc_get_topology(self.h, *ptr), "bmpc_get_topology")
        ndx, ndu, depth, parent = arr
        return np.column_stack([np.arange(self.nbranch), depth, ndx, ndu, parent]).astype(np.int64)

    def reset(self, episode_ids=None):
        if episode_ids is None:
            self._check(self.lib.bmpc_reset(self.h, None, 0), "bmpc_reset")
        else:
            ids = np.ascontiguousarray(episode_ids, dtype=np.int64)
            self._check(self.lib.bmpc_reset(self.h, ids.ctypes.data_as(C.POINTER(C.c_int64)), len(ids)), "bmpc_reset")

    # -- device path -----------------------------------------------------------------------------------------
    def device_outputs(self, count, names):
        import torch
        dev = torch.device("cuda", self.cfg.device)
        key = (count, tuple(names))
        if key not in self._dev_out:
            bufs = {}
            for k in names:
                dt = torch.int64 if k == "cycles" else (torch.int32 if k in _INT_OUTPUTS else torch.float64)
                bufs[k] = torch.zeros((count,) + _OUT_SHAPES[k](self), dtype=dt, device=dev)
            self._dev_out = {key: bufs}          # keep only the latest shape
        return self._dev_out[key]

    def solve(self, x0, z0, xref, policy_params=None, outputs=LIGHT_OUTPUTS, stream=None):
        """x0, z0, xref: CUDA float64 tensors (count, n).  Returns a dict of CUDA tensors (reused between calls).
        Asynchronous: the kernel is enqueued on `stream` (default: torch's current stream)."""
        import torch
        count = x0.shape[0]
        for t in (x0, z0, xref):
            if not (t.is_cuda and t.dtype == torch.float64 and t.is_contiguous() and t.shape == (count, self.cfg.n)):
                raise ValueError("inputs must be contiguous CUDA float64 tensors of shape (count, n)")
        pp = None
        if policy_params is not None:
            if not (policy_params.is_cuda and policy_params.dtype == torch.float64 and policy_params.is_contiguous()
                    and policy_params.numel() == count * self.cfg.m * 4):
                raise ValueError("policy_params must be a contiguous CUDA float64 tensor of (count, m, 4)")
            pp = policy_params.data_ptr()
        bufs = self.device_outputs(count, outputs)
        out = abi.Outputs(**{k: bufs[k].data_ptr() for k in bufs})
        if stream is None:
            stream = torch.cuda.current_stream(x0.device).cuda_stream
        self._check(self.lib.bmpc_solve(self.h, x0.data_ptr(), z0.data_ptr(), xref.data_ptr(), pp, count,
                                        C.byref(out), C.c_void_p(stream)), "bmpc_solve")
        return bufs

    def plant_step(self, x, u, z, obstacle_policy=0, policy_params=None, stream=None):
        """In-place Euler step of ego (input u) and obstacle (backup policy index) on CUDA tensors."""
        import torch
        count = x.shape[0]
        if stream is None:
            stream = torch.cuda.current_stream(x.device).cuda_stream
        pp = None if policy_params is None else policy_params.data_ptr()
        self._check(self.lib.bmpc_plant_step(self.h, x.data_ptr(), u.data_ptr(), None if z is None else z.data_ptr(),
                                             int(obstacle_policy), pp, count, C.c_void_p(stream)), "bmpc_plant_step")

    # -- host path (the call the drop-in classes make) ---------------------------------------------------------
    def solve_host(self, x0, z0, xref, policy_params=None, outputs=ALL_OUTPUTS):
        x0 = np.ascontiguousarray(np.atleast_2d(x0), dtype=np.float64)
        z0 = np.ascontiguousarray(np.atleast_2d(z0), dtype=np.float64)
        xref = np.ascontiguousarray(np.atleast_2d(xref), dtype=np.float64)
        count = x0.shape[0]
        if x0.shape != (count, self.cfg.n) or z0.shape != x0.shape or xref.shape != x0.shape:
            raise ValueError("x0, z0, xref must have shape (count, n)")
        pp = None
        if policy_params is not None:
            pp = np.ascontiguousarray(policy_params, dtype=np.float64).reshape(count, self.cfg.m, 4)
        res = {}
        for k in outputs:
            dt = np.int64 if k == "cycles" else (np.int32 if k in _INT_OUTPUTS else np.float64)
            res[k] = np.empty((count,) + _OUT_SHAPES[k](self), dtype=dt)
        out = abi.Outputs(**{k: v.ctypes.data for k, v in res.items()})
        self._check(self.lib.bmpc_solve_host(self.h, x0.ctypes.data, z0.ctypes.data, xref.ctypes.data,
                                             None if pp is None else pp.ctypes.data, count, C.byref(out)),
                    "bmpc_solve_host")
        return res

    def solve_host_views(self, x0, z0, xref, policy_params=None, outputs=ALL_OUTPUTS):
        """As solve_host, without the copy out of the library's pinned result block: the returned arrays are views that
        stay valid until this handle's next host solve (the drop-in controllers copy what they keep)."""
        x0 = np.ascontiguousarray(np.atleast_2d(x0), dtype=np.float64)
        z0 = np.ascontiguousarray(np.atleast_2d(z0), dtype=np.float64)
        xref = np.ascontiguousarray(np.atleast_2d(xref), dtype=np.float64)
        count = x0.shape[0]
        if x0.shape != (count, self.cfg.n) or z0.shape != x0.shape or xref.shape != x0.shape:
            raise ValueError("x0, z0, xref must have shape (count, n)")
        pp = None
        if policy_params is not None:
            pp = np.ascontiguousarray(policy_params, dtype=np.float64).reshape(count, self.cfg.m, 4)
        want = abi.Outputs(**{k: 1 for k in outputs})
        views = abi.Outputs()
        self._check(self.lib.bmpc_solve_host_views(self.h, x0.ctypes.data, z0.ctypes.data, xref.ctypes.data,
                                                   None if pp is None else pp.ctypes.data, count, C.byref(want),
                                                   C.byref(views)), "bmpc_solve_host_views")
        res = {}
        for k in outputs:
            dt = np.int64 if k == "cycles" else (np.int32 if k in _INT_OUTPUTS else np.float64)
            shape = (count,) + _OUT_SHAPES[k](self)
            nbytes = int(np.prod(shape)) * np.dtype(dt).itemsize
            buf = (C.c_char * nbytes).from_address(getattr(views, k))
            res[k] = np.frombuffer(buf, dtype=dt).reshape(shape)
        return res

    # -- merge scenario: BranchMPC_CVaR.solve(x, z, xRef, S, Fx=None, bx) -----------------------------------------
    def set_lookup_table(self, xs, ys):
        """psiref(x) of the ramp policies (casadi interpolant 'linear'): grid xs (strictly increasing) and values ys."""
        xs = np.ascontiguousarray(xs, np.float64).reshape(-1)
        ys = np.ascontiguousarray(ys, np.float64).reshape(-1)
        if xs.shape != ys.shape:
            raise ValueError("lookup grid and values disagree")
        self._check(self.lib.bmpc_set_lookup_table(self.h, xs.ctypes.data, ys.ctypes.data, len(xs)), "bmpc_set_lookup_table")

    def solve_transformed(self, x0, z0, xref, S=None, state_bounds=None, policy_params=None, outputs=LIGHT_OUTPUTS, stream=None):
        """bmpc_solve_transformed on CUDA float64 tensors: S (count, n, n), state_bounds (count, n_rows, 2), either None."""
        import torch
        count = x0.shape[0]
        for t in (x0, z0, xref):
            if not (t.is_cuda and t.dtype == torch.float64 and t.is_contiguous() and t.shape == (count, self.cfg.n)):
                raise ValueError("inputs must be contiguous CUDA float64 tensors of shape (count, n)")
        for t, numel in ((S, count * self.cfg.n ** 2), (state_bounds, count * self.cfg.n_rows * 2),
                         (policy_params, count * self.cfg.m * 4)):
            if t is not None and not (t.is_cuda and t.dtype == torch.float64 and t.is_contiguous() and t.numel() == numel):
                raise ValueError("S, state_bounds, policy_params must be contiguous CUDA float64 tensors of their documented shapes")
        bufs = self.device_outputs(count, outputs)
        out = abi.Outputs(**{k: bufs[k].data_ptr() for k in bufs})
        if stream is None:
            stream = torch.cuda.current_stream(x0.device).cuda_stream
        ptr = lambda t: None if t is None else t.data_ptr()
        self._check(self.lib.bmpc_solve_transformed(self.h, x0.data_ptr(), z0.data_ptr(), xref.data_ptr(), ptr(policy_params),
                                                    ptr(S), ptr(state_bounds), count, C.byref(out), C.c_void_p(stream)),
                    "bmpc_solve_transformed")
        return bufs

    def solve_transformed_host_views(self, x0, z0, xref, S=None, state_bounds=None, policy_params=None, outputs=ALL_OUTPUTS):
        """As solve_host_views for a BMPC_MODEL_MERGE handle, with the call's state transform S (count, n, n) and the (lo, hi)
        bounds of the state rows (count, n_rows, 2); either may be None."""
        x0 = np.ascontiguousarray(np.atleast_2d(x0), dtype=np.float64)
        z0 = np.ascontiguousarray(np.atleast_2d(z0), dtype=np.float64)
        xref = np.ascontiguousarray(np.atleast_2d(xref), dtype=np.float64)
        count, n = x0.shape[0], self.cfg.n
        if x0.shape != (count, n) or z0.shape != x0.shape or xref.shape != x0.shape:
            raise ValueError("x0, z0, xref must have shape (count, n)")
        pp = None
        if policy_params is not None:
            pp = np.ascontiguousarray(policy_params, dtype=np.float64).reshape(count, self.cfg.m, 4)
        if S is not None:
            S = np.ascontiguousarray(S, dtype=np.float64).reshape(count, n, n)
        if state_bounds is not None:
            state_bounds = np.ascontiguousarray(state_bounds, dtype=np.float64).reshape(count, self.cfg.n_rows, 2)
        want = abi.Outputs(**{k: 1 for k in outputs})
        views = abi.Outputs()
        self._check(self.lib.bmpc_solve_transformed_host_views(
            self.h, x0.ctypes.data, z0.ctypes.data, xref.ctypes.data, None if pp is None else pp.ctypes.data,
            None if S is None else S.ctypes.data, None if state_bounds is None else state_bounds.ctypes.data, count,
            C.byref(want), C.byref(views)), "bmpc_solve_transformed_host_views")
        res = {}
        for k in outputs:
            dt = np.int64 if k == "cycles" else (np.int32 if k in _INT_OUTPUTS else np.float64)
            shape = (count,) + _OUT_SHAPES[k](self)
            nbytes = int(np.prod(shape)) * np.dtype(dt).itemsize
            buf = (C.c_char * nbytes).from_address(getattr(views, k))
            res[k] = np.frombuffer(buf, dtype=dt).reshape(shape)
        return res

    # -- belief-state MPC -----------------------------------------------------------------------------------------
    def solve_belief_host(self, x0, b0, xbackup, xref, outputs=("u0", "uPred", "xPred", "bPred", "objective", "status", "iters")):
        """PredictiveControllers.MPC.solve(x0, b0, xbackup, xRef) for a batch: x0 (B,4), b0 (B,M,m), xbackup (B, M*m, cols),
        xref (B,4) host arrays -> dict of host arrays (bmpc_solve_belief on device tensors, then copied back)."""
        import torch
        dev = torch.device("cuda", self.cfg.device)
        x0 = np.ascontiguousarray(np.atleast_2d(x0), np.float64)
        B = x0.shape[0]
        nb = self.cfg.hmm_M * self.cfg.m
        b0 = np.ascontiguousarray(b0, np.float64).reshape(B, nb)
        xbackup = np.ascontiguousarray(xbackup, np.float64).reshape(B, nb, -1)
        xref = np.ascontiguousarray(np.broadcast_to(np.atleast_2d(xref), (B, 4)), np.float64)
        t = [torch.as_tensor(a, device=dev) for a in (x0, b0, xbackup, xref)]
        shapes = dict(_OUT_SHAPES)
        bufs = {}
        for k in outputs:
            dt = torch.int64 if k == "cycles" else (torch.int32 if k in _INT_OUTPUTS else torch.float64)
            shp = (self.totalx, nb) if k == "bPred" else shapes[k](self)
            bufs[k] = torch.zeros((B,) + shp, dtype=dt, device=dev)
        out = abi.Outputs(**{k: v.data_ptr() for k, v in bufs.items()})
        stream = torch.cuda.current_stream(dev).cuda_stream
        self._check(self.lib.bmpc_solve_belief(self.h, t[0].data_ptr(), t[1].data_ptr(), t[2].data_ptr(), xbackup.shape[-1],
                                               t[3].data_ptr(), B, C.byref(out), C.c_void_p(stream)), "bmpc_solve_belief")
        torch.cuda.synchronize(dev)
        return {k: v.cpu().numpy() for k, v in bufs.items()}

    def eval_belief(self, xb, xbackup, u):
        """regressionAndLinearization for a batch of points: returns A, B, C, h0 (K, M, m), Jh (K, M, m, n), xbp."""
        import torch
        dev = torch.device("cuda", self.cfg.device)
        M, m = self.cfg.hmm_M, self.cfg.m
        nb, n = M * m, 4 + M * m
        xb = np.ascontiguousarray(np.atleast_2d(xb), np.float64)
        K = xb.shape[0]
        t = [torch.as_tensor(np.ascontiguousarray(a, np.float64).reshape(K, -1), device=dev) for a in (xb, xbackup, u)]
        o = {"A": (K, n, n), "B": (K, n, 2), "C": (K, n), "h0": (K, nb), "Jh": (K, nb, 2), "xbp": (K, n)}
        r = {k: torch.zeros(s, dtype=torch.float64, device=dev) for k, s in o.items()}
        stream = torch.cuda.current_stream(dev).cuda_stream
        self._check(self.lib.bmpc_eval_belief(self.h, t[0].data_ptr(), t[1].data_ptr(), t[2].data_ptr(), K, r["A"].data_ptr(),
                                              r["B"].data_ptr(), r["C"].data_ptr(), r["h0"].data_ptr(), r["Jh"].data_ptr(),
                                              r["xbp"].data_ptr(), C.c_void_p(stream)), "bmpc_eval_belief")
        torch.cuda.synchronize(dev)
        res = {k: v.cpu().numpy() for k, v in r.items()}
        Jh = np.zeros((K, M, m, n))
        Jh[..., 0:2] = res["Jh"].reshape(K, M, m, 2)
        res["Jh"] = Jh
        res["h0"] = res["h0"].reshape(K, M, m)
        return res

    # -- persistent state ---------------------------------------------------------------------------------------
    def get_state(self, count=None):
        count = self.capacity if count is None else count
        st = {"uLin": np.empty((count, self.ulin_rows, self.cfg.d)), "pbest": np.empty((count, self.nbranch), np.int32),
              "old_input": np.empty((count, self.cfg.d)), "started": np.empty(count, np.int32)}
        xprev = None
        if self.cfg.controller == abi.CTRL_ROBUST:
            st["xprev"] = np.empty((count, self.totalx, self.cfg.n))
            xprev = st["xprev"].ctypes.data
        self._check(self.lib.bmpc_get_state(self.h, st["uLin"].ctypes.data, st["pbest"].ctypes.data,
                                            st["old_input"].ctypes.data, st["started"].ctypes.data, xprev, count, 1),
                    "bmpc_get_state")
        return st

    def set_state(self, st):
        count = st["uLin"].shape[0]
        arrs = [np.ascontiguousarray(st["uLin"], np.float64), np.ascontiguousarray(st["pbest"], np.int32),
                np.ascontiguousarray(st["old_input"], np.float64), np.ascontiguousarray(st["started"], np.int32)]
        xprev = np.ascontiguousarray(st["xprev"], np.float64) if "xprev" in st else None
        self._check(self.lib.bmpc_set_state(self.h, *[a.ctypes.data for a in arrs],
                                            None if xprev is None else xprev.ctypes.data, count, 1), "bmpc_set_state")

    # -- model functions (parity of rows M1-M5) -------------------------------------------------------------------
    def eval_model(self, x, z, u, policy_params=None):
        import torch
        dev = torch.device("cuda", self.cfg.device)
        n, d, m, N = self.cfg.n, self.cfg.d, self.cfg.m, self.cfg.N
        tx = torch.as_tensor(np.ascontiguousarray(x, np.float64), device=dev)
        tz = torch.as_tensor(np.ascontiguousarray(z, np.float64), device=dev)
        tu = torch.as_tensor(np.ascontiguousarray(u, np.float64), device=dev)
        K = tx.shape[0]
        pp = None
        if policy_params is not None:
            pp = torch.as_tensor(np.ascontiguousarray(policy_params, np.float64).reshape(K, m, 4), device=dev)
        o = {"A": (K, n, n), "B": (K, n, d), "C": (K, n), "xp": (K, n), "zpred": (K, N, m * n), "p": (K, m), "hlin": (K,),
             "dh": (K, n)}
        t = {k: torch.zeros(s, dtype=torch.float64, device=dev) for k, s in o.items()}
        stream = torch.cuda.current_stream(dev).cuda_stream
        self._check(self.lib.bmpc_eval_model(self.h, tx.data_ptr(), tz.data_ptr(), tu.data_ptr(),
                                             None if pp is None else pp.data_ptr(), K, t["A"].data_ptr(),
                                             t["B"].data_ptr(), t["C"].data_ptr(), t["xp"].data_ptr(),
                                             t["zpred"].data_ptr(), t["p"].data_ptr(), t["hlin"].data_ptr(),
                                             t["dh"].data_ptr(), C.c_void_p(stream)), "bmpc_eval_model")
        torch.cuda.synchronize(dev)
        return {k: v.cpu().numpy() for k, v in t.items()}

    def staging_enabled(self):
        """True if the kernel prefetches the next episode's state into shared memory with bulk copies (TMA)."""
        return bool(self.lib.bmpc_staging_enabled(self.h))

    def launch_info(self):
        mode, warps, smem, gl = C.c_int32(), C.c_int32(), C.c_int64(), C.c_int64()
        self._check(self.lib.bmpc_get_launch_info(self.h, C.byref(mode), C.byref(warps), C.byref(smem), C.byref(gl)),
                    "bmpc_get_launch_info")
        return {"slab_mode": ["auto", "shared", "split", "global"][mode.value], "warps": warps.value,
                "smem_bytes_per_warp": smem.value, "global_bytes_per_warp": gl.value}

    def last_kernel_ms(self):
        return float(self.lib.bmpc_last_kernel_ms(self.h))

    def launch_count(self):
        return int(self.lib.bmpc_launch_count(self.h))
