"""Batched closed-loop environments on the device: thousands of episodes of Highway_env_branch.Highway_env /
quadruped_env.Quad_env advanced by one `bmpc_env_step` per control period (obstacle policy, lane bookkeeping, xRef rule,
controller solve, both plants), with no host round trip between consecutive MPC steps.

PyTorch owns the state tensors; the arithmetic is in libbranchmpc.so (csrc/bmpc_env.cuh + the solve kernel).
"""
import ctypes as C

import numpy as np

from . import abi, batch


class _BatchedEnv:
    def __init__(self, mpc, x, z):
        import torch
        if not isinstance(mpc, batch.BatchedBranchMPC):
            raise TypeError("mpc must be a BatchedBranchMPC")
        self.mpc = mpc
        self.dev = torch.device("cuda", mpc.cfg.device)
        n, d = mpc.cfg.n, mpc.cfg.d
        x = np.ascontiguousarray(np.atleast_2d(x), dtype=np.float64)
        z = np.ascontiguousarray(np.atleast_2d(z), dtype=np.float64)
        if x.shape != z.shape or x.shape[1] != n:
            raise ValueError("x and z must both be (count, n)")
        self.count = x.shape[0]
        if self.count > mpc.capacity:
            raise ValueError("more episodes than the controller's batch_capacity")
        B = self.count
        self.x = torch.as_tensor(x, device=self.dev)
        self.z = torch.as_tensor(z, device=self.dev)
        self.obs_policy = torch.zeros(B, dtype=torch.int32, device=self.dev)
        self.collided = torch.zeros(B, dtype=torch.int32, device=self.dev)
        self.xref = torch.zeros((B, n), dtype=torch.float64, device=self.dev)
        self.u_obs = torch.zeros((B, d), dtype=torch.float64, device=self.dev)
        self.t = 0
        self.last = None

    def _state(self):
        raise NotImplementedError

    def _extra(self):
        return 0, None

    def step(self, outputs=batch.LIGHT_OUTPUTS, stream=None):
        """One control period of every episode; returns the controller's output tensors (device, reused between calls).
        Asynchronous on `stream` (default: torch's current stream)."""
        import torch
        mpc = self.mpc
        if "u0" not in outputs:
            outputs = ("u0",) + tuple(outputs)
        bufs = mpc.device_outputs(self.count, outputs)
        if "branch_p" in bufs:
            bufs["branch_p"].fill_(float("nan"))
        out = abi.Outputs(**{k: bufs[k].data_ptr() for k in bufs})
        if stream is None:
            stream = torch.cuda.current_stream(self.dev).cuda_stream
        st = self._state()
        n_lane, sizes = self._extra()
        mpc._check(mpc.lib.bmpc_env_step(mpc.h, C.byref(st), self.count, self.t, n_lane, sizes, C.byref(out),
                                         C.c_void_p(stream)), "bmpc_env_step")
        self.t += 1
        self.last = bufs
        return bufs

    def host(self):
        """Snapshot of the environment state as numpy arrays (synchronises)."""
        return {k: getattr(self, k).cpu().numpy() for k in self._fields}


class BatchedHighwayEnv(_BatchedEnv):
    """Highway_env_branch.py:46-184, one obstacle per episode.  x0/z0: (count, 4); lc_target: the lane-change target of
    the controller's model at construction (main_branch.py:35), rewritten per episode when the obstacle changes lane."""
    _fields = ("x", "z", "lane", "policy_params", "obs_policy", "collided", "xref", "u_obs")

    def __init__(self, mpc, x0, z0, N_lane=4, policy_params=None):
        import torch
        super().__init__(mpc, x0, z0)
        cfg = mpc.cfg
        self.N_lane = int(N_lane)
        self.lane = torch.zeros((self.count, 2), dtype=torch.int32, device=self.dev)       # vehicle(laneidx=0)
        if policy_params is None:
            pp = np.tile(np.array([[cfg.policy_param[j][k] for k in range(4)] for j in range(cfg.m)]), (self.count, 1, 1))
        else:
            pp = np.ascontiguousarray(policy_params, dtype=np.float64).reshape(self.count, cfg.m, 4)
        self.policy_params = torch.as_tensor(pp, device=self.dev)

    def _state(self):
        return abi.EnvState(x=self.x.data_ptr(), z=self.z.data_ptr(), lane=self.lane.data_ptr(),
                            policy_params=self.policy_params.data_ptr(), goal=None, obs_policy=self.obs_policy.data_ptr(),
                            collided=self.collided.data_ptr(), xref=self.xref.data_ptr(), u_obs=self.u_obs.data_ptr())

    def _extra(self):
        return self.N_lane, None


class BatchedQuadEnv(_BatchedEnv):
    """quadruped_env.py:42-130.  goal: (count, 3) desired final state x_des; sizes = (L1, L2, col_tol) of Quad_constants."""
    _fields = ("x", "z", "goal", "obs_policy", "collided", "xref", "u_obs")

    def __init__(self, mpc, x0, z0, goal, L1=0.5, L2=1.0, col_tol=0.2):
        import torch
        super().__init__(mpc, x0, z0)
        g = np.array(np.broadcast_to(np.asarray(goal, dtype=np.float64), (self.count, 3)))
        self.goal = torch.as_tensor(g, device=self.dev)
        self._sizes = (C.c_double * 3)(L1, L2, col_tol)

    def _state(self):
        return abi.EnvState(x=self.x.data_ptr(), z=self.z.data_ptr(), lane=None, policy_params=None,
                            goal=self.goal.data_ptr(), obs_policy=self.obs_policy.data_ptr(),
                            collided=self.collided.data_ptr(), xref=self.xref.data_ptr(), u_obs=self.u_obs.data_ptr())

    def _extra(self):
        return 0, C.cast(self._sizes, C.c_void_p)


class BatchedMergeEnv:
    """Highway_env_branch.Highway_env_merge (:271-380) for a batch of episodes, stepped on the device by
    `bmpc_env_step_merge`: x0, z0 (count, 4); the ramp tables are the arrays `merge_geometry` returns."""
    _fields = ("x", "z", "lane_id", "collided", "xref", "S", "state_bounds", "u_obs")

    def __init__(self, mpc, x0, z0, table_x, table_y, table_psi, N_lane=2, merge_lane=1, merge_s=50.0, v0=20.0, lane_id=1):
        import torch
        if not isinstance(mpc, batch.BatchedBranchMPC) or mpc.cfg.model != abi.MODEL_MERGE:
            raise TypeError("mpc must be a BatchedBranchMPC of the merge model")
        self.mpc = mpc
        self.dev = torch.device("cuda", mpc.cfg.device)
        x0 = np.ascontiguousarray(np.atleast_2d(x0), dtype=np.float64)
        z0 = np.ascontiguousarray(np.atleast_2d(z0), dtype=np.float64)
        B = self.count = x0.shape[0]
        if x0.shape != (B, 4) or z0.shape != (B, 4) or B > mpc.capacity:
            raise ValueError("x0, z0 must be (count, 4) with count <= batch_capacity")
        t64 = lambda a: torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64), device=self.dev)
        self.x, self.z = t64(x0), t64(z0)
        self.lane_id = torch.full((B,), int(lane_id), dtype=torch.int32, device=self.dev)
        self.collided = torch.zeros(B, dtype=torch.int32, device=self.dev)
        self.xref = torch.zeros((B, 4), dtype=torch.float64, device=self.dev)
        self.S = torch.zeros((B, 4, 4), dtype=torch.float64, device=self.dev)
        self.state_bounds = torch.zeros((B, 2, 2), dtype=torch.float64, device=self.dev)
        self.u_obs = torch.zeros((B, 2), dtype=torch.float64, device=self.dev)
        self.tab = [t64(np.asarray(a).reshape(-1)) for a in (table_x, table_y, table_psi)]
        if not (len(self.tab[0]) == len(self.tab[1]) == len(self.tab[2]) >= 2):
            raise ValueError("the three ramp tables must have the same length")
        self.N_lane, self.merge_lane, self.merge_s, self.v0 = int(N_lane), int(merge_lane), float(merge_s), float(v0)
        self.t = 0
        self.last = None

    def step(self, outputs=batch.LIGHT_OUTPUTS, stream=None):
        import torch
        mpc = self.mpc
        if "u0" not in outputs:
            outputs = ("u0",) + tuple(outputs)
        bufs = mpc.device_outputs(self.count, outputs)
        out = abi.Outputs(**{k: bufs[k].data_ptr() for k in bufs})
        if stream is None:
            stream = torch.cuda.current_stream(self.dev).cuda_stream
        st = abi.MergeEnvState(x=self.x.data_ptr(), z=self.z.data_ptr(), lane_id=self.lane_id.data_ptr(),
                               collided=self.collided.data_ptr(), xref=self.xref.data_ptr(), S=self.S.data_ptr(),
                               state_bounds=self.state_bounds.data_ptr(), u_obs=self.u_obs.data_ptr(),
                               table_x=self.tab[0].data_ptr(), table_y=self.tab[1].data_ptr(), table_psi=self.tab[2].data_ptr(),
                               table_n=len(self.tab[0]))
        mpc._check(mpc.lib.bmpc_env_step_merge(mpc.h, C.byref(st), self.count, self.N_lane, self.merge_lane, self.merge_s,
                                               self.v0, C.byref(out), C.c_void_p(stream)), "bmpc_env_step_merge")
        self.t += 1
        self.last = bufs
        return bufs

    def host(self):
        return {k: getattr(self, k).cpu().numpy() for k in self._fields}
