"""Drop-in for the reference's `Highway_env_branch.py` environment (Highway_env, vehicle, Highway_sim, sim_overtake),
stepped on the B200: one `bmpc_env_step` per control period does the obstacle's arg-max policy, the lane bookkeeping, the
xRef rule, the Branch-MPC solve and both plants for every episode of the batch (csrc/bmpc_env.cuh).

`Highway_env(NV, mpc, N_lane)` keeps the reference's constructor (Highway_env_branch.py:47-81; NV must be 2 - the
reference's initial-state table has two rows).  Extension: `x0` of shape (B, 2, 4) runs B independent episodes; the
per-vehicle attributes then carry a leading batch axis.  Plotting/animation helpers are not part of the path.
"""
import numpy as np

from _bmpc import env as _env

v0 = 20
lane_width = 3.6
lm = np.arange(0, 7) * lane_width


class vehicle:
    """View of one vehicle of the batch (Highway_env_branch.py:28-41)."""

    def __init__(self, state=(0, 0, v0, 0), v_length=4, v_width=2.4, dt=0.05, backupidx=0, laneidx=0):
        self.state = np.array(state, dtype=float)
        self.dt, self.v_length, self.v_width = dt, v_length, v_width
        self.backupidx, self.laneidx = backupidx, laneidx


class Highway_env:
    def __init__(self, NV, mpc, N_lane=6, x0=None):
        if NV != 2:
            raise ValueError("the reference environment is defined for NV = 2 (Highway_env_branch.py:67)")
        self.NV, self.N_lane, self.mpc = NV, N_lane, mpc
        self.predictiveModel = mpc.predictiveModel
        self.dt = self.predictiveModel.dt
        self.backupcons = self.predictiveModel.backupcons
        self.m = len(self.backupcons)
        self.cons = self.predictiveModel.cons
        self.LB = [self.cons.W / 2, N_lane * 3.6 - self.cons.W / 2]
        x0 = np.array([[0, 1.8, v0, 0], [5, 5.4, v0, 0]], dtype=float) if x0 is None else np.asarray(x0, dtype=float)
        self._single = x0.ndim == 2
        X = x0[None] if self._single else x0
        self._B = X.shape[0]
        solver = mpc._ensure_solver(self._B)
        pp = np.broadcast_to(self.predictiveModel.policy_params(), (self._B, self.m, 4))
        self._dev = _env.BatchedHighwayEnv(solver, X[:, 0], X[:, 1], N_lane, pp)
        self.veh_set = [vehicle(X[0, i], dt=self.dt) for i in range(NV)]
        self.desired_x = [np.array([0, X[0, i, 1], v0, 0]) for i in range(NV)]
        self.collision = np.zeros(self._B, dtype=bool)

    def step(self, t_):
        """One control period (Highway_env_branch.py:83-184).  Returns u_set, x_set, xx_set, xPred, zPred, branch_w as
        the reference does for a single episode; for a batch the arrays carry a leading batch axis and the tree lists
        are those of episode 0.  xx_set (the backup rollouts, only used for plotting) is not recomputed: None."""
        self._dev.t = int(t_)
        out = self._dev.step(outputs=("u0", "uPred", "xPred", "xLin", "zPred", "branch_w", "branch_p", "objective",
                                      "status", "iters"))
        self.mpc._absorb({k: v.cpu().numpy() for k, v in out.items()}, self._single)
        h = self._dev.host()
        pick = (lambda a: a[0]) if self._single else (lambda a: a)
        for i, key in enumerate(("x", "z")):
            self.veh_set[i].state = pick(h[key])
            self.veh_set[i].laneidx = pick(h["lane"][:, i])
        self.veh_set[1].backupidx = pick(h["obs_policy"])
        self.collision = h["collided"].astype(bool)
        self.xRef = pick(h["xref"])
        u_set = [pick(out["u0"].cpu().numpy()), pick(h["u_obs"])]
        x_set = [v.state for v in self.veh_set]
        xPred, zPred, uPred, branch_w = self.mpc.BT2array()
        return u_set, x_set, [None] * self.NV, xPred, zPred, branch_w


def Highway_sim(env, T):
    """Highway_env_branch.py:393-445 without the plotting records of the backup rollouts."""
    N = int(round(T / env.dt))
    B = env._B
    shape = (env.NV, N, 4) if env._single else (B, env.NV, N, 4)
    state_rec = np.zeros(shape)
    input_rec = np.zeros(shape[:-1] + (2,))
    backup_choice_rec = [[None] * N for _ in range(env.NV)]
    xPred_rec, zPred_rec, branch_w_rec = [None] * N, [None] * N, [None] * N
    for t in range(N):
        u_set, x_set, _, xPred, zPred, branch_w = env.step(t)
        xPred_rec[t], zPred_rec[t], branch_w_rec[t] = xPred, zPred, branch_w
        for i in range(env.NV):
            if env._single:
                input_rec[i][t], state_rec[i][t] = u_set[i], x_set[i]
            else:
                input_rec[:, i, t], state_rec[:, i, t] = u_set[i], x_set[i]
            backup_choice_rec[i][t] = env.veh_set[i].backupidx
    collision = bool(env.collision[0]) if env._single else env.collision.copy()
    return state_rec, input_rec, [None] * env.NV, backup_choice_rec, xPred_rec, zPred_rec, branch_w_rec, collision


def sim_overtake(mpc, N_lane, T=10):
    """Highway_env_branch.py:719-725 (the animation is not part of the path); returns what Highway_sim returns."""
    env = Highway_env(NV=2, mpc=mpc, N_lane=N_lane)
    return Highway_sim(env, T)
