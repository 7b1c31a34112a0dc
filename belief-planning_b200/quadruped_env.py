"""Drop-in for the reference's `quadruped_env.py` (Quad_env, robot, Robot_sim), stepped on the B200 through
`bmpc_env_step` (csrc/bmpc_env.cuh).  `Quad_env(NR, mpc, x_des)` keeps the reference's constructor (quadruped_env.py:43-65,
NR = 2); extension: `x0` of shape (B, 2, 3) and `x_des` of shape (B, 3) run B independent episodes."""
import numpy as np

from _bmpc import env as _env


class robot:
    def __init__(self, state=(0, 0, 0), L=1, W=0.5, dt=0.05, backupidx=0):
        self.state = np.array(state, dtype=float)
        self.dt, self.L, self.W, self.backupidx = dt, L, W, backupidx


class Quad_env:
    def __init__(self, NR, mpc, x_des, x0=None):
        if NR != 2:
            raise ValueError("the reference environment is defined for NR = 2 (quadruped_env.py:56)")
        self.NR, self.mpc = NR, mpc
        self.predictiveModel = mpc.predictiveModel
        self.dt = self.predictiveModel.dt
        self.backupcons = self.predictiveModel.backupcons
        self.m = len(self.backupcons)
        self.cons = self.predictiveModel.cons
        x0 = np.array([[0, 1.8, 0], [2.5, 2.5, -np.pi / 2]], dtype=float) if x0 is None else np.asarray(x0, dtype=float)
        self._single = x0.ndim == 2
        X = x0[None] if self._single else x0
        self._B = X.shape[0]
        solver = mpc._ensure_solver(self._B)
        c = self.cons
        self._dev = _env.BatchedQuadEnv(solver, X[:, 0], X[:, 1], np.asarray(x_des, dtype=float), c.L1, c.L2, c.col_tol)
        self.robot_set = [robot(X[0, 0], L=c.L1, W=c.W1, dt=self.dt), robot(X[0, 1], L=c.L2, W=c.W2, dt=self.dt)]
        self.desired_x = [np.asarray(x_des, dtype=float), X[0, 1]]

    def step(self, t_):
        """quadruped_env.py:67-130: returns u_set, x_set, xx_set (None: plotting only), xPred, zPred."""
        self._dev.t = int(t_)
        out = self._dev.step(outputs=("u0", "uPred", "xPred", "xLin", "zPred", "branch_w", "branch_p", "objective",
                                      "status", "iters"))
        self.mpc._absorb({k: v.cpu().numpy() for k, v in out.items()}, self._single)
        h = self._dev.host()
        pick = (lambda a: a[0]) if self._single else (lambda a: a)
        self.robot_set[0].state, self.robot_set[1].state = pick(h["x"]), pick(h["z"])
        self.robot_set[1].backupidx = pick(h["obs_policy"])
        self.xRef = pick(h["xref"])
        u_set = [pick(out["u0"].cpu().numpy()), pick(h["u_obs"])]
        x_set = [r.state for r in self.robot_set]
        xPred, zPred, uPred, _ = self.mpc.BT2array()
        return u_set, x_set, [None] * self.NR, xPred, zPred


def Robot_sim(env, T):
    """quadruped_env.py:133-170 (collision check on the robots' bounding circles as there)."""
    N = int(round(T / env.dt))
    state_rec = np.zeros((env.NR, N, 3)) if env._single else np.zeros((env._B, env.NR, N, 3))
    input_rec = np.zeros(state_rec.shape)
    xPred_rec, zPred_rec = [None] * N, [None] * N
    for t in range(N):
        u_set, x_set, _, xPred, zPred = env.step(t)
        xPred_rec[t], zPred_rec[t] = xPred, zPred
        for i in range(env.NR):
            if env._single:
                state_rec[i][t], input_rec[i][t] = x_set[i], u_set[i]
            else:
                state_rec[:, i, t], input_rec[:, i, t] = x_set[i], u_set[i]
    return state_rec, input_rec, xPred_rec, zPred_rec
