"""CPU restatement (float64 numpy) of the reference's closed-loop environments.  TEST INFRASTRUCTURE (see oracle/models.py).

  HighwayEnvOracle : /root/reference/Highway_env_branch.py:46-184 (Highway_env.__init__/step) with the collision check
                     of Highway_sim (:393-445).
  QuadEnvOracle    : /root/reference/quadruped_env.py:42-130 (Quad_env.__init__/step).

Pinned: tests/golden/highway_env_*.npz were recorded by running the reference's own Highway_env around the reference
BranchMPC (tests/golden/make_golden.py env); tests/test_oracle_env.py replays them.  The quadruped environment of the
reference cannot run past its first step (main_quadruped.py unpacks BT2array into the wrong number of names), so that
restatement is *unpinned* beyond the functions it shares with the pinned model.

Behaviour kept on purpose:
  * the obstacle's arg-max policy is computed from rollouts made BEFORE the lane-change target is updated in the same
    step (zpred_eval runs first, Highway_env_branch.py:100 vs :117-118);
  * those rollouts use the symbolic policy branches (brake: softmax([-7,-v],5)), the collision check on them the numeric
    veh_col (+-5 clip) and the hard min with the EGO's lane-boundary value (:145);
  * the input actually applied to the obstacle comes from the environment's ORIGINAL policy list (:60, :149): numeric
    branches (brake: softmax([-5,-v],3)) and the lane-change target of construction time;
  * the coin flips (:121-133) only write desired_x, which nothing reads - no random stream is needed.
"""
import numpy as np

from .models import softmin, softmax

LANE_W = 3.6
V0 = 20.0


def veh_col_numeric(x1, x2, size, alpha=1.0):
    """highway_branch_dyn.py:241-253: numeric branch, clipped to +-5."""
    dx = np.clip(np.abs(x1[:, 0] - x2[:, 0]) - size[0], -5.0, 5.0)
    dy = np.clip(np.abs(x1[:, 1] - x2[:, 1]) - size[1], -5.0, 5.0)
    return (dx * np.exp(alpha * dx) + dy * np.exp(alpha * dy)) / (np.exp(alpha * dx) + np.exp(alpha * dy))


def lane_bdry_numeric(x, lb, ub):
    """highway_branch_dyn.py:207-214."""
    return np.array([softmin([r[1] - lb, ub - r[1]], 5.0) for r in x])


class HighwayEnvOracle:
    def __init__(self, mpc, N_lane=4, x_ego=(0., 1.8, V0, 0.), x_obs=(5., 5.4, V0, 0.)):
        self.mpc = mpc
        self.model = mpc.model
        self.N_lane = int(N_lane)
        self.x = np.array(x_ego, dtype=float)
        self.z = np.array(x_obs, dtype=float)
        self.lane = [0, 0]                     # vehicle(..., laneidx=0), Highway_env_branch.py:29,:69
        self.backupidx = 0
        self.collision = False
        self.LB = (self.model.W / 2.0, self.N_lane * LANE_W - self.model.W / 2.0)      # :63
        self.env_policies = [tuple(p) for p in self.model.policies]                  # the list captured at :60
        self.xref = None

    # numeric branch of the policies (highway_branch_dyn.py:54-148), as the env applies them to the obstacle
    def _numeric_policy(self, i, s):
        p = self.env_policies[i]
        K = self.model.Kpsi
        if p[0] == "maintain":
            return np.array([0.0, -K * s[3]])
        if p[0] == "brake":
            return np.array([softmax([-5.0, -s[2]], 3.0), -K * s[3]])
        if p[0] == "lc":
            t = p[1]
            return np.array([-0.8558 * (s[2] - t[2]), -0.3162 * (s[1] - t[1]) - 3.9889 * (s[3] - t[3])])
        raise ValueError(p)

    def step(self, t_):
        # Highway_sim's check, before the step (:421-429): v_length 4, v_width 2.4
        dis = max(abs(self.x[0] - self.z[0]) - 4.0, abs(self.x[1] - self.z[1]) - 2.4)
        self.collision = self.collision or dis < 0
        states = [self.x, self.z]
        xx = [None, None]
        for i in range(2):
            s = states[i]
            xx[i] = self.model.zpred_eval(s)                                   # :100, before any target update
            newlane = int(round((s[1] - 1.8) / LANE_W))
            if t_ == 0 or (newlane != self.lane[i] and abs(s[1] - 1.8 - LANE_W * newlane) < 1.4):   # :104
                self.lane[i] = newlane
                if i == 1:
                    le, lo = self.lane
                    if le < lo:
                        tgt = lo - 1
                    elif le > lo:
                        tgt = lo + 1
                    else:
                        tgt = lo - 1 if lo > 0 else lo + 1
                    self._set_lc_target(np.array([0.0, 1.8 + LANE_W * tgt, V0, 0.0]))             # :109-118
        n = 4
        x1 = xx[0][:, 0:n]                                                     # ego under policy 0 (backupidx never changes)
        lane_h = lane_bdry_numeric(x1, self.LB[0], self.LB[1])
        hi = np.zeros(self.model.m)
        for j in range(self.model.m):
            col = veh_col_numeric(x1, xx[1][:, j * n:(j + 1) * n], [self.model.L + 1.0, self.model.W + 0.2])
            hi[j] = min(col.min(), lane_h.min())                               # :145
        self.backupidx = int(np.argmax(hi))
        u_obs = self._numeric_policy(self.backupidx, self.z)                   # :149
        # xRef rule (:153-167)
        if self.x[0] < self.z[0]:
            Ydes = 1.8 + self.lane[0] * LANE_W
        else:
            Ydes = self.z[1]
        if abs(self.x[1] - Ydes) < 1 and self.x[0] > self.z[0] + 3:
            vdes = V0
        else:
            vdes = self.z[2] + 1.0 * (self.z[0] + 1.5 - self.x[0])
        self.xref = np.array([0.0, Ydes, vdes, 0.0])
        u = np.array(self.mpc.solve(self.x, self.z, self.xref), dtype=float)
        self.x = self.x + self.model.dt * np.array([self.x[2] * np.cos(self.x[3]), self.x[2] * np.sin(self.x[3]), u[0], u[1]])
        self.z = self.z + self.model.dt * np.array([self.z[2] * np.cos(self.z[3]), self.z[2] * np.sin(self.z[3]),
                                                   u_obs[0], u_obs[1]])
        return u, u_obs

    def _set_lc_target(self, target):
        for k, p in enumerate(self.model.policies):
            if p[0] == "lc":
                self.model.policies[k] = ("lc", np.asarray(target, dtype=float))

    def lc_target(self):
        for p in self.model.policies:
            if p[0] == "lc":
                return np.asarray(p[1], dtype=float)
        return np.zeros(4)


class QuadEnvOracle:
    """quadruped_env.py:42-130.  Robot sizes: ego (L1, W1), obstacle (L2, W2) from Quad_constants (main_quadruped.py:31)."""

    def __init__(self, mpc, x_des, x_ego=(0., 1.8, 0.), x_obs=(2.5, 2.5, -np.pi / 2), L1=0.5, L2=1.0, col_tol=0.2, v0=0.2):
        self.mpc = mpc
        self.model = mpc.model
        self.x = np.array(x_ego, dtype=float)
        self.z = np.array(x_obs, dtype=float)
        self.x_des = np.array(x_des, dtype=float)
        self.L1, self.L2, self.col_tol, self.v0 = L1, L2, col_tol, v0
        self.backupidx = 0
        self.xref = None

    def step(self, t_):
        xx_e = self.model.zpred_eval(self.x)
        xx_o = self.model.zpred_eval(self.z)
        n = 3
        x1 = xx_e[:, 0:n]
        hi = np.zeros(self.model.m)
        for j in range(self.model.m):
            x2 = xx_o[:, j * n:(j + 1) * n]
            # robot_col numeric branch (quadruped_branch_dyn.py:146-150): Euclidean distance minus (L1+L2)/2 and the tolerance
            hi[j] = (np.sqrt((x1[:, 0] - x2[:, 0]) ** 2 + (x1[:, 1] - x2[:, 1]) ** 2) - (self.L1 + self.L2) / 2 - self.col_tol).min()
        self.backupidx = 0 if hi[0] > 0.5 else int(np.argmax(hi))              # :91-94
        u_obs = np.array([self.v0, 0.0, 0.0]) if self.backupidx == 0 else np.zeros(3)
        dx = self.x_des[0:2] - self.x[0:2]
        nd = np.linalg.norm(dx)
        dx = dx / nd * min(nd, 5.0)
        if np.linalg.norm(dx) > 0.1:
            psi = np.arctan2(dx[1], dx[0])
            while psi - self.x_des[2] > np.pi:
                psi -= 2 * np.pi
            while psi - self.x_des[2] < -np.pi:
                psi += 2 * np.pi
        else:
            psi = self.x[2]
        self.xref = np.array([self.x[0] + dx[0], self.x[1] + dx[1], psi])      # :100-114
        u = np.array(self.mpc.solve(self.x, self.z, self.xref), dtype=float)
        dt = self.model.dt

        def plant(s, v):
            return s + dt * np.array([v[0] * np.cos(s[2]) - v[1] * np.sin(s[2]), v[1] * np.cos(s[2]) + v[0] * np.sin(s[2]), v[2]])

        self.x = plant(self.x, u)
        self.z = plant(self.z, u_obs)
        return u, u_obs
