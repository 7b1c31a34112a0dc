"""Drop-in for the reference's `Highway_env_branch.py` environment (Highway_env, vehicle, Highway_sim, sim_overtake),
stepped on the B200: one `bmpc_env_step` per control period does the obstacle's arg-max policy, the lane bookkeeping, the
xRef rule, the Branch-MPC solve and both plants for every episode of the batch (csrc/bmpc_env.cuh).

The merge scenario (`merge_geometry`, `Highway_env_merge`, `sim_merge`, reference :227-380, :727-733) is stepped by the host
caller around the device controller: per control period one model evaluation per vehicle and one
`BranchMPC_CVaR.solve(x, z, xRef, S, Fx=None, bx=bx)`.

`Highway_env(NV, mpc, N_lane)` keeps the reference's constructor (Highway_env_branch.py:47-81; NV must be 2 - the
reference's initial-state table has two rows).  Extension: `x0` of shape (B, 2, 4) runs B independent episodes; the
per-vehicle attributes then carry a leading batch axis.  Plotting/animation helpers are not part of the path.
"""
import numpy as np

from _bmpc import env as _env
from highway_branch_dyn import interpolant, lane_bdry_h, veh_col

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


def merge_geometry(N_lane, merge_lane, merge_s, merge_R, merge_side=0):
    """Centre line of the ramp (Highway_env_branch.py:227-262): a straight piece sampled every 0.5 m up to the merge point
    `merge_s`, then an arc of radius `merge_R` that ends tangent to the highway.  Returns X1, X2, Y1, Y2, psi1, psi2
    (straight part, arc part); merge_side 0 joins from above (larger y), 1 from below."""
    theta = np.arccos(1 - lane_width * merge_lane / merge_R)            # heading of the straight piece
    s1 = np.linspace(0, merge_s, num=int(merge_s / 0.5), endpoint=False)
    s2 = merge_s + np.linspace(0, merge_R * theta, num=int(merge_R * theta / 0.5))
    cx = merge_s + merge_R * np.sin(theta)                              # arc centre
    x_start = merge_s - merge_s * np.cos(theta)
    X1 = x_start + s1 * np.cos(theta)
    if merge_side == 0:
        cy = (N_lane - merge_lane) * lane_width + merge_R
        Y1 = N_lane * lane_width + np.sin(theta) * merge_s - s1 * np.sin(theta)
        psi1 = -theta * np.ones(s1.shape)
        psi2 = (s2 - s2[-1]) / merge_R
        X2 = cx + np.sin(psi2) * merge_R
        Y2 = cy - np.cos(psi2) * merge_R
    else:
        cy = merge_lane * lane_width - merge_R
        Y1 = -np.sin(theta) * merge_s - lane_width * merge_lane + s1 * np.sin(theta)
        psi1 = theta * np.ones(s1.shape)
        psi2 = (s2[-1] - s2) / merge_R
        X2 = cx - np.sin(psi2) * merge_R
        Y2 = cy + np.cos(psi2) * merge_R - merge_lane * lane_width
    return X1, X2, Y1, Y2, psi1, psi2


class Highway_env_merge:
    """Highway_env_merge(NV, N_lane, mpc, pred_model, merge_lane, merge_s, merge_R, merge_side, dt) (reference :271-380): the
    ego (vehicle 0) starts on the ramp (lane id 1) and merges in front of / behind vehicle 1 on the highway.  While on the
    ramp the controller works in ramp coordinates: state transform S, reference and lane bounds built from the ramp tables at
    the ego's x (:357-363)."""

    def __init__(self, NV, N_lane, mpc, pred_model, merge_lane=2, merge_s=50, merge_R=300, merge_side=0, dt=0.05):
        if NV != 2:
            raise ValueError("the reference environment is defined for NV = 2 (Highway_env_branch.py:317)")
        self.NV, self.N_lane, self.mpc, self.pred_model, self.dt = NV, N_lane, mpc, pred_model, dt
        self.merge_lane, self.merge_s, self.merge_R, self.merge_side = merge_lane, merge_s, merge_R, merge_side
        self.laneID = [1] + [0] * (NV - 1)
        self.backupcons = [pm.backupcons for pm in pred_model]
        self.m = [len(b) for b in self.backupcons]
        self.cons = mpc.predictiveModel.cons
        self.LB = [self.cons.W / 2, N_lane * 3.6 - self.cons.W / 2]
        X1, X2, Y1, Y2, psi1, psi2 = merge_geometry(N_lane, merge_lane, merge_s, merge_R, merge_side)
        self.merge_theta = np.arccos(1 - lane_width * merge_lane / merge_R)
        self.merge_end = merge_s + merge_R * np.sin(self.merge_theta)
        self.merge_lane_ref_X1, self.merge_lane_ref_X2 = X1, X2
        self.merge_lane_ref_Y1, self.merge_lane_ref_Y2 = Y1, Y2
        self.merge_lane_ref_psi1, self.merge_lane_ref_psi2 = psi1, psi2
        self.merge_lane_ref_X = np.append(X1, X2)
        self.merge_lane_ref_Y = np.append(Y1, Y2)
        self.merge_lane_ref_psi = np.append(psi1, psi2)
        self.refY = interpolant("refY", "linear", [self.merge_lane_ref_X], self.merge_lane_ref_Y)
        self.refpsi = interpolant("refY", "linear", [self.merge_lane_ref_X], self.merge_lane_ref_psi)
        x0 = np.array([[24, 13, v0, -0.2], [15, 5.4, v0, 0]], dtype=float)
        self.veh_set = [vehicle(x0[i], dt=self.dt, backupidx=0) for i in range(NV)]
        self.desired_x = [np.array([0, x0[i, 1], v0, 0]) for i in range(NV)]
        self._B, self._single = 1, True
        self.collision = np.zeros(1, dtype=bool)

    def step(self, t_):
        """One control period (reference :324-380).  Returns u_set, x_set, xx_set, xPred, zPred, branch_w."""
        cons, n = self.cons, 4
        a, b = self.veh_set
        gap = max(abs(a.state[0] - b.state[0]) - 0.5 * (a.v_length + b.v_length),
                  abs(a.state[1] - b.state[1]) - 0.5 * (a.v_width + b.v_width))          # Highway_sim :421-429
        self.collision = self.collision | (gap < 0)
        xx_set = []
        for i, veh in enumerate(self.veh_set):
            if veh.state[0] > self.merge_s + 8:
                self.laneID[i] = 0                      # past the merge point: on the highway for good
            xx_set.append(self.pred_model[self.laneID[i]].zpred_eval(veh.state))
        idx0 = self.veh_set[0].backupidx
        x1 = xx_set[0][:, idx0 * n:(idx0 + 1) * n]
        size = [cons.L + 1, cons.W + 0.2]
        u0_set = []
        for i, veh in enumerate(self.veh_set):
            if i != 0:
                # the safest policy is evaluated as in the reference (:337-346) and then overridden: every vehicle follows policy 0
                rolls = [xx_set[i][:, j * n:(j + 1) * n] for j in range(self.m[self.laneID[i]])]
                if self.laneID[i] == 0:
                    hi = [min(np.append(veh_col(x1, r, size), lane_bdry_h(r, self.LB[0], self.LB[1]))) for r in rolls]
                else:
                    hi = [np.min(veh_col(x1, r, size)) for r in rolls]
                veh.backupidx = int(np.argmax(hi))
            veh.backupidx = 0
            u0_set.append(self.backupcons[self.laneID[i]][veh.backupidx](veh.state))
        x = self.veh_set[0].state
        if self.laneID[0] == 0:
            S = np.eye(4)
            xRef = np.array([0, (self.N_lane - 0.5) * 3.6, v0, 0])
            bx = self.mpc.param.bx
        else:
            y0, psi0 = self.refY(x[0]), self.refpsi(x[0])
            t = np.tan(psi0)
            S = np.array([[1., 0, 0, 0], [-t, 1., 0, 0], [0, 0, 1, 0], [0, 0, 0, 1]])
            xRef = np.array([0, -t * x[0] + y0 + 1.8, v0, psi0])
            bx = np.array([-t * x[0] + y0 + 3.6 * self.merge_lane - cons.W / 2, t * x[0] - y0 - cons.W / 2,
                           psi0 + self.mpc.psimax, -psi0 + self.mpc.psimax])
        self.mpc.solve(self.veh_set[0].state, self.veh_set[1].state, xRef, S, Fx=None, bx=bx)
        u_set = [np.array(self.mpc.uPred[0])] + u0_set[1:]
        xPred, zPred, uPred, branch_w = self.mpc.BT2array()
        for veh, u in zip(self.veh_set, u_set):
            s = veh.state
            veh.state = s + self.dt * np.array([s[2] * np.cos(s[3]), s[2] * np.sin(s[3]), u[0], u[1]])
        return u_set, [veh.state for veh in self.veh_set], xx_set, xPred, zPred, branch_w


def sim_merge(mpc, pred_model, N_lane, merge_lane, merge_s, merge_R, merge_side, T=6):
    """Highway_env_branch.py:727-733 (the animation is not part of the path); returns what Highway_sim returns."""
    env = Highway_env_merge(2, N_lane, mpc, pred_model, merge_lane, merge_s, merge_R, merge_side, pred_model[0].dt)
    return Highway_sim(env, T)
