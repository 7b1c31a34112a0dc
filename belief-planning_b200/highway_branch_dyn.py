"""Drop-in for the reference's `highway_branch_dyn.py`: same function and class names, same call signatures.

The reference builds CasADi graphs and evaluates them one tree node at a time; here `PredictiveModel` is a thin host
object whose evaluation methods run on the GPU through libbranchmpc (bmpc_eval_model), and which hands the kernels a
policy TABLE (kind + parameters) instead of closures.  The numeric branches of the helper functions (the ones the
simulation environments call with numpy arrays) are plain numpy.

Reference: highway_branch_dyn.py:17-34 dubin, :38 softsat, :54-148 policies, :151-162 softmin/softmax, :174-187
propagate_backup, :195-206 lane_bdry_h, :223-254 veh_col, :262-398 PredictiveModel, :400-502 PredictiveModel_merge.
`interpolant` is the one CasADi name the reference's scripts reach through this module's star import (main_branch.py:78).
"""
import numpy as np

from _bmpc import abi, batch, config, policies
from _bmpc.policies import PolicyProbe, PolicyDescriptor

__all__ = ["dubin", "softsat", "backup_maintain", "backup_maintain_trackV", "backup_brake", "backup_lc", "softmin",
           "softmax", "propagate_backup", "lane_bdry_h", "veh_col", "PredictiveModel", "PredictiveModel_merge", "interpolant"]


class interpolant:
    """casadi.interpolant(name, 'linear', [grid], values) for one dimension, the form main_branch.py:78-79 and
    Highway_env_branch.py:310-311 build (refY, refpsi): piecewise linear, end segments continued outside the grid.  Called
    with a number it returns a float; handed to a ramp policy it becomes the lookup table of the device model."""

    def __init__(self, name, kind, grid, values, *args, **kwargs):
        if kind != "linear" or len(grid) != 1:
            raise NotImplementedError("only one-dimensional linear lookup tables are built")
        self.name = name
        self.xs = np.ascontiguousarray(grid[0], dtype=float).reshape(-1)
        self.ys = np.ascontiguousarray(values, dtype=float).reshape(-1)
        if self.xs.shape != self.ys.shape or self.xs.size < 2 or np.any(np.diff(self.xs) <= 0):
            raise ValueError("a lookup table needs a strictly increasing grid and one value per grid point")

    def __call__(self, x):
        x = float(np.asarray(x, dtype=float).reshape(-1)[0])
        k = min(max(int(np.searchsorted(self.xs, x, side="right")) - 1, 0), self.xs.size - 2)
        slope = (self.ys[k + 1] - self.ys[k]) / (self.xs[k + 1] - self.xs[k])
        return float(self.ys[k] + slope * (x - self.xs[k]))


def dubin(x, u):
    """xdot of the Dubins car (reference :17-34): state (x, y, v, psi), input (a, r)."""
    return np.array([x[2] * np.cos(x[3]), x[2] * np.sin(x[3]), u[0], u[1]])


def softsat(x, s):
    return (np.exp(s * x) - 1) / (np.exp(s * x) + 1) * 0.5 + 0.5


def softmin(x, gamma=1):
    x = np.asarray(x, dtype=float)
    return np.sum(np.exp(-gamma * x) * x) / np.sum(np.exp(-gamma * x))


def softmax(x, gamma=1):
    x = np.asarray(x, dtype=float)
    return np.sum(np.exp(gamma * x) * x) / np.sum(np.exp(gamma * x))


def _no_psiref(psiref):
    if psiref is not None:
        raise NotImplementedError("backup_maintain with a heading table is not built (no reference script uses it)")


def _table(psiref):
    if not isinstance(psiref, interpolant):
        raise TypeError("psiref must be this module's `interpolant` (a one-dimensional linear lookup table)")
    return psiref


def backup_maintain(x, cons, psiref=None):
    """Keep speed, steer heading to zero (reference :54-67)."""
    _no_psiref(psiref)
    if isinstance(x, PolicyProbe):
        return PolicyDescriptor(abi.POLICY_MAINTAIN, consts={"Kpsi": cons.Kpsi})
    return np.array([0., -cons.Kpsi * x[3]])


def backup_maintain_trackV(x, cons, v0, psiref=None):
    """Track speed v0 (reference :80-96); with `psiref` the steering follows the ramp's heading table (:89-96)."""
    if isinstance(x, PolicyProbe):
        if psiref is not None:
            return PolicyDescriptor(abi.POLICY_TRACKV_REF, [v0], consts={"Kpsi": cons.Kpsi}, table=_table(psiref))
        return PolicyDescriptor(abi.POLICY_TRACKV, [v0], consts={"Kpsi": cons.Kpsi})
    return np.array([0.5 * (v0 - x[2]), (psiref(x[0]) if psiref is not None else 0.0) - cons.Kpsi * x[3]])


def backup_brake(x, cons, psiref=None):
    """Brake (reference :108-121).  The model uses the symbolic branch softmax([-7, -v], 5) (in the kernels); the
    numeric branch, which only the environment calls, is softmax([-5, -v], 3) (:121).  With `psiref` both branches are
    softmax([-5, -v], 3) and the steering follows the ramp's heading table (:122-131)."""
    if isinstance(x, PolicyProbe):
        if psiref is not None:
            return PolicyDescriptor(abi.POLICY_BRAKE_REF, consts={"Kpsi": cons.Kpsi}, table=_table(psiref))
        return PolicyDescriptor(abi.POLICY_BRAKE, consts={"Kpsi": cons.Kpsi})
    return np.array([softmax(np.array([-5., -x[2]]), 3), (psiref(x[0]) if psiref is not None else 0.0) - cons.Kpsi * x[3]])


def backup_lc(x, x0):
    """Lane change towards the state x0 (reference :136-148)."""
    if isinstance(x, PolicyProbe):
        return PolicyDescriptor(abi.POLICY_LC, list(np.asarray(x0, dtype=float).reshape(-1)[:4]))
    return np.array([-0.8558 * (x[2] - x0[2]), -0.3162 * (x[1] - x0[1]) - 3.9889 * (x[3] - x0[3])])


def propagate_backup(x, dyn, N, ts):
    """Forward-Euler rollout, row t = state after t+1 steps (reference :174-187); numeric branch."""
    x = np.asarray(x, dtype=float)
    xs = np.empty([N, x.shape[0]])
    for i in range(N):
        x = x + dyn(x) * ts
        xs[i, :] = x
    return xs


def lane_bdry_h(x, lb=0, ub=7.2):
    """softmin([y - lb, ub - y], 5) per row (reference :195-206)."""
    x = np.asarray(x, dtype=float)
    if x.ndim == 1:
        return softmin(np.array([x[1] - lb, ub - x[1]]), 5)
    return np.array([softmin(np.array([r[1] - lb, ub - r[1]]), 5) for r in x])


def veh_col(x1, x2, size, alpha=1):
    """Smooth box distance, numeric branch with its +-5 clip (reference :245-252)."""
    x1 = np.asarray(x1, dtype=float)
    x2 = np.asarray(x2, dtype=float)
    if x1.ndim == 1:
        dx = np.clip(abs(x1[0] - x2[0]) - size[0], -5, 5)
        dy = np.clip(abs(x1[1] - x2[1]) - size[1], -5, 5)
        return (dx * np.exp(alpha * dx) + dy * np.exp(dy * alpha)) / (np.exp(alpha * dx) + np.exp(dy * alpha))
    return np.array([veh_col(a, b, size, alpha) for a, b in zip(x1, x2)])


class PredictiveModel:
    """highway_branch_dyn.PredictiveModel(n, d, N, backupcons, dt, cons, N_lane=3) (reference :262-398)."""

    def __init__(self, n, d, N, backupcons, dt, cons, N_lane=3):
        if (n, d) != (4, 2):
            raise ValueError("the highway model has n=4, d=2")
        self.n, self.d, self.N, self.dt, self.cons = n, d, N, dt, cons
        self.N_lane = N_lane
        self.LB = [cons.W / 2, N_lane * 3.6 - cons.W / 2]
        self._eval = None
        self.update_backup(backupcons)

    # -- policy table -------------------------------------------------------------------------------------
    def update_backup(self, backupcons):
        """New policy list (the reference rebuilds every CasADi function, :331-334; here it is a parameter change)."""
        self.backupcons = backupcons
        self.m = len(backupcons)
        self.descriptors = policies.describe(backupcons)
        for dsc in self.descriptors:
            if "Kpsi" in dsc.consts and dsc.consts["Kpsi"] != self.cons.Kpsi:
                raise ValueError("policies must share the model's Kpsi")
        if self._eval is not None and [dd.kind for dd in self.descriptors] != self._eval_kinds:
            self._eval.close()
            self._eval = None

    def spec(self):
        return config.highway_spec(self.N, self.dt, policies.table(self.descriptors), self.cons.L, self.cons.W,
                                   self.cons.Kpsi, self.cons.s1, self.N_lane)

    def policy_params(self):
        return policies.param_array(self.descriptors)

    # -- point evaluation on the device --------------------------------------------------------------------
    def _handle(self):
        if self._eval is None:
            cfg = config.make_config(self.spec(), 4, 2, self.N, 1, np.eye(4), np.eye(2), np.empty((0, 4)), np.empty(0),
                                     np.kron(np.eye(2), np.array([1., -1.])).T, np.ones(4), np.array([0., 1.]))
            self._eval = batch.BatchedBranchMPC(cfg)
            self._eval_kinds = [dd.kind for dd in self.descriptors]
        return self._eval

    def _eval_points(self, x, z=None, u=None):
        x = np.atleast_2d(np.asarray(x, dtype=float))
        z = x if z is None else np.atleast_2d(np.asarray(z, dtype=float))
        u = np.zeros((x.shape[0], 2)) if u is None else np.atleast_2d(np.asarray(u, dtype=float))
        pp = np.broadcast_to(self.policy_params(), (x.shape[0], self.m, 4))
        return self._handle().eval_model(x, z, u, pp)

    def dyn_linearization(self, x, u):
        r = self._eval_points(x, u=u)
        return r["A"][0], r["B"][0], r["C"][0], r["xp"][0]

    def branch_eval(self, x, z):
        """(p, dp): dp never enters any QP (BranchTree.J is never written, MPC_branch.py:76,:1085-1089); it is returned
        as a central difference of the device-evaluated p."""
        x = np.asarray(x, dtype=float).reshape(-1)
        z = np.asarray(z, dtype=float).reshape(-1)
        h = 1e-6
        pts = [x] + [x + s * h * np.eye(4)[k] for k in range(4) for s in (1, -1)]
        p = self._eval_points(np.array(pts), z=np.tile(z, (len(pts), 1)))["p"]
        dp = np.column_stack([(p[1 + 2 * k] - p[2 + 2 * k]) / (2 * h) for k in range(4)])
        return p[0], dp

    def zpred_eval(self, z):
        return self._eval_points(z, z=z)["zpred"][0]

    def xpred_eval(self, x):
        x = np.asarray(x, dtype=float).reshape(-1)
        return self._eval_points(x, z=x)["zpred"][0][:, :4], backup_maintain(x, self.cons)

    def col_eval(self, x, z):
        r = self._eval_points(x, z=z)
        return r["hlin"][0], r["dh"][0]


class PredictiveModel_merge(PredictiveModel):
    """highway_branch_dyn.PredictiveModel_merge(n, d, N, backupcons, dt, cons, merge_ref, laneID, N_lane1, N_lane2) (reference
    :400-502): the highway vehicle with the merge scenario's branching probability (vehicle distance only, size [L+1, W+0.2],
    :452-456) and, for the ramp lane, policies that steer along a heading lookup table.  On the device this is
    BMPC_MODEL_MERGE; a controller built on it takes the state transform `S` and the bounds `bx` of every solve."""

    def __init__(self, n, d, N, backupcons, dt, cons, merge_ref, laneID=0, N_lane1=3, N_lane2=2):
        self.refY, self.refpsi = merge_ref
        self.laneID, self.N_lane2 = laneID, N_lane2
        self.LB1 = [cons.W / 2, N_lane1 * 3.6 - cons.W / 2]
        super().__init__(n, d, N, backupcons, dt, cons, N_lane=N_lane1)

    def update_backup(self, backupcons):
        super().update_backup(backupcons)
        tables = [dd.table for dd in self.descriptors if dd.table is not None]
        if any(t is not tables[0] for t in tables):
            raise NotImplementedError("the policies of one model share one heading table")
        self.lookup_table = tables[0] if tables else None
        if self._eval is not None and self.lookup_table is not None:
            self._eval.set_lookup_table(self.lookup_table.xs, self.lookup_table.ys)

    def spec(self):
        return config.merge_spec(self.N, self.dt, policies.table(self.descriptors), self.cons.L, self.cons.W, self.cons.Kpsi,
                                 self.cons.s1)

    def _handle(self):
        if self._eval is None:
            Fx = np.array([[0., 1., 0., 0.], [0., -1., 0., 0.], [0., 0., 0., 1.], [0., 0., 0., -1.]])
            cfg = config.make_config(self.spec(), 4, 2, self.N, 1, np.eye(4), np.eye(2), Fx, np.ones(4),
                                     np.kron(np.eye(2), np.array([1., -1.])).T, np.ones(4), np.array([0., 1.]),
                                     controller=abi.CTRL_CVAR, cvar_alpha=0.5)
            self._eval = batch.BatchedBranchMPC(cfg)
            self._eval_kinds = [dd.kind for dd in self.descriptors]
            if self.lookup_table is not None:
                self._eval.set_lookup_table(self.lookup_table.xs, self.lookup_table.ys)
        return self._eval

    def xpred_eval(self, x):
        """(ego rollout under the first policy, that policy's input at x) (reference :498-499)."""
        x = np.asarray(x, dtype=float).reshape(-1)
        return self._eval_points(x, z=x)["zpred"][0][:, :4], self.backupcons[0](x)
