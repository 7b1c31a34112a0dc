"""Drop-in for the reference's `quadruped_branch_dyn.py` (n=3: x, y, theta; d=3: vx, vy, r).

Reference: quadruped_branch_dyn.py:14-27 quad_kinetics, :34-54 policies, :135-150 robot_col, :154-248 PredictiveModel.
"""
import numpy as np

from _bmpc import abi, batch, config, policies
from _bmpc.policies import PolicyProbe, PolicyDescriptor
from numpy import arctan2   # quadruped_env.py:104 gets casadi's arctan2 through `from quadruped_branch_dyn import *`
from utils import Quad_constants

# the reference module has no __all__: `from quadruped_branch_dyn import *` (main_quadruped.py:4) also hands out the names it
# imported itself - main_quadruped.py:31 relies on Quad_constants arriving that way
__all__ = ["quad_kinetics", "softsat", "backup_forward", "backup_stop", "softmin", "softmax", "propagate_backup",
           "robot_col", "PredictiveModel", "Quad_constants", "np", "arctan2"]


def quad_kinetics(x, u):
    c, s = np.cos(x[2]), np.sin(x[2])
    return np.array([u[0] * c - u[1] * s, u[0] * s + u[1] * c, u[2]])


def softsat(x, s):
    return (np.exp(s * x) - 1) / (np.exp(s * x) + 1) * 0.5 + 0.5


def softmin(x, gamma=1):
    x = np.asarray(x, dtype=float)
    return np.sum(np.exp(-gamma * x) * x) / np.sum(np.exp(-gamma * x))


def softmax(x, gamma=1):
    x = np.asarray(x, dtype=float)
    return np.sum(np.exp(gamma * x) * x) / np.sum(np.exp(gamma * x))


def backup_forward(x, v0):
    if isinstance(x, PolicyProbe):
        return PolicyDescriptor(abi.POLICY_FORWARD, [v0])
    return np.array([v0, 0, 0])


def backup_stop(x):
    if isinstance(x, PolicyProbe):
        return PolicyDescriptor(abi.POLICY_STOP)
    return np.array([0, 0, 0])


def propagate_backup(x, dyn, N, ts):
    x = np.asarray(x, dtype=float)
    xs = np.empty([N, x.shape[0]])
    for i in range(N):
        x = x + dyn(x) * ts
        xs[i, :] = x
    return xs


def robot_col(x1, x2, L1, W1, L2, W2, tol, alpha=1):
    """Numeric branch: Euclidean distance minus margin (reference :146-150; the model's symbolic branch is L1-norm)."""
    x1 = np.atleast_2d(np.asarray(x1, dtype=float))
    x2 = np.atleast_2d(np.asarray(x2, dtype=float))
    return np.array([np.linalg.norm(a[0:2] - b[0:2]) - (L1 + L2) / 2 - tol for a, b in zip(x1, x2)])


class PredictiveModel:
    """quadruped_branch_dyn.PredictiveModel(n, d, N, backupcons, dt, cons) (reference :154-248)."""

    def __init__(self, n, d, N, backupcons, dt, cons):
        if (n, d) != (3, 3):
            raise ValueError("the quadruped model has n=3, d=3")
        self.n, self.d, self.N, self.dt, self.cons = n, d, N, dt, cons
        self._eval = None
        self.update_backup(backupcons)

    def update_backup(self, backupcons):
        self.backupcons = backupcons
        self.m = len(backupcons)
        self.descriptors = policies.describe(backupcons)
        if self._eval is not None:
            self._eval.close()
            self._eval = None

    def spec(self):
        c = self.cons
        return config.quadruped_spec(self.N, self.dt, policies.table(self.descriptors), c.L1, c.L2, c.col_tol, c.s1)

    def policy_params(self):
        return policies.param_array(self.descriptors)

    def _handle(self):
        if self._eval is None:
            cfg = config.make_config(self.spec(), 3, 3, self.N, 1, np.eye(3), np.eye(3), np.empty((0, 3)), np.empty(0),
                                     np.kron(np.eye(3), np.array([1., -1.])).T, np.ones(6), np.array([0., 1.]))
            self._eval = batch.BatchedBranchMPC(cfg)
        return self._eval

    def _eval_points(self, x, z=None, u=None):
        x = np.atleast_2d(np.asarray(x, dtype=float))
        z = x if z is None else np.atleast_2d(np.asarray(z, dtype=float))
        u = np.zeros((x.shape[0], 3)) if u is None else np.atleast_2d(np.asarray(u, dtype=float))
        pp = np.broadcast_to(self.policy_params(), (x.shape[0], self.m, 4))
        return self._handle().eval_model(x, z, u, pp)

    def dyn_linearization(self, x, u):
        r = self._eval_points(x, u=u)
        return r["A"][0], r["B"][0], r["C"][0], r["xp"][0]

    def branch_eval(self, x, z):
        x = np.asarray(x, dtype=float).reshape(-1)
        z = np.asarray(z, dtype=float).reshape(-1)
        h = 1e-6
        pts = [x] + [x + s * h * np.eye(3)[k] for k in range(3) for s in (1, -1)]
        p = self._eval_points(np.array(pts), z=np.tile(z, (len(pts), 1)))["p"]
        dp = np.column_stack([(p[1 + 2 * k] - p[2 + 2 * k]) / (2 * h) for k in range(3)])
        return p[0], dp

    def zpred_eval(self, z):
        return self._eval_points(z, z=z)["zpred"][0]

    def xpred_eval(self, x):
        x = np.asarray(x, dtype=float).reshape(-1)
        return self._eval_points(x, z=x)["zpred"][0][:, :3], self.backupcons[0](x)

    def col_eval(self, x, z):
        r = self._eval_points(x, z=z)
        return r["hlin"][0], r["dh"][0]
