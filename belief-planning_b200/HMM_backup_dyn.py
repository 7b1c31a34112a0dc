"""Drop-in for the reference's `HMM_backup_dyn.py` (belief-state model): same function and class names.

The reference module does not import as shipped (`from utils import HMM_constants`, :5, names a class its utils.py lacks); the
drop-in `utils` provides that dataclass.  Numeric helper functions (the branches the environments call with numpy arrays) are
plain numpy; `PredictiveModel` evaluates on the GPU: backup rollouts through bmpc_hmm_backup_rollout, the linearisation of the
augmented dynamics through bmpc_eval_belief, and it hands the kernels its constants instead of CasADi graphs.

Reference: dubin :30-41, softsat :94, backup_trans :96-101, backup_input_prob :103, backup_maintain :105, backup_brake :107-109,
softmin/softmax :111-115, propagate_backup :122-132, lane_bdry_h :134, veh_col :136-157, PredictiveModel :177-276.
"""
import numpy as np

from _bmpc import abi, batch, config, hmm as _hmm
from utils import HMM_constants, MPCParams  # noqa: F401  (the reference module imports both)

__all__ = ["dubin", "softsat", "backup_trans", "backup_input_prob", "backup_maintain", "backup_brake", "softmin", "softmax",
           "propagate_backup", "lane_bdry_h", "veh_col", "generate_backup_traj", "dubin_f_x", "PredictiveModel", "HMM_constants"]


def dubin(x, u):
    return np.array([x[2] * np.cos(x[3]), x[2] * np.sin(x[3]), u[0], u[1]])


def softsat(x, s):
    return (np.exp(s * x) - 1) / (np.exp(s * x) + 1) * 0.5 + 0.5


def softmin(x, y, gamma=1):
    return (np.exp(-gamma * x) * x + np.exp(-gamma * y) * y) / (np.exp(-gamma * x) + np.exp(-gamma * y))


def softmax(x, y, gamma=1):
    return (np.exp(gamma * x) * x + np.exp(gamma * y) * y) / (np.exp(gamma * x) + np.exp(gamma * y))


def backup_trans(h, cons):
    m = softsat(np.asarray(h, dtype=float), cons.s1)
    return np.kron((1 - cons.tran_diag) * np.ones([m.shape[0], 1]), m.reshape(1, -1) / np.sum(m)) + cons.tran_diag * np.eye(m.shape[0])


def backup_input_prob(cbfcond, cons):
    return softsat(cbfcond - cons.c2, cons.s2)


def backup_maintain(x, cons):
    return np.array([0, -cons.Kpsi * x[3]])


def backup_brake(x, cons):
    return np.array([softmax(-5, -x[2], 3), -cons.Kpsi * x[3]])


def propagate_backup(x, dyn, N, ts):
    x = np.asarray(x, dtype=float)
    xs = np.empty([N, x.shape[0]])
    for i in range(N):
        x = x + dyn(x) * ts
        xs[i, :] = x
    return xs


def lane_bdry_h(x, lb=0, ub=7.2):
    return softmin(x[1] - lb, ub - x[1], 5)


def veh_col(x1, x2, size, alpha=1):
    """numeric branch: normalised box distance with its +-5 clip (:145-157)."""
    x1, x2 = np.asarray(x1, dtype=float), np.asarray(x2, dtype=float)
    if x1.ndim == 1:
        dx = np.clip((abs(x1[0] - x2[0]) - size[0]) / size[0], -5, 5)
        dy = np.clip((abs(x1[1] - x2[1]) - size[1]) / size[1], -5, 5)
        return (dx * np.exp(alpha * dx) + dy * np.exp(dy * alpha)) / (np.exp(alpha * dx) + np.exp(dy * alpha))
    return np.array([veh_col(a, b, size, alpha) for a, b in zip(x1, x2)])


def dubin_f_x(x, con):
    h = 1e-6
    dudx = np.array([(con(x + h * np.eye(4)[k]) - con(x - h * np.eye(4)[k])) / 2 / h for k in range(4)])
    return np.concatenate((np.array([[0, 0, np.cos(x[3]), -x[2] * np.sin(x[3])], [0, 0, np.sin(x[3]), x[2] * np.cos(x[3])]]),
                           dudx.transpose()))


def _kind(con, cons):
    """policy closure -> device policy kind, by probing it at a point where the two policies differ."""
    probe = np.array([0.0, 0.0, 10.0, 0.3])
    u = np.asarray(con(probe), dtype=float)
    if np.allclose(u, backup_maintain(probe, cons)):
        return abi.HMM_MAINTAIN
    if np.allclose(u, backup_brake(probe, cons)):
        return abi.HMM_BRAKE
    raise NotImplementedError("backup policies of the belief-state model: backup_maintain, backup_brake")


def generate_backup_traj(x, con, stop_crit, f0, ts=0.05, sensitivity=True):
    """module-level rollout with the sensitivity matrix (:54-85); host loop, as in the reference (the batched device version is
    _bmpc.hmm.rollout_sensitivity)."""
    t, tt, xx, uu, QQ, Qt = 0, [], [], [], [], []
    x = np.asarray(x, dtype=float)
    Q = np.identity(4)
    while not stop_crit(x, t):
        u = con(x)
        xdot = np.array([x[2] * np.cos(x[3]), x[2] * np.sin(x[3]), u[0], u[1]])
        if sensitivity:
            QQ.append(Q)
            Qt.append(xdot - f0)
            Q = Q + np.matmul(dubin_f_x(x, con), Q) * ts
        tt.append(t)
        xx.append(x)
        uu.append(u)
        x = x + xdot * ts
        t = t + ts
    return tt, xx, uu, QQ, Qt


class PredictiveModel:
    """HMM_backup_dyn.PredictiveModel(n, d, M, backupcons, dt, cons) (:177-276); n is the physical state dimension (4)."""

    def __init__(self, n, d, M, backupcons, dt, cons):
        if (n, d) != (4, 2):
            raise ValueError("the belief-state model is the highway model: n = 4, d = 2")
        self.n, self.d, self.M, self.dt, self.cons = n, d, M, dt, cons
        self.m = len(backupcons)
        self.lamb = 0.0
        self.alpha = cons.alpha
        self.backupcons = backupcons
        self.kinds = [_kind(c, cons) for c in backupcons]
        self._eval = None

    def spec(self, N):
        pol = [((abi.POLICY_MAINTAIN if k == abi.HMM_MAINTAIN else abi.POLICY_BRAKE), [0, 0, 0, 0]) for k in self.kinds]
        c = self.cons
        return config.ModelSpec(abi.MODEL_HIGHWAY, 4, 2, N, self.dt, pol, veh_L=c.L, veh_W=c.W, Kpsi=c.Kpsi, s1=c.s1,
                                lane_lo=c.ylb, lane_hi=c.yub)

    def generate_backup_traj(self, x0, N):
        """x0 (M, 4) -> (M*m, N*4), row m*i+j, flattened the way casadi.reshape does (:204-214)."""
        return _hmm.backup_rollout(np.asarray(x0, dtype=float)[None], self.kinds, N, self.dt, self.cons.Kpsi)[0]

    def _handle(self):
        if self._eval is None:
            c = self.cons
            cfg = config.make_config(self.spec(2), 4, 2, 2, 1, np.eye(4), np.eye(2), np.array([[0., 1, 0, 0], [0, -1., 0, 0], [0, 0, 0, 1.], [0, 0, 0, -1.]]),
                                     np.ones(4), np.kron(np.eye(2), np.array([1., -1.])).T, np.ones(4), np.array([0., 1.]),
                                     controller=abi.CTRL_BELIEF, Qf=np.zeros((4, 4)), hmm_M=self.M, hmm_col_alpha=float(c.col_alpha),
                                     hmm_tran_diag=float(c.tran_diag), hmm_thres=0.1)
            self._eval = batch.BatchedBranchMPC(cfg)
        return self._eval

    def regressionAndLinearization(self, xb, xbackup, u):
        """A, B, C, h0, Jh of the augmented dynamics at (xb, u) with the backup states xbackup (M*m, 4) (:216-229)."""
        r = self._handle().eval_belief(np.asarray(xb, dtype=float), np.asarray(xbackup, dtype=float)[None], np.asarray(u, dtype=float))
        h0 = [r["h0"][0, i].reshape(-1, 1) for i in range(self.M)]
        Jh = [r["Jh"][0, i] for i in range(self.M)]
        return r["A"][0], r["B"][0], r["C"][0], h0, Jh
