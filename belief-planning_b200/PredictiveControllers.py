"""Drop-in for the reference's `PredictiveControllers.py` (belief-state MPC).

`MPC(mpcParameters, predictiveModel)` and `solve(x0, b0, xbackup, xRef=None)` keep the reference's interface
(PredictiveControllers.py:56-160): after a solve the object carries `xPred (N+1, n)` with n = 4 + M*m (physical state, then the
predicted beliefs), `uPred (N, d)`, `xLin`, `uLin`, `feasible`, `solverTime`, `OldInput`, `timeStep`.  The same call accepts a
batch: x0 (B, 4), b0 (B, M, m), xbackup (B, M*m, cols), and the result attributes gain a leading batch axis.

Underneath (per solve): get_xLin, computeLTVdynamics, buildIneqConstr, buildCost, buildEqConstr, the dense->CSC conversions, OSQP
setup + solve + polish and unpackSolution (:115-340) are one launch of the chain instance of the persistent solve kernel
(bmpc_solve_belief, csrc/bmpc_solver.h expand_belief / belief_outputs).  There is no CPU path.
"""
import datetime
from dataclasses import dataclass, field

import numpy as np

from _bmpc import abi, batch, config
from utils import PythonMsg

__all__ = ["PythonMsg", "MPCParams", "MPC"]


def _f():
    return field(default=None)


@dataclass
class MPCParams(PythonMsg):
    """PredictiveControllers.py:27-54."""
    n: int = _f()
    d: int = _f()
    N: int = _f()
    A: np.ndarray = _f()
    B: np.ndarray = _f()
    Q: np.ndarray = _f()
    R: np.ndarray = _f()
    Qf: np.ndarray = _f()
    dR: np.ndarray = _f()
    Qslack: float = _f()
    Fx: np.ndarray = _f()
    bx: np.ndarray = _f()
    Fu: np.ndarray = _f()
    bu: np.ndarray = _f()
    xRef: np.ndarray = _f()
    slacks: bool = field(default=True)
    timeVarying: bool = field(default=False)

    def __post_init__(self):
        if self.Qf is None:
            self.Qf = np.zeros((self.n, self.n))
        if self.dR is None:
            self.dR = np.zeros(self.d)
        if self.xRef is None:
            self.xRef = np.zeros(self.n)


class MPC:
    """PredictiveControllers.MPC (:56-340)."""

    def __init__(self, mpcParameters, predictiveModel, **solver_knobs):
        p = mpcParameters
        self.N, self.Qslack, self.Q, self.Qf, self.R, self.dR = p.N, p.Qslack, p.Q, p.Qf, p.R, p.dR
        self.n, self.d, self.A, self.B = p.n, p.d, p.A, p.B
        self.Fx, self.Fu, self.bx, self.bu, self.xRef = p.Fx, p.Fu, p.bx, p.bu, p.xRef
        self.M, self.m = predictiveModel.M, predictiveModel.m
        self.nx = self.n - self.M * self.m
        self.thres = 0.1                                                     # :76
        self.alphad = np.exp(-predictiveModel.alpha * predictiveModel.dt)
        self.slacks, self.timeVarying = p.slacks, p.timeVarying
        self.predictiveModel = predictiveModel
        if not p.slacks or not p.timeVarying:
            raise NotImplementedError("the reference builds this controller with slacks=True, timeVarying=True (Init_MPC.py:33)")
        if self.nx != 4 or self.d != 2:
            raise ValueError("the belief-state model is the highway model: nx = 4, d = 2")
        self._knobs = solver_knobs
        self._solver = None
        self._capacity = 0
        self.OldInput = np.zeros((1, 2))
        self.xPred = self.uPred = self.xLin = self.uLin = None
        self.feasible = 0
        self.solverTime = datetime.timedelta(0)
        self.linearizationTime = datetime.timedelta(0)
        self.timeStep = 0
        abi.load_library()          # fail now, loudly, if the CUDA library is missing

    def _make_solver(self, capacity):
        model = self.predictiveModel
        n4 = self.nx
        Fx = np.asarray(self.Fx, dtype=float)
        if np.abs(Fx[:, n4:]).max(initial=0.0) > 0:
            raise NotImplementedError("state constraints on the belief part are not built (initMPCParams has none, Init_MPC.py:13)")
        Q = np.asarray(self.Q, dtype=float)
        Qf = np.asarray(self.Qf, dtype=float)
        if np.abs(Q[n4:, :]).max(initial=0.0) > 0 or np.abs(Q[:, n4:]).max(initial=0.0) > 0 or np.abs(Qf[n4:, n4:]).max(initial=0.0) > 0:
            raise NotImplementedError("costs on the belief part are not built (initMPCParams has none, Init_MPC.py:27)")
        bx = np.squeeze(np.asarray(self.bx, dtype=float)).reshape(-1)          # the reference stores a 1-tuple (Init_MPC.py:15-18)
        bu = np.squeeze(np.asarray(self.bu, dtype=float)).reshape(-1)
        cfg = config.make_config(model.spec(self.N), 4, 2, self.N, 1, Q[:n4, :n4], self.R, Fx[:, :n4], bx, self.Fu, bu, self.Qslack,
                                 controller=abi.CTRL_BELIEF, Qf=Qf[:n4, :n4], dR=self.dR, batch_capacity=capacity,
                                 hmm_M=self.M, hmm_col_alpha=float(model.cons.col_alpha), hmm_tran_diag=float(model.cons.tran_diag),
                                 hmm_thres=float(self.thres), **self._knobs)
        return batch.BatchedBranchMPC(cfg)

    def solve(self, x0, b0, xbackup, xRef=None):
        """Computes the control action(s) (:130-160).  x0: (4,) or (B, 4); b0: (M, m) or (B, M, m); xbackup: (M*m, cols) or
        (B, M*m, cols), row m*i+j = agent i under policy j, column block 4k..4k+3 = backup state at step k."""
        if xRef is not None:
            self.xRef = np.append(np.asarray(xRef, dtype=float)[..., :4], np.zeros(np.shape(xRef)[:-1] + (self.M * self.m,)), axis=-1)
        x0 = np.asarray(x0, dtype=float)
        single = x0.ndim == 1
        X = np.atleast_2d(x0)
        B = X.shape[0]
        b0 = np.asarray(b0, dtype=float).reshape(B, self.M, self.m)
        xb = np.asarray(xbackup, dtype=float).reshape(B, self.M * self.m, -1)
        R = np.broadcast_to(np.atleast_2d(np.asarray(self.xRef, dtype=float))[:, :4], (B, 4))
        if self._solver is None or B > self._capacity:
            if self._solver is not None:
                self._solver.close()
            self._solver = self._make_solver(B)
            self._capacity = B
        t0 = datetime.datetime.now()
        r = self._solver.solve_belief_host(X, b0, xb, R)
        self.solverTime = datetime.datetime.now() - t0
        ok = r["status"] <= abi.STATUS_CONVERGED
        self.status = r["status"][0] if single else r["status"]
        self.feasible = int(ok[0]) if single else ok.astype(int)
        xPred = np.concatenate([r["xPred"], r["bPred"]], axis=2)
        pick = (lambda a: a[0]) if single else (lambda a: a)
        self.xPred, self.uPred = pick(xPred), pick(r["uPred"])
        self.objective = pick(r["objective"])
        self.zt, self.zt_u = self.xPred[..., -1, :], self.uPred[..., -1, :]              # feasibleStateInput :179-181
        # timeVarying (:156-158): the next linearisation is the shifted plan
        self.xLin = np.concatenate([self.xPred[..., 1:, :], self.xPred[..., -1:, :]], axis=-2)
        self.uLin = np.concatenate([self.uPred[..., 1:, :], self.uPred[..., -1:, :]], axis=-2)
        self.OldInput = self.uPred[..., 0, :]
        self.timeStep += 1

    def reset(self, episode_ids=None):
        if self._solver is not None:
            self._solver.reset(episode_ids)
