"""Drop-in for the reference's `Init_MPC.py` parameter factories (Init_MPC.py:40-94): same names, arguments and
constants, including the 1-tuple `bx` (trailing comma at :48-51, :77) that the controllers unwrap with np.squeeze."""
import numpy as np

from scipy import linalg

from MPC_branch import BranchMPCParams
from PredictiveControllers import MPC, MPCParams  # noqa: F401  (the reference module imports both, Init_MPC.py:4)


def initMPCParams(nx, d, N, M, m, ydes, vdes, am, rm, N_lane, W):
    """Belief-state MPC parameters (reference :7-34)."""
    Fx = np.array([[0., 1., 0., 0.], [0., -1., 0., 0.], [0., 0., 0., 1.], [0., 0., 0., -1.]])
    Fx = np.hstack((Fx, np.zeros([Fx.shape[0], m * M])))
    bx = np.array([[N_lane * 3.6 - W / 2], [-W / 2], [0.25], [0.25]]),      # 1-tuple, as in the reference
    Fu = np.kron(np.eye(2), np.array([1, -1])).T
    bu = np.array([[am], [0.5 * am], [rm], [rm]])
    Qx = np.diag([0., 0.5, 0.2, 5.])
    Q = linalg.block_diag(Qx, np.zeros([M * m, M * m]))
    R = np.diag([30, 100.0])
    xRef = np.append(np.array([0, ydes, vdes, 0]), np.zeros(M * m))
    Qslack = 1 * np.array([0, 1000])
    return MPCParams(n=nx + M * m, d=d, N=N, Q=Q, R=R, Fx=Fx, bx=bx, Fu=Fu, bu=bu, xRef=xRef, slacks=True, Qslack=Qslack,
                     timeVarying=True)


def initBranchMPC(n, d, N, NB, xRef, am, rm, N_lane, W):
    """Highway Branch-MPC parameters (reference :40-72)."""
    Fx = np.array([[0., 1., 0., 0.], [0., -1., 0., 0.], [0., 0., 0., 1.], [0., 0., 0., -1.]])
    bx = np.array([[N_lane * 3.6 - W / 2], [-W / 2], [0.25], [0.25]]),      # 1-tuple, as in the reference
    Fu = np.kron(np.eye(2), np.array([1, -1])).T
    bu = np.array([[am], [am], [rm], [rm]])
    Q = np.diag([0., 3, 3, 10.])
    R = np.diag([1, 100.0])
    Qslack = 1 * np.array([0, 300])
    return BranchMPCParams(n=n, d=d, N=N, NB=NB, Q=Q, R=R, Fx=Fx, bx=bx, Fu=Fu, bu=bu, xRef=xRef, slacks=True,
                           Qslack=Qslack, timeVarying=True)


def initquadBranchMPC(n, d, N, NB, xRef, vxm, vym, rm):
    """Quadruped Branch-MPC parameters (reference :74-94)."""
    Fx = np.empty([0, n])
    bx = np.empty([0, 1]),
    Fu = np.kron(np.eye(3), np.array([1, -1])).T
    bu = np.array([[vxm], [0], [vym], [vym], [rm], [rm]])
    Q = np.diag([1., 1., 1])
    R = np.diag([1., 100., 1.])
    dR = np.array([0.9, 5, 1])
    Qslack = 1 * np.array([0, 300])
    return BranchMPCParams(n=n, d=d, N=N, NB=NB, Q=Q, R=R, dR=dR, Fx=Fx, bx=bx, Fu=Fu, bu=bu, xRef=xRef, slacks=True,
                           Qslack=Qslack, timeVarying=True)
