"""Scenario constants of the reference entry scripts, restated for the oracle.  TEST INFRASTRUCTURE.

  highway   : /root/reference/main_branch.py:24-48, Init_MPC.py:40-72
  quadruped : /root/reference/main_quadruped.py:14-41, Init_MPC.py:74-94
"""
import numpy as np

from .models import HighwayModel, QuadrupedModel
from .branch_mpc import BranchMPCOracle


def highway_model(policies=None, N=8, dt=0.1, lc_target=(0.5, 1.8, 15.0, 0.0)):
    if policies is None:
        policies = ["maintain", "brake", "lc"]
    desc = []
    for p in policies:
        if p == "lc":
            desc.append(("lc", np.asarray(lc_target, dtype=float)))
        elif isinstance(p, str):
            desc.append((p,))
        else:
            desc.append(tuple(p))
    return HighwayModel(N, desc, dt, L=4.0, W=2.5, Kpsi=0.1, s1=2.0, N_lane=3)


def highway_mpc_params(am=6.0, rm=0.3, N_lane=4, W=2.5):
    Fx = np.array([[0., 1., 0., 0.], [0., -1., 0., 0.], [0., 0., 0., 1.], [0., 0., 0., -1.]])
    bx = np.array([N_lane * 3.6 - W / 2, -W / 2, 0.25, 0.25])
    Fu = np.array([[1., 0.], [-1., 0.], [0., 1.], [0., -1.]])
    bu = np.array([am, am, rm, rm])
    return dict(Q=np.diag([0., 3., 3., 10.]), R=np.diag([1., 100.]), Fx=Fx, bx=bx, Fu=Fu, bu=bu,
                Qslack=np.array([0., 300.]))


def highway_branch_mpc(policies=None, NB=2, N=8, lc_target=(0.5, 1.8, 15.0, 0.0), xRef=None, N_lane=4):
    model = highway_model(policies, N=N, lc_target=lc_target)
    par = highway_mpc_params(N_lane=N_lane)
    xRef = np.asarray(lc_target if xRef is None else xRef, dtype=float)
    return BranchMPCOracle(model, NB, xRef=xRef, variant="branch", **par)


def quadruped_model(N=25, dt=0.2, v0=0.2):
    return QuadrupedModel(N, [("forward", v0), ("stop",)], dt, L1=0.5, L2=1.0, col_tol=0.2, s1=2.0)


def quadruped_mpc_params(vxm=0.2, vym=0.1, rm=0.5):
    Fu = np.kron(np.eye(3), np.array([[1.], [-1.]]))
    bu = np.array([vxm, 0., vym, vym, rm, rm])
    return dict(Q=np.eye(3), R=np.diag([1., 100., 1.]), dR=np.array([0.9, 5., 1.]), Fx=np.empty((0, 3)),
                bx=np.empty(0), Fu=Fu, bu=bu, Qslack=np.array([0., 300.]))


def quadruped_prox_mpc(NB=2, N=25, xRef=(5., 5., 0.)):
    model = quadruped_model(N=N)
    return BranchMPCOracle(model, NB, xRef=np.asarray(xRef, dtype=float), variant="prox", **quadruped_mpc_params())


def highway_robust_mpc(policies=None, NB=2, N=8, lc_target=(0.5, 1.8, 15.0, 0.0), xRef=None, N_lane=4):
    from .robust_mpc import RobustMPCOracle
    model = highway_model(policies, N=N, lc_target=lc_target)
    par = highway_mpc_params(N_lane=N_lane)
    xRef = np.asarray(lc_target if xRef is None else xRef, dtype=float)
    return RobustMPCOracle(model, NB, xRef=xRef, **par)
