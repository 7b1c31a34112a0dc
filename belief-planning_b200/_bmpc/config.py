"""Translation of the reference's parameter objects (BranchMPCParams + PredictiveModel constants) into bmpc_config.

Reference: MPC_branch.py:27-54 (BranchMPCParams), Init_MPC.py:40-94, utils.py:25-59.
"""
import numpy as np

from . import abi


class ModelSpec:
    """What the kernels need to know about a PredictiveModel: kind, sizes, policy table and constants."""

    def __init__(self, kind, n, d, N, dt, policies, **consts):
        self.kind, self.n, self.d, self.N, self.dt = int(kind), int(n), int(d), int(N), float(dt)
        self.policies = [(int(k), [float(v) for v in par]) for k, par in policies]   # (policy kind, up to 4 params)
        self.m = len(self.policies)
        self.consts = dict(consts)


def pair_state_rows(Fx, bx, n):
    """Fx x <= bx  ->  two-sided rows lo <= f'x <= hi (opposite rows share one row)."""
    Fx = np.asarray(Fx, dtype=float).reshape(-1, n)
    bx = np.asarray(bx, dtype=float).reshape(-1)
    if Fx.shape[0] != bx.shape[0]:
        raise ValueError("Fx and bx disagree: %s vs %s" % (Fx.shape, bx.shape))
    rows, taken = [], set()
    for i in range(len(bx)):
        if i in taken:
            continue
        lo = -np.inf
        for j in range(i + 1, len(bx)):
            if j not in taken and np.array_equal(Fx[j], -Fx[i]):
                taken.add(j)
                lo = -bx[j]
                break
        rows.append((Fx[i].copy(), lo, float(bx[i])))
    return rows


def input_box(Fu, bu, d):
    """Fu u <= bu with one non-zero per row -> (u_lo, u_hi)."""
    Fu = np.asarray(Fu, dtype=float).reshape(-1, d)
    bu = np.asarray(bu, dtype=float).reshape(-1)
    lo, hi = np.full(d, -np.inf), np.full(d, np.inf)
    for row, b in zip(Fu, bu):
        nz = np.flatnonzero(row)
        if len(nz) != 1:
            raise ValueError("input constraints must be a box (one input per row of Fu)")
        k = nz[0]
        if row[k] > 0:
            hi[k] = min(hi[k], b / row[k])
        else:
            lo[k] = max(lo[k], b / row[k])
    if not (np.isfinite(lo).all() and np.isfinite(hi).all()):
        raise ValueError("every input needs a lower and an upper bound")
    return lo, hi


def make_config(model, n, d, N, NB, Q, R, Fx, bx, Fu, bu, Qslack, controller=abi.CTRL_BRANCH, Qf=None, dR=None,
                batch_capacity=1, device=0, **knobs):
    c = abi.Config()
    c.model, c.controller = model.kind, int(controller)
    if (model.n, model.d, model.N) != (n, d, N):
        raise ValueError("controller and model disagree on (n, d, N)")
    c.n, c.d, c.N, c.NB, c.m, c.dt = n, d, N, int(NB), model.m, model.dt
    if model.m > abi.MAX_POLICIES:
        raise ValueError("at most %d backup policies" % abi.MAX_POLICIES)
    for i, (kind, par) in enumerate(model.policies):
        c.policy_kind[i] = kind
        for k, v in enumerate(par[:4]):
            c.policy_param[i][k] = v
    Q = np.asarray(Q, dtype=float).reshape(n, n)
    Qf = Q if Qf is None else np.asarray(Qf, dtype=float).reshape(n, n)
    R = np.asarray(R, dtype=float).reshape(d, d)
    dR = np.zeros(d) if dR is None else np.asarray(dR, dtype=float).reshape(d)
    for i in range(n * n):
        c.Q[i], c.Qf[i] = Q.flat[i], Qf.flat[i]
    for i in range(d * d):
        c.R[i] = R.flat[i]
    for i in range(d):
        c.dR[i] = dR[i]
    Qslack = np.asarray(Qslack, dtype=float).reshape(-1)
    c.Qslack[0], c.Qslack[1] = Qslack[0], Qslack[1]
    rows = pair_state_rows(Fx, bx, n)
    if len(rows) > abi.MAX_ROWS:
        raise ValueError("at most %d two-sided state rows" % abi.MAX_ROWS)
    c.n_rows = len(rows)
    for j, (f, lo, hi) in enumerate(rows):
        for i in range(n):
            c.row_f[j][i] = f[i]
        c.row_lo[j], c.row_hi[j] = lo, hi
    lo, hi = input_box(Fu, bu, d)
    for a in range(d):
        c.u_lo[a], c.u_hi[a] = lo[a], hi[a]
    for key in ("veh_L", "veh_W", "Kpsi", "s1", "lane_lo", "lane_hi", "quad_margin"):
        setattr(c, key, float(model.consts.get(key, 0.0)))
    for key, val in knobs.items():
        if not hasattr(c, key):
            raise TypeError("unknown solver knob %r" % key)
        setattr(c, key, val)
    c.batch_capacity, c.device = int(batch_capacity), int(device)
    return c


def highway_spec(N, dt, policies, L, W, Kpsi, s1, N_lane_model=3):
    """highway_branch_dyn.PredictiveModel constants; the lane boundary inside the branching probability uses the
    model's own N_lane (default 3, highway_branch_dyn.py:264,:279)."""
    return ModelSpec(abi.MODEL_HIGHWAY, 4, 2, N, dt, policies, veh_L=L, veh_W=W, Kpsi=Kpsi, s1=s1,
                     lane_lo=W / 2.0, lane_hi=N_lane_model * 3.6 - W / 2.0)


def merge_spec(N, dt, policies, L, W, Kpsi, s1):
    """highway_branch_dyn.PredictiveModel_merge constants (:400-456): no lane-boundary term in the branching probability."""
    return ModelSpec(abi.MODEL_MERGE, 4, 2, N, dt, policies, veh_L=L, veh_W=W, Kpsi=Kpsi, s1=s1)


def quadruped_spec(N, dt, policies, L1, L2, col_tol, s1):
    return ModelSpec(abi.MODEL_QUADRUPED, 3, 3, N, dt, policies, s1=s1, quad_margin=(L1 + L2) / 2.0 + col_tol)
