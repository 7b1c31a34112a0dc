"""CPU restatement (float64 numpy) of the numeric functions of the reference's belief-state model.  TEST INFRASTRUCTURE.

Follows /root/reference/HMM_backup_dyn.py: backup_maintain :105, backup_brake :107-109 (numeric softmax(-5,-v,3)),
propagate_backup :122-132, PredictiveModel.generate_backup_traj :204-214 (rows flattened column-major, the way
casadi.reshape does), module-level generate_backup_traj with sensitivity :54-85 and dubin_f_x :43-52 (central
differences, h = 1e-6), veh_col numeric branch :145-157, lane_bdry_h :134, softmin :111, softsat :94, backup_trans
:96-101, backup_input_prob :103, and the belief update of Highway_env.py:251-256.
Pinned by tests/golden/hmm_functions.npz (produced by the unmodified module, tests/golden/make_golden.py).
"""
import numpy as np

MAINTAIN, BRAKE = 0, 1


def softsat(x, s):
    return (np.exp(s * x) - 1) / (np.exp(s * x) + 1) * 0.5 + 0.5


def softmin2(x, y, gamma):
    mn = min(x, y)
    ex, ey = np.exp(-gamma * (x - mn)), np.exp(-gamma * (y - mn))
    return (ex * x + ey * y) / (ex + ey)


def softmax2(x, y, gamma):
    mx = max(x, y)
    ex, ey = np.exp(gamma * (x - mx)), np.exp(gamma * (y - mx))
    return (ex * x + ey * y) / (ex + ey)


def policy(kind, x, Kpsi):
    if kind == MAINTAIN:
        return np.array([0.0, -Kpsi * x[3]])
    if kind == BRAKE:
        return np.array([softmax2(-5.0, -x[2], 3.0), -Kpsi * x[3]])
    raise ValueError(kind)


def dubin(x, u):
    return np.array([x[2] * np.cos(x[3]), x[2] * np.sin(x[3]), u[0], u[1]])


def backup_rollout(x0, kinds, N, dt, Kpsi):
    """x0 (M,4) -> (M*m, N*4); row m*i+j is agent i under policy j, flattened column-major (component-major)."""
    x0 = np.asarray(x0, dtype=float)
    M, m = x0.shape[0], len(kinds)
    out = np.empty((M * m, N * 4))
    for i in range(M):
        for j, kind in enumerate(kinds):
            x = x0[i].copy()
            xs = np.empty((N, 4))
            for t in range(N):
                x = x + dubin(x, policy(kind, x, Kpsi)) * dt
                xs[t] = x
            out[m * i + j] = xs.reshape(-1, order="F")
    return out


def rollout_sensitivity(x, kind, steps, ts, f0, Kpsi):
    """states, sensitivity matrices Q (dx_t/dx_0) and xdot - f0 before each of `steps` Euler steps."""
    x = np.asarray(x, dtype=float).copy()
    con = lambda s: policy(kind, s, Kpsi)
    Q = np.eye(4)
    xx, QQ, Qt = [], [], []
    h = 1e-6
    for _ in range(steps):
        QQ.append(Q.copy())
        u = con(x)
        xdot = dubin(x, u)
        dudx = np.array([(con(x + h * np.eye(4)[k]) - con(x - h * np.eye(4)[k])) / 2 / h for k in range(4)])   # (4,2)
        ja = np.vstack([[0, 0, np.cos(x[3]), -x[2] * np.sin(x[3])], [0, 0, np.sin(x[3]), x[2] * np.cos(x[3])], dudx.T])
        xx.append(x.copy())
        Qt.append(xdot - f0)
        x = x + xdot * ts
        Q = Q + ja @ Q * ts
    return np.array(xx), np.array(QQ), np.array(Qt)


def veh_col_norm(x1, x2, size, alpha=1.0, clip=True):
    dx = (abs(x1[0] - x2[0]) - size[0]) / size[0]
    dy = (abs(x1[1] - x2[1]) - size[1]) / size[1]
    if clip:
        dx, dy = np.clip(dx, -5, 5), np.clip(dy, -5, 5)
    mx = max(dx, dy)
    ex, ey = np.exp(alpha * (dx - mx)), np.exp(alpha * (dy - mx))
    return (dx * ex + dy * ey) / (ex + ey)


def safety(ego, xb, L, W, ylb, yub, col_alpha, clip=True):
    """h of one (agent, policy): softmin(veh_col(ego, backup state), lane_bdry_h(backup state), col_alpha) (:255)."""
    lane = softmin2(xb[1] - ylb, yub - xb[1], 5.0)
    return softmin2(veh_col_norm(ego, xb, [L + 1.0, W + 0.2], 1.0, clip), lane, col_alpha)


def backup_trans(h, s1, tran_diag):
    mh = softsat(np.asarray(h, dtype=float), s1)
    m = len(mh)
    return np.kron((1 - tran_diag) * np.ones((m, 1)), (mh / mh.sum())[None, :]) + tran_diag * np.eye(m)


def belief_step(b, H, cbf=None, c2=0.5, s2=3.0):
    """b+ = b H; with cbf (Highway_env.py:253-256): b_j <- b_j softsat(cbf_j - c2, s2), normalised."""
    bn = np.asarray(b, dtype=float) @ H
    if cbf is not None:
        bn = bn * softsat(np.asarray(cbf, dtype=float) - c2, s2)
        bn = bn / bn.sum()
    return bn
