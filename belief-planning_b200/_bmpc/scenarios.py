"""Seeded synthetic scenario batches; distributions anchored on the reference environments' initial
conditions and respawn windows (Highway_env_branch.py:67, Highway_env.py:225, quadruped_env.py:58).

Host-side numpy only: the arrays are handed to the library as float64 (device copies are made by the caller).
"""
import numpy as np

from . import abi, config

LANE_W = 3.6


def highway_batch(count, seed=1237, n_lanes=4):
    """Returns x0, z0, xref (count,4) and per-episode policy parameters (count,3,4) for [maintain, brake, lc]."""
    rng = np.random.default_rng(seed)
    lane_e = rng.integers(0, n_lanes, count)
    lane_o = rng.integers(0, n_lanes, count)
    x0 = np.column_stack([np.zeros(count), 1.8 + LANE_W * lane_e + rng.normal(0, 0.3, count), rng.uniform(15, 25, count),
                          np.clip(rng.normal(0, 0.03, count), -0.2, 0.2)])
    z0 = np.column_stack([rng.uniform(-15, 25, count), 1.8 + LANE_W * lane_o + rng.normal(0, 0.3, count),
                          rng.uniform(15, 25, count), rng.normal(0, 0.03, count)])
    xref = np.column_stack([np.zeros(count), 1.8 + LANE_W * lane_e, z0[:, 2] + rng.uniform(-3, 3, count), np.zeros(count)])
    # lane-change target of the obstacle: an adjacent lane (Highway_env_branch.py:96-118 picks it per step)
    step = np.where(lane_o == 0, 1, np.where(lane_o == n_lanes - 1, -1, rng.choice([-1, 1], count)))
    tgt = np.column_stack([np.zeros(count), 1.8 + LANE_W * (lane_o + step), z0[:, 2], np.zeros(count)])
    pp = np.zeros((count, 3, 4))
    pp[:, 2, :] = tgt
    return x0, z0, xref, pp


def highway_policies(names=("maintain", "brake", "lc"), lc_target=(0.5, 1.8, 15.0, 0.0), v0=20.0):
    table = {"maintain": (abi.POLICY_MAINTAIN, [0, 0, 0, 0]), "brake": (abi.POLICY_BRAKE, [0, 0, 0, 0]),
             "lc": (abi.POLICY_LC, list(lc_target)), "trackv": (abi.POLICY_TRACKV, [v0, 0, 0, 0]),
             "trackv_ref": (abi.POLICY_TRACKV_REF, [v0, 0, 0, 0]), "brake_ref": (abi.POLICY_BRAKE_REF, [0, 0, 0, 0])}
    return [table[n] for n in names]


def highway_config(policies=("maintain", "brake", "lc"), NB=2, N=8, lc_target=(0.5, 1.8, 15.0, 0.0), am=6.0, rm=0.3,
                   N_lane=4, W=2.5, L=4.0, batch_capacity=1, device=0, **knobs):
    """main_branch.py:24-48 + Init_MPC.initBranchMPC (:40-72) as a bmpc_config."""
    spec = config.highway_spec(N, 0.1, highway_policies(policies, lc_target), L, W, 0.1, 2.0)
    Fx = np.array([[0., 1., 0., 0.], [0., -1., 0., 0.], [0., 0., 0., 1.], [0., 0., 0., -1.]])
    bx = np.array([N_lane * LANE_W - W / 2, -W / 2, 0.25, 0.25])
    Fu = np.kron(np.eye(2), np.array([1., -1.])).T
    bu = np.array([am, am, rm, rm])
    return config.make_config(spec, 4, 2, N, NB, np.diag([0., 3., 3., 10.]), np.diag([1., 100.]), Fx, bx, Fu, bu,
                              np.array([0., 300.]), batch_capacity=batch_capacity, device=device, **knobs)


def merge_config(policies=("trackv", "brake"), NB=1, N=40, v0=20.0, am=7.0, rm=0.3, N_lane=2, W=2.5, L=4.0, ralpha=0.1,
                 batch_capacity=1, device=0, **knobs):
    """sim_merge of main_branch.py (:53-88): PredictiveModel_merge without lookup-table policies (`backupcons_normal`),
    initBranchMPC and BranchMPC_CVaR(ralpha=0.1) as a bmpc_config."""
    spec = config.merge_spec(N, 0.1, highway_policies(policies, v0=v0), L, W, 0.1, 2.0)
    Fx = np.array([[0., 1., 0., 0.], [0., -1., 0., 0.], [0., 0., 0., 1.], [0., 0., 0., -1.]])
    bx = np.array([N_lane * LANE_W - W / 2, -W / 2, 0.25, 0.25])
    Fu = np.kron(np.eye(2), np.array([1., -1.])).T
    bu = np.array([am, am, rm, rm])
    return config.make_config(spec, 4, 2, N, NB, np.diag([0., 3., 3., 10.]), np.diag([1., 100.]), Fx, bx, Fu, bu,
                              np.array([0., 300.]), controller=abi.CTRL_CVAR, cvar_alpha=ralpha,
                              batch_capacity=batch_capacity, device=device, **knobs)


def merge_batch(count, seed=1242, N_lane=2, merge_lane=1, merge_s=50.0, merge_R=300.0, v0=20.0, W=2.5, psimax=0.25):
    """Synthetic ramp scenes of the merge scenario: the ego somewhere along the ramp's centre line (offset 1.8 m as the
    reference's initial state, Highway_env_branch.py:315), an obstacle in the highway lane next to it, and the reference,
    state transform and bounds Highway_env_merge.step builds from the ramp tables at the ego's x (:357-363).
    Returns x0, z0, xref (count, 4), S (count, 4, 4), state bounds (count, 2, 2)."""
    from Highway_env_branch import merge_geometry
    rng = np.random.default_rng(seed)
    X1, X2, Y1, Y2, P1, P2 = merge_geometry(N_lane, merge_lane, merge_s, merge_R, 0)
    gx, gy, gpsi = np.append(X1, X2), np.append(Y1, Y2), np.append(P1, P2)
    xe = rng.uniform(15.0, merge_s, count)
    y0, psi0 = np.interp(xe, gx, gy), np.interp(xe, gx, gpsi)
    x0 = np.column_stack([xe, y0 + 1.8 + rng.normal(0, 0.15, count), rng.uniform(14.0, 22.0, count), psi0 + rng.normal(0, 0.02, count)])
    z0 = np.column_stack([xe + rng.uniform(-15.0, 12.0, count), np.full(count, 5.4) + rng.normal(0, 0.1, count),
                          rng.uniform(16.0, 22.0, count), np.zeros(count)])
    t = np.tan(psi0)
    S = np.tile(np.eye(4), (count, 1, 1))
    S[:, 1, 0] = -t
    xref = np.column_stack([np.zeros(count), -t * xe + y0 + 1.8, np.full(count, v0), psi0])
    bx = np.column_stack([-t * xe + y0 + 3.6 * merge_lane - W / 2, t * xe - y0 - W / 2, psi0 + psimax, -psi0 + psimax])
    return x0, z0, xref, S, bounds_from_bx(bx)


def bounds_from_bx(bx):
    """The reference's bx of the four one-sided rows [y <= b0, -y <= b1, psi <= b2, -psi <= b3] as (lo, hi) pairs."""
    bx = np.asarray(bx, dtype=float).reshape(-1, 4)
    return np.stack([np.stack([-bx[:, 1], bx[:, 0]], axis=1), np.stack([-bx[:, 3], bx[:, 2]], axis=1)], axis=1)


def euler_highway(x, u, dt=0.1):
    """vehicle plant (Highway_env_branch.py:39-41), batched."""
    out = np.array(x, dtype=float, copy=True)
    out[:, 0] += dt * x[:, 2] * np.cos(x[:, 3])
    out[:, 1] += dt * x[:, 2] * np.sin(x[:, 3])
    out[:, 2] += dt * u[:, 0]
    out[:, 3] += dt * u[:, 1]
    return out


# ------------------------------------------------------------------------------------------------------------
# quadruped (main_quadruped.py:14-41, Init_MPC.initquadBranchMPC :74-94)
# ------------------------------------------------------------------------------------------------------------
def quadruped_config(NB=2, N=25, v0=0.2, vxm=0.2, vym=0.1, rm=0.5, batch_capacity=1, device=0,
                     controller=abi.CTRL_PROX, **knobs):
    pol = [(abi.POLICY_FORWARD, [v0, 0, 0, 0]), (abi.POLICY_STOP, [0, 0, 0, 0])]
    spec = config.quadruped_spec(N, 0.2, pol, 0.5, 1.0, 0.2, 2.0)
    Fu = np.kron(np.eye(3), np.array([1., -1.])).T
    bu = np.array([vxm, 0., vym, vym, rm, rm])
    return config.make_config(spec, 3, 3, N, NB, np.eye(3), np.diag([1., 100., 1.]), np.empty((0, 3)), np.empty(0), Fu, bu,
                              np.array([0., 300.]), controller=controller, dR=np.array([0.9, 5., 1.]),
                              batch_capacity=batch_capacity, device=device, **knobs)


def quadruped_batch(count, seed=1238, goal=(5.0, -3.0, 0.0)):
    """ego near the origin, obstacle 1-4 m away at a random bearing (quadruped_env.py:58), goal as xRef."""
    rng = np.random.default_rng(seed)
    x0 = np.column_stack([rng.uniform(-1, 1, count), rng.uniform(-1, 1, count), rng.uniform(-np.pi, np.pi, count)])
    dist, bearing = rng.uniform(1, 4, count), rng.uniform(-np.pi, np.pi, count)
    z0 = np.column_stack([x0[:, 0] + dist * np.cos(bearing), x0[:, 1] + dist * np.sin(bearing),
                          rng.uniform(-np.pi, np.pi, count)])
    xref = np.tile(np.asarray(goal, dtype=float), (count, 1))
    return x0, z0, xref


def euler_quadruped(x, u, dt=0.2):
    """robot plant (quadruped_env.py:24-40), batched."""
    out = np.array(x, dtype=float, copy=True)
    c, s = np.cos(x[:, 2]), np.sin(x[:, 2])
    out[:, 0] += dt * (u[:, 0] * c - u[:, 1] * s)
    out[:, 1] += dt * (u[:, 0] * s + u[:, 1] * c)
    out[:, 2] += dt * u[:, 2]
    return out


# ------------------------------------------------------------------------------------------------------------
# belief-state MPC (Init_MPC.initMPCParams :7-34 + the HMM_constants the belief model reads)
# ------------------------------------------------------------------------------------------------------------
def belief_config(N=10, M=2, m=2, am=6.0, rm=0.3, N_lane=2, W=2.5, L=4.0, ylb=0.0, yub=7.2, col_alpha=5.0, s1=2.0,
                  tran_diag=0.3, thres=0.1, dt=0.1, Kpsi=0.1, batch_capacity=1, device=0, **knobs):
    pol = [(abi.POLICY_MAINTAIN, [0, 0, 0, 0]), (abi.POLICY_BRAKE, [0, 0, 0, 0])][:m]
    spec = config.ModelSpec(abi.MODEL_HIGHWAY, 4, 2, N, dt, pol, veh_L=L, veh_W=W, Kpsi=Kpsi, s1=s1, lane_lo=ylb, lane_hi=yub)
    Fx = np.array([[0., 1., 0., 0.], [0., -1., 0., 0.], [0., 0., 0., 1.], [0., 0., 0., -1.]])
    bx = np.array([N_lane * LANE_W - W / 2, -W / 2, 0.25, 0.25])
    Fu = np.kron(np.eye(2), np.array([1., -1.])).T
    bu = np.array([am, 0.5 * am, rm, rm])
    return config.make_config(spec, 4, 2, N, 1, np.diag([0., 0.5, 0.2, 5.]), np.diag([30., 100.]), Fx, bx, Fu, bu,
                              np.array([0., 1000.]), controller=abi.CTRL_BELIEF, Qf=np.zeros((4, 4)), hmm_M=M,
                              hmm_col_alpha=col_alpha, hmm_tran_diag=tran_diag, hmm_thres=thres,
                              batch_capacity=batch_capacity, device=device, **knobs)
