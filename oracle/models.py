"""Closed-form CPU restatement (float64 numpy) of the reference's predictive models.

TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this package.  The product path (the CUDA library behind
`belief-planning_b200`) never does.

The reference builds these functions as CasADi expression graphs and differentiates them
automatically; here they are written out by hand.  They are pinned against the reference's own
graphs evaluated through `tests/golden/shims/casadi.py` (tests/golden/model_functions.npz,
tests/test_oracle_models.py).

Reference lines followed:
  highway  - /root/reference/highway_branch_dyn.py:17-34 (dubin), 38-39 (softsat), 54-148 (policies),
             151-162 (softmin/softmax), 174-187 (propagate_backup), 195-206 (lane_bdry_h),
             223-235 (veh_col, symbolic branch), 284-325 (evaluation methods), 337-398 (graph build)
  quadruped- /root/reference/quadruped_branch_dyn.py:14-27, 34-54, 135-145, 175-248
Policies are plain descriptors instead of Python closures:
  ("maintain",) ("brake",) ("lc", [x,y,v,psi] target) ("trackv", v0) | ("forward", v0) ("stop",)
"""
import numpy as np


def softmin(v, gamma):
    """sum(exp(-g v) v)/sum(exp(-g v)), evaluated with a max-shift (mathematically identical)."""
    v = np.asarray(v, dtype=float)
    e = np.exp(-gamma * (v - v.min()))
    return float((e * v).sum() / e.sum())


def softmax(v, gamma):
    v = np.asarray(v, dtype=float)
    e = np.exp(gamma * (v - v.max()))
    return float((e * v).sum() / e.sum())


def sigmoid(h):
    # softsat(h, 1) = (e^h - 1)/(e^h + 1)/2 + 1/2   (highway_branch_dyn.py:38-39)
    return 1.0 / (1.0 + np.exp(-h))


def soft_box_distance(dx, dy):
    """(dx e^dx + dy e^dy)/(e^dx + e^dy) and its partials w.r.t. dx, dy (veh_col core)."""
    mx = max(dx, dy)
    ex = np.exp(dx - mx)
    ey = np.exp(dy - mx)
    wx = ex / (ex + ey)
    wy = 1.0 - wx
    h = wx * dx + wy * dy
    return h, wx * (1.0 + dx - h), wy * (1.0 + dy - h)


class HighwayModel:
    """Dubins-car ego + obstacle with backup policies (n=4: x,y,v,psi; d=2: a,r)."""

    n = 4
    d = 2

    def __init__(self, N, policies, dt, L=4.0, W=2.5, Kpsi=0.1, s1=2.0, N_lane=3):
        self.N = int(N)
        self.policies = [tuple(p) for p in policies]
        self.m = len(self.policies)
        self.dt = float(dt)
        self.L = float(L)
        self.W = float(W)
        self.Kpsi = float(Kpsi)
        self.s1 = float(s1)
        # lane boundary used inside the branching probability (highway_branch_dyn.py:279; N_lane defaults to 3)
        self.LB = (self.W / 2.0, N_lane * 3.6 - self.W / 2.0)

    # -- dynamics ---------------------------------------------------------------------------
    def f(self, x, u):
        return np.array([x[2] * np.cos(x[3]), x[2] * np.sin(x[3]), u[0], u[1]])

    def step(self, x, u):
        return x + self.f(x, u) * self.dt

    def dyn_linearization(self, x, u):
        x = np.asarray(x, dtype=float)
        u = np.asarray(u, dtype=float)
        c, s, v, dt = np.cos(x[3]), np.sin(x[3]), x[2], self.dt
        A = np.eye(4)
        A[0, 2] = dt * c
        A[0, 3] = -dt * v * s
        A[1, 2] = dt * s
        A[1, 3] = dt * v * c
        B = np.zeros((4, 2))
        B[2, 0] = dt
        B[3, 1] = dt
        xp = self.step(x, u)
        C = xp - A @ x - B @ u
        return A, B, C, xp

    # -- backup policies (symbolic branch of each) ---------------------------------------------
    def policy(self, i, x):
        p = self.policies[i]
        kind = p[0]
        if kind == "maintain":
            return np.array([0.0, -self.Kpsi * x[3]])
        if kind == "brake":
            return np.array([softmax([-7.0, -x[2]], 5.0), -self.Kpsi * x[3]])
        if kind == "lc":
            t = p[1]
            return np.array([-0.8558 * (x[2] - t[2]),
                             -0.3162 * (x[1] - t[1]) - 3.9889 * (x[3] - t[3])])
        if kind == "trackv":
            return np.array([0.5 * (p[1] - x[2]), -self.Kpsi * x[3]])
        raise ValueError("unknown highway policy %r" % (kind,))

    def rollout(self, i, x):
        x = np.asarray(x, dtype=float)
        out = np.empty((self.N, 4))
        for t in range(self.N):
            x = self.step(x, self.policy(i, x))
            out[t] = x
        return out

    def zpred_eval(self, z):
        return np.hstack([self.rollout(i, z) for i in range(self.m)])

    def xpred_eval(self, x):
        return self.rollout(0, x), self.policy(0, np.asarray(x, dtype=float))

    # -- collision function -------------------------------------------------------------------
    def h_and_grad(self, x, z):
        ex = x[0] - z[0]
        ey = x[1] - z[1]
        h, gx, gy = soft_box_distance(abs(ex) - (self.L + 1.0), abs(ey) - (self.W + 0.2))
        dh = np.array([np.sign(ex) * gx, np.sign(ey) * gy, 0.0, 0.0])
        return h, dh

    def col_eval(self, x, z):
        x = np.asarray(x, dtype=float)
        h, dh = self.h_and_grad(x, np.asarray(z, dtype=float))
        return h - dh @ x, dh

    # -- branching probabilities ---------------------------------------------------------------
    def policy_safety(self, x, z):
        x1 = self.rollout(0, x)                      # ego under policy 0
        hi = np.empty(self.m)
        for i in range(self.m):
            z2 = self.rollout(i, z)
            vals = np.empty(2 * self.N)
            for t in range(self.N):
                vals[t], _, _ = soft_box_distance(abs(z2[t, 0] - x1[t, 0]) - (self.L + 2.0),
                                                  abs(z2[t, 1] - x1[t, 1]) - (self.W + 0.2))
                vals[self.N + t] = softmin([z2[t, 1] - self.LB[0], self.LB[1] - z2[t, 1]], 5.0)
            hi[i] = softmin(vals, 5.0)
        return hi

    def prob(self, x, z):
        e = np.exp(self.s1 * sigmoid(self.policy_safety(x, z)))
        return e / e.sum()

    def branch_eval(self, x, z, with_dp=False):
        x = np.asarray(x, dtype=float)
        z = np.asarray(z, dtype=float)
        p = self.prob(x, z)
        if not with_dp:
            return p, None
        # dp never enters any QP (BranchTree.J is never written, MPC_branch.py:76,:1085-1089);
        # provided by central differences for API completeness only.
        dp = np.empty((self.m, self.n))
        for k in range(self.n):
            e = np.zeros(self.n)
            e[k] = 1e-6
            dp[:, k] = (self.prob(x + e, z) - self.prob(x - e, z)) / 2e-6
        return p, dp


class QuadrupedModel:
    """Planar robot with body-frame velocities (n=3: x,y,theta; d=3: vx,vy,r)."""

    n = 3
    d = 3

    def __init__(self, N, policies, dt, L1=0.5, L2=1.0, col_tol=0.2, s1=2.0):
        self.N = int(N)
        self.policies = [tuple(p) for p in policies]
        self.m = len(self.policies)
        self.dt = float(dt)
        self.margin = (L1 + L2) / 2.0 + col_tol
        self.s1 = float(s1)

    def f(self, x, u):
        c, s = np.cos(x[2]), np.sin(x[2])
        return np.array([u[0] * c - u[1] * s, u[0] * s + u[1] * c, u[2]])

    def step(self, x, u):
        return x + self.f(x, u) * self.dt

    def dyn_linearization(self, x, u):
        x = np.asarray(x, dtype=float)
        u = np.asarray(u, dtype=float)
        c, s, dt = np.cos(x[2]), np.sin(x[2]), self.dt
        A = np.eye(3)
        A[0, 2] = dt * (-u[0] * s - u[1] * c)
        A[1, 2] = dt * (u[0] * c - u[1] * s)
        B = dt * np.array([[c, -s, 0.0], [s, c, 0.0], [0.0, 0.0, 1.0]])
        xp = self.step(x, u)
        C = xp - A @ x - B @ u
        return A, B, C, xp

    def policy(self, i, x):
        p = self.policies[i]
        if p[0] == "forward":
            return np.array([p[1], 0.0, 0.0])
        if p[0] == "stop":
            return np.zeros(3)
        raise ValueError("unknown quadruped policy %r" % (p[0],))

    def rollout(self, i, x):
        x = np.asarray(x, dtype=float)
        out = np.empty((self.N, 3))
        for t in range(self.N):
            x = self.step(x, self.policy(i, x))
            out[t] = x
        return out

    def zpred_eval(self, z):
        return np.hstack([self.rollout(i, z) for i in range(self.m)])

    def xpred_eval(self, x):
        return self.rollout(0, x), self.policy(0, np.asarray(x, dtype=float))

    def h_and_grad(self, x, z):
        ex = x[0] - z[0]
        ey = x[1] - z[1]
        return abs(ex) + abs(ey) - self.margin, np.array([np.sign(ex), np.sign(ey), 0.0])

    def col_eval(self, x, z):
        x = np.asarray(x, dtype=float)
        h, dh = self.h_and_grad(x, np.asarray(z, dtype=float))
        return h - dh @ x, dh

    def policy_safety(self, x, z):
        x1 = self.rollout(0, x)
        hi = np.empty(self.m)
        for i in range(self.m):
            z2 = self.rollout(i, z)
            d1 = np.abs(z2[:, 0] - x1[:, 0]) + np.abs(z2[:, 1] - x1[:, 1]) - self.margin
            hi[i] = softmin(d1, 5.0)
        return hi

    def prob(self, x, z):
        hi = self.policy_safety(x, z)
        e = np.exp(self.s1 * (hi - hi.max()))
        return e / e.sum()

    def branch_eval(self, x, z, with_dp=False):
        x = np.asarray(x, dtype=float)
        z = np.asarray(z, dtype=float)
        p = self.prob(x, z)
        if not with_dp:
            return p, None
        dp = np.empty((self.m, self.n))
        for k in range(self.n):
            e = np.zeros(self.n)
            e[k] = 1e-6
            dp[:, k] = (self.prob(x + e, z) - self.prob(x - e, z)) / 2e-6
        return p, dp
