// Closed-form predictive models evaluated on the device (rows M1-M5 of SURVEY.md 8a).
// The reference builds these as CasADi graphs and differentiates them automatically
// (highway_branch_dyn.py:363-398, quadruped_branch_dyn.py:218-248); here they are written out.
#pragma once
#include "bmpc_params.h"

// softmin/softmax accumulators with a running shift: sum(bmpc_exp(-g v) v)/sum(bmpc_exp(-g v)) is invariant to
// subtracting a constant from the exponent, so the shifted form equals the reference's unshifted one
// (highway_branch_dyn.py:151-162) without its overflow.
struct SoftMinAcc {
  real vmin, num, den, gamma;
  BMPC_D SoftMinAcc(real g) : vmin(1e300), num(0), den(0), gamma(g) {}
  BMPC_D void add(real v) {
    real e = 1.0;   // exp(0) when v is the new minimum
    if (v < vmin) {
      const real sc = (den > 0) ? bmpc_exp(-gamma * (vmin - v)) : 0.0;
      num *= sc;
      den *= sc;
      vmin = v;
    } else {
      e = bmpc_exp(-gamma * (v - vmin));
    }
    num += e * v;
    den += e;
  }
  BMPC_D real value() const { return bmpc_div(num, den); }
};

// (dx e^dx + dy e^dy)/(e^dx + e^dy) and partials; veh_col core (highway_branch_dyn.py:231-234)
BMPC_D void soft_box(real dx, real dy, real& h, real& gx, real& gy) {
  // one of the two shifted exponentials is exp(0) = 1: a single call
  const real e = bmpc_exp(-fabs(dx - dy));
  const real ex = (dx >= dy) ? 1.0 : e, ey = (dx >= dy) ? e : 1.0;
  const real wx = bmpc_div(ex, ex + ey), wy = 1.0 - wx;
  h = wx * dx + wy * dy;
  gx = wx * (1.0 + dx - h);
  gy = wy * (1.0 + dy - h);
}

BMPC_D real sgn(real v) { return (v > 0) - (v < 0); }

// piecewise-linear lookup table of the handle (casadi interpolant 'linear', main_branch.py:78-79): segment by bisection, the
// end segments continue outside the grid
BMPC_D real bmpc_lookup(const KParams& P, real x) {
  if (P.lut_n < 2) return 0.0;
  int lo = 0, hi = P.lut_n - 1;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (P.lut_x[mid] <= x) lo = mid; else hi = mid;
  }
  const real x0 = P.lut_x[lo], y0 = P.lut_y[lo];
  return y0 + bmpc_div(P.lut_y[lo + 1] - y0, P.lut_x[lo + 1] - x0) * (x - x0);
}

// ------------------------------------------------------------------------------------------
// Highway: x = (x, y, v, psi), u = (a, r).   highway_branch_dyn.py
// ------------------------------------------------------------------------------------------
struct HighwayModel {
  static constexpr int NX = 4;    // Riccati state dimension
  static constexpr int NXP = 4;   // physical state dimension
  static constexpr bool RATE = false;
  static constexpr int NU = 2;
  static constexpr int NLIN = 4;  // dt*cos, -dt*v*sin, dt*sin, dt*v*cos  (the four non-trivial entries of A)
  static constexpr int NCC = 2;   // C has two non-zero entries
  static constexpr bool kStateTransform = false;   // per-episode state transform S and bounds bx (merge scenario only)

  // one Euler step of dubin (:17-34, :369)
  static constexpr int HEADING = 3;   // index of the heading angle: the only argument of a transcendental in step / linearize
  BMPC_D static void step(const KParams& P, const real* x, const real* u, real* xn) {
    real s, c;
    bmpc_sincos(x[3], &s, &c);
    step_sc(P, x, u, s, c, xn);
  }
  // ... with sine and cosine of the heading supplied by the caller (Solver::rollouts evaluates three headings side by side)
  BMPC_D static void step_sc(const KParams& P, const real* x, const real* u, real s, real c, real* xn) {
    xn[0] = x[0] + P.dt * (x[2] * c);
    xn[1] = x[1] + P.dt * (x[2] * s);
    xn[2] = x[2] + P.dt * u[0];
    xn[3] = x[3] + P.dt * u[1];
  }

  // symbolic branch of each backup policy (:54-148)
  BMPC_DN static void policy(const KParams& P, int kind, const real* par, const real* x, real* u) {
    switch (kind) {
      case BMPC_POLICY_MAINTAIN:
        u[0] = 0.0;
        u[1] = -P.Kpsi * x[3];
        break;
      case BMPC_POLICY_BRAKE: {
        // softmax([-7, -v], 5) = (a e^{5a} + b e^{5b})/(e^{5a} + e^{5b})
        const real a = -7.0, b = -x[2];
        const real e = bmpc_exp(-5.0 * fabs(a - b));
        const real ea = (a >= b) ? 1.0 : e, eb = (a >= b) ? e : 1.0;
        u[0] = bmpc_div(ea * a + eb * b, ea + eb);
        u[1] = -P.Kpsi * x[3];
        break;
      }
      case BMPC_POLICY_LC:
        u[0] = -0.8558 * (x[2] - par[2]);
        u[1] = -0.3162 * (x[1] - par[1]) - 3.9889 * (x[3] - par[3]);
        break;
      case BMPC_POLICY_TRACKV:
        u[0] = 0.5 * (par[0] - x[2]);
        u[1] = -P.Kpsi * x[3];
        break;
      case BMPC_POLICY_TRACKV_REF:   // ramp policies of the merge scenario (:89-96, :122-131)
        u[0] = 0.5 * (par[0] - x[2]);
        u[1] = bmpc_lookup(P, x[0]) - P.Kpsi * x[3];
        break;
      case BMPC_POLICY_BRAKE_REF: {
        const real a = -5.0, b = -x[2];
        const real e = bmpc_exp(-3.0 * fabs(a - b));
        const real ea = (a >= b) ? 1.0 : e, eb = (a >= b) ? e : 1.0;
        u[0] = bmpc_div(ea * a + eb * b, ea + eb);
        u[1] = bmpc_lookup(P, x[0]) - P.Kpsi * x[3];
        break;
      }
      default:
        u[0] = 0.0;
        u[1] = 0.0;
    }
  }

  // The same for the rollouts of the tree expansion, in line and on values: the out-of-line switch above is an indirect
  // branch with its arguments in local memory, paid at every step of a rollout.  `par` is the caller's register copy of the
  // policy's parameters; only the lookup-table policies of the merge scenario go through a call.
  struct Input2 { real a, r; };
  BMPC_DN static Input2 policy_ref(const KParams& P, int kind, real par0, real x0, real v, real psi) {
    const real xs[4] = {x0, 0.0, v, psi}, ps[4] = {par0, 0.0, 0.0, 0.0};
    real u[2];
    policy(P, kind, ps, xs, u);
    return Input2{u[0], u[1]};
  }
  BMPC_D static void policy_fast(const KParams& P, int kind, const real* par, const real* x, real* u) {
    if (kind == BMPC_POLICY_MAINTAIN) {
      u[0] = 0.0;
      u[1] = -P.Kpsi * x[3];
    } else if (kind == BMPC_POLICY_BRAKE) {
      const real a = -7.0, b = -x[2];
      const real e = bmpc_exp(-5.0 * fabs(a - b));
      const real ea = (a >= b) ? 1.0 : e, eb = (a >= b) ? e : 1.0;
      u[0] = bmpc_div(ea * a + eb * b, ea + eb);
      u[1] = -P.Kpsi * x[3];
    } else if (kind == BMPC_POLICY_LC) {
      u[0] = -0.8558 * (x[2] - par[2]);
      u[1] = -0.3162 * (x[1] - par[1]) - 3.9889 * (x[3] - par[3]);
    } else if (kind == BMPC_POLICY_TRACKV) {
      u[0] = 0.5 * (par[0] - x[2]);
      u[1] = -P.Kpsi * x[3];
    } else {
      const Input2 r = policy_ref(P, kind, par[0], x[0], x[2], x[3]);
      u[0] = r.a;
      u[1] = r.r;
    }
  }

  // dyn_linearization (:284-291): compressed A, C and the successor state
  BMPC_D static void linearize(const KParams& P, const real* x, const real* u, real* lin, real* cc, real* xn) {
    real s, c;
    bmpc_sincos(x[3], &s, &c);
    linearize_sc(P, x, u, s, c, lin, cc, xn);
  }
  BMPC_D static void linearize_sc(const KParams& P, const real* x, const real* u, real s, real c, real* lin, real* cc, real* xn) {
    lin_only(P, x, u, s, c, lin, cc);
    step_sc(P, x, u, s, c, xn);
  }
  // compressed A and C alone (the successor state is step_sc's)
  BMPC_D static void lin_only(const KParams& P, const real* x, const real*, real s, real c, real* lin, real* cc) {
    const real dt = P.dt, v = x[2];
    lin[0] = dt * c;
    lin[1] = -dt * v * s;
    lin[2] = dt * s;
    lin[3] = dt * v * c;
    // C = xp - A x - B u; rows 2,3 vanish identically
    cc[0] = dt * v * x[3] * s;
    cc[1] = -dt * v * x[3] * c;
  }

  BMPC_D static void mulA(const KParams&, const real* lin, const real* x, real* y) {
    y[0] = x[0] + lin[0] * x[2] + lin[1] * x[3];
    y[1] = x[1] + lin[2] * x[2] + lin[3] * x[3];
    y[2] = x[2];
    y[3] = x[3];
  }
  BMPC_D static void mulAT(const KParams&, const real* lin, const real* g, real* y) {
    y[0] = g[0];
    y[1] = g[1];
    y[2] = g[2] + lin[0] * g[0] + lin[2] * g[1];
    y[3] = g[3] + lin[1] * g[0] + lin[3] * g[1];
  }
  // y += B u ;  r = B' g   (B = dt [0;0;I2], constant)
  BMPC_D static void addBu(const KParams& P, const real*, const real* u, real* y) {
    y[2] += P.dt * u[0];
    y[3] += P.dt * u[1];
  }
  BMPC_D static void mulBT(const KParams& P, const real*, const real* g, real* r) {
    r[0] = P.dt * g[2];
    r[1] = P.dt * g[3];
  }
  BMPC_D static void addC(const real* cc, real* y) {
    y[0] += cc[0];
    y[1] += cc[1];
  }
  BMPC_D static void expandC(const real* cc, real* C) {
    C[0] = cc[0];
    C[1] = cc[1];
    C[2] = 0.0;
    C[3] = 0.0;
  }
  BMPC_D static void denseA(const KParams&, const real* lin, real* A) {
    for (int i = 0; i < 16; ++i) A[i] = 0.0;
    A[0] = A[5] = A[10] = A[15] = 1.0;
    A[2] = lin[0];
    A[3] = lin[1];
    A[6] = lin[2];
    A[7] = lin[3];
  }
  BMPC_D static void denseB(const KParams& P, const real*, real* B) {
    for (int i = 0; i < 8; ++i) B[i] = 0.0;
    B[4] = P.dt;
    B[7] = P.dt;
  }

  // collision function h(x,z) with size [L+1, W+0.2] and its gradient in (x,y) (:223-235, :386, :392)
  BMPC_D static void collision(const KParams& P, const real* x, const real* zxy, real& h, real& dhx, real& dhy) {
    const real ex = x[0] - zxy[0], ey = x[1] - zxy[1];
    real gx, gy;
    soft_box(fabs(ex) - (P.veh_L + 1.0), fabs(ey) - (P.veh_W + 0.2), h, gx, gy);
    dhx = sgn(ex) * gx;
    dhy = sgn(ey) * gy;
  }

  // Safety value of one policy: BF_traj (:337-349) over the obstacle rollout under `kind` against
  // the ego rollout under policy 0; also returns the obstacle trajectory through `emit(t, z)`.
  template <class Emit>
  BMPC_D static real policy_safety(const KParams& P, int kind, const real* par, int kind0, const real* par0,
                                       const real* xe0, const real* z0, real* zlast, int nsteps, Emit emit) {
    real xe[4] = {xe0[0], xe0[1], xe0[2], xe0[3]};
    real z[4] = {z0[0], z0[1], z0[2], z0[3]};
    SoftMinAcc acc(5.0);
#pragma unroll 1
    for (int t = 0; t < nsteps; ++t) {
      real u[2], xn[4];
      policy(P, kind0, par0, xe, u);
      step(P, xe, u, xn);
      for (int i = 0; i < 4; ++i) xe[i] = xn[i];
      policy(P, kind, par, z, u);
      step(P, z, u, xn);
      for (int i = 0; i < 4; ++i) z[i] = xn[i];
      emit(t, z);
      real h, gx, gy;
      soft_box(fabs(z[0] - xe[0]) - (P.veh_L + 2.0), fabs(z[1] - xe[1]) - (P.veh_W + 0.2), h, gx, gy);
      acc.add(h);
      // lane_bdry_h (:195-206): softmin([y - lb, ub - y], 5)
      const real a = z[1] - P.lane_lo, b = P.lane_hi - z[1];
      const real e = bmpc_exp(-5.0 * fabs(a - b));
      const real ea = (a <= b) ? 1.0 : e, eb = (a <= b) ? e : 1.0;
      acc.add(bmpc_div(ea * a + eb * b, ea + eb));
    }
    for (int i = 0; i < 4; ++i) zlast[i] = z[i];
    return acc.value();
  }

  // the two safety terms of one step of BF_traj (:337-349): vehicle distance veh_col(z, x1, [L+2, W+0.2]) and lane
  // boundary softmin([y - lb, ub - y], 5) (:195-206); their soft-min over the N steps is the safety value of the policy
  static constexpr int NSAFE = 2;
  BMPC_D static void safety_terms(const KParams& P, const real* z, real xe0, real xe1, real* v) {
    real gx, gy;
    soft_box(fabs(z[0] - xe0) - (P.veh_L + 2.0), fabs(z[1] - xe1) - (P.veh_W + 0.2), v[0], gx, gy);
    const real a = z[1] - P.lane_lo, b = P.lane_hi - z[1];
    const real e = bmpc_exp(-5.0 * fabs(a - b));
    const real ea = (a <= b) ? 1.0 : e, eb = (a <= b) ? e : 1.0;
    v[1] = bmpc_div(ea * a + eb * b, ea + eb);
  }

  // un-normalised branch weight bmpc_exp(s1 * softsat(hi, 1)) (:355-359); softsat(h,1) == sigmoid(h)
  BMPC_D static real branch_weight(const KParams& P, real hi, real /*himax*/) {
    return bmpc_exp(bmpc_div(P.s1, 1.0 + bmpc_exp(-hi)));
  }
  static constexpr bool kWeightNeedsMax = false;
};

// ------------------------------------------------------------------------------------------
// Merge scenario: PredictiveModel_merge (highway_branch_dyn.py:400-502).  Vehicle, policies, linearisation and collision
// row are the highway model's; the safety value of a policy measures the vehicle distance only, with size [L+1, W+0.2]
// (BF_traj :452-456, no lane-boundary term), and the controller is called with a state transform S and its own state
// bounds per solve (Highway_env_branch.py:352-366), which the solver keeps per episode (kStateTransform).
// ------------------------------------------------------------------------------------------
struct MergeModel : HighwayModel {
  static constexpr bool kStateTransform = true;
  static constexpr int NSAFE = 1;
  BMPC_D static void safety_terms(const KParams& P, const real* z, real xe0, real xe1, real* v) {
    real gx, gy;
    soft_box(fabs(z[0] - xe0) - (P.veh_L + 1.0), fabs(z[1] - xe1) - (P.veh_W + 0.2), v[0], gx, gy);
  }
  template <class Emit>
  BMPC_D static real policy_safety(const KParams& P, int kind, const real* par, int kind0, const real* par0,
                                   const real* xe0, const real* z0, real* zlast, int nsteps, Emit emit) {
    real xe[4] = {xe0[0], xe0[1], xe0[2], xe0[3]};
    real z[4] = {z0[0], z0[1], z0[2], z0[3]};
    SoftMinAcc acc(5.0);
#pragma unroll 1
    for (int t = 0; t < nsteps; ++t) {
      real u[2], xn[4];
      policy(P, kind0, par0, xe, u);
      step(P, xe, u, xn);
      for (int i = 0; i < 4; ++i) xe[i] = xn[i];
      policy(P, kind, par, z, u);
      step(P, z, u, xn);
      for (int i = 0; i < 4; ++i) z[i] = xn[i];
      emit(t, z);
      real v;
      safety_terms(P, z, xe[0], xe[1], &v);
      acc.add(v);
    }
    for (int i = 0; i < 4; ++i) zlast[i] = z[i];
    return acc.value();
  }
};

// ------------------------------------------------------------------------------------------
// Quadruped: x = (x, y, theta), u = (vx, vy, r).   quadruped_branch_dyn.py
// ------------------------------------------------------------------------------------------
struct QuadrupedModel {
  static constexpr int NX = 3;
  static constexpr int NXP = 3;
  static constexpr bool RATE = false;
  static constexpr bool kStateTransform = false;
  static constexpr int NU = 3;
  static constexpr int NLIN = 4;  // dt*cos, dt*sin, A[0][2], A[1][2]
  static constexpr int NCC = 2;

  static constexpr int HEADING = 2;
  BMPC_D static void step(const KParams& P, const real* x, const real* u, real* xn) {
    real s, c;
    bmpc_sincos(x[2], &s, &c);
    step_sc(P, x, u, s, c, xn);
  }
  BMPC_D static void step_sc(const KParams& P, const real* x, const real* u, real s, real c, real* xn) {
    xn[0] = x[0] + P.dt * (u[0] * c - u[1] * s);
    xn[1] = x[1] + P.dt * (u[0] * s + u[1] * c);
    xn[2] = x[2] + P.dt * u[2];
  }
  BMPC_D static void policy(const KParams&, int kind, const real* par, const real*, real* u) {
    u[0] = (kind == BMPC_POLICY_FORWARD) ? par[0] : 0.0;
    u[1] = 0.0;
    u[2] = 0.0;
  }
  BMPC_D static void policy_fast(const KParams& P, int kind, const real* par, const real* x, real* u) { policy(P, kind, par, x, u); }
  BMPC_D static void linearize(const KParams& P, const real* x, const real* u, real* lin, real* cc, real* xn) {
    real s, c;
    bmpc_sincos(x[2], &s, &c);
    linearize_sc(P, x, u, s, c, lin, cc, xn);
  }
  BMPC_D static void linearize_sc(const KParams& P, const real* x, const real* u, real s, real c, real* lin, real* cc, real* xn) {
    lin_only(P, x, u, s, c, lin, cc);
    step_sc(P, x, u, s, c, xn);
  }
  BMPC_D static void lin_only(const KParams& P, const real* x, const real* u, real s, real c, real* lin, real* cc) {
    const real dt = P.dt;
    lin[0] = dt * c;
    lin[1] = dt * s;
    lin[2] = dt * (-u[0] * s - u[1] * c);
    lin[3] = dt * (u[0] * c - u[1] * s);
    // C = xp - A x - B u = -A[:,2]*theta in rows 0,1
    cc[0] = -lin[2] * x[2];
    cc[1] = -lin[3] * x[2];
  }
  BMPC_D static void mulA(const KParams&, const real* lin, const real* x, real* y) {
    y[0] = x[0] + lin[2] * x[2];
    y[1] = x[1] + lin[3] * x[2];
    y[2] = x[2];
  }
  BMPC_D static void mulAT(const KParams&, const real* lin, const real* g, real* y) {
    y[0] = g[0];
    y[1] = g[1];
    y[2] = g[2] + lin[2] * g[0] + lin[3] * g[1];
  }
  BMPC_D static void addBu(const KParams& P, const real* lin, const real* u, real* y) {
    y[0] += lin[0] * u[0] - lin[1] * u[1];
    y[1] += lin[1] * u[0] + lin[0] * u[1];
    y[2] += P.dt * u[2];
  }
  BMPC_D static void mulBT(const KParams& P, const real* lin, const real* g, real* r) {
    r[0] = lin[0] * g[0] + lin[1] * g[1];
    r[1] = -lin[1] * g[0] + lin[0] * g[1];
    r[2] = P.dt * g[2];
  }
  BMPC_D static void addC(const real* cc, real* y) {
    y[0] += cc[0];
    y[1] += cc[1];
  }
  BMPC_D static void expandC(const real* cc, real* C) {
    C[0] = cc[0];
    C[1] = cc[1];
    C[2] = 0.0;
  }
  BMPC_D static void denseA(const KParams&, const real* lin, real* A) {
    for (int i = 0; i < 9; ++i) A[i] = 0.0;
    A[0] = A[4] = A[8] = 1.0;
    A[2] = lin[2];
    A[5] = lin[3];
  }
  BMPC_D static void denseB(const KParams& P, const real* lin, real* B) {
    B[0] = lin[0]; B[1] = -lin[1]; B[2] = 0.0;
    B[3] = lin[1]; B[4] = lin[0];  B[5] = 0.0;
    B[6] = 0.0;    B[7] = 0.0;     B[8] = P.dt;
  }
  // robot_col, symbolic branch: L1 distance minus margin (:135-145)
  BMPC_D static void collision(const KParams& P, const real* x, const real* zxy, real& h, real& dhx, real& dhy) {
    const real ex = x[0] - zxy[0], ey = x[1] - zxy[1];
    h = fabs(ex) + fabs(ey) - P.quad_margin;
    dhx = sgn(ex);
    dhy = sgn(ey);
  }
  template <class Emit>
  BMPC_D static real policy_safety(const KParams& P, int kind, const real* par, int kind0, const real* par0,
                                       const real* xe0, const real* z0, real* zlast, int nsteps, Emit emit) {
    real xe[3] = {xe0[0], xe0[1], xe0[2]};
    real z[3] = {z0[0], z0[1], z0[2]};
    SoftMinAcc acc(5.0);
#pragma unroll 1
    for (int t = 0; t < nsteps; ++t) {
      real u[3], xn[3];
      policy(P, kind0, par0, xe, u);
      step(P, xe, u, xn);
      for (int i = 0; i < 3; ++i) xe[i] = xn[i];
      policy(P, kind, par, z, u);
      step(P, z, u, xn);
      for (int i = 0; i < 3; ++i) z[i] = xn[i];
      emit(t, z);
      acc.add(fabs(z[0] - xe[0]) + fabs(z[1] - xe[1]) - P.quad_margin);
    }
    for (int i = 0; i < 3; ++i) zlast[i] = z[i];
    return acc.value();
  }
  static constexpr int NSAFE = 1;   // robot_col of the step (:204-210)
  BMPC_D static void safety_terms(const KParams& P, const real* z, real xe0, real xe1, real* v) {
    v[0] = fabs(z[0] - xe0) + fabs(z[1] - xe1) - P.quad_margin;
    v[1] = 0.0;
  }
  // bmpc_exp(s1*hi), shifted by the group maximum (normalisation cancels the shift) (:211-216)
  BMPC_D static real branch_weight(const KParams& P, real hi, real himax) { return bmpc_exp(P.s1 * (hi - himax)); }
  static constexpr bool kWeightNeedsMax = true;
};

// ------------------------------------------------------------------------------------------
// Belief-state model (HMM_backup_dyn.PredictiveModel.calc_xp_expr, :238-267): the other agents follow one of m backup
// policies; the belief b[i][:] over agent i's policies moves with the transition matrix H_i(x) (backup_trans :96-101) built
// from the safety values h_ij = softmin(veh_col(x, xb_ij), lane_bdry_h(xb_ij), col_alpha) (:255).  Closed forms of the
// CasADi graphs; the symbolic veh_col branch is normalised by the box size and not clipped (:136-143).
// ------------------------------------------------------------------------------------------
struct BeliefModel {
  // h and its gradient with respect to the ego position (x, y)
  BMPC_D static real safety(const KParams& P, const real* x, const real* xb, real& gx, real& gy) {
    const real s0 = P.veh_L + 1.0, s1 = P.veh_W + 0.2;
    const real ex = x[0] - xb[0], ey = x[1] - xb[1];
    const real dx = bmpc_div(fabs(ex) - s0, s0), dy = bmpc_div(fabs(ey) - s1, s1);
    real v, wgx, wgy;
    soft_box(dx, dy, v, wgx, wgy);                          // alpha = 1: (dx e^dx + dy e^dy)/(e^dx + e^dy) and its partials
    const real gvx = bmpc_div(wgx * sgn(ex), s0), gvy = bmpc_div(wgy * sgn(ey), s1);
    // lane_bdry_h: softmin(y - ylb, yub - y, 5)
    const real a = xb[1] - P.lane_lo, b = P.lane_hi - xb[1];
    const real e = bmpc_exp(-5.0 * fabs(a - b));
    const real ea = (a <= b) ? 1.0 : e, eb = (a <= b) ? e : 1.0;
    const real lane = bmpc_div(ea * a + eb * b, ea + eb);
    // softmin(v, lane, col_alpha)
    const real g = P.hmm_col_alpha;
    const real e2 = bmpc_exp(-g * fabs(v - lane));
    const real ev = (v <= lane) ? 1.0 : e2, el = (v <= lane) ? e2 : 1.0;
    const real wv = bmpc_div(ev, ev + el);
    const real h = wv * v + (1.0 - wv) * lane;
    const real dh = wv * (1.0 - g * (v - h));
    gx = dh * gvx;
    gy = dh * gvy;
    return h;
  }
  // b+ (column-major flat belief, entry k M + i = agent i, policy k) from b and the ego position; optionally the
  // Jacobian blocks: dbx[q][2] = d b+_q / d(x, y), and Hm[i][r][k] = H_i[r][k] (d b+_(i,k) / d b_(i,r))
  BMPC_D static void transition(const KParams& P, const real* x, const real* xbackup, int xb_stride, int slice, const real* b,
                                real* bp, real* dbx, real* Hm) {
    const int M = P.hmm_M, m = P.zm;
    for (int i = 0; i < M; ++i) {
      real mh[BMPC_MAX_POLICIES], gh[BMPC_MAX_POLICIES][2], msum = 0.0, bsum = 0.0;
      for (int j = 0; j < m; ++j) {
        const real* xb = xbackup + (size_t)(m * i + j) * xb_stride + 4 * slice;
        const real h = safety(P, x, xb, gh[j][0], gh[j][1]);
        mh[j] = bmpc_div(1.0, 1.0 + bmpc_exp(-P.s1 * h));   // softsat(h, s1) == sigmoid(s1 h)
        msum += mh[j];
        bsum += b[j * M + i];
      }
      const real tau = P.hmm_tran_diag;
      for (int k = 0; k < m; ++k) {
        const real pik = bmpc_div(mh[k], msum);
        bp[k * M + i] = (1.0 - tau) * bsum * pik + tau * b[k * M + i];
        if (Hm)
          for (int r = 0; r < m; ++r) Hm[(i * m + r) * m + k] = (1.0 - tau) * pik + (r == k ? tau : 0.0);
        if (dbx) {
          real ax = 0.0, ay = 0.0;
          for (int l = 0; l < m; ++l) {
            const real dpi = bmpc_div(((l == k) ? 1.0 : 0.0) - pik, msum);
            const real c = (1.0 - tau) * bsum * dpi * P.s1 * mh[l] * (1.0 - mh[l]);
            ax += c * gh[l][0];
            ay += c * gh[l][1];
          }
          dbx[(k * M + i) * 2] = ax;
          dbx[(k * M + i) * 2 + 1] = ay;
        }
      }
    }
  }
};

// ------------------------------------------------------------------------------------------
// Input-rate costs (BranchMPCProx, MPC_branch.py:280-297) couple consecutive inputs.  The Riccati recursion carries the
// previous input as extra state: xi = (x, v), v_{k+1} = u_k, i.e. A~ = [A 0; 0 0], B~ = [B; I], C~ = [C; 0].  This
// wrapper presents a physical model M in that augmented form; everything that concerns the physical state (rollouts,
// collision, policies, linearisation data) is forwarded unchanged.
// ------------------------------------------------------------------------------------------
template <class M>
struct RateAug {
  static constexpr int NXP = M::NXP;
  static constexpr int NU = M::NU;
  static constexpr int NX = M::NXP + M::NU;
  static constexpr int NLIN = M::NLIN;
  static constexpr int NCC = M::NCC;
  static constexpr bool RATE = true;
  static constexpr bool kStateTransform = false;

  static constexpr int HEADING = M::HEADING;
  BMPC_D static void step(const KParams& P, const real* x, const real* u, real* xn) { M::step(P, x, u, xn); }
  BMPC_D static void step_sc(const KParams& P, const real* x, const real* u, real s, real c, real* xn) { M::step_sc(P, x, u, s, c, xn); }
  BMPC_D static void linearize_sc(const KParams& P, const real* x, const real* u, real s, real c, real* lin, real* cc, real* xn) {
    M::linearize_sc(P, x, u, s, c, lin, cc, xn);
  }
  BMPC_D static void lin_only(const KParams& P, const real* x, const real* u, real s, real c, real* lin, real* cc) { M::lin_only(P, x, u, s, c, lin, cc); }
  BMPC_D static void policy(const KParams& P, int kind, const real* par, const real* x, real* u) { M::policy(P, kind, par, x, u); }
  BMPC_D static void policy_fast(const KParams& P, int kind, const real* par, const real* x, real* u) { M::policy_fast(P, kind, par, x, u); }
  BMPC_D static void linearize(const KParams& P, const real* x, const real* u, real* lin, real* cc, real* xn) {
    M::linearize(P, x, u, lin, cc, xn);
  }
  BMPC_D static void mulA(const KParams& P, const real* lin, const real* x, real* y) {
    M::mulA(P, lin, x, y);
#pragma unroll
    for (int a = 0; a < NU; ++a) y[NXP + a] = 0.0;
  }
  BMPC_D static void mulAT(const KParams& P, const real* lin, const real* g, real* y) {
    M::mulAT(P, lin, g, y);
#pragma unroll
    for (int a = 0; a < NU; ++a) y[NXP + a] = 0.0;
  }
  BMPC_D static void addBu(const KParams& P, const real* lin, const real* u, real* y) {
    M::addBu(P, lin, u, y);
#pragma unroll
    for (int a = 0; a < NU; ++a) y[NXP + a] += u[a];
  }
  BMPC_D static void mulBT(const KParams& P, const real* lin, const real* g, real* r) {
    M::mulBT(P, lin, g, r);
#pragma unroll
    for (int a = 0; a < NU; ++a) r[a] += g[NXP + a];
  }
  BMPC_D static void addC(const real* cc, real* y) { M::addC(cc, y); }
  BMPC_D static void expandC(const real* cc, real* C) {
    M::expandC(cc, C);
#pragma unroll
    for (int a = 0; a < NU; ++a) C[NXP + a] = 0.0;
  }
  BMPC_D static void denseA(const KParams& P, const real* lin, real* A) { M::denseA(P, lin, A); }   // physical block only
  BMPC_D static void denseB(const KParams& P, const real* lin, real* B) {
    real Bp[NXP * NU];
    M::denseB(P, lin, Bp);
#pragma unroll
    for (int i = 0; i < NXP * NU; ++i) B[i] = Bp[i];
#pragma unroll
    for (int a = 0; a < NU; ++a)
#pragma unroll
      for (int b = 0; b < NU; ++b) B[(NXP + a) * NU + b] = (a == b) ? 1.0 : 0.0;
  }
  BMPC_D static void collision(const KParams& P, const real* x, const real* zxy, real& h, real& dhx, real& dhy) {
    M::collision(P, x, zxy, h, dhx, dhy);
  }
  template <class Emit>
  BMPC_D static real policy_safety(const KParams& P, int kind, const real* par, int kind0, const real* par0,
                                   const real* xe0, const real* z0, real* zlast, int nsteps, Emit emit) {
    return M::policy_safety(P, kind, par, kind0, par0, xe0, z0, zlast, nsteps, emit);
  }
  BMPC_D static real branch_weight(const KParams& P, real hi, real himax) { return M::branch_weight(P, hi, himax); }
  static constexpr int NSAFE = M::NSAFE;
  BMPC_D static void safety_terms(const KParams& P, const real* z, real xe0, real xe1, real* v) { M::safety_terms(P, z, xe0, xe1, v); }
  static constexpr bool kWeightNeedsMax = M::kWeightNeedsMax;
};
