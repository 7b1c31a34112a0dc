// Closed-loop environment step on the device (row f1 of SURVEY.md 8f): everything the reference does between two
// controller calls, one thread per episode, so that consecutive MPC steps never leave the GPU.
//   highway  : Highway_env_branch.py:83-184 (Highway_env.step) + the collision flag of Highway_sim (:421-429)
//   quadruped: quadruped_env.py:67-130 (Quad_env.step)
//   merge    : Highway_env_branch.py:324-380 (Highway_env_merge.step): lane id of the ego, ramp coordinates (state transform S,
//              reference and bounds from the ramp's lookup tables at the ego's x), the obstacle follows its first policy
// `pre` runs before the solve (obstacle arg-max policy, lane bookkeeping, lane-change target, xRef rule); `post` applies
// both Euler plants.  The reference's quirks are kept (see oracle/env.py for the list and the pinning fixtures):
// rollouts use the symbolic policy branches and the lane-change target from BEFORE this step's update; the collision
// value is the numeric veh_col (clipped to +-5) in a hard min with the EGO rollout's lane value; the obstacle is driven by
// the numeric branch of the environment's ORIGINAL policy list (the handle's policy table, not the per-episode one).
#pragma once
#include "bmpc_models.h"

struct EnvArgs {
  real* x;            // [count][n] ego state (in/out over pre+post)
  real* z;            // [count][n] obstacle state
  int* lane;          // [count][2] highway: lane index of ego / obstacle
  real* polpar;       // [count][m][4] per-episode policy parameters seen by the controller (lane-change row rewritten)
  const real* goal;   // [count][n] quadruped: desired final state
  int* obs_policy;    // [count] out
  int* collided;      // [count] in/out, sticky
  real* xref;         // [count][n] out
  real* u_obs;        // [count][d] out
  int count, t, n_lane;
};

BMPC_D real env_softmin2(real a, real b, real g) {
  const real mn = fmin(a, b);
  const real ea = bmpc_exp(-g * (a - mn)), eb = bmpc_exp(-g * (b - mn));
  return (ea * a + eb * b) / (ea + eb);
}

// numeric branch of the highway policies (highway_branch_dyn.py:54-148): brake is softmax([-5, -v], 3) there
BMPC_D void env_highway_policy_numeric(const KParams& P, int kind, const real* par, const real* s, real* u) {
  if (kind == BMPC_POLICY_BRAKE) {
    const real a = -5.0, b = -s[2];
    const real mx = fmax(a, b);
    const real ea = bmpc_exp(3.0 * (a - mx)), eb = bmpc_exp(3.0 * (b - mx));
    u[0] = (ea * a + eb * b) / (ea + eb);
    u[1] = -P.Kpsi * s[3];
  } else {
    HighwayModel::policy(P, kind, par, s, u);
  }
}

__global__ void bmpc_env_pre_highway(const __grid_constant__ KParams P, const EnvArgs a) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.count) return;
  using M = HighwayModel;
  real x[4], z[4];
  for (int q = 0; q < 4; ++q) { x[q] = a.x[(size_t)i * 4 + q]; z[q] = a.z[(size_t)i * 4 + q]; }
  // Highway_sim's collision flag, evaluated before the step (vehicle length 4, width 2.4)
  const real dis = fmax(fabs(x[0] - z[0]) - 4.0, fabs(x[1] - z[1]) - 2.4);
  if (dis < 0.0) a.collided[i] = 1;
  // obstacle policy: arg-max over j of min_t( veh_col_numeric(ego rollout, obstacle rollout_j), lane value of the ego rollout )
  const real lb = P.veh_W / 2.0, ub = a.n_lane * 3.6 - P.veh_W / 2.0;
  const real* pp = a.polpar + (size_t)i * P.m * 4;
  real zr[BMPC_MAX_POLICIES][4], hcol[BMPC_MAX_POLICIES];
  for (int j = 0; j < P.m; ++j) {
    for (int q = 0; q < 4; ++q) zr[j][q] = z[q];
    hcol[j] = 1e300;
  }
  real xe[4] = {x[0], x[1], x[2], x[3]};
  real hlane = 1e300;
  for (int t = 0; t < P.N; ++t) {
    real u[2], xn[4];
    M::policy(P, P.pol_kind[0], pp, xe, u);
    M::step(P, xe, u, xn);
    for (int q = 0; q < 4; ++q) xe[q] = xn[q];
    hlane = fmin(hlane, env_softmin2(xe[1] - lb, ub - xe[1], 5.0));
    for (int j = 0; j < P.m; ++j) {
      M::policy(P, P.pol_kind[j], pp + 4 * j, zr[j], u);
      M::step(P, zr[j], u, xn);
      for (int q = 0; q < 4; ++q) zr[j][q] = xn[q];
      const real dx = bmpc_clamp(fabs(xe[0] - xn[0]) - (P.veh_L + 1.0), -5.0, 5.0);
      const real dy = bmpc_clamp(fabs(xe[1] - xn[1]) - (P.veh_W + 0.2), -5.0, 5.0);
      real h, gx, gy;
      soft_box(dx, dy, h, gx, gy);
      hcol[j] = fmin(hcol[j], h);
    }
  }
  int best = 0;
  real hb = fmin(hcol[0], hlane);
  for (int j = 1; j < P.m; ++j) {
    const real hj = fmin(hcol[j], hlane);
    if (hj > hb) { hb = hj; best = j; }   // np.argmax: first maximum
  }
  a.obs_policy[i] = best;
  // lane bookkeeping and the lane-change target of the controller's model (:101-118), after the rollouts above
  int le = a.lane[2 * i], lo = a.lane[2 * i + 1];
  {
    const int nl = (int)rint((x[1] - 1.8) / 3.6);
    if (a.t == 0 || (nl != le && fabs(x[1] - 1.8 - 3.6 * nl) < 1.4)) le = nl;
    const int no = (int)rint((z[1] - 1.8) / 3.6);
    if (a.t == 0 || (no != lo && fabs(z[1] - 1.8 - 3.6 * no) < 1.4)) {
      lo = no;
      const int tgt = (le < lo) ? lo - 1 : (le > lo) ? lo + 1 : (lo > 0 ? lo - 1 : lo + 1);
      for (int j = 0; j < P.m; ++j)
        if (P.pol_kind[j] == BMPC_POLICY_LC) {
          real* row = a.polpar + ((size_t)i * P.m + j) * 4;
          row[0] = 0.0; row[1] = 1.8 + 3.6 * tgt; row[2] = 20.0; row[3] = 0.0;
        }
    }
    a.lane[2 * i] = le;
    a.lane[2 * i + 1] = lo;
  }
  // the input the obstacle really gets: numeric branch, the environment's original policy list (:60, :149)
  real uo[2];
  env_highway_policy_numeric(P, P.pol_kind[best], P.pol_par[best], z, uo);
  a.u_obs[(size_t)i * 2] = uo[0];
  a.u_obs[(size_t)i * 2 + 1] = uo[1];
  // xRef rule (:153-167)
  const real Ydes = (x[0] < z[0]) ? 1.8 + le * 3.6 : z[1];
  const real vdes = (fabs(x[1] - Ydes) < 1.0 && x[0] > z[0] + 3.0) ? 20.0 : z[2] + (z[0] + 1.5 - x[0]);
  real* r = a.xref + (size_t)i * 4;
  r[0] = 0.0; r[1] = Ydes; r[2] = vdes; r[3] = 0.0;
}

__global__ void bmpc_env_pre_quadruped(const __grid_constant__ KParams P, const EnvArgs a, real L1, real L2, real col_tol) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.count) return;
  using M = QuadrupedModel;
  real x[3], z[3];
  for (int q = 0; q < 3; ++q) { x[q] = a.x[(size_t)i * 3 + q]; z[q] = a.z[(size_t)i * 3 + q]; }
  const real* pp = a.polpar ? a.polpar + (size_t)i * P.m * 4 : nullptr;
  real zr[BMPC_MAX_POLICIES][3], hcol[BMPC_MAX_POLICIES];
  for (int j = 0; j < P.m; ++j) {
    for (int q = 0; q < 3; ++q) zr[j][q] = z[q];
    hcol[j] = 1e300;
  }
  real xe[3] = {x[0], x[1], x[2]};
  for (int t = 0; t < P.N; ++t) {
    real u[3], xn[3];
    M::policy(P, P.pol_kind[0], pp ? pp : P.pol_par[0], xe, u);
    M::step(P, xe, u, xn);
    for (int q = 0; q < 3; ++q) xe[q] = xn[q];
    for (int j = 0; j < P.m; ++j) {
      M::policy(P, P.pol_kind[j], pp ? pp + 4 * j : P.pol_par[j], zr[j], u);
      M::step(P, zr[j], u, xn);
      for (int q = 0; q < 3; ++q) zr[j][q] = xn[q];
      // robot_col, numeric branch: Euclidean distance (quadruped_branch_dyn.py:146-150)
      const real ex = xe[0] - xn[0], ey = xe[1] - xn[1];
      hcol[j] = fmin(hcol[j], sqrt(ex * ex + ey * ey) - (L1 + L2) / 2.0 - col_tol);
    }
  }
  int best = 0;
  if (!(hcol[0] > 0.5)) {                                  // quadruped_env.py:91-94
    real hb = hcol[0];
    for (int j = 1; j < P.m; ++j)
      if (hcol[j] > hb) { hb = hcol[j]; best = j; }
  }
  a.obs_policy[i] = best;
  real uo[3];
  M::policy(P, P.pol_kind[best], P.pol_par[best], z, uo);
  for (int q = 0; q < 3; ++q) a.u_obs[(size_t)i * 3 + q] = uo[q];
  // xRef: up to 5 m towards the goal, heading along that direction unwrapped about the goal heading (:100-114)
  const real* g = a.goal + (size_t)i * 3;
  real dx = g[0] - x[0], dy = g[1] - x[1];
  const real nd = sqrt(dx * dx + dy * dy);
  const real sc = fmin(nd, 5.0) / nd;
  dx *= sc;
  dy *= sc;
  real psi = x[2];
  if (sqrt(dx * dx + dy * dy) > 0.1) {
    psi = atan2(dy, dx);
    while (psi - g[2] > 3.141592653589793) psi -= 2.0 * 3.141592653589793;
    while (psi - g[2] < -3.141592653589793) psi += 2.0 * 3.141592653589793;
  }
  real* r = a.xref + (size_t)i * 3;
  r[0] = x[0] + dx; r[1] = x[1] + dy; r[2] = psi;
}

// both Euler plants (Highway_env_branch.py:39-41, quadruped_env.py:34-40): ego with the controller's first input
template <class M>
__global__ void bmpc_env_post(const __grid_constant__ KParams P, real* x, real* z, const real* u0, const real* u_obs, int count) {
  constexpr int NX = M::NX, NU = M::NU;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  real s[NX], u[NU], sn[NX];
  for (int q = 0; q < NX; ++q) s[q] = x[(size_t)i * NX + q];
  for (int q = 0; q < NU; ++q) u[q] = u0[(size_t)i * NU + q];
  M::step(P, s, u, sn);
  for (int q = 0; q < NX; ++q) x[(size_t)i * NX + q] = sn[q];
  for (int q = 0; q < NX; ++q) s[q] = z[(size_t)i * NX + q];
  for (int q = 0; q < NU; ++q) u[q] = u_obs[(size_t)i * NU + q];
  M::step(P, s, u, sn);
  for (int q = 0; q < NX; ++q) z[(size_t)i * NX + q] = sn[q];
}

// ---- merge scenario ------------------------------------------------------------------------------------------------
struct MergeEnvArgs {
  real* x;            // [count][4] ego (in/out over pre+post)
  real* z;            // [count][4] obstacle
  int* lane_id;       // [count] ego: 1 = on the ramp, 0 = on the highway (sticky once x > merge_s + 8, :329-330)
  int* collided;      // [count] in/out, sticky (Highway_sim :421-429)
  real* xref;         // [count][4] out
  real* S;            // [count][4][4] out
  real* bounds;       // [count][2][2] out: (lo, hi) of the y row and of the psi row
  real* u_obs;        // [count][2] out
  const real* tab_x;  // ramp centre line: grid, y and heading (merge_geometry :227-262), n points
  const real* tab_y;
  const real* tab_psi;
  int tab_n;
  int count, n_lane, merge_lane;
  real merge_s, v0, psimax;
};

BMPC_D real env_table(const real* gx, const real* gv, int n, real x) {
  int lo = 0, hi = n - 1;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (gx[mid] <= x) lo = mid; else hi = mid;
  }
  return gv[lo] + (gv[lo + 1] - gv[lo]) / (gx[lo + 1] - gx[lo]) * (x - gx[lo]);
}

__global__ void bmpc_env_pre_merge(const __grid_constant__ KParams P, const MergeEnvArgs a) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.count) return;
  real x[4], z[4];
  for (int q = 0; q < 4; ++q) { x[q] = a.x[(size_t)i * 4 + q]; z[q] = a.z[(size_t)i * 4 + q]; }
  const real dis = fmax(fabs(x[0] - z[0]) - 4.0, fabs(x[1] - z[1]) - 2.4);
  if (dis < 0.0) a.collided[i] = 1;
  int lane = a.lane_id[i];
  if (x[0] > a.merge_s + 8.0) lane = 0;
  a.lane_id[i] = lane;
  real* S = a.S + (size_t)i * 16;
  real* r = a.xref + (size_t)i * 4;
  real* bd = a.bounds + (size_t)i * 4;
  for (int q = 0; q < 16; ++q) S[q] = (q % 5 == 0) ? 1.0 : 0.0;
  if (lane == 0) {
    r[0] = 0.0; r[1] = (a.n_lane - 0.5) * 3.6; r[2] = a.v0; r[3] = 0.0;          // :351-353
    bd[0] = P.rlo[0]; bd[1] = P.rhi[0]; bd[2] = P.rlo[1]; bd[3] = P.rhi[1];        // mpc.param.bx
  } else {
    const real y0 = env_table(a.tab_x, a.tab_y, a.tab_n, x[0]), psi0 = env_table(a.tab_x, a.tab_psi, a.tab_n, x[0]);
    const real t = tan(psi0);
    S[4] = -t;                                                                     // :357
    r[0] = 0.0; r[1] = -t * x[0] + y0 + 1.8; r[2] = a.v0; r[3] = psi0;             // :358
    bd[0] = -(t * x[0] - y0 - P.veh_W / 2.0);                                      // -bx[1]
    bd[1] = -t * x[0] + y0 + 3.6 * a.merge_lane - P.veh_W / 2.0;                   //  bx[0]
    bd[2] = -(-psi0 + a.psimax);                                                   // -bx[3]
    bd[3] = psi0 + a.psimax;                                                       //  bx[2]
  }
  // every vehicle but the ego follows its first policy (:347-348); the obstacle drives on the highway (numeric branch)
  real uo[2];
  env_highway_policy_numeric(P, P.pol_kind[0], P.pol_par[0], z, uo);
  a.u_obs[(size_t)i * 2] = uo[0];
  a.u_obs[(size_t)i * 2 + 1] = uo[1];
}
