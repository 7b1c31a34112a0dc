// Host-side translation of a bmpc_config into the kernel parameter block: validation, tree numbering
// (MPC_branch.py:928-981 in closed form) and solver defaults.  Shared by the CUDA library (bmpc_api.cu)
// and the single-lane host build used by the CPU tests (tests/hostsim).
#pragma once
#include <math.h>
#include <stdio.h>
#include <string.h>

#include <string>

#include "bmpc_params.h"

namespace bmpc {

inline int model_dims(int model, int* n, int* d) {
  if (model == BMPC_MODEL_HIGHWAY || model == BMPC_MODEL_MERGE) { *n = 4; *d = 2; return 0; }
  if (model == BMPC_MODEL_QUADRUPED) { *n = 3; *d = 3; return 0; }
  return -1;
}

// Fills every field of KParams that does not depend on a particular call.  Returns BMPC_OK or an error code with
// a message in `err`.
inline int make_params(const bmpc_config& c, KParams* out, std::string* err) {
  KParams P;
  memset(&P, 0, sizeof(P));
  int n = 0, d = 0;
  if (model_dims(c.model, &n, &d) != 0) { *err = "unknown model kind"; return BMPC_E_INVALID; }
  if (c.n != n || c.d != d) { *err = "n/d do not match the model kind"; return BMPC_E_INVALID; }
  if (c.N < 2 || c.N > 64) { *err = "N must be in [2, 64]"; return BMPC_E_INVALID; }
  if (c.NB < 1 || c.NB > BMPC_MAX_NB) { *err = "NB must be in [1, BMPC_MAX_NB]"; return BMPC_E_INVALID; }
  if (c.m < 1 || c.m > BMPC_MAX_POLICIES) { *err = "m must be in [1, BMPC_MAX_POLICIES]"; return BMPC_E_INVALID; }
  if (!(c.dt > 0.0)) { *err = "dt must be positive"; return BMPC_E_INVALID; }
  if (c.n_rows < 0 || c.n_rows > BMPC_MAX_ROWS) { *err = "n_rows out of range"; return BMPC_E_INVALID; }
  if (c.controller != BMPC_CTRL_BRANCH && c.controller != BMPC_CTRL_PROX && c.controller != BMPC_CTRL_ROBUST &&
      c.controller != BMPC_CTRL_CVAR && c.controller != BMPC_CTRL_BELIEF) {
    *err = "unknown controller kind";
    return BMPC_E_INVALID;
  }
  if (c.controller == BMPC_CTRL_CVAR) {
    if (c.model != BMPC_MODEL_HIGHWAY && c.model != BMPC_MODEL_MERGE) { *err = "BranchMPC_CVaR is built for the highway models"; return BMPC_E_UNSUPPORTED; }
    if (!(c.cvar_alpha > 0.0 && c.cvar_alpha <= 1.0)) { *err = "cvar_alpha (ralpha) must be in (0, 1]"; return BMPC_E_INVALID; }
    for (int a = 0; a < d; ++a)
      if (c.dR[a] != 0.0) { *err = "BranchMPC_CVaR ignores input-rate costs; dR must be 0"; return BMPC_E_UNSUPPORTED; }
  }
  if (c.model == BMPC_MODEL_MERGE && (c.controller != BMPC_CTRL_CVAR || c.n_rows != 2)) {
    *err = "the merge model is built for BranchMPC_CVaR with the reference's two state rows (main_branch.py:87)";
    return BMPC_E_UNSUPPORTED;
  }
  const bool belief = c.controller == BMPC_CTRL_BELIEF;
  if (belief) {
    if (c.model != BMPC_MODEL_HIGHWAY) { *err = "the belief-state MPC is built for the highway model"; return BMPC_E_UNSUPPORTED; }
    if (c.hmm_M < 1 || c.hmm_M * c.m > 9) { *err = "hmm_M * m must be in [1, 9]"; return BMPC_E_INVALID; }
    if (c.NB != 1) { *err = "the belief-state MPC plans one chain: NB must be 1"; return BMPC_E_INVALID; }
    if (!(c.hmm_tran_diag >= 0.0 && c.hmm_tran_diag <= 1.0) || !(c.hmm_col_alpha > 0.0)) { *err = "bad belief-model constants"; return BMPC_E_INVALID; }
    for (int a = 0; a < d; ++a)
      if (c.dR[a] != 0.0) { *err = "belief-state MPC with input-rate costs (dR != 0) is not built"; return BMPC_E_UNSUPPORTED; }
  }
  const bool robust = c.controller == BMPC_CTRL_ROBUST || belief;   // one ego chain with a dummy stage for the terminal state
  if (c.controller == BMPC_CTRL_ROBUST) {
    for (int a = 0; a < d; ++a)
      if (c.dR[a] != 0.0) { *err = "robustMPC with input-rate costs (dR != 0) is not built"; return BMPC_E_UNSUPPORTED; }
  }
  if (c.Qslack[0] != 0.0) { *err = "quadratic slack weight Qslack[0] must be 0 (the reference uses 0)"; return BMPC_E_UNSUPPORTED; }
  if (!(c.Qslack[1] > 0.0)) { *err = "linear slack weight Qslack[1] must be positive"; return BMPC_E_INVALID; }
  for (int i = 0; i < c.m; ++i) {
    const int k = c.policy_kind[i];
    const bool hw = (k == BMPC_POLICY_MAINTAIN || k == BMPC_POLICY_BRAKE || k == BMPC_POLICY_LC || k == BMPC_POLICY_TRACKV ||
                     k == BMPC_POLICY_TRACKV_REF || k == BMPC_POLICY_BRAKE_REF);
    const bool qd = (k == BMPC_POLICY_FORWARD || k == BMPC_POLICY_STOP);
    if ((c.model != BMPC_MODEL_QUADRUPED && !hw) || (c.model == BMPC_MODEL_QUADRUPED && !qd)) {
      *err = "policy kind does not belong to the model";
      return BMPC_E_INVALID;
    }
  }
  for (int a = 0; a < d; ++a) {
    if (!(c.u_lo[a] <= c.u_hi[a])) { *err = "empty input box"; return BMPC_E_INVALID; }
    if (!(c.R[a * d + a] > 0.0)) { *err = "R must have a positive diagonal"; return BMPC_E_INVALID; }
  }
  // obstacle scenario tree
  P.zm = c.m;
  P.zNB = c.NB;
  P.zN = c.N;
  {
    int pw = 1, off = 0;
    for (int k = 0; k <= c.NB; ++k) {
      P.zpw[k] = pw;
      P.zoff[k] = off;
      off += pw;
      pw *= c.m;
    }
    P.zoff[c.NB + 1] = off;
    P.znbranch = off;
  }
  // ego tree: the scenario tree itself, or (robustMPC, MPC_branch.py:1301-1302) one chain of N*NB+1 input nodes plus an
  // internal dummy stage that carries the terminal state's cost and soft rows
  P.m = robust ? 1 : c.m;
  P.NB = robust ? 1 : c.NB;
  P.N = belief ? c.N : (robust ? c.N * c.NB + 1 : c.N);   // belief chain: root + N nodes = N + 1 states (PredictiveControllers.py:199)
  int pw = 1, off = 0;
  for (int k = 0; k <= P.NB; ++k) {
    P.pw[k] = pw;
    P.off[k] = off;
    off += pw;
    pw *= P.m;
  }
  P.off[P.NB + 1] = off;
  P.nbranch = off;
  P.totalu = 1 + P.N * (P.nbranch - 1);
  P.totalx = robust ? P.totalu : P.totalu + P.pw[P.NB];
  P.pub_totalu = robust ? P.totalu - 1 : P.totalu;
  P.pub_totalx = P.totalx;
  P.nup = P.totalu + P.nbranch;
  P.nbx = P.nbranch > P.znbranch ? P.nbranch : P.znbranch;
  P.dt = c.dt;
  P.veh_L = c.veh_L;
  P.veh_W = c.veh_W;
  P.Kpsi = c.Kpsi;
  P.s1 = c.s1;
  P.lane_lo = c.lane_lo;
  P.lane_hi = c.lane_hi;
  P.quad_margin = c.quad_margin;
  for (int i = 0; i < BMPC_MAX_POLICIES; ++i) {
    P.pol_kind[i] = c.policy_kind[i];
    for (int k = 0; k < 4; ++k) P.pol_par[i][k] = c.policy_param[i][k];
  }
  P.ctrl = c.controller;
  for (int i = 0; i < n * n; ++i) { P.Q[i] = c.Q[i]; P.Qf[i] = c.Qf[i]; }
  for (int i = 0; i < d * d; ++i) P.R[i] = c.R[i];
  for (int i = 0; i < d; ++i) { P.dR[i] = c.dR[i]; P.ulo[i] = c.u_lo[i]; P.uhi[i] = c.u_hi[i]; }
  // MPC_branch.py:271 / :1070 / none in robustMPC and in BranchMPC_CVaR (its cones carry (x - xRef)'Q(x - xRef) only, :1953-1957)
  P.dq_scale = (c.controller == BMPC_CTRL_PROX) ? 3.0 : ((robust || c.controller == BMPC_CTRL_CVAR) ? 0.0 : 0.5);
  P.hmm_M = c.hmm_M;
  P.hmm_col_alpha = c.hmm_col_alpha;
  P.hmm_tran_diag = c.hmm_tran_diag;
  P.hmm_thres = c.hmm_thres;
  P.bel_reals = belief ? (size_t)(P.totalu + 2) * 16 : 0;
  P.cvar_alpha = c.cvar_alpha;
  P.cvar_floor = 1.0e-6;
  P.cvar_tol = 1.0e-8;
  P.cvar_max_cuts = 32;
  {
    // master LP of the risk multipliers: variables nu_c (c = 1..nbranch-1) and t; rows: one per non-leaf branch, the
    // bdim + m - 1 coupling rows of the dual-CVaR equalities (with the reference's index rule) and the cuts
    const int bdim = P.off[P.NB];
    const int nvar = P.nbranch;                              // nbranch - 1 multipliers + t
    const int nrow = bdim + (bdim + P.m - 1) + P.cvar_max_cuts;
    P.cv_rows = nrow + 1;
    P.cv_cols = nvar + nrow + 1;
    P.cv_reals = (size_t)2 * P.nbranch + (size_t)bdim * P.m + (size_t)P.cvar_max_cuts * P.nbranch +
                 (size_t)P.cv_rows * P.cv_cols + (size_t)(P.cv_rows + 3) / 2 + 2;   // multipliers, p, cuts, tableau, basis (ints)
  }
  P.lam_lin = c.Qslack[1];
  P.nrows = c.n_rows;
  for (int j = 0; j < c.n_rows; ++j) {
    for (int i = 0; i < n; ++i) P.rf[j][i] = c.row_f[j][i];
    int nz = 0, one = -1;
    for (int i = 0; i < n; ++i)
      if (c.row_f[j][i] != 0.0) { ++nz; one = i; }
    P.rf_one[j] = (nz == 1) ? one : -1;
    P.rlo[j] = isfinite(c.row_lo[j]) ? c.row_lo[j] : -1.0e300;
    P.rhi[j] = isfinite(c.row_hi[j]) ? c.row_hi[j] : 1.0e300;
    if (!(P.rlo[j] <= P.rhi[j])) { *err = "empty state row range"; return BMPC_E_INVALID; }
  }
  P.max_iter = c.max_iter > 0 ? c.max_iter : 400;
  P.polish_first = c.polish_first > 0 ? c.polish_first : 10;
  P.polish_every = c.polish_every > 0 ? c.polish_every : 10;
  P.polish_passes = c.polish_passes > 0 ? c.polish_passes : 6;
  P.polish_al_iters = c.polish_al_iters > 0 ? c.polish_al_iters : 24;
  P.polish_careful = c.polish_careful > 0 ? c.polish_careful : 0;   // off: the interior-point fallback handles cycling sets
  P.warm_polish = c.warm_polish >= 0 ? 1 : 0;
  P.warm_passes = c.warm_polish > 0 ? c.warm_polish : (c.model == BMPC_MODEL_QUADRUPED ? 6 : 3);   // measured (profiles/r01_knob_matrix.md)
  P.rho_refresh = c.rho_refresh > 0 ? c.rho_refresh : (c.rho_refresh < 0 ? 0 : 16);   // 8 / 16 / 32: 5.06 / 5.01 / 4.97 ms per highway step (profiles/r02_staging_ab.md)
  P.alpha = c.alpha > 0.0 ? c.alpha : 1.6;
  P.theta = c.theta > 0.0 ? c.theta : 1.0;
  P.theta_u = c.theta_u > 0.0 ? c.theta_u : 1.0;
  P.inv_N = nextafterf(1.0f / (float)P.N, 2.0f);
  P.inv_m = nextafterf(1.0f / (float)P.m, 2.0f);
  P.inv_theta = 1.0 / P.theta;
  P.inv_theta_u = 1.0 / P.theta_u;
  P.eps_abs = c.eps_abs > 0.0 ? c.eps_abs : 1.0e-6;
  P.polish_big = c.polish_big > 0.0 ? c.polish_big : 1.0e4;
  P.polish_mult = c.polish_mult > 0.0 ? c.polish_mult : 1.0e5;   // measured: one KKT solve fewer per polish pass than 1e4
  P.check_every = 5;
  P.polish_stable = c.reserved[2];
  // measured on the B200 (profiles/r01_knob_matrix.md): the highway problems are cheapest when the ADMM runs until its
  // implied active set is stable and the interior point only rescues repeated failures; the long, narrow quadruped trees
  // (ADMM contracts slowly along 51 stages) when the polish is forced early and its first failure goes to the interior point
  const bool quad = c.model == BMPC_MODEL_QUADRUPED;
  P.polish_force = c.reserved[3] > 0 ? c.reserved[3] : (quad ? 20 : 80);
  P.rho_u_feedback = 1.0;
  P.ipm_after = c.reserved[4] > 0 ? c.reserved[4] : (c.reserved[4] < 0 ? 0 : (quad ? 1 : 3));
  P.ipm_max_iter = c.reserved[5] > 0 ? c.reserved[5] : 40;
  P.cycles_mode = c.reserved[7] > 0 ? c.reserved[7] : 0;
  P.ipm_mu_tol = 1.0e-10;
  P.ipm_s0 = 1.0;
  P.ipm_y0 = 0.5;
  P.ipm = nullptr;
  P.ipm_reals = 0;
  P.rebalance = (c.reserved[0] & 1) ? 0 : 1;   // experimental switches
  P.warm_on_refresh = (c.reserved[0] & 2) ? 0 : 1;
  P.warm_backoff = (c.reserved[0] & 4) ? 0 : 1;
  {
    const int ws = (c.reserved[0] >> 4) & 15;    // 0 = default, 15 = never skip
    // quadruped: a failed attempt does not predict the next one (its hard episodes are the interior-point ones): skipping
    // costs 4 % in the first steps of a closed loop and gains nothing later (profiles/r02_staging_ab.md)
    P.warm_skip = ws == 0 ? (quad ? 0 : 3) : (ws == 15 ? 0 : ws);
  }
  *out = P;
  return BMPC_OK;
}

// (model, number of soft rows incl. the collision row) -> which Solver instantiation runs
inline bool supported_instance(int model, int n_rows, int controller = BMPC_CTRL_BRANCH, int obstacle_leaves = 1) {
  // robustMPC is instantiated for the highway model with the reference's two state rows and up to 9 obstacle nodes per slot
  if (controller == BMPC_CTRL_ROBUST || controller == BMPC_CTRL_BELIEF)
    return model == BMPC_MODEL_HIGHWAY && n_rows == 2 && obstacle_leaves <= 9;
  if (model == BMPC_MODEL_HIGHWAY) return n_rows >= 0 && n_rows <= 2;
  if (model == BMPC_MODEL_QUADRUPED) return n_rows == 0;
  if (model == BMPC_MODEL_MERGE) return controller == BMPC_CTRL_CVAR && n_rows == 2;
  return false;
}

}  // namespace bmpc
