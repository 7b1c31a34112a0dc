// Kernel parameter block of libbranchmpc (passed by value; fits the 4 KB kernel-parameter space) and the
// closed-form tree numbering of MPC_branch.py:928-981.  See include/branchmpc.h for the ABI.
#pragma once
#include "bmpc_portable.h"

#include "branchmpc.h"

struct KParams {
  // ---- tree topology (BFS numbering of MPC_branch.py:928-981; all closed form) ----
  int m, NB, N;
  int nbranch, totalu, totalx;
  int nup;                      // padded node count = totalu + nbranch (one pad slot per branch)
  int off[BMPC_MAX_NB + 2];     // first branch id of each depth; off[NB+1] = nbranch
  int pw[BMPC_MAX_NB + 1];      // m^depth
  // obstacle scenario tree (equal to the ego tree for the branch controllers; robustMPC plans one ego chain against it)
  int zm, zNB, zN, znbranch;
  int zoff[BMPC_MAX_NB + 2], zpw[BMPC_MAX_NB + 1];
  int nbx;                      // branches the per-branch arrays of the slab are sized for = max(nbranch, znbranch)
  int pub_totalu, pub_totalx;   // rows of uPred / xPred the caller sees (the robust chain carries one internal dummy stage)
  // ---- model ----
  real dt, veh_L, veh_W, Kpsi, s1, lane_lo, lane_hi, quad_margin;
  int pol_kind[BMPC_MAX_POLICIES];
  real pol_par[BMPC_MAX_POLICIES][4];
  const real* lut_x;            // lookup table psiref(x) of the *_REF policies: grid (strictly increasing) and values, or null
  const real* lut_y;
  int lut_n;
  // ---- cost / constraints ----
  int ctrl;
  real Q[BMPC_MAX_N * BMPC_MAX_N], Qf[BMPC_MAX_N * BMPC_MAX_N], R[BMPC_MAX_D * BMPC_MAX_D], dR[BMPC_MAX_D];
  real dq_scale;                // dQ = dq_scale * Q  (0.5 BranchMPC :1070, 3 BranchMPCProx :271)
  real lam_lin;                 // Qslack[1]
  int nrows;                    // two-sided state rows (collision row not counted)
  real rf[BMPC_MAX_ROWS][BMPC_MAX_N], rlo[BMPC_MAX_ROWS], rhi[BMPC_MAX_ROWS];
  int rf_one[BMPC_MAX_ROWS];    // index of the row's only non-zero entry, or -1 for a general row
  real ulo[BMPC_MAX_D], uhi[BMPC_MAX_D];
  // ---- solver ----
  int warm_on_refresh;          // 1: the warm polish is also tried on the solves that refresh rho
  int warm_backoff;             // 1: that skip doubles with every further failed attempt in a row (at most 24 solves)
  int warm_skip;                // solves without a warm-polish attempt after one that ended on the ADMM path
  int max_iter, polish_first, polish_every, polish_passes, polish_al_iters, polish_careful, warm_polish, rebalance, rho_refresh, warm_passes, check_every, polish_stable, polish_force;
  real alpha, theta, theta_u, eps_abs, polish_big, polish_mult, rho_u_feedback;
  float inv_N, inv_m;           // reciprocals, rounded up, for exact small-integer division (bmpc_idiv)
  real inv_theta, inv_theta_u;  // reciprocals (the polish passes scale by them per row)
  int cycles_mode;              // 0: `cycles` = whole solve; k > 0: time spent in phase k (see Solver::prof_begin)
  int ipm_after, ipm_max_iter;  // interior-point fallback: after this many failed polish attempts (0 = never), iteration cap
  real ipm_mu_tol, ipm_s0, ipm_y0;
  // ---- BranchMPC_CVaR: cutting-plane loop over the risk multipliers ----
  real cvar_alpha;              // ralpha
  real cvar_floor;              // smallest branch weight handed to the inner tree QP (a cone without risk weight)
  real cvar_tol;                // relative gap between the cutting-plane model and the value at the current multipliers
  int cvar_max_cuts;            // cap on inner solves per MPC step
  int cv_rows, cv_cols;         // master LP tableau (rows incl. objective, columns incl. right-hand side)
  // ---- belief-state MPC (PredictiveControllers.MPC): chain of N stages, state augmented with M x m beliefs ----
  int hmm_M, xb_cols;           // agents; columns of one xbackup row
  real hmm_col_alpha, hmm_tran_diag, hmm_thres;
  const real* b0;               // [count][M][m]
  const real* xbackup;          // [count][M*m][xb_cols]
  real* bel;                    // per-team scratch: linearisation trajectory of the augmented state
  size_t bel_reals;
  real* nu_cache;               // persistent [cap][nbranch]: multipliers of the last solved step (warm start)
  real* cv;                     // per-team scratch of the risk master problem
  size_t cv_reals;
  // ---- batch ----
  int count;
  const real* x0;
  const real* z0;
  const real* xref;
  const real* polpar;           // [count][m][4] or null
  const real* xform;            // [count][n][n] state transform S of this call, or null (merge scenario, MPC_branch.py:2043)
  const real* xbounds;          // [count][nrows][2] (lo, hi) of the state rows of this call, or null
  real* uLin;                   // persistent [cap][totalu+1][d]
  int* pbest;                   // persistent [cap][nbranch]
  real* oldin;                  // persistent [cap][d]
  real* xprev;                  // persistent [cap][pub_totalx][n]: previous predicted states (robustMPC's LTV shift), else null
  int* started;                 // persistent [cap]
  real* rho_cache;              // solver cache [cap][totalu][rows+inputs]: curvature-matched rho of the last refresh
  long long* code_cache;              // solver cache [cap][totalu]: active-set codes of the last certified optimum
  int* cache_state;             // solver cache [cap][2]: age of rho_cache (-1 none), code_cache valid
  bmpc_outputs out;
  int* cost;                    // solver cache [cap]: cycles >> 10 of the previous solve (scheduling hint), may be null
  const int* order;             // work order of this launch (longest expected first), or null = natural order
  int stage_on;                 // 1: the next episode's uLin / active-set codes / rho cache are staged into shared memory by bulk
                                //    copies while the current episode is being solved (shared slab placement, tree controllers)
  int* counter;                 // work queue head
  real* gws;                    // global workspace (only when the per-problem slab does not fit shared memory)
  size_t slab_reals;            // reals per problem slab
  size_t factor_reals;          // reals of the per-warp factor-field region (split placement)
  real* ipm;                    // per-warp scratch of the interior-point fallback
  size_t ipm_reals;
};

// floor(q / d) for 0 <= q < 2^16 through a single-precision multiply with the rounded-up reciprocal `inv` of d (exact:
// tests/test_host_logic.py checks every q < 65536 for every d the ABI admits); the integer division instruction sequence is
// ~25 instructions and sat in every node-parallel loop
// one ego chain instead of a tree: robustMPC and the belief-state MPC
BMPC_HD inline bool bmpc_is_chain(int ctrl) { return ctrl == BMPC_CTRL_ROBUST || ctrl == BMPC_CTRL_BELIEF; }
BMPC_HD inline int bmpc_idiv(int q, float inv) { return (int)((float)q * inv); }
BMPC_HD inline int bmpc_ndu(const KParams& P, int b) { return b == 0 ? 0 : 1 + P.N * (b - 1); }
BMPC_HD inline int bmpc_ndx(const KParams& P, int b) {
  const int ol = P.off[P.NB];   // first leaf branch
  return b < ol ? bmpc_ndu(P, b) : 1 + P.N * (ol - 1) + (P.N + 1) * (b - ol);
}
BMPC_HD inline int bmpc_depth(const KParams& P, int b) {
  int d = 0;
  while (b >= P.off[d + 1]) ++d;
  return d;
}
BMPC_HD inline int bmpc_parent(const KParams& P, int b, int d) { return P.off[d - 1] + bmpc_idiv(b - P.off[d], P.inv_m); }
BMPC_HD inline int bmpc_first_child(const KParams& P, int b, int d) { return P.off[d + 1] + (b - P.off[d]) * P.m; }
