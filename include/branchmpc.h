/*
 * branchmpc.h - C ABI of libbranchmpc.so: batched Branch-MPC solves on one B200 (sm_100a).
 *
 * The reference (Gavinli-lgf/belief-planning) has no FFI layer: its boundary is the duck-typed
 * Python pair  PredictiveModel + BranchMPC.solve(x, z, xRef)  (MPC_branch.py:883, :1171;
 * highway_branch_dyn.py:264).  This library is what a ctypes binding under that interface calls
 * (see INTEGRATION.md).  One call solves `count` independent scenario-tree MPC problems.
 *
 * Conventions
 *  - every data pointer passed to bmpc_solve() is a DEVICE pointer owned by the caller
 *    (PyTorch allocates); the library borrows it for the duration of the call;
 *    bmpc_solve_host() is the same call on HOST buffers (copies included);
 *  - all floating point data is IEEE float64, row-major, batch index outermost;
 *  - functions return 0 on success or a negative BMPC_E_* code; nothing throws or aborts across
 *    the ABI; per-problem solver status is an OUTPUT, not an error;
 *  - a handle is bound to one device and is not thread-safe; work is enqueued on `stream`
 *    (a cudaStream_t passed as void*; NULL = legacy default stream) and the call returns
 *    without synchronising.
 */
#ifndef BRANCHMPC_H_
#define BRANCHMPC_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BMPC_VERSION 201

#define BMPC_MAX_N 4          /* state dimension supported by this build            */
#define BMPC_MAX_D 3          /* input dimension                                    */
#define BMPC_MAX_POLICIES 4   /* branching factor m                                 */
#define BMPC_MAX_ROWS 4       /* two-sided soft state rows (besides the collision row) */
#define BMPC_MAX_NB 3         /* branching depth                                    */

/* model kinds: which closed-form predictive model the kernels evaluate */
#define BMPC_MODEL_HIGHWAY 0   /* highway_branch_dyn.PredictiveModel   (n=4, d=2) */
#define BMPC_MODEL_QUADRUPED 1 /* quadruped_branch_dyn.PredictiveModel (n=3, d=3) */
#define BMPC_MODEL_MERGE 2     /* highway_branch_dyn.PredictiveModel_merge (:400-502; n=4, d=2): the highway vehicle with the
                                  merge scenario's safety value (vehicle distance only, size [L+1, W+0.2], :452-456); handles of
                                  this model take a per-episode state transform and state bounds (bmpc_solve_transformed);
                                  BMPC_CTRL_CVAR with n_rows = 2 (what main_branch.py:87 builds) */

/* controller kinds */
#define BMPC_CTRL_BRANCH 0 /* MPC_branch.BranchMPC (effective, second definition, :881) */
#define BMPC_CTRL_PROX 1   /* MPC_branch.BranchMPCProx (:82)                             */
#define BMPC_CTRL_ROBUST 2 /* MPC_branch.robustMPC (:1275): total_u = N*NB+1, total_x = N*NB+2; xLin/zPred outputs unused */
#define BMPC_CTRL_BELIEF 4 /* PredictiveControllers.MPC (:56-340) on HMM_backup_dyn.PredictiveModel (:177-276): one ego chain of N
                              stages whose state is augmented with the beliefs over the other agents' policies; solved through
                              bmpc_solve_belief (x0, b0, xbackup, xRef); highway model, m * hmm_M <= 9 */
#define BMPC_CTRL_CVAR 3   /* MPC_branch.BranchMPC_CVaR (:1598): nested-CVaR objective (the controller main_branch.py:48 builds);
                              highway model; solved as a cutting-plane loop over the risk multipliers of the cones whose
                              inner problems are branch-weighted tree QPs */

/* backup-policy kinds (symbolic branch of each reference policy) */
#define BMPC_POLICY_MAINTAIN 0 /* highway_branch_dyn.backup_maintain        :54  */
#define BMPC_POLICY_BRAKE 1    /* highway_branch_dyn.backup_brake           :108 */
#define BMPC_POLICY_LC 2       /* highway_branch_dyn.backup_lc              :136, param = target state */
#define BMPC_POLICY_TRACKV 3   /* highway_branch_dyn.backup_maintain_trackV :80,  param[0] = v0 */
#define BMPC_POLICY_FORWARD 4  /* quadruped_branch_dyn.backup_forward       :34,  param[0] = v0 */
#define BMPC_POLICY_STOP 5     /* quadruped_branch_dyn.backup_stop          :46  */
#define BMPC_POLICY_TRACKV_REF 6 /* backup_maintain_trackV(x, cons, v0, psiref) :89-96: steering psiref(x) - Kpsi psi with the
                                    handle's lookup table (bmpc_set_lookup_table), param[0] = v0 */
#define BMPC_POLICY_BRAKE_REF 7  /* backup_brake(x, cons, psiref) :122-131: softmax([-5, -v], 3), steering as above */

/* working-set placement (bmpc_config.slab_mode) */
#define BMPC_SLAB_AUTO 0
#define BMPC_SLAB_SHARED 1 /* everything in shared memory (falls back to GLOBAL when it does not fit) */
#define BMPC_SLAB_SPLIT 2  /* iterate fields in shared memory, factor fields in a per-warp L2-resident region */
#define BMPC_SLAB_GLOBAL 3 /* everything in global memory */

/* error codes */
#define BMPC_OK 0
#define BMPC_E_INVALID (-1)     /* bad argument / unsupported configuration        */
#define BMPC_E_CUDA (-2)        /* a CUDA runtime call failed (see bmpc_last_error) */
#define BMPC_E_CAPACITY (-3)    /* count exceeds the handle's batch capacity        */
#define BMPC_E_UNSUPPORTED (-4) /* valid in the reference, not built yet            */

/* per-problem solver status (bmpc_outputs.status) */
#define BMPC_STATUS_POLISHED 0  /* active set verified: exact optimum of the QP (to round-off)   */
#define BMPC_STATUS_CONVERGED 1 /* ADMM residuals below eps_abs, or the interior-point fallback converged (complementarity
                                   gap <= 1e-9 Qslack[1], ~1e-7 from the optimum); active set not verified by the polish */
#define BMPC_STATUS_MAXITER 2   /* iteration caps of ADMM and of the interior point reached; best iterate returned */
#define BMPC_STATUS_NUMERIC 3   /* non-finite data met; previous plan kept (reference: feasible=0) */

typedef struct bmpc_config {
  int32_t model;      /* BMPC_MODEL_*  */
  int32_t controller; /* BMPC_CTRL_*   */
  int32_t n, d;       /* state / input dimension (PredictiveModel.n, .d)               */
  int32_t N;          /* steps per branch      (PredictiveModel.N, main_branch.py:24)   */
  int32_t NB;         /* branching depth       (BranchMPCParams.NB, main_branch.py:30)  */
  int32_t m;          /* number of backup policies = len(backupcons)                    */
  double dt;
  int32_t policy_kind[BMPC_MAX_POLICIES];
  double policy_param[BMPC_MAX_POLICIES][4]; /* defaults; per-problem override in bmpc_solve */

  /* cost, BranchMPCParams (MPC_branch.py:27-54), row-major */
  double Q[BMPC_MAX_N * BMPC_MAX_N];
  double Qf[BMPC_MAX_N * BMPC_MAX_N];
  double R[BMPC_MAX_D * BMPC_MAX_D];
  double dR[BMPC_MAX_D];
  double Qslack[2]; /* [quadratic, linear] as the code uses them (:1105-1106); quadratic must be 0 */

  /* soft state rows lo <= f'x <= hi (Fx x <= bx with opposite rows paired), weight Qslack[1]*w */
  int32_t n_rows;
  double row_f[BMPC_MAX_ROWS][BMPC_MAX_N];
  double row_lo[BMPC_MAX_ROWS]; /* -inf allowed */
  double row_hi[BMPC_MAX_ROWS]; /* +inf allowed */
  /* hard input box (Fu u <= bu) */
  double u_lo[BMPC_MAX_D];
  double u_hi[BMPC_MAX_D];

  /* model constants: Branch_constants / Quad_constants (utils.py:25-59) */
  double veh_L, veh_W; /* highway: L, W                                               */
  double Kpsi;         /* highway: heading gain of maintain/brake                     */
  double s1;           /* branching-probability sharpness                             */
  double lane_lo, lane_hi; /* highway model lane boundary LB (highway_branch_dyn.py:279) */
  double quad_margin;  /* quadruped: (L1+L2)/2 + col_tol                              */

  /* solver knobs (0 = library default) */
  int32_t max_iter;        /* ADMM iteration cap (default 400)                                   */
  int32_t polish_first;    /* first polish attempt after this many iterations (10)               */
  int32_t polish_every;    /* then this many later, doubling after each failed attempt (10)                          */
  int32_t polish_passes;   /* active-set passes per polish attempt (6)                           */
  int32_t polish_al_iters; /* augmented-Lagrangian refinements per pass (24)                     */
  int32_t polish_careful;  /* extra one-change-at-a-time passes when the set iteration cycles (0 = off, the default) */
  int32_t warm_polish;     /* warm solves first try a polish from the previous optimum's shifted active set:
                              number of active-set passes for that attempt (3; quadruped 6; <0 = off)        */
  int32_t rho_refresh;     /* warm solves reuse the cached rho for this many steps (16; <0 = recompute every solve)  */
  double alpha;            /* over-relaxation (1.6)                                              */
  double theta, theta_u;   /* curvature-matched rho scale for state rows / inputs (1)            */
  double eps_abs;          /* ADMM residual tolerance for STATUS_CONVERGED (1e-6)                */
  double polish_big;       /* lower bound of the stiff penalty, times branch weight (1e4)        */
  double polish_mult;      /* stiff penalty = polish_mult x curvature-matched stiffness (1e5)    */
  double cvar_alpha;       /* BMPC_CTRL_CVAR: the `ralpha` of BranchMPC_CVaR (MPC_branch.py:1601, :1798); in (0, 1]        */
  /* BMPC_CTRL_BELIEF: constants of the belief-state model (HMM_backup_dyn.py:238-267; lane_lo / lane_hi hold ylb / yub, s1 the
     sharpness of softsat) and the belief threshold above which a policy's row is imposed (PredictiveControllers.py:76, :213) */
  int32_t hmm_M;           /* number of uncontrolled agents                                       */
  double hmm_col_alpha, hmm_tran_diag, hmm_thres;

  int32_t slab_mode;      /* BMPC_SLAB_*: where a problem's working set lives (0 = library picks) */
  int32_t batch_capacity; /* maximum number of episodes (persistent state slots)      */
  int32_t device;         /* CUDA device ordinal                                      */
  int32_t reserved[8];    /* experiment switches, 0 = default: [0] bit 0 = no residual balancing of rho, bit 1 = no warm-polish attempt on
                             the solves that refresh rho, bit 2 = no doubling of the skip below, bits 4..7 = k: an episode whose warm-polish attempt
                             ended on the ADMM path skips the attempt on its next k solves (0 = the default: 3, the quadruped model never skips; 15 = never skip; with
                             bit 2 clear k doubles with every further failed attempt in a row, up to 24); [1] cap of resident
                             warps per SM; [2] polish when at most this many nodes changed their implied set between
                             checks; [3] polish at the latest every this many ADMM iterations (80; quadruped 20); [4] interior-point
                             fallback after this many failed polish attempts (3; quadruped 1; <0 = never; 100 = always, without a polish attempt first); [5] its iteration cap (40);
                             [6] bit 0 = natural work order instead of longest-first, bit 1 = stage the next episode's state into shared memory by bulk copies while the current one is solved (off by default: measured slower), bit 2 = bmpc_solve_host* packs the results into a device block and copies it back in one DMA instead of letting the kernel write them into the pinned host block; [7] k > 0 = the cycles output counts phase k only
                             (1 interior point, 2 expansion, 3 rho, 4 factorisations, 5 sweeps, 6 polish passes,
                             7 adjoint, 8 ADMM rows, 9 final pass) */
} bmpc_config;

/* Output device pointers; any may be NULL to skip that output. */
typedef struct bmpc_outputs {
  double* u0;        /* [count][d]            first applied input  (mpc.uPred[0])           */
  double* uPred;     /* [count][totalu][d]    unpackSolution, MPC_branch.py:1226            */
  double* xPred;     /* [count][totalx][n]    unpackSolution, MPC_branch.py:1225            */
  double* xLin;      /* [count][totalu][n]    linearisation states per input node (xtraj)   */
  double* zPred;     /* [count][totalu][n]    obstacle prediction per input node (ztraj)    */
  double* branch_w;  /* [count][nbranch]      branch weights w                               */
  double* branch_p;  /* [count][nbranch][m]   child probabilities of non-leaf branches       */
  double* objective; /* [count]               QP objective (1/2 z'Pz + q'z, slacks eliminated) */
  int32_t* status;   /* [count]               BMPC_STATUS_*                                  */
  int32_t* iters;    /* [count]               ADMM iterations used                           */
  int32_t* nfact;    /* [count]               Riccati factorisations used                    */
  int32_t* nsolve;   /* [count]               KKT solves (backward+forward sweeps): ADMM + polish */
  int64_t* cycles;   /* [count]               SM clock cycles the owning warp spent on the problem */
  double* bPred;     /* [count][totalx][m*hmm_M]  BMPC_CTRL_BELIEF: belief part of the predicted augmented state (xPred[:, 4:]) */
} bmpc_outputs;

typedef struct bmpc_handle bmpc_handle;

int bmpc_version(void);

/* Create a solver for one configuration.  Allocates the persistent per-episode state
 * (warm-start inputs uLin, arg-max child per branch, OldInput) for batch_capacity episodes. */
int bmpc_create(const bmpc_config* cfg, bmpc_handle** out);
int bmpc_destroy(bmpc_handle* h);

/* Forget the persistent state of the given episode slots (NULL = all): their next solve is a
 * first solve (inittree: zero inputs, MPC_branch.py:932-957).  One kernel (one block per listed slot) on the stream of
 * the handle's last solve; episode_ids is a HOST array. */
int bmpc_reset(bmpc_handle* h, const int64_t* episode_ids, int64_t count);

/* Tree numbering tables (BFS order, MPC_branch.py:928-981).  Each array has bmpc_num_branches()
 * entries; NULL pointers are skipped.  Host pointers. */
int bmpc_num_branches(const bmpc_handle* h);
int bmpc_total_x(const bmpc_handle* h);
int bmpc_total_u(const bmpc_handle* h);
int bmpc_get_topology(const bmpc_handle* h, int32_t* ndx, int32_t* ndu, int32_t* depth, int32_t* parent);
/* rows of the persistent uLin array per episode (bmpc_get_state / bmpc_set_state): total_u + 1, or total_u + 2 for
 * robustMPC, whose chain carries one internal stage for the terminal state */
int bmpc_ulin_rows(const bmpc_handle* h);

/* One MPC step for episodes 0..count-1 (episode i uses persistent slot i).
 *   x0, z0, xref : [count][n] device float64   (solve(x, z, xRef), MPC_branch.py:1171)
 *   policy_params: [count][m][4] device float64 or NULL (per-episode lane-change target etc.;
 *                  the reference rebuilds its model for this, highway_branch_dyn.py:331)
 * The kernel's parameter block travels through one constant-memory symbol per device: the call uploads it stream-ordered
 * and orders itself behind the previous solve launch on that device (other handles, other streams), so solves of
 * different handles on one device do not overlap.
 * The call may be recorded into a CUDA graph (stream capture): the recorded launch keeps a private pinned copy of its
 * parameter block, and every pointer argument must stay valid and in place for the replays; a replayed graph must not run
 * concurrently with other solves on the same device.  bmpc_last_kernel_ms() does not see replayed launches. */
int bmpc_solve(bmpc_handle* h, const double* x0, const double* z0, const double* xref,
               const double* policy_params, int64_t count, const bmpc_outputs* out, void* stream);

/* BranchMPC_CVaR.solve(x, z, xRef, S, Fx=None, bx) (MPC_branch.py:2043-2059) - the call Highway_env_merge.step makes
 * (Highway_env_branch.py:364) - for episodes 0..count-1 of a BMPC_MODEL_MERGE handle.  As bmpc_solve, plus per episode
 *   S            : [count][n][n] device float64, row-major state transform: stage cost (S x)' Q (S x) - 2 xRef' Q x (:1938,
 *                  :1962 - the reference transforms the quadratic term only) and state rows  lo <= Fx S x <= hi  (:1899)
 *   state_bounds : [count][n_rows][2] device float64, (lo, hi) of the handle's state rows for this call (the reference's bx
 *                  holds them as [hi_0, -lo_0, hi_1, -lo_1])
 * Either may be NULL: identity / the bounds of the configuration. */
int bmpc_solve_transformed(bmpc_handle* h, const double* x0, const double* z0, const double* xref,
                           const double* policy_params, const double* S, const double* state_bounds, int64_t count,
                           const bmpc_outputs* out, void* stream);
/* The same call on HOST buffers with zero-copy result views (see bmpc_solve_host_views). */
int bmpc_solve_transformed_host_views(bmpc_handle* h, const double* x0, const double* z0, const double* xref,
                                      const double* policy_params, const double* S, const double* state_bounds,
                                      int64_t count, const bmpc_outputs* want, bmpc_outputs* views);

/* 1-D piecewise-linear lookup table psiref(x) of the merge scenario's ramp policies (casadi `interpolant('refpsi', 'linear',
 * [X], psi)`, main_branch.py:79; the end segments continue outside the grid).  HOST pointers, xs strictly increasing,
 * 2 <= n <= 4096; copied to the device, replaces the handle's previous table.  Required before the first solve / model
 * evaluation of a handle whose policy table holds a *_REF kind. */
int bmpc_set_lookup_table(bmpc_handle* h, const double* xs, const double* ys, int32_t n);

/* One step of the belief-state MPC (BMPC_CTRL_BELIEF) for episodes 0..count-1: PredictiveControllers.MPC.solve(x0, b0, xbackup,
 * xRef) (:130-160).  Device pointers, float64:
 *   x0 [count][4], b0 [count][hmm_M][m] (the array the reference caller passes), xref [count][4],
 *   xbackup [count][hmm_M*m][xbackup_cols]: row m*i+j = agent i under policy j, time-major, column block 4k..4k+3 = backup state
 *   at step k (Highway_env.py:135-142 hands over N+1 states; at least N are read).
 * Outputs: u0, uPred [count][N][2], xPred [count][N+1][4] (physical part), bPred [count][N+1][hmm_M*m], objective, status ... */
int bmpc_solve_belief(bmpc_handle* h, const double* x0, const double* b0, const double* xbackup, int32_t xbackup_cols,
                      const double* xref, int64_t count, const bmpc_outputs* out, void* stream);

/* Point-wise belief-state model (HMM_backup_dyn.PredictiveModel.regressionAndLinearization, :216-229) on a BMPC_CTRL_BELIEF
 * handle, device pointers: xb [count][4+nb] augmented states (belief column-major, nb = hmm_M*m), xbackup [count][nb][4] the
 * backup state of every (agent, policy) row, u [count][2] -> A [count][n][n], B [count][n][2], C [count][n], next state xbp
 * [count][n], and per row the safety value's linearisation: h0 [count][nb] (h - Jh x) and Jh [count][nb][2] (d h / d(x, y)). */
int bmpc_eval_belief(bmpc_handle* h, const double* xb, const double* xbackup, const double* u, int64_t count, double* A,
                     double* B, double* C, double* h0, double* Jh, double* xbp, void* stream);

/* Same step on HOST buffers: copies inputs to the device, solves, copies every non-NULL output
 * back and synchronises.  `out` holds HOST pointers here.  Inputs travel in ONE transfer through pinned staging owned by
 * the handle (runs on the stream of the handle's last solve); input arrays that already lie in page-locked memory
 * (cudaHostAlloc / cudaHostRegister / torch pin_memory) are sent from where they lie instead.  The kernel writes the requested
 * results straight into a pinned host block of the handle (no transfer after the launch), from which they are copied into
 * the caller's arrays. */
int bmpc_solve_host(bmpc_handle* h, const double* x0, const double* z0, const double* xref,
                    const double* policy_params, int64_t count, const bmpc_outputs* out);
/* Zero-copy variant (the call the drop-in BranchMPC.solve makes): `want` marks the requested outputs with non-NULL
 * members (values ignored); `views` receives HOST pointers into one of the handle's two pinned result blocks, [count] rows
 * each; the blocks alternate, so the views of a call stay intact during the NEXT bmpc_solve_host* call on this handle and
 * are overwritten by the one after it. */
int bmpc_solve_host_views(bmpc_handle* h, const double* x0, const double* z0, const double* xref,
                          const double* policy_params, int64_t count, const bmpc_outputs* want, bmpc_outputs* views);

/* Persistent state access for tests / checkpointing (device or host pointers, see `on_host`).
 *   uLin [count][bmpc_ulin_rows][d], pbest [count][nbranch] (int32), old_input [count][d], started [count] (int32),
 *   xprev [count][totalx][n]: robustMPC only (its linearisation is the previous predicted trajectory shifted,
 *   MPC_branch.py:1429-1431), NULL otherwise; restoring `started` into a robustMPC handle requires xprev.
 * Like bmpc_reset they are ordered behind the handle's last solve (they run on its stream and synchronise it). */
int bmpc_get_state(bmpc_handle* h, double* uLin, int32_t* pbest, double* old_input, int32_t* started, double* xprev,
                   int64_t count, int on_host);
int bmpc_set_state(bmpc_handle* h, const double* uLin, const int32_t* pbest, const double* old_input,
                   const int32_t* started, const double* xprev, int64_t count, int on_host);

/* Model functions evaluated on the device for a batch of points (parity tests of rows M1-M5):
 *   dyn_linearization: A [count][n][n], B [count][n][d], C [count][n], xp [count][n]
 *   zpred            : [count][N][m*n]
 *   branch_eval p    : [count][m]
 *   col_eval         : hlin [count], dh [count][n]
 * Device pointers; NULL outputs are skipped. */
int bmpc_eval_model(bmpc_handle* h, const double* x, const double* z, const double* u, const double* policy_params,
                    int64_t count, double* A, double* B, double* C, double* xp, double* zpred, double* p,
                    double* hlin, double* dh, void* stream);

/* ---- belief-state model (HMM_backup_dyn.py), rows H1/H2 of the hot-path table -------------------------------------
 * Device pointers, float64; no handle needed (no persistent state).  Policies of that module: */
#define BMPC_HMM_MAINTAIN 0 /* HMM_backup_dyn.backup_maintain :105 */
#define BMPC_HMM_BRAKE 1    /* HMM_backup_dyn.backup_brake :107-109 (numeric softmax(-5, -v, 3)) */

typedef struct bmpc_hmm_params {
  double Kpsi, L, W, ylb, yub, col_alpha, s1, s2, c2, tran_diag; /* utils HMM/Branch constants */
} bmpc_hmm_params;

/* PredictiveModel.generate_backup_traj (:204-214): x0 [count][M][4] -> xbackup [count][M*m][N*4], row m*i+j = agent i
 * under policy j, flattened component-major as casadi.reshape does.  policy_kind: m HOST ints. */
int bmpc_hmm_backup_rollout(const double* x0, int64_t count, int32_t M, int32_t m, const int32_t* policy_kind, int32_t N,
                            double dt, double Kpsi, double* xbackup, int32_t device, void* stream);

/* module-level generate_backup_traj with sensitivity (:54-85): x0 [count][4]; outputs for every (point, policy):
 * xx [count][m][steps][4], QQ [count][m][steps][16] (dx_t/dx_0), Qt [count][m][steps][4] (xdot - f0). f0: 4 device doubles. */
int bmpc_hmm_rollout_sensitivity(const double* x0, int64_t count, int32_t m, const int32_t* policy_kind, int32_t steps,
                                 double ts, double Kpsi, const double* f0, double* xx, double* QQ, double* Qt,
                                 int32_t device, void* stream);

/* belief transition (backup_trans :96-101 inside calc_xp_expr :249-257) and, when cbf != NULL, the environment's update
 * (Highway_env.py:251-256): ego [count][4], xb [count][M][m][4] (backup state of every agent/policy at the evaluated
 * time), b [count][M][m], cbf [count][M][m] or NULL.  clip=1: numeric veh_col (+-5 clip, :145-147), clip=0: symbolic.
 * Outputs (NULL = skip): h [count][M][m], H [count][M][m][m], b_next [count][M][m]. */
int bmpc_hmm_belief_update(const double* ego, const double* xb, const double* b, const double* cbf, int64_t count,
                           int32_t M, int32_t m, const bmpc_hmm_params* p, int32_t clip, double* h, double* H,
                           double* b_next, int32_t device, void* stream);

/* Plant step of the reference environments, batched on the device (vehicle.step, Highway_env_branch.py:39-41;
 * robot.step, quadruped_env.py:24-40): x <- x + dt f(x, u) for the ego with the applied input u [count][d], and the
 * obstacle z under backup policy `obstacle_policy` (index into the handle's policy table, per-episode parameters from
 * policy_params as in bmpc_solve).  In place; device pointers; either state pointer may be NULL. */
int bmpc_plant_step(bmpc_handle* h, double* x, const double* u, double* z, int32_t obstacle_policy,
                    const double* policy_params, int64_t count, void* stream);

/* Closed-loop environment state of a batch of episodes (device pointers, caller-owned).  Replaces the per-episode
 * Python objects of Highway_env_branch.py:28-81 (vehicle, Highway_env) and quadruped_env.py:24-65 (robot, Quad_env). */
typedef struct bmpc_env_state {
  double* x;             /* [count][n] ego state, advanced in place                                              */
  double* z;             /* [count][n] obstacle state, advanced in place                                         */
  int32_t* lane;         /* [count][2] highway: lane index of ego / obstacle (vehicle.laneidx); NULL for the quadruped */
  double* policy_params; /* [count][m][4] per-episode policy parameters handed to the controller; the lane-change row is
                            rewritten when the obstacle changes lane (update_backup, Highway_env_branch.py:117-118)   */
  const double* goal;    /* [count][n] quadruped: desired final state x_des (quadruped_env.py:57); NULL on the highway */
  int32_t* obs_policy;   /* [count] out: arg-max backup policy of the obstacle (veh_set[1].backupidx)             */
  int32_t* collided;     /* [count] in/out: sticky collision flag of Highway_sim (Highway_env_branch.py:421-429)   */
  double* xref;          /* [count][n] out: the reference this step handed to the controller                       */
  double* u_obs;         /* [count][d] out: the input applied to the obstacle                                      */
} bmpc_env_state;

/* One closed-loop step of every episode, all on `stream`: Highway_env.step (Highway_env_branch.py:83-184) or
 * Quad_env.step (quadruped_env.py:67-130) - obstacle arg-max policy from the numeric collision functions, lane
 * bookkeeping and lane-change target, the xRef rule, the controller solve (as bmpc_solve with the state's x, z, xref and
 * policy_params; `out` as there, out->u0 is required) and both Euler plants.  t is the step counter t_ of the reference
 * (t == 0 initialises the lane indices), n_lane the environment's lane count (highway).  quad_sizes = {L1, L2, col_tol}
 * for the quadruped (Quad_constants), NULL on the highway. */
int bmpc_env_step(bmpc_handle* h, const bmpc_env_state* env, int64_t count, int32_t t, int32_t n_lane,
                  const double* quad_sizes, const bmpc_outputs* out, void* stream);

/* Merge scenario, one control period of Highway_env_merge.step (Highway_env_branch.py:324-380) for every episode, on the
 * device: collision flag, the ego's lane id (1 = ramp, 0 = highway, sticky once x > merge_s + 8), the reference / state
 * transform / state bounds of the call (ramp coordinates from the lookup tables refY, refpsi at the ego's x while on the ramp),
 * the controller solve (as bmpc_solve_transformed), and both plants (the obstacle follows the handle's first policy).
 * BMPC_MODEL_MERGE handles.  All pointers of the state are DEVICE pointers; the tables are device arrays of table_n points
 * (grid strictly increasing). */
typedef struct bmpc_merge_env_state {
  double* x;            /* [count][4] ego state, in/out */
  double* z;            /* [count][4] obstacle state, in/out */
  int32_t* lane_id;     /* [count] in/out */
  int32_t* collided;    /* [count] in/out, sticky */
  double* xref;         /* [count][4] out */
  double* S;            /* [count][4][4] out */
  double* state_bounds; /* [count][2][2] out */
  double* u_obs;        /* [count][2] out */
  const double* table_x;   /* ramp centre line (merge_geometry :227-262): grid */
  const double* table_y;   /*   y of the centre line   (refY)   */
  const double* table_psi; /*   heading                (refpsi) */
  int32_t table_n;
} bmpc_merge_env_state;
int bmpc_env_step_merge(bmpc_handle* h, const bmpc_merge_env_state* env, int64_t count, int32_t n_lane, int32_t merge_lane,
                        double merge_s, double v0, const bmpc_outputs* out, void* stream);

/* 1 if the solve kernel of this handle stages the NEXT episode's persistent state (uLin, active-set codes, rho cache) into
 * shared memory with bulk copies (cp.async.bulk + mbarrier) while it solves the current one; 0 if it reads them from global
 * memory at the start of each solve (the default; staging is requested with reserved[6] bit 1 and applies to the tree
 * controllers in the shared slab placement). */
int bmpc_staging_enabled(const bmpc_handle* h);

/* How the solve kernel is launched for this handle: resolved BMPC_SLAB_* placement, number of persistent warps
 * (= thread blocks of 32), dynamic shared memory per warp, bytes of the per-warp global region. */
int bmpc_get_launch_info(const bmpc_handle* h, int32_t* slab_mode, int32_t* warps, int64_t* smem_bytes,
                         int64_t* global_bytes_per_warp);

/* Kernel launch accounting since creation (for bench.py's gpu_launches). */
int64_t bmpc_launch_count(const bmpc_handle* h);

/* Measured FP64 FMA throughput of this GPU in TFLOP/s (a dependent-chain-free DFMA loop over all
 * SMs; used as the roofline denominator by bench.py). Returns < 0 on error. */
double bmpc_measure_fp64_peak(int device, int iters);

/* Time of the last bmpc_solve kernel in milliseconds measured with CUDA events on its stream
 * (valid after the stream has been synchronised). */
float bmpc_last_kernel_ms(bmpc_handle* h);

const char* bmpc_last_error(const bmpc_handle* h);

#ifdef __cplusplus
}
#endif
#endif /* BRANCHMPC_H_ */
