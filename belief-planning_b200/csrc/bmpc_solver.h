// Warp-per-problem tree-QP solver: ADMM whose KKT step is a Riccati sweep over the branch tree, finished by a
// primal-dual active-set polish that certifies the exact optimum.  One warp owns one scenario-tree MPC problem; the
// per-node data lives in that warp's shared-memory slab (field-major, one pad slot per branch so that the lanes of a
// sweep hit distinct banks); lanes are the parallel branches of a tree level during sweeps and the nodes during the
// row phase.  Replaces buildCost + buildEqConstr + buildIneqConstr + osqp_solve_qp + unpackSolution of the reference
// (MPC_branch.py:984-1274) and the per-node CasADi calls of inittree/updatetree (:928-981, :1024-1061).
//
// The QP is solved in its slack-free exact-penalty form: the reference's slack variables
// s >= 0 with linear cost lam = Qslack[1]*w are eliminated, leaving lam*max(0, f'x - hi) per soft row, whose prox is
// closed form.  Opposite rows of Fx are paired into one two-sided row lo <= f'x <= hi.
//
// ADMM bookkeeping: per row only the scaled Moreau variable sh = rho*(v + y/rho) is stored; the multiplier is
//   y(sh) = clamp(sh - rho*hi, 0, lam)  above,  clamp(sh - rho*lo, -lam, 0)  below,  0 inside,
// and rho*v = sh - y, so no division appears in the iteration.
#pragma once
#include "bmpc_models.h"

enum { ROW_INACTIVE = 0, ROW_UP_KINK = 1, ROW_UP_LIN = 2, ROW_LO_KINK = 3, ROW_LO_LIN = 4, ROW_IGNORED = 7 };
enum { IN_FREE = 0, IN_AT_HI = 1, IN_AT_LO = 2 };
enum { FACT_ADMM = 0, FACT_FREE = 1, FACT_POLISH = 2, FACT_IPM = 3 };
enum { IPM_PRED_ASM = 0, IPM_PRED_STEP = 1, IPM_CORR_ASM = 2, IPM_CORR_STEP = 3 };
typedef long long code_t;   // per-node active-set code (see Solver::row_of / in_of)

#define BMPC_NOLO (-1.0e300)

#if defined(BMPC_HOSTSIM_TRACE) && !defined(__CUDACC__)
#include <stdio.h>
#define BMPC_TRACE(...) fprintf(stderr, __VA_ARGS__)
#else
#define BMPC_TRACE(...) ((void)0)
#endif

// multiplier of a two-sided soft row from the scaled Moreau variable (rlo = rho*lo, rhi = rho*hi)
BMPC_D real row_dual(real sh, real rlo, real rhi, real lam) {
  if (sh > rhi) return fmin(sh - rhi, lam);
  if (sh < rlo) return fmax(sh - rlo, -lam);
  return 0.0;
}

// Working-set placement (MODE = BMPC_SLAB_*), all node-major (one node's fields are contiguous, so a node step
// addresses them as base + immediate offset):
//   SHARED: every field in the warp's shared-memory slab (record stride odd -> the lanes of a sweep and of the row
//           phase hit distinct banks);
//   SPLIT : the factor fields (written by a factorisation, read-only during the ADMM sweeps) in a per-warp global/L2
//           region, the iterate fields in shared memory (~3x smaller shared footprint, ~3x more resident warps);
//   GLOBAL: everything in the per-warp global region (trees too large for shared memory).
#if defined(__CUDACC__)
extern __shared__ __align__(16) real bmpc_smem[];
// parameter block of the solve kernel in flight on this device (written stream-ordered before every launch, bmpc_solve)
__constant__ KParams bmpc_cP;
#endif

// NR = soft rows per node: NC collision rows (1 for the branch controllers; robustMPC has one per obstacle node of the
// time slot, up to m^NB) followed by the two-sided state rows of the configuration.
template <class M, int NR, int MODE = BMPC_SLAB_SHARED, int NC = 1>
struct Solver {
  static constexpr bool SPLIT = (MODE == BMPC_SLAB_SPLIT);
  static constexpr int NX = M::NX, NU = M::NU;   // Riccati state (physical, or physical + previous input) / input
  static constexpr int NXP = M::NXP;             // physical state dimension
  static constexpr bool RATE = M::RATE;          // input-rate costs carried through the augmented state
  static constexpr int NS = NX * (NX + 1) / 2;
  static constexpr int NSU = NU * (NU + 1) / 2;
  // field offsets (each field is one real per padded node)
  // Multi-real fields start on even offsets of a record with even stride: records are 16-byte aligned, so the compiler
  // fetches neighbouring reals with one LDS.128 (a backward-sweep step: 14 shared loads instead of 27).  The order keeps the
  // padding small (highway, 3 rows: 49 reals in a record of 50).
#if defined(BMPC_ODD_FIELDS)
  static constexpr int ev(int v) { return v; }
#else
  static constexpr int ev(int v) { return (v + 1) & ~1; }
#endif
  static constexpr int F_LIN = 0;
  static constexpr int F_CC = ev(F_LIN + M::NLIN);
  static constexpr int F_Q = ev(F_CC + M::NCC);
  static constexpr int F_FC = ev(F_Q + NXP);       // collision row f (x,y components); holds the obstacle (x,y) before setup
  static constexpr int F_K = ev(F_FC + 2 * NC);    // feedback gain
  static constexpr int F_H0 = ev(F_K + NU * NX);   // P+ C
  static constexpr int F_SI = ev(F_H0 + NX);       // S^-1, packed upper triangle
  static constexpr int F_HC = F_SI + NSU;          // collision row upper bounds
  static constexpr int F_RHO = F_HC + NC;          // rho of the NR soft rows then of the NU inputs
  // ---- iterate fields (kept in the slab in every placement) ----
  static constexpr int F_XQ = ev(F_RHO + NR + NU); // x after a forward sweep / q~x before a backward sweep
  static constexpr int F_UQ = ev(F_XQ + NX);       // u (kff between the sweeps) / q~u before a backward sweep
  static constexpr int F_SU = ev(F_UQ + NU);       // ADMM state of the inputs: scaled Moreau variable
  static constexpr int F_S = ev(F_SU + NU);        // same per soft row
  static constexpr int F_Y = F_S + NR;             // polish multipliers (rows then inputs)
  static constexpr int NF = F_Y + NR + NU;
  static constexpr int BR = 1 + NS + 3 * NX; // per-branch reals: w, exchange(NS), x last, z last, x after last
  static constexpr int NFA = F_XQ;                // factor fields per node
  static constexpr int NFAP = (NFA + 1) & ~1;     // ... padded to an even count (16-byte aligned records in global)
  static constexpr int NFW = SPLIT ? NF - NFA : NF;   // fields kept in the slab
#if defined(BMPC_ODD_FIELDS)
  static constexpr int NFWP = NFW | 1;                // round-1 layout: odd stride (conflict-free for 64-bit accesses)
#else
  static constexpr int NFWP = (NFW + 1) & ~1;         // even stride (see ev above); 128-bit accesses of 8 lanes cover the 32 banks
#endif

  // merge scenario: the episode's transformed stage Hessian S'QS, state rows Fx S and their bounds (kStateTransform)
  static constexpr bool XF = M::kStateTransform;
  static constexpr int XFR = XF ? NXP * NXP + (NR - NC) * (NXP + 2) : 0;
  BMPC_HD static size_t slab_reals(int nup, int nbranch) {   // nbranch = KParams::nbx
    return (size_t)NFWP * nup + (size_t)nup + (size_t)BR * nbranch + XFR;
  }
  BMPC_HD static size_t factor_reals(int nup) { return SPLIT ? (size_t)NFAP * nup : 0; }
  // per-episode strides of the solver caches in global memory, padded to 16 bytes so that one episode's block can travel
  // as one bulk copy; the stage behind the slab holds the NEXT episode's uLin | active-set codes | rho cache
  BMPC_HD static size_t code_stride(int totalu) { return ((size_t)totalu + 1) & ~(size_t)1; }
  BMPC_HD static size_t rho_stride(int totalu) { return ((size_t)totalu * (NR + NU) + 1) & ~(size_t)1; }
  BMPC_HD static size_t ulin_reals(int totalu) { return (size_t)(totalu + 1) * NU; }
  BMPC_HD static bool stage_possible(int totalu) { return (ulin_reals(totalu) & 1) == 0; }
  BMPC_HD static size_t stage_offset(int nup, int nbranch) { return (slab_reals(nup, nbranch) + 1) & ~(size_t)1; }
  BMPC_HD static size_t stage_reals(int totalu) { return ulin_reals(totalu) + code_stride(totalu) + rho_stride(totalu); }
  // interior-point fallback: per-node scratch in this warp's global region (rare path, see ipm_solve)
  static constexpr int IP_X = 0;                    // current iterate x, u
  static constexpr int IP_U = IP_X + NX;
  static constexpr int IP_RS = IP_U + NU;           // per soft row: s+, y+, s-, y-
  static constexpr int IP_IN = IP_RS + 4 * NR;      // per input: y+, y-
  static constexpr int IP_KAP = IP_IN + 2 * NU;     // barrier curvature of the rows / inputs (penalties of FACT_IPM)
  static constexpr int IP_COR = IP_KAP + NR + NU;   // per complementarity side: two carried values (see ipm_side)
  static constexpr int NIP = IP_COR + 4 * NR + 4 * NU;
  BMPC_HD static size_t ipm_reals(int nup) { return (size_t)NIP * nup; }

#if defined(__CUDA_ARCH__)
  // device: every phase function reads the parameter block straight from constant memory (LDC with immediate offsets)
  // instead of through a reference member (a generic load per field in each out-of-line function)
#define PP bmpc_cP
#else
  const KParams& P_host;
#define PP P_host
#endif
  real* ws;     // slab base (shared memory is addressed through bmpc_smem on the device so that loads are LDS)
  real* fa;     // factor-field region (SPLIT only)
  real* ip;     // interior-point scratch (global)
  // offsets (in reals) of the per-branch arrays inside the slab: recomputed from the constant-memory block where needed
  // (a member would be a generic load through `this` in every out-of-line function)
  BMPC_D int oSt() const { return NFWP * PP.nup; }
  BMPC_D int oWb() const { return oSt() + PP.nup; }
  BMPC_D int oEX() const { return oWb() + PP.nbx; }
  BMPC_D int oEXL() const { return oEX() + NS * PP.nbx; }
  BMPC_D int oEXZ() const { return oEXL() + NX * PP.nbx; }
  BMPC_D int oEXX() const { return oEXZ() + NX * PP.nbx; }
  BMPC_D int oXF() const { return oEXX() + NX * PP.nbx; }
#if defined(__CUDA_ARCH__)
#define BMPC_LANE_ID ((int)threadIdx.x)
#else
  int lane_host;
#define BMPC_LANE_ID lane_host
#endif
  int prob;
#if defined(__CUDA_ARCH__)
  // staging of the next episode (P.stage_on): mbarrier + hand-over slots in static shared memory of the team's block
  unsigned long long* stage_bar;
  int* stage_slot;        // [0] work index claimed for the next solve, [1] 1 if bulk copies for it are in flight, [2] wait verdict
  unsigned stage_parity;  // phase of the mbarrier the next wait looks for
  bool stage_have;        // copies for THIS episode were issued during the previous solve
  bool stage_broken;      // a wait timed out: staging stays off for the rest of the launch (this team)
  bool stage_pending;     // the hand-over (claim + copies for the next episode) has not run yet in this solve
  bool staged;            // this episode's uLin / codes / rho are read from the stage
#endif
  bool use_codes;   // this solve starts its polish from the cached active set of the previous step
  real gap_r, stp_r, gap_u, stp_u;   // last residual check: max primal gap |f'x - v| and max step |v+ - v| (rows / inputs)
  int set_changes;                   // last residual check: nodes whose implied active set differs from the previous check
  int nsolve;   // KKT solves (one backward + one forward sweep each) of the current problem
  int ipm_iters;   // interior-point iterations of the current problem
  real ipm_acc, ipm_ratio;   // per-lane results of the last ipm_pass
  real rlin;  // linear cost on every component of the root input: -2 * OldInput . dR  (MPC_branch.py:1099)
  const real* polpar;

#if defined(__CUDA_ARCH__)
  BMPC_D Solver(const KParams& P_, real* slab, real* factor, real* ipm, int lane_) : ws(slab), fa(factor), ip(ipm) {
    (void)lane_;
#else
  BMPC_D Solver(const KParams& P_, real* slab, real* factor, real* ipm, int lane_) : P_host(P_), ws(slab), fa(factor), ip(ipm), lane_host(lane_) {
#endif
    prob = 0;
#if defined(__CUDA_ARCH__)
    stage_bar = nullptr;
    stage_slot = nullptr;
    stage_parity = 0;
    stage_have = stage_broken = stage_pending = staged = false;
#endif
    use_codes = false;
    nsolve = 0;
    rlin = 0.0;
    polpar = nullptr;
  }

  BMPC_D real* slab() {
#if defined(__CUDA_ARCH__)
    if (MODE != BMPC_SLAB_GLOBAL) return bmpc_smem;   // one warp per block: the slab starts at shared offset 0
#endif
    return ws;
  }
  BMPC_D real& F(int field, int kp) {
    if (SPLIT) {
      if (field < NFA) return fa[kp * NFAP + field];
      return slab()[kp * NFWP + (field - NFA)];
    }
    return slab()[kp * NFWP + field];
  }
  // phase accounting (development aid): with reserved[7] = k > 0 the `cycles` output reports the time spent in phase k
  // (1 interior point, 2 tree expansion, 3 rho selection/cache, 4 factorisations, 5 backward+forward sweeps, 6 node-parallel
  // polish passes, 7 adjoint sweep, 8 ADMM row phase, 9 final pass + caches)
  long long t_phase;
  BMPC_D long long prof_begin(int k) {
#if defined(__CUDA_ARCH__)
    if (PP.cycles_mode == k) return clock64();
#endif
    return 0;
  }
  BMPC_D void prof_end(int k, long long t0) {
#if defined(__CUDA_ARCH__)
    if (PP.cycles_mode == k) t_phase += clock64() - t0;
#endif
  }
  BMPC_D real& IPF(int field, int kp) { return ip[(size_t)field * PP.nup + kp]; }   // field-major: the lanes of a node-parallel pass read consecutive addresses
  BMPC_D code_t* stp() { return reinterpret_cast<code_t*>(slab() + oSt()); }
  BMPC_D real* Wbp() { return slab() + oWb(); }
  BMPC_D real* EXp() { return slab() + oEX(); }
  BMPC_D real* EXLp() { return slab() + oEXL(); }
  BMPC_D real* EXZp() { return slab() + oEXZ(); }
  BMPC_D real* EXXp() { return slab() + oEXX(); }
  // stage Hessian (NXP x NXP), state rows and their bounds of this problem: the handle's, or the episode's transformed ones
  BMPC_D real* XFp() { return slab() + oXF(); }
  BMPC_D const real* QH() { if constexpr (XF) return XFp(); else return PP.Q; }
  BMPC_D real RF(int r, int i) { if constexpr (XF) return XFp()[NXP * NXP + r * NXP + i]; else return PP.rf[r][i]; }
  BMPC_D real RLO(int r) { if constexpr (XF) return XFp()[NXP * NXP + (NR - NC) * NXP + r]; else return PP.rlo[r]; }
  BMPC_D real RHI(int r) { if constexpr (XF) return XFp()[NXP * NXP + (NR - NC) * (NXP + 1) + r]; else return PP.rhi[r]; }
  BMPC_D int RF_ONE(int r) { if constexpr (XF) return -1; else return PP.rf_one[r]; }
  // active-set code of a node: 3 bits per soft row, then 2 bits per input
  BMPC_D static int row_of(code_t c, int j) { return (int)((c >> (3 * j)) & 7); }
  BMPC_D static int in_of(code_t c, int a) { return (int)((c >> (3 * NR + 2 * a)) & 3); }
  BMPC_D static code_t row_bits(int v, int j) { return (code_t)v << (3 * j); }
  BMPC_D static code_t in_bits(int v, int a) { return (code_t)v << (3 * NR + 2 * a); }
  static_assert(3 * NR + 2 * NU <= 62, "active-set code does not fit 64 bits");

  BMPC_D int kp_of(int b, int t) const { return bmpc_ndu(PP, b) + t + b; }
  // Node-parallel passes: slot i of the team's lanes works on input node i + 1, the last slot on the root, so that the
  // non-root nodes (a multiple of N, e.g. 96) fill whole rounds of the team and the root (whose soft rows are inert:
  // its state is fixed) is one short extra round of lane 0.
  BMPC_D int node_at(int i) const { return (i + 1 < PP.totalu) ? i + 1 : 0; }
#define BMPC_FOR_NODES(k) \
  for (int k##_i = BMPC_LANE_ID, k = node_at(k##_i); k##_i < PP.totalu; k##_i += BMPC_LANES, k = node_at(k##_i))
  BMPC_D void node_of(int k, int& b, int& t) const {
    if (k == 0) { b = 0; t = 0; } else { const int q = bmpc_idiv(k - 1, PP.inv_N); b = 1 + q; t = (k - 1) - q * PP.N; }
  }
  BMPC_D const real* pol_par(int i) const { return polpar ? polpar + 4 * i : PP.pol_par[i]; }
  // persistent per-episode blocks: from the stage when this episode was prefetched, else from global memory
  BMPC_D real* stage_base() { return slab() + stage_offset(PP.nup, PP.nbx); }
  BMPC_D const real* ep_uLin() {
#if defined(__CUDA_ARCH__)
    if (staged) return stage_base();
#endif
    return PP.uLin + (size_t)prob * ulin_reals(PP.totalu);
  }
  BMPC_D const code_t* ep_codes() {
#if defined(__CUDA_ARCH__)
    if (staged) return reinterpret_cast<const code_t*>(stage_base() + ulin_reals(PP.totalu));
#endif
    return PP.code_cache + (size_t)prob * code_stride(PP.totalu);
  }
  BMPC_D const real* ep_rho() {
#if defined(__CUDA_ARCH__)
    if (staged) return stage_base() + ulin_reals(PP.totalu) + code_stride(PP.totalu);
#endif
    return PP.rho_cache + (size_t)prob * rho_stride(PP.totalu);
  }
#if defined(__CUDA_ARCH__)
  // Start of a solve: wait (bounded) for the bulk copies issued for this episode during the previous solve.
  BMPC_DN void stage_acquire() {
    staged = false;
    if (stage_have) {
      stage_have = false;
      if (BMPC_LANE_ID == 0) {
        bool ok = false;
        const long long t0 = clock64();
        while (!(ok = bmpc_mbar_try_wait(stage_bar, stage_parity)) && clock64() - t0 < (4ll << 20)) {}
        stage_slot[2] = ok ? 1 : 0;
      }
      team_sync();
      stage_parity ^= 1u;
      staged = stage_slot[2] == 1;
      if (!staged) stage_broken = true;   // never observed; the episode is then read from global memory as without staging
    }
    stage_pending = PP.stage_on && !stage_broken;
  }
  // Hand-over, once per solve behind the team barrier that follows the last read of the stage: claim the next work item
  // and start the three bulk copies (TMA, 1-D) of its episode's uLin | codes | rho; they land while this episode is solved.
  BMPC_DN void stage_next() {
    stage_pending = false;
    if (BMPC_LANE_ID == 0) {
      // claiming ahead fixes a team's next item one solve early; near the end of the queue that would cost the balance of
      // the tail (measured: -3 %), so the last gridDim.x items are claimed when their team is actually free
      int nxt = -1, issued = 0;
      if (*(volatile int*)PP.counter + (int)gridDim.x < PP.count) nxt = atomicAdd(PP.counter, 1);
      if (nxt >= 0 && nxt < PP.count) {
        const size_t e = (size_t)(PP.order ? PP.order[nxt] : nxt);
        const unsigned b0 = (unsigned)(ulin_reals(PP.totalu) * sizeof(real)), b1 = (unsigned)(code_stride(PP.totalu) * sizeof(code_t)),
                       b2 = (unsigned)(rho_stride(PP.totalu) * sizeof(real));
        real* st = stage_base();
        bmpc_fence_proxy_async();
        bmpc_mbar_expect_tx(stage_bar, b0 + b1 + b2);
        bmpc_bulk_g2s(st, PP.uLin + e * ulin_reals(PP.totalu), b0, stage_bar);
        bmpc_bulk_g2s(st + ulin_reals(PP.totalu), PP.code_cache + e * code_stride(PP.totalu), b1, stage_bar);
        bmpc_bulk_g2s(st + ulin_reals(PP.totalu) + code_stride(PP.totalu), PP.rho_cache + e * rho_stride(PP.totalu), b2, stage_bar);
        issued = 1;
      }
      stage_slot[0] = nxt;
      stage_slot[1] = issued;
    }
  }
#endif

  // ========================================================================================
  // Tree expansion: obstacle rollouts, branch probabilities/weights, ego linearisation rollouts,
  // per-node linearisation + collision linearisation + cost vectors   (kernels K1-K3 of SURVEY.md)
  // ========================================================================================
  // collision rows (col_eval, :1114-1168: -dh x - s <= h - dh xbar) against the `ncol` obstacle positions parked in the
  // FC slots, and the ADMM start of every soft row (unscaled, multiplied by rho in choose_rho / load_rho): rows at the
  // linearisation point, projected on their bounds
  BMPC_D void rows_setup(int kp, const real* xbar, int ncol) {
#pragma unroll
    for (int j = 0; j < NC; ++j) {
      if (j < ncol) {
        real zxy[2] = {F(F_FC + 2 * j, kp), F(F_FC + 2 * j + 1, kp)};
        real h, dhx, dhy;
        M::collision(PP, xbar, zxy, h, dhx, dhy);
        const real hi0 = h - (dhx * xbar[0] + dhy * xbar[1]);
        if constexpr (XF) {
          // updateIneqConstr with a state transform (MPC_branch.py:2027-2031): on every solve after the first the x-component
          // of the collision gradient is pushed away from zero AFTER col_eval formed the row's offset with the true one
          if (PP.xform && PP.started[prob]) dhx = sgn(dhx) * fmax(0.1, fabs(dhx));
        }
        const real fx = -dhx, fy = -dhy;
        F(F_FC + 2 * j, kp) = fx;
        F(F_FC + 2 * j + 1, kp) = fy;
        F(F_HC + j, kp) = hi0;
        F(F_S + j, kp) = fmin(fx * xbar[0] + fy * xbar[1], hi0);
      } else {
        F(F_FC + 2 * j, kp) = 0.0;       // no obstacle in this slot: a zero row gets rho = 0 and is ignored
        F(F_FC + 2 * j + 1, kp) = 0.0;
        F(F_HC + j, kp) = 1.0e30;
        F(F_S + j, kp) = 0.0;
      }
    }
#pragma unroll
    for (int j = NC; j < NR; ++j) {
      real v = 0.0;
#pragma unroll
      for (int i = 0; i < NXP; ++i) v += RF(j - NC, i) * xbar[i];
      F(F_S + j, kp) = bmpc_clamp(v, RLO(j - NC), RHI(j - NC));
    }
  }

  // Expansion scratch, parked in the gain fields (dead until the first factorisation): the obstacle state of the node's
  // time step, the position of the ego "maintain" rollout, and the node's safety terms.
  static constexpr int F_ZS = F_K;
  static constexpr int F_ZE = F_ZS + NXP;
  static constexpr int F_ZV = F_ZE + 2;
  static_assert(NXP + 4 <= NU * NX, "expansion scratch does not fit the gain fields");

  // per-node data that does not depend on the rollout order: linear cost, collision row, ADMM start, xLin output
  // (the linearisation itself was stored by the rollout sweep, which needs the successor state anyway)
  BMPC_DN void node_setup_all() {
    const real* xref = PP.xref + (size_t)prob * NXP;
#pragma unroll 1
    BMPC_FOR_NODES(k) {
      int b, t;
      node_of(k, b, t);
      const int kp = kp_of(b, t);
      const real w = Wbp()[b];
      const bool leaf_last = (b >= PP.off[PP.NB] && t == PP.N - 1);
      real xbar[NXP];
#pragma unroll
      for (int i = 0; i < NXP; ++i) xbar[i] = F(F_XQ + i, kp);
      // linear state cost -2 w (xRef' Q° + xbar' dQ): Q° = Qf on the last node of a leaf branch of BranchMPC (:1095)
      const real* Ql = (leaf_last && PP.ctrl == BMPC_CTRL_BRANCH) ? PP.Qf : PP.Q;
#pragma unroll
      for (int j = 0; j < NXP; ++j) {
        real a = 0.0, c = 0.0;
#pragma unroll
        for (int i = 0; i < NXP; ++i) {
          a += xref[i] * Ql[i * NXP + j];
          c += xbar[i] * PP.Q[i * NXP + j];
        }
        F(F_Q + j, kp) = -2.0 * w * (a + PP.dq_scale * c);
      }
      rows_setup(kp, xbar, 1);
#pragma unroll
      for (int a = 0; a < NU; ++a) F(F_SU + a, kp) = bmpc_clamp(F(F_UQ + a, kp), PP.ulo[a], PP.uhi[a]);
      if (PP.out.xLin) {
        real* o = PP.out.xLin + ((size_t)prob * PP.totalu + k) * NXP;
#pragma unroll
        for (int i = 0; i < NXP; ++i) o[i] = xbar[i];
      }
    }
    lanes_sync();
  }

  // Order of the expansion (everything that is not a recursion in time runs node-parallel over the whole team):
  //   1. node-parallel: time-shifted previous inputs and active-set codes of every node (updatetree :1025-1033)
  //   2. per tree level, lanes = (chain, child branch): the three rollouts of a branch - obstacle under its policy
  //      (zpred_eval), ego under policy 0 (BF_traj) and the ego linearisation trajectory with its per-node A, B, C
  //      (:1048-1059) - are stepped by one instruction stream; they do not depend on the branch probabilities
  //   3. node-parallel: safety terms of every (branch, step); soft-min per branch in two passes (minimum, then the
  //      shifted exponentials - identical to the reference's unshifted form, highway_branch_dyn.py:151-162)
  //   4. lanes = parents: probabilities, weights, arg-max children (branch_eval)
  //   5. node-parallel: cost vectors, collision rows (col_eval), ADMM start
  // first warp of the team only; the root by lane 0, then level by level (see the level loop)
  BMPC_DN void rollouts() {
    const real* x0 = PP.x0 + (size_t)prob * NXP;
    const real* z0 = PP.z0 + (size_t)prob * NXP;
    const int m = PP.m;
    if (BMPC_LANE_ID == 0) {
      const int kp = kp_of(0, 0);
      real xb[NXP], ub[NU], xn[NXP], lin[M::NLIN], cc[M::NCC];
#pragma unroll
      for (int i = 0; i < NXP; ++i) xb[i] = x0[i];
#pragma unroll
      for (int a = 0; a < NU; ++a) ub[a] = F(F_UQ + a, kp);
      M::linearize(PP, xb, ub, lin, cc, xn);
#pragma unroll
      for (int i = 0; i < M::NLIN; ++i) F(F_LIN + i, kp) = lin[i];
#pragma unroll
      for (int i = 0; i < M::NCC; ++i) F(F_CC + i, kp) = cc[i];
      F(F_FC, kp) = z0[0];
      F(F_FC + 1, kp) = z0[1];
      Wbp()[0] = 1.0;
      if (PP.out.zPred) {
        real* o = PP.out.zPred + (size_t)prob * PP.totalu * NXP;
#pragma unroll
        for (int i = 0; i < NXP; ++i) o[i] = z0[i];
      }
#pragma unroll
      for (int i = 0; i < NXP; ++i) {
        F(F_XQ + i, kp) = xb[i];
        EXLp()[i] = xb[i];
        EXZp()[i] = z0[i];
        EXXp()[i] = xn[i];
      }
    }
    bsync();
#pragma unroll 1
    for (int d = 0; d < PP.NB; ++d) {
      // lanes = (chain, child branch): chain 0 the ego under policy 0 (BF_traj), chain 1 the obstacle under the child's policy
      // (zpred_eval), chain 2 the ego linearisation trajectory under the shifted previous inputs (:1048-1059).  The three
      // chains of a branch are the same recursion x+ = f(x, u(x)) with different inputs, so one instruction stream steps all
      // of them: the leader warp's time follows its instruction count (profiles/r02_staging_ab.md), and a lane that walked the
      // three chains one after the other issued the step three times.
      const int cnt = PP.pw[d] * m;
#pragma unroll 1
      for (int idx = BMPC_LANE_ID; idx < 3 * cnt; idx += BMPC_BLANES) {
        const int ch = (idx >= 2 * cnt) ? 2 : (idx >= cnt ? 1 : 0);
        const int ci = idx - ch * cnt;
        const int bq = bmpc_idiv(ci, PP.inv_m);
        const int b = PP.off[d] + bq;
        const int i = ci - bq * m;
        const int c = bmpc_first_child(PP, b, d) + i;
        const int kc = bmpc_ndu(PP, c);
        const int kpc = kp_of(c, 0);
        const real* parg = pol_par(ch == 0 ? 0 : i);
        const real par[4] = {parg[0], parg[1], parg[2], parg[3]};   // in registers for the whole rollout
        const int kind = PP.pol_kind[ch == 0 ? 0 : i];
        real* zout = (ch == 1 && PP.out.zPred) ? PP.out.zPred + ((size_t)prob * PP.totalu + kc) * NXP : nullptr;
        const real* start = (ch == 0 ? EXLp() : (ch == 1 ? EXZp() : EXXp())) + NX * b;
        real x[NXP];
#pragma unroll
        for (int q = 0; q < NXP; ++q) x[q] = start[q];
#pragma unroll 1
        for (int t = 0; t < PP.N; ++t) {
          const int kp = kpc + t;
          real u[NU], xn[NXP], sn, cs;
          bmpc_sincos_inline(x[M::HEADING], sn, cs);
          if (ch == 2) {
            real lin[M::NLIN], cc[M::NCC];
#pragma unroll
            for (int a = 0; a < NU; ++a) u[a] = F(F_UQ + a, kp);
            M::lin_only(PP, x, u, sn, cs, lin, cc);
#pragma unroll
            for (int q = 0; q < M::NLIN; ++q) F(F_LIN + q, kp) = lin[q];
#pragma unroll
            for (int q = 0; q < M::NCC; ++q) F(F_CC + q, kp) = cc[q];
#pragma unroll
            for (int q = 0; q < NXP; ++q) F(F_XQ + q, kp) = x[q];
            if (t == PP.N - 1) {
#pragma unroll
              for (int q = 0; q < NXP; ++q) EXLp()[NX * c + q] = x[q];
            }
          } else {
            M::policy_fast(PP, kind, par, x, u);
          }
          M::step_sc(PP, x, u, sn, cs, xn);
#pragma unroll
          for (int q = 0; q < NXP; ++q) x[q] = xn[q];
          if (ch == 0) {
            F(F_ZE, kp) = x[0];
            F(F_ZE + 1, kp) = x[1];
          } else if (ch == 1) {
#pragma unroll
            for (int q = 0; q < NXP; ++q) {
              F(F_ZS + q, kp) = x[q];
              if (zout) zout[t * NXP + q] = x[q];
            }
            F(F_FC, kp) = x[0];
            F(F_FC + 1, kp) = x[1];
          }
        }
        if (ch != 0) {
          real* end = (ch == 1 ? EXZp() : EXXp()) + NX * c;
#pragma unroll
          for (int q = 0; q < NXP; ++q) end[q] = x[q];
        }
      }
      bsync();
    }
  }

  BMPC_DN void expand_tree() {
    const long long prof_t0 = prof_begin(2) + prof_begin(11);   // (11: staging of the shifted inputs, 10: rollouts, 13: safety values
                                                                // and probabilities, 12: node set-up - parts of phase 2)
    const int started = PP.started[prob];
    const real* uLin = ep_uLin();
    int* pbest = PP.pbest + (size_t)prob * PP.nbranch;
    const code_t* codes = ep_codes();
    const int m = PP.m;
    if constexpr (XF) {
      // solve(x, z, xRef, S, Fx, bx) (MPC_branch.py:2043-2059): stage Hessian S'QS (:1938), state rows Fx S (:1899) and the
      // bounds of this call; identity / the handle's bounds where the caller passes none
      const real* Sx = PP.xform ? PP.xform + (size_t)prob * NXP * NXP : nullptr;
      const real* bd = PP.xbounds ? PP.xbounds + (size_t)prob * (NR - NC) * 2 : nullptr;
      real* xf = XFp();
      for (int e = BMPC_LANE_ID; e < XFR; e += BMPC_LANES) {
        real v = 0.0;
        if (e < NXP * NXP) {
          const int i = e / NXP, j = e - i * NXP;
          for (int a2 = 0; a2 < NXP; ++a2)
            for (int b2 = 0; b2 < NXP; ++b2) {
              const real sa = Sx ? Sx[a2 * NXP + i] : (a2 == i ? 1.0 : 0.0), sb = Sx ? Sx[b2 * NXP + j] : (b2 == j ? 1.0 : 0.0);
              v += sa * PP.Q[a2 * NXP + b2] * sb;
            }
        } else if (e < NXP * NXP + (NR - NC) * NXP) {
          const int r = (e - NXP * NXP) / NXP, i = (e - NXP * NXP) - r * NXP;
          for (int a2 = 0; a2 < NXP; ++a2) v += PP.rf[r][a2] * (Sx ? Sx[a2 * NXP + i] : (a2 == i ? 1.0 : 0.0));
        } else if (e < NXP * NXP + (NR - NC) * (NXP + 1)) {
          const int r = e - NXP * NXP - (NR - NC) * NXP;
          v = bd ? bd[2 * r] : PP.rlo[r];
        } else {
          const int r = e - NXP * NXP - (NR - NC) * (NXP + 1);
          v = bd ? bd[2 * r + 1] : PP.rhi[r];
        }
        xf[e] = v;
      }
    }
#pragma unroll 1
    BMPC_FOR_NODES(k) {
      int b, t;
      node_of(k, b, t);
      const int kp = kp_of(b, t);
      int ksrc;
      if (k == 0) {
        // root input: previous first input of the most likely child (updatetree :1029-1031)
        ksrc = bmpc_ndu(PP, bmpc_first_child(PP, 0, 0) + (started ? pbest[0] : 0));
      } else if (t < PP.N - 1) {
        ksrc = k + 1;
      } else if (b >= PP.off[PP.NB]) {
        ksrc = k;   // leaf: repeat the shifted last input (:1033)
      } else {
        ksrc = bmpc_ndu(PP, bmpc_first_child(PP, b, bmpc_depth(PP, b)) + (started ? pbest[b] : 0));
      }
#pragma unroll
      for (int a = 0; a < NU; ++a) F(F_UQ + a, kp) = started ? uLin[ksrc * NU + a] : 0.0;   // zero on the first solve
      if (use_codes) stp()[kp] = codes[ksrc];   // the active set shifts in time like the inputs
    }
    team_sync();
    prof_end(11, prof_t0);
    const long long prof_t1 = prof_begin(10);
    if (team_leader()) rollouts();
    team_sync();
    prof_end(10, prof_t1);
    const long long prof_t2 = prof_begin(13);
    // safety value of every child branch: soft-min (gamma = 5) over the safety terms of its N steps
#pragma unroll 1
    BMPC_FOR_NODES(k) {
      if (k == 0) continue;
      int b, t;
      node_of(k, b, t);
      const int kp = kp_of(b, t);
      real z[NXP], v[2];
#pragma unroll
      for (int q = 0; q < NXP; ++q) z[q] = F(F_ZS + q, kp);
      M::safety_terms(PP, z, F(F_ZE, kp), F(F_ZE + 1, kp), v);
      F(F_ZV, kp) = v[0];
      if (M::NSAFE > 1) F(F_ZV + 1, kp) = v[1];
    }
    team_sync();
    if (team_leader()) {
#pragma unroll 1
    for (int c = 1 + BMPC_LANE_ID; c < PP.nbranch; c += BMPC_BLANES) {
      const int kpc = kp_of(c, 0);
      real vmin = 1e300;
#pragma unroll 1
      for (int t = 0; t < PP.N; ++t) {
        vmin = fmin(vmin, F(F_ZV, kpc + t));
        if (M::NSAFE > 1) vmin = fmin(vmin, F(F_ZV + 1, kpc + t));
      }
      EXp()[NS * c + 2] = vmin;
    }
    }
    team_sync();
#pragma unroll 1
    BMPC_FOR_NODES(k) {
      if (k == 0) continue;
      int b, t;
      node_of(k, b, t);
      const int kp = kp_of(b, t);
      const real vmin = EXp()[NS * b + 2];
      F(F_ZS, kp) = bmpc_exp(-5.0 * (F(F_ZV, kp) - vmin));
      if (M::NSAFE > 1) F(F_ZS + 1, kp) = bmpc_exp(-5.0 * (F(F_ZV + 1, kp) - vmin));
    }
    team_sync();
    if (team_leader()) {
#pragma unroll 1
    for (int c = 1 + BMPC_LANE_ID; c < PP.nbranch; c += BMPC_BLANES) {
      const int kpc = kp_of(c, 0);
      real num = 0.0, den = 0.0;
#pragma unroll 1
      for (int t = 0; t < PP.N; ++t) {
        const real e0 = F(F_ZS, kpc + t);
        num += e0 * F(F_ZV, kpc + t);
        den += e0;
        if (M::NSAFE > 1) {
          const real e1 = F(F_ZS + 1, kpc + t);
          num += e1 * F(F_ZV + 1, kpc + t);
          den += e1;
        }
      }
      EXp()[NS * c] = bmpc_div(num, den);   // exchange slot: safety value of child c
    }
    bsync();
    // un-normalised weight of every child branch (two exponentials), lanes = children of all levels at once: a parent lane
    // that walked its children paid them one after the other, on every level
#pragma unroll 1
    for (int c = 1 + BMPC_LANE_ID; c < PP.nbranch; c += BMPC_BLANES) {
      real himax = 0.0;
      if (M::kWeightNeedsMax) {
        const int dc = bmpc_depth(PP, c);
        const int fc = bmpc_first_child(PP, bmpc_parent(PP, c, dc), dc - 1);
        himax = -1e300;
#pragma unroll 1
        for (int j = 0; j < m; ++j) himax = fmax(himax, EXp()[NS * (fc + j)]);
      }
      EXp()[NS * c + 1] = M::branch_weight(PP, EXp()[NS * c], himax);
    }
    bsync();
#pragma unroll 1
    for (int d = 0; d < PP.NB; ++d) {
      // probabilities p = softmax over the siblings, weights w = w_parent p, arg-max child (lanes = parents)
#pragma unroll 1
      for (int b = PP.off[d] + BMPC_LANE_ID; b < PP.off[d + 1]; b += BMPC_BLANES) {
        const int fc = bmpc_first_child(PP, b, d);
        real sum = 0.0;
#pragma unroll 1
        for (int j = 0; j < m; ++j) sum += EXp()[NS * (fc + j) + 1];
        int best = 0;
        real pb = -1.0;
        const real rsum = bmpc_div(1.0, sum);
#pragma unroll 1
        for (int j = 0; j < m; ++j) {
          const real p = EXp()[NS * (fc + j) + 1] * rsum;
          Wbp()[fc + j] = Wbp()[b] * p;
          if (PP.out.branch_p) PP.out.branch_p[((size_t)prob * PP.nbranch + b) * m + j] = p;
          if (PP.ctrl == BMPC_CTRL_CVAR) cvP()[b * m + j] = p;
          if (p > pb) { pb = p; best = j; }
        }
        pbest[b] = best;   // the old value was consumed by the input shift above
      }
      bsync();
    }
    }
    team_sync();
    if (PP.out.branch_w) {
      for (int b = BMPC_LANE_ID; b < PP.nbranch; b += BMPC_LANES) PP.out.branch_w[(size_t)prob * PP.nbranch + b] = Wbp()[b];
    }
    if (PP.ctrl == BMPC_CTRL_CVAR) cvar_first_weights();
    prof_end(13, prof_t2);
    const long long prof_t3 = prof_begin(12);
    node_setup_all();
    prof_end(12, prof_t3);
    rlin = 0.0;
    if (PP.ctrl != BMPC_CTRL_CVAR) {
#pragma unroll
      for (int a = 0; a < NU; ++a) rlin += -2.0 * PP.oldin[(size_t)prob * NU + a] * PP.dR[a];
    }
    prof_end(2, prof_t0);
  }

  // ----------------------------------------------------------------------------------------
  // robustMPC (MPC_branch.py:1275-1595): ONE ego chain (root + N*NB input nodes + an internal dummy stage for the
  // terminal state) against every obstacle node of the scenario tree.  Linearisation: zero-input nonlinear rollout on
  // the first solve (get_xLin :1326-1335), afterwards the previous QP solution shifted by one step (:1429-1431; note that
  // the shifted trajectory does NOT restart from the measured state).  Obstacle nodes are collected per time slot
  // t = (depth-1) N + i + 1 in BFS order (:1336-1383); slot t gives node t its collision rows.
  // ----------------------------------------------------------------------------------------
  BMPC_DN void expand_chain() {
    const long long prof_t0 = prof_begin(2);
    const int started = PP.started[prob];
    const int Nx = PP.totalu, Nu = PP.totalu - 1;   // states incl. the terminal one / real inputs
    const real* uPrev = PP.uLin + (size_t)prob * (PP.totalu + 1) * NU;
    const real* xPrev = PP.xprev + (size_t)prob * PP.pub_totalx * NXP;
    const code_t* codes = PP.code_cache + (size_t)prob * code_stride(PP.totalu);
    const real* x0 = PP.x0 + (size_t)prob * NXP;
    const real* z0 = PP.z0 + (size_t)prob * NXP;
    auto kp_chain = [&](int k) { return k == 0 ? kp_of(0, 0) : kp_of(1, k - 1); };
    // 1. linearisation trajectory, parked in the sweep vectors
    if (!started) {
      if (BMPC_LANE_ID == 0) {
        real x[NXP], xn[NXP], u[NU];
#pragma unroll
        for (int i = 0; i < NXP; ++i) x[i] = x0[i];
#pragma unroll
        for (int a = 0; a < NU; ++a) u[a] = 0.0;
        for (int k = 0; k < Nx; ++k) {
          const int kp = kp_chain(k);
#pragma unroll
          for (int i = 0; i < NXP; ++i) F(F_XQ + i, kp) = x[i];
#pragma unroll
          for (int a = 0; a < NU; ++a) F(F_UQ + a, kp) = 0.0;
          M::step(PP, x, u, xn);
#pragma unroll
          for (int i = 0; i < NXP; ++i) x[i] = xn[i];
        }
      }
    } else {
      for (int k = BMPC_LANE_ID; k < Nx; k += BMPC_LANES) {
        const int kp = kp_chain(k);
        const int kx = (k + 1 < Nx) ? k + 1 : Nx - 1;
        const int ku = (k + 1 < Nu) ? k + 1 : Nu - 1;
#pragma unroll
        for (int i = 0; i < NXP; ++i) F(F_XQ + i, kp) = xPrev[(size_t)kx * NXP + i];
#pragma unroll
        for (int a = 0; a < NU; ++a) F(F_UQ + a, kp) = (k < Nu) ? uPrev[(size_t)ku * NU + a] : 0.0;
        if (use_codes) stp()[kp] = codes[(k + 1 < Nu) ? k + 1 : k];
      }
    }
    // 2. obstacle scenario tree -> collision-row slots of the chain nodes
    if (BMPC_LANE_ID == 0) {
      const int kp = kp_chain(0);
      F(F_FC, kp) = z0[0];
      F(F_FC + 1, kp) = z0[1];
#pragma unroll
      for (int i = 0; i < NXP; ++i) EXZp()[i] = z0[i];
    }
    lanes_sync();
    const int m = PP.zm;
    for (int d = 0; d < PP.zNB; ++d) {
      const int cnt = PP.zpw[d] * m;
      for (int idx = BMPC_LANE_ID; idx < cnt; idx += BMPC_LANES) {
        const int bq = bmpc_idiv(idx, nextafterf(1.0f / (float)m, 2.0f));   // obstacle tree of the robust chain: m = zm here
        const int b = PP.zoff[d] + bq;
        const int i = idx - bq * m;
        const int c = PP.zoff[d + 1] + (b - PP.zoff[d]) * m + i;
        const int j = c - PP.zoff[d + 1];                 // position inside the slot (BFS order)
        real zl[NXP];
        M::policy_safety(PP, PP.pol_kind[i], pol_par(i), PP.pol_kind[0], pol_par(0), x0, EXZp() + NX * b, zl, PP.zN,
                         [&](int t, const real* z) {
                           const int kp = kp_chain(d * PP.zN + t + 1);
                           F(F_FC + 2 * (j < NC ? j : 0), kp) = z[0];
                           F(F_FC + 2 * (j < NC ? j : 0) + 1, kp) = z[1];
                         });
#pragma unroll
        for (int q = 0; q < NXP; ++q) EXZp()[NX * c + q] = zl[q];
      }
      lanes_sync();
    }
    // 3. per-node data: every node is linearised about its own (x, u) of the shifted trajectory - no rollout dependence
    const real* xref = PP.xref + (size_t)prob * NXP;
    for (int k = BMPC_LANE_ID; k < Nx; k += BMPC_LANES) {
      const int kp = kp_chain(k);
      real xb[NXP], ub[NU], xn[NXP], lin[M::NLIN], cc[M::NCC];
#pragma unroll
      for (int i = 0; i < NXP; ++i) xb[i] = F(F_XQ + i, kp);
#pragma unroll
      for (int a = 0; a < NU; ++a) ub[a] = F(F_UQ + a, kp);
      M::linearize(PP, xb, ub, lin, cc, xn);
#pragma unroll
      for (int i = 0; i < M::NLIN; ++i) F(F_LIN + i, kp) = lin[i];
#pragma unroll
      for (int i = 0; i < M::NCC; ++i) F(F_CC + i, kp) = cc[i];
      const real* Ql = (k == Nx - 1) ? PP.Qf : PP.Q;       // buildCost :1541-1556: q = -2 xRef' blockdiag(Q.., Qf)
#pragma unroll
      for (int jj = 0; jj < NXP; ++jj) {
        real a = 0.0;
#pragma unroll
        for (int i = 0; i < NXP; ++i) a += xref[i] * Ql[i * NXP + jj];
        F(F_Q + jj, kp) = -2.0 * a;
      }
      const int depth = (k == 0) ? 0 : (k - 1) / PP.zN + 1;
      const int ncol = (k <= PP.zN * PP.zNB) ? PP.zpw[depth] : 0;   // the terminal state has state rows only (:1470-1482)
      rows_setup(kp, xb, ncol);
#pragma unroll
      for (int a = 0; a < NU; ++a) F(F_SU + a, kp) = bmpc_clamp(ub[a], PP.ulo[a], PP.uhi[a]);
    }
    if (BMPC_LANE_ID == 0) {
      Wbp()[0] = 1.0;     // slack weights are not branch-weighted in robustMPC (:1562)
      Wbp()[1] = 1.0;
    }
    rlin = 0.0;
    lanes_sync();
    prof_end(2, prof_t0);
  }

  // ----------------------------------------------------------------------------------------
  // Belief-state MPC (PredictiveControllers.MPC, :56-340): ONE ego chain of N stages; the state of the reference's QP is
  // [x; b] with b the beliefs over the other agents' policies (HMM_backup_dyn.PredictiveModel).  The beliefs carry no cost
  // and no constraint, so they do not influence the optimal (x, u): the chain QP is solved in the physical state, and the
  // belief part of xPred is propagated afterwards through the same linearised dynamics the reference's equality rows
  // hold (belief_outputs).  What the beliefs DO decide is which rows exist: the row of (agent j, policy k) on node i+1 is
  // imposed only if the belief entry of the linearisation trajectory exceeds hmm_thres (:211-216).
  //   get_xLin (:115-127)            nonlinear rollout of [x; b] under the shifted previous inputs
  //   computeLTVdynamics (:162-166)  stage i linearised about xLin[i+1] with the backup states of step i
  //   buildIneqConstr (:195-242)     state rows on nodes 0..N-1, gated rows on nodes 1..N-1 from linearisation i+1
  // ----------------------------------------------------------------------------------------
  BMPC_D real* belX(int k) { return PP.bel + (size_t)cv_team() * PP.bel_reals + (size_t)k * 16; }   // xLin[k]: x (4), b (<= 9)
  BMPC_DN void expand_belief() {
    const long long prof_t0 = prof_begin(2);
    const int started = PP.started[prob];
    const int Nst = PP.totalu - 1;                          // stages N; chain states 0..N
    const int nb = PP.hmm_M * PP.zm;
    const real* uPrev = PP.uLin + (size_t)prob * (PP.totalu + 1) * NU;
    const code_t* codes = PP.code_cache + (size_t)prob * code_stride(PP.totalu);
    const real* x0 = PP.x0 + (size_t)prob * NXP;
    const real* b0 = PP.b0 + (size_t)prob * nb;
    const real* xbk = PP.xbackup + (size_t)prob * nb * PP.xb_cols;
    auto kp_chain = [&](int k) { return k == 0 ? kp_of(0, 0) : kp_of(1, k - 1); };
    // uLin[i] of this call = previous uPred[min(i + 1, N - 1)] (:156-158, :118), zero on the first call
    auto ulin = [&](int i, real* u) {
      const int src = (i + 1 < Nst) ? i + 1 : Nst - 1;
#pragma unroll
      for (int a = 0; a < NU; ++a) u[a] = started ? uPrev[(size_t)src * NU + a] : 0.0;
    };
    if (BMPC_LANE_ID == 0) {
      real x[NXP], b[9], bn[9], u[NU], xn[NXP];
#pragma unroll
      for (int i = 0; i < NXP; ++i) x[i] = x0[i];
      // the rollout reads the belief column-major (np.reshape(b0, -1, 1) with the legacy integer order, :121)
      for (int i = 0; i < PP.hmm_M; ++i)
        for (int j = 0; j < PP.zm; ++j) b[j * PP.hmm_M + i] = b0[i * PP.zm + j];
      for (int k = 0; k <= Nst; ++k) {
        real* o = belX(k);
#pragma unroll
        for (int i = 0; i < NXP; ++i) o[i] = x[i];
        for (int q = 0; q < nb; ++q) o[4 + q] = b[q];
        if (k == Nst) break;
        ulin(k, u);
        BeliefModel::transition(PP, x, xbk, PP.xb_cols, k, b, bn, nullptr, nullptr);
        M::step(PP, x, u, xn);
#pragma unroll
        for (int i = 0; i < NXP; ++i) x[i] = xn[i];
        for (int q = 0; q < nb; ++q) b[q] = bn[q];
      }
    }
    team_sync();
    const real* xref = PP.xref + (size_t)prob * NXP;
#pragma unroll 1
    for (int k = BMPC_LANE_ID; k <= Nst; k += BMPC_LANES) {
      const int kp = kp_chain(k);
      const int kl = (k < Nst) ? k + 1 : Nst;               // stage k is linearised about xLin[k+1]
      real xb[NXP], ub[NU], xn[NXP], lin[M::NLIN], cc[M::NCC];
#pragma unroll
      for (int i = 0; i < NXP; ++i) xb[i] = belX(kl)[i];
      ulin(kl, ub);
      M::linearize(PP, xb, ub, lin, cc, xn);
#pragma unroll
      for (int i = 0; i < M::NLIN; ++i) F(F_LIN + i, kp) = lin[i];
#pragma unroll
      for (int i = 0; i < M::NCC; ++i) F(F_CC + i, kp) = cc[i];
      const real* Ql = (k == Nst) ? PP.Qf : PP.Q;           // buildCost :271-280: q = -2 xRef' blockdiag(Q.., Qf)
#pragma unroll
      for (int jj = 0; jj < NXP; ++jj) {
        real a = 0.0;
#pragma unroll
        for (int i = 0; i < NXP; ++i) a += xref[i] * Ql[i * NXP + jj];
        F(F_Q + jj, kp) = -2.0 * a;
      }
      // gated rows of node k (1 <= k <= N-1): belief of xLin[k] in the row-major view, h0 / Jh of linearisation k
      // (about xLin[k+1], backup states of step k)
      int slot = 0;
      const real* xk = belX(k);        // ADMM start: rows evaluated at the node's own linearisation state
      if (k >= 1 && k <= Nst - 1) {
        for (int j = 0; j < PP.hmm_M; ++j)
          for (int q = 0; q < PP.zm; ++q) {
            if (belX(k)[4 + j * PP.zm + q] > PP.hmm_thres && slot < NC) {
              real gx, gy;
              const real h = BeliefModel::safety(PP, xb, xbk + (size_t)(PP.zm * j + q) * PP.xb_cols + 4 * k, gx, gy);
              const real hi0 = h - (gx * xb[0] + gy * xb[1]);
              F(F_FC + 2 * slot, kp) = -gx;
              F(F_FC + 2 * slot + 1, kp) = -gy;
              F(F_HC + slot, kp) = hi0;
              F(F_S + slot, kp) = fmin(-gx * xk[0] - gy * xk[1], hi0);
              ++slot;
            }
          }
      }
      for (; slot < NC; ++slot) {
        F(F_FC + 2 * slot, kp) = 0.0;       // empty slot: a zero row gets rho = 0 and is ignored
        F(F_FC + 2 * slot + 1, kp) = 0.0;
        F(F_HC + slot, kp) = 1.0e30;
        F(F_S + slot, kp) = 0.0;
      }
#pragma unroll
      for (int j = NC; j < NR; ++j) {
        real v = 0.0;
#pragma unroll
        for (int i = 0; i < NXP; ++i) v += RF(j - NC, i) * xk[i];
        F(F_S + j, kp) = bmpc_clamp(v, RLO(j - NC), RHI(j - NC));
      }
      real uk[NU];
      ulin(k, uk);
#pragma unroll
      for (int a = 0; a < NU; ++a) F(F_SU + a, kp) = bmpc_clamp(uk[a], PP.ulo[a], PP.uhi[a]);
      if (use_codes) stp()[kp] = codes[(k + 1 < Nst) ? k + 1 : k];
    }
    if (BMPC_LANE_ID == 0) {
      Wbp()[0] = 1.0;     // slack weights are not weighted (:287-289)
      Wbp()[1] = 1.0;
    }
    rlin = 0.0;
    team_sync();
    prof_end(2, prof_t0);
  }

  // belief part of xPred: b_0 = b0 in the row-major flattening the equality rows use (:147), then
  // b_(i+1) = C_b + (d b+/d(x,y)) x_i + H' b_i with the Jacobians of linearisation i (about xLin[i+1], backup states of step i)
  BMPC_DN void belief_outputs() {
    if (!PP.out.bPred) return;
    if (BMPC_LANE_ID == 0) {
      const int Nst = PP.totalu - 1, nb = PP.hmm_M * PP.zm, Mh = PP.hmm_M, m = PP.zm;
      const real* b0 = PP.b0 + (size_t)prob * nb;
      const real* xbk = PP.xbackup + (size_t)prob * nb * PP.xb_cols;
      real* out = PP.out.bPred + (size_t)prob * PP.pub_totalx * nb;
      auto kp_chain = [&](int k) { return k == 0 ? kp_of(0, 0) : kp_of(1, k - 1); };
      real b[9], bn[9], bl[9], dbx[18], Hm[36];
      for (int q = 0; q < nb; ++q) { b[q] = b0[q]; out[q] = b[q]; }
      for (int i = 0; i < Nst; ++i) {
        const real* xl = belX(i + 1);
        BeliefModel::transition(PP, xl, xbk, PP.xb_cols, i, xl + 4, bl, dbx, Hm);
        const real x0s = F(F_XQ, kp_chain(i)), x1s = F(F_XQ + 1, kp_chain(i));
        for (int q = 0; q < nb; ++q) {
          const int ag = q % Mh, k = q / Mh;               // column-major position -> (agent, policy)
          real v = bl[q] + dbx[2 * q] * (x0s - xl[0]) + dbx[2 * q + 1] * (x1s - xl[1]);
          for (int r = 0; r < m; ++r) v += Hm[(ag * m + r) * m + k] * (b[r * Mh + ag] - xl[4 + r * Mh + ag]);
          bn[q] = v;
        }
        for (int q = 0; q < nb; ++q) { b[q] = bn[q]; out[(size_t)(i + 1) * nb + q] = bn[q]; }
      }
    }
    team_sync();
  }

  // ========================================================================================
  // Riccati factorisation over the tree
  // ========================================================================================
  BMPC_D void unpack_sym(const real* src, real* Pm) {
    int q = 0;
#pragma unroll
    for (int i = 0; i < NX; ++i)
#pragma unroll
      for (int j = i; j < NX; ++j) {
        Pm[i * NX + j] = src[q];
        Pm[j * NX + i] = src[q];
        ++q;
      }
  }
  BMPC_D void pack_sym(const real* Pm, real* dst) {
    int q = 0;
#pragma unroll
    for (int i = 0; i < NX; ++i)
#pragma unroll
      for (int j = i; j < NX; ++j) dst[q++] = Pm[i * NX + j];
  }

  BMPC_D static void invert_spd(const real* S, real* Si) {
    if constexpr (NU == 2) {
      const real det = S[0] * S[3] - S[1] * S[2];
      const real id = bmpc_div(1.0, det);
      Si[0] = S[3] * id;
      Si[1] = -S[1] * id;
      Si[2] = -S[2] * id;
      Si[3] = S[0] * id;
    } else {
      // 3x3 symmetric: cofactors
      const real a = S[0], b = S[1], c = S[2], d = S[4], e = S[5], f = S[8];
      const real A = d * f - e * e, B = c * e - b * f, C = b * e - c * d;
      const real det = a * A + b * B + c * C;
      const real id = bmpc_div(1.0, det);
      Si[0] = A * id;
      Si[1] = B * id;
      Si[2] = C * id;
      Si[3] = B * id;
      Si[4] = (a * f - c * c) * id;
      Si[5] = (b * c - a * e) * id;
      Si[6] = C * id;
      Si[7] = (b * c - a * e) * id;
      Si[8] = (a * d - b * b) * id;
    }
  }

  // fixed-input helpers of the polish: value of input a when the active-set code pins it to a bound, else 0
  BMPC_D real pinned_value(code_t code, int a) {
    const int ca = in_of(code, a);
    return ca == IN_AT_HI ? PP.uhi[a] : (ca == IN_AT_LO ? PP.ulo[a] : 0.0);
  }
  BMPC_D bool pinned(code_t code, int a) { return in_of(code, a) != IN_FREE; }

  // Stiff penalty of a guessed-active row in the polish: polish_mult times the row's curvature-matched stiffness
  // (rho/theta = 1/(f' Sigma f)), at least polish_big*w, so that every augmented-Lagrangian step contracts strongly.
  BMPC_D real big_row(int kp, int j, real w) {
    return fmin(fmax(PP.polish_big * w, PP.polish_mult * F(F_RHO + j, kp) * PP.inv_theta), 1.0e12 * w);
  }
  // penalties of the node's soft rows and inputs for the requested factorisation
  BMPC_D void penalties(int kp, real w, int mode, real* pr, real* pu) {
    if (mode == FACT_ADMM) {
#pragma unroll
      for (int j = 0; j < NR; ++j) pr[j] = F(F_RHO + j, kp);
#pragma unroll
      for (int a = 0; a < NU; ++a) pu[a] = F(F_RHO + NR + a, kp);
    } else if (mode == FACT_IPM) {
#pragma unroll
      for (int j = 0; j < NR; ++j) pr[j] = IPF(IP_KAP + j, kp);
#pragma unroll
      for (int a = 0; a < NU; ++a) pu[a] = IPF(IP_KAP + NR + a, kp);
    } else if (mode == FACT_FREE) {
#pragma unroll
      for (int j = 0; j < NR; ++j) pr[j] = 0.0;
#pragma unroll
      for (int a = 0; a < NU; ++a) pu[a] = 0.0;
    } else {
      const code_t code = stp()[kp];
#pragma unroll
      for (int j = 0; j < NR; ++j) {
        const int cj = row_of(code, j);
        pr[j] = (cj == ROW_UP_KINK || cj == ROW_LO_KINK) ? big_row(kp, j, w) : 0.0;
      }
#pragma unroll
      for (int a = 0; a < NU; ++a) {
        const int ca = in_of(code, a);
        pu[a] = 0.0;   // inputs at a bound are eliminated exactly (see node_factor), not penalised
      }
    }
  }

  // one backward Riccati step; Pn (full NX x NX, symmetric) is the successor value Hessian on entry
  // and this node's on exit
  // rate = weight of the input-rate pair (previous input, this input), sig = 0 where the reference drops the pair's
  // own-input term (leaf last node, MPC_branch.py:303), root = the root-input quirks apply (:311-312)
  BMPC_D void node_factor(int kp, real w, real* Pn, int mode, real rate, real sig, bool root, bool use_qf = false) {
    const real* Qs = use_qf ? PP.Qf : QH();   // robustMPC: the dummy stage carries the terminal cost Qf
    real lin[M::NLIN], cc[M::NCC];
#pragma unroll
    for (int i = 0; i < M::NLIN; ++i) lin[i] = F(F_LIN + i, kp);
#pragma unroll
    for (int i = 0; i < M::NCC; ++i) cc[i] = F(F_CC + i, kp);
    real pr[NR], pu[NU];
    penalties(kp, w, mode, pr, pu);
    // h0 = P+ C  (polish: C + B u_pinned, the inputs held at their bounds act as a known offset of the dynamics)
    const code_t pcode = (mode == FACT_POLISH) ? stp()[kp] : 0;
    {
      real C[NX];
      M::expandC(cc, C);
      if (mode == FACT_POLISH) {
        real up[NU];
#pragma unroll
        for (int a = 0; a < NU; ++a) up[a] = pinned_value(pcode, a);
        M::addBu(PP, lin, up, C);
      }
#pragma unroll
      for (int i = 0; i < NX; ++i) {
        real a = 0.0;
#pragma unroll
        for (int j = 0; j < NX; ++j) a += Pn[i * NX + j] * C[j];
        F(F_H0 + i, kp) = a;
      }
    }
    real T[NX * NX];  // P+ A  (row i = A' applied to row i of P+)
#pragma unroll
    for (int i = 0; i < NX; ++i) M::mulAT(PP, lin, Pn + i * NX, T + i * NX);
    real G[NU * NX];  // B' P+ A
#pragma unroll
    for (int j = 0; j < NX; ++j) {
      real col[NX], g2[NU];
#pragma unroll
      for (int i = 0; i < NX; ++i) col[i] = T[i * NX + j];
      M::mulBT(PP, lin, col, g2);
#pragma unroll
      for (int a = 0; a < NU; ++a) G[a * NX + j] = g2[a];
    }
    real S[NU * NU];
    {
      real PB[NX * NU];  // P+ B (row i = B' applied to row i of P+)
#pragma unroll
      for (int i = 0; i < NX; ++i) M::mulBT(PP, lin, Pn + i * NX, PB + i * NU);
#pragma unroll
      for (int b2 = 0; b2 < NU; ++b2) {
        real col[NX], g2[NU];
#pragma unroll
        for (int i = 0; i < NX; ++i) col[i] = PB[i * NU + b2];
        M::mulBT(PP, lin, col, g2);
#pragma unroll
        for (int a = 0; a < NU; ++a) S[a * NU + b2] = g2[a] + w * (PP.R[a * NU + b2] + PP.R[b2 * NU + a]);
      }
#pragma unroll
      for (int a = 0; a < NU; ++a) S[a * NU + a] += pu[a];
      if (RATE) {
#pragma unroll
        for (int a = 0; a < NU; ++a) S[a * NU + a] += 2.0 * rate * sig * PP.dR[a];
        if (root) {
          // Hu[0:d,0:d] += dR broadcasts the vector over the rows (:312); OSQP keeps the upper triangle
#pragma unroll
          for (int a = 0; a < NU; ++a)
#pragma unroll
            for (int b2 = 0; b2 < NU; ++b2) S[a * NU + b2] += 2.0 * PP.dR[a > b2 ? a : b2];
        }
#pragma unroll
        for (int a = 0; a < NU; ++a) G[a * NX + NXP + a] += -2.0 * rate * PP.dR[a];   // cross term previous input x input
      }
    }
    real Si[NU * NU];
    if (mode == FACT_POLISH) {
      // inverse on the free inputs only: pinned rows/columns are cut out (K and kff vanish there)
#pragma unroll
      for (int a = 0; a < NU; ++a) {
        if (pinned(pcode, a)) {
#pragma unroll
          for (int b2 = 0; b2 < NU; ++b2) { S[a * NU + b2] = 0.0; S[b2 * NU + a] = 0.0; }
          S[a * NU + a] = 1.0;
        }
      }
    }
    invert_spd(S, Si);
    if (mode == FACT_POLISH) {
#pragma unroll
      for (int a = 0; a < NU; ++a)
        if (pinned(pcode, a)) Si[a * NU + a] = 0.0;
    }
    {
      int q = 0;
#pragma unroll
      for (int a = 0; a < NU; ++a)
#pragma unroll
        for (int b2 = a; b2 < NU; ++b2) F(F_SI + (q++), kp) = 0.5 * (Si[a * NU + b2] + Si[b2 * NU + a]);
    }
    real K[NU * NX];
#pragma unroll
    for (int a = 0; a < NU; ++a)
#pragma unroll
      for (int j = 0; j < NX; ++j) {
        real v = 0.0;
#pragma unroll
        for (int b2 = 0; b2 < NU; ++b2) v += Si[a * NU + b2] * G[b2 * NX + j];
        K[a * NX + j] = v;
        F(F_K + a * NX + j, kp) = v;
      }
    if constexpr (NX > 4) {
    // P = Q~ + A' T - G' K, upper triangle only (the unused entries of every product below are dead code after unrolling)
      real Pnew[NX * NX];
      const real qs = w * (1.0 + PP.dq_scale);
  #pragma unroll
      for (int j = 0; j < NX; ++j) {
        real col[NX], o[NX];
  #pragma unroll
        for (int i = 0; i < NX; ++i) col[i] = T[i * NX + j];
        M::mulAT(PP, lin, col, o);
  #pragma unroll
        for (int i = 0; i <= j; ++i) {
          real v = o[i];
          if (i < NXP && j < NXP) v += qs * (Qs[i * NXP + j] + Qs[j * NXP + i]);
          if (RATE && i == j && i >= NXP) v += 2.0 * rate * PP.dR[i - NXP];
  #pragma unroll
          for (int a = 0; a < NU; ++a) v -= G[a * NX + i] * K[a * NX + j];
          Pnew[i * NX + j] = v;
        }
      }
      // soft-row penalties: the collision row touches (x,y) only; a state row with a single entry (the usual box on one
      // state, flagged by rf_one) touches one diagonal element
      {
  #pragma unroll
        for (int j = 0; j < NC; ++j) {
          const real fx = F(F_FC + 2 * j, kp), fy = F(F_FC + 2 * j + 1, kp);
          Pnew[0] += pr[j] * fx * fx;
          Pnew[1] += pr[j] * fx * fy;
          Pnew[NX + 1] += pr[j] * fy * fy;
        }
  #pragma unroll
        for (int j = NC; j < NR; ++j) {
          const int one = RF_ONE(j - NC);
          if (one >= 0) {
  #pragma unroll
            for (int i = 0; i < NXP; ++i)
              if (i == one) Pnew[i * NX + i] += pr[j] * RF(j - NC, i) * RF(j - NC, i);
          } else {
  #pragma unroll
            for (int i = 0; i < NXP; ++i)
  #pragma unroll
              for (int i2 = i; i2 < NXP; ++i2) Pnew[i * NX + i2] += pr[j] * RF(j - NC, i) * RF(j - NC, i2);
          }
        }
      }
  #pragma unroll
      for (int i = 0; i < NX; ++i)
  #pragma unroll
        for (int j = i; j < NX; ++j) {
          Pn[i * NX + j] = Pnew[i * NX + j];
          Pn[j * NX + i] = Pnew[i * NX + j];
        }
      return;
    }
    // P = Q~ + A' T - G' K (full matrix, symmetrised at the end: for the 3- and 4-state models this form compiles to
    // faster code than the triangle-only one above, measured: highway -3.5 %, rate-augmented quadruped +5 %)
    real Pnew[NX * NX];
    const real qs = w * (1.0 + PP.dq_scale);
#pragma unroll
    for (int j = 0; j < NX; ++j) {
      real col[NX], o[NX];
#pragma unroll
      for (int i = 0; i < NX; ++i) col[i] = T[i * NX + j];
      M::mulAT(PP, lin, col, o);
#pragma unroll
      for (int i = 0; i < NX; ++i) {
        real v = o[i];
        if (i < NXP && j < NXP) v += qs * (Qs[i * NXP + j] + Qs[j * NXP + i]);
        if (RATE && i == j && i >= NXP) v += 2.0 * rate * PP.dR[i - NXP];
#pragma unroll
        for (int a = 0; a < NU; ++a) v -= G[a * NX + i] * K[a * NX + j];
        Pnew[i * NX + j] = v;
      }
    }
    // soft-row penalties: the collision row touches (x,y) only
    {
#pragma unroll
      for (int j = 0; j < NC; ++j) {
        const real fx = F(F_FC + 2 * j, kp), fy = F(F_FC + 2 * j + 1, kp);
        Pnew[0] += pr[j] * fx * fx;
        Pnew[1] += pr[j] * fx * fy;
        Pnew[NX] += pr[j] * fx * fy;
        Pnew[NX + 1] += pr[j] * fy * fy;
      }
#pragma unroll
      for (int j = NC; j < NR; ++j)
#pragma unroll
        for (int i = 0; i < NXP; ++i)
#pragma unroll
          for (int i2 = 0; i2 < NXP; ++i2) Pnew[i * NX + i2] += pr[j] * RF(j - NC, i) * RF(j - NC, i2);
    }
#pragma unroll
    for (int i = 0; i < NX; ++i)
#pragma unroll
      for (int j = 0; j < NX; ++j) Pn[i * NX + j] = 0.5 * (Pnew[i * NX + j] + Pnew[j * NX + i]);
  }

  BMPC_D void sum_children(int b, int d, real* Pn) {
    const int fc = bmpc_first_child(PP, b, d);
    unpack_sym(EXp() + NS * fc, Pn);
#pragma unroll 1
    for (int c = 1; c < PP.m; ++c) {
      real Pc[NX * NX];
      unpack_sym(EXp() + NS * (fc + c), Pc);
#pragma unroll
      for (int i = 0; i < NX * NX; ++i) Pn[i] += Pc[i];
    }
  }

  // The root is "depth 0, one node": every sweep runs one loop nest over the levels, so that each node step is
  // instantiated once (code size matters: the instruction cache is the first bottleneck of this kernel, profiles/).
  BMPC_DN void factorize_sweep(int mode) {
    const long long prof_t0 = prof_begin(4);
#pragma unroll 1
    for (int d = PP.NB; d >= 0; --d) {
      const int nt = (d == 0) ? 1 : PP.N;
#pragma unroll 1
      for (int b = PP.off[d] + BMPC_LANE_ID; b < PP.off[d + 1]; b += BMPC_BLANES) {
        const real w = Wbp()[b];
        real Pn[NX * NX];
        if (d == PP.NB) {
          // terminal node: x' (w Qf) x in the reference's H, doubled by H <- 2H (:1094, :1112)
#pragma unroll
          for (int i = 0; i < NX; ++i)
#pragma unroll
            for (int j = 0; j < NX; ++j)
              Pn[i * NX + j] = (i < NXP && j < NXP && !bmpc_is_chain(PP.ctrl) && PP.ctrl != BMPC_CTRL_CVAR)
                                   ? w * (PP.Qf[i * NXP + j] + PP.Qf[j * NXP + i]) : 0.0;   // no terminal cost in robustMPC / BranchMPC_CVaR
        } else {
          sum_children(b, d, Pn);
        }
#pragma unroll 1
        for (int t = nt - 1; t >= 0; --t)
          node_factor(kp_of(b, t), w, Pn, mode, (d == 0) ? 0.0 : w, (d == PP.NB && t == nt - 1) ? 0.0 : 1.0, d == 0,
                      bmpc_is_chain(PP.ctrl) && d == PP.NB && t == nt - 1);
        pack_sym(Pn, EXp() + NS * b);
      }
      bsync();
    }
    prof_end(4, prof_t0);
  }
  BMPC_D void factorize(int mode) {
    if (team_leader()) factorize_sweep(mode);
    team_sync();
  }

  // ========================================================================================
  // Curvature-matched rho: rho_row = theta / (f' Sigma f), rho_u = theta_u / (Sigma_u)_aa where Sigma is the
  // covariance-like forward recursion of the unconstrained LQ problem (Sigma+ = Acl Sigma Acl' + B S^-1 B'),
  // i.e. the diagonal of F H^-1 F' restricted to the dynamics: a Jacobi preconditioner of the ADMM dual problem.
  // Also scales the ADMM start written by node_setup (sh = rho * v).
  // ========================================================================================
  BMPC_D void node_cov(int kp, real w, real* Sg) {
    real lin[M::NLIN], K[NU * NX], Si[NU * NU];
#pragma unroll
    for (int i = 0; i < M::NLIN; ++i) lin[i] = F(F_LIN + i, kp);
#pragma unroll
    for (int i = 0; i < NU * NX; ++i) K[i] = F(F_K + i, kp);
    {
      int q = 0;
#pragma unroll
      for (int a = 0; a < NU; ++a)
#pragma unroll
        for (int b2 = a; b2 < NU; ++b2) {
          const real v = F(F_SI + (q++), kp);
          Si[a * NU + b2] = v;
          Si[b2 * NU + a] = v;
        }
    }
    const real rho_max = 1.0e9 * w;
    {
#pragma unroll
      for (int j = 0; j < NC; ++j) {
        const real fx = F(F_FC + 2 * j, kp), fy = F(F_FC + 2 * j + 1, kp);
        const real q0 = fx * fx * Sg[0] + 2.0 * fx * fy * Sg[1] + fy * fy * Sg[NX + 1];
        const real r0 = (q0 > 1e-12) ? fmin(PP.theta / q0, rho_max) : 0.0;
        F(F_RHO + j, kp) = r0;
        F(F_S + j, kp) *= r0;
      }
      const bool no_state_rows = PP.ctrl == BMPC_CTRL_BELIEF && kp == kp_of(1, PP.N - 1);   // last state: unconstrained (:199)
#pragma unroll
      for (int j = NC; j < NR; ++j) {
        real q = 0.0;
#pragma unroll
        for (int i = 0; i < NXP; ++i)
#pragma unroll
          for (int i2 = 0; i2 < NXP; ++i2) q += RF(j - NC, i) * RF(j - NC, i2) * Sg[i * NX + i2];
        const real rj = (q > 1e-12 && !no_state_rows) ? fmin(PP.theta / q, rho_max) : 0.0;
        F(F_RHO + j, kp) = rj;
        F(F_S + j, kp) *= rj;
      }
    }
#pragma unroll
    for (int a = 0; a < NU; ++a) {
      real var = Si[a * NU + a];
#pragma unroll
      for (int i = 0; i < NX; ++i) {
        real ks = 0.0;
#pragma unroll
        for (int i2 = 0; i2 < NX; ++i2) ks += K[a * NX + i2] * Sg[i2 * NX + i];
        var += PP.rho_u_feedback * ks * K[a * NX + i];
      }
      const real ra = fmin(PP.theta_u / var, rho_max);
      F(F_RHO + NR + a, kp) = ra;
      F(F_SU + a, kp) *= ra;
    }
    // Sigma+ = Acl Sigma Acl' + B Si B',  Acl v = A v - B (K v)
    real T[NX * NX];
#pragma unroll
    for (int i = 0; i < NX; ++i) {
      real kv[NU];
#pragma unroll
      for (int a = 0; a < NU; ++a) {
        real v = 0.0;
#pragma unroll
        for (int j = 0; j < NX; ++j) v += K[a * NX + j] * Sg[i * NX + j];
        kv[a] = -v;
      }
      M::mulA(PP, lin, Sg + i * NX, T + i * NX);
      M::addBu(PP, lin, kv, T + i * NX);
    }
    real Bd[NX * NU];
    M::denseB(PP, lin, Bd);
    real Sn[NX * NX];
#pragma unroll
    for (int j = 0; j < NX; ++j) {
      real col[NX], o[NX], kv[NU];
#pragma unroll
      for (int i = 0; i < NX; ++i) col[i] = T[i * NX + j];
#pragma unroll
      for (int a = 0; a < NU; ++a) {
        real v = 0.0;
#pragma unroll
        for (int i = 0; i < NX; ++i) v += K[a * NX + i] * col[i];
        kv[a] = -v;
      }
      M::mulA(PP, lin, col, o);
      M::addBu(PP, lin, kv, o);
#pragma unroll
      for (int i = 0; i < NX; ++i) Sn[i * NX + j] = o[i];
    }
#pragma unroll
    for (int i = 0; i < NX; ++i)
#pragma unroll
      for (int j = 0; j < NX; ++j) {
        real v = 0.0;
#pragma unroll
        for (int a = 0; a < NU; ++a)
#pragma unroll
          for (int b2 = 0; b2 < NU; ++b2) v += Bd[i * NU + a] * Si[a * NU + b2] * Bd[j * NU + b2];
        Sn[i * NX + j] += v;
      }
#pragma unroll
    for (int i = 0; i < NX; ++i)
#pragma unroll
      for (int j = 0; j < NX; ++j) Sg[i * NX + j] = 0.5 * (Sn[i * NX + j] + Sn[j * NX + i]);
  }

  BMPC_DN void choose_rho_sweep() {
    const long long prof_t0 = prof_begin(3);
#pragma unroll 1
    for (int d = 0; d <= PP.NB; ++d) {
      const int nt = (d == 0) ? 1 : PP.N;
#pragma unroll 1
      for (int b = PP.off[d] + BMPC_LANE_ID; b < PP.off[d + 1]; b += BMPC_BLANES) {
        real Sg[NX * NX];
        if (d == 0) {
#pragma unroll
          for (int i = 0; i < NX * NX; ++i) Sg[i] = 0.0;
        } else {
          unpack_sym(EXp() + NS * bmpc_parent(PP, b, d), Sg);
        }
        const real w = Wbp()[b];
#pragma unroll 1
        for (int t = 0; t < nt; ++t) node_cov(kp_of(b, t), w, Sg);
        pack_sym(Sg, EXp() + NS * b);
      }
      bsync();
    }
    prof_end(3, prof_t0);
  }
  BMPC_D void choose_rho() {
    if (team_leader()) choose_rho_sweep();
    team_sync();
  }

  // ========================================================================================
  // Vector sweeps (the hot loop): backward for the feed-forward terms, forward for (x,u)
  // ========================================================================================
  BMPC_D void bw_step(int kp, real* pn) {
    real lin[M::NLIN], g[NX], r[NU], kff[NU];
#pragma unroll
    for (int i = 0; i < M::NLIN; ++i) lin[i] = F(F_LIN + i, kp);
#pragma unroll
    for (int i = 0; i < NX; ++i) g[i] = pn[i] + F(F_H0 + i, kp);
    M::mulBT(PP, lin, g, r);
#pragma unroll
    for (int a = 0; a < NU; ++a) r[a] += F(F_UQ + a, kp);
    {
      real Si[NSU];
#pragma unroll
      for (int q = 0; q < NSU; ++q) Si[q] = F(F_SI + q, kp);
      if constexpr (NU == 2) {
        kff[0] = -(Si[0] * r[0] + Si[1] * r[1]);
        kff[1] = -(Si[1] * r[0] + Si[2] * r[1]);
      } else {
        kff[0] = -(Si[0] * r[0] + Si[1] * r[1] + Si[2] * r[2]);
        kff[1] = -(Si[1] * r[0] + Si[3] * r[1] + Si[4] * r[2]);
        kff[2] = -(Si[2] * r[0] + Si[4] * r[1] + Si[5] * r[2]);
      }
    }
#pragma unroll
    for (int a = 0; a < NU; ++a) F(F_UQ + a, kp) = kff[a];
    real p[NX];
    M::mulAT(PP, lin, g, p);
#pragma unroll
    for (int i = 0; i < NX; ++i) {
      real v = p[i] + F(F_XQ + i, kp);
#pragma unroll
      for (int a = 0; a < NU; ++a) v -= F(F_K + a * NX + i, kp) * r[a];
      pn[i] = v;
    }
  }

  BMPC_DN void backward_sweep() {
    const long long prof_t0 = prof_begin(5);
#pragma unroll 1
    for (int d = PP.NB; d >= 0; --d) {
      const int nt = (d == 0) ? 1 : PP.N;
#pragma unroll 1
      for (int b = PP.off[d] + BMPC_LANE_ID; b < PP.off[d + 1]; b += BMPC_BLANES) {
        real pn[NX];
        if (d == PP.NB) {
#pragma unroll
          for (int i = 0; i < NX; ++i) pn[i] = 0.0;   // BranchMPC terminal node: no linear term (MPC_branch.py:1094)
          if (PP.ctrl == BMPC_CTRL_PROX) {
            // BranchMPCProx: -2 w xRef' Qf (:308)
            const real* xref = PP.xref + (size_t)prob * NXP;
            const real w = Wbp()[b];
#pragma unroll
            for (int j = 0; j < NXP; ++j) {
              real a = 0.0;
#pragma unroll
              for (int i = 0; i < NXP; ++i) a += xref[i] * PP.Qf[i * NXP + j];
              pn[j] = -2.0 * w * a;
            }
          }
        } else {
          const int fc = bmpc_first_child(PP, b, d);
#pragma unroll
          for (int i = 0; i < NX; ++i) pn[i] = EXp()[NS * fc + i];
#pragma unroll 1
          for (int c = 1; c < PP.m; ++c)
#pragma unroll
            for (int i = 0; i < NX; ++i) pn[i] += EXp()[NS * (fc + c) + i];
        }
        const int kp0 = kp_of(b, 0);
#pragma unroll 1
        for (int t = nt - 1; t >= 0; --t) bw_step(kp0 + t, pn);
#pragma unroll
        for (int i = 0; i < NX; ++i) EXp()[NS * b + i] = pn[i];
      }
      bsync();
    }
    prof_end(5, prof_t0);
  }
  BMPC_D void backward() {
    if (team_leader()) backward_sweep();
    team_sync();
  }

  BMPC_D void fw_step(int kp, real* x) {
    real lin[M::NLIN], cc[M::NCC], u[NU], xn[NX];
#pragma unroll
    for (int i = 0; i < M::NLIN; ++i) lin[i] = F(F_LIN + i, kp);
#pragma unroll
    for (int i = 0; i < M::NCC; ++i) cc[i] = F(F_CC + i, kp);
#pragma unroll
    for (int a = 0; a < NU; ++a) {
      real v = F(F_UQ + a, kp);
#pragma unroll
      for (int i = 0; i < NX; ++i) v -= F(F_K + a * NX + i, kp) * x[i];
      u[a] = v;
      F(F_UQ + a, kp) = v;
    }
#pragma unroll
    for (int i = 0; i < NX; ++i) F(F_XQ + i, kp) = x[i];
    M::mulA(PP, lin, x, xn);
    M::addBu(PP, lin, u, xn);
    M::addC(cc, xn);
#pragma unroll
    for (int i = 0; i < NX; ++i) x[i] = xn[i];
  }

  BMPC_DN void forward_sweep() {
    const long long prof_t0 = prof_begin(5);
#pragma unroll 1
    for (int d = 0; d <= PP.NB; ++d) {
      const int nt = (d == 0) ? 1 : PP.N;
#pragma unroll 1
      for (int b = PP.off[d] + BMPC_LANE_ID; b < PP.off[d + 1]; b += BMPC_BLANES) {
        real x[NX];
        if (d == 0) {
#pragma unroll
          for (int i = 0; i < NX; ++i) x[i] = (i < NXP) ? PP.x0[(size_t)prob * NXP + (i < NXP ? i : 0)] : 0.0;
        } else {
          const int pa = bmpc_parent(PP, b, d);
#pragma unroll
          for (int i = 0; i < NX; ++i) x[i] = EXXp()[NX * pa + i];
        }
        const int kp0 = kp_of(b, 0);
#pragma unroll 1
        for (int t = 0; t < nt; ++t) fw_step(kp0 + t, x);
#pragma unroll
        for (int i = 0; i < NX; ++i) EXXp()[NX * b + i] = x[i];
      }
      bsync();
    }
    prof_end(5, prof_t0);
  }
  BMPC_D void forward() {
    ++nsolve;
    if (team_leader()) forward_sweep();
    team_sync();
  }
  // one KKT solve without anything between the two sweeps (ADMM, interior point): the team meets once
  BMPC_D void kkt_solve() {
    ++nsolve;
    if (team_leader()) {
      backward_sweep();
      forward_sweep();
    }
    team_sync();
  }

  // ========================================================================================
  // Row phase (node-parallel): ADMM z/y update through the scaled Moreau variable, residuals, and
  // the linear terms of the next KKT solve.
  // ========================================================================================
  BMPC_D real row_value(int kp, int j, const real* x) {
    if (j < NC) return F(F_FC + 2 * j, kp) * x[0] + F(F_FC + 2 * j + 1, kp) * x[1];
    real v = 0.0;
#pragma unroll
    for (int i = 0; i < NXP; ++i) v += RF(j - NC, i) * x[i];
    return v;
  }
  BMPC_D void row_bounds(int kp, int j, real& lo, real& hi) {
    if (j < NC) { lo = BMPC_NOLO; hi = F(F_HC + j, kp); } else { lo = RLO(j - NC); hi = RHI(j - NC); }
  }
  BMPC_D void add_row_grad(int kp, int j, real gcoef, real* qx) {
    if (j < NC) {
      qx[0] += F(F_FC + 2 * j, kp) * gcoef;
      qx[1] += F(F_FC + 2 * j + 1, kp) * gcoef;
    } else {
#pragma unroll
      for (int i = 0; i < NXP; ++i) qx[i] += RF(j - NC, i) * gcoef;
    }
  }

  // update=false: only assemble q~ from the current state (start / after a failed polish)
  // check=true : also return max(primal residual, scaled dual residual) over this lane's nodes
  // (measured: one out-of-line copy with runtime switches shrinks the kernel's main body from 45 KB to 13 KB but is 8 %
  // slower - the specialised no-check variant is the hot one and must stay compact)
  template <bool UPDATE, bool CHECK>
  BMPC_D real admm_rows() {
    const long long prof_t0 = prof_begin(8);
    real res = 0.0;
    if (CHECK) { gap_r = 0.0; stp_r = 0.0; gap_u = 0.0; stp_u = 0.0; set_changes = 0; }
#pragma unroll 1
    BMPC_FOR_NODES(k) {
      int b, t;
      node_of(k, b, t);
      const int kp = kp_of(b, t);
      const real w = Wbp()[b];
      const real lam = PP.lam_lin * w;
      code_t ncode = 0;   // CHECK: the active set the ADMM state currently implies (compared with the previous check)
      const real iw100 = CHECK ? bmpc_div(0.01, w) : 0.0;
      (void)iw100;
      real x[NX], u[NU], qx[NX], qu[NU];
#pragma unroll
      for (int i = 0; i < NX; ++i) { x[i] = F(F_XQ + i, kp); qx[i] = (i < NXP) ? F(F_Q + (i < NXP ? i : 0), kp) : 0.0; }
#pragma unroll
      for (int a = 0; a < NU; ++a) { u[a] = F(F_UQ + a, kp); qu[a] = (k == 0) ? rlin : 0.0; }
#pragma unroll 1
      for (int j = 0; j < NR; ++j) {
        const real rho = F(F_RHO + j, kp);
        if (rho > 0.0) {
          real lo, hi;
          row_bounds(kp, j, lo, hi);
          const real rlo = rho * lo, rhi = rho * hi;
          real sh = F(F_S + j, kp);
          real y = row_dual(sh, rlo, rhi, lam);
          if (UPDATE) {
            const real rfx = rho * row_value(kp, j, x);
            const real rv = sh - y;
            const real shn = PP.alpha * rfx + (1.0 - PP.alpha) * rv + y;
            const real yn = row_dual(shn, rlo, rhi, lam);
            if (CHECK) {
              const real rvn = shn - yn;
              const real irho = bmpc_div(1.0, rho);        // one reciprocal per row instead of four divisions
              res = fmax(res, fmax(fabs(rfx - rvn) * irho, fabs(rvn - rv) * iw100));
              gap_r = fmax(gap_r, fabs(rfx - rvn) * irho);
              stp_r = fmax(stp_r, fabs(rvn - rv) * irho);
              const int cj = (shn > rhi + lam) ? ROW_UP_LIN : (shn > rhi) ? ROW_UP_KINK : (shn >= rlo) ? ROW_INACTIVE
                             : (shn >= rlo - lam) ? ROW_LO_KINK : ROW_LO_LIN;
              ncode |= row_bits(cj, j);
              if (fabs(rfx - rvn) / rho > 5e-2 || fabs(rvn - rv) / (100.0 * w) > 5e-2) BMPC_TRACE("      [row] k %d j %d rho %.3e prim %.3e dual %.3e fx %.4f hi %.4f w %.3e\n", k, j, rho, fabs(rfx - rvn) / rho, fabs(rvn - rv) / (100.0 * w), rfx / rho, hi, w);
            }
            F(F_S + j, kp) = shn;
            sh = shn;
            y = yn;
          }
          add_row_grad(kp, j, -(sh - 2.0 * y), qx);
        }
      }
#pragma unroll
      for (int a = 0; a < NU; ++a) {
        const real rho = F(F_RHO + NR + a, kp);
        real sh = F(F_SU + a, kp);
        real rv = bmpc_clamp(sh, rho * PP.ulo[a], rho * PP.uhi[a]);
        if (UPDATE) {
          const real ru = rho * u[a];
          const real shn = PP.alpha * ru + (1.0 - PP.alpha) * rv + (sh - rv);
          const real rvn = bmpc_clamp(shn, rho * PP.ulo[a], rho * PP.uhi[a]);
          if (CHECK) {
            const real irho = bmpc_div(1.0, rho);
            res = fmax(res, fmax(fabs(ru - rvn) * irho, fabs(rvn - rv) * iw100));
            gap_u = fmax(gap_u, fabs(ru - rvn) * irho);
            stp_u = fmax(stp_u, fabs(rvn - rv) * irho);
            ncode |= in_bits((shn > rho * PP.uhi[a]) ? IN_AT_HI : (shn < rho * PP.ulo[a]) ? IN_AT_LO : IN_FREE, a);
            if (fabs(ru - rvn) / rho > 5e-2 || fabs(rvn - rv) / (100.0 * w) > 5e-2) BMPC_TRACE("      [in] k %d a %d rho %.3e prim %.3e dual %.3e u %.4f w %.3e\n", k, a, rho, fabs(ru - rvn) / rho, fabs(rvn - rv) / (100.0 * w), u[a], w);
          }
          F(F_SU + a, kp) = shn;
          sh = shn;
          rv = rvn;
        }
        qu[a] -= 2.0 * rv - sh;
      }
#pragma unroll
      for (int i = 0; i < NX; ++i) F(F_XQ + i, kp) = qx[i];
#pragma unroll
      for (int a = 0; a < NU; ++a) F(F_UQ + a, kp) = qu[a];
      if (UPDATE && CHECK) {
        // rows without a penalty (rho = 0) stay "ignored"
#pragma unroll
        for (int j = 0; j < NR; ++j)
          if (!(F(F_RHO + j, kp) > 0.0)) ncode |= row_bits(ROW_IGNORED, j);
        set_changes += (ncode != stp()[kp]);
        stp()[kp] = ncode;
      }
    }
    lanes_sync();
    { const auto prof_rv = res; prof_end(8, prof_t0); return prof_rv; }
  }

  BMPC_DN void admm_assemble() { admm_rows<false, false>(); }   // cold variant, two call sites: one out-of-line copy

  // rho cache: the curvature-matched rho changes slowly from one MPC step to the next, so warm solves reuse the
  // values of the previous step (refreshed every P.rho_refresh solves) and skip the free factorisation + covariance sweep.
  BMPC_DN void store_rho() {
    const long long prof_t0 = prof_begin(3);
    real* cache = PP.rho_cache + (size_t)prob * rho_stride(PP.totalu);
#pragma unroll 1
    BMPC_FOR_NODES(k) {
      int b, t;
      node_of(k, b, t);
      const int kp = kp_of(b, t);
#pragma unroll
      for (int j = 0; j < NR + NU; ++j) cache[(size_t)k * (NR + NU) + j] = F(F_RHO + j, kp);
    }
    lanes_sync();
    prof_end(3, prof_t0);
  }
  BMPC_DN void load_rho() {
    const long long prof_t0 = prof_begin(3);
    const real* cache = ep_rho();
#pragma unroll 1
    BMPC_FOR_NODES(k) {
      int b, t;
      node_of(k, b, t);
      const int kp = kp_of(b, t);
#pragma unroll
      for (int j = 0; j < NR; ++j) {
        const real r = cache[(size_t)k * (NR + NU) + j];
        F(F_RHO + j, kp) = r;
        F(F_S + j, kp) *= r;      // ADMM start written by node_setup, scaled as in choose_rho
      }
#pragma unroll
      for (int a = 0; a < NU; ++a) {
        const real r = cache[(size_t)k * (NR + NU) + NR + a];
        F(F_RHO + NR + a, kp) = r;
        F(F_SU + a, kp) *= r;
      }
    }
    lanes_sync();
    prof_end(3, prof_t0);
  }
  // shifted codes of the previous optimum -> a consistent starting guess for the polish (multipliers start at their
  // natural values: 0 on kinks)
  BMPC_DN void guess_from_codes() {
    const long long prof_t0 = prof_begin(6);
#pragma unroll 1
    BMPC_FOR_NODES(k) {
      int b, t;
      node_of(k, b, t);
      const int kp = kp_of(b, t);
      const code_t code = stp()[kp];
      code_t ncode = 0;
#pragma unroll 1
      for (int j = 0; j < NR; ++j) {
        int cj = row_of(code, j);
        if (!(F(F_RHO + j, kp) > 0.0)) cj = ROW_IGNORED;
        else if (cj == ROW_IGNORED) cj = ROW_INACTIVE;
        ncode |= row_bits(cj, j);
        F(F_Y + j, kp) = 0.0;
      }
#pragma unroll 1
      for (int a = 0; a < NU; ++a) {
        ncode |= in_bits(in_of(code, a), a);
        F(F_Y + NR + a, kp) = 0.0;
      }
      stp()[kp] = ncode;
    }
    lanes_sync();
    prof_end(6, prof_t0);
  }
  BMPC_DN void store_codes() {
    const long long prof_t0 = prof_begin(9);
    code_t* codes = PP.code_cache + (size_t)prob * code_stride(PP.totalu);
#pragma unroll 1
    BMPC_FOR_NODES(k) {
      int b, t;
      node_of(k, b, t);
      codes[k] = stp()[kp_of(b, t)];
    }
    lanes_sync();
    prof_end(9, prof_t0);
  }

  // Residual balancing (Boyd et al. 2011, 3.4.1) on top of the curvature-matched rho: with rho in matched units the
  // primal gap |f'x - v| and the step |v+ - v| are commensurable; when one dominates, the whole group (state rows /
  // inputs) is rescaled.  The ADMM state (v, y) is kept: sh' = s (sh - y) + y.  The caller refactorises.
  BMPC_DN bool rebalance_rho() {
    const real gr = lanes_max(gap_r), sr = lanes_max(stp_r), gu = lanes_max(gap_u), su = lanes_max(stp_u);
    real fr = sqrt(gr / fmax(sr, 1e-30)), fu = sqrt(gu / fmax(su, 1e-30));
    fr = (gr > 1e-9 || sr > 1e-9) ? bmpc_clamp(fr, 0.2, 5.0) : 1.0;
    fu = (gu > 1e-9 || su > 1e-9) ? bmpc_clamp(fu, 0.2, 5.0) : 1.0;
    if (fr > 0.5 && fr < 2.0) fr = 1.0;
    if (fu > 0.5 && fu < 2.0) fu = 1.0;
    BMPC_TRACE("    rebalance: rows gap %.2e step %.2e -> x%.2f   inputs gap %.2e step %.2e -> x%.2f\n", gr, sr, fr, gu, su, fu);
    if (fr == 1.0 && fu == 1.0) return false;
#pragma unroll 1
    BMPC_FOR_NODES(k) {
      int b, t;
      node_of(k, b, t);
      const int kp = kp_of(b, t);
      const real lam = PP.lam_lin * Wbp()[b];
#pragma unroll
      for (int j = 0; j < NR; ++j) {
        const real rho = F(F_RHO + j, kp);
        if (rho > 0.0 && fr != 1.0) {
          real lo, hi;
          row_bounds(kp, j, lo, hi);
          const real sh = F(F_S + j, kp);
          const real y = row_dual(sh, rho * lo, rho * hi, lam);
          F(F_S + j, kp) = fr * (sh - y) + y;
          F(F_RHO + j, kp) = fr * rho;
        }
      }
      if (fu != 1.0) {
#pragma unroll
        for (int a = 0; a < NU; ++a) {
          const real rho = F(F_RHO + NR + a, kp);
          const real sh = F(F_SU + a, kp);
          const real rv = bmpc_clamp(sh, rho * PP.ulo[a], rho * PP.uhi[a]);
          F(F_SU + a, kp) = fu * rv + (sh - rv);
          F(F_RHO + NR + a, kp) = fu * rho;
        }
      }
    }
    lanes_sync();
    return true;
  }

  // ========================================================================================
  // Active-set polish (primal-dual active set on the exact-penalty QP; equalities of a guessed set are
  // imposed by a stiff penalty + augmented-Lagrangian refinement so that the same tree Riccati solves them)
  // ========================================================================================
  BMPC_DN void polish_guess() {
    const long long prof_t0 = prof_begin(6);
#pragma unroll 1
    BMPC_FOR_NODES(k) {
      int b, t;
      node_of(k, b, t);
      const int kp = kp_of(b, t);
      const real lam = PP.lam_lin * Wbp()[b];
      code_t code = 0;
#pragma unroll 1
      for (int j = 0; j < NR; ++j) {
        const real rho = F(F_RHO + j, kp);
        int cj = ROW_IGNORED;
        real y = 0.0;
        if (rho > 0.0) {
          real lo, hi;
          row_bounds(kp, j, lo, hi);
          const real sh = F(F_S + j, kp);
          const real rlo = rho * lo, rhi = rho * hi;
          if (sh > rhi + lam) cj = ROW_UP_LIN;
          else if (sh > rhi) { cj = ROW_UP_KINK; y = sh - rhi; }
          else if (sh >= rlo) cj = ROW_INACTIVE;
          else if (sh >= rlo - lam) { cj = ROW_LO_KINK; y = sh - rlo; }
          else cj = ROW_LO_LIN;
        }
        code |= row_bits(cj, j);
        F(F_Y + j, kp) = y;
      }
#pragma unroll 1
      for (int a = 0; a < NU; ++a) {
        const real rho = F(F_RHO + NR + a, kp);
        const real sh = F(F_SU + a, kp);
        int ca = IN_FREE;
        real y = 0.0;
        if (sh > rho * PP.uhi[a]) { ca = IN_AT_HI; y = sh - rho * PP.uhi[a]; }
        else if (sh < rho * PP.ulo[a]) { ca = IN_AT_LO; y = sh - rho * PP.ulo[a]; }
        code |= in_bits(ca, a);
        F(F_Y + NR + a, kp) = y;
      }
      stp()[kp] = code;
    }
    lanes_sync();
    prof_end(6, prof_t0);
  }

  BMPC_DN void polish_assemble() {
    const long long prof_t0 = prof_begin(6);
#pragma unroll 1
    BMPC_FOR_NODES(k) {
      int b, t;
      node_of(k, b, t);
      const int kp = kp_of(b, t);
      const real w = Wbp()[b];
      const real lam = PP.lam_lin * w;
      const code_t code = stp()[kp];
      real qx[NX], qu[NU];
#pragma unroll
      for (int i = 0; i < NX; ++i) qx[i] = (i < NXP) ? F(F_Q + (i < NXP ? i : 0), kp) : 0.0;
#pragma unroll 1
      for (int j = 0; j < NR; ++j) {
        const int cj = row_of(code, j);
        real lo, hi;
        row_bounds(kp, j, lo, hi);
        real g = 0.0;
        if (cj == ROW_UP_LIN) g = lam;
        else if (cj == ROW_LO_LIN) g = -lam;
        else if (cj == ROW_UP_KINK) g = F(F_Y + j, kp) - big_row(kp, j, w) * hi;
        else if (cj == ROW_LO_KINK) g = F(F_Y + j, kp) - big_row(kp, j, w) * lo;
        if (g != 0.0) add_row_grad(kp, j, g, qx);
      }
      // u = u_pinned + delta: the stage terms that couple delta (and the previous input) with the pinned part
      real up[NU];
#pragma unroll
      for (int a = 0; a < NU; ++a) { up[a] = pinned_value(code, a); qu[a] = (k == 0) ? rlin : 0.0; }
      const real rate = (k == 0) ? 0.0 : w;
      const real sig = (b >= PP.off[PP.NB] && t == PP.N - 1) ? 0.0 : 1.0;
#pragma unroll
      for (int a = 0; a < NU; ++a) {
#pragma unroll
        for (int b2 = 0; b2 < NU; ++b2) {
          real r2 = w * (PP.R[a * NU + b2] + PP.R[b2 * NU + a]);
          if (RATE && k == 0) r2 += 2.0 * PP.dR[a > b2 ? a : b2];
          qu[a] += r2 * up[b2];
        }
        if (RATE) {
          qu[a] += 2.0 * rate * sig * PP.dR[a] * up[a];
          qx[NXP + (RATE ? a : 0)] += -2.0 * rate * PP.dR[a] * up[a];
        }
      }
#pragma unroll
      for (int i = 0; i < NX; ++i) F(F_XQ + i, kp) = qx[i];
#pragma unroll
      for (int a = 0; a < NU; ++a) F(F_UQ + a, kp) = qu[a];
    }
    lanes_sync();
    prof_end(6, prof_t0);
  }

  // multiplier (augmented-Lagrangian) update on the guessed-active rows; returns this lane's max residual
  BMPC_DN real polish_multipliers() {
    const long long prof_t0 = prof_begin(6);
    real res = 0.0;
#pragma unroll 1
    BMPC_FOR_NODES(k) {
      int b, t;
      node_of(k, b, t);
      const int kp = kp_of(b, t);
      const real w = Wbp()[b];
      const code_t code = stp()[kp];
      real x[NX];
#pragma unroll
      for (int i = 0; i < NX; ++i) x[i] = F(F_XQ + i, kp);
#pragma unroll 1
      for (int j = 0; j < NR; ++j) {
        const int cj = row_of(code, j);
        if (cj == ROW_UP_KINK || cj == ROW_LO_KINK) {
          real lo, hi;
          row_bounds(kp, j, lo, hi);
          const real r = row_value(kp, j, x) - (cj == ROW_UP_KINK ? hi : lo);
          if (fabs(r) > 1e-6) BMPC_TRACE("      k %d row %d code %d r %.2e y %.3e lam %.3e big %.2e f=(%.3f,%.3f)\n", k, j, cj, r, F(F_Y + j, kp), PP.lam_lin * w, big_row(kp, j, w), F(F_FC, kp), F(F_FC + 1, kp));
          const real yn = F(F_Y + j, kp) + big_row(kp, j, w) * r;
          F(F_Y + j, kp) = yn;
          res = fmax(res, fabs(r));
        }
      }
    }
    lanes_sync();
    { const auto prof_rv = res; prof_end(6, prof_t0); return prof_rv; }
  }

  // pinned inputs take their bound value after the backward sweep (their gain rows and feed-forward are zero)
  BMPC_DN void polish_inject() {
    const long long prof_t0 = prof_begin(6);
#pragma unroll 1
    BMPC_FOR_NODES(k) {
      int b, t;
      node_of(k, b, t);
      const int kp = kp_of(b, t);
      const code_t code = stp()[kp];
#pragma unroll
      for (int a = 0; a < NU; ++a)
        if (pinned(code, a)) F(F_UQ + a, kp) = pinned_value(code, a);
    }
    lanes_sync();
    prof_end(6, prof_t0);
  }

  // Adjoint (costate) sweep of the equality-constrained solution in XQ/UQ: lam_k = dstage/dxi + A~' lam_{k+1}.  The
  // multiplier of a pinned input is minus the Lagrangian gradient with respect to it; for the free inputs that
  // gradient is zero (stationarity), whose largest violation is returned as a certificate.
  // Only lam itself is a recursion: the stage gradients (state cost, soft-row terms, input cost) are computed for all
  // nodes at once by adjoint_stage (node-parallel) and parked in the gain fields F_K, which are dead between the last
  // KKT solve of a polish pass and the next factorisation; the sweep then adds A~' lam and B~' lam per node.
  static constexpr int F_AX = F_K;        // stage gradient with respect to the (augmented) state
  static constexpr int F_AU = F_K + NX;   // stage gradient with respect to the input, without B~' lam
  static_assert(NX + NU <= NU * NX, "adjoint scratch does not fit the gain fields");
  BMPC_DN void adjoint_stage() {
#pragma unroll 1
    BMPC_FOR_NODES(k) {
      int b, t;
      node_of(k, b, t);
      const int kp = kp_of(b, t);
      const real w = Wbp()[b];
      real x[NX], u[NU], st_x[NX];
#pragma unroll
      for (int i = 0; i < NX; ++i) x[i] = F(F_XQ + i, kp);
#pragma unroll
      for (int a = 0; a < NU; ++a) u[a] = F(F_UQ + a, kp);
      const code_t code = stp()[kp];
      const real rate = (k == 0) ? 0.0 : w;
      const real sig = (b >= PP.off[PP.NB] && t == PP.N - 1) ? 0.0 : 1.0;
#pragma unroll
      for (int a = 0; a < NU; ++a) {
        real g = (k == 0) ? rlin : 0.0;
#pragma unroll
        for (int b2 = 0; b2 < NU; ++b2) {
          real r2 = w * (PP.R[a * NU + b2] + PP.R[b2 * NU + a]);
          if (RATE && k == 0) r2 += 2.0 * PP.dR[a > b2 ? a : b2];
          g += r2 * u[b2];
        }
        if (RATE) g += 2.0 * rate * PP.dR[a] * (sig * u[a] - x[NXP + (RATE ? a : 0)]);
        F(F_AU + a, kp) = g;
      }
      const real qs = w * (1.0 + PP.dq_scale);
      const real lamw = PP.lam_lin * w;
#pragma unroll
      for (int i = 0; i < NX; ++i) {
        real v = 0.0;
        if (i < NXP) {
          v = F(F_Q + (i < NXP ? i : 0), kp);
#pragma unroll
          for (int j = 0; j < NXP; ++j) v += qs * (QH()[(i < NXP ? i : 0) * NXP + j] + QH()[j * NXP + (i < NXP ? i : 0)]) * x[j];
        } else if (RATE) {
          v = 2.0 * rate * PP.dR[i - NXP] * (x[i] - u[i - NXP]);
        }
        st_x[i] = v;
      }
#pragma unroll 1
      for (int j = 0; j < NR; ++j) {
        const int cj = row_of(code, j);
        real g = 0.0;
        if (cj == ROW_UP_LIN) g = lamw;
        else if (cj == ROW_LO_LIN) g = -lamw;
        else if (cj == ROW_UP_KINK || cj == ROW_LO_KINK) g = F(F_Y + j, kp);
        if (g != 0.0) add_row_grad(kp, j, g, st_x);
      }
#pragma unroll
      for (int i = 0; i < NX; ++i) F(F_AX + i, kp) = st_x[i];
    }
    lanes_sync();
  }
  BMPC_D real adjoint_step(int kp, real* lam) {
    real lin[M::NLIN], gu[NU];
#pragma unroll
    for (int i = 0; i < M::NLIN; ++i) lin[i] = F(F_LIN + i, kp);
    const code_t code = stp()[kp];
    M::mulBT(PP, lin, lam, gu);
    real viol = 0.0;
#pragma unroll
    for (int a = 0; a < NU; ++a) {
      const real g = gu[a] + F(F_AU + a, kp);
      if (pinned(code, a)) F(F_Y + NR + a, kp) = -g;
      else viol = fmax(viol, fabs(g));
    }
    real al[NX];
    M::mulAT(PP, lin, lam, al);
#pragma unroll
    for (int i = 0; i < NX; ++i) lam[i] = F(F_AX + i, kp) + al[i];
    return viol;
  }

  BMPC_DN real polish_adjoint() {
    const long long prof_t0 = prof_begin(7);
    adjoint_stage();
    real viol = 0.0;
    if (team_leader()) {
#pragma unroll 1
    for (int d = PP.NB; d >= 0; --d) {
      const int nt = (d == 0) ? 1 : PP.N;
#pragma unroll 1
      for (int b = PP.off[d] + BMPC_LANE_ID; b < PP.off[d + 1]; b += BMPC_BLANES) {
        const real w = Wbp()[b];
        real lam[NX];
        if (d == PP.NB) {
          // terminal costate: 2 w Qf x_T (BranchMPCProx: - 2 w Qf' xRef as well; robustMPC: none, its terminal cost sits on
          // the chain's last stage); x_T was left in EXX by forward()
          const real* xT = EXXp() + NX * b;
          const real* xref = PP.xref + (size_t)prob * NXP;
#pragma unroll
          for (int i = 0; i < NX; ++i) {
            real v = 0.0;
            if (i < NXP && !bmpc_is_chain(PP.ctrl) && PP.ctrl != BMPC_CTRL_CVAR) {
#pragma unroll
              for (int j = 0; j < NXP; ++j) {
                v += w * (PP.Qf[(i < NXP ? i : 0) * NXP + j] + PP.Qf[j * NXP + (i < NXP ? i : 0)]) * xT[j];
                if (PP.ctrl == BMPC_CTRL_PROX) v -= 2.0 * w * PP.Qf[j * NXP + (i < NXP ? i : 0)] * xref[j];
              }
            }
            lam[i] = v;
          }
        } else {
          const int fc = bmpc_first_child(PP, b, d);
#pragma unroll
          for (int i = 0; i < NX; ++i) lam[i] = EXp()[NS * fc + i];
#pragma unroll 1
          for (int c = 1; c < PP.m; ++c)
#pragma unroll
            for (int i = 0; i < NX; ++i) lam[i] += EXp()[NS * (fc + c) + i];
        }
        const int kp0 = kp_of(b, 0);
#pragma unroll 1
        for (int t = nt - 1; t >= 0; --t) viol = fmax(viol, adjoint_step(kp0 + t, lam));
#pragma unroll
        for (int i = 0; i < NX; ++i) EXp()[NS * b + i] = lam[i];
      }
      bsync();
    }
    }
    team_sync();
    { const auto prof_rv = viol; prof_end(7, prof_t0); return prof_rv; }
  }

  // Primal-dual active-set update from the last equality-constrained solve.  Every candidate change carries a score in
  // multiplier units (primal violations are scaled by the row's curvature-matched stiffness), and only candidates
  // with score >= thresh are applied; score_max returns the largest score seen (for the one-change-at-a-time mode).
  // restricted=true (stalled refinement): only kink rows whose multiplier is out of range by more than the last
  // multiplier step are revised; everything else is left for a settled solve.
  BMPC_DN int polish_update_sets(bool restricted, real thresh, bool apply, real& score_max) {
    const long long prof_t0 = prof_begin(6);
    int changes = 0;
    real smax = 0.0;
    const real tol = 1e-7;
#pragma unroll 1
    BMPC_FOR_NODES(k) {
      int b, t;
      node_of(k, b, t);
      const int kp = kp_of(b, t);
      const real w = Wbp()[b];
      const real lam = PP.lam_lin * w;
      const real ytol = 1e-9 * lam;
      const code_t code = stp()[kp];
      code_t ncode = 0;
      real x[NX];
#pragma unroll
      for (int i = 0; i < NX; ++i) x[i] = F(F_XQ + i, kp);
#pragma unroll 1
      for (int j = 0; j < NR; ++j) {
        const int cj = row_of(code, j);
        int nj = cj;
        real ynew = 0.0, score = 0.0;
        if (cj != ROW_IGNORED) {
          real lo, hi;
          row_bounds(kp, j, lo, hi);
          const real fx = row_value(kp, j, x);
          const real y = F(F_Y + j, kp);
          const real stiff = F(F_RHO + j, kp) * PP.inv_theta;
          if (restricted) {
            if (cj == ROW_UP_KINK || cj == ROW_LO_KINK) {
              const real m2 = 2.0 * big_row(kp, j, w) * fabs(fx - (cj == ROW_UP_KINK ? hi : lo));
              const real ys = (cj == ROW_UP_KINK) ? y : -y;     // in [0, lam] when consistent
              if (ys + m2 < -ytol) { nj = ROW_INACTIVE; score = -ys; }
              else if (ys - m2 > lam + ytol) { nj = (cj == ROW_UP_KINK) ? ROW_UP_LIN : ROW_LO_LIN; score = ys - lam; }
            }
          } else if (cj == ROW_INACTIVE) {
            if (fx > hi + tol) { nj = ROW_UP_KINK; score = stiff * (fx - hi); }
            else if (fx < lo - tol) { nj = ROW_LO_KINK; score = stiff * (lo - fx); }
          } else if (cj == ROW_UP_KINK) {
            if (y < -ytol) { nj = ROW_INACTIVE; score = -y; }
            else if (y > lam + ytol) { nj = ROW_UP_LIN; score = y - lam; }
          } else if (cj == ROW_LO_KINK) {
            if (y > ytol) { nj = ROW_INACTIVE; score = y; }
            else if (y < -lam - ytol) { nj = ROW_LO_LIN; score = -y - lam; }
          } else if (cj == ROW_UP_LIN) {
            if (fx < hi - tol) { nj = ROW_UP_KINK; ynew = lam; score = stiff * (hi - fx); }
          } else if (cj == ROW_LO_LIN) {
            if (fx > lo + tol) { nj = ROW_LO_KINK; ynew = -lam; score = stiff * (fx - lo); }
          }
        }
        if (nj != cj) {
          smax = fmax(smax, score);
          if (apply && score >= thresh) {
            BMPC_TRACE("        k %d row %d: %d -> %d score %.3e\n", k, j, cj, nj, score);
            F(F_Y + j, kp) = ynew;
            ++changes;
          } else {
            nj = cj;
          }
        }
        ncode |= row_bits(nj, j);
      }
#pragma unroll 1
      for (int a = 0; a < NU; ++a) {
        const int ca = in_of(code, a);
        int na = ca;
        real score = 0.0;
        const real u = F(F_UQ + a, kp);
        const real y = F(F_Y + NR + a, kp);
        const real utol = 1e-9 * w;
        const real stiff = F(F_RHO + NR + a, kp) * PP.inv_theta_u;
        if (restricted) {
        } else if (ca == IN_FREE) {
          if (u > PP.uhi[a] + tol) { na = IN_AT_HI; score = stiff * (u - PP.uhi[a]); }
          else if (u < PP.ulo[a] - tol) { na = IN_AT_LO; score = stiff * (PP.ulo[a] - u); }
        } else if (ca == IN_AT_HI) {
          if (y < -utol) { na = IN_FREE; score = -y; }
        } else {
          if (y > utol) { na = IN_FREE; score = y; }
        }
        if (na != ca) {
          smax = fmax(smax, score);
          if (apply && score >= thresh) {
            BMPC_TRACE("        k %d input %d: %d -> %d score %.3e\n", k, a, ca, na, score);
            F(F_Y + NR + a, kp) = 0.0;
            ++changes;
          } else {
            na = ca;
          }
        }
        ncode |= in_bits(na, a);
      }
      if (apply) stp()[kp] = ncode;
    }
    lanes_sync();
    score_max = smax;
    { const auto prof_rv = changes; prof_end(6, prof_t0); return prof_rv; }
  }

  // returns true when the guessed active set was verified (then XQ/UQ hold the optimal x,u).
  // Per pass: factorise with stiff penalties on the guessed-active rows, refine the multipliers a few
  // augmented-Lagrangian steps at a time and let the primal-dual rules revise the sets; a set that stays
  // unchanged while the equality residual falls below 1e-7 is the verified optimum.  Conflicting guesses
  // (e.g. the collision rows of sibling branches, which see the same position one step after the branching
  // point) show up as multipliers running past their bounds and are revised without waiting for convergence.
  BMPC_DN bool polish(int& nfact, bool allow_careful, bool from_admm_state) {
    if (from_admm_state) polish_guess();
    // between the cutting-plane iterations of BranchMPC_CVaR the branch weights jump from one vertex of the multiplier set to
    // another: the verified active set of the last inner problem is several changes away, still far cheaper than ADMM
    const int base_passes = from_admm_state ? PP.polish_passes : PP.warm_passes * (PP.ctrl == BMPC_CTRL_CVAR ? 3 : 1);
    int prev_changes = 1 << 30;
    bool careful = false;
    const int max_passes = base_passes + PP.polish_careful;
    for (int pass = 0; pass < max_passes; ++pass) {
      if (!careful && pass >= base_passes) return false;
      factorize(FACT_POLISH);
      ++nfact;
      real res = 1.0, prev = 1e300;
      bool stalled = false;
      int al = 0;
      for (; al < PP.polish_al_iters; ++al) {
        polish_assemble();
        backward();
        polish_inject();
        forward();
        res = lanes_max(polish_multipliers());
        if (res < 1e-9) break;
        if (al >= 3 && res > 0.5 * prev) { stalled = true; break; }   // not contracting: conflicting or very weak set
        prev = res;
      }
      const bool settled = res < 1e-7;
      if (!(res < 1e30)) return false;              // non-finite data
      const real stat = lanes_max(polish_adjoint());   // multipliers of the pinned inputs + stationarity certificate
      BMPC_TRACE("    stationarity of the free inputs: %.2e\n", stat);
      (void)stat;
      // Set decisions need multipliers that are accurate (error ~ penalty x residual); a stalled refinement only
      // revises kink rows whose multipliers have plainly left their range.
      real smax = 0.0;
      int changes;
      if (careful && settled) {
        // one change at a time (the classical active-set rule): the most violated condition only
        polish_update_sets(false, 0.0, false, smax);
        smax = lanes_max(smax);
        changes = lanes_sum_int(polish_update_sets(false, smax * (1.0 - 1e-12), true, smax));
      } else {
        changes = lanes_sum_int(polish_update_sets(!settled, -1.0, true, smax));
      }
      BMPC_TRACE("    polish pass %d: al %d res %.2e stalled %d careful %d changes %d\n", pass, al, res, (int)stalled,
                 (int)careful, changes);
      if (changes == 0) return settled;
      if (settled && !careful) {
        // the all-at-once primal-dual iteration is not contracting: fall back to one change per pass
        if (changes > prev_changes) {
          if (!allow_careful || PP.polish_careful <= 0) return false;
          careful = true;
        }
        prev_changes = changes;
      }
    }
    return false;
  }

  // ========================================================================================
  // Interior-point fallback (primal-dual, Mehrotra predictor-corrector) for problems whose active set the ADMM + polish
  // combination does not settle (degenerate or conflicting sets: a fraction of a percent of the highway batches, a few
  // percent of the quadruped ones).  Same exact-penalty QP, same tree Riccati: every soft row side gets a slack
  //   upper:  a = hi + s - f'x > 0,  s >= 0,  multipliers y of the row and v = lam - y of s >= 0,  a y = s v = mu
  //   lower:  a = f'x - lo + s > 0,  ...                                  inputs:  a = uhi - u | u - ulo,  a y = mu
  // and the Newton step eliminates (a, s, y) per row: the row enters the stage Hessian with curvature
  // kappa = y / (a + y s / v) (the penalties of FACT_IPM) and the linear term with sg (y + c) - kappa f'x, so that one
  // factorisation + one backward/forward sweep return the full Newton target z+ (XQ/UQ); iterates move by
  // z <- z + alpha (z+ - z), which keeps the dynamics residual shrinking by (1 - alpha) per iteration.
  // The iteration state lives in this warp's global scratch (IPF): it is a rare path and must not cost shared memory.
  // ========================================================================================
  // One complementarity side.  soft: the side has a slack s (soft rows) or not (hard input bounds).
  // c0/c1 carry, per phase: PRED_STEP -> (dy, ds) of the predictor; CORR_ASM -> (rc1, rc2) of the corrector;
  // CORR_STEP -> (dy, ds) of the corrector, which the next PRED_ASM applies with step alpha before it assembles.
  BMPC_DN void ipm_side(int phase, bool soft, real sg, real gap, real dt, real lam, real& s_io, real& y_io, real& c0,
                        real& c1, real sigmu, real alpha, real& gsum, real& ksum) {
    real ss = soft ? s_io : 0.0;
    real y = y_io;
    if (phase == IPM_PRED_ASM && alpha > 0.0) {
      y += alpha * c0;
      y_io = y;
      if (soft) {
        ss += alpha * c1;
        s_io = ss;
      }
    }
    const real a = gap + ss;
    const real v = soft ? lam - y : 1.0;
    const real rv = 1.0 / v;
    const real rD = 1.0 / (a + y * ss * rv);
    const real kap = y * rD;
    real rc1, rc2;
    if (phase == IPM_PRED_ASM || phase == IPM_PRED_STEP) {
      rc1 = -a * y;
      rc2 = -ss * v;
    } else if (phase == IPM_CORR_ASM) {
      const real dy = c0, ds = c1, da = -sg * dt + ds;
      // second-order (Mehrotra) term weighted by `alpha` = omega in [0,1]: a short affine step means the affine direction
      // is a poor predictor and its products would throw the iterate off centre (weighted correctors, Colombo & Gondzio)
      rc1 = sigmu - a * y - alpha * da * dy;
      rc2 = soft ? sigmu - ss * v + alpha * ds * dy : 0.0;   // dv = -dy
      c0 = rc1;
      c1 = rc2;
    } else {
      rc1 = c0;
      rc2 = c1;
    }
    const real c = (rc1 - y * rc2 * rv) * rD;
    if (phase == IPM_PRED_ASM || phase == IPM_CORR_ASM) {
      gsum += sg * (y + c);
      ksum += kap;
      if (phase == IPM_PRED_ASM) ipm_acc += a * y + ss * v;
      return;
    }
    const real dy = sg * kap * dt + c;
    const real ds = soft ? (rc2 + ss * dy) * rv : 0.0;
    const real da = -sg * dt + ds;
    // largest step keeping a, y (and s, v) positive: a ratio test, single precision is plenty (the step is damped by 0.995)
    ipm_ratio = fmax(ipm_ratio, fmax(bmpc_ratio(-da, a), bmpc_ratio(-dy, y)));
    if (soft) ipm_ratio = fmax(ipm_ratio, fmax(bmpc_ratio(-ds, ss), bmpc_ratio(dy, v)));
    if (phase == IPM_PRED_STEP) ipm_acc += da * dy - (soft ? ds * dy : 0.0);   // second-order term of the predicted gap
    c0 = dy;
    c1 = ds;
  }

  // node-parallel pass of the interior-point iteration.  Leaves in ipm_acc the lane's complementarity sum (PRED_ASM) or
  // the sum of the predictor's second-order products (PRED_STEP), in ipm_ratio the lane's largest step ratio
  // max(-dq/q) over the positive quantities q (PRED_STEP, CORR_STEP).
  // PRED_ASM with alpha > 0 first moves the iterate by the corrector step of the previous iteration.
  BMPC_DN void ipm_pass(int phase, real sigmu, real alpha) {
    ipm_acc = 0.0;
    ipm_ratio = 0.0;
    const bool assemble = (phase == IPM_PRED_ASM || phase == IPM_CORR_ASM);
    const bool first = (phase == IPM_PRED_ASM && !(alpha > 0.0));
#pragma unroll 1
    BMPC_FOR_NODES(k) {
      int b, t;
      node_of(k, b, t);
      const int kp = kp_of(b, t);
      const real lam = PP.lam_lin * Wbp()[b];
      real x[NX], u[NU], xp[NX], up[NU], qx[NX], qu[NU];
#pragma unroll
      for (int i = 0; i < NX; ++i) {
        x[i] = IPF(IP_X + i, kp);
        xp[i] = first ? x[i] : F(F_XQ + i, kp);
        if (phase == IPM_PRED_ASM && !first) {
          x[i] += alpha * (xp[i] - x[i]);
          IPF(IP_X + i, kp) = x[i];
        }
        qx[i] = (i < NXP) ? F(F_Q + (i < NXP ? i : 0), kp) : 0.0;
      }
#pragma unroll
      for (int a = 0; a < NU; ++a) {
        u[a] = IPF(IP_U + a, kp);
        up[a] = first ? u[a] : F(F_UQ + a, kp);
        if (phase == IPM_PRED_ASM && !first) {
          u[a] += alpha * (up[a] - u[a]);
          IPF(IP_U + a, kp) = u[a];
        }
        qu[a] = (k == 0) ? rlin : 0.0;
      }
#pragma unroll 1
      for (int j = 0; j < NR; ++j) {
        real gs = 0.0, ks = 0.0;
        if (F(F_RHO + j, kp) > 0.0) {
          real lo, hi;
          row_bounds(kp, j, lo, hi);
          const real tv = row_value(kp, j, x);
          const real dt = row_value(kp, j, xp) - tv;
          ipm_side(phase, true, 1.0, hi - tv, dt, lam, IPF(IP_RS + 4 * j, kp), IPF(IP_RS + 4 * j + 1, kp),
                   IPF(IP_COR + 4 * j, kp), IPF(IP_COR + 4 * j + 1, kp), sigmu, alpha, gs, ks);
          if (lo > 0.5 * BMPC_NOLO)
            ipm_side(phase, true, -1.0, tv - lo, dt, lam, IPF(IP_RS + 4 * j + 2, kp), IPF(IP_RS + 4 * j + 3, kp),
                     IPF(IP_COR + 4 * j + 2, kp), IPF(IP_COR + 4 * j + 3, kp), sigmu, alpha, gs, ks);
          if (assemble) add_row_grad(kp, j, gs - ks * tv, qx);
        }
        if (assemble) IPF(IP_KAP + j, kp) = ks;
      }
#pragma unroll 1
      for (int a = 0; a < NU; ++a) {
        real gs = 0.0, ks = 0.0, dummy = 0.0;
        const real dt = (phase == IPM_PRED_ASM) ? 0.0 : up[a] - u[a];
        ipm_side(phase, false, 1.0, PP.uhi[a] - u[a], dt, 0.0, dummy, IPF(IP_IN + 2 * a, kp), IPF(IP_COR + 4 * NR + 4 * a, kp),
                 IPF(IP_COR + 4 * NR + 4 * a + 1, kp), sigmu, alpha, gs, ks);
        ipm_side(phase, false, -1.0, u[a] - PP.ulo[a], dt, 0.0, dummy, IPF(IP_IN + 2 * a + 1, kp),
                 IPF(IP_COR + 4 * NR + 4 * a + 2, kp), IPF(IP_COR + 4 * NR + 4 * a + 3, kp), sigmu, alpha, gs, ks);
        if (assemble) {
          IPF(IP_KAP + NR + a, kp) = ks;
          qu[a] += gs - ks * u[a];
        }
      }
      if (assemble) {
#pragma unroll
        for (int i = 0; i < NX; ++i) F(F_XQ + i, kp) = qx[i];
#pragma unroll
        for (int a = 0; a < NU; ++a) F(F_UQ + a, kp) = qu[a];
      }
    }
    lanes_sync();
  }

  // start: (x, u) of the last KKT solve with the inputs pulled strictly inside their box, slacks that make every row
  // side feasible with room, multipliers at the centre of their range; returns this lane's number of complementarity pairs
  BMPC_DN int ipm_init() {
    int pairs = 0;
#pragma unroll 1
    BMPC_FOR_NODES(k) {
      int b, t;
      node_of(k, b, t);
      const int kp = kp_of(b, t);
      const real w = Wbp()[b];
      const real lam = PP.lam_lin * w;
      real x[NX];
#pragma unroll
      for (int i = 0; i < NX; ++i) {
        x[i] = F(F_XQ + i, kp);
        IPF(IP_X + i, kp) = x[i];
      }
#pragma unroll
      for (int j = 0; j < NR; ++j) {
        real lo = 0.0, hi = 0.0, tv = 0.0;
        const bool on = F(F_RHO + j, kp) > 0.0;
        if (on) {
          row_bounds(kp, j, lo, hi);
          tv = row_value(kp, j, x);
          pairs += 2;
          if (lo > 0.5 * BMPC_NOLO) pairs += 2;
        }
        IPF(IP_RS + 4 * j, kp) = fmax(tv - hi, 0.0) + PP.ipm_s0;
        IPF(IP_RS + 4 * j + 1, kp) = 0.5 * lam;
        IPF(IP_RS + 4 * j + 2, kp) = fmax(lo - tv, 0.0) + PP.ipm_s0;
        IPF(IP_RS + 4 * j + 3, kp) = 0.5 * lam;
      }
#pragma unroll
      for (int a = 0; a < NU; ++a) {
        const real range = PP.uhi[a] - PP.ulo[a];
        const real uc = bmpc_clamp(F(F_UQ + a, kp), PP.ulo[a] + 0.1 * range, PP.uhi[a] - 0.1 * range);
        IPF(IP_U + a, kp) = uc;
        IPF(IP_IN + 2 * a, kp) = PP.ipm_y0 * lam * PP.ipm_s0 / (PP.uhi[a] - uc);
        IPF(IP_IN + 2 * a + 1, kp) = PP.ipm_y0 * lam * PP.ipm_s0 / (uc - PP.ulo[a]);
        pairs += 2;
      }
    }
    lanes_sync();
    return pairs;
  }

  // active set + multipliers implied by the converged interior-point state -> starting guess of the polish
  BMPC_DN void ipm_guess() {
#pragma unroll 1
    BMPC_FOR_NODES(k) {
      int b, t;
      node_of(k, b, t);
      const int kp = kp_of(b, t);
      const real lam = PP.lam_lin * Wbp()[b];
      real x[NX];
#pragma unroll
      for (int i = 0; i < NX; ++i) x[i] = IPF(IP_X + i, kp);
      code_t code = 0;
#pragma unroll
      for (int j = 0; j < NR; ++j) {
        int cj = ROW_IGNORED;
        real ym = 0.0;
        if (F(F_RHO + j, kp) > 0.0) {
          real lo, hi;
          row_bounds(kp, j, lo, hi);
          const real tv = row_value(kp, j, x);
          cj = ROW_INACTIVE;
          {
            const real s = IPF(IP_RS + 4 * j, kp), y = IPF(IP_RS + 4 * j + 1, kp);
            const real a = hi - tv + s;
            if (s * lam > (lam - y)) cj = ROW_UP_LIN;
            else if (a * lam < y) { cj = ROW_UP_KINK; ym = y; }
          }
          if (lo > 0.5 * BMPC_NOLO) {
            const real s = IPF(IP_RS + 4 * j + 2, kp), y = IPF(IP_RS + 4 * j + 3, kp);
            const real a = tv - lo + s;
            if (s * lam > (lam - y)) cj = ROW_LO_LIN;
            else if (a * lam < y) { cj = ROW_LO_KINK; ym = -y; }
          }
        }
        code |= row_bits(cj, j);
        F(F_Y + j, kp) = ym;
      }
#pragma unroll
      for (int a = 0; a < NU; ++a) {
        const real u = IPF(IP_U + a, kp);
        const real range = PP.uhi[a] - PP.ulo[a];
        const real yh = IPF(IP_IN + 2 * a, kp), yl = IPF(IP_IN + 2 * a + 1, kp);
        int ca = IN_FREE;
        real ym = 0.0;
        if ((PP.uhi[a] - u) * lam < yh * range) { ca = IN_AT_HI; ym = yh; }
        else if ((u - PP.ulo[a]) * lam < yl * range) { ca = IN_AT_LO; ym = -yl; }
        code |= in_bits(ca, a);
        F(F_Y + NR + a, kp) = ym;
      }
      stp()[kp] = code;
    }
    lanes_sync();
  }

  BMPC_DN void ipm_export() {   // interior-point iterate -> XQ/UQ (what finish() reads)
#pragma unroll 1
    BMPC_FOR_NODES(k) {
      int b, t;
      node_of(k, b, t);
      const int kp = kp_of(b, t);
#pragma unroll
      for (int i = 0; i < NX; ++i) F(F_XQ + i, kp) = IPF(IP_X + i, kp);
#pragma unroll
      for (int a = 0; a < NU; ++a) F(F_UQ + a, kp) = IPF(IP_U + a, kp);
    }
    lanes_sync();
  }

  // returns true when the complementarity gap and the carried residual factor reached their tolerances
  BMPC_DN bool ipm_solve(int& nfact, int& iters) {
    const int pairs = lanes_sum_int(ipm_init());
    const real mu_tol = PP.ipm_mu_tol * PP.lam_lin;
    real resfac = 1.0, alpha = 0.0;
    for (int it = 0; it <= PP.ipm_max_iter; ++it) {
      ipm_pass(bmpc_opaque(IPM_PRED_ASM), 0.0, alpha);   // applies the previous step, then assembles the predictor
      const real mu = lanes_sum(ipm_acc) / pairs;
      BMPC_TRACE("    ipm %d: mu %.3e resfac %.2e\n", it, mu, resfac);
      if (!(mu < 1e300)) return false;
      // the end game is superlinear, so the tight tolerance usually costs one more iteration; degenerate problems (pairs with
      // both factors vanishing) stall near 1e-9 lam, which is still 1e-7 from the optimum, and are accepted there
      if (mu < (it < 20 ? mu_tol : 10.0 * mu_tol) && resfac < 1e-7) return true;
      if (it == PP.ipm_max_iter) break;
      factorize(FACT_IPM);
      ++nfact;
      kkt_solve();
      ipm_pass(bmpc_opaque(IPM_PRED_STEP), 0.0, 0.0);
      const real a_aff = fmin(1.0, 1.0 / fmax(lanes_max(ipm_ratio), 1e-300));
      // gap after the affine step: every product moves to (1 - a) q1 q2 + a^2 dq1 dq2
      const real mu_aff = fmax((1.0 - a_aff) * mu + a_aff * a_aff * lanes_sum(ipm_acc) / pairs, 0.0);
      real sigma = mu_aff / mu;
      sigma = sigma * sigma * sigma;
      ipm_pass(bmpc_opaque(IPM_CORR_ASM), sigma * mu, a_aff);
      kkt_solve();
      ipm_pass(bmpc_opaque(IPM_CORR_STEP), 0.0, 0.0);
      alpha = fmin(1.0, 0.995 / fmax(lanes_max(ipm_ratio), 1e-300));
      BMPC_TRACE("      a_aff %.3f sigma %.2e alpha %.4f\n", a_aff, sigma, alpha);
      if (!(alpha > 0.0)) return false;
      resfac *= (1.0 - alpha);
      ++iters;
      ++ipm_iters;
    }
    return false;
  }

  // ========================================================================================
  // Final pass: clamp inputs, roll the linear dynamics out, write outputs and persistent state
  // ========================================================================================
  // The only recursion of the final pass is the rollout of the linear dynamics under the clamped inputs (finish_step,
  // lanes = branches); objective terms and every output are written by the node-parallel emit_nodes afterwards.
  BMPC_D void finish_step(int kp, real* x) {
    real lin[M::NLIN], cc[M::NCC], u[NU], xn[NX];
#pragma unroll
    for (int i = 0; i < M::NLIN; ++i) lin[i] = F(F_LIN + i, kp);
#pragma unroll
    for (int i = 0; i < M::NCC; ++i) cc[i] = F(F_CC + i, kp);
#pragma unroll
    for (int a = 0; a < NU; ++a) {
      u[a] = bmpc_clamp(F(F_UQ + a, kp), PP.ulo[a], PP.uhi[a]);
      F(F_UQ + a, kp) = u[a];
    }
#pragma unroll
    for (int i = 0; i < NX; ++i) F(F_XQ + i, kp) = x[i];
    M::mulA(PP, lin, x, xn);
    M::addBu(PP, lin, u, xn);
    M::addC(cc, xn);
#pragma unroll
    for (int i = 0; i < NX; ++i) x[i] = xn[i];
  }

  // objective of one node (slacks eliminated) + its rows of uPred / xPred / uLin / OldInput / xprev
  BMPC_DN real emit_nodes() {
    real J = 0.0;
    real* uLin = PP.uLin + (size_t)prob * (PP.totalu + 1) * NU;
    real* xP = PP.out.xPred ? PP.out.xPred + (size_t)prob * PP.pub_totalx * NXP : nullptr;
    real* xprev = PP.xprev ? PP.xprev + (size_t)prob * PP.pub_totalx * NXP : nullptr;   // robustMPC's LTV shift source
    const bool robust = bmpc_is_chain(PP.ctrl);
    const bool belief = PP.ctrl == BMPC_CTRL_BELIEF;
#pragma unroll 1
    BMPC_FOR_NODES(k) {
      int b, t;
      node_of(k, b, t);
      const int kp = kp_of(b, t);
      const real w = Wbp()[b];
      const bool leaf_last = (b >= PP.off[PP.NB] && t == PP.N - 1);
      const real rate = (k == 0) ? 0.0 : w;
      const real sig = leaf_last ? 0.0 : 1.0;
      const real* Qs = (robust && leaf_last) ? PP.Qf : QH();   // robustMPC: the dummy stage carries the terminal cost Qf
      real x[NX], u[NU];
#pragma unroll
      for (int i = 0; i < NX; ++i) x[i] = F(F_XQ + i, kp);
#pragma unroll
      for (int a = 0; a < NU; ++a) u[a] = F(F_UQ + a, kp);
      const real qs = w * (1.0 + PP.dq_scale);
#pragma unroll
      for (int i = 0; i < NXP; ++i) {
        real a = 0.0;
#pragma unroll
        for (int j = 0; j < NXP; ++j) a += Qs[i * NXP + j] * x[j];
        J += qs * x[i] * a + F(F_Q + i, kp) * x[i];
      }
      if (RATE) {
        // input-rate pair (previous input v = x[NXP..], this input u): rate [v'dR v - 2 v'dR u + sig u'dR u]
#pragma unroll
        for (int a = 0; a < NU; ++a) {
          const real v = x[NXP + (RATE ? a : 0)];
          J += rate * PP.dR[a] * (v * v - 2.0 * v * u[a] + sig * u[a] * u[a]);
        }
        if (k == 0) {
#pragma unroll
          for (int a = 0; a < NU; ++a)
#pragma unroll
            for (int b2 = 0; b2 < NU; ++b2) J += PP.dR[a > b2 ? a : b2] * u[a] * u[b2];   // root quirk (:312)
        }
      }
#pragma unroll
      for (int a = 0; a < NU; ++a) {
        real v = 0.0;
#pragma unroll
        for (int b2 = 0; b2 < NU; ++b2) v += PP.R[a * NU + b2] * u[b2];
        J += w * u[a] * v + ((k == 0) ? rlin * u[a] : 0.0);
      }
      const real lam = (belief && leaf_last) ? 0.0 : PP.lam_lin * w;   // the belief chain's last state carries no rows (:199)
#pragma unroll 1
      for (int j = 0; j < NR; ++j) {
        real lo, hi;
        row_bounds(kp, j, lo, hi);
        const real fx = row_value(kp, j, x);
        J += lam * (fmax(fx - hi, 0.0) + fmax(lo - fx, 0.0));
      }
      const int kx = bmpc_ndx(PP, b) + t;
      if (xP) {
#pragma unroll
        for (int i = 0; i < NXP; ++i) xP[(size_t)kx * NXP + i] = x[i];
      }
      if (xprev) {
#pragma unroll
        for (int i = 0; i < NXP; ++i) xprev[(size_t)kx * NXP + i] = x[i];
      }
      if (PP.out.uPred && k < PP.pub_totalu) {
        real* o = PP.out.uPred + ((size_t)prob * PP.pub_totalu + k) * NU;
#pragma unroll
        for (int a = 0; a < NU; ++a) o[a] = u[a];
      }
#pragma unroll
      for (int a = 0; a < NU; ++a) uLin[(size_t)k * NU + a] = u[a];
      if (k == PP.totalu - 1) {
#pragma unroll
        for (int a = 0; a < NU; ++a) uLin[(size_t)(k + 1) * NU + a] = u[a];   // uLin gets the last row twice (:1229)
      }
      if (k == 0) {
#pragma unroll
        for (int a = 0; a < NU; ++a) {
          PP.oldin[(size_t)prob * NU + a] = u[a];
          if (PP.out.u0) PP.out.u0[(size_t)prob * NU + a] = u[a];
        }
      }
    }
    return J;
  }

  BMPC_DN real finish() {
    const long long prof_t0 = prof_begin(9);
    real J = 0.0;
    real* xP = PP.out.xPred ? PP.out.xPred + (size_t)prob * PP.pub_totalx * NXP : nullptr;
    const bool robust = bmpc_is_chain(PP.ctrl);
    if (team_leader()) {
#pragma unroll 1
    for (int d = 0; d <= PP.NB; ++d) {
      const int nt = (d == 0) ? 1 : PP.N;
#pragma unroll 1
      for (int b = PP.off[d] + BMPC_LANE_ID; b < PP.off[d + 1]; b += BMPC_BLANES) {
        real x[NX];
        if (d == 0) {
#pragma unroll
          for (int i = 0; i < NX; ++i) x[i] = (i < NXP) ? PP.x0[(size_t)prob * NXP + (i < NXP ? i : 0)] : 0.0;
        } else {
          const int pa = bmpc_parent(PP, b, d);
#pragma unroll
          for (int i = 0; i < NX; ++i) x[i] = EXXp()[NX * pa + i];
        }
        const int kp0 = kp_of(b, 0);
#pragma unroll 1
        for (int t = 0; t < nt; ++t) finish_step(kp0 + t, x);
        if (d == PP.NB && !robust) {
          const real w = Wbp()[b];
          if (xP) {
            const int kx = bmpc_ndx(PP, b);
#pragma unroll
            for (int i = 0; i < NXP; ++i) xP[(size_t)(kx + nt) * NXP + i] = x[i];
          }
          const real* xref = PP.xref + (size_t)prob * NXP;
#pragma unroll
          for (int i = 0; i < NXP; ++i) {
            real a = 0.0, c = 0.0;
#pragma unroll
            for (int j = 0; j < NXP; ++j) {
              a += PP.Qf[i * NXP + j] * x[j];
              c += PP.Qf[j * NXP + i] * xref[j];
            }
            // terminal x' (w Qf) x; BranchMPC has no linear term there (:1094), BranchMPCProx has -2 w xRef' Qf (:308)
            J += w * x[i] * a - ((PP.ctrl == BMPC_CTRL_PROX) ? 2.0 * w * c * x[i] : 0.0);
          }
        }
#pragma unroll
        for (int i = 0; i < NX; ++i) EXXp()[NX * b + i] = x[i];
      }
      bsync();
    }
    }
    team_sync();
    J += emit_nodes();
    if (PP.ctrl == BMPC_CTRL_BELIEF) belief_outputs();
    { const auto prof_rv = lanes_sum(J); prof_end(9, prof_t0); return prof_rv; }
  }

  // A failed solve returns the plan the episode already had: previous inputs (uLin rows, OldInput; zeros before the first
  // success) and, where the library keeps them (robustMPC), the previous predicted states; other states read NaN.
  BMPC_DN void keep_plan() {
    const real* uLin = PP.uLin + (size_t)prob * (PP.totalu + 1) * NU;
    const real* xprev = PP.xprev ? PP.xprev + (size_t)prob * PP.pub_totalx * NXP : nullptr;
#pragma unroll 1
    BMPC_FOR_NODES(k) {
      if (PP.out.uPred && k < PP.pub_totalu) {
#pragma unroll
        for (int a = 0; a < NU; ++a) PP.out.uPred[((size_t)prob * PP.pub_totalu + k) * NU + a] = uLin[(size_t)k * NU + a];
      }
      if (k == 0 && PP.out.u0) {
#pragma unroll
        for (int a = 0; a < NU; ++a) PP.out.u0[(size_t)prob * NU + a] = PP.oldin[(size_t)prob * NU + a];
      }
    }
    if (PP.out.xPred) {
#pragma unroll 1
      for (int q = BMPC_LANE_ID; q < PP.pub_totalx * NXP; q += BMPC_LANES)
        PP.out.xPred[(size_t)prob * PP.pub_totalx * NXP + q] = xprev ? xprev[q] : bmpc_nan();
    }
    lanes_sync();
  }

  // any non-finite input left by the last forward sweep?
  BMPC_DN bool solution_is_finite() {
    int bad = 0;
#pragma unroll 1
    BMPC_FOR_NODES(k) {
      int b, t;
      node_of(k, b, t);
      const int kp = kp_of(b, t);
#pragma unroll
      for (int a = 0; a < NU; ++a) {
        const real v = F(F_UQ + a, kp);
        bad |= !(fabs(v) < 1e300);
      }
    }
    return lanes_or_int(bad) == 0;
  }

  // ========================================================================================
  // BranchMPC_CVaR (MPC_branch.py:1598-2152).  The reference minimises the epigraph variable J of a cone program:
  //   root cone (:1969-1984):            u0' R u0 + lam 1's_root + rho_0 <= J
  //   cone of child i of branch b (:1940-1967):   C_c(x,u,s) + rho_c [c not a leaf] + sigma_b + mu+_k - mu-_k <= 0,  k = idx_b + i
  //       C_c = sum over the child's N nodes of (x - xRef)'Q(x - xRef) + u'Ru + lam 1's
  //   dual-CVaR equality of branch b (:1790-1804):  rho_b + sigma_b = (1/alpha) sum_i p_bi mu-_(idx_b m + i),   rho, mu+- >= 0
  // (note the two index rules for mu: k = idx_b + i in the cones, idx_b m + i in the equalities - the reference's own
  // inconsistency, SURVEY 8a-Q7, reproduced).  With multipliers nu_c >= 0 on the cones the trajectory part is the tree QP of
  // this solver with branch weights nu_c, and eliminating (J, rho, sigma, mu) leaves the concave problem
  //   max over nu in E of  V(nu) = min over trajectories of  r0 + sum_c nu_c C_c
  //   E: sum_i nu_(b,i) <= nu_b (1 at the root);   sum over {(b,i): idx_b + i = k} nu_(b,i) <= (p_(b',i')/alpha) sum_j nu_(b',j),  (b',i') = divmod(k, m)
  // solved by Kelley's cutting planes: every inner solve at nu^k returns the branch costs C^k, i.e. the cut
  // V(nu) <= r0^k + nu . C^k; the next multipliers maximise the model over E (a small LP, dense simplex below); the gap
  // between the model's maximum and V at the current multipliers certifies the optimum.  A best-response fixed point (no
  // tie between branches) needs two inner solves; ties converge like a bisection.
  // ========================================================================================
  BMPC_D real* cvNu() { return PP.cv + (size_t)cv_team() * PP.cv_reals; }   // [nbranch] current multipliers (by branch id; [0] = 1)
  BMPC_D real* cvNuNew() { return cvNu() + PP.nbranch; }
  BMPC_D real* cvP() { return cvNuNew() + PP.nbranch; }                      // [bdim][m] child probabilities of the non-leaf branches
  BMPC_D real* cvCuts() { return cvP() + PP.off[PP.NB] * PP.m; }             // [max_cuts][nbranch]: [0] = r0, [c] = C_c
  BMPC_D real* cvTab() { return cvCuts() + (size_t)PP.cvar_max_cuts * PP.nbranch; }
  BMPC_D int* cvBasis() { return reinterpret_cast<int*>(cvTab() + (size_t)PP.cv_rows * PP.cv_cols); }
  BMPC_D int cv_team() const {
#if defined(__CUDA_ARCH__)
    return (int)blockIdx.x;
#else
    return 0;
#endif
  }
  real cvar_ub, cvar_val, cvar_scale;

  // first inner problem of a step: the multipliers of the previous step if the solver kept them, else the products of the
  // branch probabilities (any positive weights give a valid first cut)
  BMPC_DN void cvar_first_weights() {
    const real* cache = PP.nu_cache + (size_t)prob * PP.nbranch;
    const bool have = PP.started[prob] && PP.cache_state[(size_t)prob * 2 + 1] == 1 && cache[0] == 1.0;
    if (team_leader()) {
#pragma unroll 1
      for (int b = BMPC_LANE_ID; b < PP.nbranch; b += BMPC_BLANES) {
        const real nu = (b == 0) ? 1.0 : (have ? cache[b] : Wbp()[b]);
        cvNu()[b] = nu;
        Wbp()[b] = (b == 0) ? 1.0 : fmax(nu, PP.cvar_floor);
      }
    }
    team_sync();
  }

  // unweighted cost of every node at the inner solution (XQ/UQ) -> per-branch sums = cut number `cut`
  BMPC_DN void cvar_costs(int cut) {
    const real* xref = PP.xref + (size_t)prob * NXP;
#pragma unroll 1
    BMPC_FOR_NODES(k) {
      int b, t;
      node_of(k, b, t);
      const int kp = kp_of(b, t);
      real x[NX], u[NU];
#pragma unroll
      for (int i = 0; i < NX; ++i) x[i] = F(F_XQ + i, kp);
#pragma unroll
      for (int a = 0; a < NU; ++a) u[a] = F(F_UQ + a, kp);
      real c = 0.0;
      if (k > 0) {
#pragma unroll
        for (int i = 0; i < NXP; ++i) {
          real a = 0.0;
          if constexpr (XF) {
            // (S x)' Q (S x) - 2 xRef' Q x + xRef' Q xRef: the reference transforms the quadratic term only (:1938, :1962)
            real g = 0.0;
#pragma unroll
            for (int j = 0; j < NXP; ++j) {
              a += QH()[i * NXP + j] * x[j];
              g += PP.Q[i * NXP + j] * xref[j];
            }
            c += x[i] * a + g * (xref[i] - 2.0 * x[i]);
          } else {
#pragma unroll
            for (int j = 0; j < NXP; ++j) a += PP.Q[i * NXP + j] * (x[j] - xref[j]);
            c += (x[i] - xref[i]) * a;
          }
        }
      }
#pragma unroll
      for (int a = 0; a < NU; ++a) {
        real v = 0.0;
#pragma unroll
        for (int b2 = 0; b2 < NU; ++b2) v += PP.R[a * NU + b2] * u[b2];
        c += u[a] * v;
      }
#pragma unroll 1
      for (int j = 0; j < NR; ++j) {
        real lo, hi;
        row_bounds(kp, j, lo, hi);
        const real fx = row_value(kp, j, x);
        c += PP.lam_lin * (fmax(fx - hi, 0.0) + fmax(lo - fx, 0.0));
      }
      F(F_AX, kp) = c;
    }
    team_sync();
    if (team_leader()) {
      real* row = cvCuts() + (size_t)cut * PP.nbranch;
#pragma unroll 1
      for (int b = BMPC_LANE_ID; b < PP.nbranch; b += BMPC_BLANES) {
        const int nt = (b == 0) ? 1 : PP.N;
        const int kp0 = kp_of(b, 0);
        real acc = 0.0;
#pragma unroll 1
        for (int t = 0; t < nt; ++t) acc += F(F_AX, kp0 + t);
        row[b] = acc;
      }
    }
    team_sync();
  }

  // Master problem: maximise t over (nu in E, t <= r0^k + nu . C^k for every cut k) by the dense tableau simplex (all
  // right-hand sides are non-negative, so the slack basis is feasible; Dantzig's rule, Bland's after 60 pivots).  First warp
  // of the team; lanes = tableau columns.  Leaves the maximiser in cvNuNew and returns the model value.
  BMPC_DN real cvar_master(int ncuts) {
    const int nb = PP.nbranch, m = PP.m, bdim = PP.off[PP.NB];
    const int nvar = nb;                       // nu_1..nu_(nb-1) in columns 0..nb-2, t in column nb-1
    const int nE = bdim + (bdim + m - 1);
    const int nrow = nE + ncuts;               // constraint rows 1..nrow; row 0 = objective
    const int ncol = nvar + nrow + 1;          // structural, slack, right-hand side
    const int W = PP.cv_cols;
    real* T = cvTab();
    int* basis = cvBasis();
    const int lane = BMPC_LANE_ID;
    real ub = 0.0;
    if (team_leader()) {
#pragma unroll 1
      for (int q = lane; q < (nrow + 1) * W; q += BMPC_BLANES) T[q] = 0.0;
      bsync();
      if (lane == 0) {
        const real* P = cvP();
        T[nb - 1] = -1.0;                      // maximise t
#pragma unroll 1
        for (int b = 0; b < bdim; ++b) {
          real* r = T + (size_t)(1 + b) * W;
          const int fc = bmpc_first_child(PP, b, bmpc_depth(PP, b));
#pragma unroll 1
          for (int i = 0; i < m; ++i) r[fc + i - 1] = 1.0;
          if (b > 0) r[b - 1] = -1.0;
          r[ncol - 1] = (b == 0) ? 1.0 : 0.0;
        }
#pragma unroll 1
        for (int k = 0; k < bdim + m - 1; ++k) {
          real* r = T + (size_t)(1 + bdim + k) * W;
#pragma unroll 1
          for (int b = 0; b < bdim; ++b) {
            const int i = k - b;
            if (i >= 0 && i < m) r[bmpc_first_child(PP, b, bmpc_depth(PP, b)) + i - 1] += 1.0;
          }
          const int bp = k / m, ip = k - bp * m;
          const real cap = P[bp * m + ip] / PP.cvar_alpha;
          const int fcp = bmpc_first_child(PP, bp, bmpc_depth(PP, bp));
#pragma unroll 1
          for (int j = 0; j < m; ++j) r[fcp + j - 1] -= cap;
        }
#pragma unroll 1
        for (int k = 0; k < ncuts; ++k) {
          real* r = T + (size_t)(1 + nE + k) * W;
          const real* cut = cvCuts() + (size_t)k * nb;
#pragma unroll 1
          for (int c = 1; c < nb; ++c) r[c - 1] = -cut[c] / cvar_scale;
          r[nb - 1] = 1.0;
          r[ncol - 1] = cut[0] / cvar_scale;
        }
#pragma unroll 1
        for (int i = 1; i <= nrow; ++i) {
          T[(size_t)i * W + nvar + i - 1] = 1.0;
          basis[i] = nvar + i - 1;
        }
      }
      bsync();
#pragma unroll 1
      for (int pivots = 0; pivots < 2000; ++pivots) {
        // entering column: most negative reduced cost (smallest index once Bland's rule is on)
        const bool bland = pivots >= 60;
        real best = -1e-11;
        int jin = -1;
#pragma unroll 1
        for (int j = lane; j < ncol - 1; j += BMPC_BLANES) {
          const real rc = T[j];
          if (bland ? (rc < -1e-11 && jin < 0) : (rc < best)) { best = rc; jin = j; }
        }
#if defined(__CUDA_ARCH__)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const real ob = __shfl_xor_sync(BMPC_FULL_MASK, best, o);
          const int oj = __shfl_xor_sync(BMPC_FULL_MASK, jin, o);
          const bool take = bland ? (oj >= 0 && (jin < 0 || oj < jin)) : (oj >= 0 && (jin < 0 || ob < best || (ob == best && oj < jin)));
          if (take) { best = ob; jin = oj; }
        }
#endif
        if (jin < 0) break;
        // leaving row: minimum ratio (ties: smallest basis index)
        real rbest = 1e300;
        int iout = -1;
#pragma unroll 1
        for (int i = 1 + lane; i <= nrow; i += BMPC_BLANES) {
          const real a = T[(size_t)i * W + jin];
          if (a > 1e-9) {
            const real ratio = T[(size_t)i * W + ncol - 1] / a;
            if (ratio < rbest - 1e-12 || (ratio <= rbest + 1e-12 && (iout < 0 || basis[i] < basis[iout]))) { rbest = ratio; iout = i; }
          }
        }
#if defined(__CUDA_ARCH__)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const real orr = __shfl_xor_sync(BMPC_FULL_MASK, rbest, o);
          const int oi = __shfl_xor_sync(BMPC_FULL_MASK, iout, o);
          if (oi >= 0 && (iout < 0 || orr < rbest - 1e-12 || (orr <= rbest + 1e-12 && basis[oi] < basis[iout]))) { rbest = orr; iout = oi; }
        }
#endif
        if (iout < 0) break;   // unbounded: cannot happen (E is bounded and t is cut from above)
        const real piv = T[(size_t)iout * W + jin];
        const real ipiv = 1.0 / piv;
        bsync();
#pragma unroll 1
        for (int j = lane; j < ncol; j += BMPC_BLANES) T[(size_t)iout * W + j] *= ipiv;
        bsync();
#pragma unroll 1
        for (int i = 0; i <= nrow; ++i) {
          if (i == iout) continue;
          const real f = T[(size_t)i * W + jin];
          if (f != 0.0) {
#pragma unroll 1
            for (int j = lane; j < ncol; j += BMPC_BLANES) T[(size_t)i * W + j] -= f * T[(size_t)iout * W + j];
          }
          bsync();
        }
        if (lane == 0) basis[iout] = jin;
        bsync();
      }
#pragma unroll 1
      for (int c = lane; c < nb; c += BMPC_BLANES) cvNuNew()[c] = (c == 0) ? 1.0 : 0.0;
      bsync();
#pragma unroll 1
      for (int i = 1 + lane; i <= nrow; i += BMPC_BLANES)
        if (basis[i] < nb - 1) cvNuNew()[basis[i] + 1] = fmax(T[(size_t)i * W + ncol - 1], 0.0);
      bsync();
      ub = T[ncol - 1] * cvar_scale;
    }
    return lanes_max(team_leader() ? ub : -1e300);
  }

  // new multipliers -> branch weights of the next inner problem: the linear cost scales with the weight, the ADMM start is
  // rebuilt from the last solution (choose_rho scales it), the active-set codes of the last polish stay as the next guess
  BMPC_DN void cvar_reweight() {
#pragma unroll 1
    BMPC_FOR_NODES(k) {
      int b, t;
      node_of(k, b, t);
      const int kp = kp_of(b, t);
      const real wold = Wbp()[b];
      const real wnew = (b == 0) ? 1.0 : fmax(cvNuNew()[b], PP.cvar_floor);
      const real ratio = bmpc_div(wnew, wold);
#pragma unroll
      for (int i = 0; i < NXP; ++i) F(F_Q + i, kp) *= ratio;
      real x[NX];
#pragma unroll
      for (int i = 0; i < NX; ++i) x[i] = F(F_XQ + i, kp);
#pragma unroll 1
      for (int j = 0; j < NR; ++j) {
        real lo, hi;
        row_bounds(kp, j, lo, hi);
        F(F_S + j, kp) = bmpc_clamp(row_value(kp, j, x), lo, hi);
      }
#pragma unroll
      for (int a = 0; a < NU; ++a) F(F_SU + a, kp) = bmpc_clamp(F(F_UQ + a, kp), PP.ulo[a], PP.uhi[a]);
    }
    team_sync();
#pragma unroll 1
    for (int b = BMPC_LANE_ID; b < PP.nbranch; b += BMPC_LANES) {
      cvNu()[b] = cvNuNew()[b];
      Wbp()[b] = (b == 0) ? 1.0 : fmax(cvNuNew()[b], PP.cvar_floor);
    }
    team_sync();
  }

  // After inner solve number `outer`: add its cut, test the gap, otherwise move to the maximiser of the cutting-plane model.
  // Returns true when the multipliers are optimal (cvar_val = objective J of the cone program) or the cut budget is spent.
  BMPC_DN bool cvar_outer(int outer, int& status) {
    cvar_costs(outer);
    const real* cut = cvCuts() + (size_t)outer * PP.nbranch;
    real val = 0.0;
#pragma unroll 1
    for (int c = 0; c < PP.nbranch; ++c) val += ((c == 0) ? 1.0 : cvNu()[c]) * cut[c];   // same order in every lane: uniform
    cvar_val = val;
    if (outer == 0) cvar_scale = fmax(val, 1.0);
    BMPC_TRACE("  cvar outer %d: value %.9f model %.9f\n", outer, val, outer > 0 ? cvar_ub : 0.0);
    if (outer > 0 && cvar_ub - val <= PP.cvar_tol * fmax(1.0, fabs(cvar_ub))) return true;
    if (outer + 1 >= PP.cvar_max_cuts) {
      if (status == BMPC_STATUS_POLISHED) status = BMPC_STATUS_CONVERGED;   // best iterate, gap not closed to cvar_tol
      return true;
    }
    cvar_ub = cvar_master(outer + 1);
    real move = 0.0;
#pragma unroll 1
    for (int c = 1; c < PP.nbranch; ++c) move = fmax(move, fabs(cvNuNew()[c] - cvNu()[c]));
    if (outer > 0 && move <= 1e-13) return true;   // best-response fixed point
    cvar_reweight();
    return false;
  }

  // ========================================================================================
  // One tree QP with the current branch weights: rho (cached or curvature-matched afresh), warm polish from the active-set
  // codes in the slab when use_codes is set, else / then ADMM with polish attempts and the interior-point fallback.
  // Returns the status; for POLISHED / CONVERGED the solution is in XQ/UQ.
  BMPC_D int inner_solve(bool reuse_rho, bool keep_rho, int& nfact, int& iters) {
    int status = BMPC_STATUS_MAXITER;
    bool have_xu = false;
    if (reuse_rho) {
      load_rho();
    } else {
      factorize(FACT_FREE);
      choose_rho();
      if (keep_rho) store_rho();
      ++nfact;
    }
#if defined(__CUDA_ARCH__)
    if (stage_pending) stage_next();   // every read of the stage (expansion, load_rho) lies behind a team barrier by now
#endif
    if (use_codes) {
      // warm solve: the previous optimum's active set, shifted in time, is usually one or two changes away
      guess_from_codes();
      if (polish(nfact, false, false)) {
        status = BMPC_STATUS_POLISHED;
        have_xu = true;
      }
    }
    int next_polish = PP.polish_first, polish_gap = PP.polish_every, next_forced = PP.polish_force;
    int nfail = 0;
    bool ipm_tried = false;
    const int it0 = iters, iter_cap = iters + PP.max_iter;
    if (!have_xu) {
      factorize(FACT_ADMM);
      ++nfact;
      admm_assemble();
    }
    while (!have_xu && iters < iter_cap) {
      kkt_solve();
      ++iters;
      const int it_here = iters - it0;   // iterations of THIS inner problem (the interior point's count as well)
      const bool check = (it_here % PP.check_every == 0);
      real res = 1e300;
      int moved = 1 << 20;
      if (check) {
        res = lanes_max(admm_rows<true, true>());
        moved = lanes_sum_int(set_changes);
      } else {
        admm_rows<true, false>();
      }
      const bool conv = res < PP.eps_abs;
      if (check) BMPC_TRACE("  it %d res %.3e set changes %d\n", iters, res, moved);
      // polish when the active set implied by the ADMM state has stopped moving (or, at the latest, every force_every)
      const bool due = check && it_here >= next_polish && (moved <= PP.polish_stable || it_here >= next_forced);
      if (due || conv) {
        next_forced = it_here + PP.polish_force;
        next_polish = it_here + polish_gap;
        polish_gap *= 2;   // back off: a problem whose active set is slow to settle should not pay for many attempts
        // ipm_after == 100 (tests): skip the polish attempt, the interior point runs at the first opportunity
        if (PP.ipm_after != 100 && polish(nfact, it_here >= 4 * PP.polish_first, true)) {
          status = BMPC_STATUS_POLISHED;
          have_xu = true;
          break;
        }
        if (conv) status = BMPC_STATUS_CONVERGED;
        ++nfail;
        if (!conv && PP.ipm_after > 0 && (nfail >= PP.ipm_after || PP.ipm_after == 100) && !ipm_tried) {
          // the active set does not settle: interior point on the same Riccati, then a polish from its (clean) active set
          ipm_tried = true;
          const long long t_ipm0 = prof_begin(1);
          const bool ipm_ok = ipm_solve(nfact, iters);
          prof_end(1, t_ipm0);
          if (ipm_ok) {
            ipm_guess();
            if (polish(nfact, false, false)) {
              status = BMPC_STATUS_POLISHED;
            } else {
              ipm_export();
              status = BMPC_STATUS_CONVERGED;
            }
            have_xu = true;
            break;
          }
        }
        if (!conv && PP.rebalance) rebalance_rho();
        factorize(FACT_ADMM);
        ++nfact;
        admm_assemble();
        if (conv) break;
      }
    }
    if (!have_xu) {
      // XQ/UQ hold q~: one more KKT solve gives the (x,u) of the final ADMM state
      kkt_solve();
    }
    return status;
  }

  BMPC_D void solve(int prob_) {
#if defined(__CUDA_ARCH__)
    const long long t_start = clock64();
#endif
    prob = prob_;
    polpar = PP.polpar ? PP.polpar + (size_t)prob * PP.zm * 4 : nullptr;
    const int warm = PP.started[prob];
    const bool cvar = PP.ctrl == BMPC_CTRL_CVAR;
    int* cstate = PP.cache_state + (size_t)prob * 2;   // [0] age of the cached rho (-1: none), [1] cached codes valid
    // the risk multipliers move the branch weights, and with them the curvature rho is matched to: no rho cache for CVaR
    const bool reuse_rho = !cvar && warm && PP.rho_refresh > 0 && cstate[0] >= 0 && cstate[0] < PP.rho_refresh;
    // cstate[1]: 0 = no valid codes; 1 = valid, start with the warm polish; k > 1 = valid, but the episode's last warm attempt
    // ended on the ADMM path - and then the next one does too, 19 times out of 20 (measured: its active set is in flux for
    // many steps in a row) - so the attempt is skipped for k - 1 more solves
    // (low byte: that countdown; above it: how many attempts in a row have failed - the skip doubles with each, PP.warm_backoff)
    const int code_state = cstate[1];
    const int cs_count = code_state > 0 ? (code_state & 0xff) : 0, cs_level = code_state > 0 ? (code_state >> 8) : 0;
    const bool skip_warm = !cvar && cs_count > 1;
    use_codes = warm && PP.warm_polish && cs_count >= 1 && !skip_warm && (reuse_rho || cvar || PP.warm_on_refresh);
    const bool tried_warm = use_codes;
    t_phase = 0;
#if defined(__CUDA_ARCH__)
    stage_acquire();
#endif
    if (PP.ctrl == BMPC_CTRL_ROBUST) expand_chain();
    else if (PP.ctrl == BMPC_CTRL_BELIEF) expand_belief();
    else expand_tree();
    nsolve = 0;
    ipm_iters = 0;
    int nfact = 0, iters = 0, status = BMPC_STATUS_MAXITER;
    bool finite = true;
    // one pass for the QP controllers; BranchMPC_CVaR repeats the inner solve with the multipliers of its master problem
#pragma unroll 1
    for (int outer = 0;; ++outer) {
      status = inner_solve(reuse_rho && outer == 0, !cvar, nfact, iters);
      finite = solution_is_finite();
      if (!cvar || !finite || status > BMPC_STATUS_CONVERGED) break;
      if (cvar_outer(outer, status)) break;
      use_codes = (status == BMPC_STATUS_POLISHED);   // the polish left its verified active set in the slab
    }
    // Only a solved problem is adopted (the reference sets feasible = 1 for OSQP's 'solved' alone and otherwise keeps its
    // previous plan and linearisation inputs, MPC_branch.py:1224, :1269-1272): an iterate that ended on the iteration caps
    // or on non-finite data is neither returned nor used as the next warm start.
    if (finite && status <= BMPC_STATUS_CONVERGED) {
      real J = finish();
      if (cvar) J = cvar_val;   // the epigraph variable of the cone program: r0 + sum_c nu_c C_c at the optimal multipliers
      if (status == BMPC_STATUS_POLISHED) store_codes();
      if (cvar) {
        real* cache = PP.nu_cache + (size_t)prob * PP.nbranch;
        for (int b = BMPC_LANE_ID; b < PP.nbranch; b += BMPC_LANES) cache[b] = (b == 0) ? 1.0 : cvNu()[b];
      }
      if (BMPC_LANE_ID == 0) {
        if (PP.out.status) PP.out.status[prob] = status;
        if (PP.out.objective) PP.out.objective[prob] = J;
        PP.started[prob] = 1;
        cstate[0] = reuse_rho ? cstate[0] + 1 : 0;
        if (status != BMPC_STATUS_POLISHED) cstate[1] = 0;
        else if (skip_warm) cstate[1] = code_state - 1;
        else if (cvar || !tried_warm) cstate[1] = (cs_level << 8) | 1;
        else if (iters == 0) cstate[1] = 1;
        else {
          const int lvl = !PP.warm_backoff ? 1 : (cs_level < 3 ? cs_level + 1 : 4);
          const int skip = PP.warm_skip << (lvl - 1);
          cstate[1] = (lvl << 8) | (1 + (skip < 24 ? skip : 24));
        }
      }
    } else {
      keep_plan();
      if (BMPC_LANE_ID == 0) {
        if (PP.out.status) PP.out.status[prob] = finite ? status : BMPC_STATUS_NUMERIC;
        if (PP.out.objective) PP.out.objective[prob] = bmpc_nan();
        cstate[1] = 0;
      }
    }
    if (BMPC_LANE_ID == 0) {
      if (PP.out.iters) PP.out.iters[prob] = iters;
      if (PP.out.nfact) PP.out.nfact[prob] = nfact;
      if (PP.out.nsolve) PP.out.nsolve[prob] = nsolve;
#if defined(__CUDA_ARCH__)
      if (PP.out.cycles) PP.out.cycles[prob] = PP.cycles_mode > 0 ? (int64_t)t_phase : (int64_t)(clock64() - t_start);
      if (PP.cost) PP.cost[prob] = (int)min((long long)0x7fffffff, (clock64() - t_start) >> 10);
#else
      if (PP.out.cycles) PP.out.cycles[prob] = ipm_iters;   // host build: no clock; reports the interior-point iterations instead
#endif
    }
    lanes_sync();
  }
};
