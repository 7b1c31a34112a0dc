// TEST INFRASTRUCTURE (never shipped, never loaded by the product package): the solver text of
// belief-planning_b200/csrc/bmpc_solver.h compiled by g++ as a single-lane host program, so that the CPU test
// suite can check the device ALGORITHM (tree expansion, Riccati/ADMM, polish, outputs, warm-start state) against
// the oracle without a GPU.  The CUDA library builds the same headers with 32 lanes per problem.
#include <stdlib.h>

#include <string>
#include <vector>

#include "bmpc_host.h"
#include "bmpc_solver.h"

namespace {
bool g_split = false;
template <class M, int NR, int MODE, int NC = 1>
void run_layout(KParams& P) {
  using S = Solver<M, NR, MODE, NC>;
  P.slab_reals = S::slab_reals(P.nup, P.nbx);
  std::vector<real> slab(P.slab_reals, 0.0), factor(S::factor_reals(P.nup) + 2, 0.0), ipm(S::ipm_reals(P.nup), 0.0);
  S solver(P, slab.data(), factor.data(), ipm.data(), 0);
  for (int i = 0; i < P.count; ++i) solver.solve(i);
}
template <class M, int NR>
void run(KParams& P) {
  if (g_split) run_layout<M, NR, BMPC_SLAB_SPLIT>(P); else run_layout<M, NR, BMPC_SLAB_SHARED>(P);
}
std::string g_err;
// belief-state MPC inputs of the next hostsim_solve call (hostsim_set_belief)
const double* g_b0 = nullptr;
const double* g_xbackup = nullptr;
int g_xb_cols = 0;
// merge scenario inputs of the next hostsim_solve call (hostsim_set_transform)
const double* g_xform = nullptr;
const double* g_xbounds = nullptr;
// lookup table of the *_REF policies (hostsim_set_lookup): stays until replaced
std::vector<double> g_lut_x, g_lut_y;
void apply_lookup(KParams& P) {
  P.lut_x = g_lut_x.empty() ? nullptr : g_lut_x.data();
  P.lut_y = g_lut_y.empty() ? nullptr : g_lut_y.data();
  P.lut_n = (int)g_lut_x.size();
}
}  // namespace

extern "C" {

const char* hostsim_last_error(void) { return g_err.c_str(); }

int hostsim_sizes(const bmpc_config* cfg, int32_t* nbranch, int32_t* totalx, int32_t* totalu) {
  KParams P;
  const int rc = bmpc::make_params(*cfg, &P, &g_err);
  if (rc != BMPC_OK) return rc;
  *nbranch = P.nbranch;
  *totalx = P.pub_totalx;
  *totalu = P.pub_totalu;
  return BMPC_OK;
}

// All pointers are HOST pointers; uLin/pbest/oldin/started are the persistent state (caller-owned here).
int hostsim_solve(const bmpc_config* cfg, const double* x0, const double* z0, const double* xref,
                  const double* policy_params, int64_t count, double* uLin, int32_t* pbest, double* oldin,
                  int32_t* started, double* rho_cache, int64_t* code_cache, int32_t* cache_state,
                  double* xprev, const bmpc_outputs* out) {
  KParams P;
  const int rc = bmpc::make_params(*cfg, &P, &g_err);
  if (rc != BMPC_OK) return rc;
  if (!bmpc::supported_instance(cfg->model, cfg->n_rows, cfg->controller, P.zpw[P.zNB])) {
    g_err = "unsupported (model, n_rows, controller)";
    return BMPC_E_UNSUPPORTED;
  }
  g_split = cfg->slab_mode == BMPC_SLAB_SPLIT;
  P.count = (int)count;
  P.x0 = x0;
  P.z0 = z0;
  P.xref = xref;
  P.polpar = policy_params;
  P.xform = g_xform;
  P.xbounds = g_xbounds;
  apply_lookup(P);
  P.uLin = uLin;
  P.pbest = pbest;
  P.oldin = oldin;
  P.started = started;
  P.rho_cache = rho_cache;
  P.code_cache = reinterpret_cast<long long*>(code_cache);
  P.cache_state = cache_state;
  P.xprev = (cfg->controller == BMPC_CTRL_ROBUST) ? xprev : nullptr;
  P.out = *out;
  // BranchMPC_CVaR: master-problem scratch and (per call, not persistent here) the multiplier cache
  std::vector<real> cv(P.cv_reals + 4, 0.0), nu((size_t)count * P.nbranch + 1, 0.0);
  P.cv = cv.data();
  P.nu_cache = nu.data();
  std::vector<real> bel(P.bel_reals + 4, 0.0);
  P.bel = bel.data();
  P.b0 = g_b0;
  P.xbackup = g_xbackup;
  P.xb_cols = g_xb_cols;
  const bool prox = cfg->controller == BMPC_CTRL_PROX;
  if (bmpc_is_chain(cfg->controller)) {
    if (g_split) run_layout<HighwayModel, 11, BMPC_SLAB_SPLIT, 9>(P); else run_layout<HighwayModel, 11, BMPC_SLAB_SHARED, 9>(P);
    return BMPC_OK;
  }
  if (cfg->model == BMPC_MODEL_MERGE) {
    run<MergeModel, 3>(P);
    return BMPC_OK;
  }
  if (cfg->model == BMPC_MODEL_HIGHWAY) {
    switch (cfg->n_rows) {
      case 0: prox ? run<RateAug<HighwayModel>, 1>(P) : run<HighwayModel, 1>(P); break;
      case 1: prox ? run<RateAug<HighwayModel>, 2>(P) : run<HighwayModel, 2>(P); break;
      default: prox ? run<RateAug<HighwayModel>, 3>(P) : run<HighwayModel, 3>(P); break;
    }
  } else {
    prox ? run<RateAug<QuadrupedModel>, 1>(P) : run<QuadrupedModel, 1>(P);
  }
  return BMPC_OK;
}

int hostsim_set_belief(const double* b0, const double* xbackup, int32_t cols) {
  g_b0 = b0;
  g_xbackup = xbackup;
  g_xb_cols = cols;
  return BMPC_OK;
}

int hostsim_set_lookup(const double* xs, const double* ys, int32_t n) {
  g_lut_x.assign(xs, xs + (n > 0 ? n : 0));
  g_lut_y.assign(ys, ys + (n > 0 ? n : 0));
  return BMPC_OK;
}

int hostsim_set_transform(const double* S, const double* state_bounds) {
  g_xform = S;
  g_xbounds = state_bounds;
  return BMPC_OK;
}

int hostsim_topology(const bmpc_config* cfg, int32_t* ndx, int32_t* ndu, int32_t* depth, int32_t* parent) {
  KParams P;
  const int rc = bmpc::make_params(*cfg, &P, &g_err);
  if (rc != BMPC_OK) return rc;
  for (int b = 0; b < P.nbranch; ++b) {
    const int d = bmpc_depth(P, b);
    ndx[b] = bmpc_ndx(P, b);
    ndu[b] = bmpc_ndu(P, b);
    depth[b] = d;
    parent[b] = d == 0 ? -1 : bmpc_parent(P, b, d);
  }
  return BMPC_OK;
}
}  // extern "C"

namespace {
// point-wise model functions with the same model text the kernels use (what bmpc_eval_model does on the device)
template <class M>
void eval_points(const KParams& P, const double* x, const double* z, const double* u, const double* polpar, int64_t count,
                 double* A, double* B, double* C, double* xp, double* zpred, double* p, double* hlin, double* dh) {
  constexpr int NX = M::NX, NU = M::NU;
  for (int64_t i = 0; i < count; ++i) {
    const real* xi = x + i * NX;
    {
      real lin[M::NLIN], cc[M::NCC], xn[NX];
      M::linearize(P, xi, u + i * NU, lin, cc, xn);
      M::denseA(P, lin, A + i * NX * NX);
      M::denseB(P, lin, B + i * NX * NU);
      M::expandC(cc, C + i * NX);
      for (int q = 0; q < NX; ++q) xp[i * NX + q] = xn[q];
    }
    const real* zi = z + i * NX;
    real h, dhx, dhy;
    M::collision(P, xi, zi, h, dhx, dhy);
    hlin[i] = h - (dhx * xi[0] + dhy * xi[1]);
    for (int q = 0; q < NX; ++q) dh[i * NX + q] = 0.0;
    dh[i * NX] = dhx;
    dh[i * NX + 1] = dhy;
    real hi[BMPC_MAX_POLICIES], himax = -1e300;
    for (int k = 0; k < P.zm; ++k) {
      const real* par = polpar ? polpar + (i * P.zm + k) * 4 : P.pol_par[k];
      const real* par0 = polpar ? polpar + (i * P.zm) * 4 : P.pol_par[0];
      real zl[NX];
      real* zo = zpred + i * P.zN * P.zm * NX;
      hi[k] = M::policy_safety(P, P.pol_kind[k], par, P.pol_kind[0], par0, xi, zi, zl, P.zN, [&](int t, const real* zz) {
        for (int q = 0; q < NX; ++q) zo[((size_t)t * P.zm + k) * NX + q] = zz[q];
      });
      himax = fmax(himax, hi[k]);
    }
    real sum = 0.0;
    for (int k = 0; k < P.zm; ++k) sum += M::branch_weight(P, hi[k], himax);
    for (int k = 0; k < P.zm; ++k) p[i * P.zm + k] = M::branch_weight(P, hi[k], himax) / sum;
  }
}
}  // namespace

extern "C" {
int hostsim_eval_model(const bmpc_config* cfg, const double* x, const double* z, const double* u, const double* polpar,
                       int64_t count, double* A, double* B, double* C, double* xp, double* zpred, double* p, double* hlin,
                       double* dh) {
  KParams P;
  const int rc = bmpc::make_params(*cfg, &P, &g_err);
  if (rc != BMPC_OK) return rc;
  apply_lookup(P);
  if (cfg->model == BMPC_MODEL_HIGHWAY) eval_points<HighwayModel>(P, x, z, u, polpar, count, A, B, C, xp, zpred, p, hlin, dh);
  else if (cfg->model == BMPC_MODEL_MERGE) eval_points<MergeModel>(P, x, z, u, polpar, count, A, B, C, xp, zpred, p, hlin, dh);
  else eval_points<QuadrupedModel>(P, x, z, u, polpar, count, A, B, C, xp, zpred, p, hlin, dh);
  return BMPC_OK;
}
}
