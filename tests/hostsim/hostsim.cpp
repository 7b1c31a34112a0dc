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
  const bool prox = cfg->controller == BMPC_CTRL_PROX;
  if (cfg->controller == BMPC_CTRL_ROBUST) {
    if (g_split) run_layout<HighwayModel, 11, BMPC_SLAB_SPLIT, 9>(P); else run_layout<HighwayModel, 11, BMPC_SLAB_SHARED, 9>(P);
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
}
