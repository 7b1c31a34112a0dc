// libbranchmpc.so: C ABI (include/branchmpc.h) over the sm_100a kernels.  Nothing here throws or aborts across the
// ABI; every failure is an integer code plus a message in bmpc_last_error().
#include <cuda_runtime.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "bmpc_env.cuh"
#include "bmpc_hmm.cuh"
#include "bmpc_host.h"
#include "bmpc_solver.h"

// ------------------------------------------------------------------------------------------------------------
// kernels
// ------------------------------------------------------------------------------------------------------------

// Persistent warps: every warp pulls problem indices from a global counter and solves them one at a time in its
// own slab (shared memory when it fits, global/L2 otherwise).
// MODE = BMPC_SLAB_SHARED (whole slab in shared memory), BMPC_SLAB_SPLIT (iterate fields in shared memory, factor fields
// in this warp's global region, which stays L2-resident), BMPC_SLAB_GLOBAL (everything in the global region).
template <class M, int NR, int MODE, int NC = 1>
// Launch bounds 384 x 1 = at most 168 registers per thread: four teams of 96 lanes, two of 192 or one of 256 fit an SM; the
// block size is the team size the host picked for the tree (configure_instance).
__global__ void __launch_bounds__(384, 1) bmpc_solve_kernel() {
  const KParams& P = bmpc_cP;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x;
  constexpr bool SPLIT = (MODE == BMPC_SLAB_SPLIT);
  (void)smem_raw;
  real* slab;
  real* factor = nullptr;
  if (MODE == BMPC_SLAB_GLOBAL) {
    slab = P.gws + (size_t)blockIdx.x * P.slab_reals;
  } else {
    slab = reinterpret_cast<real*>(smem_raw);
    if (SPLIT) factor = P.gws + (size_t)blockIdx.x * P.factor_reals;
  }
  Solver<M, NR, MODE, NC> S(P, slab, factor, P.ipm + (size_t)blockIdx.x * P.ipm_reals, lane);
  __shared__ int next_problem[4];                       // [0] next work index, [1] its state is being staged, [2] wait verdict
  __shared__ __align__(8) unsigned long long stage_bar; // completion barrier of the staging copies
  const bool staging = (MODE == BMPC_SLAB_SHARED) && P.stage_on;
  if (staging) {
    if (lane == 0) bmpc_mbar_init(&stage_bar, 1);
    S.stage_bar = &stage_bar;
    S.stage_slot = next_problem;
  }
  if (lane == 0) next_problem[0] = atomicAdd(P.counter, 1);
  __syncthreads();
  int idx = next_problem[0];
  while (idx < P.count) {
    S.solve(P.order ? P.order[idx] : idx);   // ends with a team barrier
    if (staging && !S.stage_broken && next_problem[0] >= 0) {
      // the solve claimed the next work item itself and started the copies of its episode's state (Solver::stage_next)
      idx = next_problem[0];
      S.stage_have = next_problem[1] != 0;
    } else {
      __syncthreads();
      if (lane == 0) next_problem[0] = atomicAdd(P.counter, 1);
      __syncthreads();
      idx = next_problem[0];
    }
  }
}

// Longest-processing-time-first work order: episodes sorted by the cycles their previous solve took (descending,
// 64 logarithmic buckets), so that the few expensive problems start first and overlap with the bulk instead of forming
// the tail of the launch.  Single block; counting sort in shared memory.
__global__ void __launch_bounds__(1024) bmpc_order_kernel(const int* __restrict__ cost, int count, int* __restrict__ order) {
  __shared__ int hist[64];
  __shared__ int start[64];
  if (threadIdx.x < 64) hist[threadIdx.x] = 0;
  __syncthreads();
  for (int i = threadIdx.x; i < count; i += blockDim.x) {
    const int c = cost[i];
    const int bkt = c <= 0 ? 0 : min(63, 2 * (31 - __clz(c)) + ((c >> max(0, 30 - __clz(c))) & 1));
    atomicAdd(&hist[bkt], 1);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int acc = 0;
    for (int b = 63; b >= 0; --b) { start[b] = acc; acc += hist[b]; }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < count; i += blockDim.x) {
    const int c = cost[i];
    const int bkt = c <= 0 ? 0 : min(63, 2 * (31 - __clz(c)) + ((c >> max(0, 30 - __clz(c))) & 1));
    order[atomicAdd(&start[bkt], 1)] = i;
  }
}

// Forget the persistent state of a list of episode slots: one block per listed episode (bmpc_reset with ids).
struct ResetArgs {
  real* uLin; int* pbest; real* oldin; real* xprev; int* started; int* cache_state; int* cost;
  int ulin_reals, nbranch, d, xprev_reals;
};
__global__ void bmpc_reset_kernel(const ResetArgs a, const long long* __restrict__ ids, int count) {
  const int i = blockIdx.x;
  if (i >= count) return;
  const size_t e = (size_t)ids[i];
  for (int q = threadIdx.x; q < a.ulin_reals; q += blockDim.x) a.uLin[e * a.ulin_reals + q] = 0.0;
  for (int q = threadIdx.x; q < a.nbranch; q += blockDim.x) a.pbest[e * a.nbranch + q] = 0;
  for (int q = threadIdx.x; q < a.d; q += blockDim.x) a.oldin[e * a.d + q] = 0.0;
  if (a.xprev)
    for (int q = threadIdx.x; q < a.xprev_reals; q += blockDim.x) a.xprev[e * a.xprev_reals + q] = 0.0;
  if (threadIdx.x == 0) {
    a.started[e] = 0;
    a.cache_state[2 * e] = -1;
    a.cache_state[2 * e + 1] = -1;
    a.cost[e] = 0;
  }
}

// Point-wise model functions for the parity tests of rows M1-M5 (one thread per point).
struct EvalArgs {
  const real *x, *z, *u, *polpar;
  real *A, *B, *C, *xp, *zpred, *p, *hlin, *dh;
  int count;
};

template <class M>
__global__ void bmpc_eval_kernel(const __grid_constant__ KParams P, const EvalArgs a) {
  constexpr int NX = M::NX, NU = M::NU;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.count) return;
  const real* x = a.x + (size_t)i * NX;
  if (a.u && (a.A || a.B || a.C || a.xp)) {
    const real* u = a.u + (size_t)i * NU;
    real lin[M::NLIN], cc[M::NCC], xn[NX];
    M::linearize(P, x, u, lin, cc, xn);
    if (a.A) M::denseA(P, lin, a.A + (size_t)i * NX * NX);
    if (a.B) M::denseB(P, lin, a.B + (size_t)i * NX * NU);
    if (a.C) M::expandC(cc, a.C + (size_t)i * NX);
    if (a.xp)
      for (int q = 0; q < NX; ++q) a.xp[(size_t)i * NX + q] = xn[q];
  }
  if (a.z && (a.hlin || a.dh)) {
    const real* z = a.z + (size_t)i * NX;
    real h, dhx, dhy;
    M::collision(P, x, z, h, dhx, dhy);
    if (a.hlin) a.hlin[i] = h - (dhx * x[0] + dhy * x[1]);
    if (a.dh) {
      for (int q = 0; q < NX; ++q) a.dh[(size_t)i * NX + q] = 0.0;
      a.dh[(size_t)i * NX] = dhx;
      a.dh[(size_t)i * NX + 1] = dhy;
    }
  }
  if (a.z && (a.zpred || a.p)) {
    const real* z = a.z + (size_t)i * NX;
    real hi[BMPC_MAX_POLICIES];
    real himax = -1e300;
    for (int k = 0; k < P.zm; ++k) {
      const real* par = a.polpar ? a.polpar + ((size_t)i * P.zm + k) * 4 : P.pol_par[k];
      const real* par0 = a.polpar ? a.polpar + ((size_t)i * P.zm) * 4 : P.pol_par[0];
      real zl[NX];
      real* zo = a.zpred ? a.zpred + (size_t)i * P.zN * P.zm * NX : nullptr;
      hi[k] = M::policy_safety(P, P.pol_kind[k], par, P.pol_kind[0], par0, x, z, zl, P.zN, [&](int t, const real* zz) {
        if (zo)
          for (int q = 0; q < NX; ++q) zo[((size_t)t * P.zm + k) * NX + q] = zz[q];
      });
      himax = fmax(himax, hi[k]);
    }
    if (a.p) {
      real sum = 0.0;
      for (int k = 0; k < P.zm; ++k) sum += M::branch_weight(P, hi[k], himax);
      for (int k = 0; k < P.zm; ++k) a.p[(size_t)i * P.zm + k] = M::branch_weight(P, hi[k], himax) / sum;
    }
  }
}

// Point-wise belief-state model: HMM_backup_dyn.PredictiveModel.regressionAndLinearization (:216-229) for a batch of points.
struct BeliefEvalArgs {
  const real *xb, *xbackup, *u;
  real *A, *B, *C, *h0, *Jh, *xbp;
  int count;
};
__global__ void bmpc_belief_eval_kernel(const __grid_constant__ KParams P, const BeliefEvalArgs a) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= a.count) return;
  const int M = P.hmm_M, m = P.zm, nb = M * m, n = 4 + nb;
  const real* xb = a.xb + (size_t)e * n;
  const real* xbk = a.xbackup + (size_t)e * nb * 4;
  const real* u = a.u + (size_t)e * 2;
  real lin[HighwayModel::NLIN], cc[HighwayModel::NCC], xn[4], bp[9];
  HighwayModel::linearize(P, xb, u, lin, cc, xn);
  real* A = a.A + (size_t)e * n * n;
  for (int q = 0; q < n * n; ++q) A[q] = 0.0;
  real A4[16];
  HighwayModel::denseA(P, lin, A4);
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) A[i * n + j] = A4[i * 4 + j];
  // belief rows, agent by agent (same formulas as BeliefModel::transition, written straight into the dense matrix)
  const real tau = P.hmm_tran_diag;
  for (int i = 0; i < M; ++i) {
    real mh[BMPC_MAX_POLICIES], gx[BMPC_MAX_POLICIES], gy[BMPC_MAX_POLICIES], msum = 0.0, bsum = 0.0;
    for (int j = 0; j < m; ++j) {
      const real hv = BeliefModel::safety(P, xb, xbk + (size_t)(m * i + j) * 4, gx[j], gy[j]);
      mh[j] = 1.0 / (1.0 + exp(-P.s1 * hv));
      msum += mh[j];
      bsum += xb[4 + j * M + i];
    }
    for (int k = 0; k < m; ++k) {
      const int q = k * M + i;
      const real pik = mh[k] / msum;
      bp[q] = (1.0 - tau) * bsum * pik + tau * xb[4 + q];
      for (int r = 0; r < m; ++r) A[(4 + q) * n + 4 + r * M + i] = (1.0 - tau) * pik + (r == k ? tau : 0.0);
      real ax = 0.0, ay = 0.0;
      for (int l = 0; l < m; ++l) {
        const real c = (1.0 - tau) * bsum * ((((l == k) ? 1.0 : 0.0) - pik) / msum) * P.s1 * mh[l] * (1.0 - mh[l]);
        ax += c * gx[l];
        ay += c * gy[l];
      }
      A[(4 + q) * n] = ax;
      A[(4 + q) * n + 1] = ay;
    }
  }
  real* B = a.B + (size_t)e * n * 2;
  for (int q = 0; q < n * 2; ++q) B[q] = 0.0;
  B[2 * 2] = P.dt;
  B[3 * 2 + 1] = P.dt;
  real* xbp = a.xbp + (size_t)e * n;
  for (int i = 0; i < 4; ++i) xbp[i] = xn[i];
  for (int q = 0; q < nb; ++q) xbp[4 + q] = bp[q];
  real* C = a.C + (size_t)e * n;
  for (int i = 0; i < n; ++i) {
    real v = xbp[i];
    for (int j = 0; j < n; ++j) v -= A[i * n + j] * xb[j];
    v -= B[i * 2] * u[0] + B[i * 2 + 1] * u[1];
    C[i] = v;
  }
  for (int i = 0; i < M; ++i)
    for (int j = 0; j < m; ++j) {
      real gx, gy;
      const real h = BeliefModel::safety(P, xb, xbk + (size_t)(m * i + j) * 4, gx, gy);
      a.h0[(size_t)e * nb + i * m + j] = h - (gx * xb[0] + gy * xb[1]);
      a.Jh[((size_t)e * nb + i * m + j) * 2] = gx;
      a.Jh[((size_t)e * nb + i * m + j) * 2 + 1] = gy;
    }
}

// Euler plant step of ego (applied input) and obstacle (one of the backup policies), one thread per episode.
template <class M>
__global__ void bmpc_plant_kernel(const __grid_constant__ KParams P, real* x, const real* u, real* z, int pol,
                                  const real* polpar, int count) {
  constexpr int NX = M::NX, NU = M::NU;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  if (x && u) {
    real xs[NX], us[NU], xn[NX];
    for (int q = 0; q < NX; ++q) xs[q] = x[(size_t)i * NX + q];
    for (int a = 0; a < NU; ++a) us[a] = u[(size_t)i * NU + a];
    M::step(P, xs, us, xn);
    for (int q = 0; q < NX; ++q) x[(size_t)i * NX + q] = xn[q];
  }
  if (z) {
    real zs[NX], us[NU], zn[NX];
    for (int q = 0; q < NX; ++q) zs[q] = z[(size_t)i * NX + q];
    const real* par = polpar ? polpar + ((size_t)i * P.m + pol) * 4 : P.pol_par[pol];
    M::policy(P, P.pol_kind[pol], par, zs, us);
    M::step(P, zs, us, zn);
    for (int q = 0; q < NX; ++q) z[(size_t)i * NX + q] = zn[q];
  }
}

// Dependent-chain-free DFMA loop: 8 independent accumulators per thread.
__global__ void bmpc_dfma_kernel(double* out, int iters) {
  double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double b = 1.0000001, c = 1e-9;
  for (int i = 0; i < iters; ++i) {
    a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
    a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

// ------------------------------------------------------------------------------------------------------------
// handle
// ------------------------------------------------------------------------------------------------------------
#define BMPC_PSTAGE 8
#define BMPC_MAX_CAPTURED 32   // launches of one handle that may be recorded into CUDA graphs
struct bmpc_handle {
  bmpc_config cfg;
  KParams P;            // call-independent part
  int device = 0;
  int num_sms = 0;
  size_t slab_bytes = 0;
  int mode = BMPC_SLAB_SHARED;   // resolved BMPC_SLAB_* placement
  size_t gws_bytes_per_warp = 0;
  int grid = 0;         // persistent warps (= blocks)
  // persistent per-episode state
  real* uLin = nullptr;
  int* pbest = nullptr;
  real* oldin = nullptr;
  real* xprev = nullptr;   // robustMPC only: previous predicted states
  int* started = nullptr;
  // solver caches (not part of the reference-visible state): rho of the last refresh, active set of the last optimum
  real* rho_cache = nullptr;
  long long* code_cache = nullptr;
  int* cache_state = nullptr;
  int* cost = nullptr;    // cycles >> 10 of each episode's previous solve (0 = unknown)
  int* order = nullptr;   // work order of the current launch
  int* counter = nullptr;
  real* gws = nullptr;
  real* ipm_ws = nullptr;   // per-warp scratch of the interior-point fallback
  real* nu_cache = nullptr; // BranchMPC_CVaR: risk multipliers of each episode's last step
  real* cv_ws = nullptr;    // BranchMPC_CVaR: per-team scratch of the master problem
  real* lut = nullptr;      // lookup table of the *_REF policies: grid then values
  int team_lanes = 96;      // threads per block = lanes of one team
  real* bel_ws = nullptr;   // belief-state MPC: per-team linearisation trajectory of the augmented state
  KParams* captured = nullptr;            // pinned parameter blocks for launches recorded into CUDA graphs (BMPC_MAX_CAPTURED,
  int captured_used = 0;                  // allocated at create: nothing may be allocated while a stream is capturing)
  KParams* pstage = nullptr;              // pinned staging ring of parameter blocks (source of the constant-memory upload)
  cudaEvent_t pstage_evt[BMPC_PSTAGE] = {};
  unsigned pstage_next = 0;
  // staging for bmpc_solve_host: device and pinned host mirrors, one DMA each way per call
  real* stage_in = nullptr;   // x0 | z0 | xref | polpar
  real* stage_in_host = nullptr;
  void* stage_out = nullptr;
  void* stage_out_host[2] = {nullptr, nullptr};   // alternating: the views of one call survive the next call
  unsigned stage_out_turn = 0;
  size_t stage_out_bytes = 0;
  long long* reset_ids = nullptr;   // device copy of the ids of the last bmpc_reset
  size_t reset_ids_cap = 0;
  cudaStream_t last_stream = nullptr;   // stream of the last solve (persistent-state accessors order themselves behind it)
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  bool timed = false;
  int64_t launches = 0;
  std::string err;
};

static std::string g_create_error;

struct ConstSlot { cudaEvent_t done = nullptr; cudaStream_t stream = nullptr; bool used = false; std::mutex mu; };
static ConstSlot g_const[16];   // per device: the last solve launch that reads bmpc_cP

#define BMPC_CK(h, call)                                                                           \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess) {                                                                       \
      (h)->err = std::string(#call) + ": " + cudaGetErrorString(e_);                               \
      return BMPC_E_CUDA;                                                                          \
    }                                                                                              \
  } while (0)

template <class M, int NR, int MODE, int NC>
static int try_mode(bmpc_handle* h, int max_optin, int* per_sm) {
  using S = Solver<M, NR, MODE, NC>;
  const size_t slab = S::slab_reals(h->P.nup, h->P.nbx) * sizeof(real);
  const size_t smem = (MODE == BMPC_SLAB_GLOBAL) ? 0 : slab;
  *per_sm = 0;
  if (smem > (size_t)max_optin) return BMPC_OK;   // does not fit: caller falls through to the next mode
  if (smem > 0)
    BMPC_CK(h, cudaFuncSetAttribute(bmpc_solve_kernel<M, NR, MODE, NC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  BMPC_CK(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(per_sm, bmpc_solve_kernel<M, NR, MODE, NC>, 96, smem));
  return BMPC_OK;
}

template <class M, int NR, int NC = 1>
static int configure_instance(bmpc_handle* h) {
  int max_optin = 0;
  BMPC_CK(h, cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, h->device));
  int want = h->cfg.slab_mode;
  if (want == BMPC_SLAB_AUTO) want = BMPC_SLAB_SHARED;   // measured: the shared slab wins whenever it fits (profiles/)
  int per_sm = 0, rc = BMPC_OK;
  h->mode = 0;
  if (want == BMPC_SLAB_SHARED) {
    rc = try_mode<M, NR, BMPC_SLAB_SHARED, NC>(h, max_optin, &per_sm);
    if (rc != BMPC_OK) return rc;
    if (per_sm > 0) h->mode = BMPC_SLAB_SHARED;
  }
  if (h->mode == 0 && want != BMPC_SLAB_GLOBAL) {
    rc = try_mode<M, NR, BMPC_SLAB_SPLIT, NC>(h, max_optin, &per_sm);
    if (rc != BMPC_OK) return rc;
    if (per_sm > 0) h->mode = BMPC_SLAB_SPLIT;
  }
  if (h->mode == 0) {
    rc = try_mode<M, NR, BMPC_SLAB_GLOBAL, NC>(h, max_optin, &per_sm);
    if (rc != BMPC_OK) return rc;
    if (per_sm > 8) per_sm = 8;   // keep the global slabs of the resident warps inside L2
    if (per_sm > 0) h->mode = BMPC_SLAB_GLOBAL;
  }
  if (per_sm < 1 || h->mode == 0) { h->err = "kernel does not fit on an SM"; return BMPC_E_CUDA; }
  if (h->cfg.reserved[1] > 0 && per_sm > h->cfg.reserved[1]) per_sm = h->cfg.reserved[1];   // occupancy cap (experiments)
  h->grid = per_sm * h->num_sms;
  // Team size: three warps when four teams share an SM (the default highway tree); trees that leave the SM emptier get more
  // lanes for their node-parallel passes, as far as nodes, registers (168 per thread) and the placement's occupancy allow.
  h->team_lanes = 96;
  {
    int want_lanes = per_sm >= 3 ? 96 : (per_sm == 2 ? 192 : 256);   // measured (profiles/r02_staging_ab.md)
    if (const char* e = getenv("BMPC_TEAM_LANES")) want_lanes = atoi(e);   // experiments
    const int node_lanes = ((h->P.totalu + 31) / 32) * 32;
    if (want_lanes > node_lanes) want_lanes = node_lanes < 96 ? 96 : node_lanes;
    if (want_lanes < 32 || want_lanes > 32 * BMPC_MAX_TEAM_WARPS || (want_lanes & 31)) want_lanes = 96;
    if (want_lanes != 96) {
      int fit = 0;
      const size_t smem_now = (h->mode == BMPC_SLAB_GLOBAL) ? 0 : (h->mode == BMPC_SLAB_SPLIT
          ? Solver<M, NR, BMPC_SLAB_SPLIT, NC>::slab_reals(h->P.nup, h->P.nbx) : Solver<M, NR, BMPC_SLAB_SHARED, NC>::slab_reals(h->P.nup, h->P.nbx)) * sizeof(real);
      cudaError_t e2 = h->mode == BMPC_SLAB_SHARED
          ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&fit, bmpc_solve_kernel<M, NR, BMPC_SLAB_SHARED, NC>, want_lanes, smem_now)
          : h->mode == BMPC_SLAB_SPLIT
              ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&fit, bmpc_solve_kernel<M, NR, BMPC_SLAB_SPLIT, NC>, want_lanes, smem_now)
              : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&fit, bmpc_solve_kernel<M, NR, BMPC_SLAB_GLOBAL, NC>, want_lanes, smem_now);
      if (e2 == cudaSuccess && fit >= per_sm) h->team_lanes = want_lanes;
    }
  }
  if (h->mode == BMPC_SLAB_SPLIT) {
    using S = Solver<M, NR, BMPC_SLAB_SPLIT, NC>;
    h->P.slab_reals = S::slab_reals(h->P.nup, h->P.nbx);
    h->P.factor_reals = S::factor_reals(h->P.nup);
    h->gws_bytes_per_warp = h->P.factor_reals * sizeof(real);
  } else {
    using S = Solver<M, NR, BMPC_SLAB_GLOBAL, NC>;
    h->P.slab_reals = S::slab_reals(h->P.nup, h->P.nbx);
    h->P.factor_reals = 0;
    h->gws_bytes_per_warp = (h->mode == BMPC_SLAB_GLOBAL) ? h->P.slab_reals * sizeof(real) : 0;
  }
  h->slab_bytes = (h->mode == BMPC_SLAB_GLOBAL) ? 0 : h->P.slab_reals * sizeof(real);
  h->P.ipm_reals = Solver<M, NR, BMPC_SLAB_SHARED, NC>::ipm_reals(h->P.nup);
  // Staging of the next episode's state (uLin | codes | rho cache) behind the slab by bulk copies (cp.async.bulk + mbarrier):
  // shared placement, tree controllers.  Opt-in (reserved[6] bit 1): measured, it does not pay - the data it prefetches is
  // 8 KB per episode out of L2, read once by 96 lanes, while the hand-over (claim, fences, three copies, one more team
  // barrier per solve) costs 2-10 % on small trees, and on the default highway tree the 6 KB of extra shared memory moves the
  // SM to the next carve-out step, which alone costs 3 % in L1 capacity (profiles/r02_staging_ab.md).
  h->P.stage_on = 0;
  {
    using S = Solver<M, NR, BMPC_SLAB_SHARED, NC>;
    if (h->mode == BMPC_SLAB_SHARED && !bmpc_is_chain(h->cfg.controller) && S::stage_possible(h->P.totalu) &&
        (h->cfg.reserved[6] & 2) != 0) {
      const size_t with_stage = (S::stage_offset(h->P.nup, h->P.nbx) + S::stage_reals(h->P.totalu)) * sizeof(real);
      int fit = 0;
      if (with_stage <= (size_t)max_optin) {
        BMPC_CK(h, cudaFuncSetAttribute(bmpc_solve_kernel<M, NR, BMPC_SLAB_SHARED, NC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)with_stage));
        BMPC_CK(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&fit, bmpc_solve_kernel<M, NR, BMPC_SLAB_SHARED, NC>, h->team_lanes, with_stage));
      }
      const int resident = h->grid / h->num_sms;
      if (fit >= resident) {
        h->P.stage_on = 1;
        h->slab_bytes = with_stage;
      }
    }
  }
#ifdef BMPC_DEV_EXTRA_SMEM
  if (!h->P.stage_on && h->mode == BMPC_SLAB_SHARED) h->slab_bytes += BMPC_DEV_EXTRA_SMEM;   // experiment: shared-memory carve-out vs L1
#endif
  return BMPC_OK;
}

template <class M, int NR, int NC = 1>
static int launch_instance(bmpc_handle* h, const KParams& P, int grid, cudaStream_t s) {
  // the dynamic shared-memory limit is a property of the kernel instance, not of the handle: another handle of the same
  // instance with a smaller tree may have lowered it since this one was created
  if (h->mode == BMPC_SLAB_SHARED) {
    BMPC_CK(h, cudaFuncSetAttribute(bmpc_solve_kernel<M, NR, BMPC_SLAB_SHARED, NC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->slab_bytes));
    bmpc_solve_kernel<M, NR, BMPC_SLAB_SHARED, NC><<<grid, h->team_lanes, h->slab_bytes, s>>>();
  } else if (h->mode == BMPC_SLAB_SPLIT) {
    BMPC_CK(h, cudaFuncSetAttribute(bmpc_solve_kernel<M, NR, BMPC_SLAB_SPLIT, NC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->slab_bytes));
    bmpc_solve_kernel<M, NR, BMPC_SLAB_SPLIT, NC><<<grid, h->team_lanes, h->slab_bytes, s>>>();
  } else {
    bmpc_solve_kernel<M, NR, BMPC_SLAB_GLOBAL, NC><<<grid, h->team_lanes, 0, s>>>();
  }
  BMPC_CK(h, cudaGetLastError());
  return BMPC_OK;
}

#define BMPC_DISPATCH_M(h, fn, HW, QD, ...)                                              \
  ((h)->cfg.model == BMPC_MODEL_HIGHWAY                                                  \
       ? ((h)->cfg.n_rows == 0 ? fn<HW, 1>(__VA_ARGS__)                                  \
                               : (h)->cfg.n_rows == 1 ? fn<HW, 2>(__VA_ARGS__)           \
                                                      : fn<HW, 3>(__VA_ARGS__))          \
       : fn<QD, 1>(__VA_ARGS__))
// BranchMPCProx carries the previous input through the Riccati state (RateAug); BranchMPC does not need to.
// robustMPC: highway model, 2 state rows + up to 9 obstacle nodes per time slot
#define BMPC_DISPATCH(h, fn, ...)                                                                       \
  (bmpc_is_chain((h)->cfg.controller) ? fn<HighwayModel, 11, 9>(__VA_ARGS__) :                     \
   (h)->cfg.model == BMPC_MODEL_MERGE ? fn<MergeModel, 3>(__VA_ARGS__) :                                \
   (h)->cfg.controller == BMPC_CTRL_PROX                                                                \
       ? BMPC_DISPATCH_M(h, fn, RateAug<HighwayModel>, RateAug<QuadrupedModel>, __VA_ARGS__)            \
       : BMPC_DISPATCH_M(h, fn, HighwayModel, QuadrupedModel, __VA_ARGS__))

#ifdef BMPC_DEV_HIGHWAY_ONLY
// development builds (tools/build_variant.sh ... -DBMPC_DEV_HIGHWAY_ONLY): only the default highway instance, seconds to compile
#undef BMPC_DISPATCH
#define BMPC_DISPATCH(h, fn, ...) fn<HighwayModel, 3>(__VA_ARGS__)
#endif

static void free_handle(bmpc_handle* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  cudaFree(h->uLin);
  cudaFree(h->pbest);
  cudaFree(h->oldin);
  cudaFree(h->xprev);
  cudaFree(h->started);
  cudaFree(h->rho_cache);
  cudaFree(h->code_cache);
  cudaFree(h->cache_state);
  cudaFree(h->cost);
  cudaFree(h->order);
  cudaFree(h->counter);
  cudaFree(h->gws);
  cudaFree(h->ipm_ws);
  cudaFree(h->nu_cache);
  cudaFree(h->cv_ws);
  cudaFree(h->lut);
  cudaFree(h->bel_ws);
  if (h->pstage) cudaFreeHost(h->pstage);
  if (h->captured) cudaFreeHost(h->captured);
  for (int i = 0; i < BMPC_PSTAGE; ++i) if (h->pstage_evt[i]) cudaEventDestroy(h->pstage_evt[i]);
  cudaFree(h->stage_in);
  cudaFree(h->stage_out);
  cudaFree(h->reset_ids);
  if (h->stage_in_host) cudaFreeHost(h->stage_in_host);
  for (int i = 0; i < 2; ++i)
    if (h->stage_out_host[i]) cudaFreeHost(h->stage_out_host[i]);
  if (h->ev0) cudaEventDestroy(h->ev0);
  if (h->ev1) cudaEventDestroy(h->ev1);
  delete h;
}

static int create_impl(const bmpc_config* cfg, bmpc_handle* h) {
  h->cfg = *cfg;
  int rc = bmpc::make_params(*cfg, &h->P, &h->err);
  if (rc != BMPC_OK) return rc;
  if (!bmpc::supported_instance(cfg->model, cfg->n_rows, cfg->controller, h->P.zpw[h->P.zNB])) {
    h->err = "no kernel instance for this (model, n_rows, controller)";
    return BMPC_E_UNSUPPORTED;
  }
  if (cfg->batch_capacity < 1) { h->err = "batch_capacity must be >= 1"; return BMPC_E_INVALID; }
  h->device = cfg->device;
  BMPC_CK(h, cudaSetDevice(h->device));
  BMPC_CK(h, cudaDeviceGetAttribute(&h->num_sms, cudaDevAttrMultiProcessorCount, h->device));
  rc = BMPC_DISPATCH(h, configure_instance, h);
  if (rc != BMPC_OK) return rc;
  const size_t cap = (size_t)cfg->batch_capacity;
  const KParams& P = h->P;
  BMPC_CK(h, cudaMalloc(&h->uLin, cap * (P.totalu + 1) * cfg->d * sizeof(real)));
  BMPC_CK(h, cudaMalloc(&h->pbest, cap * P.nbranch * sizeof(int)));
  BMPC_CK(h, cudaMalloc(&h->oldin, cap * cfg->d * sizeof(real)));
  BMPC_CK(h, cudaMalloc(&h->started, cap * sizeof(int)));
  if (cfg->controller == BMPC_CTRL_ROBUST) BMPC_CK(h, cudaMalloc(&h->xprev, cap * P.pub_totalx * cfg->n * sizeof(real)));
  BMPC_CK(h, cudaMalloc(&h->rho_cache, cap * P.totalu * 16 * sizeof(real))  /* up to 11 rows + 3 inputs per node */);
  BMPC_CK(h, cudaMalloc(&h->code_cache, cap * (P.totalu + 1) * sizeof(long long)));   // per-episode stride padded to 16 bytes
  BMPC_CK(h, cudaMalloc(&h->cache_state, cap * 2 * sizeof(int)));
  BMPC_CK(h, cudaMalloc(&h->cost, cap * sizeof(int)));
  BMPC_CK(h, cudaMalloc(&h->order, cap * sizeof(int)));
  BMPC_CK(h, cudaMalloc(&h->counter, sizeof(int)));
  if (h->gws_bytes_per_warp) BMPC_CK(h, cudaMalloc(&h->gws, (size_t)h->grid * h->gws_bytes_per_warp));
  BMPC_CK(h, cudaMalloc(&h->ipm_ws, (size_t)h->grid * h->P.ipm_reals * sizeof(real)));
  if (cfg->controller == BMPC_CTRL_BELIEF) BMPC_CK(h, cudaMalloc(&h->bel_ws, (size_t)h->grid * h->P.bel_reals * sizeof(real)));
  if (cfg->controller == BMPC_CTRL_CVAR) {
    BMPC_CK(h, cudaMalloc(&h->nu_cache, cap * P.nbranch * sizeof(real)));
    BMPC_CK(h, cudaMemset(h->nu_cache, 0, cap * P.nbranch * sizeof(real)));
    BMPC_CK(h, cudaMalloc(&h->cv_ws, (size_t)h->grid * h->P.cv_reals * sizeof(real)));
  }
  BMPC_CK(h, cudaMallocHost(&h->pstage, BMPC_PSTAGE * sizeof(KParams)));
  BMPC_CK(h, cudaMallocHost(&h->captured, BMPC_MAX_CAPTURED * sizeof(KParams)));
  BMPC_CK(h, cudaEventCreate(&h->ev0));
  BMPC_CK(h, cudaEventCreate(&h->ev1));
  return bmpc_reset(h, nullptr, 0);
}

extern "C" {

int bmpc_version(void) { return BMPC_VERSION; }

int bmpc_create(const bmpc_config* cfg, bmpc_handle** out) {
  if (!cfg || !out) { g_create_error = "null argument"; return BMPC_E_INVALID; }
  *out = nullptr;
  bmpc_handle* h = new (std::nothrow) bmpc_handle();
  if (!h) { g_create_error = "out of host memory"; return BMPC_E_INVALID; }
  const int rc = create_impl(cfg, h);
  if (rc != BMPC_OK) {
    g_create_error = h->err;
    free_handle(h);
    return rc;
  }
  *out = h;
  return BMPC_OK;
}

int bmpc_destroy(bmpc_handle* h) {
  free_handle(h);
  return BMPC_OK;
}

// The accessors of the persistent state (reset / get_state / set_state) are ordered behind the handle's last solve:
// they run on that solve's stream (the legacy default stream before the first solve).
int bmpc_reset(bmpc_handle* h, const int64_t* episode_ids, int64_t count) {
  if (!h) return BMPC_E_INVALID;
  BMPC_CK(h, cudaSetDevice(h->device));
  const KParams& P = h->P;
  const size_t cap = (size_t)h->cfg.batch_capacity;
  const size_t ulin_row = (size_t)(P.totalu + 1) * h->cfg.d * sizeof(real);
  const size_t xprev_row = (size_t)P.pub_totalx * h->cfg.n * sizeof(real);
  cudaStream_t s = h->last_stream;
  if (!episode_ids) {
    BMPC_CK(h, cudaMemsetAsync(h->uLin, 0, cap * ulin_row, s));
    BMPC_CK(h, cudaMemsetAsync(h->pbest, 0, cap * P.nbranch * sizeof(int), s));
    BMPC_CK(h, cudaMemsetAsync(h->oldin, 0, cap * h->cfg.d * sizeof(real), s));
    BMPC_CK(h, cudaMemsetAsync(h->started, 0, cap * sizeof(int), s));
    BMPC_CK(h, cudaMemsetAsync(h->cache_state, 0xff, cap * 2 * sizeof(int), s));
    BMPC_CK(h, cudaMemsetAsync(h->cost, 0, cap * sizeof(int), s));
    if (h->xprev) BMPC_CK(h, cudaMemsetAsync(h->xprev, 0, cap * xprev_row, s));
    return BMPC_OK;
  }
  if (count <= 0) return BMPC_OK;
  for (int64_t i = 0; i < count; ++i)
    if (episode_ids[i] < 0 || (size_t)episode_ids[i] >= cap) { h->err = "episode id out of range"; return BMPC_E_INVALID; }
  if ((size_t)count > h->reset_ids_cap) {
    BMPC_CK(h, cudaStreamSynchronize(s));
    cudaFree(h->reset_ids);
    h->reset_ids = nullptr;
    h->reset_ids_cap = 0;
    const size_t want = (size_t)count > cap ? (size_t)count : cap;
    BMPC_CK(h, cudaMalloc(&h->reset_ids, want * sizeof(long long)));
    h->reset_ids_cap = want;
  }
  // pageable source: the copy has returned from the host buffer when the call returns
  BMPC_CK(h, cudaMemcpyAsync(h->reset_ids, episode_ids, count * sizeof(long long), cudaMemcpyHostToDevice, s));
  ResetArgs a{h->uLin, h->pbest, h->oldin, h->xprev, h->started, h->cache_state, h->cost,
              (int)(ulin_row / sizeof(real)), P.nbranch, h->cfg.d, (int)(xprev_row / sizeof(real))};
  bmpc_reset_kernel<<<(int)count, 64, 0, s>>>(a, h->reset_ids, (int)count);
  BMPC_CK(h, cudaGetLastError());
  h->launches += 1;
  return BMPC_OK;
}

int bmpc_num_branches(const bmpc_handle* h) { return h ? h->P.nbranch : BMPC_E_INVALID; }
int bmpc_total_x(const bmpc_handle* h) { return h ? h->P.pub_totalx : BMPC_E_INVALID; }
int bmpc_total_u(const bmpc_handle* h) { return h ? h->P.pub_totalu : BMPC_E_INVALID; }

int bmpc_ulin_rows(const bmpc_handle* h) { return h ? h->P.totalu + 1 : BMPC_E_INVALID; }

int bmpc_get_topology(const bmpc_handle* h, int32_t* ndx, int32_t* ndu, int32_t* depth, int32_t* parent) {
  if (!h) return BMPC_E_INVALID;
  const KParams& P = h->P;
  for (int b = 0; b < P.nbranch; ++b) {
    const int d = bmpc_depth(P, b);
    if (ndx) ndx[b] = bmpc_ndx(P, b);
    if (ndu) ndu[b] = bmpc_ndu(P, b);
    if (depth) depth[b] = d;
    if (parent) parent[b] = d == 0 ? -1 : bmpc_parent(P, b, d);
  }
  return BMPC_OK;
}

struct BeliefArgs { const double* b0; const double* xbackup; int cols; };
struct XformArgs { const double* S; const double* bounds; };
static int solve_impl(bmpc_handle* h, const double* x0, const double* z0, const double* xref, const double* policy_params,
                      int64_t count, const bmpc_outputs* out, void* stream, const BeliefArgs* bel, const XformArgs* xf = nullptr);

int bmpc_solve(bmpc_handle* h, const double* x0, const double* z0, const double* xref, const double* policy_params,
               int64_t count, const bmpc_outputs* out, void* stream) {
  if (!h) return BMPC_E_INVALID;
  if (h->cfg.controller == BMPC_CTRL_BELIEF) { h->err = "the belief-state MPC is solved through bmpc_solve_belief"; return BMPC_E_INVALID; }
  if (!z0) { h->err = "null argument"; return BMPC_E_INVALID; }
  return solve_impl(h, x0, z0, xref, policy_params, count, out, stream, nullptr);
}

int bmpc_solve_belief(bmpc_handle* h, const double* x0, const double* b0, const double* xbackup, int32_t xbackup_cols,
                      const double* xref, int64_t count, const bmpc_outputs* out, void* stream) {
  if (!h) return BMPC_E_INVALID;
  if (h->cfg.controller != BMPC_CTRL_BELIEF) { h->err = "not a belief-state MPC handle"; return BMPC_E_INVALID; }
  if (!b0 || !xbackup || xbackup_cols < 4 * h->cfg.N) { h->err = "b0, xbackup with at least 4 N columns per row are required"; return BMPC_E_INVALID; }
  const BeliefArgs bel{b0, xbackup, xbackup_cols};
  return solve_impl(h, x0, x0, xref, nullptr, count, out, stream, &bel);
}

int bmpc_set_lookup_table(bmpc_handle* h, const double* xs, const double* ys, int32_t n) {
  if (!h) return BMPC_E_INVALID;
  if (!xs || !ys || n < 2 || n > 4096) { h->err = "lookup table needs 2..4096 points"; return BMPC_E_INVALID; }
  for (int i = 1; i < n; ++i)
    if (!(xs[i] > xs[i - 1])) { h->err = "lookup grid must be strictly increasing"; return BMPC_E_INVALID; }
  BMPC_CK(h, cudaSetDevice(h->device));
  BMPC_CK(h, cudaStreamSynchronize(h->last_stream));   // a solve in flight may still read the old table
  real* t = nullptr;
  BMPC_CK(h, cudaMalloc(&t, (size_t)2 * n * sizeof(real)));
  cudaError_t e = cudaMemcpy(t, xs, (size_t)n * sizeof(real), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(t + n, ys, (size_t)n * sizeof(real), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) { cudaFree(t); h->err = cudaGetErrorString(e); return BMPC_E_CUDA; }
  cudaFree(h->lut);
  h->lut = t;
  h->P.lut_x = t;
  h->P.lut_y = t + n;
  h->P.lut_n = n;
  return BMPC_OK;
}

static bool needs_lookup(const bmpc_handle* h) {
  for (int i = 0; i < h->cfg.m; ++i)
    if (h->cfg.policy_kind[i] == BMPC_POLICY_TRACKV_REF || h->cfg.policy_kind[i] == BMPC_POLICY_BRAKE_REF) return h->P.lut_n < 2;
  return false;
}

int bmpc_solve_transformed(bmpc_handle* h, const double* x0, const double* z0, const double* xref, const double* policy_params,
                           const double* S, const double* state_bounds, int64_t count, const bmpc_outputs* out, void* stream) {
  if (!h) return BMPC_E_INVALID;
  if (h->cfg.model != BMPC_MODEL_MERGE) { h->err = "state transforms are built for BMPC_MODEL_MERGE handles"; return BMPC_E_INVALID; }
  if (!z0) { h->err = "null argument"; return BMPC_E_INVALID; }
  const XformArgs xf{S, state_bounds};
  return solve_impl(h, x0, z0, xref, policy_params, count, out, stream, nullptr, &xf);
}

static int solve_impl(bmpc_handle* h, const double* x0, const double* z0, const double* xref, const double* policy_params,
                      int64_t count, const bmpc_outputs* out, void* stream, const BeliefArgs* bel, const XformArgs* xf) {
  if (!h) return BMPC_E_INVALID;
  if (!x0 || !z0 || !xref || !out || count < 0) { h->err = "null argument"; return BMPC_E_INVALID; }
  if (count > h->cfg.batch_capacity) { h->err = "count exceeds batch_capacity"; return BMPC_E_CAPACITY; }
  if (needs_lookup(h)) { h->err = "the policy table uses a lookup-table policy: call bmpc_set_lookup_table first"; return BMPC_E_INVALID; }
  if (count == 0) return BMPC_OK;
  BMPC_CK(h, cudaSetDevice(h->device));
  cudaStream_t s = (cudaStream_t)stream;
  KParams P = h->P;
  P.count = (int)count;
  P.x0 = x0;
  P.z0 = z0;
  P.xref = xref;
  P.polpar = policy_params;
  P.xform = xf ? xf->S : nullptr;
  P.xbounds = xf ? xf->bounds : nullptr;
  P.uLin = h->uLin;
  P.pbest = h->pbest;
  P.oldin = h->oldin;
  P.xprev = h->xprev;
  P.started = h->started;
  P.rho_cache = h->rho_cache;
  P.code_cache = h->code_cache;
  P.cache_state = h->cache_state;
  P.cost = h->cost;
  P.order = nullptr;
  if (count > h->grid && (h->cfg.reserved[6] & 1) == 0) {
    bmpc_order_kernel<<<1, 1024, 0, s>>>(h->cost, (int)count, h->order);
    BMPC_CK(h, cudaGetLastError());
    P.order = h->order;
    h->launches += 1;
  }
  P.out = *out;
  P.counter = h->counter;
  P.gws = h->gws;
  P.ipm = h->ipm_ws;
  P.nu_cache = h->nu_cache;
  P.cv = h->cv_ws;
  P.bel = h->bel_ws;
  if (bel) { P.b0 = bel->b0; P.xbackup = bel->xbackup; P.xb_cols = bel->cols; }
  BMPC_CK(h, cudaMemsetAsync(h->counter, 0, sizeof(int), s));
  // rows of branch_p that belong to leaf branches carry no probabilities: NaN pattern, on the solve's own stream
  if (out->branch_p) BMPC_CK(h, cudaMemsetAsync(out->branch_p, 0xff, (size_t)count * h->P.nbranch * h->cfg.m * sizeof(real), s));
  h->last_stream = s;
  const int grid = (int)(count < h->grid ? count : h->grid);
  // the parameter block travels through constant memory: one symbol per device, so launches from other streams or handles
  // are ordered behind the previous solve kernel of this device before the symbol is rewritten
  ConstSlot& cs = g_const[h->device & 15];
  std::lock_guard<std::mutex> lock(cs.mu);   // distinct handles may be driven from distinct host threads
  cudaStreamCaptureStatus capturing = cudaStreamCaptureStatusNone;
  if (s) BMPC_CK(h, cudaStreamIsCapturing(s, &capturing));
  if (capturing == cudaStreamCaptureStatusActive) {
    // Recorded into a CUDA graph: the upload node reads its pinned source at every replay, so the block gets a buffer of its
    // own that is never reused; nothing here may wait on the host, and the timing / ordering events stay out of the graph.
    // The captured graph must not run concurrently with other solves of this device (they share the constant symbol).
    if (h->captured_used >= BMPC_MAX_CAPTURED) { h->err = "too many captured launches for this handle"; return BMPC_E_CAPACITY; }
    KParams* own = h->captured + h->captured_used++;
    *own = P;
    BMPC_CK(h, cudaMemcpyToSymbolAsync(bmpc_cP, own, sizeof(KParams), 0, cudaMemcpyHostToDevice, s));
    const int rc = BMPC_DISPATCH(h, launch_instance, h, P, grid, s);
    if (rc != BMPC_OK) return rc;
    h->launches += 1;
    return BMPC_OK;
  }
  if (!cs.done) BMPC_CK(h, cudaEventCreateWithFlags(&cs.done, cudaEventDisableTiming));
  if (cs.used && cs.stream != s) BMPC_CK(h, cudaStreamWaitEvent(s, cs.done, 0));
  KParams* stage = h->pstage + (h->pstage_next++ % BMPC_PSTAGE);
  if (h->pstage_evt[stage - h->pstage]) BMPC_CK(h, cudaEventSynchronize(h->pstage_evt[stage - h->pstage]));
  else BMPC_CK(h, cudaEventCreateWithFlags(&h->pstage_evt[stage - h->pstage], cudaEventDisableTiming));
  *stage = P;
  BMPC_CK(h, cudaMemcpyToSymbolAsync(bmpc_cP, stage, sizeof(KParams), 0, cudaMemcpyHostToDevice, s));
  BMPC_CK(h, cudaEventRecord(h->pstage_evt[stage - h->pstage], s));
  BMPC_CK(h, cudaEventRecord(h->ev0, s));
  const int rc = BMPC_DISPATCH(h, launch_instance, h, P, grid, s);
  if (rc != BMPC_OK) return rc;
  BMPC_CK(h, cudaEventRecord(h->ev1, s));
  BMPC_CK(h, cudaEventRecord(cs.done, s));
  cs.used = true;
  cs.stream = s;
  h->timed = true;
  h->launches += 1;
  return BMPC_OK;
}

// Host-buffer step.  Inputs are gathered into one pinned block and travel in one DMA; the requested outputs are packed
// back to back on the device ([count] rows each), come back in one DMA into a pinned mirror and are handed out either as
// views into that mirror (bmpc_solve_host_views: zero copy, valid until the handle's next host call) or copied into
// the caller's arrays (bmpc_solve_host).
static const int kNumOut = 14;
static void out_sizes(const bmpc_handle* h, size_t* sz) {
  const KParams& P = h->P;
  const size_t n = h->cfg.n, d = h->cfg.d, m = h->cfg.m;
  const size_t nbel = (h->cfg.controller == BMPC_CTRL_BELIEF) ? (size_t)h->cfg.hmm_M * m : 0;
  const size_t v[kNumOut] = {d * 8, (size_t)P.pub_totalu * d * 8, (size_t)P.pub_totalx * n * 8, (size_t)P.pub_totalu * n * 8,
                             (size_t)P.pub_totalu * n * 8, (size_t)P.nbranch * 8, (size_t)P.nbranch * m * 8, 8, 4, 4, 4, 4, 8,
                             (size_t)P.pub_totalx * nbel * 8};
  for (int i = 0; i < kNumOut; ++i) sz[i] = v[i];
}
static void** out_slots(bmpc_outputs* o, void*** slots) {
  void** v[kNumOut] = {(void**)&o->u0, (void**)&o->uPred, (void**)&o->xPred, (void**)&o->xLin, (void**)&o->zPred,
                       (void**)&o->branch_w, (void**)&o->branch_p, (void**)&o->objective, (void**)&o->status,
                       (void**)&o->iters, (void**)&o->nfact, (void**)&o->nsolve, (void**)&o->cycles, (void**)&o->bPred};
  for (int i = 0; i < kNumOut; ++i) slots[i] = v[i];
  return nullptr;
}

// want: which outputs to produce (non-NULL members); views: receives host pointers into the pinned mirror
static int solve_host_impl(bmpc_handle* h, const double* x0, const double* z0, const double* xref,
                           const double* policy_params, int64_t count, const bmpc_outputs* want, bmpc_outputs* views,
                           const XformArgs* xf = nullptr) {
  if (!h) return BMPC_E_INVALID;
  if (!x0 || !z0 || !xref || !want || !views || count < 0) { h->err = "null argument"; return BMPC_E_INVALID; }
  if (count > h->cfg.batch_capacity) { h->err = "count exceeds batch_capacity"; return BMPC_E_CAPACITY; }
  bmpc_outputs w = *want;
  void** wslot[kNumOut];
  void** vslot[kNumOut];
  out_slots(&w, wslot);
  out_slots(views, vslot);
  for (int i = 0; i < kNumOut; ++i) *vslot[i] = nullptr;
  if (count == 0) return BMPC_OK;
  BMPC_CK(h, cudaSetDevice(h->device));
  const size_t cap = (size_t)h->cfg.batch_capacity, n = h->cfg.n, m = h->cfg.m;
  size_t sz[kNumOut], per = 0;
  out_sizes(h, sz);
  for (int i = 0; i < kNumOut; ++i) per += sz[i];
  const size_t nrw = (size_t)h->cfg.n_rows;
  const size_t in_reals = cap * (3 * n + 4 * m + n * n + 2 * nrw);
  if (!h->stage_in) {
    BMPC_CK(h, cudaMalloc(&h->stage_in, in_reals * sizeof(real)));
    BMPC_CK(h, cudaMallocHost(&h->stage_in_host, in_reals * sizeof(real)));
  }
  if (!h->stage_out) {
    h->stage_out_bytes = cap * per;
    BMPC_CK(h, cudaMalloc(&h->stage_out, h->stage_out_bytes));
    BMPC_CK(h, cudaMemset(h->stage_out, 0, h->stage_out_bytes));
    BMPC_CK(h, cudaMallocHost(&h->stage_out_host[0], h->stage_out_bytes));
    BMPC_CK(h, cudaMallocHost(&h->stage_out_host[1], h->stage_out_bytes));
    memset(h->stage_out_host[0], 0, h->stage_out_bytes);
    memset(h->stage_out_host[1], 0, h->stage_out_bytes);
  }
  cudaStream_t s = h->last_stream;
  BMPC_CK(h, cudaStreamSynchronize(s));
  char* host_out = (char*)h->stage_out_host[h->stage_out_turn++ & 1];   // the views handed out two calls ago die here
  // inputs: x0 | z0 | xref | polpar, [count] rows each, contiguous on the device.  Arrays the caller keeps in page-locked
  // memory travel straight from there (one DMA each); pageable ones are gathered into the pinned block first and travel in one
  // DMA.  Either way the stream is drained before this call returns, so the caller's arrays are free again.
  real* hin = h->stage_in_host;
  const size_t rows = (size_t)count;
  const void* src[6] = {x0, z0, xref, policy_params, xf ? xf->S : nullptr, xf ? xf->bounds : nullptr};
  const size_t len[6] = {rows * n, rows * n, rows * n, rows * m * 4, rows * n * n, rows * nrw * 2};
  bool direct = rows * n * 8 >= 4096;   // small batches: one gathered copy is cheaper than several DMAs
  for (int i = 0; i < 6 && direct; ++i) {
    if (!src[i]) continue;
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, src[i]) != cudaSuccess) { cudaGetLastError(); direct = false; }
    else if (attr.type != cudaMemoryTypeHost) direct = false;
  }
  size_t in_used = 0;
  real* dS = nullptr;
  real* dbd = nullptr;
  for (int i = 0; i < 6; ++i) {
    if (!src[i]) continue;
    if (direct) BMPC_CK(h, cudaMemcpyAsync(h->stage_in + in_used, src[i], len[i] * sizeof(real), cudaMemcpyHostToDevice, s));
    else memcpy(hin + in_used, src[i], len[i] * sizeof(real));
    if (i == 4) dS = h->stage_in + in_used;     // merge scenario: state transform and state bounds of the call, each optional
    if (i == 5) dbd = h->stage_in + in_used;
    in_used += len[i];
  }
  if (!direct) BMPC_CK(h, cudaMemcpyAsync(h->stage_in, hin, in_used * sizeof(real), cudaMemcpyHostToDevice, s));
  real* dx0 = h->stage_in;
  real* dz0 = dx0 + rows * n;
  real* dxr = dz0 + rows * n;
  real* dpp = dxr + rows * n;
  // outputs: only the requested ones, packed; 8-byte outputs first keeps every block aligned.  The kernel writes them straight
  // into the pinned host block (page-locked memory is device-addressable under unified addressing): every team's rows cross the
  // bus while the other teams are still solving, so the 190 MB of a full-interface step of 16 384 episodes cost no transfer time
  // after the launch.  reserved[6] bit 2: through the device block and one DMA instead (the former path; tests compare the two).
  static const bool via_device_env = getenv("BMPC_HOST_VIA_DEVICE") != nullptr;   // experiments (A/B of the two paths)
  const bool direct_out = (h->cfg.reserved[6] & 4) == 0 && !via_device_env;
  bmpc_outputs dout;
  void** dslot[kNumOut];
  out_slots(&dout, dslot);
  size_t off = 0;
  size_t offs[kNumOut];
  for (int pass = 0; pass < 2; ++pass)
    for (int i = 0; i < kNumOut; ++i) {
      const bool wide = (sz[i] % 8) == 0;
      if (wide != (pass == 0)) continue;
      if (*wslot[i]) {
        *dslot[i] = (direct_out ? host_out : (char*)h->stage_out) + off;
        offs[i] = off;
        off += rows * sz[i];
      } else {
        *dslot[i] = nullptr;
      }
    }
  const int rc = xf ? bmpc_solve_transformed(h, dx0, dz0, dxr, policy_params ? dpp : nullptr, dS, dbd, count, &dout, s)
                    : bmpc_solve(h, dx0, dz0, dxr, policy_params ? dpp : nullptr, count, &dout, s);
  if (rc != BMPC_OK) return rc;
  if (!direct_out) BMPC_CK(h, cudaMemcpyAsync(host_out, h->stage_out, off, cudaMemcpyDeviceToHost, s));
  BMPC_CK(h, cudaStreamSynchronize(s));
  for (int i = 0; i < kNumOut; ++i)
    if (*wslot[i]) *vslot[i] = host_out + offs[i];
  return BMPC_OK;
}

int bmpc_solve_host_views(bmpc_handle* h, const double* x0, const double* z0, const double* xref,
                          const double* policy_params, int64_t count, const bmpc_outputs* want, bmpc_outputs* views) {
  return solve_host_impl(h, x0, z0, xref, policy_params, count, want, views);
}

int bmpc_solve_transformed_host_views(bmpc_handle* h, const double* x0, const double* z0, const double* xref,
                                      const double* policy_params, const double* S, const double* state_bounds, int64_t count,
                                      const bmpc_outputs* want, bmpc_outputs* views) {
  if (!h) return BMPC_E_INVALID;
  if (h->cfg.model != BMPC_MODEL_MERGE) { h->err = "state transforms are built for BMPC_MODEL_MERGE handles"; return BMPC_E_INVALID; }
  const XformArgs xf{S, state_bounds};
  return solve_host_impl(h, x0, z0, xref, policy_params, count, want, views, &xf);
}

int bmpc_solve_host(bmpc_handle* h, const double* x0, const double* z0, const double* xref,
                    const double* policy_params, int64_t count, const bmpc_outputs* out) {
  if (!out) { if (h) h->err = "null argument"; return BMPC_E_INVALID; }
  bmpc_outputs views;
  const int rc = solve_host_impl(h, x0, z0, xref, policy_params, count, out, &views);
  if (rc != BMPC_OK || count == 0) return rc;
  bmpc_outputs o = *out;
  void** oslot[kNumOut];
  void** vslot[kNumOut];
  out_slots(&o, oslot);
  out_slots(&views, vslot);
  size_t sz[kNumOut];
  out_sizes(h, sz);
  for (int i = 0; i < kNumOut; ++i)
    if (*oslot[i]) memcpy(*oslot[i], *vslot[i], (size_t)count * sz[i]);
  return BMPC_OK;
}

int bmpc_get_state(bmpc_handle* h, double* uLin, int32_t* pbest, double* old_input, int32_t* started, double* xprev,
                   int64_t count, int on_host) {
  if (!h) return BMPC_E_INVALID;
  if (count < 0 || count > h->cfg.batch_capacity) { h->err = "count out of range"; return BMPC_E_CAPACITY; }
  if (xprev && !h->xprev) { h->err = "xprev is robustMPC state"; return BMPC_E_INVALID; }
  BMPC_CK(h, cudaSetDevice(h->device));
  const cudaMemcpyKind k = on_host ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
  const KParams& P = h->P;
  cudaStream_t s = h->last_stream;
  if (uLin) BMPC_CK(h, cudaMemcpyAsync(uLin, h->uLin, count * (P.totalu + 1) * h->cfg.d * sizeof(real), k, s));
  if (pbest) BMPC_CK(h, cudaMemcpyAsync(pbest, h->pbest, count * P.nbranch * sizeof(int), k, s));
  if (old_input) BMPC_CK(h, cudaMemcpyAsync(old_input, h->oldin, count * h->cfg.d * sizeof(real), k, s));
  if (started) BMPC_CK(h, cudaMemcpyAsync(started, h->started, count * sizeof(int), k, s));
  if (xprev) BMPC_CK(h, cudaMemcpyAsync(xprev, h->xprev, count * P.pub_totalx * h->cfg.n * sizeof(real), k, s));
  BMPC_CK(h, cudaStreamSynchronize(s));
  return BMPC_OK;
}

int bmpc_set_state(bmpc_handle* h, const double* uLin, const int32_t* pbest, const double* old_input,
                   const int32_t* started, const double* xprev, int64_t count, int on_host) {
  if (!h) return BMPC_E_INVALID;
  if (count < 0 || count > h->cfg.batch_capacity) { h->err = "count out of range"; return BMPC_E_CAPACITY; }
  if (xprev && !h->xprev) { h->err = "xprev is robustMPC state"; return BMPC_E_INVALID; }
  if (h->xprev && started && !xprev) {
    h->err = "robustMPC linearises about the previous predicted states: restoring `started` needs xprev as well";
    return BMPC_E_INVALID;
  }
  BMPC_CK(h, cudaSetDevice(h->device));
  const cudaMemcpyKind k = on_host ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice;
  const KParams& P = h->P;
  cudaStream_t s = h->last_stream;
  if (uLin) BMPC_CK(h, cudaMemcpyAsync(h->uLin, uLin, count * (P.totalu + 1) * h->cfg.d * sizeof(real), k, s));
  if (pbest) BMPC_CK(h, cudaMemcpyAsync(h->pbest, pbest, count * P.nbranch * sizeof(int), k, s));
  if (old_input) BMPC_CK(h, cudaMemcpyAsync(h->oldin, old_input, count * h->cfg.d * sizeof(real), k, s));
  if (started) BMPC_CK(h, cudaMemcpyAsync(h->started, started, count * sizeof(int), k, s));
  if (xprev) BMPC_CK(h, cudaMemcpyAsync(h->xprev, xprev, count * P.pub_totalx * h->cfg.n * sizeof(real), k, s));
  // a caller-supplied warm start invalidates the solver's own caches for those episodes
  BMPC_CK(h, cudaMemsetAsync(h->cache_state, 0xff, count * 2 * sizeof(int), s));
  BMPC_CK(h, cudaStreamSynchronize(s));
  return BMPC_OK;
}

int bmpc_eval_model(bmpc_handle* h, const double* x, const double* z, const double* u, const double* policy_params,
                    int64_t count, double* A, double* B, double* C, double* xp, double* zpred, double* p, double* hlin,
                    double* dh, void* stream) {
  if (!h) return BMPC_E_INVALID;
  if (!x || count < 0) { h->err = "null argument"; return BMPC_E_INVALID; }
  if (needs_lookup(h)) { h->err = "the policy table uses a lookup-table policy: call bmpc_set_lookup_table first"; return BMPC_E_INVALID; }
  if (count == 0) return BMPC_OK;
  BMPC_CK(h, cudaSetDevice(h->device));
  EvalArgs a{x, z, u, policy_params, A, B, C, xp, zpred, p, hlin, dh, (int)count};
  const int threads = 128, blocks = (int)((count + threads - 1) / threads);
  cudaStream_t s = (cudaStream_t)stream;
  if (h->cfg.model == BMPC_MODEL_HIGHWAY) bmpc_eval_kernel<HighwayModel><<<blocks, threads, 0, s>>>(h->P, a);
  else if (h->cfg.model == BMPC_MODEL_MERGE) bmpc_eval_kernel<MergeModel><<<blocks, threads, 0, s>>>(h->P, a);
  else bmpc_eval_kernel<QuadrupedModel><<<blocks, threads, 0, s>>>(h->P, a);
  BMPC_CK(h, cudaGetLastError());
  h->launches += 1;
  return BMPC_OK;
}

int bmpc_eval_belief(bmpc_handle* h, const double* xb, const double* xbackup, const double* u, int64_t count, double* A,
                     double* B, double* C, double* h0, double* Jh, double* xbp, void* stream) {
  if (!h) return BMPC_E_INVALID;
  if (h->cfg.controller != BMPC_CTRL_BELIEF) { h->err = "not a belief-state MPC handle"; return BMPC_E_INVALID; }
  if (!xb || !xbackup || !u || !A || !B || !C || !h0 || !Jh || !xbp || count < 0) { h->err = "null argument"; return BMPC_E_INVALID; }
  if (count == 0) return BMPC_OK;
  BMPC_CK(h, cudaSetDevice(h->device));
  BeliefEvalArgs a{xb, xbackup, u, A, B, C, h0, Jh, xbp, (int)count};
  bmpc_belief_eval_kernel<<<(int)((count + 63) / 64), 64, 0, (cudaStream_t)stream>>>(h->P, a);
  BMPC_CK(h, cudaGetLastError());
  h->launches += 1;
  return BMPC_OK;
}

int bmpc_plant_step(bmpc_handle* h, double* x, const double* u, double* z, int32_t obstacle_policy,
                    const double* policy_params, int64_t count, void* stream) {
  if (!h) return BMPC_E_INVALID;
  if (count < 0 || obstacle_policy < 0 || obstacle_policy >= h->cfg.m) { h->err = "bad argument"; return BMPC_E_INVALID; }
  if (count == 0) return BMPC_OK;
  BMPC_CK(h, cudaSetDevice(h->device));
  const int threads = 128, blocks = (int)((count + threads - 1) / threads);
  cudaStream_t s = (cudaStream_t)stream;
  if (h->cfg.model == BMPC_MODEL_HIGHWAY || h->cfg.model == BMPC_MODEL_MERGE)
    bmpc_plant_kernel<HighwayModel><<<blocks, threads, 0, s>>>(h->P, x, u, z, obstacle_policy, policy_params, (int)count);
  else
    bmpc_plant_kernel<QuadrupedModel><<<blocks, threads, 0, s>>>(h->P, x, u, z, obstacle_policy, policy_params, (int)count);
  BMPC_CK(h, cudaGetLastError());
  h->launches += 1;
  return BMPC_OK;
}

int bmpc_env_step(bmpc_handle* h, const bmpc_env_state* env, int64_t count, int32_t t, int32_t n_lane,
                  const double* quad_sizes, const bmpc_outputs* out, void* stream) {
  if (!h) return BMPC_E_INVALID;
  if (!env || !out || count < 0 || t < 0) { h->err = "null argument"; return BMPC_E_INVALID; }
  if (!env->x || !env->z || !env->obs_policy || !env->collided || !env->xref || !env->u_obs || !out->u0) {
    h->err = "bmpc_env_step needs x, z, obs_policy, collided, xref, u_obs and out->u0";
    return BMPC_E_INVALID;
  }
  if (h->cfg.model == BMPC_MODEL_MERGE) { h->err = "the merge environment (Highway_env_merge) is stepped by the host caller"; return BMPC_E_UNSUPPORTED; }
  const bool highway = h->cfg.model == BMPC_MODEL_HIGHWAY;
  if (highway && (!env->lane || !env->policy_params || n_lane < 1)) {
    h->err = "highway environment needs lane, policy_params and n_lane";
    return BMPC_E_INVALID;
  }
  if (!highway && (!env->goal || !quad_sizes)) { h->err = "quadruped environment needs goal and quad_sizes"; return BMPC_E_INVALID; }
  if (bmpc_is_chain(h->cfg.controller)) { h->err = "the environment drives the branch controllers"; return BMPC_E_UNSUPPORTED; }
  if (count == 0) return BMPC_OK;
  if (count > h->cfg.batch_capacity) { h->err = "count exceeds batch_capacity"; return BMPC_E_CAPACITY; }
  BMPC_CK(h, cudaSetDevice(h->device));
  cudaStream_t s = (cudaStream_t)stream;
  EnvArgs a;
  a.x = env->x; a.z = env->z; a.lane = env->lane; a.polpar = env->policy_params; a.goal = env->goal;
  a.obs_policy = env->obs_policy; a.collided = env->collided; a.xref = env->xref; a.u_obs = env->u_obs;
  a.count = (int)count; a.t = t; a.n_lane = n_lane;
  const int threads = 128, blocks = (int)((count + threads - 1) / threads);
  if (highway) bmpc_env_pre_highway<<<blocks, threads, 0, s>>>(h->P, a);
  else bmpc_env_pre_quadruped<<<blocks, threads, 0, s>>>(h->P, a, quad_sizes[0], quad_sizes[1], quad_sizes[2]);
  BMPC_CK(h, cudaGetLastError());
  h->launches += 1;
  const int rc = bmpc_solve(h, env->x, env->z, env->xref, env->policy_params, count, out, stream);
  if (rc != BMPC_OK) return rc;
  if (highway) bmpc_env_post<HighwayModel><<<blocks, threads, 0, s>>>(h->P, env->x, env->z, out->u0, env->u_obs, (int)count);
  else bmpc_env_post<QuadrupedModel><<<blocks, threads, 0, s>>>(h->P, env->x, env->z, out->u0, env->u_obs, (int)count);
  BMPC_CK(h, cudaGetLastError());
  h->launches += 1;
  return BMPC_OK;
}

int bmpc_env_step_merge(bmpc_handle* h, const bmpc_merge_env_state* env, int64_t count, int32_t n_lane, int32_t merge_lane,
                        double merge_s, double v0, const bmpc_outputs* out, void* stream) {
  if (!h) return BMPC_E_INVALID;
  if (h->cfg.model != BMPC_MODEL_MERGE) { h->err = "bmpc_env_step_merge drives BMPC_MODEL_MERGE handles"; return BMPC_E_INVALID; }
  if (!env || !out || count < 0 || n_lane < 1 || merge_lane < 1) { h->err = "bad argument"; return BMPC_E_INVALID; }
  if (!env->x || !env->z || !env->lane_id || !env->collided || !env->xref || !env->S || !env->state_bounds || !env->u_obs ||
      !env->table_x || !env->table_y || !env->table_psi || env->table_n < 2 || !out->u0) {
    h->err = "bmpc_env_step_merge needs every state array, the three tables and out->u0";
    return BMPC_E_INVALID;
  }
  if (count == 0) return BMPC_OK;
  if (count > h->cfg.batch_capacity) { h->err = "count exceeds batch_capacity"; return BMPC_E_CAPACITY; }
  BMPC_CK(h, cudaSetDevice(h->device));
  cudaStream_t s = (cudaStream_t)stream;
  MergeEnvArgs a;
  a.x = env->x; a.z = env->z; a.lane_id = env->lane_id; a.collided = env->collided; a.xref = env->xref; a.S = env->S;
  a.bounds = env->state_bounds; a.u_obs = env->u_obs; a.tab_x = env->table_x; a.tab_y = env->table_y; a.tab_psi = env->table_psi;
  a.tab_n = env->table_n; a.count = (int)count; a.n_lane = n_lane; a.merge_lane = merge_lane; a.merge_s = merge_s; a.v0 = v0;
  a.psimax = h->P.rhi[1];   // mpc.psimax = bx[0][2][0] (MPC_branch.py:1621): the upper bound of the heading row
  const int threads = 128, blocks = (int)((count + threads - 1) / threads);
  bmpc_env_pre_merge<<<blocks, threads, 0, s>>>(h->P, a);
  BMPC_CK(h, cudaGetLastError());
  h->launches += 1;
  const int rc = bmpc_solve_transformed(h, env->x, env->z, env->xref, nullptr, env->S, env->state_bounds, count, out, stream);
  if (rc != BMPC_OK) return rc;
  bmpc_env_post<HighwayModel><<<blocks, threads, 0, s>>>(h->P, env->x, env->z, out->u0, env->u_obs, (int)count);
  BMPC_CK(h, cudaGetLastError());
  h->launches += 1;
  return BMPC_OK;
}

int bmpc_staging_enabled(const bmpc_handle* h) { return h ? h->P.stage_on : BMPC_E_INVALID; }

int bmpc_get_launch_info(const bmpc_handle* h, int32_t* slab_mode, int32_t* warps, int64_t* smem_bytes,
                         int64_t* global_bytes_per_warp) {
  if (!h) return BMPC_E_INVALID;
  if (slab_mode) *slab_mode = h->mode;
  if (warps) *warps = h->grid;
  if (smem_bytes) *smem_bytes = (int64_t)h->slab_bytes;
  if (global_bytes_per_warp) *global_bytes_per_warp = (int64_t)h->gws_bytes_per_warp;
  return BMPC_OK;
}

static int hmm_kinds(const int32_t* kinds, int m, hmm::Kinds* out) {
  if (!kinds || m < 1 || m > BMPC_MAX_POLICIES) return BMPC_E_INVALID;
  for (int j = 0; j < BMPC_MAX_POLICIES; ++j) out->v[j] = 0;
  for (int j = 0; j < m; ++j) {
    if (kinds[j] != BMPC_HMM_MAINTAIN && kinds[j] != BMPC_HMM_BRAKE) return BMPC_E_INVALID;
    out->v[j] = kinds[j];
  }
  return BMPC_OK;
}

// The three belief-state entry points only enqueue on the caller's stream: no allocation, no synchronisation.
int bmpc_hmm_backup_rollout(const double* x0, int64_t count, int32_t M, int32_t m, const int32_t* policy_kind, int32_t N,
                            double dt, double Kpsi, double* xbackup, int32_t device, void* stream) {
  if (!x0 || !xbackup || count < 0 || M < 1 || N < 1) { g_create_error = "bad argument"; return BMPC_E_INVALID; }
  if (count == 0) return BMPC_OK;
  if (cudaSetDevice(device) != cudaSuccess) return BMPC_E_CUDA;
  hmm::Kinds dk;
  const int rc = hmm_kinds(policy_kind, m, &dk);
  if (rc != BMPC_OK) return rc;
  const int64_t n = count * M * m;
  hmm::rollout_kernel<<<(int)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(x0, (int)count, M, m, dk, N, dt, Kpsi, xbackup);
  return cudaGetLastError() == cudaSuccess ? BMPC_OK : BMPC_E_CUDA;
}

int bmpc_hmm_rollout_sensitivity(const double* x0, int64_t count, int32_t m, const int32_t* policy_kind, int32_t steps,
                                 double ts, double Kpsi, const double* f0, double* xx, double* QQ, double* Qt,
                                 int32_t device, void* stream) {
  if (!x0 || !f0 || !xx || !QQ || !Qt || count < 0 || steps < 1) { g_create_error = "bad argument"; return BMPC_E_INVALID; }
  if (count == 0) return BMPC_OK;
  if (cudaSetDevice(device) != cudaSuccess) return BMPC_E_CUDA;
  hmm::Kinds dk;
  const int rc = hmm_kinds(policy_kind, m, &dk);
  if (rc != BMPC_OK) return rc;
  const int64_t n = count * m;
  hmm::sensitivity_kernel<<<(int)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(x0, (int)count, m, dk, steps, ts, Kpsi, f0,
                                                                                     xx, QQ, Qt);
  return cudaGetLastError() == cudaSuccess ? BMPC_OK : BMPC_E_CUDA;
}

int bmpc_hmm_belief_update(const double* ego, const double* xb, const double* b, const double* cbf, int64_t count,
                           int32_t M, int32_t m, const bmpc_hmm_params* p, int32_t clip, double* h, double* H,
                           double* b_next, int32_t device, void* stream) {
  if (!ego || !xb || !b || !p || !b_next || count < 0 || M < 1 || m < 1 || m > BMPC_MAX_POLICIES) {
    g_create_error = "bad argument";
    return BMPC_E_INVALID;
  }
  if (count == 0) return BMPC_OK;
  if (cudaSetDevice(device) != cudaSuccess) return BMPC_E_CUDA;
  const int64_t n = count * M;
  hmm::belief_kernel<<<(int)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(ego, xb, b, cbf, (int)count, M, m, *p, clip, h, H,
                                                                                b_next);
  return cudaGetLastError() == cudaSuccess ? BMPC_OK : BMPC_E_CUDA;
}

int64_t bmpc_launch_count(const bmpc_handle* h) { return h ? h->launches : 0; }

double bmpc_measure_fp64_peak(int device, int iters) {
  if (cudaSetDevice(device) != cudaSuccess) return -1.0;
  int sms = 0;
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) return -1.0;
  if (iters < 1) iters = 4096;
  const int threads = 256, blocks = sms * 8;
  double* out = nullptr;
  if (cudaMalloc(&out, (size_t)threads * blocks * sizeof(double)) != cudaSuccess) return -1.0;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  bmpc_dfma_kernel<<<blocks, threads>>>(out, iters);   // warm-up
  double best = -1.0;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(e0);
    bmpc_dfma_kernel<<<blocks, threads>>>(out, iters);
    cudaEventRecord(e1);
    if (cudaEventSynchronize(e1) != cudaSuccess) { best = -1.0; break; }
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    const double tf = 2.0 * 8.0 * (double)iters * threads * blocks / (ms * 1e-3) * 1e-12;
    if (tf > best) best = tf;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(out);
  return best;
}

float bmpc_last_kernel_ms(bmpc_handle* h) {
  if (!h || !h->timed) return -1.f;
  float ms = -1.f;
  if (cudaEventSynchronize(h->ev1) != cudaSuccess) return -1.f;
  if (cudaEventElapsedTime(&ms, h->ev0, h->ev1) != cudaSuccess) return -1.f;
  return ms;
}

const char* bmpc_last_error(const bmpc_handle* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

}  // extern "C"
