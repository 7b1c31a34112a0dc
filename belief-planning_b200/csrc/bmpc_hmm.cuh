// Belief-state model kernels (rows H1, H2 of SURVEY.md 8a): backup-policy rollouts of the other agents (optionally
// with the sensitivity matrix) and the belief transition / update.  Numeric functions of HMM_backup_dyn.py:
// backup_maintain :105, backup_brake :107-109, propagate_backup :122-132, generate_backup_traj :54-85 and :204-214,
// veh_col :136-157, lane_bdry_h :134, backup_trans :96-101, backup_input_prob :103; Highway_env.py:251-256.
#pragma once
#include <cuda_runtime.h>

#include "branchmpc.h"

namespace hmm {

__device__ __forceinline__ double softsat(double x, double s) { const double e = exp(s * x); return (e - 1.0) / (e + 1.0) * 0.5 + 0.5; }
__device__ __forceinline__ double softmin2(double x, double y, double g) {
  const double mn = fmin(x, y);
  const double ex = exp(-g * (x - mn)), ey = exp(-g * (y - mn));
  return (ex * x + ey * y) / (ex + ey);
}
__device__ __forceinline__ double softmax2(double x, double y, double g) {
  const double mx = fmax(x, y);
  const double ex = exp(g * (x - mx)), ey = exp(g * (y - mx));
  return (ex * x + ey * y) / (ex + ey);
}
__device__ __forceinline__ void policy(int kind, const double* x, double Kpsi, double* u) {
  u[0] = (kind == BMPC_HMM_BRAKE) ? softmax2(-5.0, -x[2], 3.0) : 0.0;
  u[1] = -Kpsi * x[3];
}
__device__ __forceinline__ void xdot(const double* x, const double* u, double* f) {
  double s, c;
  sincos(x[3], &s, &c);
  f[0] = x[2] * c;
  f[1] = x[2] * s;
  f[2] = u[0];
  f[3] = u[1];
}

// one thread per (episode, agent, policy): row (m*i + j) of xbackup, flattened component-major (casadi.reshape)
struct Kinds { int v[BMPC_MAX_POLICIES]; };   // policy table by value: no device allocation per call
__global__ void rollout_kernel(const double* __restrict__ x0, int count, int M, int m, const Kinds kinds, int N,
                               double dt, double Kpsi, double* __restrict__ xbackup) {
  const int id = blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= count * M * m) return;
  const int j = id % m, i = (id / m) % M, e = id / (m * M);
  double x[4], u[2], f[4];
  for (int q = 0; q < 4; ++q) x[q] = x0[((size_t)e * M + i) * 4 + q];
  double* row = xbackup + ((size_t)e * M * m + (size_t)m * i + j) * N * 4;
  const int kind = kinds.v[j];
  for (int t = 0; t < N; ++t) {
    policy(kind, x, Kpsi, u);
    xdot(x, u, f);
    for (int q = 0; q < 4; ++q) {
      x[q] += f[q] * dt;
      row[(size_t)q * N + t] = x[q];
    }
  }
}

// one thread per (point, policy): states, sensitivity matrices dx_t/dx_0 and xdot - f0 BEFORE each of `steps` steps
__global__ void sensitivity_kernel(const double* __restrict__ x0, int count, int m, const Kinds kinds, int steps,
                                   double ts, double Kpsi, const double* __restrict__ f0, double* __restrict__ xx,
                                   double* __restrict__ QQ, double* __restrict__ Qt) {
  const int id = blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= count * m) return;
  const int j = id % m, e = id / m;
  const int kind = kinds.v[j];
  double x[4], Q[16];
  for (int q = 0; q < 4; ++q) x[q] = x0[(size_t)e * 4 + q];
  for (int q = 0; q < 16; ++q) Q[q] = (q % 5 == 0) ? 1.0 : 0.0;
  const double h = 1e-6;
  for (int t = 0; t < steps; ++t) {
    const size_t o = (size_t)id * steps + t;
    double u[2], f[4], ja[16];
    policy(kind, x, Kpsi, u);
    xdot(x, u, f);
    for (int q = 0; q < 16; ++q) QQ[o * 16 + q] = Q[q];
    for (int q = 0; q < 4; ++q) { xx[o * 4 + q] = x[q]; Qt[o * 4 + q] = f[q] - f0[q]; }
    // Jacobian of the closed loop: rows 0,1 analytic, rows 2,3 = policy Jacobian by central differences (dubin_f_x :43-52)
    double s, c;
    sincos(x[3], &s, &c);
    for (int q = 0; q < 16; ++q) ja[q] = 0.0;
    ja[2] = c; ja[3] = -x[2] * s; ja[6] = s; ja[7] = x[2] * c;
    for (int k = 0; k < 4; ++k) {
      double xp[4] = {x[0], x[1], x[2], x[3]}, xm[4] = {x[0], x[1], x[2], x[3]}, up[2], um[2];
      xp[k] += h;
      xm[k] -= h;
      policy(kind, xp, Kpsi, up);
      policy(kind, xm, Kpsi, um);
      ja[8 + k] = (up[0] - um[0]) / 2 / h;
      ja[12 + k] = (up[1] - um[1]) / 2 / h;
    }
    double Qn[16];
    for (int r = 0; r < 4; ++r)
      for (int cc = 0; cc < 4; ++cc) {
        double a = 0.0;
        for (int k = 0; k < 4; ++k) a += ja[r * 4 + k] * Q[k * 4 + cc];
        Qn[r * 4 + cc] = Q[r * 4 + cc] + a * ts;
      }
    for (int q = 0; q < 16; ++q) Q[q] = Qn[q];
    for (int q = 0; q < 4; ++q) x[q] += f[q] * ts;
  }
}

// one thread per (episode, agent): h_j, transition matrix H, b+ = b H, optional input-probability update + normalisation
__global__ void belief_kernel(const double* __restrict__ ego, const double* __restrict__ xb, const double* __restrict__ b,
                              const double* __restrict__ cbf, int count, int M, int m, bmpc_hmm_params p, int clip,
                              double* __restrict__ h_out, double* __restrict__ H_out, double* __restrict__ b_next) {
  const int id = blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= count * M) return;
  const int e = id / M;
  const double* x = ego + (size_t)e * 4;
  double hv[BMPC_MAX_POLICIES], mh[BMPC_MAX_POLICIES], msum = 0.0;
  for (int j = 0; j < m; ++j) {
    const double* z = xb + ((size_t)id * m + j) * 4;
    double dx = (fabs(x[0] - z[0]) - (p.L + 1.0)) / (p.L + 1.0);
    double dy = (fabs(x[1] - z[1]) - (p.W + 0.2)) / (p.W + 0.2);
    if (clip) { dx = fmin(fmax(dx, -5.0), 5.0); dy = fmin(fmax(dy, -5.0), 5.0); }
    const double mx = fmax(dx, dy);
    const double ex = exp(dx - mx), ey = exp(dy - mx);
    const double col = (dx * ex + dy * ey) / (ex + ey);
    const double lane = softmin2(z[1] - p.ylb, p.yub - z[1], 5.0);
    hv[j] = softmin2(col, lane, p.col_alpha);
    mh[j] = softsat(hv[j], p.s1);
    msum += mh[j];
    if (h_out) h_out[(size_t)id * m + j] = hv[j];
  }
  double bn[BMPC_MAX_POLICIES];
  for (int j = 0; j < m; ++j) bn[j] = 0.0;
  for (int r = 0; r < m; ++r)
    for (int j = 0; j < m; ++j) {
      const double Hrj = (1.0 - p.tran_diag) * mh[j] / msum + ((r == j) ? p.tran_diag : 0.0);
      if (H_out) H_out[((size_t)id * m + r) * m + j] = Hrj;
      bn[j] += b[(size_t)id * m + r] * Hrj;
    }
  if (cbf) {
    double s = 0.0;
    for (int j = 0; j < m; ++j) {
      bn[j] *= softsat(cbf[(size_t)id * m + j] - p.c2, p.s2);
      s += bn[j];
    }
    for (int j = 0; j < m; ++j) bn[j] /= s;
  }
  for (int j = 0; j < m; ++j) b_next[(size_t)id * m + j] = bn[j];
}

}  // namespace hmm
