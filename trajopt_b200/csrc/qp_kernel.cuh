// QP subproblem kernel: one warp per trajectory.
//
// Replaces OSQPModel::optimize() -> osqp_setup/osqp_solve (trajopt_sco/src/osqp_interface.cpp:283-615) for
// every trajectory of the batch: the l1-penalty QP of optimizers.cpp:781-799 (Appendix B of SURVEY.md) is
// assembled from the fixed-layout convexification rows, equilibrated (Ruiz), and solved with the
// OSQP-equivalent ADMM (rho_eq = 1e3 rho, sigma, alpha relaxation, residual tests every 25 iterations,
// adaptive rho with refactorisation, polish).  Linear algebra: the KKT solve is done in its reduced form
//   (P + sigma I + A' diag(rho) A) x = rhs
// after eliminating, row by row and in closed form, the hinge / abs auxiliary variables (each couples to
// exactly one row): what is left is an N x N symmetric positive definite matrix, N = T*D, with half
// bandwidth 2*D (block tridiagonal in time).  It is factored once per rho in shared memory and every
// ADMM iteration is two banded triangular solves + one pass over the rows (projection, dual update).
#pragma once
#include "device_types.cuh"
#include "eval_kernel.cuh"

namespace tb200 {

constexpr double kOsqpInf = 1e30;
constexpr double kMinScaling = 1e-4, kMaxScaling = 1e4;
constexpr double kVerifyTol = 1e-9;  // KKT verification of the polished point (deviation D2)
constexpr int kVerifyRounds = 3;
constexpr double kRhoMin = 1e-6, kRhoMax = 1e6, kRhoTol = 1e-4, kRhoEqOverIneq = 1e3;
enum { QPS_UNSOLVED = 0, QPS_SOLVED = 1, QPS_SOLVED_INACC = 2, QPS_PINF = 3, QPS_PINF_INACC = 4, QPS_DINF = 5,
       QPS_DINF_INACC = 6, QPS_MAXITER = 7, QPS_NONCVX = 8 };

__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {  // fixed tree => deterministic
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ int warp_sum_int(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double limit_scaling(double v) {
  v = v < kMinScaling ? 1.0 : v;
  return v > kMaxScaling ? kMaxScaling : v;
}

struct QpSmem {
  int Kb, Dz, Eb, qs, lbs, ubs, x, zb, yb, v1, v2, invd, ints, total;
};
__host__ __device__ inline QpSmem qp_smem_layout(int N, int HB, int T, int D) {
  QpSmem s;
  int o = 0;
  s.Kb = o;   o += N * (HB + 1);
  s.Dz = o;   o += N;
  s.Eb = o;   o += N;
  s.qs = o;   o += N;
  s.lbs = o;  o += N;
  s.ubs = o;  o += N;
  s.x = o;    o += N;
  s.zb = o;   o += N;
  s.yb = o;   o += N;
  s.v1 = o;   o += N;
  s.v2 = o;   o += N;
  s.invd = o; o += N;
  s.ints = o; o += (T + D + 2 + 1) / 2 + 1;
  s.total = o;
  return s;
}

// Per-warp solver context.  All lanes hold identical copies of the scalar members.
struct QpCtx {
  // geometry
  int N, HB, T, D, CN, RS, lane;
  // shared memory
  double *Kb, *Dz, *Eb, *qs, *lbs, *ubs, *x, *zb, *yb, *v1, *v2, *invd;
  int* ls;  // list starts: T dense slots then D sparse slots, +1
  // global memory
  double* rows;
  int* rints;
  const int* list;
  const double* Pband;
  double* scratch;
  // scalars
  int nrows;
  double c, cinv, rho, rho_eq, sigma, alpha;

  __device__ __forceinline__ double* R(int r) const { return rows + static_cast<size_t>(r) * RS; }
  __device__ __forceinline__ const int* I(int r) const { return rints + static_cast<size_t>(r) * RI_NINTS; }

  template <class F>
  __device__ __forceinline__ void for_rows(F f) const {
    for (int slot = lane; slot < T; slot += 32)
      for (int k = ls[slot]; k < ls[slot + 1]; ++k) f(list[k]);
    __syncwarp();
    for (int slot = lane; slot < D; slot += 32)
      for (int k = ls[T + slot]; k < ls[T + slot + 1]; ++k) f(list[k]);
    __syncwarp();
  }
};

// scaled view of one row
struct RowV {
  int base, cnt, stride, naux;
  double E, u0, u1, b0, b1, lo, up, qa0, qa1, rho;
};
__device__ __forceinline__ void row_view(const QpCtx& q, const double* R, const int* I, RowV& v) {
  v.base = I[RI_BASE];
  v.cnt = I[RI_CNT];
  v.stride = I[RI_STRIDE];
  const int aux = I[RI_AUX];
  const double* F = R + q.CN;
  v.E = F[F_E];
  v.u0 = v.u1 = v.b0 = v.b1 = v.qa0 = v.qa1 = 0.0;
  v.naux = aux;  // AUX_NONE 0, HINGE 1, ABS 2 == number of aux variables
  v.up = -F[F_C] * v.E;
  if (aux == AUX_HINGE) {
    v.u0 = -v.E * F[F_DA0];
    v.b0 = F[F_EA0] * F[F_DA0];
    v.qa0 = q.c * F[F_DA0] * F[F_W];
    v.lo = -kOsqpInf * v.E;
    v.rho = q.rho;
  } else {
    if (aux == AUX_ABS) {
      v.u0 = v.E * F[F_DA0];
      v.u1 = -v.E * F[F_DA1];
      v.b0 = F[F_EA0] * F[F_DA0];
      v.b1 = F[F_EA1] * F[F_DA1];
      v.qa0 = q.c * F[F_DA0] * F[F_W];
      v.qa1 = q.c * F[F_DA1] * F[F_W];
    }
    v.lo = v.up;
    v.rho = q.rho_eq;
  }
}

// ---------------------------------------------------------------------------------------------------
// banded Cholesky (in place in Kb, lower band: Kb[i*(HB+1)+k] = K(i, i-k)) and the two triangular solves
__device__ inline bool band_factor(const QpCtx& q) {
  const int N = q.N, HB = q.HB, W = HB + 1, lane = q.lane;
  bool ok = true;
  for (int j = 0; j < N; ++j) {
    const double djj = q.Kb[j * W];
    if (!(djj > 0.0)) ok = false;
    const double d = sqrt(djj > 0.0 ? djj : 1.0);
    const double inv = 1.0 / d;
    __syncwarp();
    for (int k = 1 + lane; k <= HB && j + k < N; k += 32) q.Kb[(j + k) * W + k] *= inv;
    if (lane == 0) {
      q.Kb[j * W] = d;
      q.invd[j] = inv;
    }
    __syncwarp();
    // trailing update: K(j+a, j+b) -= L(j+a,j) L(j+b,j), 1 <= b <= a <= HB
    const int m = min(HB, N - 1 - j);
    const int npairs = m * (m + 1) / 2;
    for (int pidx = lane; pidx < npairs; pidx += 32) {
      // invert pidx -> (a,b): a*(a-1)/2 + (b-1) with 1<=b<=a
      int a = static_cast<int>((sqrt(8.0 * pidx + 1.0) - 1.0) * 0.5) + 1;
      while (a * (a - 1) / 2 > pidx) --a;
      while ((a + 1) * a / 2 <= pidx) ++a;
      const int b = pidx - a * (a - 1) / 2 + 1;
      q.Kb[(j + a) * W + (a - b)] -= q.Kb[(j + a) * W + a] * q.Kb[(j + b) * W + b];
    }
    __syncwarp();
  }
  return ok;
}
// solves K v = v in place (v in shared memory)
__device__ inline void band_solve(const QpCtx& q, double* v) {
  const int N = q.N, HB = q.HB, W = HB + 1, lane = q.lane;
  __syncwarp();
  for (int j = 0; j < N; ++j) {  // forward, column oriented
    const double yj = v[j] * q.invd[j];
    __syncwarp();
    if (lane == 0) v[j] = yj;
    for (int k = 1 + lane; k <= HB && j + k < N; k += 32) v[j + k] -= q.Kb[(j + k) * W + k] * yj;
    __syncwarp();
  }
  for (int j = N - 1; j >= 0; --j) {  // backward: x_j = (y_j - sum_k L(j+k,j) x_{j+k}) / L_jj
    double s = 0.0;
    for (int k = 1 + lane; k <= HB && j + k < N; k += 32) s += q.Kb[(j + k) * W + k] * v[j + k];
    s = warp_sum(s);
    __syncwarp();
    if (lane == 0) v[j] = (v[j] - s) * q.invd[j];
    __syncwarp();
  }
}

// scaled P (band) times a shared vector: out = c * Dz .* (P (Dz .* in))
__device__ inline void p_matvec(const QpCtx& q, const double* in, double* out) {
  const int N = q.N, HB = q.HB, W = HB + 1;
  for (int i = q.lane; i < N; i += 32) {
    double s = 0.0;
    for (int k = 0; k <= HB && k <= i; ++k) s += q.Pband[i * W + k] * q.Dz[i - k] * in[i - k];
    for (int k = 1; k <= HB && i + k < N; ++k) s += q.Pband[(i + k) * W + k] * q.Dz[i + k] * in[i + k];
    out[i] = q.c * q.Dz[i] * s;
  }
  __syncwarp();
}

// weights of the linear system: ADMM (rho vector, sigma) or polish (1/delta on the active set, delta)
struct SysW {
  bool polish;
  double sig, rho_aux;  // ADMM: rho on the aux bound rows
};
// Weights of one row in the current linear system.  With the aux block K_aa = diag(g) + Wr u u' the
// closed forms below are written cancellation free (den = det(K_aa) / 1, expanded analytically): the polish
// system has Wr = 1/delta and g = delta, where the textbook Sherman-Morrison form loses ~12 digits.
__device__ __forceinline__ void row_weights(const QpCtx& q, const SysW& w, const double* F, const RowV& v, double& Wr,
                                            double& g0, double& g1, double& den) {
  double wa0, wa1;
  if (w.polish) {
    Wr = fabs(F[F_PW]);
    wa0 = fabs(F[F_PWA0]);
    wa1 = fabs(F[F_PWA1]);
  } else {
    Wr = v.rho;
    wa0 = wa1 = w.rho_aux;
  }
  g0 = (v.naux >= 1) ? w.sig + wa0 * v.b0 * v.b0 : 1.0;
  g1 = (v.naux == 2) ? w.sig + wa1 * v.b1 * v.b1 : 1.0;
  den = g0 * g1 + Wr * (v.u0 * v.u0 * g1 + v.u1 * v.u1 * g0);
}
__device__ __forceinline__ double xbound_weight(const QpCtx& q, const SysW& w, int j) {
  if (w.polish) return fabs(q.zb[j]);  // zb holds the signed polish weights during polish
  return (q.ubs[j] - q.lbs[j] < kRhoTol) ? q.rho_eq : q.rho;
}

// K = P + sig I + A' W A with the aux variables eliminated; then factor.
__device__ inline bool assemble_factor(const QpCtx& q, const SysW& w) {
  const int N = q.N, HB = q.HB, Wd = HB + 1;
  for (int i = q.lane; i < N; i += 32) {
    for (int k = 0; k <= HB; ++k) q.Kb[i * Wd + k] = (k <= i) ? q.c * q.Dz[i] * q.Pband[i * Wd + k] * q.Dz[i - k] : 0.0;
    const double beta = q.Eb[i] * q.Dz[i];
    q.Kb[i * Wd] += w.sig + xbound_weight(q, w, i) * beta * beta;
  }
  __syncwarp();
  q.for_rows([&](int r) {
    const double* R = q.R(r);
    const double* F = R + q.CN;
    RowV v;
    row_view(q, R, q.I(r), v);
    double Wr, g0, g1, den;
    row_weights(q, w, F, v, Wr, g0, g1, den);
    if (Wr == 0.0) return;
    const double wr = Wr * g0 * g1 / den;  // Schur complement weight of the row on the trajectory block
    for (int i = 0; i < v.cnt; ++i) {
      const int vi = v.base + i * v.stride;
      const double ai = v.E * R[i] * q.Dz[vi];
      if (ai == 0.0) continue;
      for (int j = 0; j <= i; ++j) {
        const int vj = v.base + j * v.stride;
        q.Kb[vi * Wd + (vi - vj)] += wr * ai * (v.E * R[j] * q.Dz[vj]);
      }
    }
  });
  return band_factor(q);
}

// Reduce step of the aux elimination: rows hold their aux right-hand sides in F_RA*; adds the Schur
// correction to the trajectory right-hand side in v1.
__device__ __forceinline__ void row_reduce_rhs(const QpCtx& q, const double* R, const double* F, const RowV& v, double Wr,
                                               double g0, double g1, double den, double zcoef) {
  // zcoef: multiplier of the row on the trajectory part supplied by the caller (s_r for ADMM, -e_r for polish)
  const double ra0 = (v.naux >= 1) ? F[F_RA0] : 0.0, ra1 = (v.naux == 2) ? F[F_RA1] : 0.0;
  const double coef = zcoef - Wr * (v.u0 * ra0 * g1 + v.u1 * ra1 * g0) / den;
  for (int i = 0; i < v.cnt; ++i) {
    const int vi = v.base + i * v.stride;
    q.v1[vi] += v.E * R[i] * q.Dz[vi] * coef;
  }
}
// Back substitution of the aux variables after the banded solve (solution in v1).
__device__ __forceinline__ void row_backsub(const QpCtx& q, const double* R, const double* F, const RowV& v, double Wr,
                                            double g0, double g1, double den, double& zeta, double& a0, double& a1) {
  zeta = 0.0;
  for (int i = 0; i < v.cnt; ++i) {
    const int vi = v.base + i * v.stride;
    zeta += v.E * R[i] * q.Dz[vi] * q.v1[vi];
  }
  a0 = a1 = 0.0;
  if (v.naux == 0) return;
  const double ra0 = F[F_RA0], ra1 = (v.naux == 2) ? F[F_RA1] : 0.0;
  a0 = (g1 * (ra0 - Wr * v.u0 * zeta) + Wr * v.u1 * (v.u1 * ra0 - v.u0 * ra1)) / den;
  if (v.naux == 2) a1 = (g0 * (ra1 - Wr * v.u1 * zeta) + Wr * v.u0 * (v.u0 * ra1 - v.u1 * ra0)) / den;
}

struct QpOut {
  int status, iters, polish;
  double rho;
  double pri_res, dua_res, pol_pri, pol_dua, c;
  int pol_factor_ok, rho_updates, rounds;
};

// The whole QP solve for the calling warp's trajectory.  `warm`: rows/ws_* hold the previous solution.
__device__ inline QpOut qp_solve_warp(QpCtx& q, const QpSettings& st, bool warm, double warm_rho, double* ws_x,
                                      double* ws_yb, int n_aux_total) {
  const int N = q.N, lane = q.lane;
  QpOut out{QPS_UNSOLVED, 0, 0, st.rho, 0, 0, 0, 0, 0, -1, 0, 0};
  // ------------------------------------------------------------------ Ruiz equilibration (scale_data) [EXT]
  q.c = 1.0;
  for (int i = lane; i < N; i += 32) {
    q.Dz[i] = 1.0;
    q.Eb[i] = 1.0;
  }
  q.for_rows([&](int r) {
    double* F = q.R(r) + q.CN;
    F[F_E] = 1.0;
    F[F_DA0] = F[F_DA1] = F[F_EA0] = F[F_EA1] = 1.0;
  });
  const int W = q.HB + 1;
  for (int pass = 0; pass < st.scaling; ++pass) {
    // column norms of [P A'; A 0] restricted to the trajectory variables -> v2
    for (int i = lane; i < N; i += 32) {
      double m = 0.0;
      for (int k = 0; k <= q.HB && k <= i; ++k) m = fmax(m, fabs(q.c * q.Dz[i] * q.Pband[i * W + k] * q.Dz[i - k]));
      for (int k = 1; k <= q.HB && i + k < N; ++k) m = fmax(m, fabs(q.c * q.Dz[i] * q.Pband[(i + k) * W + k] * q.Dz[i + k]));
      m = fmax(m, fabs(q.Eb[i] * q.Dz[i]));
      q.v2[i] = m;
    }
    __syncwarp();
    // rows: contribute to column norms, compute own row / aux norms, update own E / aux scalings
    q.for_rows([&](int r) {
      double* R = q.R(r);
      double* F = R + q.CN;
      const int* I = q.I(r);
      const int base = I[RI_BASE], cnt = I[RI_CNT], stride = I[RI_STRIDE], aux = I[RI_AUX];
      const double E = F[F_E];
      double rn = 0.0;
      for (int i = 0; i < cnt; ++i) {
        const int vi = base + i * stride;
        const double a = fabs(E * R[i] * q.Dz[vi]);
        rn = fmax(rn, a);
        q.v2[vi] = fmax(q.v2[vi], a);
      }
      double dt0 = 1.0, dt1 = 1.0, et0 = 1.0, et1 = 1.0;
      if (aux >= 1) {
        const double ua = fabs(E * F[F_DA0]), ba = fabs(F[F_EA0] * F[F_DA0]);
        rn = fmax(rn, ua);
        dt0 = 1.0 / sqrt(limit_scaling(fmax(ua, ba)));
        et0 = 1.0 / sqrt(limit_scaling(ba));
      }
      if (aux == 2) {
        const double ua = fabs(E * F[F_DA1]), ba = fabs(F[F_EA1] * F[F_DA1]);
        rn = fmax(rn, ua);
        dt1 = 1.0 / sqrt(limit_scaling(fmax(ua, ba)));
        et1 = 1.0 / sqrt(limit_scaling(ba));
      }
      F[F_RA0] = 1.0 / sqrt(limit_scaling(rn));  // E_temp, applied below once the column pass is complete
      F[F_DA0] *= dt0;
      F[F_DA1] *= dt1;
      F[F_EA0] *= et0;
      F[F_EA1] *= et1;
    });
    q.for_rows([&](int r) {
      double* F = q.R(r) + q.CN;
      F[F_E] *= F[F_RA0];
    });
    for (int i = lane; i < N; i += 32) {
      const double bn = fabs(q.Eb[i] * q.Dz[i]);
      const double dt = 1.0 / sqrt(limit_scaling(q.v2[i]));
      q.Dz[i] *= dt;
      q.Eb[i] *= 1.0 / sqrt(limit_scaling(bn));
    }
    __syncwarp();
    // cost normalisation: mean column inf-norm of the scaled P over ALL n variables (aux columns are 0)
    double csum = 0.0, qn = 0.0;
    for (int i = lane; i < N; i += 32) {
      double m = 0.0;
      for (int k = 0; k <= q.HB && k <= i; ++k) m = fmax(m, fabs(q.c * q.Dz[i] * q.Pband[i * W + k] * q.Dz[i - k]));
      for (int k = 1; k <= q.HB && i + k < N; ++k) m = fmax(m, fabs(q.c * q.Dz[i] * q.Pband[(i + k) * W + k] * q.Dz[i + k]));
      csum += m;
      qn = fmax(qn, fabs(q.c * q.Dz[i] * q.qs[i]));
    }
    double qa = 0.0;
    q.for_rows([&](int r) {
      const double* F = q.R(r) + q.CN;
      const int aux = q.I(r)[RI_AUX];
      if (aux >= 1) qa = fmax(qa, fabs(q.c * F[F_DA0] * F[F_W]));
      if (aux == 2) qa = fmax(qa, fabs(q.c * F[F_DA1] * F[F_W]));
    });
    csum = warp_sum(csum);
    qn = warp_max(fmax(qn, qa));
    const double mean = limit_scaling(csum / static_cast<double>(N + n_aux_total));
    const double ct = 1.0 / fmax(mean, limit_scaling(qn));
    q.c *= ct;
  }
  q.cinv = 1.0 / q.c;
  // scaled cost vector and bounds (qs / lbs / ubs enter unscaled)
  for (int i = lane; i < N; i += 32) {
    q.qs[i] = q.c * q.Dz[i] * q.qs[i];
    q.lbs[i] *= q.Eb[i];
    q.ubs[i] *= q.Eb[i];
  }
  __syncwarp();

  // ------------------------------------------------------------------ initial iterate
  double rho = warm ? warm_rho : st.rho;
  rho = fmin(fmax(rho, kRhoMin), kRhoMax);
  q.rho = rho;
  q.rho_eq = kRhoEqOverIneq * rho;
  q.sigma = st.sigma;
  q.alpha = st.alpha;
  if (warm) {  // osqp_warm_start: x <- Dinv x, y <- c Einv y, z <- A x
    for (int i = lane; i < N; i += 32) {
      q.x[i] = ws_x[i] / q.Dz[i];
      q.yb[i] = ws_yb[i] / q.Eb[i] * q.c;
      q.zb[i] = q.Eb[i] * q.Dz[i] * q.x[i];
    }
    __syncwarp();
    q.for_rows([&](int r) {
      double* R = q.R(r);
      double* F = R + q.CN;
      RowV v;
      row_view(q, R, q.I(r), v);
      F[F_XA0] = (v.naux >= 1) ? F[F_XA0] / F[F_DA0] : 0.0;
      F[F_XA1] = (v.naux == 2) ? F[F_XA1] / F[F_DA1] : 0.0;
      F[F_Y] = F[F_Y] / v.E * q.c;
      F[F_YA0] = (v.naux >= 1) ? F[F_YA0] / F[F_EA0] * q.c : 0.0;
      F[F_YA1] = (v.naux == 2) ? F[F_YA1] / F[F_EA1] * q.c : 0.0;
      double ax = 0.0;
      for (int i = 0; i < v.cnt; ++i) {
        const int vi = v.base + i * v.stride;
        ax += v.E * R[i] * q.Dz[vi] * q.x[vi];
      }
      F[F_Z] = ax + v.u0 * F[F_XA0] + v.u1 * F[F_XA1];
      F[F_ZA0] = v.b0 * F[F_XA0];
      F[F_ZA1] = v.b1 * F[F_XA1];
    });
  } else {
    for (int i = lane; i < N; i += 32) q.x[i] = q.zb[i] = q.yb[i] = 0.0;
    q.for_rows([&](int r) {
      double* F = q.R(r) + q.CN;
      F[F_XA0] = F[F_XA1] = F[F_Z] = F[F_Y] = F[F_ZA0] = F[F_ZA1] = F[F_YA0] = F[F_YA1] = 0.0;
    });
  }
  __syncwarp();

  SysW sysw{false, st.sigma, rho};
  if (!assemble_factor(q, sysw)) {
    out.status = QPS_NONCVX;
    return out;
  }

  // ------------------------------------------------------------------ ADMM iterations
  double* dxs = q.scratch;          // [N] last trajectory step (written on check iterations)
  double* dyb = q.scratch + N;      // [N] last dual step of the variable-bound rows
  double* st_x = q.scratch + 2 * N;   // ADMM x, zb, yb stashed while polish reuses the shared vectors
  double* st_zb = q.scratch + 3 * N;
  double* st_yb = q.scratch + 4 * N;
  double pri_res = 0.0, dua_res = 0.0;
  int status = QPS_UNSOLVED, iter = 0;
  // residuals / norms gathered by the info pass
  double n_z = 0, n_ax = 0, n_q = 0, n_aty = 0, n_px = 0, s_pri = 0, s_dua = 0, s_z = 0, s_ax = 0, s_q = 0, s_aty = 0, s_px = 0;

  auto info_pass = [&]() {  // update_info(): v1 <- P x, v2 <- A'y (trajectory part), all norms
    p_matvec(q, q.x, q.v1);
    double m_pri = 0, m_z = 0, m_ax = 0, m_dua_a = 0, m_aty_a = 0, m_q_a = 0;
    double ms_pri = 0, ms_z = 0, ms_ax = 0, ms_dua_a = 0, ms_aty_a = 0, ms_q_a = 0;
    for (int i = lane; i < N; i += 32) {
      const double beta = q.Eb[i] * q.Dz[i];
      const double ax = beta * q.x[i];
      q.v2[i] = beta * q.yb[i];
      const double einv = 1.0 / q.Eb[i];
      m_pri = fmax(m_pri, fabs(einv * (ax - q.zb[i])));
      m_z = fmax(m_z, fabs(einv * q.zb[i]));
      m_ax = fmax(m_ax, fabs(einv * ax));
      ms_pri = fmax(ms_pri, fabs(ax - q.zb[i]));
      ms_z = fmax(ms_z, fabs(q.zb[i]));
      ms_ax = fmax(ms_ax, fabs(ax));
    }
    __syncwarp();
    q.for_rows([&](int r) {
      const double* R = q.R(r);
      const double* F = R + q.CN;
      RowV v;
      row_view(q, R, q.I(r), v);
      double ax = 0.0;
      for (int i = 0; i < v.cnt; ++i) {
        const int vi = v.base + i * v.stride;
        const double a = v.E * R[i] * q.Dz[vi];
        ax += a * q.x[vi];
        q.v2[vi] += a * F[F_Y];
      }
      ax += v.u0 * F[F_XA0] + v.u1 * F[F_XA1];
      const double einv = 1.0 / v.E;
      m_pri = fmax(m_pri, fabs(einv * (ax - F[F_Z])));
      m_z = fmax(m_z, fabs(einv * F[F_Z]));
      m_ax = fmax(m_ax, fabs(einv * ax));
      ms_pri = fmax(ms_pri, fabs(ax - F[F_Z]));
      ms_z = fmax(ms_z, fabs(F[F_Z]));
      ms_ax = fmax(ms_ax, fabs(ax));
      for (int k = 0; k < v.naux; ++k) {
        const double u = k ? v.u1 : v.u0, bb = k ? v.b1 : v.b0, qa = k ? v.qa1 : v.qa0;
        const double xa = F[F_XA0 + k], za = F[F_ZA0 + k], ya = F[F_YA0 + k];
        const double da = F[F_DA0 + k], ea = F[F_EA0 + k];
        const double axb = bb * xa;
        m_pri = fmax(m_pri, fabs((axb - za) / ea));
        m_z = fmax(m_z, fabs(za / ea));
        m_ax = fmax(m_ax, fabs(axb / ea));
        ms_pri = fmax(ms_pri, fabs(axb - za));
        ms_z = fmax(ms_z, fabs(za));
        ms_ax = fmax(ms_ax, fabs(axb));
        const double aty = u * F[F_Y] + bb * ya;
        m_dua_a = fmax(m_dua_a, fabs((qa + aty) / da));
        m_aty_a = fmax(m_aty_a, fabs(aty / da));
        m_q_a = fmax(m_q_a, fabs(qa / da));
        ms_dua_a = fmax(ms_dua_a, fabs(qa + aty));
        ms_aty_a = fmax(ms_aty_a, fabs(aty));
        ms_q_a = fmax(ms_q_a, fabs(qa));
      }
    });
    double m_dua = m_dua_a, m_aty = m_aty_a, m_q = m_q_a, m_px = 0, ms_dua = ms_dua_a, ms_aty = ms_aty_a, ms_q = ms_q_a, ms_px = 0;
    for (int i = lane; i < N; i += 32) {
      const double dinv = 1.0 / q.Dz[i];
      const double d = q.qs[i] + q.v1[i] + q.v2[i];
      m_dua = fmax(m_dua, fabs(dinv * d));
      m_aty = fmax(m_aty, fabs(dinv * q.v2[i]));
      m_q = fmax(m_q, fabs(dinv * q.qs[i]));
      m_px = fmax(m_px, fabs(dinv * q.v1[i]));
      ms_dua = fmax(ms_dua, fabs(d));
      ms_aty = fmax(ms_aty, fabs(q.v2[i]));
      ms_q = fmax(ms_q, fabs(q.qs[i]));
      ms_px = fmax(ms_px, fabs(q.v1[i]));
    }
    pri_res = warp_max(m_pri);
    dua_res = warp_max(m_dua) * q.cinv;
    n_z = warp_max(m_z); n_ax = warp_max(m_ax); n_q = warp_max(m_q); n_aty = warp_max(m_aty); n_px = warp_max(m_px);
    s_pri = warp_max(ms_pri); s_dua = warp_max(ms_dua); s_z = warp_max(ms_z); s_ax = warp_max(ms_ax);
    s_q = warp_max(ms_q); s_aty = warp_max(ms_aty); s_px = warp_max(ms_px);
  };

  auto primal_infeasible = [&](double eps) -> bool {  // is_primal_infeasible [EXT]
    // projected dual step, its E-scaled norm and the support function of [l,u]
    double nd = 0.0, lhs = 0.0;
    for (int i = lane; i < N; i += 32) {  // variable-bound rows: both bounds finite
      const double d = dyb[i];
      nd = fmax(nd, fabs(q.Eb[i] * d));
      lhs += q.ubs[i] * fmax(d, 0.0) + q.lbs[i] * fmin(d, 0.0);
      q.v1[i] = q.Eb[i] * q.Dz[i] * d;  // A' dy accumulates in v1
    }
    __syncwarp();
    double na = 0.0;  // inf-norm of Dinv A'dy over the aux columns
    q.for_rows([&](int r) {
      const double* R = q.R(r);
      const double* F = R + q.CN;
      RowV v;
      row_view(q, R, q.I(r), v);
      double d = F[F_DY];
      if (v.naux == AUX_HINGE) d = fmax(d, 0.0);  // l = -inf
      nd = fmax(nd, fabs(v.E * d));
      lhs += v.up * fmax(d, 0.0) + v.lo * fmin(d, 0.0);
      for (int i = 0; i < v.cnt; ++i) {
        const int vi = v.base + i * v.stride;
        q.v1[vi] += v.E * R[i] * q.Dz[vi] * d;
      }
      for (int k = 0; k < v.naux; ++k) {
        const double da = fmin(F[F_DYA0 + k], 0.0);  // aux bound rows: u = +inf
        nd = fmax(nd, fabs(F[F_EA0 + k] * da));
        // l = 0: no contribution to lhs
        const double u = k ? v.u1 : v.u0, bb = k ? v.b1 : v.b0;
        na = fmax(na, fabs((u * d + bb * da) / F[F_DA0 + k]));
      }
    });
    nd = warp_max(nd);
    lhs = warp_sum(lhs);
    if (nd > eps) {
      if (lhs < -eps * nd) {
        double m = na;
        for (int i = lane; i < N; i += 32) m = fmax(m, fabs(q.v1[i] / q.Dz[i]));
        m = warp_max(m);
        return m < eps * nd;
      }
    }
    return false;
  };
  auto dual_infeasible = [&](double eps) -> bool {  // is_dual_infeasible [EXT]
    double ndx = 0.0, qdx = 0.0;
    for (int i = lane; i < N; i += 32) {
      ndx = fmax(ndx, fabs(q.Dz[i] * dxs[i]));
      qdx += q.qs[i] * dxs[i];
      q.v2[i] = dxs[i];
    }
    q.for_rows([&](int r) {
      const double* F = q.R(r) + q.CN;
      RowV v;
      row_view(q, q.R(r), q.I(r), v);
      for (int k = 0; k < v.naux; ++k) {
        ndx = fmax(ndx, fabs(F[F_DA0 + k] * F[F_DXA0 + k]));
        qdx += (k ? v.qa1 : v.qa0) * F[F_DXA0 + k];
      }
    });
    ndx = warp_max(ndx);
    qdx = warp_sum(qdx);
    if (!(ndx > eps)) return false;
    if (!(qdx < -q.c * eps * ndx)) return false;
    p_matvec(q, q.v2, q.v1);
    double m = 0.0;
    for (int i = lane; i < N; i += 32) m = fmax(m, fabs(q.v1[i] / q.Dz[i]));
    m = warp_max(m);
    if (!(m < q.c * eps * ndx)) return false;
    int bad = 0;
    for (int i = lane; i < N; i += 32) {  // both bounds finite
      const double vv = q.Dz[i] * dxs[i];  // Einv * (Eb Dz dx)
      if (vv > eps * ndx || vv < -eps * ndx) bad = 1;
    }
    q.for_rows([&](int r) {
      const double* R = q.R(r);
      const double* F = R + q.CN;
      RowV v;
      row_view(q, R, q.I(r), v);
      double ax = 0.0;
      for (int i = 0; i < v.cnt; ++i) {
        const int vi = v.base + i * v.stride;
        ax += v.E * R[i] * q.Dz[vi] * q.v2[vi];
      }
      ax += v.u0 * F[F_DXA0] + v.u1 * F[F_DXA1];
      const double vv = ax / v.E;
      if (vv > eps * ndx) bad = 1;                               // u finite for every row
      if (v.naux != AUX_HINGE && vv < -eps * ndx) bad = 1;       // l finite unless hinge
      for (int k = 0; k < v.naux; ++k) {
        const double va = (k ? v.b1 : v.b0) * F[F_DXA0 + k] / F[F_EA0 + k];
        if (va < -eps * ndx) bad = 1;                            // aux rows: l = 0 finite, u infinite
      }
    });
    return warp_sum_int(bad) == 0;
  };
  double eps_scale = 1.0;  // tightened by the verified-polish rounds (DESIGN.md deviation D2)
  auto check_termination = [&](bool approximate) -> int {
    double eps_abs = st.eps_abs * eps_scale, eps_rel = st.eps_rel * eps_scale, epi = st.eps_prim_inf, edi = st.eps_dual_inf;
    if (approximate) {
      eps_abs *= 10; eps_rel *= 10; epi *= 10; edi *= 10;
    }
    if (pri_res > kOsqpInf || dua_res > kOsqpInf) return QPS_NONCVX;
    const double eps_pri = eps_abs + eps_rel * fmax(n_z, n_ax);
    const double eps_dua = eps_abs + eps_rel * q.cinv * fmax(n_q, fmax(n_aty, n_px));
    const bool pri_ok = pri_res < eps_pri, dua_ok = dua_res < eps_dua;
    bool pinf = false, dinf = false;
    if (!pri_ok) pinf = primal_infeasible(epi);
    if (!dua_ok) dinf = dual_infeasible(edi);
    if (pri_ok && dua_ok) return approximate ? QPS_SOLVED_INACC : QPS_SOLVED;
    if (pinf) return approximate ? QPS_PINF_INACC : QPS_PINF;
    if (dinf) return approximate ? QPS_DINF_INACC : QPS_DINF;
    return QPS_UNSOLVED;
  };

  // ADMM iterations, continuing from the current state until a termination test fires or max_iter.
  auto run_admm = [&]() {
    status = QPS_UNSOLVED;
    while (iter < st.max_iter) {
      ++iter;
      const bool can_check = st.check_termination > 0 && (iter % st.check_termination == 0);
      const bool rho_iter = st.adaptive_rho && st.adaptive_rho_interval > 0 && (iter % st.adaptive_rho_interval == 0);
      const bool keep_steps = can_check || iter == st.max_iter;
      // ---- right-hand side:  sigma x - q + A'(rho z - y), aux part eliminated -------------------------
      for (int i = lane; i < N; i += 32) {
        const double beta = q.Eb[i] * q.Dz[i];
        const double rb = (q.ubs[i] - q.lbs[i] < kRhoTol) ? q.rho_eq : q.rho;
        q.v1[i] = q.sigma * q.x[i] - q.qs[i] + beta * (rb * q.zb[i] - q.yb[i]);
      }
      __syncwarp();
      q.for_rows([&](int r) {
        double* R = q.R(r);
        double* F = R + q.CN;
        RowV v;
        row_view(q, R, q.I(r), v);
        double Wr, g0, g1, den;
        row_weights(q, sysw, F, v, Wr, g0, g1, den);
        const double s = Wr * F[F_Z] - F[F_Y];
        if (v.naux >= 1) F[F_RA0] = q.sigma * F[F_XA0] - v.qa0 + v.u0 * s + v.b0 * (sysw.rho_aux * F[F_ZA0] - F[F_YA0]);
        if (v.naux == 2) F[F_RA1] = q.sigma * F[F_XA1] - v.qa1 + v.u1 * s + v.b1 * (sysw.rho_aux * F[F_ZA1] - F[F_YA1]);
        row_reduce_rhs(q, R, F, v, Wr, g0, g1, den, s);
      });
      band_solve(q, q.v1);
      // ---- rows: back-substitute aux, relax, project, dual update ----------------------------------------
      q.for_rows([&](int r) {
        double* R = q.R(r);
        double* F = R + q.CN;
        RowV v;
        row_view(q, R, q.I(r), v);
        double Wr, g0, g1, den, zeta, a0, a1;
        row_weights(q, sysw, F, v, Wr, g0, g1, den);
        row_backsub(q, R, F, v, Wr, g0, g1, den, zeta, a0, a1);
        const double zt = zeta + v.u0 * a0 + v.u1 * a1;
        {
          const double zr = q.alpha * zt + (1.0 - q.alpha) * F[F_Z];
          double zn = zr + F[F_Y] / Wr;
          zn = fmin(fmax(zn, v.lo), v.up);
          const double dy = Wr * (zr - zn);
          F[F_Z] = zn;
          F[F_Y] += dy;
          F[F_DY] = dy;
        }
        for (int k = 0; k < v.naux; ++k) {
          const double at = k ? a1 : a0, bb = k ? v.b1 : v.b0;
          const double xo = F[F_XA0 + k];
          const double xn = q.alpha * at + (1.0 - q.alpha) * xo;
          F[F_XA0 + k] = xn;
          F[F_DXA0 + k] = xn - xo;
          const double zr = q.alpha * (bb * at) + (1.0 - q.alpha) * F[F_ZA0 + k];
          double zn = zr + F[F_YA0 + k] / sysw.rho_aux;
          zn = fmin(fmax(zn, 0.0), kOsqpInf * F[F_EA0 + k]);
          const double dy = sysw.rho_aux * (zr - zn);
          F[F_ZA0 + k] = zn;
          F[F_YA0 + k] += dy;
          F[F_DYA0 + k] = dy;
        }
      });
      // ---- trajectory variables and their bound rows -----------------------------------------------------
      for (int i = lane; i < N; i += 32) {
        const double beta = q.Eb[i] * q.Dz[i];
        const double rb = (q.ubs[i] - q.lbs[i] < kRhoTol) ? q.rho_eq : q.rho;
        const double xt = q.v1[i];
        const double xn = q.alpha * xt + (1.0 - q.alpha) * q.x[i];
        const double zr = q.alpha * (beta * xt) + (1.0 - q.alpha) * q.zb[i];
        double zn = zr + q.yb[i] / rb;
        zn = fmin(fmax(zn, q.lbs[i]), q.ubs[i]);
        const double dy = rb * (zr - zn);
        if (keep_steps) {
          dxs[i] = xn - q.x[i];
          dyb[i] = dy;
        }
        q.x[i] = xn;
        q.zb[i] = zn;
        q.yb[i] += dy;
      }
      __syncwarp();
      if (can_check) {
        info_pass();
        status = check_termination(false);
        if (status != QPS_UNSOLVED) return;
      }
      if (rho_iter) {
        if (!can_check) info_pass();
        // compute_rho_estimate on the scaled quantities [EXT]
        const double pn = s_pri / (fmax(s_z, s_ax) + 1e-10);
        const double dn = s_dua / (fmax(s_q, fmax(s_aty, s_px)) + 1e-10);
        double rho_new = rho * sqrt(pn / (dn + 1e-10));
        rho_new = fmin(fmax(rho_new, kRhoMin), kRhoMax);
        if (rho_new > rho * st.adaptive_rho_tolerance || rho_new < rho / st.adaptive_rho_tolerance) {
          rho = rho_new;
          q.rho = rho;
          q.rho_eq = kRhoEqOverIneq * rho;
          sysw.rho_aux = rho;
          out.rho_updates++;
          if (!assemble_factor(q, sysw)) {
            status = QPS_NONCVX;
            return;
          }
        }
      }
    }
    // max_iter reached without a verdict: approximate test, then MAX_ITER_REACHED
    if (!(st.check_termination > 0 && (iter % st.check_termination == 0))) info_pass();
    status = check_termination(true);
    if (status == QPS_UNSOLVED) status = QPS_MAXITER;
  };

  // ---- polish (OSQP polish.c [EXT]) ---------------------------------------------------------------------
  // Equality-constrained QP on the guessed active set, solved as the delta-regularised KKT system with
  // iterative refinement, in its reduced form K_p = P + delta I + (1/delta) A_act' A_act (same aux
  // elimination and banded factor as the ADMM system).  Returns false when K_p could not be factored.
  // `verified`: the polished point is primal feasible to verify_tol and every active inequality row has a
  // correctly signed multiplier, i.e. it is a KKT point of the QP = the unique minimiser.
  const double wp = 1.0 / st.delta;
  const SysW pw{true, st.delta, 0.0};
  auto polish_once = [&](bool& verified, double& p_pri, double& p_dua) -> bool {
    verified = false;
    for (int i = lane; i < N; i += 32) {
      st_x[i] = q.x[i];
      st_zb[i] = q.zb[i];
      st_yb[i] = q.yb[i];
      const double z = q.zb[i], y = q.yb[i];
      double w = 0.0;
      if (z - q.lbs[i] < -y) w = -wp;           // lower active
      else if (q.ubs[i] - z < y) w = wp;        // upper active
      q.zb[i] = w;                               // signed polish weight
      q.x[i] = 0.0;                              // polish iterate
      q.yb[i] = 0.0;                             // polish multiplier
    }
    q.for_rows([&](int r) {
      double* R = q.R(r);
      double* F = R + q.CN;
      RowV v;
      row_view(q, R, q.I(r), v);
      double w = 0.0, b = 0.0;
      if (F[F_Z] - v.lo < -F[F_Y]) { w = -wp; b = v.lo; }
      else if (v.up - F[F_Z] < F[F_Y]) { w = wp; b = v.up; }
      F[F_PW] = w;
      F[F_PB] = b;
      for (int k = 0; k < 2; ++k) {
        double wa = 0.0;
        if (k < v.naux) {
          if (F[F_ZA0 + k] - 0.0 < -F[F_YA0 + k]) wa = -wp;                                    // lower (0) active
          else if (kOsqpInf * F[F_EA0 + k] - F[F_ZA0 + k] < F[F_YA0 + k]) wa = wp;            // never in practice
        }
        F[F_PWA0 + k] = wa;
        F[F_PYA0 + k] = 0.0;
        F[F_PX0 + k] = 0.0;
      }
      F[F_PY] = 0.0;
    });
    __syncwarp();
    if (!assemble_factor(q, pw)) return false;
    for (int it = 0; it <= st.polish_refine_iter + 1; ++it) {
      const bool last = (it == st.polish_refine_iter + 1);  // final pass: pending dual update + residuals only
      // v1 <- P xq ; residual rd = -(P x + q + A'y) - A' W (A x - b), y update of the previous step folded in
      p_matvec(q, q.x, q.v1);
      double m_pri = 0.0, m_dua_a = 0.0;
      int bad_sign = 0;
      for (int i = lane; i < N; i += 32) {
        const double beta = q.Eb[i] * q.Dz[i];
        const double ax = beta * q.x[i];
        const double w = fabs(q.zb[i]);
        const double bnd = q.zb[i] > 0 ? q.ubs[i] : q.lbs[i];
        if (it > 0 && w != 0.0) q.yb[i] += w * (ax - bnd);
        const double e = q.yb[i] + (last ? 0.0 : w * (ax - bnd));
        const double zc = fmin(fmax(ax, q.lbs[i]), q.ubs[i]);
        m_pri = fmax(m_pri, fabs((ax - zc) / q.Eb[i]));
        if (last && w != 0.0 && q.ubs[i] - q.lbs[i] >= kRhoTol) {
          if (q.zb[i] > 0 && q.yb[i] < -kVerifyTol) bad_sign = 1;
          if (q.zb[i] < 0 && q.yb[i] > kVerifyTol) bad_sign = 1;
        }
        q.v2[i] = q.v1[i] + q.qs[i] + beta * q.yb[i];  // dual residual (uses y only)
        q.v1[i] = -(q.v1[i] + q.qs[i]) - beta * e;
      }
      __syncwarp();
      q.for_rows([&](int r) {
        double* R = q.R(r);
        double* F = R + q.CN;
        RowV v;
        row_view(q, R, q.I(r), v);
        double Wr, g0, g1, den;
        row_weights(q, pw, F, v, Wr, g0, g1, den);
        double ax = 0.0;
        for (int i = 0; i < v.cnt; ++i) {
          const int vi = v.base + i * v.stride;
          ax += v.E * R[i] * q.Dz[vi] * q.x[vi];
        }
        ax += v.u0 * F[F_PX0] + v.u1 * F[F_PX1];
        if (it > 0 && Wr != 0.0) F[F_PY] += Wr * (ax - F[F_PB]);
        const double e = F[F_PY] + (last ? 0.0 : Wr * (ax - F[F_PB]));
        const double zc = fmin(fmax(ax, v.lo), v.up);
        m_pri = fmax(m_pri, fabs((ax - zc) / v.E));
        if (last && Wr != 0.0 && v.naux == AUX_HINGE) {  // inequality row (l = -inf): upper active needs y >= 0
          if (F[F_PY] < -kVerifyTol) bad_sign = 1;
        }
        double ea[2] = {0.0, 0.0};
        for (int k = 0; k < v.naux; ++k) {
          const double bb = k ? v.b1 : v.b0, u = k ? v.u1 : v.u0, qa = k ? v.qa1 : v.qa0;
          const double wa = fabs(F[F_PWA0 + k]);
          const double axb = bb * F[F_PX0 + k];
          if (it > 0 && wa != 0.0) F[F_PYA0 + k] += wa * (axb - 0.0);
          ea[k] = F[F_PYA0 + k] + (last ? 0.0 : wa * axb);
          const double zca = fmax(axb, 0.0);
          m_pri = fmax(m_pri, fabs((axb - zca) / F[F_EA0 + k]));
          if (last && wa != 0.0 && F[F_PYA0 + k] > kVerifyTol) bad_sign = 1;  // aux >= 0 held at 0 needs y <= 0
          m_dua_a = fmax(m_dua_a, fabs((qa + u * F[F_PY] + bb * F[F_PYA0 + k]) / F[F_DA0 + k]));
          F[F_RA0 + k] = -qa - u * e - bb * ea[k];
        }
        for (int i = 0; i < v.cnt; ++i) {
          const int vi = v.base + i * v.stride;
          q.v2[vi] += v.E * R[i] * q.Dz[vi] * F[F_PY];
        }
        if (!last) row_reduce_rhs(q, R, F, v, Wr, g0, g1, den, -e);
      });
      if (last) {
        double m_dua = m_dua_a;
        for (int i = lane; i < N; i += 32) m_dua = fmax(m_dua, fabs(q.v2[i] / q.Dz[i]));
        p_pri = warp_max(m_pri);
        p_dua = warp_max(m_dua) * q.cinv;
        const int nbad = warp_sum_int(bad_sign);
        verified = (nbad == 0) && (p_pri <= kVerifyTol) && isfinite(p_pri) && isfinite(p_dua);
        break;
      }
      band_solve(q, q.v1);
      q.for_rows([&](int r) {
        double* R = q.R(r);
        double* F = R + q.CN;
        RowV v;
        row_view(q, R, q.I(r), v);
        double Wr, g0, g1, den, zeta, a0, a1;
        row_weights(q, pw, F, v, Wr, g0, g1, den);
        row_backsub(q, R, F, v, Wr, g0, g1, den, zeta, a0, a1);
        F[F_PX0] += a0;
        F[F_PX1] += a1;
      });
      for (int i = lane; i < N; i += 32) q.x[i] += q.v1[i];
      __syncwarp();
    }
    return true;
  };
  auto restore_admm_state = [&](bool keep_polished_x) {
    for (int i = lane; i < N; i += 32) {
      if (!keep_polished_x) q.x[i] = st_x[i];
      q.zb[i] = st_zb[i];
      q.yb[i] = st_yb[i];
    }
    __syncwarp();
  };

  // ---- main loop: ADMM -> polish -> verify; on a failed verification ADMM continues with 10x tighter ------
  // tolerances (DESIGN.md deviation D2; verify_rounds = 0 is plain OSQP behaviour).
  int round = 0;
  while (true) {
    run_admm();
    out.pri_res = pri_res;
    out.dua_res = dua_res;
    if (status != QPS_SOLVED || !st.polishing) break;
    bool verified = false;
    double p_pri = 0.0, p_dua = 0.0;
    const bool factored = polish_once(verified, p_pri, p_dua);
    out.pol_factor_ok = factored ? 1 : 0;
    out.pol_pri = p_pri;
    out.pol_dua = p_dua;
    out.rounds = round;
    if (factored && verified) {
      out.polish = 1;
      break;
    }
    if (round >= kVerifyRounds || iter >= st.max_iter) {  // OSQP's acceptance rule
      const bool ok = factored && ((p_pri < pri_res && p_dua < dua_res) || (p_pri < pri_res && dua_res < 1e-10) ||
                                   (p_dua < dua_res && pri_res < 1e-10)) && isfinite(p_pri) && isfinite(p_dua);
      out.polish = ok ? 2 : -1;
      break;
    }
    ++round;
    eps_scale *= 0.1;
    restore_admm_state(false);
    if (!assemble_factor(q, sysw)) {  // back to the ADMM factor
      status = QPS_NONCVX;
      break;
    }
  }
  out.iters = iter;
  out.status = status;
  out.rho = rho;
  out.c = q.c;
  if (out.polish != 0) {
    // Adopt the polished PRIMAL point when accepted.  The duals kept for the next warm start are always the
    // ADMM duals: polished duals are non-unique on degenerate active sets (DESIGN.md deviation D1).
    if (out.polish > 0) {
      q.for_rows([&](int r) {
        double* F = q.R(r) + q.CN;
        for (int k = 0; k < 2; ++k) F[F_XA0 + k] = F[F_PX0 + k];
      });
    }
    restore_admm_state(out.polish > 0);
  }
  return out;
}

}  // namespace tb200

namespace tb200 {

// ---------------------------------------------------------------------------------------------------
// Kernel: QP assembly (optimizers.cpp:781-799 + osqp_interface.cpp:170-281 in fixed layout) + solve.
// grid = B, block = 32 (one warp per trajectory).
__global__ void __launch_bounds__(32) qp_kernel(DevProblem p, const double* x_override /*kernel-level API*/,
                                                const double* trust_override, int* admm_iters_out,
                                                int* polish_out) {
  extern __shared__ double sm[];
  const int b = blockIdx.x;
  const int lane = threadIdx.x;
  if (!x_override && p.status[b] != 5) return;
  const int N = p.N, T = p.T, D = p.D, HB = p.HB;
  const QpSmem S = qp_smem_layout(N, HB, T, D);
  QpCtx q;
  q.N = N; q.HB = HB; q.T = T; q.D = D; q.lane = lane;
  q.CN = p.row_stride - F_NFIELDS;
  q.RS = p.row_stride;
  q.Kb = sm + S.Kb; q.Dz = sm + S.Dz; q.Eb = sm + S.Eb; q.qs = sm + S.qs; q.lbs = sm + S.lbs; q.ubs = sm + S.ubs;
  q.x = sm + S.x; q.zb = sm + S.zb; q.yb = sm + S.yb; q.v1 = sm + S.v1; q.v2 = sm + S.v2; q.invd = sm + S.invd;
  q.ls = reinterpret_cast<int*>(sm + S.ints);
  q.rows = p.rows + static_cast<size_t>(b) * p.max_rows * p.row_stride;
  q.rints = p.row_ints + static_cast<size_t>(b) * p.max_rows * RI_NINTS;
  int* mylist = p.lists + static_cast<size_t>(b) * (2 * p.max_rows + p.n_costs + p.n_cnts + 2);
  int* obj_start = mylist + 2 * p.max_rows;  // [n_costs + n_cnts + 1]
  q.list = mylist;
  q.Pband = p.Pband;
  q.scratch = p.scratch + static_cast<size_t>(b) * 5 * N;

  const double* xc = (x_override ? x_override : p.x) + static_cast<size_t>(b) * N;
  const double trust = trust_override ? trust_override[b] : p.trust[b];
  const double* mu = p.merit_coeffs + static_cast<size_t>(b) * p.n_cnts;
  const int buf = x_override ? 0 : p.cur_buf[b];
  const size_t slot = static_cast<size_t>(buf) * p.B + b;
  const double* cart_err = p.cart_err + slot * p.n_cart_rows;
  const double* cart_jac = p.cart_jac + slot * static_cast<size_t>(p.n_cart_rows) * p.cart_stride;
  const double* coll_rows = p.coll_rows + slot * static_cast<size_t>(p.n_coll_cand) * p.coll_stride;
  const unsigned long long* coll_mask = p.coll_mask + slot * static_cast<size_t>(p.n_coll_objs) * p.coll_words;

  // ---- trajectory part: x, trust box (setTrustBoxConstraints, optimizers.cpp:151-170), linear cost -----
  for (int i = lane; i < N; i += 32) {
    const double lb = p.lower[i % D], ub = p.upper[i % D];
    const double xi = fmin(fmax(xc[i], lb), ub);
    q.lbs[i] = fmax(fmax(xi - trust, lb), -kOsqpInf);
    q.ubs[i] = fmin(fmin(xi + trust, ub), kOsqpInf);
    q.qs[i] = p.qlin[i];
    q.x[i] = xc[i];  // linearisation point (until the solver takes over x)
  }
  __syncwarp();

  // ---- rows in the reference's canonical order: permanent rows, cost rows, penalised constraint rows -----
  int nr = 0, n_aux = 0, nnzA = 0;
  // (a) fixed_timesteps / fixed_dofs rows: x_k - init_k == 0
  for (int f = lane; f < p.n_fixed; f += 32) {
    const int var = p.fixed_vars[f];
    double* R = q.R(nr + f);
    int* I = q.rints + static_cast<size_t>(nr + f) * RI_NINTS;
    R[0] = 1.0;
    R[q.CN + F_C] = -p.init_traj[static_cast<size_t>(b) * N + var];
    R[q.CN + F_W] = 0.0;
    I[RI_BASE] = var; I[RI_CNT] = 1; I[RI_STRIDE] = D; I[RI_AUX] = AUX_NONE; I[RI_OBJ] = -1; I[RI_PAD] = -1 - (var % D);
  }
  nr += p.n_fixed;
  nnzA += p.n_fixed;
  // (b) objects
  const int n_obj = p.n_costs + p.n_cnts;
  int coll_obj_counter = 0;
  for (int oi = 0; oi < n_obj; ++oi) {
    const bool is_cnt = oi >= p.n_costs;
    const DevObj o = is_cnt ? p.cnt_objs[oi - p.n_costs] : p.cost_objs[oi];
    if (lane == 0) obj_start[oi] = nr;
    const double w_aux = is_cnt ? mu[oi - p.n_costs] : 1.0;
    if (o.kind == OBJ_JOINT_EQ_COST) continue;
    if (o.kind == OBJ_JOINT_EQ_CNT || o.kind == OBJ_JOINT_INEQ_CNT || o.kind == OBJ_JOINT_INEQ_COST) {
      const DevJointTerm& jt = p.joint_terms[o.term];
      const int per = (o.kind == OBJ_JOINT_EQ_CNT) ? 1 : 2;
      const int total = o.n_steps * D * per;
      const double wst[3][3] = {{1, 0, 0}, {-1, 1, 0}, {1, -2, 1}};
      for (int k = lane; k < total; k += 32) {
        const int t = o.first + k / (D * per), d = (k / per) % D, side = k % per;
        double* R = q.R(nr + k);
        int* I = q.rints + static_cast<size_t>(nr + k) * RI_NINTS;
        const double cd = jt.coeffs[d];
        double sgn = cd, cst;
        if (per == 1) cst = -jt.targets[d] * cd;
        else if (side == 0) cst = (-jt.targets[d] - jt.upper[d]) * cd;        // (e - upper) * c
        else { sgn = -cd; cst = (jt.lower[d] + jt.targets[d]) * cd; }         // (lower - e) * c
        for (int i = 0; i <= o.order; ++i) R[i] = wst[o.order][i] * sgn;
        R[q.CN + F_C] = cst;
        R[q.CN + F_W] = w_aux;
        I[RI_BASE] = t * D + d; I[RI_CNT] = o.order + 1; I[RI_STRIDE] = D;
        I[RI_AUX] = (per == 1) ? AUX_ABS : AUX_HINGE; I[RI_OBJ] = oi; I[RI_PAD] = -1 - d;
      }
      nr += total;
      n_aux += total * ((per == 1) ? 2 : 1);
      nnzA += total * (o.order + 1 + ((per == 1) ? 2 : 1));
    } else if (o.kind == OBJ_CART_POSE) {
      const DevCartTerm& ct = p.cart_terms[o.term];
      int nz = 0;
      for (int k = lane; k < o.n_rows; k += 32) {
        double* R = q.R(nr + k);
        int* I = q.rints + static_cast<size_t>(nr + k) * RI_NINTS;
        const double* J = cart_jac + static_cast<size_t>(o.src_off + k) * p.cart_stride;
        const double thr = 1e-7 * fabs(ct.coeff[k]);  // cleanupAff acts on the unscaled gradient (modeling_utils.cpp:31-39)
        double dot = 0.0;
        for (int j = 0; j < D; ++j) {
          dot += J[j] * q.x[o.first * D + j];
          const double a = (fabs(J[j]) > thr) ? J[j] : 0.0;
          R[j] = a;
          nz += (a != 0.0);
        }
        R[q.CN + F_C] = cart_err[o.src_off + k] - dot;
        R[q.CN + F_W] = w_aux;
        I[RI_BASE] = o.first * D; I[RI_CNT] = D; I[RI_STRIDE] = 1; I[RI_AUX] = AUX_ABS; I[RI_OBJ] = oi; I[RI_PAD] = o.first;
      }
      nz = warp_sum_int(nz);
      nr += o.n_rows;
      n_aux += 2 * o.n_rows;
      nnzA += nz + 2 * o.n_rows;
    } else if (o.kind == OBJ_COLL) {
      // active candidates of this timestep, in candidate order (ballot compaction)
      const unsigned long long* mw = coll_mask + static_cast<size_t>(coll_obj_counter) * p.coll_words;
      ++coll_obj_counter;
      int nz = 0, count = 0;
      for (int c0 = 0; c0 < o.n_rows; c0 += 32) {
        const int c = c0 + lane;
        const bool act = (c < o.n_rows) && ((mw[c / 64] >> (c % 64)) & 1ull);
        const unsigned bal = __ballot_sync(0xffffffffu, act);
        if (act) {
          const int pos = nr + count + __popc(bal & ((1u << lane) - 1u));
          double* R = q.R(pos);
          int* I = q.rints + static_cast<size_t>(pos) * RI_NINTS;
          const double* cr = coll_rows + static_cast<size_t>(o.src_off + c) * p.coll_stride;
          // dist(q) ~ d0 + g.(q - q0);  constraint: coeff*(margin - dist) <= 0;  cost: hinge(margin - dist)*coeff
          const double scale = is_cnt ? cr[D + 2] : 1.0;
          double dot = 0.0;
          for (int j = 0; j < D; ++j) {
            dot += cr[j] * q.x[o.first * D + j];
            const double a = -cr[j] * scale;
            R[j] = a;
            nz += (a != 0.0);
          }
          R[q.CN + F_C] = (cr[D + 1] - cr[D] + dot) * scale;
          R[q.CN + F_W] = is_cnt ? w_aux : cr[D + 2];
          I[RI_BASE] = o.first * D; I[RI_CNT] = D; I[RI_STRIDE] = 1; I[RI_AUX] = AUX_HINGE; I[RI_OBJ] = oi; I[RI_PAD] = o.first;
        }
        count += __popc(bal);
      }
      nz = warp_sum_int(nz);
      nr += count;
      n_aux += count;
      nnzA += nz + count;
    }
  }
  if (lane == 0) obj_start[n_obj] = nr;
  q.nrows = nr;
  nnzA += N + n_aux;  // identity rows carrying the variable bounds
  __syncwarp();
  __threadfence_block();

  // ---- per-lane row lists: dense slot t (RI_PAD == t), sparse slot d (RI_PAD == -1-d) -----------------
  {
    // counts
    int* ls = q.ls;
    for (int sidx = lane; sidx < T + D; sidx += 32) {
      const int key = (sidx < T) ? sidx : -1 - (sidx - T);
      int cnt = 0;
      for (int r = 0; r < nr; ++r) cnt += (q.rints[static_cast<size_t>(r) * RI_NINTS + RI_PAD] == key);
      ls[sidx + 1] = cnt;
    }
    if (lane == 0) ls[0] = 0;
    __syncwarp();
    if (lane == 0)
      for (int sidx = 0; sidx < T + D; ++sidx) ls[sidx + 1] += ls[sidx];
    __syncwarp();
    for (int sidx = lane; sidx < T + D; sidx += 32) {
      const int key = (sidx < T) ? sidx : -1 - (sidx - T);
      int pos = ls[sidx];
      for (int r = 0; r < nr; ++r)
        if (q.rints[static_cast<size_t>(r) * RI_NINTS + RI_PAD] == key) mylist[pos++] = r;
    }
    __syncwarp();
    __threadfence_block();
  }

  // ---- warm start decision (createOrUpdateSolver, osqp_interface.cpp:283-370) ---------------------------
  int* meta = p.ws_meta + static_cast<size_t>(b) * 4;
  const bool warm = !x_override && p.qp.warm_starting && meta[3] == 1 && meta[0] == n_aux && meta[1] == nr && meta[2] == nnzA;
  const double warm_rho = p.ws_rho[b];

  QpOut res = qp_solve_warp(q, p.qp, warm, warm_rho, p.ws_x + static_cast<size_t>(b) * N, p.ws_yb + static_cast<size_t>(b) * N, n_aux);

  // ---- unscale, store the solution (and the warm-start state), model values -----------------------------
  double* nx = p.new_x + static_cast<size_t>(b) * N;
  for (int i = lane; i < N; i += 32) {
    const double xu = q.Dz[i] * q.x[i];
    nx[i] = xu;
    q.v1[i] = xu;  // unscaled solution for the model-value pass
    p.ws_x[static_cast<size_t>(b) * N + i] = xu;
    p.ws_yb[static_cast<size_t>(b) * N + i] = q.cinv * q.Eb[i] * q.yb[i];
  }
  __syncwarp();
  q.for_rows([&](int r) {
    double* R = q.R(r);
    double* F = R + q.CN;
    const int* I = q.I(r);
    const int aux = I[RI_AUX];
    F[F_Y] = q.cinv * F[F_E] * F[F_Y];
    for (int k = 0; k < 2; ++k) {
      F[F_XA0 + k] = (k < aux) ? F[F_DA0 + k] * F[F_XA0 + k] : 0.0;
      F[F_YA0 + k] = (k < aux) ? q.cinv * F[F_EA0 + k] * F[F_YA0 + k] : 0.0;
    }
    double val = F[F_C];
    for (int i = 0; i < I[RI_CNT]; ++i) val += R[i] * q.v1[I[RI_BASE] + i * I[RI_STRIDE]];
    // ConvexConstraints::violations (modeling.cpp:132-142) for constraint rows; hinge/abs cost = w * aux values
    F[F_MV] = (aux == AUX_ABS || aux == AUX_NONE) ? fabs(val) : fmax(val, 0.0);
  });
  // per object sums, canonical order, one lane per object (deterministic)
  for (int oi = lane; oi < n_obj; oi += 32) {
    const bool is_cnt = oi >= p.n_costs;
    double s = 0.0;
    if (!is_cnt && p.cost_objs[oi].kind == OBJ_JOINT_EQ_COST) s = joint_obj_value(p, p.cost_objs[oi], q.v1);  // exact quadratic
    for (int r = obj_start[oi]; r < obj_start[oi + 1]; ++r) {
      const double* F = q.R(r) + q.CN;
      if (is_cnt) s += F[F_MV];
      else s += F[F_W] * (F[F_XA0] + F[F_XA1]);  // ConvexObjective::value: the penalty terms use the aux values
    }
    if (is_cnt) p.model_cnt_viols[static_cast<size_t>(b) * p.n_cnts + (oi - p.n_costs)] = s;
    else p.model_cost_vals[static_cast<size_t>(b) * p.n_costs + oi] = s;
  }
  if (lane == 0) {
    // status map of osqp_interface.cpp:565-614
    int cvx = 2;
    if (res.status == QPS_SOLVED || res.status == QPS_SOLVED_INACC) cvx = 0;
    else if (res.status >= QPS_PINF && res.status <= QPS_DINF_INACC) cvx = 1;
    p.qp_status[b] = cvx;
    meta[0] = n_aux; meta[1] = nr; meta[2] = nnzA; meta[3] = (cvx == 0) ? 1 : 0;
    p.ws_rho[b] = res.rho;
    if (!x_override) p.n_admm_iters[b] += res.iters;
    if (admm_iters_out) admm_iters_out[b] = res.iters;
    if (polish_out) polish_out[b] = res.polish;
    double* g = p.dbg + static_cast<size_t>(b) * 16;
    g[0] = res.status; g[1] = res.iters; g[2] = res.polish; g[3] = res.rho; g[4] = res.pri_res; g[5] = res.dua_res;
    g[6] = res.pol_pri; g[7] = res.pol_dua; g[8] = res.c; g[9] = res.pol_factor_ok; g[10] = res.rho_updates;
    g[11] = nr; g[12] = n_aux; g[13] = nnzA; g[14] = warm ? 1 : 0; g[15] = res.rounds;
  }
}

}  // namespace tb200
