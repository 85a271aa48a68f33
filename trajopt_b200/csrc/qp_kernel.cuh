// QP subproblem kernel: one warp per trajectory, time sliced.
//
// Replaces OSQPModel::optimize() -> osqp_setup/osqp_solve (trajopt_sco/src/osqp_interface.cpp:283-615) for
// every trajectory of the batch: the l1-penalty QP of optimizers.cpp:781-799 (Appendix B of SURVEY.md) is
// assembled from the fixed-layout convexification rows, equilibrated (Ruiz), and solved with the
// OSQP-equivalent ADMM (rho_eq = 1e3 rho, sigma, alpha relaxation, residual tests every 25 iterations,
// adaptive rho with refactorisation, polish).
//
// Linear algebra.  The KKT solve is done in its reduced form (P + sigma I + A' diag(rho) A) x = rhs after
// eliminating, row by row and in closed form, the hinge / abs auxiliary variables (each couples to exactly
// one row).  What is left is an N x N symmetric positive definite matrix, N = T*D, with half bandwidth
// nb = 2*D: block tridiagonal with nb x nb blocks.  It is factored in shared memory (band Cholesky), the
// diagonal blocks of the factor are inverted once per factorisation, and every triangular solve is then a
// chain of 2*M small triangular mat-vecs (M = N/nb) done with warp shuffles: two lanes per block row, the
// running block vector lives in registers.
//
// Work distribution inside the warp.  Row-local work (aux back-substitution, projection, dual update) runs
// one lane per row; accumulations into trajectory variables (A'v, the Hessian assembly, column norms) run one
// lane per variable over per-column entry lists, so no atomics are needed and every sum has a fixed order.
//
// Convergence discipline.  Under independent thread scheduling a warp that splits inside the iteration loop
// (e.g. a `for (i = lane; i < N; i += 32)` loop whose trip count differs between lanes) stays split across
// iterations, and every later shuffle / __syncwarp then takes the slow WARPSYNC.COLLECTIVE path (measured: 36k
// instructions per ADMM iteration instead of 3k).  Everything inside the solver loop is therefore written with
// warp-uniform trip counts and predicated bodies: rows are padded to CN coefficients and two aux slots, lane
// loops run ceil(n/32) times for every lane, column walks run to the longest column of their 32-variable chunk.
//
// Time slicing.  A launch advances each unfinished QP by at most `slice` ADMM iterations and parks its state
// in HBM; finished QPs raise qp_done and are consumed by eval_convexify_decide_kernel.  Trajectories
// therefore walk through their own SQP state machines asynchronously: a slow QP (up to max_iter = 8192
// iterations) no longer stalls the batch.
#pragma once
#include "device_types.cuh"
#include "eval_kernel.cuh"

namespace tb200 {

#ifdef TB200_PROFILE
__device__ unsigned long long g_prof[16];
#define PROF_T0() const long long prof_t0_ = clock64()
#define PROF_ADD(slot) do { if (q.lane == 0) atomicAdd(&g_prof[slot], (unsigned long long)(clock64() - prof_t0_)); } while (0)
#else
#define PROF_T0()
#define PROF_ADD(slot)
#endif

constexpr double kOsqpInf = 1e30;
constexpr double kMinScaling = 1e-4, kMaxScaling = 1e4;
constexpr double kVerifyTol = 1e-9;  // KKT verification of the polished point (deviation D2)
constexpr int kVerifyRounds = 3;
constexpr double kRhoMin = 1e-6, kRhoMax = 1e6, kRhoTol = 1e-4, kRhoEqOverIneq = 1e3;
enum { QPS_UNSOLVED = 0, QPS_SOLVED = 1, QPS_SOLVED_INACC = 2, QPS_PINF = 3, QPS_PINF_INACC = 4, QPS_DINF = 5,
       QPS_DINF_INACC = 6, QPS_MAXITER = 7, QPS_NONCVX = 8, QPS_YIELD = 100 };

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
__device__ __forceinline__ int warp_max_int(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ double limit_scaling(double v) {
  v = v < kMinScaling ? 1.0 : v;
  return v > kMaxScaling ? kMaxScaling : v;
}

// ---- shared memory layout (doubles) ------------------------------------------------------------------
struct QpSmem {
  int Kb, Linv, beta, x, zb, yb, v1, qs, lbs, ubs, xch, colptr, maxcol, total;
};
__host__ __device__ inline int qp_block_count(int N, int nb) { return (N + nb - 1) / nb; }
__host__ __device__ inline QpSmem qp_smem_layout(int N, int nb) {
  const int M = qp_block_count(N, nb), Np = M * nb, Wd = nb + 2;
  QpSmem s;
  int o = 0;
  s.Kb = o;   o += Np * Wd;
  s.Linv = o; o += M * (nb * (nb + 1) / 2);
  s.beta = o; o += Np;
  s.x = o;    o += Np;
  s.zb = o;   o += Np;
  s.yb = o;   o += Np;
  s.v1 = o;   o += Np;
  s.qs = o;   o += Np;
  s.lbs = o;  o += Np;
  s.ubs = o;  o += Np;
  s.xch = o;  o += 32;  // two 16-double exchange buffers of the block solve
  s.colptr = o; o += (Np + 2) / 2 + 1;
  s.maxcol = o; o += ((Np + 31) / 32 + 1) / 2 + 1;
  s.total = o;
  return s;
}

// ---- per-row record (global memory): CN raw coefficients, CN scaled coefficients, then these fields ----
// Every row is padded to CN coefficients (zeros) and two aux slots (absent aux: u = b = qa = 0, scalings 1).
enum RowF {
  R_C = 0, R_W,                         // raw: constant, aux cost
  R_E, R_DA0, R_DA1, R_EA0, R_EA1,      // Ruiz scalings
  R_U0, R_U1, R_B0, R_B1, R_LO, R_UP, R_QA0, R_QA1, R_RHO,  // scaled view (R_RHO: 1 = equality row)
  R_XA0, R_XA1, R_Z, R_Y, R_ZA0, R_ZA1, R_YA0, R_YA1,        // ADMM state
  R_RA0, R_RA1, R_COEF, R_WR, R_G0, R_G1, R_DEN, R_WRR,      // per-solve temporaries (WRR: raw row weight)
  R_IDEN, R_IWRR,                                            // reciprocals of DEN / WRR (0 when WRR == 0)
  R_DY, R_DYA0, R_DYA1, R_DXA0, R_DXA1,
  R_PW, R_PWA0, R_PWA1, R_PB, R_PY, R_PYA0, R_PYA1, R_PX0, R_PX1,  // polish
  R_MV,
  R_NF
};
__host__ __device__ inline int qp_row_stride(int CN) { return 2 * CN + R_NF; }

struct QpCtx {
  int N, Np, nb, M, Wd, T, D, CN, RS, lane, nrows;
  int npl, nrl;          // uniform trip counts of the per-variable / per-row lane loops
  double *Kb, *Linv, *beta, *x, *zb, *yb, *v1, *qs, *lbs, *ubs, *xch;   // shared
  int* colptr;           // shared [Np+1]
  int* maxcol;           // shared [npl]: longest column of each 32-variable chunk
  double *Dz, *v2;       // global [Np] (used by the residual / polish passes only)
  double* rows;          // global
  int* rints;
  const int* colent;     // global entries: (row << 5) | k
  const double* Pband;   // global [N][2D+1]
  double* scratch;
  double c, cinv, rho, rho_eq, sigma, alpha;
  __device__ __forceinline__ double* R(int r) const { return rows + static_cast<size_t>(r) * RS; }
  __device__ __forceinline__ double* F(int r) const { return rows + static_cast<size_t>(r) * RS + 2 * CN; }
  __device__ __forceinline__ const int* I(int r) const { return rints + static_cast<size_t>(r) * RI_NINTS; }
};

// weights of the linear system: ADMM (rho vector, sigma) or polish (1/delta on the active set, delta)
struct SysW {
  bool polish;
  double sig, rho_aux;
};

// ---------------------------------------------------------------------------------------------------
// Band Cholesky in place (lower band, Kb[i*Wd + k] = K(i, i-k), Wd = nb+2 keeps the column walks conflict
// free) followed by the inversion of the nb x nb diagonal blocks of the factor.
__device__ inline bool band_factor(const QpCtx& q) {
  const int Np = q.Np, HB = q.nb, W = q.Wd, lane = q.lane;
  bool ok = true;
  const int full_pairs = HB * (HB + 1) / 2;
  const int pair_trips = (full_pairs + 31) / 32;
  for (int j = 0; j < Np; ++j) {
    const double djj = q.Kb[j * W];
    ok = ok && (djj > 0.0);
    const double d = sqrt(djj > 0.0 ? djj : 1.0);
    const double inv = 1.0 / d;
    __syncwarp();
    const int m = min(HB, Np - 1 - j);
    if (lane == 0) q.Kb[j * W] = d;
    if (lane >= 1 && lane <= m) q.Kb[(j + lane) * W + lane] *= inv;
    __syncwarp();
    // trailing update: K(j+a, j+b) -= L(j+a,j) L(j+b,j), 1 <= b <= a <= m
    const int npairs = m * (m + 1) / 2;
    for (int t = 0; t < pair_trips; ++t) {
      const int pidx = lane + 32 * t;
      if (pidx < npairs) {
        int a = static_cast<int>((sqrt(8.0 * pidx + 1.0) - 1.0) * 0.5) + 1;
        a -= (a * (a - 1) / 2 > pidx);
        a += ((a + 1) * a / 2 <= pidx);
        const int b = pidx - a * (a - 1) / 2 + 1;
        q.Kb[(j + a) * W + (a - b)] -= q.Kb[(j + a) * W + a] * q.Kb[(j + b) * W + b];
      }
    }
    __syncwarp();
  }
  // Linv_i = inv(L_ii), packed lower triangle, one (block, column) task per lane, uniform trip counts
  const int nb = q.nb, tri = nb * (nb + 1) / 2;
  const int ntask = q.M * nb, task_trips = (ntask + 31) / 32;
  for (int t = 0; t < task_trips; ++t) {
    const int task_raw = lane + 32 * t;
    const bool act = task_raw < ntask;
    const int task = act ? task_raw : 0;
    const int blk = task / nb, c = task % nb;
    double* Li = q.Linv + blk * tri;
    const double* Lb = q.Kb + static_cast<size_t>(blk) * nb * W;  // row r of the block: Lb[r*W + (r - col)]
    for (int r = 0; r < nb; ++r) {  // solve L z = e_c by forward substitution; z_r for r >= c
      double s = (r == c) ? 1.0 : 0.0;
      for (int k = 0; k < nb; ++k) {
        const bool on = k >= c && k < r;
        const int kc = on ? k : c;
        const double lv = Lb[r * W + (on ? r - k : 0)], zv = Li[kc * (kc + 1) / 2 + c];
        s -= on ? lv * zv : 0.0;
      }
      if (act && r >= c) Li[r * (r + 1) / 2 + c] = s / Lb[r * W];
    }
  }
  __syncwarp();
  return ok;
}

// Solves K v = v in place (v in shared memory, length Np) with the block factor.
__device__ inline void block_solve_generic(const QpCtx& q, double* v) {
  const int nb = q.nb, W = q.Wd, M = q.M, tri = nb * (nb + 1) / 2;
  const int halves = (nb <= 16) ? 2 : 1;
  const int r = (halves == 2) ? (q.lane & 15) : q.lane;      // block row handled by this lane
  const int h = (halves == 2) ? (q.lane >> 4) : 0;           // which half of the columns
  const int chunk = (nb + halves - 1) / halves;
  const int c0 = h * chunk;
  const bool rowok = r < nb;
  const int rr = rowok ? r : 0;
  __syncwarp();
  // ---- forward: y_i = Linv_i (b_i - W_i y_{i-1}),  W_i[r][c] = Kb[(i nb + r) W + nb + r - c], c >= r
  double yprev = 0.0;
  for (int i = 0; i < M; ++i) {
    double t = rowok ? v[i * nb + rr] : 0.0;
    if (i > 0) {
      double acc = 0.0;
      const double* Wrow = q.Kb + static_cast<size_t>(i * nb + rr) * W + nb + rr;
      for (int j = 0; j < chunk; ++j) {
        const int c = c0 + j;
        const double yv = __shfl_sync(0xffffffffu, yprev, c & 31);
        const bool on = rowok && c < nb && c >= rr;
        const double wv = Wrow[-(on ? c : rr)];
        acc += (on ? wv : 0.0) * yv;
      }
      if (halves == 2) acc += __shfl_xor_sync(0xffffffffu, acc, 16);
      t -= acc;
    }
    double acc = 0.0;
    const double* Lrow = q.Linv + i * tri + rr * (rr + 1) / 2;
    for (int j = 0; j < chunk; ++j) {
      const int c = c0 + j;
      const double tv = __shfl_sync(0xffffffffu, t, c & 31);
      const bool on = rowok && c <= rr;
      const double lv = Lrow[on ? c : 0];
      acc += (on ? lv : 0.0) * tv;
    }
    if (halves == 2) acc += __shfl_xor_sync(0xffffffffu, acc, 16);
    yprev = acc;
    if (rowok && h == 0) v[i * nb + rr] = acc;
  }
  // ---- backward: x_i = Linv_i' (y_i - W_{i+1}' x_{i+1})
  double xnext = 0.0;
  for (int i = M - 1; i >= 0; --i) {
    double t = rowok ? ((i == M - 1) ? yprev : v[i * nb + rr]) : 0.0;
    if (i < M - 1) {
      double acc = 0.0;
      for (int j = 0; j < chunk; ++j) {
        const int c = c0 + j;
        const double xv = __shfl_sync(0xffffffffu, xnext, c & 31);
        // W_{i+1}[c][r] = Kb[((i+1) nb + c) W + nb + c - r], nonzero iff r >= c
        const bool on = rowok && c <= rr;
        const int cc = on ? c : rr;
        const double wv = q.Kb[static_cast<size_t>((i + 1) * nb + cc) * W + nb + cc - rr];
        acc += (on ? wv : 0.0) * xv;
      }
      if (halves == 2) acc += __shfl_xor_sync(0xffffffffu, acc, 16);
      t -= acc;
    }
    double acc = 0.0;
    for (int j = 0; j < chunk; ++j) {
      const int c = c0 + j;
      const double tv = __shfl_sync(0xffffffffu, t, c & 31);
      const bool on = rowok && c < nb && c >= rr;
      const int cc = on ? c : rr;
      const double lv = q.Linv[i * tri + cc * (cc + 1) / 2 + rr];
      acc += (on ? lv : 0.0) * tv;
    }
    if (halves == 2) acc += __shfl_xor_sync(0xffffffffu, acc, 16);
    xnext = acc;
    if (rowok && h == 0) v[i * nb + rr] = acc;
  }
  __syncwarp();
}

// Fast path of the block solve for nb = NB <= 16 (two lanes per block row, CH columns each), fully unrolled.
// The block vectors travel through two small shared exchange buffers (half 0 at [0,CH), half 1 at [8,8+CH)).
template <int NB>
__device__ __forceinline__ void block_solve_t(const QpCtx& q, double* v) {
  constexpr int CH = (NB + 1) / 2;
  constexpr int TRI = NB * (NB + 1) / 2;
  const int W = q.Wd, M = q.M;
  const int r = q.lane & 15, h = q.lane >> 4;
  const bool rowok = r < NB;
  const int ra = rowok ? r : 0;             // row used for addressing
  const int rge = rowok ? r : 1 << 20;      // "c >= r" never true for idle lanes
  const int rle = rowok ? r : -1;           // "c <= r" never true for idle lanes
  const int c0 = h * CH;
  const bool writer = rowok && h == 0;
  const int xi = (ra < CH) ? ra : 8 + ra - CH;   // slot of element `ra` in an exchange buffer
  double* xa = q.xch;
  double* xb = q.xch + 16;
  const double* seg_a = xa + h * 8;
  const double* seg_b = xb + h * 8;
  int ltri[CH];                              // packed-triangle offsets of Linv[c][ra], c = c0 + j
#pragma unroll
  for (int j = 0; j < CH; ++j) ltri[j] = (c0 + j) * (c0 + j + 1) / 2 + ra;
  __syncwarp();
  // ---- forward: y_i = Linv_i (b_i - W_i y_{i-1})
  double yseg[CH];
#pragma unroll
  for (int j = 0; j < CH; ++j) yseg[j] = 0.0;
  const double* Wrow = q.Kb + ra * W + NB + ra - c0;        // W_i[ra][c0+j] = Wrow[i*NB*W - j]
  const double* Lrow = q.Linv + ra * (ra + 1) / 2 + c0;     // Linv_i[ra][c0+j] = Lrow[i*TRI + j]
  for (int i = 0; i < M; ++i) {
    double t = v[i * NB + ra];
    if (i > 0) {
      const double* wp = Wrow + i * NB * W;
      double acc = 0.0;
#pragma unroll
      for (int j = 0; j < CH; ++j)
        if (c0 + j >= rge && c0 + j < NB) acc = fma(wp[-j], yseg[j], acc);
      acc += __shfl_xor_sync(0xffffffffu, acc, 16);
      t -= acc;
    }
    if (writer) xa[xi] = t;
    __syncwarp();
    const double* lp = Lrow + i * TRI;
    double acc = 0.0;
#pragma unroll
    for (int j = 0; j < CH; ++j)
      if (c0 + j <= rle) acc = fma(lp[j], seg_a[j], acc);
    acc += __shfl_xor_sync(0xffffffffu, acc, 16);
    if (writer) {
      v[i * NB + ra] = acc;
      xb[xi] = acc;
    }
    __syncwarp();
#pragma unroll
    for (int j = 0; j < CH; ++j) yseg[j] = seg_b[j];
  }
  // ---- backward: x_i = Linv_i' (y_i - W_{i+1}' x_{i+1});  yseg now holds y_{M-1}, reused as x_{i+1} below
  const double* Wcol = q.Kb + c0 * (W + 1) + NB - ra;        // W_{i+1}[c0+j][ra] = Wcol[(i+1)*NB*W + j*(W+1)]
  for (int i = M - 1; i >= 0; --i) {
    double t = v[i * NB + ra];
    if (i < M - 1) {
      const double* wp = Wcol + (i + 1) * NB * W;
      double acc = 0.0;
#pragma unroll
      for (int j = 0; j < CH; ++j)
        if (c0 + j <= rle) acc = fma(wp[j * (W + 1)], yseg[j], acc);
      acc += __shfl_xor_sync(0xffffffffu, acc, 16);
      t -= acc;
    }
    if (writer) xa[xi] = t;
    __syncwarp();
    const double* lp = q.Linv + i * TRI;
    double acc = 0.0;
#pragma unroll
    for (int j = 0; j < CH; ++j)
      if (c0 + j >= rge && c0 + j < NB) acc = fma(lp[ltri[j]], seg_a[j], acc);
    acc += __shfl_xor_sync(0xffffffffu, acc, 16);
    if (writer) {
      v[i * NB + ra] = acc;
      xb[xi] = acc;
    }
    __syncwarp();
#pragma unroll
    for (int j = 0; j < CH; ++j) yseg[j] = seg_b[j];
  }
}

__device__ __forceinline__ void block_solve(const QpCtx& q, double* v) {
  if (q.nb == 14) block_solve_t<14>(q, v);
  else block_solve_generic(q, v);
}

// variable index of coefficient k of a row (padding coefficients alias the last real one; their value is 0)
__device__ __forceinline__ int row_var(const int* I, int k) {
  return I[RI_BASE] + min(k, I[RI_CNT] - 1) * I[RI_STRIDE];
}

// scaled P (band) times a vector: out = c * Dz .* (P (Dz .* in)); uniform trip counts
__device__ inline void p_matvec(const QpCtx& q, const double* in, double* out) {
  const int N = q.N, HB = 2 * q.D, W = HB + 1;
  for (int kk = 0; kk < q.npl; ++kk) {
    const int iraw = q.lane + 32 * kk;
    const bool act = iraw < N;
    const int i = act ? iraw : 0;
    double s = 0.0;
    for (int k = 0; k <= HB; ++k) {
      const bool on = k <= i;
      const int j = on ? i - k : 0;
      s += (on ? q.Pband[i * W + k] : 0.0) * q.Dz[j] * in[j];
    }
    for (int k = 1; k <= HB; ++k) {
      const bool on = i + k < N;
      const int j = on ? i + k : 0;
      s += (on ? q.Pband[j * W + k] : 0.0) * q.Dz[j] * in[j];
    }
    s *= q.c * q.Dz[i];
    if (iraw < q.Np) out[iraw] = act ? s : 0.0;
  }
  __syncwarp();
}

// per-row weights of the current linear system -> R_WRR (raw row weight), R_G0/G1, R_DEN, R_WR (Schur weight).
// With the aux block K_aa = diag(g) + Wr u u' everything is written cancellation free (den = det K_aa
// expanded analytically); the polish system has Wr = 1/delta and g = delta.
__device__ inline void rows_prepare_weights(const QpCtx& q, const SysW& w) {
  for (int kk = 0; kk < q.nrl; ++kk) {
    const int rraw = q.lane + 32 * kk;
    const bool act = rraw < q.nrows;
    double* F = q.F(act ? rraw : 0);
    const int naux = q.I(act ? rraw : 0)[RI_AUX];
    const double Wr = w.polish ? fabs(F[R_PW]) : ((F[R_RHO] != 0.0) ? q.rho_eq : q.rho);
    const double wa0 = w.polish ? fabs(F[R_PWA0]) : w.rho_aux;
    const double wa1 = w.polish ? fabs(F[R_PWA1]) : w.rho_aux;
    const double g0 = (naux >= 1) ? w.sig + wa0 * F[R_B0] * F[R_B0] : 1.0;
    const double g1 = (naux == 2) ? w.sig + wa1 * F[R_B1] * F[R_B1] : 1.0;
    const double den = g0 * g1 + Wr * (F[R_U0] * F[R_U0] * g1 + F[R_U1] * F[R_U1] * g0);
    if (act) {
      F[R_WRR] = Wr;
      F[R_G0] = g0;
      F[R_G1] = g1;
      F[R_DEN] = den;
      F[R_IDEN] = 1.0 / den;
      F[R_IWRR] = (Wr != 0.0) ? 1.0 / Wr : 0.0;
      F[R_WR] = Wr * g0 * g1 / den;
    }
  }
  __syncwarp();
}
__device__ __forceinline__ double xbound_weight(const QpCtx& q, const SysW& w, int j) {
  const double adm = (q.ubs[j] - q.lbs[j] < kRhoTol) ? q.rho_eq : q.rho;
  return w.polish ? fabs(q.zb[j]) : adm;  // zb holds the signed polish weights during polish
}

// K = P + sig I + A' W A with the aux variables eliminated (one lane per matrix row); then factor.
__device__ inline bool assemble_factor(const QpCtx& q, const SysW& w) {
  rows_prepare_weights(q, w);
  const int N = q.N, HB = 2 * q.D, PW = HB + 1, Wd = q.Wd, CN = q.CN;
  for (int kk = 0; kk < q.npl; ++kk) {
    const int iraw = q.lane + 32 * kk;
    const bool inP = iraw < q.Np, act = iraw < N;
    const int i = act ? iraw : 0;
    double* Ki = q.Kb + static_cast<size_t>(inP ? iraw : 0) * Wd;
    if (inP)
      for (int k = 0; k < Wd; ++k) Ki[k] = 0.0;
    if (inP && !act) Ki[0] = 1.0;  // padding variable
    if (act) {
      for (int k = 0; k <= HB; ++k)
        if (k <= i) Ki[k] = q.c * q.Dz[i] * q.Pband[i * PW + k] * q.Dz[i - k];
      Ki[0] += w.sig + xbound_weight(q, w, i) * q.beta[i] * q.beta[i];
    }
    const int e0 = act ? q.colptr[i] : 0, e1 = act ? q.colptr[i + 1] : 0;
    const int trips = q.maxcol[kk];
    for (int t = 0; t < trips; ++t) {
      const bool on = e0 + t < e1;
      const int ent = q.colent[on ? e0 + t : 0], r = ent >> 5, k = ent & 31;
      const double* R = q.R(r);
      const double wr = on ? R[2 * CN + R_WR] : 0.0;
      const int stride = q.I(r)[RI_STRIDE];
      const double* as = R + CN;
      const double ai = wr * as[k];
      for (int k2 = 0; k2 < CN; ++k2)
        if (on && k2 <= k) Ki[(k - k2) * stride] += ai * as[k2];
    }
  }
  __syncwarp();
  return band_factor(q);
}

struct QpOut {
  int status, iters, polish;
  double rho;
  double pri_res, dua_res, pol_pri, pol_dua, c;
  int pol_factor_ok, rho_updates, rounds;
};

// Persistent solver state of one QP between time slices.
struct QpResume {
  int iter, round, rho_updates, status;
  double rho, eps_scale, c;
};

// Scatter pass: v1[i] = base(i) + sum over the column entries of as[k] * R_COEF(row); uniform trip counts.
template <class Base>
__device__ __forceinline__ void scatter_columns(const QpCtx& q, Base base) {
  const int CN = q.CN;
  for (int kk = 0; kk < q.npl; ++kk) {
    const int iraw = q.lane + 32 * kk;
    const bool act = iraw < q.N;
    const int i = act ? iraw : 0;
    double s = base(i);
    const int e0 = act ? q.colptr[i] : 0, e1 = act ? q.colptr[i + 1] : 0;
    const int trips = q.maxcol[kk];
    for (int t = 0; t < trips; ++t) {
      const bool on = e0 + t < e1;
      const int ent = q.colent[on ? e0 + t : 0], r = ent >> 5, k = ent & 31;
      const double* R = q.R(r);
      const double a = R[CN + k], cf = R[2 * CN + R_COEF];
      s += on ? a * cf : 0.0;
    }
    if (iraw < q.Np) q.v1[iraw] = act ? s : 0.0;
  }
  __syncwarp();
}
// zeta_r = as . v(vars of the row); all CN (zero padded) coefficients
__device__ __forceinline__ double row_dot(const QpCtx& q, const double* R, const int* I, const double* v) {
  double z = 0.0;
  const int base = I[RI_BASE], stride = I[RI_STRIDE], last = I[RI_CNT] - 1;
  for (int k = 0; k < q.CN; ++k) z += R[q.CN + k] * v[base + min(k, last) * stride];
  return z;
}
// aux back-substitution (cancellation free; absent aux slots have u = ra = 0 and g = 1 and come out 0)
__device__ __forceinline__ void row_backsub(const double* F, double zeta, double& a0, double& a1) {
  const double Wr = F[R_WRR], ra0 = F[R_RA0], ra1 = F[R_RA1], u0 = F[R_U0], u1 = F[R_U1];
  const double iden = F[R_IDEN];
  a0 = (F[R_G1] * (ra0 - Wr * u0 * zeta) + Wr * u1 * (u1 * ra0 - u0 * ra1)) * iden;
  a1 = (F[R_G0] * (ra1 - Wr * u1 * zeta) + Wr * u0 * (u0 * ra1 - u1 * ra0)) * iden;
}
__device__ __forceinline__ double row_reduce_coef(const double* F, double ra0, double ra1, double zcoef) {
  return zcoef - F[R_WRR] * (F[R_U0] * ra0 * F[R_G1] + F[R_U1] * ra1 * F[R_G0]) * F[R_IDEN];
}

// Ruiz equilibration (scale_data of OSQP [EXT]); leaves the scaled view of every row in its record and the
// scaled trajectory cost / bounds in q.qs / q.lbs / q.ubs, Dz (global) / beta (shared).  Runs once per QP,
// outside the iteration loop.
__device__ inline void qp_scale(QpCtx& q, const QpSettings& st, int n_aux_total) {
  double *qs = q.qs, *lbs = q.lbs, *ubs = q.ubs;
  const int N = q.N, lane = q.lane, HB = 2 * q.D, W = HB + 1;
  double* Eb = q.zb;  // bound-row scalings live in zb during scaling
  q.c = 1.0;
  for (int i = lane; i < q.Np; i += 32) {
    q.Dz[i] = 1.0;
    Eb[i] = 1.0;
  }
  for (int r = lane; r < q.nrows; r += 32) {
    double* F = q.F(r);
    F[R_E] = F[R_DA0] = F[R_DA1] = F[R_EA0] = F[R_EA1] = 1.0;
  }
  __syncwarp();
  for (int pass = 0; pass < st.scaling; ++pass) {
    // row norms (one lane per row) -> E_temp in R_RA0; aux column / bound-row scalings updated in place
    for (int r = lane; r < q.nrows; r += 32) {
      double* R = q.R(r);
      double* F = q.F(r);
      const int* I = q.I(r);
      const int base = I[RI_BASE], cnt = I[RI_CNT], stride = I[RI_STRIDE], aux = I[RI_AUX];
      const double E = F[R_E];
      double rn = 0.0;
      for (int k = 0; k < cnt; ++k) rn = fmax(rn, fabs(E * R[k] * q.Dz[base + k * stride]));
      double dt0 = 1.0, dt1 = 1.0, et0 = 1.0, et1 = 1.0;
      if (aux >= 1) {
        const double ua = fabs(E * F[R_DA0]), ba = fabs(F[R_EA0] * F[R_DA0]);
        rn = fmax(rn, ua);
        dt0 = 1.0 / sqrt(limit_scaling(fmax(ua, ba)));
        et0 = 1.0 / sqrt(limit_scaling(ba));
      }
      if (aux == 2) {
        const double ua = fabs(E * F[R_DA1]), ba = fabs(F[R_EA1] * F[R_DA1]);
        rn = fmax(rn, ua);
        dt1 = 1.0 / sqrt(limit_scaling(fmax(ua, ba)));
        et1 = 1.0 / sqrt(limit_scaling(ba));
      }
      F[R_RA0] = 1.0 / sqrt(limit_scaling(rn));
      F[R_RA1] = E;  // E before this pass (the column pass below must still see the old value)
      F[R_DA0] *= dt0;
      F[R_DA1] *= dt1;
      F[R_EA0] *= et0;
      F[R_EA1] *= et1;
    }
    __syncwarp();
    // column norms of [P A'; A 0] restricted to the trajectory variables (one lane per variable)
    for (int i = lane; i < N; i += 32) {
      double m = 0.0;
      for (int k = 0; k <= HB && k <= i; ++k) m = fmax(m, fabs(q.c * q.Dz[i] * q.Pband[i * W + k] * q.Dz[i - k]));
      for (int k = 1; k <= HB && i + k < N; ++k) m = fmax(m, fabs(q.c * q.Dz[i] * q.Pband[(i + k) * W + k] * q.Dz[i + k]));
      const double bn = fabs(Eb[i] * q.Dz[i]);
      m = fmax(m, bn);
      for (int e = q.colptr[i]; e < q.colptr[i + 1]; ++e) {
        const int ent = q.colent[e], r = ent >> 5, k = ent & 31;
        const double* R = q.R(r);
        m = fmax(m, fabs(R[2 * q.CN + R_RA1] * R[k] * q.Dz[i]));
      }
      q.v1[i] = 1.0 / sqrt(limit_scaling(m));
      Eb[i] *= 1.0 / sqrt(limit_scaling(bn));
    }
    __syncwarp();
    for (int i = lane; i < N; i += 32) q.Dz[i] *= q.v1[i];
    for (int r = lane; r < q.nrows; r += 32) {
      double* F = q.F(r);
      F[R_E] = F[R_RA1] * F[R_RA0];
    }
    __syncwarp();
    // cost normalisation: mean column inf-norm of the scaled P over ALL n variables (aux columns are 0)
    double csum = 0.0, qn = 0.0;
    for (int i = lane; i < N; i += 32) {
      double m = 0.0;
      for (int k = 0; k <= HB && k <= i; ++k) m = fmax(m, fabs(q.c * q.Dz[i] * q.Pband[i * W + k] * q.Dz[i - k]));
      for (int k = 1; k <= HB && i + k < N; ++k) m = fmax(m, fabs(q.c * q.Dz[i] * q.Pband[(i + k) * W + k] * q.Dz[i + k]));
      csum += m;
      qn = fmax(qn, fabs(q.c * q.Dz[i] * qs[i]));
    }
    for (int r = lane; r < q.nrows; r += 32) {
      const double* F = q.F(r);
      const int aux = q.I(r)[RI_AUX];
      if (aux >= 1) qn = fmax(qn, fabs(q.c * F[R_DA0] * F[R_W]));
      if (aux == 2) qn = fmax(qn, fabs(q.c * F[R_DA1] * F[R_W]));
    }
    __syncwarp();
    csum = warp_sum(csum);
    qn = warp_max(qn);
    const double mean = limit_scaling(csum / static_cast<double>(N + n_aux_total));
    q.c *= 1.0 / fmax(mean, limit_scaling(qn));
  }
  q.cinv = 1.0 / q.c;
  for (int i = lane; i < q.Np; i += 32) {
    if (i < N) {
      qs[i] = q.c * q.Dz[i] * qs[i];
      lbs[i] *= Eb[i];
      ubs[i] *= Eb[i];
      q.beta[i] = Eb[i] * q.Dz[i];
    } else {
      qs[i] = 0.0; lbs[i] = -1.0; ubs[i] = 1.0; q.beta[i] = 1.0; q.Dz[i] = 1.0;
    }
  }
  // scaled view of every row (padding coefficients stay exactly 0)
  for (int r = lane; r < q.nrows; r += 32) {
    double* R = q.R(r);
    double* F = q.F(r);
    const int* I = q.I(r);
    const int aux = I[RI_AUX];
    const double E = F[R_E];
    for (int k = 0; k < q.CN; ++k) R[q.CN + k] = E * R[k] * q.Dz[row_var(I, k)];
    F[R_U0] = F[R_U1] = F[R_B0] = F[R_B1] = F[R_QA0] = F[R_QA1] = 0.0;
    F[R_UP] = -F[R_C] * E;
    if (aux == AUX_HINGE) {
      F[R_U0] = -E * F[R_DA0];
      F[R_B0] = F[R_EA0] * F[R_DA0];
      F[R_QA0] = q.c * F[R_DA0] * F[R_W];
      F[R_LO] = -kOsqpInf * E;
      F[R_RHO] = 0.0;
    } else {
      if (aux == AUX_ABS) {
        F[R_U0] = E * F[R_DA0];
        F[R_U1] = -E * F[R_DA1];
        F[R_B0] = F[R_EA0] * F[R_DA0];
        F[R_B1] = F[R_EA1] * F[R_DA1];
        F[R_QA0] = q.c * F[R_DA0] * F[R_W];
        F[R_QA1] = q.c * F[R_DA1] * F[R_W];
      }
      F[R_LO] = F[R_UP];
      F[R_RHO] = 1.0;
    }
  }
  __syncwarp();
}

// The QP solve for the calling warp's trajectory.  `fresh`: start a new solve (initial iterate from the warm
// start or zero); otherwise resume from `rs`.  Returns status QPS_YIELD when the slice budget ran out.
__device__ inline QpOut qp_solve_warp(QpCtx& q, const QpSettings& st, bool fresh, bool warm, double warm_rho,
                                      const double* ws_x, const double* ws_yb, QpResume& rs, int slice) {
  const int N = q.N, lane = q.lane;
  QpOut out{QPS_UNSOLVED, 0, 0, st.rho, 0, 0, 0, 0, 0, -1, 0, 0};
  double rho;
  double eps_scale;
  int iter, round;
  q.sigma = st.sigma;
  q.alpha = st.alpha;
  if (fresh) {
    rho = warm ? warm_rho : st.rho;
    rho = fmin(fmax(rho, kRhoMin), kRhoMax);
    eps_scale = 1.0;
    iter = 0;
    round = 0;
    q.rho = rho;
    q.rho_eq = kRhoEqOverIneq * rho;
    if (warm) {  // osqp_warm_start: x <- Dinv x, y <- c Einv y, z <- A x
      for (int i = lane; i < q.Np; i += 32) {
        if (i < N) {
          q.x[i] = ws_x[i] / q.Dz[i];
          q.yb[i] = ws_yb[i] * q.Dz[i] / q.beta[i] * q.c;   // Eb = beta / Dz
          q.zb[i] = q.beta[i] * q.x[i];
        } else {
          q.x[i] = q.yb[i] = q.zb[i] = 0.0;
        }
      }
      __syncwarp();
      for (int r = lane; r < q.nrows; r += 32) {
        double* R = q.R(r);
        double* F = q.F(r);
        const int naux = q.I(r)[RI_AUX];
        F[R_XA0] = (naux >= 1) ? F[R_XA0] / F[R_DA0] : 0.0;
        F[R_XA1] = (naux == 2) ? F[R_XA1] / F[R_DA1] : 0.0;
        F[R_Y] = F[R_Y] / F[R_E] * q.c;
        F[R_YA0] = (naux >= 1) ? F[R_YA0] / F[R_EA0] * q.c : 0.0;
        F[R_YA1] = (naux == 2) ? F[R_YA1] / F[R_EA1] * q.c : 0.0;
        F[R_Z] = row_dot(q, R, q.I(r), q.x) + F[R_U0] * F[R_XA0] + F[R_U1] * F[R_XA1];
        F[R_ZA0] = F[R_B0] * F[R_XA0];
        F[R_ZA1] = F[R_B1] * F[R_XA1];
      }
    } else {
      for (int i = lane; i < q.Np; i += 32) q.x[i] = q.zb[i] = q.yb[i] = 0.0;
      for (int r = lane; r < q.nrows; r += 32) {
        double* F = q.F(r);
        F[R_XA0] = F[R_XA1] = F[R_Z] = F[R_Y] = F[R_ZA0] = F[R_ZA1] = F[R_YA0] = F[R_YA1] = 0.0;
      }
    }
    __syncwarp();
  } else {
    rho = rs.rho;
    eps_scale = rs.eps_scale;
    iter = rs.iter;
    round = rs.round;
    out.rho_updates = rs.rho_updates;
    q.rho = rho;
    q.rho_eq = kRhoEqOverIneq * rho;
  }
  SysW sysw{false, st.sigma, rho};
  bool factor_ok;
  { PROF_T0(); factor_ok = assemble_factor(q, sysw); PROF_ADD(6); }

  double* dxs = q.scratch;              // [Np] last trajectory step (written on check iterations)
  double* dyb = q.scratch + q.Np;       // [Np] last dual step of the variable-bound rows
  double* st_x = q.scratch + 2 * q.Np;  // ADMM x, zb, yb stashed while polish reuses the shared vectors
  double* st_zb = q.scratch + 3 * q.Np;
  double* st_yb = q.scratch + 4 * q.Np;
  double pri_res = 0.0, dua_res = 0.0;
  int status = factor_ok ? QPS_UNSOLVED : QPS_NONCVX, budget = slice;
  double n_z = 0, n_ax = 0, n_q = 0, n_aty = 0, n_px = 0, s_pri = 0, s_dua = 0, s_z = 0, s_ax = 0, s_q = 0, s_aty = 0, s_px = 0;

  // ---------------------------------------------------------------- update_info(): residuals and norms
  auto info_pass = [&]() {
    p_matvec(q, q.x, q.v2);  // v2 <- P x
    double m_pri = 0, m_z = 0, m_ax = 0, m_dua = 0, m_aty = 0, m_q = 0, m_px = 0;
    double ms_pri = 0, ms_z = 0, ms_ax = 0, ms_dua = 0, ms_aty = 0, ms_q = 0, ms_px = 0;
    for (int kk = 0; kk < q.nrl; ++kk) {
      const int rraw = lane + 32 * kk;
      const bool act = rraw < q.nrows;
      const int r = act ? rraw : 0;
      const double* R = q.R(r);
      double* F = q.F(r);
      const double on = act ? 1.0 : 0.0;
      const double ax = row_dot(q, R, q.I(r), q.x) + F[R_U0] * F[R_XA0] + F[R_U1] * F[R_XA1];
      const double einv = 1.0 / F[R_E];
      m_pri = fmax(m_pri, on * fabs(einv * (ax - F[R_Z])));
      m_z = fmax(m_z, on * fabs(einv * F[R_Z]));
      m_ax = fmax(m_ax, on * fabs(einv * ax));
      ms_pri = fmax(ms_pri, on * fabs(ax - F[R_Z]));
      ms_z = fmax(ms_z, on * fabs(F[R_Z]));
      ms_ax = fmax(ms_ax, on * fabs(ax));
#pragma unroll
      for (int k = 0; k < 2; ++k) {  // absent aux slots contribute exact zeros
        const double u = F[R_U0 + k], bb = F[R_B0 + k], qa = F[R_QA0 + k];
        const double xa = F[R_XA0 + k], za = F[R_ZA0 + k], ya = F[R_YA0 + k];
        const double da = F[R_DA0 + k], ea = F[R_EA0 + k];
        const double axb = bb * xa;
        m_pri = fmax(m_pri, on * fabs((axb - za) / ea));
        m_z = fmax(m_z, on * fabs(za / ea));
        m_ax = fmax(m_ax, on * fabs(axb / ea));
        ms_pri = fmax(ms_pri, on * fabs(axb - za));
        ms_z = fmax(ms_z, on * fabs(za));
        ms_ax = fmax(ms_ax, on * fabs(axb));
        const double aty = u * F[R_Y] + bb * ya;
        m_dua = fmax(m_dua, on * fabs((qa + aty) / da));
        m_aty = fmax(m_aty, on * fabs(aty / da));
        m_q = fmax(m_q, on * fabs(qa / da));
        ms_dua = fmax(ms_dua, on * fabs(qa + aty));
        ms_aty = fmax(ms_aty, on * fabs(aty));
        ms_q = fmax(ms_q, on * fabs(qa));
      }
      if (act) F[R_COEF] = F[R_Y];
    }
    __syncwarp();
    scatter_columns(q, [&](int i) { return q.beta[i] * q.yb[i]; });  // v1 <- A'y (trajectory part)
    for (int kk = 0; kk < q.npl; ++kk) {
      const int iraw = lane + 32 * kk;
      const bool act = iraw < N;
      const int i = act ? iraw : 0;
      const double on = act ? 1.0 : 0.0;
      const double dz = q.Dz[i], beta = q.beta[i];
      const double ax = beta * q.x[i], aty = q.v1[i], px = q.v2[i];
      const double einv = dz / beta, dinv = 1.0 / dz;
      m_pri = fmax(m_pri, on * fabs(einv * (ax - q.zb[i])));
      m_z = fmax(m_z, on * fabs(einv * q.zb[i]));
      m_ax = fmax(m_ax, on * fabs(einv * ax));
      ms_pri = fmax(ms_pri, on * fabs(ax - q.zb[i]));
      ms_z = fmax(ms_z, on * fabs(q.zb[i]));
      ms_ax = fmax(ms_ax, on * fabs(ax));
      const double qv = q.qs[i], d = qv + px + aty;
      m_dua = fmax(m_dua, on * fabs(dinv * d));
      m_aty = fmax(m_aty, on * fabs(dinv * aty));
      m_q = fmax(m_q, on * fabs(dinv * qv));
      m_px = fmax(m_px, on * fabs(dinv * px));
      ms_dua = fmax(ms_dua, on * fabs(d));
      ms_aty = fmax(ms_aty, on * fabs(aty));
      ms_q = fmax(ms_q, on * fabs(qv));
      ms_px = fmax(ms_px, on * fabs(px));
    }
    __syncwarp();
    pri_res = warp_max(m_pri);
    dua_res = warp_max(m_dua) * q.cinv;
    n_z = warp_max(m_z); n_ax = warp_max(m_ax); n_q = warp_max(m_q); n_aty = warp_max(m_aty); n_px = warp_max(m_px);
    s_pri = warp_max(ms_pri); s_dua = warp_max(ms_dua); s_z = warp_max(ms_z); s_ax = warp_max(ms_ax);
    s_q = warp_max(ms_q); s_aty = warp_max(ms_aty); s_px = warp_max(ms_px);
  };

  auto primal_infeasible = [&](double eps) -> bool {  // is_primal_infeasible [EXT]
    double nd = 0.0, lhs = 0.0, na = 0.0;
    for (int kk = 0; kk < q.nrl; ++kk) {
      const int rraw = lane + 32 * kk;
      const bool act = rraw < q.nrows;
      const int r = act ? rraw : 0;
      double* F = q.F(r);
      const double on = act ? 1.0 : 0.0;
      const int naux = q.I(r)[RI_AUX];
      double d = F[R_DY];
      d = (naux == AUX_HINGE) ? fmax(d, 0.0) : d;  // l = -inf
      nd = fmax(nd, on * fabs(F[R_E] * d));
      lhs += on * (F[R_UP] * fmax(d, 0.0) + F[R_LO] * fmin(d, 0.0));
      if (act) F[R_COEF] = d;  // projected dual step, consumed by the column pass
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const double da = fmin(F[R_DYA0 + k], 0.0);  // aux bound rows: u = +inf, l = 0
        nd = fmax(nd, on * fabs(F[R_EA0 + k] * da));
        na = fmax(na, on * fabs((F[R_U0 + k] * d + F[R_B0 + k] * da) / F[R_DA0 + k]));
      }
    }
    for (int kk = 0; kk < q.npl; ++kk) {  // variable-bound rows: both bounds finite
      const int iraw = lane + 32 * kk;
      const bool act = iraw < N;
      const int i = act ? iraw : 0;
      const double on = act ? 1.0 : 0.0;
      const double d = dyb[i];
      nd = fmax(nd, on * fabs(q.beta[i] / q.Dz[i] * d));
      lhs += on * (q.ubs[i] * fmax(d, 0.0) + q.lbs[i] * fmin(d, 0.0));
    }
    __syncwarp();
    nd = warp_max(nd);
    lhs = warp_sum(lhs);
    na = warp_max(na);
    scatter_columns(q, [&](int i) { return q.beta[i] * dyb[i]; });
    double m = na;
    for (int kk = 0; kk < q.npl; ++kk) {
      const int iraw = lane + 32 * kk;
      const int i = iraw < N ? iraw : 0;
      m = fmax(m, (iraw < N ? 1.0 : 0.0) * fabs(q.v1[i] / q.Dz[i]));
    }
    m = warp_max(m);
    return (nd > eps) && (lhs < -eps * nd) && (m < eps * nd);
  };
  auto dual_infeasible = [&](double eps) -> bool {  // is_dual_infeasible [EXT]
    double ndx = 0.0, qdx = 0.0;
    for (int kk = 0; kk < q.npl; ++kk) {
      const int iraw = lane + 32 * kk;
      const bool act = iraw < N;
      const int i = act ? iraw : 0;
      const double dxi = act ? dxs[i] : 0.0;
      ndx = fmax(ndx, fabs(q.Dz[i] * dxi));
      qdx += q.qs[i] * dxi;
      if (iraw < q.Np) q.v1[iraw] = dxi;
    }
    for (int kk = 0; kk < q.nrl; ++kk) {
      const int rraw = lane + 32 * kk;
      const bool act = rraw < q.nrows;
      const double* F = q.F(act ? rraw : 0);
      const double on = act ? 1.0 : 0.0;
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        ndx = fmax(ndx, on * fabs(F[R_DA0 + k] * F[R_DXA0 + k]));
        qdx += on * F[R_QA0 + k] * F[R_DXA0 + k];
      }
    }
    __syncwarp();
    ndx = warp_max(ndx);
    qdx = warp_sum(qdx);
    p_matvec(q, q.v1, q.v2);  // v2 <- P dx
    double m = 0.0;
    int bad = 0;
    for (int kk = 0; kk < q.npl; ++kk) {
      const int iraw = lane + 32 * kk;
      const bool act = iraw < N;
      const int i = act ? iraw : 0;
      m = fmax(m, (act ? 1.0 : 0.0) * fabs(q.v2[i] / q.Dz[i]));
      const double vv = q.Dz[i] * q.v1[i];  // Einv * (Eb Dz dx); both bounds finite
      bad |= (act && (vv > eps * ndx || vv < -eps * ndx)) ? 1 : 0;
    }
    m = warp_max(m);
    for (int kk = 0; kk < q.nrl; ++kk) {
      const int rraw = lane + 32 * kk;
      const bool act = rraw < q.nrows;
      const int r = act ? rraw : 0;
      const double* R = q.R(r);
      const double* F = q.F(r);
      const int naux = q.I(r)[RI_AUX];
      const double ax = row_dot(q, R, q.I(r), q.v1) + F[R_U0] * F[R_DXA0] + F[R_U1] * F[R_DXA1];
      const double vv = ax / F[R_E];
      bad |= (act && vv > eps * ndx) ? 1 : 0;                              // u finite for every row
      bad |= (act && naux != AUX_HINGE && vv < -eps * ndx) ? 1 : 0;        // l finite unless hinge
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const double va = F[R_B0 + k] * F[R_DXA0 + k] / F[R_EA0 + k];
        bad |= (act && k < naux && va < -eps * ndx) ? 1 : 0;               // aux rows: l = 0 finite, u infinite
      }
    }
    const int nbad = warp_sum_int(bad);
    return (ndx > eps) && (qdx < -q.c * eps * ndx) && (m < q.c * eps * ndx) && (nbad == 0);
  };
  auto check_termination = [&](bool approximate) -> int {
    double eps_abs = st.eps_abs * eps_scale, eps_rel = st.eps_rel * eps_scale, epi = st.eps_prim_inf, edi = st.eps_dual_inf;
    if (approximate) {
      eps_abs *= 10; eps_rel *= 10; epi *= 10; edi *= 10;
    }
    const double eps_pri = eps_abs + eps_rel * fmax(n_z, n_ax);
    const double eps_dua = eps_abs + eps_rel * q.cinv * fmax(n_q, fmax(n_aty, n_px));
    const bool pri_ok = pri_res < eps_pri, dua_ok = dua_res < eps_dua;
    int res = QPS_UNSOLVED;
    if (pri_res > kOsqpInf || dua_res > kOsqpInf) res = QPS_NONCVX;
    else if (pri_ok && dua_ok) res = approximate ? QPS_SOLVED_INACC : QPS_SOLVED;
    else {  // warp-uniform branch; the certificates are only evaluated when their residual test failed
      const bool pinf = pri_ok ? false : primal_infeasible(epi);
      const bool dinf = dua_ok ? false : dual_infeasible(edi);
      if (pinf) res = approximate ? QPS_PINF_INACC : QPS_PINF;
      else if (dinf) res = approximate ? QPS_DINF_INACC : QPS_DINF;
    }
    return res;
  };

  // ---------------------------------------------------------------- one ADMM iteration (branch free in the lanes)
  auto admm_iteration = [&](bool keep_steps) {
    { PROF_T0();
    // rows: aux right-hand sides and the row multipliers of the reduced system
    for (int kk = 0; kk < q.nrl; ++kk) {
      const int rraw = lane + 32 * kk;
      const bool act = rraw < q.nrows;
      double* F = q.F(act ? rraw : 0);
      const double s = F[R_WRR] * F[R_Z] - F[R_Y];
      const double ra0 = q.sigma * F[R_XA0] - F[R_QA0] + F[R_U0] * s + F[R_B0] * (sysw.rho_aux * F[R_ZA0] - F[R_YA0]);
      const double ra1 = q.sigma * F[R_XA1] - F[R_QA1] + F[R_U1] * s + F[R_B1] * (sysw.rho_aux * F[R_ZA1] - F[R_YA1]);
      const double cf = row_reduce_coef(F, ra0, ra1, s);
      if (act) {
        F[R_RA0] = ra0;
        F[R_RA1] = ra1;
        F[R_COEF] = cf;
      }
    }
    __syncwarp();
    PROF_ADD(0); }
    { PROF_T0();
    // right-hand side  sigma x - q + A'(rho z - y)  (one lane per variable)
    scatter_columns(q, [&](int i) {
      const double rb = (q.ubs[i] - q.lbs[i] < kRhoTol) ? q.rho_eq : q.rho;
      return q.sigma * q.x[i] - q.qs[i] + q.beta[i] * (rb * q.zb[i] - q.yb[i]);
    });
    PROF_ADD(1); }
    { PROF_T0();
    block_solve(q, q.v1);
    PROF_ADD(2); }
    { PROF_T0();
    // rows: back-substitute aux, relax, project, dual update
    for (int kk = 0; kk < q.nrl; ++kk) {
      const int rraw = lane + 32 * kk;
      const bool act = rraw < q.nrows;
      const int r = act ? rraw : 0;
      const double* R = q.R(r);
      double* F = q.F(r);
      const double zeta = row_dot(q, R, q.I(r), q.v1);
      double a0, a1;
      row_backsub(F, zeta, a0, a1);
      const double zt = zeta + F[R_U0] * a0 + F[R_U1] * a1;
      const double Wr = F[R_WRR];
      const double zr = q.alpha * zt + (1.0 - q.alpha) * F[R_Z];
      double zn = zr + F[R_Y] * F[R_IWRR];
      zn = fmin(fmax(zn, F[R_LO]), F[R_UP]);
      const double dy = Wr * (zr - zn);
      double xn[2], dxa[2], zan[2], dya[2];
      const double inv_rho_aux = 1.0 / sysw.rho_aux;
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const double at = k ? a1 : a0, bb = F[R_B0 + k];
        const double xo = F[R_XA0 + k];
        xn[k] = q.alpha * at + (1.0 - q.alpha) * xo;
        dxa[k] = xn[k] - xo;
        const double zra = q.alpha * (bb * at) + (1.0 - q.alpha) * F[R_ZA0 + k];
        double z2 = zra + F[R_YA0 + k] * inv_rho_aux;
        z2 = fmin(fmax(z2, 0.0), kOsqpInf * F[R_EA0 + k]);
        zan[k] = z2;
        dya[k] = sysw.rho_aux * (zra - z2);
      }
      if (act) {
        F[R_Z] = zn;
        F[R_Y] += dy;
        F[R_DY] = dy;
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          F[R_XA0 + k] = xn[k];
          F[R_DXA0 + k] = dxa[k];
          F[R_ZA0 + k] = zan[k];
          F[R_YA0 + k] += dya[k];
          F[R_DYA0 + k] = dya[k];
        }
      }
    }
    PROF_ADD(3); }
    { PROF_T0();
    // trajectory variables and their bound rows
    const double inv_rho = 1.0 / q.rho, inv_rho_eq = 1.0 / q.rho_eq;
    for (int kk = 0; kk < q.npl; ++kk) {
      const int iraw = lane + 32 * kk;
      const bool act = iraw < N;
      const int i = act ? iraw : 0;
      const double beta = q.beta[i];
      const double lb = q.lbs[i], ub = q.ubs[i];
      const bool beq = ub - lb < kRhoTol;
      const double rb = beq ? q.rho_eq : q.rho, irb = beq ? inv_rho_eq : inv_rho;
      const double xt = q.v1[i];
      const double xn = q.alpha * xt + (1.0 - q.alpha) * q.x[i];
      const double zr = q.alpha * (beta * xt) + (1.0 - q.alpha) * q.zb[i];
      double zn = zr + q.yb[i] * irb;
      zn = fmin(fmax(zn, lb), ub);
      const double dy = rb * (zr - zn);
      if (act && keep_steps) {
        dxs[i] = xn - q.x[i];
        dyb[i] = dy;
      }
      if (act) {
        q.x[i] = xn;
        q.zb[i] = zn;
        q.yb[i] += dy;
      }
    }
    __syncwarp();
    PROF_ADD(4); }
  };

  // ADMM iterations, continuing from the current state until a termination test fires, max_iter, or the
  // slice budget is exhausted (status QPS_YIELD).
  auto run_admm = [&]() {
    status = QPS_UNSOLVED;
    bool stop = false;
    while (!stop) {
      if (iter >= st.max_iter) {  // max_iter reached without a verdict: approximate test, then MAX_ITER_REACHED
        if (!(st.check_termination > 0 && (iter % st.check_termination == 0))) info_pass();
        status = check_termination(true);
        if (status == QPS_UNSOLVED) status = QPS_MAXITER;
        stop = true;
      } else if (budget <= 0) {
        status = QPS_YIELD;
        stop = true;
      } else {
        --budget;
        ++iter;
        const bool can_check = st.check_termination > 0 && (iter % st.check_termination == 0);
        const bool rho_iter = st.adaptive_rho && st.adaptive_rho_interval > 0 && (iter % st.adaptive_rho_interval == 0);
        admm_iteration(can_check || iter == st.max_iter);
        if (can_check) {
          PROF_T0();
          info_pass();
          status = check_termination(false);
          PROF_ADD(5);
          if (status != QPS_UNSOLVED) stop = true;
        }
        if (!stop && rho_iter) {
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
              stop = true;
            }
          }
        }
      }
    }
  };

  // ---------------------------------------------------------------- polish (OSQP polish.c [EXT])
  // Equality-constrained QP on the guessed active set, solved as the delta-regularised KKT system with
  // iterative refinement, in its reduced form K_p = P + delta I + (1/delta) A_act' A_act (same aux
  // elimination and block factor as the ADMM system).  Returns false when K_p could not be factored.
  // `verified`: the polished point is primal feasible to kVerifyTol and every active inequality row has a
  // correctly signed multiplier, i.e. it is a KKT point of the QP = the unique minimiser.
  const double wp = 1.0 / st.delta;
  const SysW pw{true, st.delta, 0.0};
  auto polish_once = [&](bool& verified, double& p_pri, double& p_dua) -> bool {
    verified = false;
    for (int kk = 0; kk < q.npl; ++kk) {
      const int iraw = lane + 32 * kk;
      const bool inP = iraw < q.Np, act = iraw < N;
      const int i = inP ? iraw : 0;
      const double z = q.zb[i], y = q.yb[i];
      double w = 0.0;
      w = (act && (q.ubs[i] - z < y)) ? wp : w;        // upper active
      w = (act && (z - q.lbs[i] < -y)) ? -wp : w;      // lower active (tested first by OSQP)
      if (inP) {
        st_x[i] = q.x[i];
        st_zb[i] = z;
        st_yb[i] = y;
        q.zb[i] = w;                                   // signed polish weight
        q.x[i] = 0.0;                                  // polish iterate
        q.yb[i] = 0.0;                                 // polish multiplier
      }
    }
    for (int kk = 0; kk < q.nrl; ++kk) {
      const int rraw = lane + 32 * kk;
      const bool act = rraw < q.nrows;
      double* F = q.F(act ? rraw : 0);
      const int naux = q.I(act ? rraw : 0)[RI_AUX];
      double w = 0.0, b = 0.0;
      const bool up = F[R_UP] - F[R_Z] < F[R_Y], lo = F[R_Z] - F[R_LO] < -F[R_Y];
      w = up ? wp : w; b = up ? F[R_UP] : b;
      w = lo ? -wp : w; b = lo ? F[R_LO] : b;
      double wa[2];
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        double v = 0.0;
        v = (k < naux && (kOsqpInf * F[R_EA0 + k] - F[R_ZA0 + k] < F[R_YA0 + k])) ? wp : v;   // never in practice
        v = (k < naux && (F[R_ZA0 + k] - 0.0 < -F[R_YA0 + k])) ? -wp : v;                       // lower (0) active
        wa[k] = v;
      }
      if (act) {
        F[R_PW] = w;
        F[R_PB] = b;
        F[R_PY] = 0.0;
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          F[R_PWA0 + k] = wa[k];
          F[R_PYA0 + k] = 0.0;
          F[R_PX0 + k] = 0.0;
        }
      }
    }
    __syncwarp();
    if (!assemble_factor(q, pw)) return false;
    for (int it = 0; it <= st.polish_refine_iter + 1; ++it) {
      const bool last = (it == st.polish_refine_iter + 1);  // final pass: pending dual update + residuals only
      p_matvec(q, q.x, q.v2);  // v2 <- P xq
      double m_pri = 0.0, m_dua = 0.0;
      int bad_sign = 0;
      // rows: residual of the row, pending multiplier update, aux right-hand sides, row multiplier for A'
      for (int kk = 0; kk < q.nrl; ++kk) {
        const int rraw = lane + 32 * kk;
        const bool act = rraw < q.nrows;
        const int r = act ? rraw : 0;
        const double* R = q.R(r);
        double* F = q.F(r);
        const int naux = q.I(r)[RI_AUX];
        const double on = act ? 1.0 : 0.0;
        const double Wr = F[R_WRR];
        const double ax = row_dot(q, R, q.I(r), q.x) + F[R_U0] * F[R_PX0] + F[R_U1] * F[R_PX1];
        const double py = F[R_PY] + ((it > 0) ? Wr * (ax - F[R_PB]) : 0.0);
        const double e = py + (last ? 0.0 : Wr * (ax - F[R_PB]));
        const double zc = fmin(fmax(ax, F[R_LO]), F[R_UP]);
        m_pri = fmax(m_pri, on * fabs((ax - zc) / F[R_E]));
        bad_sign |= (act && last && Wr != 0.0 && naux == AUX_HINGE && py < -kVerifyTol) ? 1 : 0;  // upper active needs y >= 0
        double pya[2], ra[2];
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          const double bb = F[R_B0 + k], u = F[R_U0 + k], qa = F[R_QA0 + k];
          const double wa = fabs(F[R_PWA0 + k]);
          const double axb = bb * F[R_PX0 + k];
          pya[k] = F[R_PYA0 + k] + ((it > 0) ? wa * axb : 0.0);
          const double ea = pya[k] + (last ? 0.0 : wa * axb);
          m_pri = fmax(m_pri, on * fabs((axb - fmax(axb, 0.0)) / F[R_EA0 + k]));
          bad_sign |= (act && last && wa != 0.0 && pya[k] > kVerifyTol) ? 1 : 0;  // aux >= 0 held at 0 needs y <= 0
          m_dua = fmax(m_dua, on * fabs((qa + u * py + bb * pya[k]) / F[R_DA0 + k]));
          ra[k] = -qa - u * e - bb * ea;
        }
        const double cf = last ? py : row_reduce_coef(F, ra[0], ra[1], -e);
        if (act) {
          F[R_PY] = py;
          F[R_PYA0] = pya[0];
          F[R_PYA1] = pya[1];
          F[R_RA0] = ra[0];
          F[R_RA1] = ra[1];
          F[R_COEF] = cf;
        }
      }
      __syncwarp();
      if (last) {
        // dual residual  P x + q + A'y  over the trajectory variables
        scatter_columns(q, [&](int i) { return q.v2[i] + q.qs[i] + q.beta[i] * q.yb[i]; });
      } else {
        // rd = -(P x + q) - beta * (y + W (A x - b)) + A' coef
        scatter_columns(q, [&](int i) {
          const double beta = q.beta[i];
          const double w = fabs(q.zb[i]);
          const double bnd = q.zb[i] > 0 ? q.ubs[i] : q.lbs[i];
          return -(q.v2[i] + q.qs[i]) - beta * (q.yb[i] + w * (beta * q.x[i] - bnd));
        });
      }
      for (int kk = 0; kk < q.npl; ++kk) {
        const int iraw = lane + 32 * kk;
        const bool act = iraw < N;
        const int i = act ? iraw : 0;
        const double on = act ? 1.0 : 0.0;
        const double beta = q.beta[i];
        const double ax = beta * q.x[i];
        const double w = fabs(q.zb[i]);
        const double lb = q.lbs[i], ub = q.ubs[i];
        const double zc = fmin(fmax(ax, lb), ub);
        m_pri = fmax(m_pri, on * fabs((ax - zc) * q.Dz[i] / beta));
        const bool ineq = act && last && w != 0.0 && (ub - lb >= kRhoTol);
        bad_sign |= (ineq && q.zb[i] > 0 && q.yb[i] < -kVerifyTol) ? 1 : 0;
        bad_sign |= (ineq && q.zb[i] < 0 && q.yb[i] > kVerifyTol) ? 1 : 0;
        m_dua = fmax(m_dua, (act && last) ? fabs(q.v1[i] / q.Dz[i]) : 0.0);
      }
      __syncwarp();
      if (last) {
        p_pri = warp_max(m_pri);
        p_dua = warp_max(m_dua) * q.cinv;
        const int nbad = warp_sum_int(bad_sign);
        verified = (nbad == 0) && (p_pri <= kVerifyTol) && isfinite(p_pri) && isfinite(p_dua);
      } else {
        block_solve(q, q.v1);
        for (int kk = 0; kk < q.nrl; ++kk) {
          const int rraw = lane + 32 * kk;
          const bool act = rraw < q.nrows;
          const int r = act ? rraw : 0;
          const double* R = q.R(r);
          double* F = q.F(r);
          double a0, a1;
          row_backsub(F, row_dot(q, R, q.I(r), q.v1), a0, a1);
          if (act) {
            F[R_PX0] += a0;
            F[R_PX1] += a1;
          }
        }
        for (int kk = 0; kk < q.npl; ++kk) {
          const int iraw = lane + 32 * kk;
          const bool act = iraw < N;
          const int i = act ? iraw : 0;
          const double xn = q.x[i] + q.v1[i];
          // multiplier update of the variable-bound rows with the new iterate (the rows do theirs at the
          // start of the next pass, where A x is recomputed anyway)
          const double w = fabs(q.zb[i]);
          const double yn = q.yb[i] + w * (q.beta[i] * xn - (q.zb[i] > 0 ? q.ubs[i] : q.lbs[i]));
          if (act) {
            q.x[i] = xn;
            q.yb[i] = yn;
          }
        }
        __syncwarp();
      }
    }
    return true;
  };
  auto restore_admm_state = [&](bool keep_polished_x) {
    for (int kk = 0; kk < q.npl; ++kk) {
      const int iraw = lane + 32 * kk;
      if (iraw < q.Np) {
        if (!keep_polished_x) q.x[iraw] = st_x[iraw];
        q.zb[iraw] = st_zb[iraw];
        q.yb[iraw] = st_yb[iraw];
      }
    }
    __syncwarp();
  };

  // ---- main loop: ADMM -> polish -> verify; on a failed verification ADMM continues with 10x tighter ------
  // tolerances (DESIGN.md deviation D2).
  bool done = !factor_ok;
  while (!done) {
    run_admm();
    out.pri_res = pri_res;
    out.dua_res = dua_res;
    if (status != QPS_SOLVED || !st.polishing) {
      done = true;
    } else {
      bool verified = false;
      double p_pri = 0.0, p_dua = 0.0;
      const bool factored = polish_once(verified, p_pri, p_dua);
      out.pol_factor_ok = factored ? 1 : 0;
      out.pol_pri = p_pri;
      out.pol_dua = p_dua;
      out.rounds = round;
      if (factored && verified) {
        out.polish = 1;
        done = true;
      } else if (round >= kVerifyRounds || iter >= st.max_iter) {  // OSQP's acceptance rule
        const bool ok = factored && ((p_pri < pri_res && p_dua < dua_res) || (p_pri < pri_res && dua_res < 1e-10) ||
                                     (p_dua < dua_res && pri_res < 1e-10)) && isfinite(p_pri) && isfinite(p_dua);
        out.polish = ok ? 2 : -1;
        done = true;
      } else {
        ++round;
        eps_scale *= 0.1;
        restore_admm_state(false);
        if (!assemble_factor(q, sysw)) {  // back to the ADMM factor
          status = QPS_NONCVX;
          done = true;
        }
      }
    }
  }
  out.iters = iter;
  out.status = status;
  out.rho = rho;
  out.c = q.c;
  rs.iter = iter;
  rs.round = round;
  rs.rho = rho;
  rs.eps_scale = eps_scale;
  rs.rho_updates = out.rho_updates;
  rs.c = q.c;
  if (out.polish != 0) {
    // Adopt the polished PRIMAL point when accepted.  The duals kept for the next warm start are always the
    // ADMM duals: polished duals are non-unique on degenerate active sets (DESIGN.md deviation D1).
    if (out.polish > 0) {
      for (int r = lane; r < q.nrows; r += 32) {
        double* F = q.F(r);
        for (int k = 0; k < 2; ++k) F[R_XA0 + k] = F[R_PX0 + k];
      }
    }
    restore_admm_state(out.polish > 0);
  }
  return out;
}

// ---------------------------------------------------------------------------------------------------
// Kernel: QP assembly (optimizers.cpp:781-799 + osqp_interface.cpp:170-281 in fixed layout) + solve slice.
// grid = B, block = 32 (one warp per trajectory).
__global__ void __launch_bounds__(32, 4) qp_kernel(DevProblem p, const double* x_override /*kernel-level API*/,
                                                   const double* trust_override, int* admm_iters_out,
                                                   int* polish_out, int slice) {
  extern __shared__ double sm[];
  const int b = blockIdx.x;
  const int lane = threadIdx.x;
  if (!x_override && (p.status[b] != 5 || p.qp_done[b] != 0)) return;  // finished, or waiting for its evaluation
  const int N = p.N, T = p.T, D = p.D;
  QpCtx q;
  q.N = N; q.T = T; q.D = D; q.lane = lane;
  q.nb = 2 * D;
  q.M = qp_block_count(N, q.nb);
  q.Np = q.M * q.nb;
  q.Wd = q.nb + 2;
  q.npl = (q.Np + 31) / 32;
  const QpSmem S = qp_smem_layout(N, q.nb);
  q.CN = (p.row_stride - R_NF) / 2;
  q.RS = p.row_stride;
  q.Kb = sm + S.Kb; q.Linv = sm + S.Linv; q.beta = sm + S.beta;
  q.x = sm + S.x; q.zb = sm + S.zb; q.yb = sm + S.yb; q.v1 = sm + S.v1;
  q.qs = sm + S.qs; q.lbs = sm + S.lbs; q.ubs = sm + S.ubs; q.xch = sm + S.xch;
  q.colptr = reinterpret_cast<int*>(sm + S.colptr);
  q.maxcol = reinterpret_cast<int*>(sm + S.maxcol);
  q.rows = p.rows + static_cast<size_t>(b) * p.max_rows * p.row_stride;
  q.rints = p.row_ints + static_cast<size_t>(b) * p.max_rows * RI_NINTS;
  int* mylist = p.lists + static_cast<size_t>(b) * p.list_stride;
  int* colptr = mylist;                                   // [Np+1] master copy (shared copy in q.colptr)
  int* colent = mylist + q.Np + 1;                        // [max_rows*CN]
  int* obj_start = colent + static_cast<size_t>(p.max_rows) * q.CN;  // [n_objs+1]
  q.colent = colent;
  q.Pband = p.Pband;
  // per-trajectory global vectors: dxs dyb st_x st_zb st_yb | scaled qs lbs ubs (master) | Dz | v2
  double* gvec = p.scratch + static_cast<size_t>(b) * 10 * q.Np;
  q.scratch = gvec;
  double* g_qs = gvec + 5 * q.Np;
  double* g_lbs = gvec + 6 * q.Np;
  double* g_ubs = gvec + 7 * q.Np;
  q.Dz = gvec + 8 * q.Np;
  q.v2 = gvec + 9 * q.Np;
  double* park = p.park + static_cast<size_t>(b) * 4 * q.Np;     // x zb yb beta of a parked solve
  double *qs = q.qs, *lbs = q.lbs, *ubs = q.ubs;
  int* meta = p.ws_meta + static_cast<size_t>(b) * 8;  // 0..3 warm-start key, 4 phase, 5 nrows, 6 n_aux, 7 nnzA
  const int n_obj = p.n_costs + p.n_cnts;
  QpResume rs{};
  const bool resume = !x_override && meta[4] == 1;
  int nr = 0, n_aux = 0, nnzA = 0;
  bool warm = false;

  if (!resume) {
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
      lbs[i] = fmax(fmax(xi - trust, lb), -kOsqpInf);
      ubs[i] = fmin(fmin(xi + trust, ub), kOsqpInf);
      qs[i] = p.qlin[i];
      q.x[i] = xc[i];  // linearisation point (until the solver takes over x)
    }
    __syncwarp();

    // ---- rows in the reference's canonical order: permanent rows, cost rows, penalised constraint rows -----
    // (every record is padded: CN coefficients, zeros beyond the row's own count)
    for (int f = lane; f < p.n_fixed; f += 32) {  // fixed_timesteps / fixed_dofs rows: x_k - init_k == 0
      const int var = p.fixed_vars[f];
      double* R = q.R(f);
      int* I = q.rints + static_cast<size_t>(f) * RI_NINTS;
      for (int k = 0; k < q.CN; ++k) R[k] = (k == 0) ? 1.0 : 0.0;
      R[2 * q.CN + R_C] = -p.init_traj[static_cast<size_t>(b) * N + var];
      R[2 * q.CN + R_W] = 0.0;
      I[RI_BASE] = var; I[RI_CNT] = 1; I[RI_STRIDE] = D; I[RI_AUX] = AUX_NONE; I[RI_OBJ] = -1;
    }
    nr += p.n_fixed;
    nnzA += p.n_fixed;
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
          for (int i = 0; i < q.CN; ++i) R[i] = (i <= o.order) ? wst[o.order][i] * sgn : 0.0;
          R[2 * q.CN + R_C] = cst;
          R[2 * q.CN + R_W] = w_aux;
          I[RI_BASE] = t * D + d; I[RI_CNT] = o.order + 1; I[RI_STRIDE] = D;
          I[RI_AUX] = (per == 1) ? AUX_ABS : AUX_HINGE; I[RI_OBJ] = oi;
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
          for (int j = 0; j < q.CN; ++j) {
            const double Jj = (j < D) ? J[j] : 0.0;
            dot += Jj * q.x[o.first * D + min(j, D - 1)];
            const double a = (fabs(Jj) > thr) ? Jj : 0.0;
            R[j] = a;
            nz += (a != 0.0);
          }
          R[2 * q.CN + R_C] = cart_err[o.src_off + k] - dot;
          R[2 * q.CN + R_W] = w_aux;
          I[RI_BASE] = o.first * D; I[RI_CNT] = D; I[RI_STRIDE] = 1; I[RI_AUX] = AUX_ABS; I[RI_OBJ] = oi;
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
            for (int j = 0; j < q.CN; ++j) {
              const double g = (j < D) ? cr[j] : 0.0;
              dot += g * q.x[o.first * D + min(j, D - 1)];
              const double a = -g * scale;
              R[j] = (j < D) ? a : 0.0;
              nz += (j < D && a != 0.0);
            }
            R[2 * q.CN + R_C] = (cr[D + 1] - cr[D] + dot) * scale;
            R[2 * q.CN + R_W] = is_cnt ? w_aux : cr[D + 2];
            I[RI_BASE] = o.first * D; I[RI_CNT] = D; I[RI_STRIDE] = 1; I[RI_AUX] = AUX_HINGE; I[RI_OBJ] = oi;
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
    nnzA += N + n_aux;  // identity rows carrying the variable bounds
    __syncwarp();

    // ---- per-column entry lists (canonical row order inside every column; real coefficients only) ---------
    for (int i = lane; i <= q.Np; i += 32) colptr[i] = 0;
    __syncwarp();
    for (int r = lane; r < nr; r += 32) {
      const int* I = q.I(r);
      for (int k = 0; k < I[RI_CNT]; ++k) atomicAdd(&colptr[I[RI_BASE] + k * I[RI_STRIDE] + 1], 1);
    }
    __syncwarp();
    if (lane == 0)
      for (int i = 0; i < q.Np; ++i) colptr[i + 1] += colptr[i];
    __syncwarp();
    {
      int* fill = reinterpret_cast<int*>(q.v1);  // Np ints of scratch
      for (int i = lane; i < q.Np; i += 32) fill[i] = colptr[i];
      __syncwarp();
      for (int r = 0; r < nr; ++r) {  // serial over rows keeps every column in canonical row order
        const int* I = q.I(r);
        if (lane < I[RI_CNT]) {
          const int var = I[RI_BASE] + lane * I[RI_STRIDE];
          colent[fill[var]++] = (r << 5) | lane;
        }
        __syncwarp();
      }
    }
    q.nrows = nr;
  } else {
    nr = meta[5];
    n_aux = meta[6];
    nnzA = meta[7];
    q.nrows = nr;
  }
  q.nrl = (nr + 31) / 32;
  // shared copies of the column pointers and the per-chunk longest column (uniform inner trip counts)
  for (int i = lane; i <= q.Np; i += 32) q.colptr[i] = colptr[i];
  __syncwarp();
  for (int kk = 0; kk < q.npl; ++kk) {
    const int i = lane + 32 * kk;
    const int len = (i < q.Np) ? q.colptr[i + 1] - q.colptr[i] : 0;
    const int mx = warp_max_int(len);
    if (lane == 0) q.maxcol[kk] = mx;
  }
  __syncwarp();

  if (!resume) {
    // ---- warm start decision (createOrUpdateSolver, osqp_interface.cpp:283-370) ---------------------------
    warm = !x_override && p.qp.warm_starting && meta[3] == 1 && meta[0] == n_aux && meta[1] == nr && meta[2] == nnzA;
    { PROF_T0(); qp_scale(q, p.qp, n_aux); PROF_ADD(7); }
    for (int i = lane; i < q.Np; i += 32) {  // master copies for the resume path
      g_qs[i] = qs[i];
      g_lbs[i] = lbs[i];
      g_ubs[i] = ubs[i];
    }
  } else {
    rs.iter = p.rs_int[b * 4 + 0];
    rs.round = p.rs_int[b * 4 + 1];
    rs.rho_updates = p.rs_int[b * 4 + 2];
    rs.rho = p.rs_dbl[b * 4 + 0];
    rs.eps_scale = p.rs_dbl[b * 4 + 1];
    rs.c = p.rs_dbl[b * 4 + 2];
    q.c = rs.c;
    q.cinv = 1.0 / q.c;
    for (int i = lane; i < q.Np; i += 32) {
      q.x[i] = park[i];
      q.zb[i] = park[q.Np + i];
      q.yb[i] = park[2 * q.Np + i];
      q.beta[i] = park[3 * q.Np + i];
      qs[i] = g_qs[i];
      lbs[i] = g_lbs[i];
      ubs[i] = g_ubs[i];
    }
  }
  __syncwarp();

  PROF_T0();
  QpOut res = qp_solve_warp(q, p.qp, !resume, warm, p.ws_rho[b], p.ws_x + static_cast<size_t>(b) * N,
                            p.ws_yb + static_cast<size_t>(b) * N, rs, x_override ? (1 << 30) : slice);
  PROF_ADD(8);
#ifdef TB200_PROFILE
  if (lane == 0) { atomicAdd(&g_prof[9], 1ull); }
#endif
  __syncwarp();

  if (res.status == QPS_YIELD) {  // park the solve
    for (int i = lane; i < q.Np; i += 32) {
      park[i] = q.x[i];
      park[q.Np + i] = q.zb[i];
      park[2 * q.Np + i] = q.yb[i];
      park[3 * q.Np + i] = q.beta[i];
    }
    if (lane == 0) {
      meta[4] = 1; meta[5] = nr; meta[6] = n_aux; meta[7] = nnzA;
      p.rs_int[b * 4 + 0] = rs.iter; p.rs_int[b * 4 + 1] = rs.round; p.rs_int[b * 4 + 2] = rs.rho_updates;
      p.rs_dbl[b * 4 + 0] = rs.rho; p.rs_dbl[b * 4 + 1] = rs.eps_scale; p.rs_dbl[b * 4 + 2] = rs.c;
    }
    return;
  }

  // ---- unscale, store the solution (and the warm-start state), model values -----------------------------
  double* nx = p.new_x + static_cast<size_t>(b) * N;
  for (int i = lane; i < N; i += 32) {
    const double xu = q.Dz[i] * q.x[i];
    nx[i] = xu;
    q.v1[i] = xu;  // unscaled solution for the model-value pass
    p.ws_x[static_cast<size_t>(b) * N + i] = xu;
    p.ws_yb[static_cast<size_t>(b) * N + i] = q.cinv * (q.beta[i] / q.Dz[i]) * q.yb[i];
  }
  __syncwarp();
  for (int r = lane; r < nr; r += 32) {
    double* R = q.R(r);
    double* F = q.F(r);
    const int* I = q.I(r);
    const int aux = I[RI_AUX];
    F[R_Y] = q.cinv * F[R_E] * F[R_Y];
    for (int k = 0; k < 2; ++k) {
      F[R_XA0 + k] = (k < aux) ? F[R_DA0 + k] * F[R_XA0 + k] : 0.0;
      F[R_YA0 + k] = (k < aux) ? q.cinv * F[R_EA0 + k] * F[R_YA0 + k] : 0.0;
    }
    double val = F[R_C];
    for (int i = 0; i < I[RI_CNT]; ++i) val += R[i] * q.v1[I[RI_BASE] + i * I[RI_STRIDE]];
    // ConvexConstraints::violations (modeling.cpp:132-142) for constraint rows; hinge/abs cost = w * aux values
    F[R_MV] = (aux == AUX_ABS || aux == AUX_NONE) ? fabs(val) : fmax(val, 0.0);
  }
  __syncwarp();
  for (int oi = lane; oi < n_obj; oi += 32) {  // per object sums, canonical order, one lane per object
    const bool is_cnt = oi >= p.n_costs;
    double s = 0.0;
    if (!is_cnt && p.cost_objs[oi].kind == OBJ_JOINT_EQ_COST) s = joint_obj_value(p, p.cost_objs[oi], q.v1);  // exact quadratic
    for (int r = obj_start[oi]; r < obj_start[oi + 1]; ++r) {
      const double* F = q.F(r);
      if (is_cnt) s += F[R_MV];
      else s += F[R_W] * (F[R_XA0] + F[R_XA1]);  // ConvexObjective::value: the penalty terms use the aux values
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
    meta[0] = n_aux; meta[1] = nr; meta[2] = nnzA; meta[3] = (cvx == 0) ? 1 : 0; meta[4] = 0;
    p.ws_rho[b] = res.rho;
    if (!x_override) {
      p.n_admm_iters[b] += res.iters;
      p.qp_done[b] = 1;
    }
    if (admm_iters_out) admm_iters_out[b] = res.iters;
    if (polish_out) polish_out[b] = res.polish;
    double* g = p.dbg + static_cast<size_t>(b) * 16;
    g[0] = res.status; g[1] = res.iters; g[2] = res.polish; g[3] = res.rho; g[4] = res.pri_res; g[5] = res.dua_res;
    g[6] = res.pol_pri; g[7] = res.pol_dua; g[8] = res.c; g[9] = res.pol_factor_ok; g[10] = res.rho_updates;
    g[11] = nr; g[12] = n_aux; g[13] = nnzA; g[14] = warm ? 1 : 0; g[15] = res.rounds;
  }
}

}  // namespace tb200
