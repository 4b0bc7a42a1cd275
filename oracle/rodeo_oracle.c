/*
 * CPU port of the rodeo filtering hot path in plain C + OpenMP  --  TEST / BASELINE INFRASTRUCTURE ONLY.
 *
 * Role: the `cpu_baseline` ("port") and `--impl reference` legs of bench.py, and a fast second checker for
 * large-batch parity tests.  It restates, with dense full-matrix arithmetic in the reference's operation order,
 * the same functions as oracle/rodeo_oracle.py (which is the canonical, LAPACK-backed restatement and the one the
 * known-answer tests pin): predict/update/forecast (reference src/rodeo/kalmantv/standard.py:31-103,308-336),
 * _smooth/smooth_mv (:160-217), multivariate_normal_logpdf with the 1e-8 cut-off (src/rodeo/utils.py:60-78),
 * interrogate_{kramer,schober,rodeo} (src/rodeo/interrogate.py:50-115), _solve_filter / solve_mv
 * (src/rodeo/solve.py:31-122,208-302) and dalton (src/rodeo/inference/dalton.py:39-235).
 * Parity chain: real JAX cannot run in this image (no jax / jaxlib, no network), but the reference's own source runs
 * over oracle/jaxshim and its outputs are committed (tests/golden/reference_vectors.npz); tests/test_reference_golden.py
 * pins the NumPy oracle to them, and tests/test_oracle_c.py pins this port to the NumPy oracle at 1e-11.
 *
 * Nothing under rodeo_b200/ links or loads this file.
 *
 * Restrictions: n_bmeas = 1, n_bobs = 1, n_block <= 6, n_bstate <= 4; built-in models only.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* arithmetic type: double (the baseline / checker) or, with -DRODEO_LD, x87 long double (64-bit mantissa): the same
 * recursion evaluated ~2000x more finely, used by the tests to measure the float64 noise floor of dalton per theta.
 * The exported interface is double arrays either way. */
#ifdef RODEO_LD
typedef long double real;
#define R_FABS fabsl
#define R_SQRT sqrtl
#define R_LOG logl
#define R_EXP expl
#define R_SIN sinl
#else
typedef double real;
#define R_FABS fabs
#define R_SQRT sqrt
#define R_LOG log
#define R_EXP exp
#define R_SIN sin
#endif

#define NBMAX 6
#define PMAX 4

enum { M_FITZ = 0, M_LORENZ = 1, M_SECOND = 2, M_HES1 = 3, M_SEIRAH = 4 };
enum { I_KRAMER = 0, I_CHKREBTII = 1, I_SCHOBER = 2, I_RODEO = 3 };

/* f (nb) and the block-diagonal Jacobian J (nb x p, own block only) at X (nb x p) */
static void ode_eval(int model, int nb, int p, const real* X, real t, const real* th, real* f, real* J) {
  memset(J, 0, sizeof(real) * nb * p);
  switch (model) {
    case M_FITZ: {
      real a = th[0], b = th[1], c = th[2], V = X[0], R = X[p];
      f[0] = c * (V - V * V * V / 3 + R);
      f[1] = -1 / c * (V - a + b * R);
      J[0] = c * (1 - V * V);
      J[p] = -1 / c * b;
    } break;
    case M_LORENZ: {
      real rho = th[0], sg = th[1], be = th[2], x = X[0], y = X[p], z = X[2 * p];
      f[0] = -sg * x + sg * y; f[1] = rho * x - y - x * z; f[2] = -be * z + x * y;
      J[0] = -sg; J[p] = -1.0; J[2 * p] = -be;
    } break;
    case M_SECOND: {
      f[0] = R_SIN(th[0] * t) - th[1] * X[0];
      J[0] = -th[1];
    } break;
    case M_HES1: {
      real P = R_EXP(X[0]), Mm = R_EXP(X[p]), H = R_EXP(X[2 * p]);
      real a = th[0], b = th[1], c = th[2], d = th[3], e = th[4], ff = th[5], g = th[6];
      f[0] = -a * H + b * Mm / P - c;
      f[1] = -d + e / (1 + P * P) / Mm;
      f[2] = -a * P + ff / (1 + P * P) / H - g;
      J[0] = -b * Mm / P; J[p] = -e / (1 + P * P) / Mm; J[2 * p] = -ff / (1 + P * P) / H;
    } break;
    case M_SEIRAH: {
      real S = X[0], E = X[p], I = X[2 * p], R = X[3 * p], A = X[4 * p], H = X[5 * p];
      real b = th[0], r = th[1], al = th[2], De = th[3], DI = th[4], Dq = th[5], Dh = 30.0;
      real N = S + E + I + R + A + H, g = b * (I + al * A);
      f[0] = -g * S / N; f[1] = g * S / N - E / De; f[2] = r * E / De - I / Dq - I / DI;
      f[3] = (I + A) / DI + H / Dh; f[4] = (1 - r) * E / De - A / DI; f[5] = I / Dq - H / Dh;
      J[0] = -g / N + g * S / (N * N); J[p] = -g * S / (N * N) - 1 / De; J[2 * p] = -1 / Dq - 1 / DI;
      J[3 * p] = 0.0; J[4 * p] = -1 / DI; J[5 * p] = -1 / Dh;
    } break;
  }
}

/* ---- dense helpers on p x p row-major matrices -------------------------------------------------------------- */
static void predict(int p, const real* Q, const real* R, const real* mu, const real* S, real* mp, real* Sp) {
  real A[PMAX * PMAX];
  for (int i = 0; i < p; ++i) {
    real m = 0;
    for (int j = 0; j < p; ++j) m += Q[i * p + j] * mu[j];
    mp[i] = m;
    for (int k = 0; k < p; ++k) {
      real a = 0;
      for (int j = 0; j < p; ++j) a += Q[i * p + j] * S[j * p + k];
      A[i * p + k] = a;
    }
  }
  for (int i = 0; i < p; ++i)
    for (int j = 0; j < p; ++j) {
      real a = 0;
      for (int k = 0; k < p; ++k) a += A[i * p + k] * Q[j * p + k];
      Sp[i * p + j] = a + R[i * p + j];
    }
}

/* X = S^{-1} B, S mm x mm (mm <= 2), B mm x p; partial-pivot Gaussian elimination as LAPACK getrf/getrs */
static void lu_solve(int mm, int p, const real* S, real* Bm) {
  if (mm == 1) { for (int i = 0; i < p; ++i) Bm[i] /= S[0]; return; }
  real A[4] = {S[0], S[1], S[2], S[3]};
  if (R_FABS(A[2]) > R_FABS(A[0])) {
    real t;
    t = A[0]; A[0] = A[2]; A[2] = t; t = A[1]; A[1] = A[3]; A[3] = t;
    for (int i = 0; i < p; ++i) { t = Bm[i]; Bm[i] = Bm[p + i]; Bm[p + i] = t; }
  }
  real l = A[2] / A[0];
  A[3] -= l * A[1];
  for (int i = 0; i < p; ++i) Bm[p + i] -= l * Bm[i];
  for (int i = 0; i < p; ++i) {
    Bm[p + i] /= A[3];
    Bm[i] = (Bm[i] - A[1] * Bm[p + i]) / A[0];
  }
}

static const real LOG2PI = (real)1.8378770664093454835606594728112353L;

static real logpdf_term(real w, real z) {
  if (R_FABS(w) <= (real)1e-8) return 0.0;               /* ~isclose(w, 0, rtol=1e-300): default atol 1e-8 */
  return -0.5 * (z * z / w + R_LOG(w)) - 0.5 * LOG2PI;
}

/* log N(res + mean; mean, S) for mm <= 2 via the symmetric eigen-decomposition */
static real logpdf(int mm, const real* S, const real* res) {
  if (mm == 1) return logpdf_term(S[0], res[0]);
  real a = S[0], b = S[1], c = S[3];
  real tr = a + c, df = a - c, rt = R_SQRT(df * df + 4 * b * b);
  real w1 = 0.5 * (tr + (tr >= 0 ? rt : -rt));  /* larger |.| root first, then the product for the other */
  real w2 = (w1 != 0.0) ? (a * c - b * b) / w1 : 0.0;
  real v0, v1;
  if (R_FABS(w1 - a) + R_FABS(b) > R_FABS(w1 - c) + R_FABS(b)) { v0 = b; v1 = w1 - a; if (R_FABS(v0) + R_FABS(v1) == 0) { v0 = 1; v1 = 0; } }
  else { v0 = w1 - c; v1 = b; if (R_FABS(v0) + R_FABS(v1) == 0) { v0 = 1; v1 = 0; } }
  real nrm = R_SQRT(v0 * v0 + v1 * v1);
  v0 /= nrm; v1 /= nrm;
  real z1 = v0 * res[0] + v1 * res[1], z2 = -v1 * res[0] + v0 * res[1];
  return logpdf_term(w1, z1) + logpdf_term(w2, z2);
}

/* update with mm rows wm (mm x p), offsets d, noise V (mm x mm), observed x; returns the forecast log-pdf */
static real update(int p, int mm, real* mu, real* S, const real* wm, const real* d, const real* V,
                     const real* x, int want_lp) {
  real wS[2 * PMAX], Sm[4], Kt[2 * PMAX], res[2];
  for (int r = 0; r < mm; ++r) {
    real mz = 0;
    for (int i = 0; i < p; ++i) mz += wm[r * p + i] * mu[i];
    res[r] = x[r] - (mz + d[r]);
    for (int j = 0; j < p; ++j) {
      real a = 0;
      for (int i = 0; i < p; ++i) a += wm[r * p + i] * S[i * p + j];
      wS[r * p + j] = a;                           /* var_meas_state_pred */
    }
  }
  for (int r = 0; r < mm; ++r)
    for (int s = 0; s < mm; ++s) {
      real a = 0;
      for (int j = 0; j < p; ++j) a += wS[r * p + j] * wm[s * p + j];
      Sm[r * mm + s] = a + V[r * mm + s];
    }
  real lp = want_lp ? logpdf(mm, Sm, res) : 0.0;
  for (int r = 0; r < mm; ++r)                      /* (S_p wm^T)^T = rows r: S_p wm[r] */
    for (int i = 0; i < p; ++i) {
      real a = 0;
      for (int j = 0; j < p; ++j) a += S[i * p + j] * wm[r * p + j];
      Kt[r * p + i] = a;
    }
  lu_solve(mm, p, Sm, Kt);
  for (int i = 0; i < p; ++i) {
    real m = 0;
    for (int r = 0; r < mm; ++r) m += Kt[r * p + i] * res[r];
    mu[i] += m;
  }
  real Sn[PMAX * PMAX];
  for (int i = 0; i < p; ++i)
    for (int j = 0; j < p; ++j) {
      real a = 0;
      for (int r = 0; r < mm; ++r) a += Kt[r * p + i] * wS[r * p + j];
      Sn[i * p + j] = S[i * p + j] - a;
    }
  memcpy(S, Sn, sizeof(real) * p * p);
  return lp;
}

/* interrogation for all blocks: wm (nb x p), d (nb), V (nb) */
static void interrogate(int model, int interr, int nb, int p, const real* W, const real* mu, const real* S,
                        real t, const real* th, real* wm, real* d, real* V) {
  real f[NBMAX], J[NBMAX * PMAX];
  ode_eval(model, nb, p, mu, t, th, f, J);
  for (int b = 0; b < nb; ++b) {
    if (interr == I_KRAMER) {
      real jm = 0;
      for (int j = 0; j < p; ++j) { wm[b * p + j] = W[b * p + j] + (-J[b * p + j]); jm += J[b * p + j] * mu[b * p + j]; }
      d[b] = -f[b] + jm; V[b] = 0.0;
    } else {
      for (int j = 0; j < p; ++j) wm[b * p + j] = W[b * p + j];
      d[b] = -f[b];
      V[b] = 0.0;
      if (interr == I_RODEO) {
        real a = 0;
        for (int i = 0; i < p; ++i) {
          real u = 0;
          for (int j = 0; j < p; ++j) u += W[b * p + j] * S[(b * p + j) * p + i];
          a += u * W[b * p + i];
        }
        V[b] = a;
      }
    }
  }
}

static real step_time(real t_min, real t_max, int n, int N) { return t_min + (t_max - t_min) * (n + 1) / N; }

/* one filter step for all blocks; obs_i >= 0 adds the observation rows (dalton zy_update); returns sum log-pdf */
static real filter_step(int model, int interr, int nb, int p, const real* W, const real* Q, const real* R,
                          real* mu, real* S, real t, const real* th, int obs_i, const real* obs_data,
                          const real* obs_weight, const real* obs_var, int want_lp) {
  real mp[NBMAX * PMAX], Sp[NBMAX * PMAX * PMAX], wm[NBMAX * PMAX], d[NBMAX], V[NBMAX];
  for (int b = 0; b < nb; ++b) predict(p, Q + b * p * p, R + b * p * p, mu + b * p, S + b * p * p, mp + b * p, Sp + b * p * p);
  interrogate(model, interr, nb, p, W, mp, Sp, t, th, wm, d, V);
  real lp = 0.0;
  for (int b = 0; b < nb; ++b) {
    if (obs_i < 0) {
      real x = 0.0;
      lp += update(p, 1, mp + b * p, Sp + b * p * p, wm + b * p, d + b, V + b, &x, want_lp);
    } else {
      real wa[2 * PMAX], da[2] = {d[b], 0.0}, Va[4] = {V[b], 0.0, 0.0, obs_var[obs_i * nb + b]};
      real xa[2] = {0.0, obs_data[obs_i * nb + b]};
      memcpy(wa, wm + b * p, sizeof(real) * p);
      memcpy(wa + p, obs_weight + (obs_i * nb + b) * p, sizeof(real) * p);
      lp += update(p, 2, mp + b * p, Sp + b * p * p, wa, da, Va, xa, want_lp);
    }
  }
  memcpy(mu, mp, sizeof(real) * nb * p);
  memcpy(S, Sp, sizeof(real) * nb * p * p);
  return lp;
}

int rodeo_oracle_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* dalton log-likelihood per theta */
static void to_real(real* dst, const double* src, long n) { for (long k = 0; k < n; ++k) dst[k] = (real)src[k]; }

int rodeo_oracle_dalton(int model, int interr, long B, int N, int nb, int p, int n_theta, double t_min, double t_max,
                        const double* W_, const double* Q_, const double* R_, const double* X0, const double* theta,
                        int n_obs, const int* obs_ind, const double* obs_data_, const double* obs_weight_,
                        const double* obs_var_, double* out, int n_threads) {
  if (nb > NBMAX || p > PMAX || n_theta > 16 || interr == I_CHKREBTII) return 1;
#ifdef _OPENMP
  if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
  real W[NBMAX * PMAX], Q[NBMAX * PMAX * PMAX], R[NBMAX * PMAX * PMAX];
  to_real(W, W_, (long)nb * p); to_real(Q, Q_, (long)nb * p * p); to_real(R, R_, (long)nb * p * p);
  real* obs_data = (real*)malloc(sizeof(real) * (size_t)n_obs * nb);
  real* obs_weight = (real*)malloc(sizeof(real) * (size_t)n_obs * nb * p);
  real* obs_var = (real*)malloc(sizeof(real) * (size_t)n_obs * nb);
  to_real(obs_data, obs_data_, (long)n_obs * nb); to_real(obs_weight, obs_weight_, (long)n_obs * nb * p);
  to_real(obs_var, obs_var_, (long)n_obs * nb);
#pragma omp parallel for schedule(static)
  for (long k = 0; k < B; ++k) {
    real th[16];
    to_real(th, theta + k * n_theta, n_theta);
    real mzy[NBMAX * PMAX], Szy[NBMAX * PMAX * PMAX], mz[NBMAX * PMAX], Sz[NBMAX * PMAX * PMAX];
    to_real(mzy, X0 + k * nb * p, (long)nb * p);
    memcpy(mz, mzy, sizeof(real) * nb * p);
    memset(Szy, 0, sizeof(real) * nb * p * p);
    memset(Sz, 0, sizeof(real) * nb * p * p);
    real ll_zy = 0.0, ll_z = 0.0;
    int i = 0;
    if (obs_ind[0] == 0) {
      for (int b = 0; b < nb; ++b) {
        real m = 0;
        for (int j = 0; j < p; ++j) m += obs_weight[b * p + j] * mzy[b * p + j];
        real res = obs_data[b] - m;
        ll_zy += logpdf(1, obs_var + b, &res);
      }
      i = 1;
    }
    for (int n = 0; n < N; ++n) {
      real t = step_time(t_min, t_max, n, N);
      int ic = i < n_obs ? i : n_obs - 1;
      if (n + 1 == obs_ind[ic]) {
        ll_zy += filter_step(model, interr, nb, p, W, Q, R, mzy, Szy, t, th, ic, obs_data, obs_weight, obs_var, 1);
        ++i;
      } else {
        ll_zy += filter_step(model, interr, nb, p, W, Q, R, mzy, Szy, t, th, -1, 0, 0, 0, 1);
      }
      ll_z += filter_step(model, interr, nb, p, W, Q, R, mz, Sz, t, th, -1, 0, 0, 0, 1);
    }
    out[k] = (double)(ll_zy - ll_z);
  }
  free(obs_data); free(obs_weight); free(obs_var);
  return 0;
}

#ifndef RODEO_LD   /* the extended-precision build only serves the dalton noise-floor measurement */
/* p x p LU solve with partial pivoting: X = A^{-1} Bm (Bm p x p row-major, columns are right-hand sides) */
static void lu_solve_pp(int p, const double* Ain, double* Bm) {
  double A[PMAX * PMAX];
  memcpy(A, Ain, sizeof(double) * p * p);
  for (int k = 0; k < p; ++k) {
    int piv = k;
    for (int r = k + 1; r < p; ++r) if (fabs(A[r * p + k]) > fabs(A[piv * p + k])) piv = r;
    if (piv != k)
      for (int c = 0; c < p; ++c) {
        double t = A[k * p + c]; A[k * p + c] = A[piv * p + c]; A[piv * p + c] = t;
        t = Bm[k * p + c]; Bm[k * p + c] = Bm[piv * p + c]; Bm[piv * p + c] = t;
      }
    for (int r = k + 1; r < p; ++r) {
      double l = A[r * p + k] / A[k * p + k];
      for (int c = k + 1; c < p; ++c) A[r * p + c] -= l * A[k * p + c];
      for (int c = 0; c < p; ++c) Bm[r * p + c] -= l * Bm[k * p + c];
    }
  }
  for (int k = p - 1; k >= 0; --k)
    for (int c = 0; c < p; ++c) {
      double s = Bm[k * p + c];
      for (int j = k + 1; j < p; ++j) s -= A[k * p + j] * Bm[j * p + c];
      Bm[k * p + c] = s / A[k * p + k];
    }
}

/* solve_mv: mean_out (B, N+1, nb, p), var_out (B, N+1, nb, p, p) */
int rodeo_oracle_solve_mv(int model, int interr, long B, int N, int nb, int p, int n_theta, double t_min, double t_max,
                          const double* W, const double* Q, const double* R, const double* X0, const double* theta,
                          double* mean_out, double* var_out, int n_threads) {
  if (nb > NBMAX || p > PMAX || interr == I_CHKREBTII) return 1;
#ifdef _OPENMP
  if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
  const long rm = (long)nb * p, rv = (long)nb * p * p;
#pragma omp parallel
  {
    double* mf = (double*)malloc(sizeof(double) * (N + 1) * rm);
    double* vf = (double*)malloc(sizeof(double) * (N + 1) * rv);
#pragma omp for schedule(static)
    for (long k = 0; k < B; ++k) {
      const double* th = theta + k * n_theta;
      double* ms = mean_out + k * (N + 1) * rm;
      double* vs = var_out + k * (N + 1) * rv;
      memcpy(mf, X0 + k * rm, sizeof(double) * rm);
      memset(vf, 0, sizeof(double) * rv);
      for (int n = 0; n < N; ++n) {
        memcpy(mf + (n + 1) * rm, mf + n * rm, sizeof(double) * rm);
        memcpy(vf + (n + 1) * rv, vf + n * rv, sizeof(double) * rv);
        filter_step(model, interr, nb, p, W, Q, R, mf + (n + 1) * rm, vf + (n + 1) * rv,
                    step_time(t_min, t_max, n, N), th, -1, 0, 0, 0, 0);
      }
      memcpy(ms, X0 + k * rm, sizeof(double) * rm);
      memset(vs, 0, sizeof(double) * rv);
      memcpy(ms + N * rm, mf + N * rm, sizeof(double) * rm);
      memcpy(vs + N * rv, vf + N * rv, sizeof(double) * rv);
      for (int n = N - 1; n >= 1; --n)
        for (int b = 0; b < nb; ++b) {
          const double *Qb = Q + b * p * p, *Rb = R + b * p * p;
          const double *muf = mf + n * rm + b * p, *Sf = vf + n * rv + b * p * p;
          double mp[PMAX], Sp[PMAX * PMAX], Gt[PMAX * PMAX];
          predict(p, Qb, Rb, muf, Sf, mp, Sp);
          /* G = (Sp^{-1} (Sf Q^T)^T)^T : solve Sp X = Q Sf, G = X^T */
          for (int i = 0; i < p; ++i)
            for (int j = 0; j < p; ++j) {
              double a = 0;
              for (int c = 0; c < p; ++c) a += Qb[i * p + c] * Sf[c * p + j];
              Gt[i * p + j] = a;
            }
          lu_solve_pp(p, Sp, Gt);
          const double *msn = ms + (n + 1) * rm + b * p, *vsn = vs + (n + 1) * rv + b * p * p;
          double* mso = ms + n * rm + b * p;
          double* vso = vs + n * rv + b * p * p;
          double GD[PMAX * PMAX];
          for (int i = 0; i < p; ++i) {
            double m = 0;
            for (int j = 0; j < p; ++j) m += Gt[j * p + i] * (msn[j] - mp[j]);
            mso[i] = muf[i] + m;
            for (int c = 0; c < p; ++c) {
              double a = 0;
              for (int j = 0; j < p; ++j) a += Gt[j * p + i] * (vsn[j * p + c] - Sp[j * p + c]);
              GD[i * p + c] = a;
            }
          }
          for (int i = 0; i < p; ++i)
            for (int j = 0; j < p; ++j) {
              double a = 0;
              for (int c = 0; c < p; ++c) a += GD[i * p + c] * Gt[c * p + j];
              vso[i * p + j] = Sf[i * p + j] + a;
            }
        }
    }
    free(mf); free(vf);
  }
  return 0;
}
#endif /* !RODEO_LD */
