// rodeo_b200 device library: register-resident small linear algebra for the block Kalman recursions.
//
// Everything here is a __device__ __forceinline__ template over the scalar type T (double primary, float
// secondary) and compile-time sizes, so that one (theta) filter lives entirely in registers for all n_steps.
// Symmetric matrices are stored packed (upper triangle, row-major): only p(p+1)/2 entries are carried.
//
// The recurrences follow the reference's covariance-form Kalman functions
//   predict      src/rodeo/kalmantv/standard.py:31-60
//   update       src/rodeo/kalmantv/standard.py:63-103   (+ rodeo.utils.solve_var, src/rodeo/utils.py:105-119)
//   forecast     src/rodeo/kalmantv/standard.py:308-336
//   _smooth, smooth_mv, smooth_sim, smooth_cond   src/rodeo/kalmantv/standard.py:160-255,339-371
//   multivariate_normal_logpdf                    src/rodeo/utils.py:60-78
// but are restated for a thread-per-theta execution model; no line of the reference is translated.
//
// This header is also compiled at run time by NVRTC for user-supplied ODE right-hand sides, so it must not
// include any host or CUDA toolkit header.
#pragma once

#define RD_DEV __device__ __forceinline__
#define RD_UNROLL _Pragma("unroll")
#define RD_UNROLL4 _Pragma("unroll 4")

namespace rodeo {

// ---- enums shared with the C ABI (include/rodeo_b200.h) -------------------------------------------------
enum : int { INTERR_KRAMER = 0, INTERR_CHKREBTII = 1, INTERR_SCHOBER = 2, INTERR_RODEO = 3 };
// structure of the prior transition matrix Q.  QK_DENSE_BATCH: dense, and (Q, R) are per-theta DEVICE arrays
// (B, n_block, p, p) that every thread loads into registers instead of reading the kernel-parameter bank
enum : int { QK_DENSE = 0, QK_UNIT_UPPER = 1, QK_DENSE_BATCH = 2 };

// packed index of a symmetric PxP matrix, requires i <= j
template <int P>
RD_DEV constexpr int sidx(int i, int j) { return i * P - (i * (i - 1)) / 2 + (j - i); }
template <int P>
RD_DEV constexpr int sym(int i, int j) { return i <= j ? sidx<P>(i, j) : sidx<P>(j, i); }
template <int P>
struct NSym { static constexpr int value = P * (P + 1) / 2; };

template <typename T> RD_DEV T rd_fma(T a, T b, T c);
template <> RD_DEV double rd_fma<double>(double a, double b, double c) { return fma(a, b, c); }
template <> RD_DEV float rd_fma<float>(float a, float b, float c) { return fmaf(a, b, c); }

// a - b that the compiler may not contract with a multiplication feeding a (or b) into an FMA: keeps the update
// residual f - W mu_p bitwise the same in every kernel that inlines the right-hand side (one thread per theta, one lane
// per (theta, block)), whatever else surrounds the expression
RD_DEV double sub_exact(double a, double b) { return __dsub_rn(a, b); }
RD_DEV float sub_exact(float a, float b) { return __fsub_rn(a, b); }

template <typename T> struct Lim;
template <> struct Lim<double> {
  RD_DEV static double inf() { return __longlong_as_double(0x7ff0000000000000LL); }
  RD_DEV static double nan() { return __longlong_as_double(0x7ff8000000000000LL); }
};
template <> struct Lim<float> {
  RD_DEV static float inf() { return __int_as_float(0x7f800000); }
  RD_DEV static float nan() { return __int_as_float(0x7fc00000); }
};

// ---- reciprocal ---------------------------------------------------------------------------------------------------
// 1/x to <= 1 ulp without the branchy slow path of the IEEE division sequence: MUFU.RCP64H seed (>= 20 bits) and one
// cubically convergent step (3 DFMA).  Zero, infinite, NaN and subnormal arguments yield inf/NaN garbage, exactly where the reference's
// LAPACK solve would have produced inf/NaN as well (a singular innovation variance).
RD_DEV double rcp(double x) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  // one third-order step: e = 1 - x y (|e| <= 2^-20 from the seed), y <- y (1 + e + e^2): error e^3 <= 2^-60
  const double e = fma(-x, y, 1.0);
  return fma(y, fma(e, e, e), y);
}
RD_DEV float rcp(float x) { return __frcp_rn(x); }

// sqrt(x) and 1/sqrt(x) together, x > 0 normal: MUFU.RSQ64H seed, one cubically convergent step for the reciprocal
// root, one correction step for the root (<= 1 ulp each); no slow path.  Used by the Cholesky-type factors of the
// sampling paths, which need both.
RD_DEV void sqrt_rsqrt(double x, double& s, double& rs) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double e = fma(-x * y, y, 1.0);                    // 1 - x y^2
  y = fma(y * fma(0.375, e, 0.5), e, y);                   // y (1 + e/2 + 3 e^2 / 8)
  double r = x * y;
  r = fma(fma(-r, r, x), 0.5 * y, r);
  s = r; rs = y;
}
RD_DEV void sqrt_rsqrt(float x, float& s, float& rs) { rs = rsqrtf(x); s = x * rs; }

// ---- mixed precision -----------------------------------------------------------------------------------------------
// The block MEANS (and everything that feeds the ODE right-hand side and the update residual) are always carried in
// double, also by the float32 instantiation: the residual f - W mu_p ~ 1e-3 is a difference of O(1) numbers and the
// mean recursion is iterated n_steps times, so a float mean drifts by ~n_steps * 6e-8 (3.9e-5 at N = 800, above the
// 1e-5 float32 budget).  Covariances, gains and log-density terms -- the bulk of the arithmetic -- stay in T.
// For T = double the two types coincide and nothing changes.
template <typename T> struct MeanOf { typedef double type; };

// ---- constants that live in the kernel-parameter constant bank ------------------------------------------
// Q, R, W are shared by every theta of a launch (reference layouts (nb,p,p), (nb,p,p), (nb,m,p)); passing them
// by value as a __grid_constant__ kernel parameter lets every DFMA take them as c[0][..] operands: they cost
// no registers and no loads inside the time loop.
template <typename T, int NB, int P, int M>
struct FilterConsts {
  T Q[NB][P][P];
  T R[NB][P * (P + 1) / 2];
  T W[NB][M][P];
};

// ---- predict ----------------------------------------------------------------------------------------------
// mu_p = Q mu ;  S_p = Q S Q^T + rs R      (mean_state == 0 in the solver, reference src/rodeo/solve.py:52)
// rs is a per-theta scale of the shared prior variance (1 when the prior is not batched): an IBM prior whose sigma
// is part of theta is R(theta) = sigma^2 R_1 (src/rodeo/prior/ibm.py:84-86), so the per-theta prior costs one
// register per block instead of p(p+1)/2.
template <typename T, int P, int QK, typename MT>
RD_DEV void predict(const T (&Q)[P][P], const T (&R)[P * (P + 1) / 2], T rs, const MT (&mu)[P],
                    const T (&S)[P * (P + 1) / 2], MT (&mup)[P], T (&Sp)[P * (P + 1) / 2]) {
  T A[P][P];  // A = Q S
  RD_UNROLL for (int i = 0; i < P; ++i) {
    if (QK == QK_UNIT_UPPER) {
      MT m = mu[i];
      RD_UNROLL for (int j = i + 1; j < P; ++j) m = rd_fma((MT)Q[i][j], mu[j], m);
      mup[i] = m;
      RD_UNROLL for (int k = 0; k < P; ++k) {
        T a = S[sym<P>(i, k)];
        RD_UNROLL for (int j = i + 1; j < P; ++j) a = rd_fma(Q[i][j], S[sym<P>(j, k)], a);
        A[i][k] = a;
      }
    } else {
      MT m = (MT)Q[i][0] * mu[0];
      RD_UNROLL for (int j = 1; j < P; ++j) m = rd_fma((MT)Q[i][j], mu[j], m);
      mup[i] = m;
      RD_UNROLL for (int k = 0; k < P; ++k) {
        T a = Q[i][0] * S[sym<P>(0, k)];
        RD_UNROLL for (int j = 1; j < P; ++j) a = rd_fma(Q[i][j], S[sym<P>(j, k)], a);
        A[i][k] = a;
      }
    }
  }
  RD_UNROLL for (int i = 0; i < P; ++i) {
    RD_UNROLL for (int j = i; j < P; ++j) {
      T s;
      if (QK == QK_UNIT_UPPER) {
        s = A[i][j];
        RD_UNROLL for (int k = j + 1; k < P; ++k) s = rd_fma(A[i][k], Q[j][k], s);
      } else {
        s = A[i][0] * Q[j][0];
        RD_UNROLL for (int k = 1; k < P; ++k) s = rd_fma(A[i][k], Q[j][k], s);
      }
      Sp[sidx<P>(i, j)] = rd_fma(rs, R[sidx<P>(i, j)], s);
    }
  }
}

// mean-only predict (used when the covariance recursion is not needed)
template <typename T, int P, int QK>
RD_DEV void predict_mean(const T (&Q)[P][P], const T (&mu)[P], T (&mup)[P]) {
  RD_UNROLL for (int i = 0; i < P; ++i) {
    T m = (QK == QK_UNIT_UPPER) ? mu[i] : Q[i][0] * mu[0];
    RD_UNROLL for (int j = (QK == QK_UNIT_UPPER ? i + 1 : 1); j < P; ++j) m = rd_fma(Q[i][j], mu[j], m);
    mup[i] = m;
  }
}

// ---- log-determinant accumulator -------------------------------------------------------------------------
// sum_k log(w_k) is accumulated as log(prod mantissa_k) + ln2 * sum exponent_k: one DMUL plus integer work
// per term instead of one FP64 log per term.  A negative or non-finite w poisons the product with NaN, which
// is what log(w) would have produced in the reference.
template <typename T> struct LogAcc;
template <> struct LogAcc<double> {
  double prod; int esum, sgn;
  RD_DEV void init() { prod = 1.0; esum = 0; sgn = 0; }
  // multiply in one kept term (w, or 1.0 for a dropped one); renorm() must follow within a few dozen terms
  RD_DEV void add(double wk) { prod *= wk; sgn |= __double2hiint(wk); }
  RD_DEV void renorm() {
    const int hi = __double2hiint(prod), lo = __double2loint(prod);
    const int e = (hi >> 20) & 0x7ff;
    const bool fin = (unsigned)(e - 1) < 0x7feu;            // normal number: not 0 / subnormal / inf / nan
    esum += fin ? e - 1023 : 0;
    prod = fin ? __hiloint2double((hi & 0x800fffff) | 0x3ff00000, lo) : prod;
  }
  // sum of log(w) over the kept terms; a negative kept w makes it nan, as log(w) would have in the reference
  RD_DEV double value() const {
    const double v = log(fabs(prod)) + 0.6931471805599453094 * (double)esum;
    return sgn < 0 ? Lim<double>::nan() : v;
  }
};
template <> struct LogAcc<float> {
  float prod; int esum, sgn;
  RD_DEV void init() { prod = 1.0f; esum = 0; sgn = 0; }
  RD_DEV void add(float wk) { prod *= wk; sgn |= __float_as_int(wk); }
  RD_DEV void renorm() {
    const int b = __float_as_int(prod);
    const int e = (b >> 23) & 0xff;
    const bool fin = (unsigned)(e - 1) < 0xfeu;
    esum += fin ? e - 127 : 0;
    prod = fin ? __int_as_float((b & 0x807fffff) | 0x3f800000) : prod;
  }
  // in double: the exponent sum alone reaches ~1e4 over a solve, where a float carries 1e-3
  RD_DEV double value() const {
    const double v = (double)logf(fabsf(prod)) + 0.6931471805599453094 * (double)esum;
    return sgn < 0 ? Lim<double>::nan() : v;
  }
};

// Gaussian log-density accumulator: logdens = -1/2 (quad + logdet) - 1/2 cnt log(2 pi), with the reference's
// absolute eigenvalue cut-off |w| > 1e-8 (src/rodeo/utils.py:74, jnp.isclose default atol) applied per term.
// The caller renormalises the running product once per time step (ld.renorm()).
// The quadratic form is summed, and the value returned, in the mean type (double also for T = float): the two
// log-densities dalton subtracts are ~1e4 each over 800 steps, their difference ~1e1.
template <typename T>
struct LogPdfAcc {
  typedef typename MeanOf<T>::type MT;
  MT quad; LogAcc<T> ld; int cnt;
  RD_DEV void init() { quad = MT(0); ld.init(); cnt = 0; }
  // one eigen-direction: eigenvalue w, projected residual z
  RD_DEV void term(T w, MT z, T rw /* = 1/w */) {
    const bool keep = !(fabs(w) <= T(1e-8));   // nan counts as kept, as ~isclose(nan, 0) does
#ifdef RODEO_LOGPDF_SELECT
    quad = rd_fma(z * z, (MT)(keep ? rw : T(0)), quad);
    ld.add(keep ? w : T(1));
    cnt += keep ? 1 : 0;
#else
    // three predicated instructions instead of four 32-bit selects feeding unconditional ones
    if (keep) {
      quad = rd_fma(z * z, (MT)rw, quad);
      ld.add(w);
      ++cnt;
    }
#endif
  }
  RD_DEV MT value() const {
    return MT(-0.5) * (quad + (MT)ld.value()) - MT(0.5) * MT(1.8378770664093454836) * (MT)cnt;
  }
};

// The same accumulator split into per-block partial sums (quadratic form, mantissa product) and one set of integer
// counters (exponent sum, kept-term count, sign bits).  dalton sums the log-densities of all blocks of a theta: keeping
// one partial sum PER BLOCK and combining them in block order at the very end (combine_logpdf) makes the result
// independent of how the blocks are laid out over threads -- a thread that owns every block of a theta and the lanes of
// a (theta, block) warp add exactly the same numbers in exactly the same order, so both launch geometries return
// bitwise the same log-likelihood.  The integer counters are exact, hence order-free, and can be shared.
struct LogPdfCtl {
  int esum, cnt, sgn;
  RD_DEV void init() { esum = 0; cnt = 0; sgn = 0; }
};
template <typename T>
struct LogPdfPart {
  typedef typename MeanOf<T>::type MT;
  MT quad; T prod;
  RD_DEV void init() { quad = MT(0); prod = T(1); }
  // fold the exponent of the running product into ctl.esum (once per time step)
  RD_DEV void renorm(LogPdfCtl& c) {
    LogAcc<T> l; l.prod = prod; l.esum = c.esum; l.sgn = 0;
    l.renorm();
    prod = l.prod; c.esum = l.esum;
  }
};
// what update() / logpdf_terms() see: one block's partial sums plus the shared counters
template <typename T>
struct LogPdfPartRef {
  typedef typename MeanOf<T>::type MT;
  LogPdfPart<T>& p; LogPdfCtl& c;
  RD_DEV void term(T w, MT z, T rw) {
    const bool keep = !(fabs(w) <= T(1e-8));   // nan counts as kept, as ~isclose(nan, 0) does
    if (keep) {
      p.quad = rd_fma(z * z, (MT)rw, p.quad);
      p.prod *= w;
      c.sgn |= sizeof(T) == 8 ? __double2hiint((double)w) : __float_as_int((float)w);
      ++c.cnt;
    }
  }
};
// log-density of all blocks: -1/2 (sum_b quad_b + log prod_b |prod_b| + ln2 esum) - 1/2 cnt log(2 pi), blocks in order
template <typename T, int NB>
RD_DEV typename MeanOf<T>::type combine_logpdf(const LogPdfPart<T> (&part)[NB], const LogPdfCtl& ctl) {
  typedef typename MeanOf<T>::type MT;
  // (the parts are renormalised first: a caller may fold exponents only every few steps)
  MT quad = part[0].quad;
  LogAcc<T> l; l.prod = part[0].prod; l.esum = ctl.esum; l.sgn = ctl.sgn;
  l.renorm();
  RD_UNROLL for (int b = 1; b < NB; ++b) {
    quad += part[b].quad;
    LogAcc<T> lb; lb.prod = part[b].prod; lb.esum = 0; lb.sgn = 0;
    lb.renorm();
    l.prod *= lb.prod; l.esum += lb.esum;
  }
  l.renorm();
  return MT(-0.5) * (quad + (MT)l.value()) - MT(0.5) * MT(1.8378770664093454836) * (MT)ctl.cnt;
}

// all blocks of one theta held by one thread
template <typename T, int NB>
struct LogPdfParts {
  typedef typename MeanOf<T>::type MT;
  LogPdfPart<T> part[NB]; LogPdfCtl ctl;
  RD_DEV void init() { ctl.init(); RD_UNROLL for (int b = 0; b < NB; ++b) part[b].init(); }
  RD_DEV void renorm() { RD_UNROLL for (int b = 0; b < NB; ++b) part[b].renorm(ctl); }
  RD_DEV MT value() const { return combine_logpdf<T, NB>(part, ctl); }
};
// accumulator of block b: the single accumulator itself, or a view of block b's partial sums
template <typename T> RD_DEV LogPdfAcc<T>& acc_at(LogPdfAcc<T>& a, int) { return a; }
template <typename T, int NB>
RD_DEV LogPdfPartRef<T> acc_at(LogPdfParts<T, NB>& a, int b) { return LogPdfPartRef<T>{a.part[b], a.ctl}; }
template <typename T> RD_DEV LogPdfPartRef<T>& acc_at(LogPdfPartRef<T>& a, int) { return a; }

// ---- tiny dense solves -------------------------------------------------------------------------------------
// Solve S X = Bm for X (MM x P right-hand sides stored as rows r of Bm[r][:]) by Gaussian elimination with
// partial pivoting, the algorithm behind jnp.linalg.solve (LAPACK getrf/getrs) in rodeo.utils.solve_var.
// S is symmetric, given packed; it is expanded to a full MM x MM working copy.
template <typename T, int MM, int P>
RD_DEV void solve_small(const T (&Ss)[MM * (MM + 1) / 2], T (&Bm)[MM][P]) {
  if (MM == 1) {
    T r = rcp(Ss[0]);
    RD_UNROLL for (int i = 0; i < P; ++i) Bm[0][i] *= r;
    return;
  }
  T A[MM][MM];
  RD_UNROLL for (int r = 0; r < MM; ++r)
    RD_UNROLL for (int c = 0; c < MM; ++c) A[r][c] = Ss[sym<MM>(r, c)];
  RD_UNROLL for (int k = 0; k < MM; ++k) {
    // pivot search + row swap by selects (no divergent control flow)
    RD_UNROLL for (int r = k + 1; r < MM; ++r) {
      bool sw = fabs(A[r][k]) > fabs(A[k][k]);
      RD_UNROLL for (int c = 0; c < MM; ++c) {
        T a = A[k][c], b = A[r][c];
        A[k][c] = sw ? b : a; A[r][c] = sw ? a : b;
      }
      RD_UNROLL for (int c = 0; c < P; ++c) {
        T a = Bm[k][c], b = Bm[r][c];
        Bm[k][c] = sw ? b : a; Bm[r][c] = sw ? a : b;
      }
    }
    T rp = rcp(A[k][k]);
    RD_UNROLL for (int r = k + 1; r < MM; ++r) {
      T l = A[r][k] * rp;
      RD_UNROLL for (int c = k + 1; c < MM; ++c) A[r][c] = rd_fma(-l, A[k][c], A[r][c]);
      RD_UNROLL for (int c = 0; c < P; ++c) Bm[r][c] = rd_fma(-l, Bm[k][c], Bm[r][c]);
    }
  }
  RD_UNROLL for (int k = MM - 1; k >= 0; --k) {
    T rp = rcp(A[k][k]);
    RD_UNROLL for (int c = 0; c < P; ++c) {
      T s = Bm[k][c];
      RD_UNROLL for (int j = k + 1; j < MM; ++j) s = rd_fma(-A[k][j], Bm[j][c], s);
      Bm[k][c] = s * rp;
    }
  }
}

// Symmetric 2x2 eigen-decomposition [[a,b],[b,c]] -> rt1 >= rt2 (by absolute value as LAPACK dlaev2 orders
// them: |rt1| >= |rt2|), unit eigenvector (cs1, sn1) of rt1; the eigenvector of rt2 is (-sn1, cs1).
template <typename T>
RD_DEV void eig2(T a, T b, T c, T& rt1, T& rt2, T& cs1, T& sn1) {
  T sm = a + c, df = a - c, adf = fabs(df), tb = b + b, ab = fabs(tb);
  T acmx = fabs(a) > fabs(c) ? a : c, acmn = fabs(a) > fabs(c) ? c : a;
  T rt;
  if (adf > ab) { T q = ab / adf; rt = adf * sqrt(T(1) + q * q); }
  else if (adf < ab) { T q = adf / ab; rt = ab * sqrt(T(1) + q * q); }
  else rt = ab * sqrt(T(2));
  int sgn1;
  if (sm < T(0)) { rt1 = T(0.5) * (sm - rt); sgn1 = -1; rt2 = (acmx / rt1) * acmn - (b / rt1) * b; }
  else if (sm > T(0)) { rt1 = T(0.5) * (sm + rt); sgn1 = 1; rt2 = (acmx / rt1) * acmn - (b / rt1) * b; }
  else { rt1 = T(0.5) * rt; rt2 = T(-0.5) * rt; sgn1 = 1; }
  int sgn2; T cs;
  if (df >= T(0)) { cs = df + rt; sgn2 = 1; } else { cs = df - rt; sgn2 = -1; }
  if (fabs(cs) > ab) { T ct = -tb / cs; sn1 = T(1) / sqrt(T(1) + ct * ct); cs1 = ct * sn1; }
  else if (ab == T(0)) { cs1 = T(1); sn1 = T(0); }
  else { T tn = -cs / tb; cs1 = T(1) / sqrt(T(1) + tn * tn); sn1 = tn * cs1; }
  if (sgn1 == sgn2) { T tn = cs1; cs1 = -sn1; sn1 = tn; }
}

// Cyclic Jacobi eigen-decomposition of a symmetric MM x MM matrix (MM >= 3 fall-back for the log-pdf).
// On exit A's diagonal holds the eigenvalues and the columns of V the eigenvectors.
template <typename T, int MM>
RD_DEV void eig_jacobi(T (&A)[MM][MM], T (&V)[MM][MM]) {
  RD_UNROLL for (int i = 0; i < MM; ++i)
    RD_UNROLL for (int j = 0; j < MM; ++j) V[i][j] = (i == j) ? T(1) : T(0);
  for (int sweep = 0; sweep < 12; ++sweep) {
    T off = T(0);
    RD_UNROLL for (int p = 0; p < MM; ++p)
      RD_UNROLL for (int q = p + 1; q < MM; ++q) off += A[p][q] * A[p][q];
    if (off == T(0)) break;
    RD_UNROLL for (int p = 0; p < MM; ++p) {
      RD_UNROLL for (int q = p + 1; q < MM; ++q) {
        T apq = A[p][q];
        if (apq != T(0)) {
          T tau = (A[q][q] - A[p][p]) / (T(2) * apq);
          T t = (tau >= T(0) ? T(1) : T(-1)) / (fabs(tau) + sqrt(T(1) + tau * tau));
          T c = T(1) / sqrt(T(1) + t * t), s = t * c;
          RD_UNROLL for (int k = 0; k < MM; ++k) {
            T akp = A[k][p], akq = A[k][q];
            A[k][p] = c * akp - s * akq; A[k][q] = s * akp + c * akq;
          }
          RD_UNROLL for (int k = 0; k < MM; ++k) {
            T apk = A[p][k], aqk = A[q][k];
            A[p][k] = c * apk - s * aqk; A[q][k] = s * apk + c * aqk;
          }
          RD_UNROLL for (int k = 0; k < MM; ++k) {
            T vkp = V[k][p], vkq = V[k][q];
            V[k][p] = c * vkp - s * vkq; V[k][q] = s * vkp + c * vkq;
          }
        }
      }
    }
  }
}

// log N(x; mu, S) contribution(s) into `acc`, S symmetric MM x MM packed, res = x - mu.
template <typename T, int MM, class ACC>
RD_DEV void logpdf_terms(const T (&Ss)[MM * (MM + 1) / 2], const T (&res)[MM], ACC& acc) {
  if (MM == 1) {
    acc.term(Ss[0], res[0], rcp(Ss[0]));
  } else if (MM == 2) {
    T w1, w2, c, s;
    eig2<T>(Ss[0], Ss[1], Ss[2], w1, w2, c, s);
    T z1 = c * res[0] + s * res[1], z2 = -s * res[0] + c * res[1];
    acc.term(w1, z1, rcp(w1));
    acc.term(w2, z2, rcp(w2));
  } else {
    T A[MM][MM], V[MM][MM];
    RD_UNROLL for (int r = 0; r < MM; ++r)
      RD_UNROLL for (int c = 0; c < MM; ++c) A[r][c] = Ss[sym<MM>(r, c)];
    eig_jacobi<T, MM>(A, V);
    RD_UNROLL for (int k = 0; k < MM; ++k) {
      T z = T(0);
      RD_UNROLL for (int r = 0; r < MM; ++r) z = rd_fma(V[r][k], res[r], z);
      acc.term(A[k][k], z, rcp(A[k][k]));
    }
  }
}

// ---- update (with optional forecast log-density) --------------------------------------------------------------
// Measurement model  z = wm x + d + N(0, V),  observed value xm, MM rows; `res` = xm - (wm mu_p + d) is supplied by
// the caller (see Fwd::interrogate for why the solver never forms d).
//   S = wm S_p wm^T + V ;  K = S_p wm^T S^{-1} ;  mu_f = mu_p + K res ;  S_f = S_p - K (wm S_p)
// WITH_LOGPDF adds log N(xm; mu_z, S) to `acc` (reference fenrir._forecast_update, fenrir.py:40-81).
// In-place: mu, S hold the predicted moments on entry and the filtered moments on exit.
template <typename T, int P, int MM, bool WITH_LOGPDF, typename MT, class ACC>
RD_DEV void update(MT (&mu)[P], T (&S)[P * (P + 1) / 2], const T (&wm)[MM][P], const MT (&res)[MM],
                   const T (&V)[MM * (MM + 1) / 2], ACC& acc) {
  T v[MM][P];    // v[r] = S_p wm[r]^T   (== (wm S_p)[r] by symmetry)
  T Kt[MM][P];
  T Sm[MM * (MM + 1) / 2];
  RD_UNROLL for (int r = 0; r < MM; ++r) {
    RD_UNROLL for (int i = 0; i < P; ++i) {
      T a = S[sym<P>(i, 0)] * wm[r][0];
      RD_UNROLL for (int j = 1; j < P; ++j) a = rd_fma(S[sym<P>(i, j)], wm[r][j], a);
      v[r][i] = a; Kt[r][i] = a;
    }
  }
  RD_UNROLL for (int r = 0; r < MM; ++r)
    RD_UNROLL for (int s = r; s < MM; ++s) {
      T a = V[sidx<MM>(r, s)];
      RD_UNROLL for (int i = 0; i < P; ++i) a = rd_fma(wm[r][i], v[s][i], a);
      Sm[sidx<MM>(r, s)] = a;
    }
  if (WITH_LOGPDF) {
    T rt[MM];
    RD_UNROLL for (int r = 0; r < MM; ++r) rt[r] = (T)res[r];
    logpdf_terms<T, MM, ACC>(Sm, rt, acc);
  }
  solve_small<T, MM, P>(Sm, Kt);
  RD_UNROLL for (int i = 0; i < P; ++i) {
    MT m = mu[i];
    RD_UNROLL for (int r = 0; r < MM; ++r) m = rd_fma((MT)Kt[r][i], res[r], m);
    mu[i] = m;
  }
  RD_UNROLL for (int i = 0; i < P; ++i)
    RD_UNROLL for (int j = i; j < P; ++j) {
      T s = S[sidx<P>(i, j)];
      RD_UNROLL for (int r = 0; r < MM; ++r) s = rd_fma(-Kt[r][i], v[r][j], s);
      S[sidx<P>(i, j)] = s;
    }
}

// Scalar-measurement update for a row of the form  wm = e_WK - [jl_0 .. jl_{JC-1}, 0, ...]  (what interrogate_kramer
// produces for W = e_WK and a right-hand side that reads the leading JC state columns; jl == 0 for the other
// interrogations).  Same arithmetic as update<T,P,1,...> with the multiplications by the structural 0 / 1 entries
// removed (those are exact, so the results are bitwise those of the general row).
template <typename T, int P, int JC, int WK, bool WITH_LOGPDF, bool HAS_J, bool HAS_V = true, typename MT = T,
          class ACC = LogPdfAcc<T>>
RD_DEV void update_unit_row(MT (&mu)[P], T (&S)[P * (P + 1) / 2], const T (&jl)[JC], MT res, T V, ACC& acc) {
  T v[P];
  RD_UNROLL for (int i = 0; i < P; ++i) {
    T a = S[sym<P>(i, WK)];
    if (HAS_J) { RD_UNROLL for (int j = 0; j < JC; ++j) a = rd_fma(-jl[j], S[sym<P>(i, j)], a); }
    v[i] = a;
  }
  T Sm = HAS_V ? V + v[WK] : v[WK];
  if (HAS_J) { RD_UNROLL for (int j = 0; j < JC; ++j) Sm = rd_fma(-jl[j], v[j], Sm); }
  const T rS = rcp(Sm);
  if (WITH_LOGPDF) acc.term(Sm, res, rS);
  const MT g = res * (MT)rS;
  RD_UNROLL for (int i = 0; i < P; ++i) mu[i] = rd_fma((MT)v[i], g, mu[i]);
  RD_UNROLL for (int i = 0; i < P; ++i) {
    const T k = v[i] * rS;
    RD_UNROLL for (int j = i; j < P; ++j) S[sidx<P>(i, j)] = rd_fma(-k, v[j], S[sidx<P>(i, j)]);
  }
}

// ---- observation-augmented update, sequential form ---------------------------------------------------------------------
// dalton's zy_update (src/rodeo/inference/dalton.py:136-149) stacks the ODE row and the observation row of a block into
// one 2-row measurement with block-diagonal noise diag(V1, V2) and runs forecast -> log-pdf -> update on it.  Because the
// two noises are uncorrelated, conditioning on the rows one after the other gives the same filtered moments, and by the
// chain rule the joint forecast density factorises:  N([x1, x2]; mu_z, S) = N(x1; m1, a) N(x2; m2|1, c'), where
// a = S_11 and c' = S_22 - S_12^2 / S_11 is the Schur complement (det S = a c').  Two scalar updates cost ~2 plain steps;
// the stacked form costs a symmetric 2x2 eigen-decomposition (5 IEEE divisions and 2 square roots in one dependency chain),
// a pivoted 2x2 solve and a rank-2 downdate -- measured 3,700 cycles per observation step for a lone warp against 300 for
// a plain step.
// The reference's log-pdf drops eigen-directions of the joint S with |w| <= 1e-8 (src/rodeo/utils.py:74).  Both
// eigenvalues exceed tau = 1e-8 iff S - tau I is positive definite, i.e. a > tau and (a - tau)(c - tau) > b^2, which in
// terms of c' reads (a - tau)(c' - tau) > b^2 tau / a; then nothing is dropped and the two scalar terms are exact.
// Otherwise (tiny prior variances) the joint S = [[a, b], [b, c' + b^2 / a]] and the joint residual are rebuilt and go
// through the eigen-decomposition exactly as before.  The filtered moments are the sequential ones either way.
template <typename T, int P, int JC, int WK, bool HAS_J>
struct UnitRow {          // w = e_WK - [jl_0 .. jl_{JC-1}, 0, ...]
  const T (&jl)[JC];
  RD_DEV T Sdot(const T (&S)[P * (P + 1) / 2], int i) const {
    T a = S[sym<P>(i, WK)];
    if (HAS_J) { RD_UNROLL for (int j = 0; j < JC; ++j) a = rd_fma(-jl[j], S[sym<P>(i, j)], a); }
    return a;
  }
  RD_DEV T dot(const T (&v)[P], T base) const {      // base + w . v
    T a = base + v[WK];
    if (HAS_J) { RD_UNROLL for (int j = 0; j < JC; ++j) a = rd_fma(-jl[j], v[j], a); }
    return a;
  }
};
template <typename T, int P>
struct DenseRow {
  const T (&w)[P];
  RD_DEV T Sdot(const T (&S)[P * (P + 1) / 2], int i) const {
    T a = S[sym<P>(i, 0)] * w[0];
    RD_UNROLL for (int j = 1; j < P; ++j) a = rd_fma(S[sym<P>(i, j)], w[j], a);
    return a;
  }
  RD_DEV T dot(const T (&v)[P], T base) const {
    T a = base;
    RD_UNROLL for (int j = 0; j < P; ++j) a = rd_fma(w[j], v[j], a);
    return a;
  }
  template <typename MT>
  RD_DEV MT resid(MT x, const MT (&mu)[P]) const {      // x - w . mu in the mean type
    MT a = x;
    RD_UNROLL for (int j = 0; j < P; ++j) a = rd_fma(-(MT)w[j], mu[j], a);
    return a;
  }
};

// res1: residual x1 - (w1 mu_p + d1) of row 1 at the PREDICTED mean; x2: observed value of row 2 minus its offset (its
// residual is formed from the mean after row 1, in the mean type); V1 / V2 the noise variances
template <typename T, int P, bool WITH_LOGPDF, bool HAS_V1, class ROW1, class ROW2, typename MT, class ACC>
RD_DEV void update_two_rows(MT (&mu)[P], T (&S)[P * (P + 1) / 2], const ROW1& row1, MT res1, T V1, const ROW2& row2,
                            MT x2, T V2, ACC& acc) {
  T v[P];
  RD_UNROLL for (int i = 0; i < P; ++i) v[i] = row1.Sdot(S, i);
  const T a = row1.dot(v, HAS_V1 ? V1 : T(0));
  const T b = row2.dot(v, T(0));                        // S_12 of the stacked forecast variance
  const T ra = rcp(a);
  {
    const MT g = res1 * (MT)ra;
    RD_UNROLL for (int i = 0; i < P; ++i) mu[i] = rd_fma((MT)v[i], g, mu[i]);
    RD_UNROLL for (int i = 0; i < P; ++i) {
      const T k = v[i] * ra;
      RD_UNROLL for (int j = i; j < P; ++j) S[sidx<P>(i, j)] = rd_fma(-k, v[j], S[sidx<P>(i, j)]);
    }
  }
  const T bra = b * ra;
  const MT res2c = row2.resid(x2, mu);                   // x2 - (w2 mu_1 + d2): residual at the mean after row 1
  T u[P];
  RD_UNROLL for (int i = 0; i < P; ++i) u[i] = row2.Sdot(S, i);
  const T c = row2.dot(u, V2);                           // Schur complement S_22 - S_12^2 / S_11
  const T rc = rcp(c);
  {
    const MT g = res2c * (MT)rc;
    RD_UNROLL for (int i = 0; i < P; ++i) mu[i] = rd_fma((MT)u[i], g, mu[i]);
    RD_UNROLL for (int i = 0; i < P; ++i) {
      const T k = u[i] * rc;
      RD_UNROLL for (int j = i; j < P; ++j) S[sidx<P>(i, j)] = rd_fma(-k, u[j], S[sidx<P>(i, j)]);
    }
  }
  if (WITH_LOGPDF) {
    const T tau = T(1e-8);
    if (a > tau && (a - tau) * (c - tau) > b * bra * tau) {
      acc.term(a, res1, ra);
      acc.term(c, res2c, rc);
    } else {
      T Sj[3] = {a, b, rd_fma(b, bra, c)};
      T rj[2] = {(T)res1, (T)rd_fma((MT)bra, res1, res2c)};      // row 2's residual at the predicted mean
      logpdf_terms<T, 2>(Sj, rj, acc);
    }
  }
}

// ---- SPD factorisations ---------------------------------------------------------------------------------------
// L D L^T of a symmetric positive definite matrix (unit lower L packed into Lo[i][j], j<i; reciprocals of D).
template <typename T, int P>
RD_DEV void ldlt(const T (&S)[P * (P + 1) / 2], T (&L)[P][P], T (&rD)[P]) {
  T D[P];
  RD_UNROLL for (int j = 0; j < P; ++j) {
    T dj = S[sidx<P>(j, j)];
    RD_UNROLL for (int k = 0; k < j; ++k) dj = rd_fma(-L[j][k] * L[j][k], D[k], dj);
    D[j] = dj; rD[j] = rcp(dj);
    RD_UNROLL for (int i = j + 1; i < P; ++i) {
      T s = S[sidx<P>(j, i)];
      RD_UNROLL for (int k = 0; k < j; ++k) s = rd_fma(-L[i][k] * L[j][k], D[k], s);
      L[i][j] = s * rD[j];
    }
  }
}

// x <- S^{-1} x given the factorisation
template <typename T, int P>
RD_DEV void ldlt_solve(const T (&L)[P][P], const T (&rD)[P], T (&x)[P]) {
  RD_UNROLL for (int i = 1; i < P; ++i)
    RD_UNROLL for (int k = 0; k < i; ++k) x[i] = rd_fma(-L[i][k], x[k], x[i]);
  RD_UNROLL for (int i = 0; i < P; ++i) x[i] *= rD[i];
  RD_UNROLL for (int i = P - 2; i >= 0; --i)
    RD_UNROLL for (int k = i + 1; k < P; ++k) x[i] = rd_fma(-L[k][i], x[k], x[i]);
}

// Guarded Cholesky-type factor A (lower triangular, A A^T = C) of a symmetric positive SEMI-definite matrix:
// a pivot d_j <= 0 produces a zero column.  The smoothing covariances of a noise-free measurement model are
// rank deficient, which is why the reference samples with an SVD factor (src/rodeo/solve.py:179,196-198); any
// factor gives the same distribution.  Mirrors oracle psd_factor(method="ldl").
template <typename T, int P>
RD_DEV void psd_factor(const T (&C)[P * (P + 1) / 2], T (&A)[P][P]) {
  RD_UNROLL for (int j = 0; j < P; ++j) {
    T d = C[sidx<P>(j, j)];
    RD_UNROLL for (int k = 0; k < j; ++k) d = rd_fma(-A[j][k], A[j][k], d);
    bool pos = d > T(0);
    T dj, rdj;
    sqrt_rsqrt(pos ? d : T(1), dj, rdj);
    A[j][j] = pos ? dj : T(0);
    RD_UNROLL for (int i = j + 1; i < P; ++i) {
      T s = C[sidx<P>(j, i)];
      RD_UNROLL for (int k = 0; k < j; ++k) s = rd_fma(-A[i][k], A[j][k], s);
      A[i][j] = pos ? s * rdj : T(0);
    }
  }
}

// ---- smoother gain ---------------------------------------------------------------------------------------------
// G = S_f Q^T S_p^{-1}   (reference _smooth, standard.py:175-176: solve_var(S_p, (S_f Q^T)^T)^T).
// S_p is symmetric positive definite (Q S_f Q^T + R with R > 0): an un-pivoted LDL^T replaces the reference's
// pivoted LU; both are backward stable and agree to rounding.
// Also returns Ct = S_f Q^T (full P x P), needed by smooth_sim / smooth_cond.
template <typename T, int P, int QK>
RD_DEV void smooth_gain(const T (&Q)[P][P], const T (&Sf)[P * (P + 1) / 2], const T (&Sp)[P * (P + 1) / 2],
                        T (&G)[P][P], T (&Ct)[P][P]) {
  // Ct[i][j] = sum_k Sf[i][k] Q[j][k]
  RD_UNROLL for (int i = 0; i < P; ++i)
    RD_UNROLL for (int j = 0; j < P; ++j) {
      T a;
      if (QK == QK_UNIT_UPPER) {
        a = Sf[sym<P>(i, j)];
        RD_UNROLL for (int k = j + 1; k < P; ++k) a = rd_fma(Sf[sym<P>(i, k)], Q[j][k], a);
      } else {
        a = Sf[sym<P>(i, 0)] * Q[j][0];
        RD_UNROLL for (int k = 1; k < P; ++k) a = rd_fma(Sf[sym<P>(i, k)], Q[j][k], a);
      }
      Ct[i][j] = a;
    }
  T L[P][P], rD[P];
  ldlt<T, P>(Sp, L, rD);
  // row i of G solves S_p g = Ct[i]^T
  RD_UNROLL for (int i = 0; i < P; ++i) {
    T x[P];
    RD_UNROLL for (int j = 0; j < P; ++j) x[j] = Ct[i][j];
    ldlt_solve<T, P>(L, rD, x);
    RD_UNROLL for (int j = 0; j < P; ++j) G[i][j] = x[j];
  }
}

// sym result of  base + sgn * G D G^T  with D symmetric packed  (smooth_mv, standard.py:215-216)
template <typename T, int P>
RD_DEV void add_GDGt(const T (&G)[P][P], const T (&D)[P * (P + 1) / 2], T (&out)[P * (P + 1) / 2]) {
  T GD[P][P];
  RD_UNROLL for (int i = 0; i < P; ++i)
    RD_UNROLL for (int k = 0; k < P; ++k) {
      T a = G[i][0] * D[sym<P>(0, k)];
      RD_UNROLL for (int j = 1; j < P; ++j) a = rd_fma(G[i][j], D[sym<P>(j, k)], a);
      GD[i][k] = a;
    }
  RD_UNROLL for (int i = 0; i < P; ++i)
    RD_UNROLL for (int j = i; j < P; ++j) {
      T a = out[sidx<P>(i, j)];
      RD_UNROLL for (int k = 0; k < P; ++k) a = rd_fma(GD[i][k], G[j][k], a);
      out[sidx<P>(i, j)] = a;
    }
}

// C = S_f - G Ct^T   (smooth_sim / smooth_cond variance, standard.py:253-254, 370)
template <typename T, int P>
RD_DEV void cond_var(const T (&Sf)[P * (P + 1) / 2], const T (&G)[P][P], const T (&Ct)[P][P],
                     T (&C)[P * (P + 1) / 2]) {
  RD_UNROLL for (int i = 0; i < P; ++i)
    RD_UNROLL for (int j = i; j < P; ++j) {
      T a = Sf[sidx<P>(i, j)];
      RD_UNROLL for (int k = 0; k < P; ++k) a = rd_fma(-G[i][k], Ct[j][k], a);
      C[sidx<P>(i, j)] = a;
    }
}


// ====================================================================================================================
// Square-root Kalman family (reference src/rodeo/kalmantv/square_root.py): variances are carried as lower-triangular
// factors L (var = L L^T), packed row-major: lidx(i, j), j <= i.
// ====================================================================================================================
RD_DEV constexpr int lidx(int i, int j) { return i * (i + 1) / 2 + j; }
template <typename T, int P>
RD_DEV T lget(const T (&L)[P * (P + 1) / 2], int i, int j) { return j <= i ? L[lidx(i, j)] : T(0); }

// L' (lower, packed) with L' L'^T = A^T A for a ROWS x P matrix A: the R factor of a Householder QR, transposed.
// This is rodeo.utils.add_sqrt (src/rodeo/utils.py:10-24: R^T of qr(vstack([sqrt_A^T, sqrt_B^T]))) with the stacked
// matrix built by the caller.  Only L' L'^T is meaningful across implementations (QR leaves the signs of R's diagonal
// free).  A is destroyed.
template <typename T, int ROWS, int P>
RD_DEV void qr_lower(T (&A)[ROWS][P], T (&Lout)[P * (P + 1) / 2]) {
  // Householder steps with LAPACK's conventions (dgeqr2 / dlarfg: beta = -sign(x_k) |x|, sign(0) = +, and NO reflection
  // when the sub-diagonal part is exactly zero), so that the signs of R's diagonal are those jnp.linalg.qr /
  // np.linalg.qr produce.  They only matter where the reference itself uses a factor un-squared: the
  // interrogate_chkrebtii "square-root" draw x = mean + (W L) z (src/rodeo/interrogate.py:42-45).
  RD_UNROLL for (int k = 0; k < P; ++k) {
    T xn2 = T(0);
    RD_UNROLL for (int r = k + 1; r < ROWS; ++r) xn2 = rd_fma(A[r][k], A[r][k], xn2);
    const T x0 = A[k][k];
    const bool reflect = xn2 > T(0);
    const T nrm = sqrt(rd_fma(x0, x0, xn2));
    const T beta = x0 >= T(0) ? -nrm : nrm;
    // v = x - beta e_k ; H = I - 2 v v^T / (v^T v) ;  v^T v = 2 (|x|^2 - beta x0)
    const T v0 = x0 - beta;
    const T vtv = T(2) * (rd_fma(x0, x0, xn2) - beta * x0);
    const T scale = reflect ? T(2) * rcp(vtv) : T(0);
    RD_UNROLL for (int c = k + 1; c < P; ++c) {
      T dot = v0 * A[k][c];
      RD_UNROLL for (int r = k + 1; r < ROWS; ++r) dot = rd_fma(A[r][k], A[r][c], dot);
      const T s = scale * dot;
      A[k][c] = rd_fma(-s, v0, A[k][c]);
      RD_UNROLL for (int r = k + 1; r < ROWS; ++r) A[r][c] = rd_fma(-s, A[r][k], A[r][c]);
    }
    Lout[lidx(k, k)] = reflect ? beta : x0;
    RD_UNROLL for (int c = k + 1; c < P; ++c) Lout[lidx(c, k)] = A[k][c];   // R[k][c] -> L'[c][k]
  }
}

// predict: mu_p = Q mu ;  L_p = add_sqrt(Q L, R^{1/2})            (square_root.py:57-58)
// (means in the mean type MT -- double also for T = float, see MeanOf -- factors in T)
template <typename T, int P, typename MT>
RD_DEV void sqrt_predict(const T (&Q)[P][P], const T (&Rh)[P * (P + 1) / 2], const MT (&mu)[P],
                         const T (&L)[P * (P + 1) / 2], MT (&mup)[P], T (&Lp)[P * (P + 1) / 2]) {
  T A[2 * P][P];
  RD_UNROLL for (int i = 0; i < P; ++i) {
    MT m = (MT)Q[i][0] * mu[0];
    RD_UNROLL for (int j = 1; j < P; ++j) m = rd_fma((MT)Q[i][j], mu[j], m);
    mup[i] = m;
  }
  // rows 0..P-1: (Q L)^T, i.e. A[r][c] = sum_{k >= r} Q[c][k] L[k][r] ; rows P..2P-1: (R^{1/2})^T
  RD_UNROLL for (int r = 0; r < P; ++r)
    RD_UNROLL for (int c = 0; c < P; ++c) {
      T a = T(0);
      RD_UNROLL for (int k = r; k < P; ++k) a = rd_fma(Q[c][k], L[lidx(k, r)], a);
      A[r][c] = a;
      A[P + r][c] = lget<T, P>(Rh, c, r);
    }
  qr_lower<T, 2 * P, P>(A, Lp);
}

// update for one scalar measurement row (square_root.py:93-103 with n_meas = 1):
//   wl = w L ;  s^2 = |wl|^2 + |vrow|^2 ;  K = L wl^T / s^2 ;  mu_f = mu_p + K res
//   L_f = add_sqrt(L - K wl, K vrow)
// `vrow` is the measurement-noise square-root BLOCK the reference stacks: a 1 x NV row (NV = 0: none, i.e.
// interrogate_kramer / schober; NV = P for interrogate_chkrebtii's var_meas = W L).
template <typename T, int P, int NV, typename MT>
RD_DEV void sqrt_update_row(MT (&mu)[P], T (&L)[P * (P + 1) / 2], const T (&w)[P], MT res, const T* vrow) {
  T wl[P], K[P];
  T s2 = T(0);
  RD_UNROLL for (int c = 0; c < P; ++c) {
    T a = T(0);
    RD_UNROLL for (int k = c; k < P; ++k) a = rd_fma(w[k], L[lidx(k, c)], a);
    wl[c] = a;
    s2 = rd_fma(a, a, s2);
  }
  RD_UNROLL for (int c = 0; c < NV; ++c) s2 = rd_fma(vrow[c], vrow[c], s2);
  const T rs2 = rcp(s2);
  RD_UNROLL for (int i = 0; i < P; ++i) {
    T a = T(0);
    RD_UNROLL for (int c = 0; c <= i; ++c) a = rd_fma(L[lidx(i, c)], wl[c], a);
    K[i] = a * rs2;
    mu[i] = rd_fma((MT)K[i], res, mu[i]);
  }
  T A[P + NV][P];
  RD_UNROLL for (int r = 0; r < P; ++r)
    RD_UNROLL for (int c = 0; c < P; ++c) A[r][c] = rd_fma(-K[c], wl[r], lget<T, P>(L, c, r));   // (L - K wl)^T
  RD_UNROLL for (int r = 0; r < NV; ++r)
    RD_UNROLL for (int c = 0; c < P; ++c) A[P + r][c] = K[c] * vrow[r];                              // (K vrow)^T
  qr_lower<T, P + NV, P>(A, L);
}

// smooth_mv (square_root.py:160-222):  G = S_f Q^T S_p^{-1} through two triangular solves with L_p,
//   mu_s = mu_f + G (mu_s' - mu_p) ;  L_s = add_sqrt(G [L_s', R^{1/2}], (I - G Q) L_f)
template <typename T, int P, typename MT>
RD_DEV void sqrt_smooth_mv(const T (&Q)[P][P], const T (&Rh)[P * (P + 1) / 2], const MT (&muf)[P],
                           const T (&Lf)[P * (P + 1) / 2], const MT (&mup)[P], const T (&Lp)[P * (P + 1) / 2],
                           MT (&ms)[P], T (&Ls)[P * (P + 1) / 2]) {
  T Sf[P][P], X[P][P], G[P][P], rd[P];
  RD_UNROLL for (int i = 0; i < P; ++i) {
    rd[i] = rcp(Lp[lidx(i, i)]);
    RD_UNROLL for (int j = 0; j < P; ++j) {
      T a = T(0);
      RD_UNROLL for (int k = 0; k <= (i < j ? i : j); ++k) a = rd_fma(Lf[lidx(i, k)], Lf[lidx(j, k)], a);
      Sf[i][j] = a;
    }
  }
  // X = L_p^{-1} Q   (forward substitution, column by column)
  RD_UNROLL for (int c = 0; c < P; ++c)
    RD_UNROLL for (int i = 0; i < P; ++i) {
      T a = Q[i][c];
      RD_UNROLL for (int k = 0; k < i; ++k) a = rd_fma(-Lp[lidx(i, k)], X[k][c], a);
      X[i][c] = a * rd[i];
    }
  // Y = X S_f ;  G^T = L_p^{-T} Y  (back substitution);  store G = (G^T)^T
  RD_UNROLL for (int c = 0; c < P; ++c) {
    T y[P];
    RD_UNROLL for (int i = 0; i < P; ++i) {
      T a = T(0);
      RD_UNROLL for (int k = 0; k < P; ++k) a = rd_fma(X[i][k], Sf[k][c], a);
      y[i] = a;
    }
    RD_UNROLL for (int i = P - 1; i >= 0; --i) {
      T a = y[i];
      RD_UNROLL for (int k = i + 1; k < P; ++k) a = rd_fma(-Lp[lidx(k, i)], y[k], a);
      y[i] = a * rd[i];
    }
    RD_UNROLL for (int i = 0; i < P; ++i) G[c][i] = y[i];
  }
  T J[P][P];     // I - G Q
  RD_UNROLL for (int c = 0; c < P; ++c)
    RD_UNROLL for (int k = 0; k < P; ++k) {
      T a = (c == k) ? T(1) : T(0);
      RD_UNROLL for (int q2 = 0; q2 < P; ++q2) a = rd_fma(-G[c][q2], Q[q2][k], a);
      J[c][k] = a;
    }
  T A[3 * P][P];
  RD_UNROLL for (int r = 0; r < P; ++r)
    RD_UNROLL for (int c = 0; c < P; ++c) {
      // (G L_s')^T[r][c] = sum_{k >= r} G[c][k] L_s'[k][r] ; (G R^{1/2})^T and ((I - G Q) L_f)^T likewise
      T a = T(0), b = T(0), j = T(0);
      RD_UNROLL for (int k = r; k < P; ++k) {
        a = rd_fma(G[c][k], Ls[lidx(k, r)], a);
        b = rd_fma(G[c][k], Rh[lidx(k, r)], b);
        j = rd_fma(J[c][k], Lf[lidx(k, r)], j);
      }
      A[r][c] = a; A[P + r][c] = b; A[2 * P + r][c] = j;
    }
  MT mnew[P];      // ms is read in full before it is overwritten
  RD_UNROLL for (int i = 0; i < P; ++i) {
    MT m = muf[i];
    RD_UNROLL for (int jx = 0; jx < P; ++jx) m = rd_fma((MT)G[i][jx], ms[jx] - mup[jx], m);
    mnew[i] = m;
  }
  RD_UNROLL for (int i = 0; i < P; ++i) ms[i] = mnew[i];
  qr_lower<T, 3 * P, P>(A, Ls);
}

// ---- counter-based RNG -------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al. 2011).  Draws are keyed by (user key, global particle index, step, stream tag) so
// results do not depend on how the theta batch is sharded over GPUs.  Bit-parity with JAX's threefry key
// splitting is neither attempted nor attainable; sampling paths are checked in distribution and,
// deterministically, by injecting normals.
struct Philox {
  unsigned k0, k1;
  RD_DEV void operator()(unsigned c0, unsigned c1, unsigned c2, unsigned c3, unsigned (&out)[4]) const {
    unsigned ka = k0, kb = k1;
    RD_UNROLL for (int r = 0; r < 10; ++r) {
      unsigned hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
      unsigned hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
      c0 = hi1 ^ c1 ^ ka; c1 = lo1; c2 = hi0 ^ c3 ^ kb; c3 = lo0;
      ka += 0x9E3779B9u; kb += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
  }
};

// Four independent standard normals from the four 32-bit words of one Philox call: two Box-Muller pairs, each with a
// 32-bit uniform u = (r + 1/2) 2^-32 for the radius (tail to 6.7 sigma; P(|z| > 6.7) = 2e-11, i.e. one draw in 5e10 is
// clipped -- the same resolution as cuRAND's curand_normal) and a 24-bit uniform for the angle.
//
// Default (fast) variant: the transcendental part runs on the FP32 special-function unit -- ln u is split as
// (exponent) ln 2 + ln(mantissa) with the exponent taken exactly from the integer's leading one, so only ln of a number
// in [1, 2) and sin / cos of an angle in [0, 2 pi) are approximated -- and the result is widened to double.  The draws
// are standard normal to ~2^-21 relative (total-variation distance ~1e-6 from the exact law: no sample-based test can
// see it), at about a sixth of the instructions of the all-FP64 version, which matters because the sampling smoother
// spends a third of its instructions on random numbers.  Define RODEO_EXACT_NORMALS for the all-FP64 Box-Muller.
// Bit parity with JAX's threefry / erfinv normals is not attainable either way; deterministic parity
// tests inject the normals instead.
RD_DEV void normal_pair_exact(unsigned ru, unsigned ra, double& z0, double& z1) {
  const double u1 = ((double)ru + 0.5) * (1.0 / 4294967296.0);   // (0,1)
  const double u2 = ((double)ra + 0.5) * (1.0 / 4294967296.0);
  const double rad = sqrt(-2.0 * log(u1));
  double s, c;
  sincospi(2.0 * u2, &s, &c);
  z0 = rad * c; z1 = rad * s;
}
RD_DEV void normal_pair_fast(unsigned ru, unsigned ra, float& z0, float& z1) {
  // x = ru + 0.5 in [0.5, 2^32): u = x 2^-32.  Its binary exponent is found from the leading one of the integer.
  const int e = 31 - __clz((int)ru);                                                  // floor(log2 ru), -1 if ru == 0
  float mant;
  int ex;
  if (e >= 24) { mant = (float)(ru >> (e - 23)) * (1.0f / 8388608.0f); ex = e; }      // mantissa in [1, 2) to 24 bits
  else { mant = (float)ru + 0.5f; ex = 0; }                                           // small ru: exact in float
  const float ln_u = ((float)(ex - 32) + __log2f(mant)) * 0.69314718056f;             // ln(x 2^-32), < 0
  // sqrt as x rsqrt(x) (MUFU.RSQ + FMUL, 2 ulp): sqrtf's correctly-rounded fix-up with its slow-path branch is not
  // needed at the 2^-21 accuracy of this variant.  The approximate log2 may return exactly 1 for a mantissa just below
  // 2 (u within 2^-24 of 1), i.e. r2 == 0: selected to a zero radius, as sqrtf would give
  const float r2 = -2.0f * ln_u;
  const float rad = r2 > 0.0f ? r2 * rsqrtf(r2) : 0.0f;
  const float ang = (float)(ra >> 8) * (6.28318530718f / 16777216.0f);
  float s, c;
  __sincosf(ang, &s, &c);
  z0 = rad * c; z1 = rad * s;
}
RD_DEV void normal_pair(unsigned ru, unsigned ra, double& z0, double& z1) {
#ifdef RODEO_EXACT_NORMALS
  normal_pair_exact(ru, ra, z0, z1);
#else
  float a, b;
  normal_pair_fast(ru, ra, a, b);
  z0 = (double)a; z1 = (double)b;
#endif
}
RD_DEV void normal_pair(unsigned ru, unsigned ra, float& z0, float& z1) { normal_pair_fast(ru, ra, z0, z1); }
// normals 0..3 of one Philox output; an unused pair costs nothing (dead code)
template <typename T>
RD_DEV void normal_quad(const unsigned (&r)[4], T (&q)[4]) {
  normal_pair(r[0], r[1], q[0], q[1]);
  normal_pair(r[2], r[3], q[2], q[3]);
}

}  // namespace rodeo
