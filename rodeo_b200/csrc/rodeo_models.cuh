// Built-in ODE right-hand sides as __device__ functors, plus forward-mode dual numbers for user models.
//
// The reference takes `ode_fun(X, t, **params)` as a traced Python callable and differentiates it with
// jax.jacfwd inside interrogate_kramer, keeping only the block-diagonal part J[b, :, b]
// (src/rodeo/interrogate.py:75-79).  Here a model is a struct with compile-time shapes:
//
//   static constexpr int NB, P, M, NTHETA;      // n_block, n_bstate, n_bmeas, len(theta)
//   static constexpr int JCOLS;                 // f only reads X[:, 0:JCOLS]  (1 for first_order_pad systems)
//   static constexpr int WCOL;                  // the ODE is X[:, WCOL] = f(X, t): W = e_WCOL (1 for first order)
//   static constexpr bool USES_TIME, HAS_JAC;
//   template <class T> struct Par;              // per-theta constants kept in registers for the whole solve
//   template <class T> static Par<T> load(const T* theta);
//   template <class T, class X> static void rhs(const Par<T>&, T t, const X (&x)[NB][JCOLS], X (&f)[NB][M]);
//   template <class T> static void jac(const Par<T>&, T t, const T (&x)[NB][JCOLS], T (&J)[NB][M][JCOLS]);
//                                               // J[b][r][j] = d f[b][r] / d X[b][j]  (own block only)
//
// `rhs` is templated on the value type X so that a model without an analytic `jac` (HAS_JAC = false; the NVRTC
// user path) gets exactly what jax.jacfwd yields by evaluating `rhs` on dual numbers (block_jacobian below).
//
// No host / toolkit includes: this header is also fed to NVRTC.
#pragma once
#include "rodeo_core.cuh"

namespace rodeo {

// ---- forward-mode dual numbers with N tangent directions -------------------------------------------------------
template <typename T, int N>
struct Dual {
  T v; T d[N];
  RD_DEV Dual() {}
  RD_DEV Dual(T x) : v(x) { RD_UNROLL for (int i = 0; i < N; ++i) d[i] = T(0); }
};
#define RD_DUAL_BIN(op, vexpr, dexpr)                                                                       \
  template <typename T, int N> RD_DEV Dual<T, N> operator op(const Dual<T, N>& a, const Dual<T, N>& b) {    \
    Dual<T, N> r; r.v = vexpr; RD_UNROLL for (int i = 0; i < N; ++i) r.d[i] = dexpr; return r; }
RD_DUAL_BIN(+, a.v + b.v, a.d[i] + b.d[i])
RD_DUAL_BIN(-, a.v - b.v, a.d[i] - b.d[i])
RD_DUAL_BIN(*, a.v * b.v, a.d[i] * b.v + a.v * b.d[i])
RD_DUAL_BIN(/, a.v / b.v, (a.d[i] - (a.v / b.v) * b.d[i]) / b.v)
#undef RD_DUAL_BIN
template <typename T, int N> RD_DEV Dual<T, N> operator-(const Dual<T, N>& a) {
  Dual<T, N> r; r.v = -a.v; RD_UNROLL for (int i = 0; i < N; ++i) r.d[i] = -a.d[i]; return r; }
#define RD_DUAL_SCALAR(op)                                                                                  \
  template <typename T, int N> RD_DEV Dual<T, N> operator op(const Dual<T, N>& a, T b) { return a op Dual<T, N>(b); } \
  template <typename T, int N> RD_DEV Dual<T, N> operator op(T a, const Dual<T, N>& b) { return Dual<T, N>(a) op b; }
RD_DUAL_SCALAR(+) RD_DUAL_SCALAR(-) RD_DUAL_SCALAR(*) RD_DUAL_SCALAR(/)
#undef RD_DUAL_SCALAR
// make the scalar overloads visible next to the Dual ones (the Dual templates would otherwise hide ::sin etc.)
using ::exp; using ::log; using ::sin; using ::cos; using ::sqrt; using ::tanh;
#define RD_DUAL_UN(name, vexpr, dscale)                                                                     \
  template <typename T, int N> RD_DEV Dual<T, N> name(const Dual<T, N>& a) {                                \
    Dual<T, N> r; r.v = vexpr; T s = dscale; RD_UNROLL for (int i = 0; i < N; ++i) r.d[i] = s * a.d[i]; return r; }
RD_DUAL_UN(exp, exp(a.v), r.v)
RD_DUAL_UN(log, log(a.v), T(1) / a.v)
RD_DUAL_UN(sin, sin(a.v), cos(a.v))
RD_DUAL_UN(cos, cos(a.v), -sin(a.v))
RD_DUAL_UN(sqrt, sqrt(a.v), T(0.5) / r.v)
RD_DUAL_UN(tanh, tanh(a.v), T(1) - r.v * r.v)
#undef RD_DUAL_UN

// block-diagonal Jacobian of Model::rhs by dual numbers: block b is seeded with the identity on its own
// JCOLS visible columns, every other block carries zero tangents; row b of the result is kept.
template <class Model, typename T>
RD_DEV void block_jacobian_dual(const typename Model::template Par<T>& q, T t,
                                const T (&x)[Model::NB][Model::JCOLS], T (&f)[Model::NB][Model::M],
                                T (&J)[Model::NB][Model::M][Model::JCOLS]) {
  constexpr int NB = Model::NB, M = Model::M, JC = Model::JCOLS;
  typedef Dual<T, JC> D;
  RD_UNROLL for (int b = 0; b < NB; ++b) {
    D xd[NB][JC], fd[NB][M];
    RD_UNROLL for (int c = 0; c < NB; ++c)
      RD_UNROLL for (int j = 0; j < JC; ++j) {
        xd[c][j] = D(x[c][j]);
        if (c == b) xd[c][j].d[j] = T(1);
      }
    Model::template rhs<T, D>(q, t, xd, fd);
    RD_UNROLL for (int r = 0; r < M; ++r) {
      f[b][r] = fd[b][r].v;
      RD_UNROLL for (int j = 0; j < JC; ++j) J[b][r][j] = fd[b][r].d[j];
    }
  }
}

// f and block-diagonal J in one call, analytic when the model provides it
template <class Model, typename T>
RD_DEV void eval_f_jac(const typename Model::template Par<T>& q, T t, const T (&x)[Model::NB][Model::JCOLS],
                       T (&f)[Model::NB][Model::M], T (&J)[Model::NB][Model::M][Model::JCOLS]) {
  if (Model::HAS_JAC) {
    Model::template rhs<T, T>(q, t, x, f);
    Model::template jac<T>(q, t, x, J);
  } else {
    block_jacobian_dual<Model, T>(q, t, x, f, J);
  }
}

// ---- FitzHugh-Nagumo  (reference README.md:92-99; theta = (a, b, c)) -------------------------------------------
struct FitzHughNagumo {
  static constexpr int NB = 2, P = 3, M = 1, NTHETA = 3, JCOLS = 1, WCOL = 1;
  static constexpr bool USES_TIME = false, HAS_JAC = true;
  template <class T> struct Par { T a, b, c, mrc; };
  template <class T> RD_DEV static Par<T> load(const T* th) {
    Par<T> q; q.a = th[0]; q.b = th[1]; q.c = th[2]; q.mrc = T(-1) / th[2]; return q;
  }
  template <class T, class X>
  RD_DEV static void rhs(const Par<T>& q, T, const X (&x)[NB][JCOLS], X (&f)[NB][M]) {
    X V = x[0][0], R = x[1][0];
    f[0][0] = q.c * (V - V * V * V * T(1.0 / 3.0) + R);
    f[1][0] = q.mrc * (V - q.a + q.b * R);
  }
  template <class T>
  RD_DEV static void jac(const Par<T>& q, T, const T (&x)[NB][JCOLS], T (&J)[NB][M][JCOLS]) {
    T V = x[0][0];
    J[0][0][0] = q.c * (T(1) - V * V);
    J[1][0][0] = q.mrc * q.b;
  }
};

// ---- Lorenz63  (reference docs/examples/lorenz.md:95-101; theta = (rho, sigma, beta)) -------------------------
struct Lorenz63 {
  static constexpr int NB = 3, P = 3, M = 1, NTHETA = 3, JCOLS = 1, WCOL = 1;
  static constexpr bool USES_TIME = false, HAS_JAC = true;
  template <class T> struct Par { T rho, sig, beta; };
  template <class T> RD_DEV static Par<T> load(const T* th) {
    Par<T> q; q.rho = th[0]; q.sig = th[1]; q.beta = th[2]; return q;
  }
  template <class T, class X>
  RD_DEV static void rhs(const Par<T>& q, T, const X (&x)[NB][JCOLS], X (&f)[NB][M]) {
    X a = x[0][0], b = x[1][0], c = x[2][0];
    f[0][0] = -q.sig * a + q.sig * b;
    f[1][0] = q.rho * a - b - a * c;
    f[2][0] = -q.beta * c + a * b;
  }
  template <class T>
  RD_DEV static void jac(const Par<T>& q, T, const T (&)[NB][JCOLS], T (&J)[NB][M][JCOLS]) {
    J[0][0][0] = -q.sig; J[1][0][0] = T(-1); J[2][0][0] = -q.beta;
  }
};

// ---- second-order ODE  x'' = sin(omega t) - k x   (reference docs/examples/higher_order.md:47-58 with
//      theta = (omega, k) = (2, 1); n_deriv = 4) ------------------------------------------------------------------
struct SecondOrderSin {
  static constexpr int NB = 1, P = 4, M = 1, NTHETA = 2, JCOLS = 1, WCOL = 2;
  static constexpr bool USES_TIME = true, HAS_JAC = true;
  template <class T> struct Par { T om, k; };
  template <class T> RD_DEV static Par<T> load(const T* th) { Par<T> q; q.om = th[0]; q.k = th[1]; return q; }
  template <class T, class X>
  RD_DEV static void rhs(const Par<T>& q, T t, const X (&x)[NB][JCOLS], X (&f)[NB][M]) {
    f[0][0] = sin(q.om * t) - q.k * x[0][0];
  }
  template <class T>
  RD_DEV static void jac(const Par<T>& q, T, const T (&)[NB][JCOLS], T (&J)[NB][M][JCOLS]) { J[0][0][0] = -q.k; }
};

// Optional split of a right-hand side into a state-independent forcing term and the rest, f(X, t) = g(X; forcing(t)):
// the forcing of step n+1 does not depend on the filter state, so a kernel may form it off the step's dependency chain
// (fenrir_ws_kernel: an otherwise idle warp computes it ahead into shared memory).  Forcing<Model>::rhs(q, forcing(q, t),
// x, f) must be the same expression as Model::rhs(q, t, x, f): results stay bitwise identical.
template <class Model>
struct Forcing { static constexpr bool HAS = false; };
template <>
struct Forcing<SecondOrderSin> {
  static constexpr bool HAS = true;
  template <class T> RD_DEV static T eval(const SecondOrderSin::Par<T>& q, T t) { return sin(q.om * t); }
  template <class T, class X>
  RD_DEV static void rhs(const SecondOrderSin::Par<T>& q, T frc, const X (&x)[1][1], X (&f)[1][1]) {
    f[0][0] = frc - q.k * x[0][0];
  }
};
// f and block-diagonal J with the forcing term given (frc != nullptr and the model has one) or not
template <class Model, typename T>
RD_DEV void eval_f_jac(const typename Model::template Par<T>& q, T t, const T (&x)[Model::NB][Model::JCOLS],
                       T (&f)[Model::NB][Model::M], T (&J)[Model::NB][Model::M][Model::JCOLS], const T* frc) {
  if constexpr (Forcing<Model>::HAS && Model::HAS_JAC) {
    if (frc != nullptr) {
      Forcing<Model>::template rhs<T, T>(q, *frc, x, f);
      Model::template jac<T>(q, t, x, J);
      return;
    }
  }
  eval_f_jac<Model, T>(q, t, x, f, J);
}
template <class Model, typename T>
RD_DEV void eval_f(const typename Model::template Par<T>& q, T t, const T (&x)[Model::NB][Model::JCOLS],
                   T (&f)[Model::NB][Model::M], const T* frc) {
  if constexpr (Forcing<Model>::HAS) {
    if (frc != nullptr) { Forcing<Model>::template rhs<T, T>(q, *frc, x, f); return; }
  }
  Model::template rhs<T, T>(q, t, x, f);
}

// ---- Hes1 on the log scale  (reference examples/timings.py:253-262; theta = (a..g)) ----------------------------
struct Hes1 {
  static constexpr int NB = 3, P = 3, M = 1, NTHETA = 7, JCOLS = 1, WCOL = 1;
  static constexpr bool USES_TIME = false, HAS_JAC = false;   // Jacobian by dual numbers
  template <class T> struct Par { T a, b, c, d, e, f, g; };
  template <class T> RD_DEV static Par<T> load(const T* th) {
    Par<T> q; q.a = th[0]; q.b = th[1]; q.c = th[2]; q.d = th[3]; q.e = th[4]; q.f = th[5]; q.g = th[6]; return q;
  }
  template <class T, class X>
  RD_DEV static void rhs(const Par<T>& q, T, const X (&x)[NB][JCOLS], X (&f)[NB][M]) {
    X Pm = exp(x[0][0]), Mm = exp(x[1][0]), H = exp(x[2][0]);
    X den = T(1) + Pm * Pm;
    f[0][0] = -q.a * H + q.b * Mm / Pm - q.c;
    f[1][0] = -q.d + q.e / den / Mm;
    f[2][0] = -q.a * Pm + q.f / den / H - q.g;
  }
  template <class T>
  RD_DEV static void jac(const Par<T>&, T, const T (&)[NB][JCOLS], T (&)[NB][M][JCOLS]) {}
};

// ---- SEIRAH  (reference examples/timings.py:339-351; theta = (b, r, alpha, D_e, D_I, D_q)) ---------------------
struct Seirah {
  static constexpr int NB = 6, P = 3, M = 1, NTHETA = 6, JCOLS = 1, WCOL = 1;
  static constexpr bool USES_TIME = false, HAS_JAC = false;
  template <class T> struct Par { T b, r, alpha, De, DI, Dq; };
  template <class T> RD_DEV static Par<T> load(const T* th) {
    Par<T> q; q.b = th[0]; q.r = th[1]; q.alpha = th[2]; q.De = th[3]; q.DI = th[4]; q.Dq = th[5]; return q;
  }
  template <class T, class X>
  RD_DEV static void rhs(const Par<T>& q, T, const X (&x)[NB][JCOLS], X (&f)[NB][M]) {
    X S = x[0][0], E = x[1][0], I = x[2][0], R = x[3][0], A = x[4][0], H = x[5][0];
    X N = S + E + I + R + A + H;
    const T Dh = T(30);
    X inf = q.b * S * (I + q.alpha * A) / N;
    f[0][0] = -inf;
    f[1][0] = inf - E / q.De;
    f[2][0] = q.r * E / q.De - I / q.Dq - I / q.DI;
    f[3][0] = (I + A) / q.DI + H / Dh;
    f[4][0] = (T(1) - q.r) * E / q.De - A / q.DI;
    f[5][0] = I / q.Dq - H / Dh;
  }
  template <class T>
  RD_DEV static void jac(const Par<T>&, T, const T (&)[NB][JCOLS], T (&)[NB][M][JCOLS]) {}
};

// ---- two uncoupled variables in ONE block (n_bmeas = 2): state (x, x', x'', y, y', y''), x' = -a x + sin t,
//      y' = -b y^2, theta = (a, b, c) (c unused); W = [[0,1,0,0,0,0],[0,0,0,0,1,0]].  The reference's API allows any
//      n_bmeas (ode_weight is (n_block, n_bmeas, n_bstate), src/rodeo/solve.py:216-218); this is the instantiation the
//      n_bmeas = 2 golden vectors pin (tests/golden/make_reference_golden.py, pair_one_block) ------------------------
struct PairOneBlock {
  static constexpr int NB = 1, P = 6, M = 2, NTHETA = 3, JCOLS = 4, WCOL = 1;
  static constexpr bool USES_TIME = true, HAS_JAC = true;
  template <class T> struct Par { T a, b; };
  template <class T> RD_DEV static Par<T> load(const T* th) { Par<T> q; q.a = th[0]; q.b = th[1]; return q; }
  template <class T, class X>
  RD_DEV static void rhs(const Par<T>& q, T t, const X (&x)[NB][JCOLS], X (&f)[NB][M]) {
    f[0][0] = -q.a * x[0][0] + sin(t);
    f[0][1] = -q.b * x[0][3] * x[0][3];
  }
  template <class T>
  RD_DEV static void jac(const Par<T>& q, T, const T (&x)[NB][JCOLS], T (&J)[NB][M][JCOLS]) {
    RD_UNROLL for (int r = 0; r < M; ++r)
      RD_UNROLL for (int j = 0; j < JCOLS; ++j) J[0][r][j] = T(0);
    J[0][0][0] = -q.a;
    J[0][1][3] = T(-2) * q.b * x[0][3];
  }
};

}  // namespace rodeo
