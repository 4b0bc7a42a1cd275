// Batched square-root Kalman primitives with the reference's names: rodeo.kalmantv.square_root.{predict, update,
// forecast, smooth_mv, smooth_sim, smooth_cond} (src/rodeo/kalmantv/square_root.py:30-385), one thread per problem.
// Variances travel as lower-triangular factors L (var = L L^T), full row-major matrices at the ABI.  Everything is built
// on rodeo_core.cuh's Householder `qr_lower` (= rodeo.utils.add_sqrt, src/rodeo/utils.py:10-24: the R factor of the
// stacked square roots, transposed) and on triangular solves, like the reference; only L L^T is comparable across
// implementations (a QR leaves the signs of R's diagonal free), which is also all the reference's own tests compare
// (tests/test_square_root.py:11-16).  float64, n_state in 1..7, n_meas in 1..3.
#include <type_traits>

#include "rodeo_host.h"

namespace rodeo {
namespace host {

// packed lower factor -> full row-major with a zero upper triangle
template <int P>
RD_DEV void store_lower(double* __restrict__ A, const double (&L)[P * (P + 1) / 2]) {
  RD_UNROLL for (int i = 0; i < P; ++i)
    RD_UNROLL for (int j = 0; j < P; ++j) A[i * P + j] = j <= i ? L[lidx(i, j)] : 0.0;
}
template <int R, int Cn>
RD_DEV void load_full(const double* __restrict__ A, double (&M)[R][Cn]) {
  RD_UNROLL for (int i = 0; i < R; ++i)
    RD_UNROLL for (int j = 0; j < Cn; ++j) M[i][j] = A[i * Cn + j];
}

// G = S_f Q^T S_p^{-1} with S_f = L_f L_f^T, S_p = L_p L_p^T, by two triangular solves with L_p (square_root.py:160-175)
template <int P>
RD_DEV void sqrt_gain(const double (&Q)[P][P], const double (&Lf)[P][P], const double (&Lp)[P][P], double (&G)[P][P]) {
  double Sf[P][P], X[P][P];
  RD_UNROLL for (int i = 0; i < P; ++i)
    RD_UNROLL for (int j = 0; j < P; ++j) {
      double a = 0.0;
      RD_UNROLL for (int k = 0; k < P; ++k) a = fma(Lf[i][k], Lf[j][k], a);
      Sf[i][j] = a;
    }
  RD_UNROLL for (int c = 0; c < P; ++c)                        // X = L_p^{-1} Q
    RD_UNROLL for (int i = 0; i < P; ++i) {
      double a = Q[i][c];
      RD_UNROLL for (int k = 0; k < i; ++k) a = fma(-Lp[i][k], X[k][c], a);
      X[i][c] = a / Lp[i][i];
    }
  RD_UNROLL for (int c = 0; c < P; ++c) {                      // G^T = L_p^{-T} (X S_f)
    double y[P];
    RD_UNROLL for (int i = 0; i < P; ++i) {
      double a = 0.0;
      RD_UNROLL for (int k = 0; k < P; ++k) a = fma(X[i][k], Sf[k][c], a);
      y[i] = a;
    }
    RD_UNROLL for (int i = P - 1; i >= 0; --i) {
      double a = y[i];
      RD_UNROLL for (int k = i + 1; k < P; ++k) a = fma(-Lp[k][i], y[k], a);
      y[i] = a / Lp[i][i];
    }
    RD_UNROLL for (int i = 0; i < P; ++i) G[c][i] = y[i];
  }
}

// predict: mu_p = Q mu + c ; L_p = add_sqrt(Q L, R^{1/2})                      square_root.py:30-60
template <int P>
__global__ void sq_predict(i64 B, const double* mu, const double* L, const double* c, const double* Q, const double* Rh,
                           double* mup, double* Lp) {
  const i64 k = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= B) return;
  double q[P][P], l[P][P], rh[P][P], A[2 * P][P], out[P * (P + 1) / 2];
  load_full<P, P>(Q + k * P * P, q); load_full<P, P>(L + k * P * P, l); load_full<P, P>(Rh + k * P * P, rh);
  RD_UNROLL for (int i = 0; i < P; ++i) {
    double m = c[k * P + i];
    RD_UNROLL for (int j = 0; j < P; ++j) m = fma(q[i][j], mu[k * P + j], m);
    mup[k * P + i] = m;
  }
  RD_UNROLL for (int r = 0; r < P; ++r)                        // rows: (Q L)^T, then (R^{1/2})^T
    RD_UNROLL for (int cc = 0; cc < P; ++cc) {
      double a = 0.0;
      RD_UNROLL for (int j = 0; j < P; ++j) a = fma(q[cc][j], l[j][r], a);
      A[r][cc] = a;
      A[P + r][cc] = rh[cc][r];
    }
  qr_lower<double, 2 * P, P>(A, out);
  store_lower<P>(Lp + k * P * P, out);
}

// update and / or forecast                                                     square_root.py:63-103, 317-345
//   L_z = add_sqrt(W L_p, V^{1/2}) ;  K = S_p W^T (L_z L_z^T)^{-1} by two triangular solves ;  mu_f = mu_p + K (x - mu_z)
//   L_f = add_sqrt(L_p - K W L_p, K V^{1/2}) ;  forecast: (mu_z, L_z L_z^T)
template <int P, int M>
__global__ void sq_update(i64 B, const double* mup, const double* Lp, const double* x, const double* d, const double* W,
                          const double* Vh, double* muf, double* Lf, double* muz, double* Sz) {
  const i64 k = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= B) return;
  double lp[P][P], w[M][P], vh[M][M], WL[M][P], A[P + M][M], lz[M * (M + 1) / 2];
  load_full<P, P>(Lp + k * P * P, lp); load_full<M, P>(W + k * M * P, w); load_full<M, M>(Vh + k * M * M, vh);
  double res[M];
  RD_UNROLL for (int r = 0; r < M; ++r) {
    double mz = d[k * M + r];
    RD_UNROLL for (int i = 0; i < P; ++i) mz = fma(w[r][i], mup[k * P + i], mz);
    res[r] = x != nullptr ? x[k * M + r] - mz : -mz;
    if (muz) muz[k * M + r] = mz;
    RD_UNROLL for (int cc = 0; cc < P; ++cc) {
      double a = 0.0;
      RD_UNROLL for (int j = 0; j < P; ++j) a = fma(w[r][j], lp[j][cc], a);
      WL[r][cc] = a;
    }
  }
  RD_UNROLL for (int r = 0; r < P; ++r)
    RD_UNROLL for (int cc = 0; cc < M; ++cc) A[r][cc] = WL[cc][r];               // (W L_p)^T
  RD_UNROLL for (int r = 0; r < M; ++r)
    RD_UNROLL for (int cc = 0; cc < M; ++cc) A[P + r][cc] = vh[cc][r];           // (V^{1/2})^T
  qr_lower<double, P + M, M>(A, lz);
  double LZ[M][M];
  RD_UNROLL for (int i = 0; i < M; ++i)
    RD_UNROLL for (int j = 0; j < M; ++j) LZ[i][j] = j <= i ? lz[lidx(i, j)] : 0.0;
  if (Sz) {
    RD_UNROLL for (int i = 0; i < M; ++i)
      RD_UNROLL for (int j = 0; j < M; ++j) {
        double a = 0.0;
        RD_UNROLL for (int c2 = 0; c2 < M; ++c2) a = fma(LZ[i][c2], LZ[j][c2], a);
        Sz[(k * M + i) * M + j] = a;
      }
  }
  if (!muf) return;
  // T1 = L_z^{-1} (W L_p L_p^T)  (M x P) ;  K^T = L_z^{-T} T1
  double Kt[M][P];
  RD_UNROLL for (int r = 0; r < M; ++r)
    RD_UNROLL for (int i = 0; i < P; ++i) {
      double a = 0.0;
      RD_UNROLL for (int cc = 0; cc < P; ++cc) a = fma(WL[r][cc], lp[i][cc], a);
      Kt[r][i] = a;
    }
  RD_UNROLL for (int i = 0; i < P; ++i) {
    RD_UNROLL for (int r = 0; r < M; ++r) {
      double a = Kt[r][i];
      RD_UNROLL for (int c2 = 0; c2 < r; ++c2) a = fma(-LZ[r][c2], Kt[c2][i], a);
      Kt[r][i] = a / LZ[r][r];
    }
    RD_UNROLL for (int r = M - 1; r >= 0; --r) {
      double a = Kt[r][i];
      RD_UNROLL for (int c2 = r + 1; c2 < M; ++c2) a = fma(-LZ[c2][r], Kt[c2][i], a);
      Kt[r][i] = a / LZ[r][r];
    }
  }
  RD_UNROLL for (int i = 0; i < P; ++i) {
    double m = mup[k * P + i];
    RD_UNROLL for (int r = 0; r < M; ++r) m = fma(Kt[r][i], res[r], m);
    muf[k * P + i] = m;
  }
  double A2[P + M][P], out[P * (P + 1) / 2];
  RD_UNROLL for (int r = 0; r < P; ++r)                        // (L_p - K W L_p)^T
    RD_UNROLL for (int cc = 0; cc < P; ++cc) {
      double a = lp[cc][r];
      RD_UNROLL for (int m2 = 0; m2 < M; ++m2) a = fma(-Kt[m2][cc], WL[m2][r], a);
      A2[r][cc] = a;
    }
  RD_UNROLL for (int r = 0; r < M; ++r)                        // (K V^{1/2})^T
    RD_UNROLL for (int cc = 0; cc < P; ++cc) {
      double a = 0.0;
      RD_UNROLL for (int m2 = 0; m2 < M; ++m2) a = fma(Kt[m2][cc], vh[m2][r], a);
      A2[P + r][cc] = a;
    }
  qr_lower<double, P + M, P>(A2, out);
  store_lower<P>(Lf + k * P * P, out);
}

// smoothers: mode 0 smooth_mv, 1 smooth_sim, 2 smooth_cond                     square_root.py:178-315, 348-385
//   mv:   L_s = add_sqrt(G [L_next, R^{1/2}], (I - G Q) L_f) ;  sim / cond: add_sqrt(G R^{1/2}, (I - G Q) L_f)
template <int P>
__global__ void sq_smooth(i64 B, int mode, const double* xn, const double* Ln, const double* muf, const double* Lf,
                          const double* mup, const double* Lp, const double* Q, const double* Rh, double* o_mean,
                          double* o_var, double* o_wgt) {
  const i64 k = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= B) return;
  double q[P][P], lf[P][P], lp[P][P], rh[P][P], G[P][P], J[P][P];
  load_full<P, P>(Q + k * P * P, q); load_full<P, P>(Lf + k * P * P, lf); load_full<P, P>(Lp + k * P * P, lp);
  load_full<P, P>(Rh + k * P * P, rh);
  sqrt_gain<P>(q, lf, lp, G);
  RD_UNROLL for (int i = 0; i < P; ++i)
    RD_UNROLL for (int j = 0; j < P; ++j) {
      double a = (i == j) ? 1.0 : 0.0;
      RD_UNROLL for (int c2 = 0; c2 < P; ++c2) a = fma(-G[i][c2], q[c2][j], a);
      J[i][j] = a;
    }
  RD_UNROLL for (int i = 0; i < P; ++i) {
    double m = muf[k * P + i];
    if (mode == 2) { RD_UNROLL for (int j = 0; j < P; ++j) m = fma(-G[i][j], mup[k * P + j], m); }
    else { RD_UNROLL for (int j = 0; j < P; ++j) m = fma(G[i][j], xn[k * P + j] - mup[k * P + j], m); }
    o_mean[k * P + i] = m;
  }
  if (mode == 2) { RD_UNROLL for (int i = 0; i < P; ++i) RD_UNROLL for (int j = 0; j < P; ++j) o_wgt[(k * P + i) * P + j] = G[i][j]; }
  double out[P * (P + 1) / 2];
  auto GX = [&](const double (&X)[P][P], int r, int cc) {      // (G X)^T[r][cc] = sum_j G[cc][j] X[j][r]
    double a = 0.0;
    RD_UNROLL for (int j = 0; j < P; ++j) a = fma(G[cc][j], X[j][r], a);
    return a;
  };
  if (mode == 0) {
    double ln[P][P], A[3 * P][P];
    load_full<P, P>(Ln + k * P * P, ln);
    RD_UNROLL for (int r = 0; r < P; ++r)
      RD_UNROLL for (int cc = 0; cc < P; ++cc) {
        A[r][cc] = GX(ln, r, cc);
        A[P + r][cc] = GX(rh, r, cc);
        double a = 0.0;
        RD_UNROLL for (int j = 0; j < P; ++j) a = fma(J[cc][j], lf[j][r], a);
        A[2 * P + r][cc] = a;
      }
    qr_lower<double, 3 * P, P>(A, out);
  } else {
    double A[2 * P][P];
    RD_UNROLL for (int r = 0; r < P; ++r)
      RD_UNROLL for (int cc = 0; cc < P; ++cc) {
        A[r][cc] = GX(rh, r, cc);
        double a = 0.0;
        RD_UNROLL for (int j = 0; j < P; ++j) a = fma(J[cc][j], lf[j][r], a);
        A[P + r][cc] = a;
      }
    qr_lower<double, 2 * P, P>(A, out);
  }
  store_lower<P>(o_var + k * P * P, out);
}

template <typename F>
int sq_for_m(int m, F&& f) {
  switch (m) {
    case 1: return f(std::integral_constant<int, 1>());
    case 2: return f(std::integral_constant<int, 2>());
    case 3: return f(std::integral_constant<int, 3>());
  }
  set_error("n_meas = %d is not compiled (1..3)", m);
  return RODEO_ERR_UNSUPPORTED;
}
template <typename F>
int sq_for_p(int p, F&& f) {
  switch (p) {
    case 1: return f(std::integral_constant<int, 1>());
    case 2: return f(std::integral_constant<int, 2>());
    case 3: return f(std::integral_constant<int, 3>());
    case 4: return f(std::integral_constant<int, 4>());
    case 5: return f(std::integral_constant<int, 5>());
    case 6: return f(std::integral_constant<int, 6>());
    case 7: return f(std::integral_constant<int, 7>());
  }
  set_error("n_state = %d is not compiled (1..7)", p);
  return RODEO_ERR_UNSUPPORTED;
}
inline int sq_launched() { g_launches++; RODEO_CUDA_OK(cudaGetLastError()); return RODEO_OK; }

}  // namespace host
}  // namespace rodeo

using namespace rodeo;
using namespace rodeo::host;

extern "C" {

int rodeo_b200_sqrt_predict_f64(int64_t B, int n_state, const double* mean_state_past, const double* var_state_past,
                                const double* mean_state, const double* wgt_state, const double* var_state,
                                double* mean_state_pred, double* var_state_pred, void* stream) {
  if (B <= 0) return RODEO_OK;
  return sq_for_p(n_state, [&](auto Pc) {
    sq_predict<decltype(Pc)::value><<<grid_for(B, 64), 64, 0, (cudaStream_t)stream>>>(
        B, mean_state_past, var_state_past, mean_state, wgt_state, var_state, mean_state_pred, var_state_pred);
    return sq_launched();
  });
}

int rodeo_b200_sqrt_update_f64(int64_t B, int n_state, int n_meas, const double* mean_state_pred,
                               const double* var_state_pred, const double* x_meas, const double* mean_meas,
                               const double* wgt_meas, const double* var_meas, double* mean_state_filt,
                               double* var_state_filt, double* mean_fore, double* var_fore, void* stream) {
  if (B <= 0) return RODEO_OK;
  return sq_for_p(n_state, [&](auto Pc) {
    return sq_for_m(n_meas, [&](auto Mc) {
      sq_update<decltype(Pc)::value, decltype(Mc)::value><<<grid_for(B, 64), 64, 0, (cudaStream_t)stream>>>(
          B, mean_state_pred, var_state_pred, x_meas, mean_meas, wgt_meas, var_meas, mean_state_filt, var_state_filt,
          mean_fore, var_fore);
      return sq_launched();
    });
  });
}

int rodeo_b200_sqrt_smooth_f64(int64_t B, int n_state, int mode, const double* x_next, const double* var_next,
                               const double* mean_state_filt, const double* var_state_filt,
                               const double* mean_state_pred, const double* var_state_pred, const double* wgt_state,
                               const double* var_state, double* out_mean, double* out_var, double* out_wgt,
                               void* stream) {
  if (B <= 0) return RODEO_OK;
  if (mode < 0 || mode > 2) { set_error("smooth mode %d", mode); return RODEO_ERR_INVALID; }
  return sq_for_p(n_state, [&](auto Pc) {
    sq_smooth<decltype(Pc)::value><<<grid_for(B, 64), 64, 0, (cudaStream_t)stream>>>(
        B, mode, x_next, var_next, mean_state_filt, var_state_filt, mean_state_pred, var_state_pred, wgt_state, var_state,
        out_mean, out_var, out_wgt);
    return sq_launched();
  });
}

}  // extern "C"
