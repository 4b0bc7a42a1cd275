// Batched Kalman primitives with the reference's names: rodeo.kalmantv.standard.{predict, update, forecast, smooth_mv,
// smooth_sim, smooth_cond} (src/rodeo/kalmantv/standard.py) and rodeo.utils.multivariate_normal_logpdf
// (src/rodeo/utils.py:60-78), one thread per problem.  These are thin kernels over the SAME __device__ functions the
// fused solver kernels inline (rodeo_core.cuh), exposed so that the reference's own known-answer tests for the
// primitives (tests/test_standard.py: brute-force Gaussian conditioning) can be restated against the CUDA code.
// Full (un-packed) row-major matrices at the ABI, float64, n_state in 1..7, n_meas in 1..3.
#include <type_traits>

#include "rodeo_host.h"

namespace rodeo {
namespace host {

template <int P>
RD_DEV void load_sym(const double* __restrict__ A, double (&S)[P * (P + 1) / 2]) {
  RD_UNROLL for (int i = 0; i < P; ++i)
    RD_UNROLL for (int j = i; j < P; ++j) S[sidx<P>(i, j)] = A[i * P + j];
}
template <int P>
RD_DEV void store_sym(double* __restrict__ A, const double (&S)[P * (P + 1) / 2]) {
  RD_UNROLL for (int i = 0; i < P; ++i)
    RD_UNROLL for (int j = 0; j < P; ++j) A[i * P + j] = S[sym<P>(i, j)];
}
template <int R, int Cn>
RD_DEV void load_mat(const double* __restrict__ A, double (&M)[R][Cn]) {
  RD_UNROLL for (int i = 0; i < R; ++i)
    RD_UNROLL for (int j = 0; j < Cn; ++j) M[i][j] = A[i * Cn + j];
}
template <int R, int Cn>
RD_DEV void store_mat(double* __restrict__ A, const double (&M)[R][Cn]) {
  RD_UNROLL for (int i = 0; i < R; ++i)
    RD_UNROLL for (int j = 0; j < Cn; ++j) A[i * Cn + j] = M[i][j];
}

// predict: (mu, S, c, Q, R) -> (mu_p, S_p)          standard.py:31-60
template <int P>
__global__ void ktv_predict(i64 B, const double* mu, const double* S, const double* c, const double* Q, const double* R,
                            double* mup, double* Sp) {
  const i64 k = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= B) return;
  constexpr int NS = P * (P + 1) / 2;
  double m[P], s[NS], q[P][P], r[NS], mo[P], so[NS];
  RD_UNROLL for (int i = 0; i < P; ++i) m[i] = mu[k * P + i];
  load_sym<P>(S + k * P * P, s); load_mat<P, P>(Q + k * P * P, q); load_sym<P>(R + k * P * P, r);
  predict<double, P, QK_DENSE>(q, r, 1.0, m, s, mo, so);
  RD_UNROLL for (int i = 0; i < P; ++i) mup[k * P + i] = mo[i] + c[k * P + i];
  store_sym<P>(Sp + k * P * P, so);
}

// update / forecast: (mu_p, S_p, x, d, W, V) -> (mu_f, S_f) and (mu_z, S_z)     standard.py:63-103, 308-336
template <int P, int M>
__global__ void ktv_update(i64 B, const double* mup, const double* Sp, const double* x, const double* d, const double* W,
                           const double* V, double* muf, double* Sf, double* muz, double* Sz) {
  const i64 k = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= B) return;
  constexpr int NS = P * (P + 1) / 2, MS = M * (M + 1) / 2;
  double m[P], s[NS], w[M][P], v[MS], res[M];
  RD_UNROLL for (int i = 0; i < P; ++i) m[i] = mup[k * P + i];
  load_sym<P>(Sp + k * P * P, s); load_mat<M, P>(W + k * M * P, w); load_sym<M>(V + k * M * M, v);
  double sz[MS];
  RD_UNROLL for (int r = 0; r < M; ++r) {
    double mz = d[k * M + r];
    RD_UNROLL for (int i = 0; i < P; ++i) mz = fma(w[r][i], m[i], mz);
    res[r] = x[k * M + r] - mz;
    if (muz) muz[k * M + r] = mz;
  }
  if (Sz) {                                               // forecast variance W S_p W^T + V
    RD_UNROLL for (int r = 0; r < M; ++r)
      RD_UNROLL for (int c2 = r; c2 < M; ++c2) {
        double acc = v[sidx<M>(r, c2)];
        RD_UNROLL for (int i = 0; i < P; ++i)
          RD_UNROLL for (int j = 0; j < P; ++j) acc = fma(w[r][i] * s[sym<P>(i, j)], w[c2][j], acc);
        sz[sidx<M>(r, c2)] = acc;
      }
    store_sym<M>(Sz + k * M * M, sz);
  }
  if (muf) {
    LogPdfAcc<double> dummy;
    update<double, P, M, false>(m, s, w, res, v, dummy);
    RD_UNROLL for (int i = 0; i < P; ++i) muf[k * P + i] = m[i];
    store_sym<P>(Sf + k * P * P, s);
  }
}

// smoothers: mode 0 smooth_mv, 1 smooth_sim, 2 smooth_cond          standard.py:160-255, 339-371
template <int P>
__global__ void ktv_smooth(i64 B, int mode, const double* xn, const double* Sn, const double* muf, const double* Sf,
                           const double* mup, const double* Sp, const double* Q, double* o_mean, double* o_var,
                           double* o_wgt) {
  const i64 k = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= B) return;
  constexpr int NS = P * (P + 1) / 2;
  double mf[P], sf[NS], mp[P], sp[NS], q[P][P], G[P][P], Ct[P][P];
  RD_UNROLL for (int i = 0; i < P; ++i) { mf[i] = muf[k * P + i]; mp[i] = mup[k * P + i]; }
  load_sym<P>(Sf + k * P * P, sf); load_sym<P>(Sp + k * P * P, sp); load_mat<P, P>(Q + k * P * P, q);
  smooth_gain<double, P, QK_DENSE>(q, sf, sp, G, Ct);
  double m[P], out[NS];
  if (mode == 2) {                                        // A = G, b = mu_f - G mu_p, C = S_f - G (S_f Q^T)^T
    RD_UNROLL for (int i = 0; i < P; ++i) {
      double acc = mf[i];
      RD_UNROLL for (int j = 0; j < P; ++j) acc = fma(-G[i][j], mp[j], acc);
      m[i] = acc;
    }
    store_mat<P, P>(o_wgt + k * P * P, G);
    cond_var<double, P>(sf, G, Ct, out);
  } else {
    RD_UNROLL for (int i = 0; i < P; ++i) {
      double acc = mf[i];
      RD_UNROLL for (int j = 0; j < P; ++j) acc = fma(G[i][j], xn[k * P + j] - mp[j], acc);
      m[i] = acc;
    }
    if (mode == 1) cond_var<double, P>(sf, G, Ct, out);
    else {
      double sn[NS], D[NS];
      load_sym<P>(Sn + k * P * P, sn);
      RD_UNROLL for (int e = 0; e < NS; ++e) { D[e] = sn[e] - sp[e]; out[e] = sf[e]; }
      add_GDGt<double, P>(G, D, out);
    }
  }
  RD_UNROLL for (int i = 0; i < P; ++i) o_mean[k * P + i] = m[i];
  store_sym<P>(o_var + k * P * P, out);
}

template <int M>
__global__ void ktv_logpdf(i64 B, const double* x, const double* mean, const double* cov, double* out) {
  const i64 k = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= B) return;
  double s[M * (M + 1) / 2], res[M];
  load_sym<M>(cov + k * M * M, s);
  RD_UNROLL for (int r = 0; r < M; ++r) res[r] = x[k * M + r] - mean[k * M + r];
  LogPdfAcc<double> acc;
  acc.init();
  logpdf_terms<double, M>(s, res, acc);
  acc.ld.renorm();
  out[k] = acc.value();
}

// the guarded Cholesky-type factor the sampling kernels draw with (rodeo_core.cuh psd_factor): A lower, A A^T = C
template <int P>
__global__ void ktv_psd_factor(i64 B, const double* cov, double* A_out) {
  const i64 k = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= B) return;
  double c[P * (P + 1) / 2], A[P][P];
  load_sym<P>(cov + k * P * P, c);
  RD_UNROLL for (int i = 0; i < P; ++i)
    RD_UNROLL for (int j = 0; j < P; ++j) A[i][j] = 0.0;
  psd_factor<double, P>(c, A);
  RD_UNROLL for (int i = 0; i < P; ++i)
    RD_UNROLL for (int j = 0; j < P; ++j) A_out[(k * P + i) * P + j] = j <= i ? A[i][j] : 0.0;
}

template <int P, typename F>
int for_m(int m, F&& f) {
  switch (m) {
    case 1: return f(std::integral_constant<int, 1>());
    case 2: return f(std::integral_constant<int, 2>());
    case 3: return f(std::integral_constant<int, 3>());
  }
  set_error("n_meas = %d is not compiled (1..3)", m);
  return RODEO_ERR_UNSUPPORTED;
}
template <typename F>
int for_p(int p, F&& f) {
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
inline int launched() { g_launches++; RODEO_CUDA_OK(cudaGetLastError()); return RODEO_OK; }

}  // namespace host
}  // namespace rodeo

#include <type_traits>
using namespace rodeo;
using namespace rodeo::host;

extern "C" {

int rodeo_b200_ktv_predict_f64(int64_t B, int n_state, const double* mean_state_past, const double* var_state_past,
                               const double* mean_state, const double* wgt_state, const double* var_state,
                               double* mean_state_pred, double* var_state_pred, void* stream) {
  if (B <= 0) return RODEO_OK;
  return for_p(n_state, [&](auto Pc) {
    ktv_predict<decltype(Pc)::value><<<grid_for(B, 128), 128, 0, (cudaStream_t)stream>>>(
        B, mean_state_past, var_state_past, mean_state, wgt_state, var_state, mean_state_pred, var_state_pred);
    return launched();
  });
}

/* update (mean_state_filt / var_state_filt non-NULL) and / or forecast (mean_fore / var_fore non-NULL) */
int rodeo_b200_ktv_update_f64(int64_t B, int n_state, int n_meas, const double* mean_state_pred,
                              const double* var_state_pred, const double* x_meas, const double* mean_meas,
                              const double* wgt_meas, const double* var_meas, double* mean_state_filt,
                              double* var_state_filt, double* mean_fore, double* var_fore, void* stream) {
  if (B <= 0) return RODEO_OK;
  return for_p(n_state, [&](auto Pc) {
    return for_m<decltype(Pc)::value>(n_meas, [&](auto Mc) {
      ktv_update<decltype(Pc)::value, decltype(Mc)::value><<<grid_for(B, 128), 128, 0, (cudaStream_t)stream>>>(
          B, mean_state_pred, var_state_pred, x_meas, mean_meas, wgt_meas, var_meas, mean_state_filt, var_state_filt,
          mean_fore, var_fore);
      return launched();
    });
  });
}

/* mode 0: smooth_mv (x_next = mean_state_next, var_next used); 1: smooth_sim (x_next = x_state_next);
 * 2: smooth_cond (out_wgt receives A; out_mean = b; out_var = C) */
int rodeo_b200_ktv_smooth_f64(int64_t B, int n_state, int mode, const double* x_next, const double* var_next,
                              const double* mean_state_filt, const double* var_state_filt,
                              const double* mean_state_pred, const double* var_state_pred, const double* wgt_state,
                              double* out_mean, double* out_var, double* out_wgt, void* stream) {
  if (B <= 0) return RODEO_OK;
  if (mode < 0 || mode > 2) { set_error("smooth mode %d", mode); return RODEO_ERR_INVALID; }
  return for_p(n_state, [&](auto Pc) {
    ktv_smooth<decltype(Pc)::value><<<grid_for(B, 128), 128, 0, (cudaStream_t)stream>>>(
        B, mode, x_next, var_next, mean_state_filt, var_state_filt, mean_state_pred, var_state_pred, wgt_state, out_mean,
        out_var, out_wgt);
    return launched();
  });
}

int rodeo_b200_psd_factor_f64(int64_t B, int n, const double* cov, double* factor_out, void* stream) {
  if (B <= 0) return RODEO_OK;
  return for_p(n, [&](auto Pc) {
    ktv_psd_factor<decltype(Pc)::value><<<grid_for(B, 128), 128, 0, (cudaStream_t)stream>>>(B, cov, factor_out);
    return launched();
  });
}

int rodeo_b200_mvn_logpdf_f64(int64_t B, int n, const double* x, const double* mean, const double* cov, double* out,
                              void* stream) {
  if (B <= 0) return RODEO_OK;
  return for_m<1>(n, [&](auto Mc) {
    ktv_logpdf<decltype(Mc)::value><<<grid_for(B, 128), 128, 0, (cudaStream_t)stream>>>(B, x, mean, cov, out);
    return launched();
  });
}

}  // extern "C"
