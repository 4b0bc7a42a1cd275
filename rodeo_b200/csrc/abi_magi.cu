// rodeo_b200_magi_logdens_f64: batched rodeo.inference.magi_logdens
// (reference src/rodeo/inference/magi.py:6-99): the log-density p(U_{0:N}, Z = 0 | theta) of a GIVEN trajectory under
// the block-diagonal Markov prior, as a Kalman filter that observes the first n_active state entries of every block
// without noise:  per step  predict -> forecast -> log N(x; mu_fore, S_fore) -> update,  summed over steps and blocks.
//
// One thread per theta; the blocks are independent here (the ODE enters only through the trajectory the caller
// expanded), so a thread runs one block's whole filter after the other and the sum is deterministic.  The measurement
// rows are the unit rows e_0 .. e_{na-1}: forecast mean / variance are the leading entries of the predicted moments.
// The log-density is the Cholesky-type one of jax.scipy.stats.multivariate_normal.logpdf (magi.py:67-71) -- NOT the
// eigenvalue cut-off version of rodeo.utils -- evaluated through the same un-pivoted L D L^T as the smoother gains.
#include "rodeo_host.h"

namespace rodeo {
namespace host {

constexpr int MAGI_MAX_NB = 8, MAGI_MAX_P = 4;
struct MagiConsts {
  double Q[MAGI_MAX_NB][MAGI_MAX_P][MAGI_MAX_P];
  double R[MAGI_MAX_NB][MAGI_MAX_P * (MAGI_MAX_P + 1) / 2];
};

template <int P, int NA>
__global__ void magi_kernel(const __grid_constant__ MagiConsts C, i64 B, int n_steps, int n_block,
                            const double* __restrict__ X /* (B, N+1, nb, P) */, double* __restrict__ out) {
  constexpr int NS = P * (P + 1) / 2, AS = NA * (NA + 1) / 2;
  const i64 th = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (th >= B) return;
  const i64 row = (i64)n_block * P;
  const double* x = X + th * (i64)(n_steps + 1) * row;
  double total = 0.0;
  for (int b = 0; b < n_block; ++b) {
    double Q[P][P], R[NS], mu[P], S[NS];
    RD_UNROLL for (int i = 0; i < P; ++i)
      RD_UNROLL for (int j = 0; j < P; ++j) Q[i][j] = C.Q[b][i][j];
    RD_UNROLL for (int k = 0; k < NS; ++k) { R[k] = C.R[b][k]; S[k] = 0.0; }
    RD_UNROLL for (int i = 0; i < P; ++i) mu[i] = x[b * P + i];          // filter_init: (ode_state[0], 0)   magi.py:86-89
    double quad = 0.0;
    LogAcc<double> ld;
    ld.init();
    for (int n = 1; n <= n_steps; ++n) {
      double mp[P], Sp[NS];
      predict<double, P, QK_DENSE>(Q, R, 1.0, mu, S, mp, Sp);
      // forecast with W = eye(NA, P), mean_meas = 0, var_meas = 0  (magi.py:37-43, 61-66)
      double res[NA], Sf[AS], L[NA][NA], rD[NA], sol[NA];
      RD_UNROLL for (int r = 0; r < NA; ++r) {
        res[r] = x[n * row + b * P + r] - mp[r];
        sol[r] = res[r];
        RD_UNROLL for (int s = r; s < NA; ++s) Sf[sidx<NA>(r, s)] = Sp[sidx<P>(r, s)];
      }
      ldlt<double, NA>(Sf, L, rD);
      ldlt_solve<double, NA>(L, rD, sol);
      RD_UNROLL for (int r = 0; r < NA; ++r) {
        quad = fma(res[r], sol[r], quad);
        ld.add(rD[r]);                                   // log det S = -sum log(1 / d_r)
      }
      ld.renorm();
      // update with the same rows (magi.py:73-80)
      double wm[NA][P], V[AS];
      RD_UNROLL for (int r = 0; r < NA; ++r)
        RD_UNROLL for (int j = 0; j < P; ++j) wm[r][j] = (j == r) ? 1.0 : 0.0;
      RD_UNROLL for (int k = 0; k < AS; ++k) V[k] = 0.0;
      LogPdfAcc<double> dummy;
      update<double, P, NA, false>(mp, Sp, wm, res, V, dummy);
      RD_UNROLL for (int i = 0; i < P; ++i) mu[i] = mp[i];
      RD_UNROLL for (int k = 0; k < NS; ++k) S[k] = Sp[k];
    }
    total += -0.5 * (quad - ld.value()) - 0.5 * 1.8378770664093454836 * (double)NA * (double)n_steps;
  }
  out[th] = total;
}

template <int P, int NA>
int magi_launch(const MagiConsts& C, i64 B, int n_steps, int n_block, const double* X, double* out, cudaStream_t s) {
  magi_kernel<P, NA><<<grid_for(B, 64), 64, 0, s>>>(C, B, n_steps, n_block, X, out);
  g_launches++;
  RODEO_CUDA_OK(cudaGetLastError());
  return RODEO_OK;
}

}  // namespace host
}  // namespace rodeo

using namespace rodeo;
using namespace rodeo::host;

extern "C" int rodeo_b200_magi_logdens_f64(int64_t B, int n_steps, int n_block, int n_bstate, int n_active,
                                           const double* prior_weight, const double* prior_var, const double* ode_state,
                                           double* logdens_out, void* stream) {
  if (B < 0 || n_steps < 1) { set_error("magi_logdens: need B >= 0 and n_steps >= 1"); return RODEO_ERR_INVALID; }
  if (n_block < 1 || n_block > MAGI_MAX_NB || n_bstate < 2 || n_bstate > MAGI_MAX_P || n_active < 1 || n_active >= n_bstate + 1) {
    set_error("magi_logdens: n_block=%d (1..%d), n_bstate=%d (2..%d), n_active=%d (1..n_bstate) not supported", n_block,
              MAGI_MAX_NB, n_bstate, MAGI_MAX_P, n_active);
    return RODEO_ERR_UNSUPPORTED;
  }
  if (B == 0) return RODEO_OK;
  MagiConsts C{};
  for (int b = 0; b < n_block; ++b)
    for (int i = 0; i < n_bstate; ++i) {
      for (int j = 0; j < n_bstate; ++j) C.Q[b][i][j] = prior_weight[(b * n_bstate + i) * n_bstate + j];
      for (int j = i; j < n_bstate; ++j) {
        // packed upper triangle with the row length of THIS instantiation: sidx<P>(i, j) = i*P - i(i-1)/2 + (j - i)
        C.R[b][i * n_bstate - i * (i - 1) / 2 + (j - i)] = prior_var[(b * n_bstate + i) * n_bstate + j];
      }
    }
  cudaStream_t s = (cudaStream_t)stream;
#define MAGI_CASE(P_, NA_) if (n_bstate == P_ && n_active == NA_) return magi_launch<P_, NA_>(C, B, n_steps, n_block, ode_state, logdens_out, s)
  MAGI_CASE(2, 1); MAGI_CASE(2, 2);
  MAGI_CASE(3, 1); MAGI_CASE(3, 2); MAGI_CASE(3, 3);
  MAGI_CASE(4, 1); MAGI_CASE(4, 2); MAGI_CASE(4, 3); MAGI_CASE(4, 4);
#undef MAGI_CASE
  set_error("magi_logdens: unreachable shape");
  return RODEO_ERR_UNSUPPORTED;
}
