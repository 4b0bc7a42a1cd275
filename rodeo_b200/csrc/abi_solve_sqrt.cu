// rodeo_b200_solve_mv_sqrt_f64: batched rodeo.solve_mv with kalman_type="square-root"
// (reference src/rodeo/solve.py:208-302 on src/rodeo/kalmantv/square_root.py).  prior_var is the lower Cholesky
// factor of R; var_out receives lower-triangular factors L (var = L L^T), as in the reference.
#include "rodeo_host.h"

#ifndef RODEO_REAL
#define RODEO_REAL double
#define RODEO_SUFFIX _f64
#define RODEO_SQRT_F64
#endif
#define RODEO_CAT2(a, b) a##b
#define RODEO_CAT(a, b) RODEO_CAT2(a, b)
#define RODEO_FN(name) RODEO_CAT(name, RODEO_SUFFIX)
typedef RODEO_REAL real_t;

namespace rodeo {
namespace host {

template <class Model, int INTERR, int QK>
struct SolveMvSqrtRun {
  static int run(const RodeoProblem& p, const real_t* W, const real_t* Q, const real_t* Rh,
                 const CommonArgs<real_t>& a, real_t* stash, real_t* mean_out, real_t* var_out, cudaStream_t s) {
    if constexpr (INTERR == INTERR_RODEO) {
      set_error("interrogate_rodeo is not defined for kalman_type=\"square-root\" (the reference would form W L W^T)");
      return RODEO_ERR_UNSUPPORTED;
    } else {
      constexpr int NB = Model::NB, P = Model::P, M = Model::M;
      FilterConsts<real_t, NB, P, M> C;
      for (int b = 0; b < NB; ++b) {
        for (int i = 0; i < P; ++i)
          for (int j = 0; j < P; ++j) C.Q[b][i][j] = Q[(b * P + i) * P + j];
        for (int i = 0; i < P; ++i)
          for (int j = 0; j <= i; ++j) C.R[b][i * (i + 1) / 2 + j] = Rh[(b * P + i) * P + j];   // packed lower
        for (int r = 0; r < M; ++r)
          for (int j = 0; j < P; ++j) C.W[b][r][j] = W[(b * M + r) * P + j];
      }
      if (p.B == 0) return RODEO_OK;
      constexpr int SMEM = SegBuf<real_t, Fwd<real_t, Model, INTERR, QK_DENSE>>::BYTES;
      RODEO_CUDA_OK(cudaFuncSetAttribute(solve_mv_sqrt_kernel<real_t, Model, INTERR>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
      solve_mv_sqrt_kernel<real_t, Model, INTERR><<<grid_for(p.B, 32), 32, SMEM, s>>>(C, a, stash, stash_ldb(p.B),
                                                                                     mean_out, var_out);
      g_launches++;
      RODEO_CUDA_OK(cudaGetLastError());
      return RODEO_OK;
    }
  }
};

}  // namespace host
}  // namespace rodeo

using namespace rodeo;
using namespace rodeo::host;

#ifdef RODEO_SQRT_F64
// sized for float64 elements; the float32 entry point needs half of it and accepts the same size
extern "C" size_t rodeo_b200_solve_mv_sqrt_workspace_bytes(const RodeoProblem* p) {
  if (!p) return 0;
  const int nstate = nstate_of(p->n_block, p->n_bstate);
  const int K = seg_len(nstate);
  const size_t J = ((size_t)p->n_steps + K - 1) / K;
  return (J > 0 ? J - 1 : 0) * (size_t)nstate * (size_t)stash_ldb(p->B) * 8;
}
#else
extern "C" size_t rodeo_b200_solve_mv_sqrt_workspace_bytes(const RodeoProblem* p);
#endif

extern "C" int RODEO_FN(rodeo_b200_solve_mv_sqrt)(const RodeoProblem* p, const real_t* ode_weight,
                                            const real_t* prior_weight, const real_t* prior_var_sqrt,
                                            const real_t* ode_init, const real_t* theta, const real_t* z_interr,
                                            real_t* mean_out, real_t* var_sqrt_out, void* workspace,
                                            size_t workspace_bytes, void* stream) {
  if (!p) { set_error("RodeoProblem is NULL"); return RODEO_ERR_INVALID; }
  RodeoProblem q = *p;
  q.kalman_type = RODEO_KALMAN_STANDARD;            // check_common() vets the rest; this entry point IS the sqrt path
  if (int rc = check_common(&q)) return rc;
  if (p->n_bmeas != 1) { set_error("square-root path: n_bmeas must be 1"); return RODEO_ERR_UNSUPPORTED; }
  if (p->model_id >= RODEO_MODEL_USER_BASE) { set_error("square-root path is not available for user (NVRTC) models yet"); return RODEO_ERR_UNSUPPORTED; }
  const size_t need = rodeo_b200_solve_mv_sqrt_workspace_bytes(p) / (8 / sizeof(real_t));
  if (need > 0 && (workspace == nullptr || workspace_bytes < need)) {
    set_error("workspace too small: need %zu bytes, got %zu", need, workspace ? workspace_bytes : (size_t)0);
    return RODEO_ERR_WORKSPACE;
  }
  CommonArgs<real_t> a = make_common<real_t>(*p, ode_init, theta, z_interr);
  // dispatch on (model, interrogation) only: the structure argument is unused by the square-root kernels
  return dispatch_model<SolveMvSqrtRun>(*p, (const real_t*)nullptr, (const real_t*)nullptr, *p, ode_weight, prior_weight,
                                        prior_var_sqrt, a, (real_t*)workspace, mean_out, var_sqrt_out,
                                        (cudaStream_t)stream);
}
