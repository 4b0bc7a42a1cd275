// rodeo_b200_fenrir_solve_mv_f64: batched rodeo.inference.fenrir.solve_mv (reference
// src/rodeo/inference/fenrir.py:404-457): forward ODE filter, backward filter of the smoothing chain with the Gaussian
// observations, RTS pass over that chain.
#include "rodeo_host.h"

namespace rodeo {
namespace host {

template <class Model, int INTERR, int QK>
struct FenrirSolveMvRun {
  static int run(const RodeoProblem& p, const double* W, const double* Q, const double* R, const CommonArgs<double>& a,
                 const ObsArgs<double>& o, double* h1, double* h2, double* mean_out, double* var_out, cudaStream_t s) {
    FilterConsts<double, Model::NB, Model::P, Model::M> C;
    pack_consts<double, Model::NB, Model::P, Model::M>(W, Q, R, C);
    if (p.B == 0) return RODEO_OK;
    constexpr int SMEM = SegBuf<double, Fwd<double, Model, INTERR, QK>>::BYTES;
    RODEO_CUDA_OK(cudaFuncSetAttribute(fenrir_solve_mv_kernel<double, Model, INTERR, QK, 1>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    fenrir_solve_mv_kernel<double, Model, INTERR, QK, 1><<<grid_for(p.B, 32), 32, SMEM, s>>>(
        C, a, o, h1, h2, stash_ldb(p.B), mean_out, var_out);
    g_launches++;
    RODEO_CUDA_OK(cudaGetLastError());
    return RODEO_OK;
  }
};

inline size_t hist_elems(const RodeoProblem& p, size_t entries) {
  return entries * (size_t)nstate_of(p.n_block, p.n_bstate) * (size_t)stash_ldb(p.B);
}

}  // namespace host
}  // namespace rodeo

using namespace rodeo;
using namespace rodeo::host;

// workspace = [H1: filt[1..N-1] | H2: bfilt[1..N]]
extern "C" size_t rodeo_b200_fenrir_solve_mv_workspace_bytes(const RodeoProblem* p) {
  if (!p || p->n_steps < 1) return 0;
  return (hist_elems(*p, (size_t)p->n_steps - 1) + hist_elems(*p, (size_t)p->n_steps)) * 8;
}

extern "C" int rodeo_b200_fenrir_solve_mv_f64(const RodeoProblem* p, const double* ode_weight, const double* prior_weight,
                                              const double* prior_var, const double* ode_init, const double* theta,
                                              const double* z_interr, const int32_t* obs_ind, const double* obs_data,
                                              const double* obs_weight, const double* obs_var, double* mean_out,
                                              double* var_out, void* workspace, size_t workspace_bytes, void* stream) {
  if (int rc = check_common(p)) return rc;
  if (p->n_obs < 1) { set_error("fenrir.solve_mv needs n_obs >= 1"); return RODEO_ERR_INVALID; }
  if (p->n_bobs != 1) { set_error("fenrir.solve_mv: only n_bobs = 1 is compiled (got %d)", p->n_bobs); return RODEO_ERR_UNSUPPORTED; }
  if (p->model_id >= RODEO_MODEL_USER_BASE) { set_error("fenrir.solve_mv is not available for user (NVRTC) models yet"); return RODEO_ERR_UNSUPPORTED; }
  const size_t need = rodeo_b200_fenrir_solve_mv_workspace_bytes(p);
  if (workspace == nullptr || workspace_bytes < need) {
    set_error("workspace too small: need %zu bytes, got %zu", need, workspace ? workspace_bytes : (size_t)0);
    return RODEO_ERR_WORKSPACE;
  }
  double* h1 = (double*)workspace;
  double* h2 = h1 + hist_elems(*p, (size_t)p->n_steps - 1);
  CommonArgs<double> a = make_common<double>(*p, ode_init, theta, z_interr);
  ObsArgs<double> o{p->n_obs, obs_ind, obs_data, obs_weight, obs_var};
  return dispatch_model<FenrirSolveMvRun>(*p, ode_weight, prior_weight, *p, ode_weight, prior_weight, prior_var, a, o,
                                          h1, h2, mean_out, var_out, (cudaStream_t)stream);
}
