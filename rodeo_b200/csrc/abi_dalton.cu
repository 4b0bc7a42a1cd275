// rodeo_b200_dalton_f64: batched rodeo.inference.dalton (reference src/rodeo/inference/dalton.py:39-235).
#include "rodeo_host.h"

namespace rodeo {
namespace host {

template <class Model, int INTERR, int QK>
struct DaltonRun {
  static int run(const RodeoProblem& p, const double* W, const double* Q, const double* R,
                 const CommonArgs<double>& a, const ObsArgs<double>& o, double* out, cudaStream_t s) {
    FilterConsts<double, Model::NB, Model::P, Model::M> C;
    pack_consts<double, Model::NB, Model::P, Model::M>(W, Q, R, C);
    if (p.n_bobs != 1) {
      set_error("dalton: n_bobs=%d is not compiled ahead of time (only 1)", p.n_bobs);
      return RODEO_ERR_UNSUPPORTED;
    }
    if (p.B == 0) return RODEO_OK;
    dalton_kernel<double, Model, INTERR, QK, 1><<<grid_for(2 * p.B, 32), 32, 0, s>>>(C, a, o, out);
    g_launches++;
    RODEO_CUDA_OK(cudaGetLastError());
    return RODEO_OK;
  }
};

}  // namespace host
}  // namespace rodeo

using namespace rodeo;
using namespace rodeo::host;

extern "C" int rodeo_b200_dalton_f64(const RodeoProblem* p, const double* ode_weight, const double* prior_weight,
                                     const double* prior_var, const double* ode_init, const double* theta,
                                     const double* z_interr, const int32_t* obs_ind, const double* obs_data,
                                     const double* obs_weight, const double* obs_var, double* loglik_out,
                                     void* workspace, size_t workspace_bytes, void* stream) {
  (void)workspace; (void)workspace_bytes;
  if (int rc = check_common(p)) return rc;
  if (p->n_obs < 1) { set_error("dalton needs n_obs >= 1"); return RODEO_ERR_INVALID; }
  CommonArgs<double> a = make_common<double>(*p, ode_init, theta, z_interr);
  ObsArgs<double> o{p->n_obs, obs_ind, obs_data, obs_weight, obs_var};
  if (p->model_id >= RODEO_MODEL_USER_BASE) {
    if (p->n_bobs != 1) { set_error("dalton: n_bobs=%d is not supported for user models (only 1)", p->n_bobs); return RODEO_ERR_UNSUPPORTED; }
    return user_launch(*p, "dalton_kernel", ", 1", ode_weight, prior_weight, prior_var, p->user_wcol, 2 * p->B, 0,
                       {&a, &o, &loglik_out}, (cudaStream_t)stream);
  }
  return dispatch_model<DaltonRun>(*p, ode_weight, prior_weight, *p, ode_weight, prior_weight, prior_var, a, o, loglik_out,
                                   (cudaStream_t)stream);
}
