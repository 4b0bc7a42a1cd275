// rodeo_b200_fenrir_f64: batched rodeo.inference.fenrir (reference src/rodeo/inference/fenrir.py:86-328).
#include "rodeo_host.h"

namespace rodeo {
namespace host {

template <class Model, int INTERR, int QK>
struct FenrirRun {
  static int run(const RodeoProblem& p, const double* W, const double* Q, const double* R,
                 const CommonArgs<double>& a, const ObsArgs<double>& o, double* stash, double* out, cudaStream_t s) {
    FilterConsts<double, Model::NB, Model::P, Model::M> C;
    pack_consts<double, Model::NB, Model::P, Model::M>(W, Q, R, C);
    if (p.n_bobs != 1) {
      set_error("fenrir: n_bobs=%d is not compiled ahead of time (only 1)", p.n_bobs);
      return RODEO_ERR_UNSUPPORTED;
    }
    if (p.B == 0) return RODEO_OK;
    constexpr int SMEM = SegBuf<double, Fwd<double, Model, INTERR, QK>>::BYTES;
    RODEO_CUDA_OK(cudaFuncSetAttribute(fenrir_kernel<double, Model, INTERR, QK, 1>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    fenrir_kernel<double, Model, INTERR, QK, 1><<<grid_for(p.B, 32), 32, SMEM, s>>>(C, a, o, stash, stash_ldb(p.B), out);
    g_launches++;
    RODEO_CUDA_OK(cudaGetLastError());
    return RODEO_OK;
  }
};

}  // namespace host
}  // namespace rodeo

using namespace rodeo;
using namespace rodeo::host;

extern "C" int rodeo_b200_fenrir_f64(const RodeoProblem* p, const double* ode_weight, const double* prior_weight,
                                     const double* prior_var, const double* ode_init, const double* theta,
                                     const double* z_interr, const int32_t* obs_ind, const double* obs_data,
                                     const double* obs_weight, const double* obs_var, double* loglik_out,
                                     void* workspace, size_t workspace_bytes, void* stream) {
  if (int rc = check_common(p)) return rc;
  if (p->n_obs < 1) { set_error("fenrir needs n_obs >= 1"); return RODEO_ERR_INVALID; }
  const size_t need = rodeo_b200_workspace_bytes(RODEO_OP_FENRIR, p, 8);
  if (need > 0 && (workspace == nullptr || workspace_bytes < need)) {
    set_error("workspace too small: need %zu bytes, got %zu", need, workspace ? workspace_bytes : (size_t)0);
    return RODEO_ERR_WORKSPACE;
  }
  CommonArgs<double> a = make_common<double>(*p, ode_init, theta, z_interr);
  ObsArgs<double> o{p->n_obs, obs_ind, obs_data, obs_weight, obs_var};
  if (p->model_id >= RODEO_MODEL_USER_BASE) {
    if (p->n_bobs != 1) { set_error("fenrir: n_bobs=%d is not supported for user models (only 1)", p->n_bobs); return RODEO_ERR_UNSUPPORTED; }
    double* stash = (double*)workspace;
    long long ldb = stash_ldb(p->B);
    const int smem = seg_len(nstate_of(p->n_block, p->n_bstate)) * nstate_of(p->n_block, p->n_bstate) * SEG_PITCH * 8;
    return user_launch(*p, "fenrir_kernel", ", 1", ode_weight, prior_weight, prior_var, p->user_wcol, p->B, smem,
                       {&a, &o, &stash, &ldb, &loglik_out}, (cudaStream_t)stream);
  }
  return dispatch_model<FenrirRun>(*p, ode_weight, prior_weight, *p, ode_weight, prior_weight, prior_var, a, o, (double*)workspace,
                                   loglik_out, (cudaStream_t)stream);
}
