// rodeo_b200_solve_sim_f64: batched rodeo.solve_sim
// (reference src/rodeo/solve.py:125-302).
#include "rodeo_host.h"

namespace rodeo {
namespace host {

template <class Model, int INTERR, int QK>
struct SolveSimRun {
  static int run(const RodeoProblem& p, const double* W, const double* Q, const double* R,
                 const CommonArgs<double>& a, const double* z_smooth, double* stash, double* x_out, cudaStream_t s) {
    FilterConsts<double, Model::NB, Model::P, Model::M> C;
    pack_consts<double, Model::NB, Model::P, Model::M>(W, Q, R, C);
    if (p.B == 0) return RODEO_OK;
    constexpr int SMEM = SegBuf<double, Fwd<double, Model, INTERR, QK>>::BYTES;
    RODEO_CUDA_OK(cudaFuncSetAttribute(solve_sim_kernel<double, Model, INTERR, QK>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    solve_sim_kernel<double, Model, INTERR, QK><<<grid_for(p.B, 32), 32, SMEM, s>>>(C, a, z_smooth, stash,
                                                                                stash_ldb(p.B), x_out);
    g_launches++;
    RODEO_CUDA_OK(cudaGetLastError());
    return RODEO_OK;
  }
};

}  // namespace host
}  // namespace rodeo

using namespace rodeo;
using namespace rodeo::host;

static int check_ws(int op, const RodeoProblem* p, void* ws, size_t ws_bytes) {
  const size_t need = rodeo_b200_workspace_bytes(op, p, 8);
  if (need > 0 && (ws == nullptr || ws_bytes < need)) {
    set_error("workspace too small: need %zu bytes, got %zu", need, ws ? ws_bytes : (size_t)0);
    return RODEO_ERR_WORKSPACE;
  }
  return RODEO_OK;
}

extern "C" int rodeo_b200_solve_sim_f64(const RodeoProblem* p, const double* ode_weight, const double* prior_weight,
                                        const double* prior_var, const double* ode_init, const double* theta,
                                        const double* z_interr, const double* z_smooth, double* x_out,
                                        void* workspace, size_t workspace_bytes, void* stream) {
  if (int rc = check_common(p)) return rc;
  if (int rc = check_ws(RODEO_OP_SOLVE_SIM, p, workspace, workspace_bytes)) return rc;
  CommonArgs<double> a = make_common<double>(*p, ode_init, theta, z_interr);
  if (p->model_id >= RODEO_MODEL_USER_BASE) {
    double* stash = (double*)workspace;
    long long ldb = stash_ldb(p->B);
    const int smem = seg_len(nstate_of(p->n_block, p->n_bstate)) * nstate_of(p->n_block, p->n_bstate) * SEG_PITCH * 8;
    return user_launch(*p, "solve_sim_kernel", "", ode_weight, prior_weight, prior_var, p->user_wcol, p->B, smem,
                       {&a, &z_smooth, &stash, &ldb, &x_out}, (cudaStream_t)stream);
  }
  return dispatch_model<SolveSimRun>(*p, ode_weight, prior_weight, *p, ode_weight, prior_weight, prior_var, a, z_smooth,
                                     (double*)workspace, x_out, (cudaStream_t)stream);
}
