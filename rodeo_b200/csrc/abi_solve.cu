// rodeo_b200_solve_mv_f64: batched rodeo.solve_mv
// (reference src/rodeo/solve.py:125-302).
#ifndef RODEO_REAL
#define RODEO_PRIOR_BATCH      /* float64 build: also instantiate the per-theta-prior kernels (QK_DENSE_BATCH) */
#define RODEO_WIDE_MODELS      /* ... and the n_bmeas = 2 model (rodeo_host.h) */
#endif
#include "rodeo_host.h"

#ifndef RODEO_REAL
#define RODEO_REAL double
#define RODEO_SUFFIX _f64
#endif
#define RODEO_CAT2(a, b) a##b
#define RODEO_CAT(a, b) RODEO_CAT2(a, b)
#define RODEO_FN(name) RODEO_CAT(name, RODEO_SUFFIX)
typedef RODEO_REAL real_t;

namespace rodeo {
namespace host {

template <class Model, int INTERR, int QK>
struct SolveMvRun {
  static int run(const RodeoProblem& p, const real_t* W, const real_t* Q, const real_t* R,
                 const CommonArgs<real_t>& a_in, real_t* stash, real_t* mean_out, real_t* var_out, cudaStream_t s) {
    FilterConsts<real_t, Model::NB, Model::P, Model::M> C;
    // per-theta prior: Q, R are device arrays (B, n_block, p, p) the kernels read per thread; one lane per theta
    constexpr bool BATCH = QK == QK_DENSE_BATCH;
    pack_consts<real_t, Model::NB, Model::P, Model::M>(W, BATCH ? nullptr : Q, BATCH ? nullptr : R, C);
    CommonArgs<real_t> a = a_in;
    if (BATCH) { a.q_batch = Q; a.r_batch = R; }
    if (p.B == 0) return RODEO_OK;
    if constexpr (Model::NB >= 2 && !BATCH) {
      // one lane per (theta, block): fewer thetas per warp => longer contiguous output runs (rodeo_kernels.cuh)
      typedef BlockLane<real_t, Model, INTERR, QK> L;
      RODEO_CUDA_OK(cudaFuncSetAttribute(solve_mv_bl_kernel<real_t, Model, INTERR, QK>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, L::BYTES));
      solve_mv_bl_kernel<real_t, Model, INTERR, QK><<<grid_for(p.B, L::TW), 32, L::BYTES, s>>>(
          C, a, stash, stash_ldb(p.B), mean_out, var_out);
    } else {
      constexpr int SMEM = SegBuf<real_t, Fwd<real_t, Model, INTERR, QK>>::BYTES;
      RODEO_CUDA_OK(cudaFuncSetAttribute(solve_mv_kernel<real_t, Model, INTERR, QK>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
      solve_mv_kernel<real_t, Model, INTERR, QK><<<grid_for(p.B, 32), 32, SMEM, s>>>(C, a, stash, stash_ldb(p.B),
                                                                                 mean_out, var_out);
    }
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
  const size_t need = rodeo_b200_workspace_bytes(op, p, (int)sizeof(real_t));
  if (need > 0 && (ws == nullptr || ws_bytes < need)) {
    set_error("workspace too small: need %zu bytes, got %zu", need, ws ? ws_bytes : (size_t)0);
    return RODEO_ERR_WORKSPACE;
  }
  return RODEO_OK;
}

extern "C" int RODEO_FN(rodeo_b200_solve_mv)(const RodeoProblem* p, const real_t* ode_weight, const real_t* prior_weight,
                                       const real_t* prior_var, const real_t* ode_init, const real_t* theta,
                                       const real_t* z_interr, real_t* mean_out, real_t* var_out, void* workspace,
                                       size_t workspace_bytes, void* stream) {
  if (int rc = check_common(p)) return rc;
  if (int rc = check_ws(RODEO_OP_SOLVE_MV, p, workspace, workspace_bytes)) return rc;
  CommonArgs<real_t> a = make_common<real_t>(*p, ode_init, theta, z_interr);
  if (p->model_id >= RODEO_MODEL_USER_BASE) {
    if (sizeof(real_t) != 8) { set_error("user (NVRTC) models are float64 only"); return RODEO_ERR_UNSUPPORTED; }
    real_t* stash = (real_t*)workspace;
    long long ldb = stash_ldb(p->B);
    ObsHook<real_t> no_obs{};        // trailing kernel parameter of the solver kernels (unused: OBS = false)
    const int smem = seg_len(nstate_of(p->n_block, p->n_bstate)) * nstate_of(p->n_block, p->n_bstate) * SEG_PITCH * (int)sizeof(real_t);
    return user_launch(*p, "solve_mv_kernel", "", (const double*)ode_weight, (const double*)prior_weight, (const double*)prior_var, p->user_wcol, p->B, smem,
                       {&a, &stash, &ldb, &mean_out, &var_out, &no_obs}, (cudaStream_t)stream);
  }
  return dispatch_model<SolveMvRun>(*p, ode_weight, prior_weight, *p, ode_weight, prior_weight, prior_var, a, (real_t*)workspace,
                                    mean_out, var_out, (cudaStream_t)stream);
}

