// rodeo_b200_dalton_solve_mv / rodeo_b200_dalton_solve_sim: the data-adaptive solvers
// rodeo.inference.dalton.solve_mv / solve_sim (reference src/rodeo/inference/dalton.py:242-545): the forward filter
// also conditions on the Gaussian observations (augmented update at the observation steps), then the same backward
// smoother / sampler as rodeo.solve_mv / solve_sim.
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

// workspace = [history | step map (n_steps int32, 256-byte aligned)]
inline size_t history_bytes(int op, const RodeoProblem& p) { return round_up(rodeo_b200_workspace_bytes(op, &p, (int)sizeof(real_t)), 256); }

#if !defined(RODEO_ONLY_SIM)
template <class Model, int INTERR, int QK>
struct DaltonSolveMvRun {
  static int run(const RodeoProblem& p, const real_t* W, const real_t* Q, const real_t* R, const CommonArgs<real_t>& a,
                 const ObsHook<real_t>& oh, real_t* stash, real_t* mean_out, real_t* var_out, cudaStream_t s) {
    FilterConsts<real_t, Model::NB, Model::P, Model::M> C;
    pack_consts<real_t, Model::NB, Model::P, Model::M>(W, Q, R, C);
    if (p.B == 0) return RODEO_OK;
    if constexpr (Model::NB >= 2) {
      typedef BlockLane<real_t, Model, INTERR, QK> L;
      RODEO_CUDA_OK(cudaFuncSetAttribute(solve_mv_bl_kernel<real_t, Model, INTERR, QK, true>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, L::BYTES));
      solve_mv_bl_kernel<real_t, Model, INTERR, QK, true><<<grid_for(p.B, L::TW), 32, L::BYTES, s>>>(
          C, a, stash, stash_ldb(p.B), mean_out, var_out, oh);
    } else {
      constexpr int SMEM = SegBuf<real_t, Fwd<real_t, Model, INTERR, QK>>::BYTES;
      RODEO_CUDA_OK(cudaFuncSetAttribute(solve_mv_kernel<real_t, Model, INTERR, QK, true>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
      solve_mv_kernel<real_t, Model, INTERR, QK, true><<<grid_for(p.B, 32), 32, SMEM, s>>>(
          C, a, stash, stash_ldb(p.B), mean_out, var_out, oh);
    }
    g_launches++;
    RODEO_CUDA_OK(cudaGetLastError());
    return RODEO_OK;
  }
};

#endif
#if !defined(RODEO_ONLY_MV)
template <class Model, int INTERR, int QK>
struct DaltonSolveSimRun {
  static int run(const RodeoProblem& p, const real_t* W, const real_t* Q, const real_t* R, const CommonArgs<real_t>& a,
                 const ObsHook<real_t>& oh, const real_t* z_smooth, real_t* stash, real_t* x_out, cudaStream_t s) {
    FilterConsts<real_t, Model::NB, Model::P, Model::M> C;
    pack_consts<real_t, Model::NB, Model::P, Model::M>(W, Q, R, C);
    if (p.B == 0) return RODEO_OK;
    if constexpr (Model::NB >= 2) {
      typedef BlockLane<real_t, Model, INTERR, QK> L;
      constexpr int SMEM = 16 * Model::NB * Model::P * L::PITCH * (int)sizeof(real_t);
      RODEO_CUDA_OK(cudaFuncSetAttribute(solve_sim_bl_kernel<real_t, Model, INTERR, QK, true>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
      solve_sim_bl_kernel<real_t, Model, INTERR, QK, true><<<grid_for(p.B, L::TW), 32, SMEM, s>>>(
          C, a, z_smooth, stash, stash_ldb(p.B), x_out, oh);
    } else {
      constexpr int SMEM = SegBuf<real_t, Fwd<real_t, Model, INTERR, QK>>::BYTES;
      RODEO_CUDA_OK(cudaFuncSetAttribute(solve_sim_kernel<real_t, Model, INTERR, QK, true>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
      solve_sim_kernel<real_t, Model, INTERR, QK, true><<<grid_for(p.B, 32), 32, SMEM, s>>>(
          C, a, z_smooth, stash, stash_ldb(p.B), x_out, oh);
    }
    g_launches++;
    RODEO_CUDA_OK(cudaGetLastError());
    return RODEO_OK;
  }
};

#endif
static int prepare(int op, const RodeoProblem* p, const int32_t* obs_ind, void* workspace, size_t workspace_bytes,
                   cudaStream_t s, int** map_out) {
  if (int rc = check_common(p)) return rc;
  if (p->n_obs < 1) { set_error("the data-adaptive solvers need n_obs >= 1"); return RODEO_ERR_INVALID; }
  if (p->n_bobs != 1 || p->n_bmeas != 1) {
    set_error("data-adaptive solvers: only n_bmeas = n_bobs = 1 is compiled (got %d, %d)", p->n_bmeas, p->n_bobs);
    return RODEO_ERR_UNSUPPORTED;
  }
  if (p->model_id >= RODEO_MODEL_USER_BASE) { set_error("data-adaptive solvers are not available for user (NVRTC) models yet"); return RODEO_ERR_UNSUPPORTED; }
  const size_t hist = history_bytes(op, *p), need = hist + round_up((size_t)p->n_steps * 4, 256);
  if (workspace == nullptr || workspace_bytes < need) {
    set_error("workspace too small: need %zu bytes, got %zu", need, workspace ? workspace_bytes : (size_t)0);
    return RODEO_ERR_WORKSPACE;
  }
  int* map = (int*)((char*)workspace + hist);
  obs_step_map_kernel<0><<<1, 1, 0, s>>>(p->n_steps, p->n_obs, obs_ind, map);
  g_launches++;
  RODEO_CUDA_OK(cudaGetLastError());
  *map_out = map;
  return RODEO_OK;
}

}  // namespace host
}  // namespace rodeo

using namespace rodeo;
using namespace rodeo::host;

#if !defined(RODEO_ONLY_SIM)
extern "C" size_t RODEO_FN(rodeo_b200_dalton_solve_workspace_bytes)(int op, const RodeoProblem* p) {
  if (!p || (op != RODEO_OP_SOLVE_MV && op != RODEO_OP_SOLVE_SIM)) return 0;
  return history_bytes(op, *p) + round_up((size_t)p->n_steps * 4, 256);
}

extern "C" int RODEO_FN(rodeo_b200_dalton_solve_mv)(const RodeoProblem* p, const real_t* ode_weight,
                                                    const real_t* prior_weight, const real_t* prior_var,
                                                    const real_t* ode_init, const real_t* theta, const real_t* z_interr,
                                                    const int32_t* obs_ind, const real_t* obs_data,
                                                    const real_t* obs_weight, const real_t* obs_var, real_t* mean_out,
                                                    real_t* var_out, void* workspace, size_t workspace_bytes,
                                                    void* stream) {
  int* map = nullptr;
  if (int rc = prepare(RODEO_OP_SOLVE_MV, p, obs_ind, workspace, workspace_bytes, (cudaStream_t)stream, &map)) return rc;
  CommonArgs<real_t> a = make_common<real_t>(*p, ode_init, theta, z_interr);
  ObsHook<real_t> oh{{p->n_obs, obs_ind, obs_data, obs_weight, obs_var}, map};
  return dispatch_model<DaltonSolveMvRun>(*p, ode_weight, prior_weight, *p, ode_weight, prior_weight, prior_var, a, oh,
                                          (real_t*)workspace, mean_out, var_out, (cudaStream_t)stream);
}

#endif
#if !defined(RODEO_ONLY_MV)
extern "C" int RODEO_FN(rodeo_b200_dalton_solve_sim)(const RodeoProblem* p, const real_t* ode_weight,
                                                     const real_t* prior_weight, const real_t* prior_var,
                                                     const real_t* ode_init, const real_t* theta, const real_t* z_interr,
                                                     const real_t* z_smooth, const int32_t* obs_ind,
                                                     const real_t* obs_data, const real_t* obs_weight,
                                                     const real_t* obs_var, real_t* x_out, void* workspace,
                                                     size_t workspace_bytes, void* stream) {
  int* map = nullptr;
  if (int rc = prepare(RODEO_OP_SOLVE_SIM, p, obs_ind, workspace, workspace_bytes, (cudaStream_t)stream, &map)) return rc;
  CommonArgs<real_t> a = make_common<real_t>(*p, ode_init, theta, z_interr);
  ObsHook<real_t> oh{{p->n_obs, obs_ind, obs_data, obs_weight, obs_var}, map};
  return dispatch_model<DaltonSolveSimRun>(*p, ode_weight, prior_weight, *p, ode_weight, prior_weight, prior_var, a, oh,
                                           z_smooth, (real_t*)workspace, x_out, (cudaStream_t)stream);
}
#endif
