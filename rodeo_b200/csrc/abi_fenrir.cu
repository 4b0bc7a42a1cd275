// rodeo_b200_fenrir_f64: batched rodeo.inference.fenrir (reference src/rodeo/inference/fenrir.py:86-328).
#include <cstdlib>

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
struct FenrirRun {
  static int run(const RodeoProblem& p, const real_t* W, const real_t* Q, const real_t* R,
                 const CommonArgs<real_t>& a_in, const ObsArgs<real_t>& o, real_t* stash, real_t* out, cudaStream_t s) {
    FilterConsts<real_t, Model::NB, Model::P, Model::M> C;
    // per-theta prior: Q, R are device arrays (B, n_block, p, p) the kernels read per thread; one lane per theta
    constexpr bool BATCH = QK == QK_DENSE_BATCH;
    pack_consts<real_t, Model::NB, Model::P, Model::M>(W, BATCH ? nullptr : Q, BATCH ? nullptr : R, C);
    CommonArgs<real_t> a = a_in;
    if (BATCH) { a.q_batch = Q; a.r_batch = R; }
    if (p.n_bobs == 2) {
      // two observation rows per block: float64, interrogate_kramer, the one-warp kernel
      if constexpr (sizeof(real_t) == 8 && INTERR == INTERR_KRAMER && !BATCH) {
        if (p.B == 0) return RODEO_OK;
        constexpr int SMEM2 = 2 * SegBuf<real_t, Fwd<real_t, Model, INTERR, QK>>::BYTES;
        RODEO_CUDA_OK(cudaFuncSetAttribute(fenrir_kernel<real_t, Model, INTERR, QK, 2>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM2));
        fenrir_kernel<real_t, Model, INTERR, QK, 2><<<grid_for(p.B, 32), 32, SMEM2, s>>>(C, a, o, stash, stash_ldb(p.B), out);
        g_launches++;
        RODEO_CUDA_OK(cudaGetLastError());
        return RODEO_OK;
      } else {
        set_error("fenrir: n_bobs=2 is compiled for float64, interrogate_kramer and a shared prior only");
        return RODEO_ERR_UNSUPPORTED;
      }
    }
    if (p.n_bobs != 1) {
      set_error("fenrir: n_bobs=%d is not compiled ahead of time (1 or 2)", p.n_bobs);
      return RODEO_ERR_UNSUPPORTED;
    }
    if (p.B == 0) return RODEO_OK;
    // warp-specialised backward sweep (rodeo_kernels.cuh) while all of its CTAs are resident at once, i.e. while the
    // one-warp kernel would be bound by a theta's serial chain; measured on B200: second-order ODE, 16,384 thetas x 2,000
    // steps 2.65 -> 1.95 ms; FitzHugh-Nagumo, 65,536 thetas x 800 steps 3.54 -> 3.64 ms (throughput-bound: not used)
    bool ws = !BATCH && sizeof(real_t) == 8 && (long long)grid_for(p.B, 32) <= 4LL * sm_count();
    if (const char* e = getenv("RODEO_FENRIR_WS")) ws = ws && e[0] == '1';      // tuning / tests
    if constexpr (sizeof(real_t) == 8 && !BATCH) {
      if (ws) {
        constexpr int NSP = Model::P * (Model::P + 1) / 2;
        // ring + mu_p region, then the forcing buffer of models with a state-independent forcing term (two chunks)
        constexpr int SMEM_WS = (RODEO_FENRIR_SL * RODEO_FENRIR_NP * Model::NB * (Model::P * Model::P + 2 * Model::P + NSP) * 32 +
                                 (Forcing<Model>::HAS ? 2 * RODEO_FENRIR_FCH * 32 : 0)) * (int)sizeof(real_t);
        RODEO_CUDA_OK(cudaFuncSetAttribute(fenrir_ws_kernel<real_t, Model, INTERR, QK, 1>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_WS));
        RODEO_CUDA_OK(cudaFuncSetAttribute(fenrir_ws_kernel<real_t, Model, INTERR, QK, 1>,
                                           cudaFuncAttributePreferredSharedMemoryCarveout,
                                           cudaSharedmemCarveoutMaxShared));
        fenrir_ws_kernel<real_t, Model, INTERR, QK, 1><<<grid_for(p.B, 32), 32 * (1 + RODEO_FENRIR_NP), SMEM_WS, s>>>(C, a, o, stash,
                                                                                               stash_ldb(p.B), out);
      }
    }
    if (!ws) {
      constexpr int SMEM = 2 * SegBuf<real_t, Fwd<real_t, Model, INTERR, QK>>::BYTES;      // double-buffered history
      RODEO_CUDA_OK(cudaFuncSetAttribute(fenrir_kernel<real_t, Model, INTERR, QK, 1>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
      fenrir_kernel<real_t, Model, INTERR, QK, 1><<<grid_for(p.B, 32), 32, SMEM, s>>>(C, a, o, stash, stash_ldb(p.B), out);
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

extern "C" int RODEO_FN(rodeo_b200_fenrir)(const RodeoProblem* p, const real_t* ode_weight, const real_t* prior_weight,
                                     const real_t* prior_var, const real_t* ode_init, const real_t* theta,
                                     const real_t* z_interr, const int32_t* obs_ind, const real_t* obs_data,
                                     const real_t* obs_weight, const real_t* obs_var, real_t* loglik_out,
                                     void* workspace, size_t workspace_bytes, void* stream) {
  if (int rc = check_common(p)) return rc;
  if (p->n_obs < 1) { set_error("fenrir needs n_obs >= 1"); return RODEO_ERR_INVALID; }
  const size_t need = rodeo_b200_workspace_bytes(RODEO_OP_FENRIR, p, (int)sizeof(real_t));
  if (need > 0 && (workspace == nullptr || workspace_bytes < need)) {
    set_error("workspace too small: need %zu bytes, got %zu", need, workspace ? workspace_bytes : (size_t)0);
    return RODEO_ERR_WORKSPACE;
  }
  CommonArgs<real_t> a = make_common<real_t>(*p, ode_init, theta, z_interr);
  ObsArgs<real_t> o{p->n_obs, obs_ind, obs_data, obs_weight, obs_var};
  if (p->model_id >= RODEO_MODEL_USER_BASE) {
    if (sizeof(real_t) != 8) { set_error("user (NVRTC) models are float64 only"); return RODEO_ERR_UNSUPPORTED; }
    if (p->n_bobs != 1) { set_error("fenrir: n_bobs=%d is not supported for user models (only 1)", p->n_bobs); return RODEO_ERR_UNSUPPORTED; }
    real_t* stash = (real_t*)workspace;
    long long ldb = stash_ldb(p->B);
    const int smem = 2 * seg_len(nstate_of(p->n_block, p->n_bstate)) * nstate_of(p->n_block, p->n_bstate) * SEG_PITCH * (int)sizeof(real_t);
    return user_launch(*p, "fenrir_kernel", ", 1", (const double*)ode_weight, (const double*)prior_weight, (const double*)prior_var, p->user_wcol, p->B, smem,
                       {&a, &o, &stash, &ldb, &loglik_out}, (cudaStream_t)stream);
  }
  return dispatch_model<FenrirRun>(*p, ode_weight, prior_weight, *p, ode_weight, prior_weight, prior_var, a, o, (real_t*)workspace,
                                   loglik_out, (cudaStream_t)stream);
}
