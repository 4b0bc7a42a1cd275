// rodeo_b200_dalton_f64: batched rodeo.inference.dalton (reference src/rodeo/inference/dalton.py:39-235).
#include <cstdlib>

#ifndef RODEO_REAL
#define RODEO_PRIOR_BATCH      /* float64 build: also instantiate the per-theta-prior kernels (QK_DENSE_BATCH) */
#define RODEO_WIDE_MODELS      /* ... and the n_bmeas = 2 model (rodeo_host.h) */
#endif
#include "rodeo_host.h"

// (theta, filter, block) lanes while their grid has at most this many warps per SM sub-partition
#ifndef RODEO_DALTON_BL_BELOW
#define RODEO_DALTON_BL_BELOW 1.0
#endif

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
struct DaltonRun {
  static int run(const RodeoProblem& p, const real_t* W, const real_t* Q, const real_t* R,
                 const CommonArgs<real_t>& a_in, const ObsArgs<real_t>& o, real_t* out, cudaStream_t s) {
    FilterConsts<real_t, Model::NB, Model::P, Model::M> C;
    // per-theta prior: Q, R are device arrays (B, n_block, p, p) the kernels read per thread; one lane per theta
    constexpr bool BATCH = QK == QK_DENSE_BATCH;
    pack_consts<real_t, Model::NB, Model::P, Model::M>(W, BATCH ? nullptr : Q, BATCH ? nullptr : R, C);
    CommonArgs<real_t> a = a_in;
    if (BATCH) { a.q_batch = Q; a.r_batch = R; }
    if (p.n_bobs == 2) {
      // two observation rows per block (correlated noise allowed): the stacked (1 + 2)-row update with the reference's
      // eigen-decomposition log-pdf; float64, interrogate_kramer, one thread per (theta, filter)
      if constexpr (sizeof(real_t) == 8 && INTERR == INTERR_KRAMER && Model::M == 1 && !BATCH) {
        if (p.B == 0) return RODEO_OK;
        CommonArgs<real_t> ag = a;
        ag.dalton_geometry = 2;
        RODEO_CUDA_OK(cudaMemsetAsync(out, 0, (size_t)p.B * sizeof(real_t), s));
        dalton_kernel<real_t, Model, INTERR, QK, 2><<<2 * grid_for(p.B, 32), 32, 0, s>>>(C, ag, o, out);
        g_launches++;
        RODEO_CUDA_OK(cudaGetLastError());
        return RODEO_OK;
      } else {
        set_error("dalton: n_bobs=2 is compiled for float64, interrogate_kramer and a shared prior only");
        return RODEO_ERR_UNSUPPORTED;
      }
    }
    if (p.n_bobs != 1) {
      set_error("dalton: n_bobs=%d is not compiled ahead of time (1 or 2)", p.n_bobs);
      return RODEO_ERR_UNSUPPORTED;
    }
    if (p.B == 0) return RODEO_OK;
    CommonArgs<real_t> ag = a;
    if constexpr (sizeof(real_t) == 8) {
      RODEO_CUDA_OK(cudaMemsetAsync(out, 0, (size_t)p.B * sizeof(real_t), s));
      // A lone warp per SM sub-partition is bound by its own FP64 issue cadence (one DFMA per 2 cycles): spreading the
      // blocks of a filter over lanes (dalton_bl_kernel) halves the instructions per lane, which pays only while every
      // such warp still has a sub-partition to itself.  Measured on B200, FitzHugh-Nagumo N = 800, thread-per-filter /
      // block lanes: 2,048 thetas 0.146 / 0.138 ms, 4,096: 0.162 / 0.138, 8,192: 0.147 / 0.176, 16,384: 0.243 / 0.310,
      // 65,536: 0.733 / 0.976.  Both kernels return bitwise the same numbers.
      bool block_lanes = false;
      if constexpr (Model::NB >= 2 && !BATCH) {
        // (theta, filter, block) lanes while every warp of theirs still gets an SM sub-partition to itself
        typedef BlockLane<real_t, Model, INTERR, QK> L0;
        block_lanes = 2.0 * grid_for(p.B, L0::TW) <= RODEO_DALTON_BL_BELOW * 4.0 * sm_count();
        if (const char* e = getenv("RODEO_DALTON_BLOCK_LANES")) block_lanes = e[0] == '1';      // tuning / tests
        if (block_lanes) {
          typedef BlockLane<real_t, Model, INTERR, QK> L;
          dalton_bl_kernel<real_t, Model, INTERR, QK, 1><<<2 * grid_for(p.B, L::TW), 32, 0, s>>>(C, ag, o, out);
        }
      }
      if (!block_lanes) {
        // joint CTAs, then marginal CTAs, each adding +/- its log-density to the zeroed output (rodeo_kernels.cuh)
        ag.dalton_geometry = 2;
        dalton_kernel<real_t, Model, INTERR, QK, 1><<<2 * grid_for(p.B, 32), 32, 0, s>>>(C, ag, o, out);
      }
    } else {
      ag.dalton_geometry = 1;      // joint warp + marginal warp per CTA: the difference is formed in double
      dalton_kernel<real_t, Model, INTERR, QK, 1><<<grid_for(p.B, 32), 64, 0, s>>>(C, ag, o, out);
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

extern "C" int RODEO_FN(rodeo_b200_dalton)(const RodeoProblem* p, const real_t* ode_weight, const real_t* prior_weight,
                                     const real_t* prior_var, const real_t* ode_init, const real_t* theta,
                                     const real_t* z_interr, const int32_t* obs_ind, const real_t* obs_data,
                                     const real_t* obs_weight, const real_t* obs_var, real_t* loglik_out,
                                     void* workspace, size_t workspace_bytes, void* stream) {
  (void)workspace; (void)workspace_bytes;
  if (int rc = check_common(p)) return rc;
  if (p->n_obs < 1) { set_error("dalton needs n_obs >= 1"); return RODEO_ERR_INVALID; }
  CommonArgs<real_t> a = make_common<real_t>(*p, ode_init, theta, z_interr);
  ObsArgs<real_t> o{p->n_obs, obs_ind, obs_data, obs_weight, obs_var};
  if (p->model_id >= RODEO_MODEL_USER_BASE) {
    if (sizeof(real_t) != 8) { set_error("user (NVRTC) models are float64 only"); return RODEO_ERR_UNSUPPORTED; }
    if (p->n_bobs != 1) { set_error("dalton: n_bobs=%d is not supported for user models (only 1)", p->n_bobs); return RODEO_ERR_UNSUPPORTED; }
    return user_launch(*p, "dalton_kernel", ", 1", (const double*)ode_weight, (const double*)prior_weight, (const double*)prior_var, p->user_wcol, 2 * p->B, 0,
                       {&a, &o, &loglik_out}, (cudaStream_t)stream);
  }
  return dispatch_model<DaltonRun>(*p, ode_weight, prior_weight, *p, ode_weight, prior_weight, prior_var, a, o, loglik_out,
                                   (cudaStream_t)stream);
}
