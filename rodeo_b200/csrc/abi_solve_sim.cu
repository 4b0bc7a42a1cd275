// rodeo_b200_solve_sim_f64: batched rodeo.solve_sim
// (reference src/rodeo/solve.py:125-302).
#include <cstdlib>

#ifndef RODEO_REAL
#define RODEO_PRIOR_BATCH      /* float64 build: also instantiate the per-theta-prior kernels (QK_DENSE_BATCH) */
#endif
#include "rodeo_host.h"
#include "rodeo_sched.cuh"

#ifndef RODEO_REAL
#define RODEO_REAL double
#define RODEO_SUFFIX _f64
#define RODEO_SOLVE_SIM_LOGLIK
#endif
#define RODEO_CAT2(a, b) a##b
#define RODEO_CAT(a, b) RODEO_CAT2(a, b)
#define RODEO_FN(name) RODEO_CAT(name, RODEO_SUFFIX)
typedef RODEO_REAL real_t;

namespace rodeo {
namespace host {

template <class Model, int INTERR, int QK>
struct SolveSimRun {
  static int run(const RodeoProblem& p, const real_t* W, const real_t* Q, const real_t* R,
                 const CommonArgs<real_t>& a_in, const real_t* z_smooth, real_t* stash, real_t* x_out,
                 const SimLoglik<real_t>& sl, cudaStream_t s) {
    FilterConsts<real_t, Model::NB, Model::P, Model::M> C;
    // per-theta prior: Q, R are device arrays (B, n_block, p, p) the kernels read per thread; one lane per theta
    constexpr bool BATCH = QK == QK_DENSE_BATCH;
    pack_consts<real_t, Model::NB, Model::P, Model::M>(W, BATCH ? nullptr : Q, BATCH ? nullptr : R, C);
    CommonArgs<real_t> a = a_in;
    if (BATCH) { a.q_batch = Q; a.r_batch = R; }
    if (p.B == 0) return RODEO_OK;
    // State-independent interrogations: covariances, gains and factors come from a cached schedule and the kernel
    // carries the block means only (rodeo_sched.cuh).  interrogate_schober with a per-theta prior scale keeps the full
    // kernels (singular filtered variance: the sign of a rounding-noise pivot is not scale invariant).
    if constexpr (INTERR != INTERR_KRAMER && !BATCH) {
      if (sim_schedule_selected(p)) {
        typedef Sched<real_t, Model, INTERR, QK> SC;
        struct Key {
          int tag, model, interr, qk, n_steps, elem;
          FilterConsts<real_t, Model::NB, Model::P, Model::M> C;
        } key;
        memset(&key, 0, sizeof(key));
        key.tag = 1; key.model = p.model_id; key.interr = INTERR; key.qk = QK; key.n_steps = p.n_steps;
        key.elem = (int)sizeof(real_t);
        memcpy(&key.C, &C, sizeof(C));
        const int N = p.n_steps;
        auto build = [&](void* t, cudaStream_t st) -> int {
          sched_forward_kernel<real_t, Model, INTERR, QK><<<1, 32, 0, st>>>(C, N, (real_t*)t);
          RODEO_CUDA_OK(cudaGetLastError());
          sched_backward_kernel<real_t, Model, INTERR, QK><<<grid_for((long long)N * Model::NB, 128), 128, 0, st>>>(
              C, N, (real_t*)t);
          RODEO_CUDA_OK(cudaGetLastError());
          g_launches += 2;
          return RODEO_OK;
        };
        const void* tab = nullptr;
        if (int rc = sched_get(&key, sizeof(key), (size_t)SC::total(N) * sizeof(real_t), build, s, &tab)) return rc;
        // (theta, block) lanes while their grid fills at most half of the resident slots, i.e. while the launch is
        // latency-bound (rodeo_sched.cuh).  Measured on B200, FitzHugh-Nagumo N = 800, log-likelihood only, one lane per
        // theta / per (theta, block): B = 8,192: 0.897 / 0.744 ms, 16,384: 0.914 / 0.824, 32,768: 1.058 / 1.117,
        // 65,536: 1.60 / 1.90 (every lane repeats the step's Philox call and the right-hand side)
        bool bl = false;
        if constexpr (SchedSimBl<real_t, Model, INTERR, QK, false>::OK) {
          typedef SchedSimBl<real_t, Model, INTERR, QK, false> SB;
          const double slots = x_out != nullptr
              ? resident_slots(solve_sim_sched_bl_kernel<real_t, Model, INTERR, QK, true>, 32,
                               SchedSimBl<real_t, Model, INTERR, QK, true>::SMEM)
              : resident_slots(solve_sim_sched_bl_kernel<real_t, Model, INTERR, QK, false>, 32, SB::SMEM);
          bl = (double)grid_for(p.B, SB::TW) <= 0.5 * slots;
          if (const char* e = getenv("RODEO_SIM_BLOCK_LANES")) bl = e[0] == '1';      // tuning / tests
          if (bl) {
            if (x_out != nullptr)
              solve_sim_sched_bl_kernel<real_t, Model, INTERR, QK, true>
                  <<<grid_for(p.B, SB::TW), 32, SchedSimBl<real_t, Model, INTERR, QK, true>::SMEM, s>>>(
                      C, a, (const real_t*)tab, z_smooth, stash, stash_ldb(p.B), x_out, sl);
            else
              solve_sim_sched_bl_kernel<real_t, Model, INTERR, QK, false>
                  <<<grid_for(p.B, SB::TW), 32, SB::SMEM, s>>>(
                      C, a, (const real_t*)tab, z_smooth, stash, stash_ldb(p.B), x_out, sl);
          }
        }
        if (bl) {
        } else if (x_out != nullptr)
          solve_sim_sched_kernel<real_t, Model, INTERR, QK, true>
              <<<grid_for(p.B, 32), 32, SchedSim<real_t, Model, INTERR, QK, true>::SMEM, s>>>(
                  C, a, (const real_t*)tab, z_smooth, stash, stash_ldb(p.B), x_out, sl);
        else
          solve_sim_sched_kernel<real_t, Model, INTERR, QK, false>
              <<<grid_for(p.B, 32), 32, SchedSim<real_t, Model, INTERR, QK, false>::SMEM, s>>>(
                  C, a, (const real_t*)tab, z_smooth, stash, stash_ldb(p.B), x_out, sl);
        g_launches++;
        RODEO_CUDA_OK(cudaGetLastError());
        return RODEO_OK;
      }
    }
    // (theta, block) lanes shorten the serial chain of a lane by n_block and multiply the warps by n_block, at the price
    // of redundant right-hand-side / Philox work per lane (and 32 % n_block idle lanes).  They win while the launch is
    // latency-bound, i.e. while all of its warps are resident at once; measured on B200 (N = 800 / 4,000):
    //   FitzHugh-Nagumo (n_block 2)  B = 2,048 / 8,192 / 32,768: 1.11 / 1.13 / 2.23 ms against 2.00 / 2.12 / 2.53 ms
    //   Lorenz63 (n_block 3)         B = 4,096 / 16,384 / 65,536: 5.9 / 23.9 / 89.8 ms against 15.0 / 15.0 / 39.1 ms
    // so: block lanes iff their grid fits the resident slots (n_block 2), half of them (n_block >= 3: 30 of 32 lanes
    // and three redundant Jacobians make its saturated throughput a quarter of the one-theta kernel's).
    bool block_lanes = false;
    if constexpr (Model::NB >= 2 && !BATCH) {
      typedef BlockLane<real_t, Model, INTERR, QK> L;
      constexpr int SMEM_BL = 16 * Model::NB * Model::P * L::PITCH * (int)sizeof(real_t);
      RODEO_CUDA_OK(cudaFuncSetAttribute(solve_sim_bl_kernel<real_t, Model, INTERR, QK>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BL));
      const double slots = resident_slots(solve_sim_bl_kernel<real_t, Model, INTERR, QK>, 32, SMEM_BL);
      block_lanes = (double)grid_for(p.B, L::TW) <= (Model::NB == 2 ? 1.0 : 0.5) * slots;
      if (const char* e = getenv("RODEO_SIM_BLOCK_LANES")) block_lanes = e[0] == '1';      // tuning experiments
    }
    if constexpr (Model::NB >= 2 && !BATCH) if (block_lanes) {
      typedef BlockLane<real_t, Model, INTERR, QK> L;
      constexpr int SMEM = 16 * Model::NB * Model::P * L::PITCH * (int)sizeof(real_t);
      solve_sim_bl_kernel<real_t, Model, INTERR, QK><<<grid_for(p.B, L::TW), 32, SMEM, s>>>(
          C, a, z_smooth, stash, stash_ldb(p.B), x_out, ObsHook<real_t>(), sl);
    }
    if (!block_lanes) {
      constexpr int SMEM = SegBuf<real_t, Fwd<real_t, Model, INTERR, QK>>::BYTES;
      RODEO_CUDA_OK(cudaFuncSetAttribute(solve_sim_kernel<real_t, Model, INTERR, QK>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
      solve_sim_kernel<real_t, Model, INTERR, QK><<<grid_for(p.B, 32), 32, SMEM, s>>>(
          C, a, z_smooth, stash, stash_ldb(p.B), x_out, ObsHook<real_t>(), sl);
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

static int solve_sim_impl(const RodeoProblem* p, const real_t* ode_weight, const real_t* prior_weight,
                          const real_t* prior_var, const real_t* ode_init, const real_t* theta, const real_t* z_interr,
                          const real_t* z_smooth, real_t* x_out, const SimLoglik<real_t>& sl, void* workspace,
                          size_t workspace_bytes, void* stream) {
  if (int rc = check_common(p)) return rc;
  if (int rc = check_ws(RODEO_OP_SOLVE_SIM, p, workspace, workspace_bytes)) return rc;
  CommonArgs<real_t> a = make_common<real_t>(*p, ode_init, theta, z_interr);
  if (p->model_id >= RODEO_MODEL_USER_BASE) {
    if (sizeof(real_t) != 8) { set_error("user (NVRTC) models are float64 only"); return RODEO_ERR_UNSUPPORTED; }
    real_t* stash = (real_t*)workspace;
    long long ldb = stash_ldb(p->B);
    ObsHook<real_t> no_obs{};        // trailing kernel parameters of the solver kernels (OBS = false)
    SimLoglik<real_t> slv = sl;
    const int smem = seg_len(nstate_of(p->n_block, p->n_bstate)) * nstate_of(p->n_block, p->n_bstate) * SEG_PITCH * (int)sizeof(real_t);
    return user_launch(*p, "solve_sim_kernel", "", (const double*)ode_weight, (const double*)prior_weight, (const double*)prior_var, p->user_wcol, p->B, smem,
                       {&a, &z_smooth, &stash, &ldb, &x_out, &no_obs, &slv}, (cudaStream_t)stream);
  }
  return dispatch_model<SolveSimRun>(*p, ode_weight, prior_weight, *p, ode_weight, prior_weight, prior_var, a, z_smooth,
                                     (real_t*)workspace, x_out, sl, (cudaStream_t)stream);
}

extern "C" int RODEO_FN(rodeo_b200_solve_sim)(const RodeoProblem* p, const real_t* ode_weight, const real_t* prior_weight,
                                        const real_t* prior_var, const real_t* ode_init, const real_t* theta,
                                        const real_t* z_interr, const real_t* z_smooth, real_t* x_out,
                                        void* workspace, size_t workspace_bytes, void* stream) {
  return solve_sim_impl(p, ode_weight, prior_weight, prior_var, ode_init, theta, z_interr, z_smooth, x_out,
                        SimLoglik<real_t>(), workspace, workspace_bytes, stream);
}

#ifdef RODEO_SOLVE_SIM_LOGLIK     /* float64 translation unit only */
extern "C" int rodeo_b200_solve_sim_loglik_f64(const RodeoProblem* p, const double* ode_weight,
                                               const double* prior_weight, const double* prior_var,
                                               const double* ode_init, const double* theta, const double* z_interr,
                                               const double* z_smooth, const int32_t* obs_ind, const double* obs_data,
                                               double noise_sd, double* loglik_out, double* x_out, void* workspace,
                                               size_t workspace_bytes, void* stream) {
  if (!p || p->n_obs < 1 || !(noise_sd > 0.0) || !loglik_out) {
    set_error("solve_sim_loglik needs n_obs >= 1, noise_sd > 0 and an output buffer");
    return RODEO_ERR_INVALID;
  }
  SimLoglik<double> sl;
  sl.obs_ind = obs_ind; sl.obs_data = obs_data; sl.n_obs = p->n_obs; sl.noise_sd = noise_sd; sl.out = loglik_out;
  return solve_sim_impl(p, ode_weight, prior_weight, prior_var, ode_init, theta, z_interr, z_smooth, x_out, sl, workspace,
                        workspace_bytes, stream);
}
#endif
