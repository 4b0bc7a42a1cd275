/*
 * rodeo_b200 -- C ABI of the B200-native probabilistic-ODE filtering hot path.
 *
 * This is the drop-in boundary for the batched-theta path of mlysy/rodeo (reference v1.1.3).  The reference has
 * no FFI of its own (it is pure JAX); each entry point below replaces one reference *Python* function executed
 * under jax.jit(jax.vmap(...)) over a leading theta axis, and is what a jax.ffi / ctypes / cffi binding for that
 * function would call (see INTEGRATION.md for the binding stubs):
 *
 *   rodeo_b200_solve_mv_*      rodeo.solve_mv            src/rodeo/solve.py:208-302
 *   rodeo_b200_solve_sim_*     rodeo.solve_sim           src/rodeo/solve.py:125-205
 *   rodeo_b200_dalton_*        rodeo.inference.dalton    src/rodeo/inference/dalton.py:39-235
 *   rodeo_b200_fenrir_*        rodeo.inference.fenrir    src/rodeo/inference/fenrir.py:261-328
 *   rodeo_b200_basic_gather_*  Xt[searchsorted(...)] of  rodeo.inference.basic   src/rodeo/inference/basic.py:60-61
 *   rodeo_b200_ode_init_pad_*  ode_init() returned by    rodeo.utils.first_order_pad   src/rodeo/utils.py:94-96
 *
 * Conventions
 *   - every array is C-contiguous with the reference's own layout plus a leading theta axis B;
 *   - the `_f64` / `_f32` suffix is the arithmetic type of every floating-point buffer;
 *   - all buffer pointers are DEVICE pointers (cudaMalloc / torch CUDA tensor .data_ptr()) unless a parameter
 *     is documented as HOST; the `*_host` convenience wrappers take HOST buffers for everything and perform the
 *     H2D / D2H copies themselves on an internal cached arena;
 *   - calls are asynchronous and stream ordered on `stream` (a cudaStream_t passed as void*, NULL = default
 *     stream); the library never allocates on this path and keeps no pointer after return;
 *   - the caller provides `workspace` of at least rodeo_b200_workspace_bytes() bytes (may be NULL if that is 0);
 *   - return value 0 = success; non-zero = error, text via rodeo_b200_last_error() (thread local);
 *   - numerical failures (singular prior, NaN inputs) propagate as NaN/Inf exactly as in the reference -- there
 *     is no device-side exception;
 *   - there is no CPU fallback: without a CUDA device every compute entry point returns RODEO_ERR_CUDA.
 */
#ifndef RODEO_B200_H
#define RODEO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RODEO_B200_ABI_VERSION 1

/* rodeo.interrogate.interrogate_*   (src/rodeo/interrogate.py) */
enum { RODEO_INTERROGATE_KRAMER = 0, RODEO_INTERROGATE_CHKREBTII = 1, RODEO_INTERROGATE_SCHOBER = 2,
       RODEO_INTERROGATE_RODEO = 3 };
/* kalman_type (src/rodeo/solve.py:236-241).  RodeoProblem.kalman_type must be RODEO_KALMAN_STANDARD for the generic
 * entry points; the square-root form has its own entry points (rodeo_b200_solve_mv_sqrt_*, rodeo_b200_sqrt_*) */
enum { RODEO_KALMAN_STANDARD = 0, RODEO_KALMAN_SQUARE_ROOT = 1 };
/* built-in ODE right-hand sides; ids >= RODEO_MODEL_USER_BASE come from rodeo_b200_register_model_nvrtc() */
enum { RODEO_MODEL_FITZHUGH_NAGUMO = 0, RODEO_MODEL_LORENZ63 = 1, RODEO_MODEL_SECOND_ORDER_SIN = 2,
       RODEO_MODEL_HES1 = 3, RODEO_MODEL_SEIRAH = 4,
       RODEO_MODEL_PAIR_ONE_BLOCK = 5,   /* n_bmeas = 2; float64 solve_mv / dalton / fenrir only */
       RODEO_MODEL_USER_BASE = 1000 };
/* ops, for rodeo_b200_workspace_bytes() */
enum { RODEO_OP_SOLVE_MV = 0, RODEO_OP_SOLVE_SIM = 1, RODEO_OP_DALTON = 2, RODEO_OP_FENRIR = 3,
       RODEO_OP_BASIC_GATHER = 4, RODEO_OP_ODE_INIT_PAD = 5 };
enum { RODEO_OK = 0, RODEO_ERR_UNSUPPORTED = 1, RODEO_ERR_INVALID = 2, RODEO_ERR_WORKSPACE = 3,
       RODEO_ERR_CUDA = 4, RODEO_ERR_NVRTC = 5 };

/* Problem description shared by all ops (the reference's positional / keyword scalars). */
typedef struct RodeoProblem {
  int64_t B;               /* number of thetas (leading batch axis)                                        */
  int64_t particle_offset; /* global index of theta 0: keeps random streams independent of GPU sharding    */
  int32_t n_steps;         /* N                                                                            */
  int32_t n_block;         /* ode_weight.shape[0]                                                          */
  int32_t n_bstate;        /* ode_weight.shape[2]                                                          */
  int32_t n_bmeas;         /* ode_weight.shape[1]                                                          */
  int32_t n_theta;         /* len(theta)                                                                   */
  int32_t model_id;        /* RODEO_MODEL_*                                                                */
  int32_t interrogate;     /* RODEO_INTERROGATE_*                                                          */
  int32_t kalman_type;     /* RODEO_KALMAN_*                                                               */
  int32_t n_obs;           /* dalton / fenrir / basic_gather                                               */
  int32_t n_bobs;          /* obs_weight.shape[2]                                                          */
  uint32_t key[2];         /* the jax PRNG key (uint32[2]); only sampling paths read it                    */
  double t_min, t_max;
  int32_t user_wcol;       /* user (NVRTC) models only: the ODE is X[:, user_wcol] = f(X, t), i.e. W = e_user_wcol  */
  int32_t prior_batched;   /* nonzero: prior_weight / prior_var of the call are DEVICE arrays (B, n_block, p, p), one prior per
                            * theta (the reference takes any prior_pars per theta under vmap,
                            * docs/examples/parameter.md:218-236); float64 solve_mv / solve_sim / dalton / fenrir.
                            * Zero: HOST arrays (n_block, p, p) shared by every theta (see "Common inputs") */
  /* optional pointer (B, n_block), same arithmetic type as the call, or NULL: per-theta scale of the prior
   * variance, R(theta, b) = prior_var_scale[theta, b] * prior_var[b].  Covers an IBM prior whose sigma is part of
   * theta (sigma^2 R_1, src/rodeo/prior/ibm.py:84-86) without a (B, n_block, p, p) array.  A DEVICE pointer for the
   * device-pointer entry points; a HOST pointer (float64) for the `*_host` wrappers, which stage it themselves. */
  const void* prior_var_scale;
} RodeoProblem;

/* Bytes of device workspace the op needs for this problem (0 on unsupported input). */
size_t rodeo_b200_workspace_bytes(int op, const RodeoProblem* prob, int elem_bytes);

const char* rodeo_b200_last_error(void);
int rodeo_b200_abi_version(void);
/* sizeof(struct RodeoProblem) as this library was compiled: lets a binding check its own mirror of the struct */
size_t rodeo_b200_problem_sizeof(void);

/*
 * Common inputs:
 *   ode_weight   (n_block, n_bmeas, n_bstate)  HOST   W          -- tiny; copied into the kernel parameter bank
 *   prior_weight (n_block, n_bstate, n_bstate) HOST   Q
 *   prior_var    (n_block, n_bstate, n_bstate) HOST   R
 *   ode_init     (B, n_block, n_bstate)        device X0 per theta
 *   theta        (B, n_theta)                  device
 *   z_interr     optional (may be NULL) device normals for interrogate_chkrebtii,
 *                (B, n_steps, n_stream, n_block, n_bstate); n_stream = 2 for dalton (joint, marginal), else 1
 */
int rodeo_b200_solve_mv_f64(const RodeoProblem* prob, const double* ode_weight, const double* prior_weight,
                            const double* prior_var, const double* ode_init, const double* theta,
                            const double* z_interr,
                            double* mean_out /* (B, N+1, n_block, n_bstate) */,
                            double* var_out /* (B, N+1, n_block, n_bstate, n_bstate) */,
                            void* workspace, size_t workspace_bytes, void* stream);

/* z_smooth: optional (may be NULL) device normals (B, N+1, n_block, n_bstate) for the backward draws */
int rodeo_b200_solve_sim_f64(const RodeoProblem* prob, const double* ode_weight, const double* prior_weight,
                             const double* prior_var, const double* ode_init, const double* theta,
                             const double* z_interr, const double* z_smooth,
                             double* x_out /* (B, N+1, n_block, n_bstate) */,
                             void* workspace, size_t workspace_bytes, void* stream);

/*
 *   obs_ind    (n_obs) int32 device: searchsorted(linspace(t_min, t_max, N+1), obs_times), left insertion
 *   obs_data   (n_obs, n_block, n_bobs) device
 *   obs_weight (n_obs, n_block, n_bobs, n_bstate) device
 *   obs_var    (n_obs, n_block, n_bobs, n_bobs) device
 */
int rodeo_b200_dalton_f64(const RodeoProblem* prob, const double* ode_weight, const double* prior_weight,
                          const double* prior_var, const double* ode_init, const double* theta,
                          const double* z_interr,
                          const int32_t* obs_ind, const double* obs_data, const double* obs_weight,
                          const double* obs_var,
                          double* loglik_out /* (B) */,
                          void* workspace, size_t workspace_bytes, void* stream);

int rodeo_b200_fenrir_f64(const RodeoProblem* prob, const double* ode_weight, const double* prior_weight,
                          const double* prior_var, const double* ode_init, const double* theta,
                          const double* z_interr,
                          const int32_t* obs_ind, const double* obs_data, const double* obs_weight,
                          const double* obs_var,
                          double* loglik_out /* (B) */,
                          void* workspace, size_t workspace_bytes, void* stream);

/* ode_data[b, i] = Xt[b, obs_ind[i]]   (rows of n_block*n_bstate) */
int rodeo_b200_basic_gather_f64(const RodeoProblem* prob, const double* Xt /* (B, N+1, n_block, n_bstate) */,
                                const int32_t* obs_ind,
                                double* ode_data /* (B, n_obs, n_block, n_bstate) */, void* stream);

/* loglik[b] = sum_{i,k} log N(obs_data[i, k]; Xt[b, obs_ind[i], k, 0], noise_sd^2): the Gaussian measurement model of
 * the reference's parameter-inference walkthrough (docs/examples/parameter.md:192-205), fused with the gather; the
 * inner call of a pseudo-marginal MCMC step (:333-354).  obs_data: (n_obs, n_block) device. */
int rodeo_b200_gauss_obs_loglik_f64(const RodeoProblem* prob, const double* Xt, const int32_t* obs_ind,
                                    const double* obs_data, double noise_sd, double* loglik_out, void* stream);

/* The two calls above fused: one draw of the solution posterior per theta (solve_sim) and, accumulated inside the
 * backward sweep, loglik[b] = sum_{i,k} log N(obs_data[i, k]; X[b, obs_ind[i], k, 0], noise_sd^2) -- the whole
 * `logdensity_fn` of the reference's pseudo-marginal MCMC walkthrough (docs/examples/parameter.md:333-354:
 * rodeo.solve_sim, Xt[obs_ind], Gaussian log-density).  x_out may be NULL: the trajectories are then never written
 * (BASELINE configs[4]: 262,144 particles x 38 KB per iteration stay on chip).  obs_ind (n_obs) int32 device,
 * non-decreasing; obs_data (n_obs, n_block) device.  Same workspace as rodeo_b200_solve_sim_f64. */
int rodeo_b200_solve_sim_loglik_f64(const RodeoProblem* prob, const double* ode_weight, const double* prior_weight,
                                    const double* prior_var, const double* ode_init, const double* theta,
                                    const double* z_interr, const double* z_smooth, const int32_t* obs_ind,
                                    const double* obs_data, double noise_sd, double* loglik_out,
                                    double* x_out /* (B, N+1, n_block, n_bstate) or NULL */,
                                    void* workspace, size_t workspace_bytes, void* stream);

/*
 * Random-walk Metropolis-Hastings bookkeeping for many chains at once (one chain = one theta of the batched kernels):
 * the proposal and the accept / reject of the reference's pseudo-marginal step (src/rodeo/inference/pseudo_marginal.py:
 * 175-189 Gaussian proposal, :452-483 rmh_proposal.generate) around its `logdensity_fn`, which is
 * rodeo_b200_solve_sim_loglik_f64 plus the caller's prior.  position / proposal (n_chains, dim), sigma (dim): device.
 * key: HOST uint32[2] (the step's PRNG key); z (n_chains, dim) / u (n_chains) optional device arrays of injected
 * standard normals / uniforms (NULL: Philox streams keyed by (key, chain_offset + chain)).
 *   propose:  proposal = position + sigma * z
 *   accept:   log_p = new_logdensity - logdensity (NaN -> -inf); p_accept = min(1, exp(log_p)); accepted = u < p_accept;
 *             accepted chains take proposal / new_logdensity IN PLACE.
 */
int rodeo_b200_rwmh_propose_f64(int64_t n_chains, int dim, const double* position, const double* sigma,
                                const double* z, const uint32_t* key, int64_t chain_offset, double* proposal,
                                void* stream);
int rodeo_b200_rwmh_accept_f64(int64_t n_chains, int dim, double* position, double* logdensity, const double* proposal,
                               const double* new_logdensity, const double* u, const uint32_t* key, int64_t chain_offset,
                               int32_t* accepted_out, double* p_accept_out, void* stream);

/* X0[b] = [x0[b], f(x0[b], t, theta[b]), 0, ...]   x0: (B, n_block) device; X0: (B, n_block, n_bstate) device */
int rodeo_b200_ode_init_pad_f64(const RodeoProblem* prob, double t, const double* theta, const double* x0,
                                double* X0, void* stream);

/*
 * Data-adaptive solvers rodeo.inference.dalton.solve_mv / solve_sim (src/rodeo/inference/dalton.py:374-545): the
 * forward filter also conditions on the Gaussian observations.  Same arguments as solve_mv / solve_sim plus the
 * observation arrays of rodeo_b200_dalton_*; workspace of rodeo_b200_dalton_solve_workspace_bytes_*(op, prob) bytes,
 * op = RODEO_OP_SOLVE_MV or RODEO_OP_SOLVE_SIM.  float64 only.
 */
size_t rodeo_b200_dalton_solve_workspace_bytes_f64(int op, const RodeoProblem* prob);
int rodeo_b200_dalton_solve_mv_f64(const RodeoProblem* prob, const double* ode_weight, const double* prior_weight,
                                   const double* prior_var, const double* ode_init, const double* theta,
                                   const double* z_interr, const int32_t* obs_ind, const double* obs_data,
                                   const double* obs_weight, const double* obs_var, double* mean_out, double* var_out,
                                   void* workspace, size_t workspace_bytes, void* stream);
int rodeo_b200_dalton_solve_sim_f64(const RodeoProblem* prob, const double* ode_weight, const double* prior_weight,
                                    const double* prior_var, const double* ode_init, const double* theta,
                                    const double* z_interr, const double* z_smooth, const int32_t* obs_ind,
                                    const double* obs_data, const double* obs_weight, const double* obs_var,
                                    double* x_out, void* workspace, size_t workspace_bytes, void* stream);
/* rodeo.inference.fenrir.solve_mv (src/rodeo/inference/fenrir.py:404-457); same arguments as
 * rodeo_b200_dalton_solve_mv_f64; workspace of rodeo_b200_fenrir_solve_mv_workspace_bytes(prob) bytes. */
size_t rodeo_b200_fenrir_solve_mv_workspace_bytes(const RodeoProblem* prob);
int rodeo_b200_fenrir_solve_mv_f64(const RodeoProblem* prob, const double* ode_weight, const double* prior_weight,
                                   const double* prior_var, const double* ode_init, const double* theta,
                                   const double* z_interr, const int32_t* obs_ind, const double* obs_data,
                                   const double* obs_weight, const double* obs_var, double* mean_out, double* var_out,
                                   void* workspace, size_t workspace_bytes, void* stream);

/*
 * rodeo.solve_mv with kalman_type="square-root" (src/rodeo/solve.py:236-241 selecting src/rodeo/kalmantv/
 * square_root.py): prior_var_sqrt is the lower Cholesky factor of R (docs/examples/higher_order.md:108-112) and
 * var_sqrt_out receives lower-triangular factors L with var = L L^T (only L L^T is comparable across
 * implementations: QR leaves the signs free).  interrogate_kramer / schober / chkrebtii; float64.
 */
size_t rodeo_b200_solve_mv_sqrt_workspace_bytes(const RodeoProblem* prob);
int rodeo_b200_solve_mv_sqrt_f64(const RodeoProblem* prob, const double* ode_weight, const double* prior_weight,
                                 const double* prior_var_sqrt, const double* ode_init, const double* theta,
                                 const double* z_interr, double* mean_out, double* var_sqrt_out, void* workspace,
                                 size_t workspace_bytes, void* stream);

/* float32 factors (and means carried in double inside the kernel); same workspace size */
int rodeo_b200_solve_mv_sqrt_f32(const RodeoProblem* prob, const float* ode_weight, const float* prior_weight,
                                 const float* prior_var_sqrt, const float* ode_init, const float* theta,
                                 const float* z_interr, float* mean_out, float* var_sqrt_out, void* workspace,
                                 size_t workspace_bytes, void* stream);

/*
 * float32 instantiations: identical argument lists with `float` buffers (and a `float` time in ode_init_pad).
 * The reference's float width follows jax_enable_x64; its own unit tests run in float32 when tox is not used.
 */
int rodeo_b200_solve_mv_f32(const RodeoProblem* prob, const float* ode_weight, const float* prior_weight,
                            const float* prior_var, const float* ode_init, const float* theta, const float* z_interr,
                            float* mean_out, float* var_out, void* workspace, size_t workspace_bytes, void* stream);
int rodeo_b200_solve_sim_f32(const RodeoProblem* prob, const float* ode_weight, const float* prior_weight,
                             const float* prior_var, const float* ode_init, const float* theta, const float* z_interr,
                             const float* z_smooth, float* x_out, void* workspace, size_t workspace_bytes,
                             void* stream);
int rodeo_b200_dalton_f32(const RodeoProblem* prob, const float* ode_weight, const float* prior_weight,
                          const float* prior_var, const float* ode_init, const float* theta, const float* z_interr,
                          const int32_t* obs_ind, const float* obs_data, const float* obs_weight, const float* obs_var,
                          float* loglik_out, void* workspace, size_t workspace_bytes, void* stream);
int rodeo_b200_fenrir_f32(const RodeoProblem* prob, const float* ode_weight, const float* prior_weight,
                          const float* prior_var, const float* ode_init, const float* theta, const float* z_interr,
                          const int32_t* obs_ind, const float* obs_data, const float* obs_weight, const float* obs_var,
                          float* loglik_out, void* workspace, size_t workspace_bytes, void* stream);
int rodeo_b200_basic_gather_f32(const RodeoProblem* prob, const float* Xt, const int32_t* obs_ind, float* ode_data,
                                void* stream);
int rodeo_b200_ode_init_pad_f32(const RodeoProblem* prob, float t, const float* theta, const float* x0, float* X0,
                                void* stream);

/*
 * Host-buffer convenience wrapper for the headline op: every pointer is HOST memory (pinned or pageable).
 * Copies inputs to an internal cached device arena, runs rodeo_b200_dalton_f64, copies loglik back and
 * synchronises the stream.  This is the call a ctypes / cffi / jax CPU-callback binding makes.
 */
int rodeo_b200_dalton_f64_host(const RodeoProblem* prob, const double* ode_weight, const double* prior_weight,
                               const double* prior_var, const double* ode_init, const double* theta,
                               const int32_t* obs_ind, const double* obs_data, const double* obs_weight,
                               const double* obs_var, double* loglik_out);
int rodeo_b200_solve_mv_f64_host(const RodeoProblem* prob, const double* ode_weight, const double* prior_weight,
                                 const double* prior_var, const double* ode_init, const double* theta,
                                 double* mean_out, double* var_out);
/* release the cached arena of the *_host wrappers */
void rodeo_b200_host_arena_release(void);

/*
 * Batched Kalman primitives, one problem per leading index: rodeo.kalmantv.standard.{predict, update, forecast,
 * smooth_mv, smooth_sim, smooth_cond} (src/rodeo/kalmantv/standard.py:31-103, 160-255, 308-371) and
 * rodeo.utils.multivariate_normal_logpdf (src/rodeo/utils.py:60-78).  Full row-major matrices, device pointers,
 * float64, n_state in 1..7, n_meas in 1..3.  They call the same __device__ functions the fused kernels inline.
 *   update: pass mean_state_filt/var_state_filt for the update, mean_fore/var_fore for the forecast (either may be NULL);
 *   smooth: mode 0 = smooth_mv (x_next = mean_state_next, var_next = var_state_next), 1 = smooth_sim
 *           (x_next = x_state_next), 2 = smooth_cond (out_wgt = A, out_mean = b, out_var = C).
 */
int rodeo_b200_ktv_predict_f64(int64_t B, int n_state, const double* mean_state_past, const double* var_state_past,
                               const double* mean_state, const double* wgt_state, const double* var_state,
                               double* mean_state_pred, double* var_state_pred, void* stream);
int rodeo_b200_ktv_update_f64(int64_t B, int n_state, int n_meas, const double* mean_state_pred,
                              const double* var_state_pred, const double* x_meas, const double* mean_meas,
                              const double* wgt_meas, const double* var_meas, double* mean_state_filt,
                              double* var_state_filt, double* mean_fore, double* var_fore, void* stream);
int rodeo_b200_ktv_smooth_f64(int64_t B, int n_state, int mode, const double* x_next, const double* var_next,
                              const double* mean_state_filt, const double* var_state_filt,
                              const double* mean_state_pred, const double* var_state_pred, const double* wgt_state,
                              double* out_mean, double* out_var, double* out_wgt, void* stream);
int rodeo_b200_mvn_logpdf_f64(int64_t B, int n, const double* x, const double* mean, const double* cov, double* out,
                              void* stream);
/* The factor the sampling paths draw with: lower-triangular A (B, n, n) with A A^T = cov, a Cholesky factorisation that
 * zeroes the column of a non-positive pivot (positive SEMI-definite input, e.g. the singular smoothing covariances).
 * Stands where the reference calls jax.random.multivariate_normal(method='svd' / 'cholesky') (src/rodeo/solve.py:179,
 * 182-186; src/rodeo/interrogate.py:30-34): any factor with A A^T = cov gives the same distribution. */
int rodeo_b200_psd_factor_f64(int64_t B, int n, const double* cov, double* factor_out, void* stream);

/*
 * Batched square-root Kalman primitives: rodeo.kalmantv.square_root.{predict, update, forecast, smooth_mv, smooth_sim,
 * smooth_cond} (src/rodeo/kalmantv/square_root.py:30-385; rodeo.utils.add_sqrt, src/rodeo/utils.py:10-24).  Same
 * argument order as the rodeo_b200_ktv_* calls, with every variance a lower-triangular factor L (var = L L^T; full
 * row-major matrices, zero upper triangle) and the extra `var_state` = R^{1/2} the reference's smoothers take.
 * var_fore is returned as a covariance (square_root.py:343-344).  Only L L^T is comparable across implementations.
 */
int rodeo_b200_sqrt_predict_f64(int64_t B, int n_state, const double* mean_state_past, const double* var_state_past,
                                const double* mean_state, const double* wgt_state, const double* var_state,
                                double* mean_state_pred, double* var_state_pred, void* stream);
int rodeo_b200_sqrt_update_f64(int64_t B, int n_state, int n_meas, const double* mean_state_pred,
                               const double* var_state_pred, const double* x_meas, const double* mean_meas,
                               const double* wgt_meas, const double* var_meas, double* mean_state_filt,
                               double* var_state_filt, double* mean_fore, double* var_fore, void* stream);
int rodeo_b200_sqrt_smooth_f64(int64_t B, int n_state, int mode, const double* x_next, const double* var_next,
                               const double* mean_state_filt, const double* var_state_filt,
                               const double* mean_state_pred, const double* var_state_pred, const double* wgt_state,
                               const double* var_state, double* out_mean, double* out_var, double* out_wgt,
                               void* stream);

/*
 * MAGI log-density p(U_{0:N}, Z = 0 | theta) of a given trajectory under the block-diagonal Markov prior.
 * Replaces: rodeo.inference.magi_logdens (src/rodeo/inference/magi.py:6-99) AFTER its `ode_expand` call, i.e. it takes the
 * expanded solution process.  prior_weight / prior_var (n_block, n_bstate, n_bstate) HOST; ode_state
 * (B, n_steps + 1, n_block, n_bstate) device; logdens_out (B) device.  n_block <= 8, n_bstate in 2..4,
 * n_active in 1..n_bstate; kalman_type "standard".
 */
int rodeo_b200_magi_logdens_f64(int64_t B, int n_steps, int n_block, int n_bstate, int n_active,
                                const double* prior_weight, const double* prior_var, const double* ode_state,
                                double* logdens_out, void* stream);

/*
 * Register a user ODE right-hand side given as CUDA source defining `struct UserModel` with the functor interface
 * documented in rodeo_b200/csrc/rodeo_models.cuh (rodeo_b200.models.CudaOde generates it from a one-line rhs).  The
 * kernels are compiled for sm_100a with NVRTC on first use.  Replaces: an arbitrary `ode_fun` callable passed to
 * rodeo.solve_mv & co. (src/rodeo/solve.py:219).  Returns the model id to put in RodeoProblem.model_id.
 */
int rodeo_b200_register_model_nvrtc(const char* name, const char* src, int n_block, int n_bstate, int n_bmeas,
                                    int n_theta, int* model_id);

/* measured FP64 FMA throughput (TFLOP/s) of the current device: the roofline denominator bench.py reports against */
int rodeo_b200_fp64_peak_probe(int reps, double* tflops_out);

/* number of kernel launches issued by this library since load (for bench.py's gpu_launches) */
int64_t rodeo_b200_launch_count(void);

/*
 * Covariance schedules (rodeo_b200/csrc/rodeo_sched.cuh).  Under interrogate_chkrebtii / _schober / _rodeo
 * (src/rodeo/interrogate.py:13-60, 87-115: wgt_meas == 0, var_meas a function of the predicted variance) every
 * variance, gain and sampling factor of solve_sim depends on (prior, ode_weight, n_steps) only; the library computes
 * them once into a device table it owns and re-uses the table in later calls (any stream, any theta batch, any key).
 *   rodeo_b200_schedule_builds  number of tables built since load (a repeated call must not increase it)
 *   rodeo_b200_schedule_clear   drop every cached table (synchronises the device)
 * RODEO_SIM_SCHEDULE=0 in the environment sends solve_sim through the full kernels instead (A/B checks).
 */
int64_t rodeo_b200_schedule_builds(void);
void rodeo_b200_schedule_clear(void);

/*
 * Peer-memory all-gather of per-theta log-likelihoods (rodeo_b200/csrc/abi_peer.cu), one process per GPU on one node.
 * The reference has no multi-device path; this replaces the NCCL all-gather that follows a log-likelihood kernel
 *  where the step's latency matters: one kernel per rank stores its shard into every rank's gathered vector
 * through CUDA-IPC mappings (NVLink / NVSwitch peer stores), publishes a flag per peer and waits -- bounded -- for theirs.
 *   peer_region_bytes  size of a rank's region for n_total elements: [2 slots][n_total] doubles + flags
 *   peer_alloc         cudaMalloc a zeroed region and export its IPC handle (64 bytes) for the other ranks
 *   peer_open / close  map / unmap another rank's region from its handle
 *   peer_free          free a region obtained from peer_alloc
 *   peer_allgather     regions: HOST array of `world` device pointers (own region at [rank]); the caller's shard
 *                      local[0 .. n_local) lands at element `offset` of slot (epoch & 1) of every region.  epoch >= 1 and
 *                      increases by one per call on every rank -- or epoch = 0 and epoch_counter a DEVICE counter (zeroed
 *                      by the caller) that the kernel itself advances, so that the launch can be replayed from a CUDA
 *                      graph; the result (own region, that slot) must be consumed on `stream` before the next call.
 *                      *status (a DEVICE int the caller zeroes) becomes 1 if a peer's flag did not arrive within
 *                      spin_limit polls (0 = default, about a second): the kernel never hangs.
 */
size_t rodeo_b200_peer_region_bytes(long long n_total, int world);
int rodeo_b200_peer_alloc(size_t bytes, void** ptr, unsigned char* handle64);
int rodeo_b200_peer_open(const unsigned char* handle64, void** ptr);
int rodeo_b200_peer_close(void* ptr);
int rodeo_b200_peer_free(void* ptr);
int rodeo_b200_peer_allgather_f64(const double* local, long long n_local, long long offset, long long n_total, int rank,
                                  int world, void* const* regions, unsigned epoch, unsigned* epoch_counter,
                                  unsigned long long spin_limit, int* status, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* RODEO_B200_H */
