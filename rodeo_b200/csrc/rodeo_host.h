// Host-side plumbing shared by the C-ABI translation units: error reporting, packing of (W, Q, R) into the
// kernel-parameter struct, and the (model, interrogation, Q-structure) dispatch over the ahead-of-time
// instantiations.  One translation unit per op keeps nvcc compile times parallel.
#pragma once
#include <cstdlib>
#include <cuda_runtime.h>

#include <atomic>
#include <functional>
#include <cstdarg>
#include <cstdio>
#include <cstring>

#include "../../include/rodeo_b200.h"
#include "rodeo_kernels.cuh"

namespace rodeo {
namespace host {

void set_error(const char* fmt, ...);
extern std::atomic<long long> g_launches;

// Schedule cache (abi_common.cu): *table = the device table for `key`, built by `build(table, s)` on first use.
int sched_get(const void* key, size_t key_bytes, size_t bytes, const std::function<int(void*, cudaStream_t)>& build,
              cudaStream_t s, const void** table);
long long sched_builds();
void sched_clear();

inline int cuda_fail(cudaError_t e, const char* what) {
  set_error("%s: %s", what, cudaGetErrorString(e));
  return RODEO_ERR_CUDA;
}

#define RODEO_CUDA_OK(call)                                             \
  do {                                                                  \
    cudaError_t e__ = (call);                                           \
    if (e__ != cudaSuccess) return ::rodeo::host::cuda_fail(e__, #call); \
  } while (0)

// The structured instantiation (QK_UNIT_UPPER) needs, exactly:
//   Q unit upper triangular for every block (every IBM prior is: src/rodeo/prior/ibm.py:55-57), and
//   W = e_wcol (one row, a single 1) for every block (what first_order_pad builds: src/rodeo/utils.py:100-101).
// Anything else runs the dense instantiation.
template <typename T>
inline int detect_structure(const T* Q, const T* W, int nb, int p, int m, int wcol) {
  for (int b = 0; b < nb; ++b)
    for (int i = 0; i < p; ++i)
      for (int j = 0; j <= i; ++j) {
        T v = Q[(b * p + i) * p + j];
        if (i == j ? (v != T(1)) : (v != T(0))) return QK_DENSE;
      }
  if (m != 1) return QK_DENSE;
  for (int b = 0; b < nb; ++b)
    for (int j = 0; j < p; ++j)
      if (W[b * p + j] != (j == wcol ? T(1) : T(0))) return QK_DENSE;
  return QK_UNIT_UPPER;
}

// Q == nullptr (batched prior: Q, R are per-theta device arrays the kernels read themselves) packs W only
template <typename T, int NB, int P, int M>
inline void pack_consts(const T* W, const T* Q, const T* R, FilterConsts<T, NB, P, M>& C) {
  if (Q == nullptr || R == nullptr) {
    memset(&C, 0, sizeof(C));
    for (int b = 0; b < NB; ++b)
      for (int r = 0; r < M; ++r)
        for (int j = 0; j < P; ++j) C.W[b][r][j] = W[(b * M + r) * P + j];
    return;
  }
  for (int b = 0; b < NB; ++b) {
    for (int i = 0; i < P; ++i)
      for (int j = 0; j < P; ++j) C.Q[b][i][j] = Q[(b * P + i) * P + j];
    int k = 0;
    for (int i = 0; i < P; ++i)
      for (int j = i; j < P; ++j) C.R[b][k++] = R[(b * P + i) * P + j];
    for (int r = 0; r < M; ++r)
      for (int j = 0; j < P; ++j) C.W[b][r][j] = W[(b * M + r) * P + j];
  }
}

template <typename T>
inline CommonArgs<T> make_common(const RodeoProblem& p, const T* ode_init, const T* theta, const T* z_interr) {
  CommonArgs<T> a;
  a.B = p.B; a.particle_offset = p.particle_offset; a.n_steps = p.n_steps;
  a.t_min = (T)p.t_min; a.t_max = (T)p.t_max;
  a.theta = theta; a.ode_init = ode_init; a.key0 = p.key[0]; a.key1 = p.key[1]; a.z_interr = z_interr;
  a.r_scale = (const T*)p.prior_var_scale;
  a.dalton_geometry = 0;
  a.q_batch = nullptr; a.r_batch = nullptr;
  return a;
}

template <class Model>
inline bool dims_match(const RodeoProblem& p) {
  return p.n_block == Model::NB && p.n_bstate == Model::P && p.n_bmeas == Model::M && p.n_theta == Model::NTHETA;
}

inline int check_common(const RodeoProblem* p) {
  if (!p) { set_error("RodeoProblem is NULL"); return RODEO_ERR_INVALID; }
  if (p->B < 0 || p->n_steps < 1) { set_error("need B >= 0 and n_steps >= 1 (B=%lld, n_steps=%d)", (long long)p->B, p->n_steps); return RODEO_ERR_INVALID; }
  if (p->kalman_type != RODEO_KALMAN_STANDARD) {
    // reference src/rodeo/solve.py:236-241 raises NotImplementedError for unknown kalman_type
    set_error("kalman_type %d is not built (only \"standard\")", p->kalman_type);
    return RODEO_ERR_UNSUPPORTED;
  }
  if (p->prior_batched && p->model_id >= RODEO_MODEL_USER_BASE) {
    set_error("a per-theta prior (prior_batched) is not supported for user (NVRTC) models");
    return RODEO_ERR_UNSUPPORTED;
  }
  return RODEO_OK;
}

// solve_sim runs over a cached covariance schedule (rodeo_sched.cuh) -- and then keeps a history of block MEANS only --
// under the state-independent interrogations of built-in models with a shared prior; interrogate_schober with a
// per-theta prior scale keeps the full kernels (singular filtered variance: the sign of a rounding-noise pivot is not
// scale invariant).  One predicate for the dispatch (abi_solve_sim.cu) and the workspace size (abi_common.cu).
inline bool sim_schedule_selected(const RodeoProblem& p) {
  if (p.interrogate == RODEO_INTERROGATE_KRAMER || p.prior_batched || p.model_id >= RODEO_MODEL_USER_BASE) return false;
  if (p.interrogate == RODEO_INTERROGATE_SCHOBER && p.prior_var_scale != nullptr) return false;
  if (const char* e = getenv("RODEO_SIM_SCHEDULE")) { if (e[0] == '0') return false; }
  return true;
}

inline size_t round_up(size_t x, size_t m) { return (x + m - 1) / m * m; }
// leading dimension of the theta-innermost stash
inline long long stash_ldb(long long B) { return (long long)round_up((size_t)(B > 0 ? B : 1), 32); }

// ---- ahead-of-time instantiation set -----------------------------------------------------------------------------
// X(ModelType, RODEO_MODEL_id)
#if defined(RODEO_FAST_BUILD) && defined(RODEO_FAST_SECOND_ORDER)      /* development builds: one or two models */
#define RODEO_AOT_MODELS(X) X(FitzHughNagumo, RODEO_MODEL_FITZHUGH_NAGUMO) X(SecondOrderSin, RODEO_MODEL_SECOND_ORDER_SIN)
#elif defined(RODEO_FAST_BUILD)
#define RODEO_AOT_MODELS(X) X(FitzHughNagumo, RODEO_MODEL_FITZHUGH_NAGUMO)
#else
#define RODEO_AOT_MODELS(X)                              \
  X(FitzHughNagumo, RODEO_MODEL_FITZHUGH_NAGUMO)         \
  X(Lorenz63, RODEO_MODEL_LORENZ63)                      \
  X(SecondOrderSin, RODEO_MODEL_SECOND_ORDER_SIN)        \
  X(Hes1, RODEO_MODEL_HES1)                              \
  X(Seirah, RODEO_MODEL_SEIRAH)
#endif
// n_bmeas = 2 (one block, two measured variables): instantiated by the translation units that define RODEO_WIDE_MODELS
// (float64 solve_mv, dalton, fenrir), dense Q / W only
#if defined(RODEO_WIDE_MODELS) && !defined(RODEO_FAST_BUILD)
#define RODEO_AOT_MODELS_WIDE(X) X(PairOneBlock, RODEO_MODEL_PAIR_ONE_BLOCK)
#else
#define RODEO_AOT_MODELS_WIDE(X)
#endif

// Calls FN<Model, INTERR, QK>::run(args...) for the runtime (model_id, interr, qk); RODEO_ERR_UNSUPPORTED otherwise.
template <template <class, int, int> class FN, class Model, int INTERR, typename... A>
inline int dispatch_qk(int qk, A&&... args) {
  if constexpr (Model::M != 1) {
    // vector measurements per block: the dense instantiation only (the structured one needs W = e_wcol, one row)
    if (qk == QK_DENSE_BATCH) {
      set_error("a per-theta prior (prior_batched) is compiled for n_bmeas = 1 models only");
      return RODEO_ERR_UNSUPPORTED;
    }
    return FN<Model, INTERR, QK_DENSE>::run(static_cast<A&&>(args)...);
  } else {
  if (qk == QK_DENSE_BATCH) {
#ifdef RODEO_PRIOR_BATCH     /* translation units that instantiate the per-theta-prior kernels */
    return FN<Model, INTERR, QK_DENSE_BATCH>::run(static_cast<A&&>(args)...);
#else
    set_error("a per-theta prior (RodeoProblem.prior_batched) is compiled for the float64 solve_mv, solve_sim, dalton "
              "and fenrir entry points only");
    return RODEO_ERR_UNSUPPORTED;
#endif
  }
  if (qk == QK_UNIT_UPPER) return FN<Model, INTERR, QK_UNIT_UPPER>::run(static_cast<A&&>(args)...);
  return FN<Model, INTERR, QK_DENSE>::run(static_cast<A&&>(args)...);
  }
}
template <template <class, int, int> class FN, class Model, typename... A>
inline int dispatch_interr(int interr, int qk, A&&... args) {
  switch (interr) {
    case RODEO_INTERROGATE_KRAMER: return dispatch_qk<FN, Model, INTERR_KRAMER>(qk, static_cast<A&&>(args)...);
    case RODEO_INTERROGATE_CHKREBTII: return dispatch_qk<FN, Model, INTERR_CHKREBTII>(qk, static_cast<A&&>(args)...);
    case RODEO_INTERROGATE_SCHOBER: return dispatch_qk<FN, Model, INTERR_SCHOBER>(qk, static_cast<A&&>(args)...);
    case RODEO_INTERROGATE_RODEO: return dispatch_qk<FN, Model, INTERR_RODEO>(qk, static_cast<A&&>(args)...);
  }
  set_error("unknown interrogate id %d", interr);
  return RODEO_ERR_UNSUPPORTED;
}
// W and Q are the host arrays of the call; the structure test needs the model's WCOL, so it happens here.
template <template <class, int, int> class FN, typename R, typename... A>
inline int dispatch_model(const RodeoProblem& p, const R* W, const R* Q, A&&... args) {
  switch (p.model_id) {
#define RODEO_CASE(MODEL, ID)                                                                                 \
  case ID:                                                                                                    \
    if (!dims_match<MODEL>(p)) {                                                                              \
      set_error(#MODEL " expects (n_block,n_bstate,n_bmeas,n_theta)=(%d,%d,%d,%d), got (%d,%d,%d,%d)",        \
                MODEL::NB, MODEL::P, MODEL::M, MODEL::NTHETA, p.n_block, p.n_bstate, p.n_bmeas, p.n_theta);   \
      return RODEO_ERR_INVALID;                                                                               \
    }                                                                                                         \
    return dispatch_interr<FN, MODEL>(p.interrogate,                                                          \
                                      p.prior_batched ? QK_DENSE_BATCH :                                      \
                                      (W && Q) ? detect_structure<R>(Q, W, MODEL::NB, MODEL::P, MODEL::M, \
                                                                          MODEL::WCOL) : QK_DENSE,             \
                                      static_cast<A&&>(args)...);
    RODEO_AOT_MODELS(RODEO_CASE)
    RODEO_AOT_MODELS_WIDE(RODEO_CASE)
#undef RODEO_CASE
  }
  set_error("model id %d is not compiled into the library (register it with rodeo_b200_register_model_nvrtc)", p.model_id);
  return RODEO_ERR_UNSUPPORTED;
}

inline unsigned grid_for(long long B, int block) { return (unsigned)((B + block - 1) / block); }

inline int sm_count() {
  static int n_sm = 0;
  if (n_sm == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n_sm < 1) n_sm = 148;
  }
  return n_sm;
}

// CTAs of `kernel` that are resident on the whole GPU at once (registers and dynamic shared memory taken into account)
template <class K>
inline double resident_slots(K kernel, int block, size_t smem) {
  const int n_sm = sm_count();
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, block, smem) != cudaSuccess || per_sm < 1 || n_sm < 1)
    return 12.0 * 148.0;
  return (double)per_sm * n_sm;
}

}  // namespace host
}  // namespace rodeo

#include <vector>
namespace rodeo {
namespace host {
// NVRTC path for user models (abi_nvrtc.cu): compile-on-first-use of the same kernel templates, driver-API launch.
// `args` point at the kernel parameters after the leading FilterConsts; `extra_targs` are trailing template arguments
// (", 1" for the NOBS of dalton / fenrir).
int user_launch(const RodeoProblem& p, const char* kernel, const char* extra_targs, const double* W, const double* Q,
                const double* R, int wcol, long long threads, int smem, std::vector<void*> args, cudaStream_t s);
int user_launch_raw(int model_id, const char* expr, long long threads, int block, std::vector<void*> args,
                    cudaStream_t s);

}  // namespace host
}  // namespace rodeo
