// XLA FFI adapter over the rodeo_b200 C ABI  --  NOT BUILT IN THIS IMAGE (no jaxlib, no xla/ffi/api/*.h).
//
// What a rodeo maintainer adds so that jax.jit / jax.vmap over theta dispatch to the sm_100a kernels: one handler
// per entry point of include/rodeo_b200.h.  The handler is a thin argument shuffle; every buffer XLA hands over is
// already a device pointer in the layout the C ABI expects, and the stream is XLA's.
//
//   g++ -std=c++17 -shared -fPIC -I$(python -c "import jax; print(jax.ffi.include_dir())") \
//       -Iinclude integration/xla_ffi_adapter.cc -Lrodeo_b200 -lrodeo_b200 -o rodeo_b200_xla.so
#include <cstdint>

#include "xla/ffi/api/ffi.h"

#include "rodeo_b200.h"

namespace ffi = xla::ffi;

// dalton: operands theta (B, n_theta), ode_init (B, nb, p), obs_ind (n_obs), obs_data, obs_weight, obs_var;
// attributes carry the static problem description and the (tiny, static under jit) W, Q, R as flat spans.
static ffi::Error DaltonImpl(cudaStream_t stream, ffi::Buffer<ffi::F64> theta, ffi::Buffer<ffi::F64> ode_init,
                             ffi::Buffer<ffi::S32> obs_ind, ffi::Buffer<ffi::F64> obs_data,
                             ffi::Buffer<ffi::F64> obs_weight, ffi::Buffer<ffi::F64> obs_var,
                             ffi::Span<const double> W, ffi::Span<const double> Q, ffi::Span<const double> R,
                             int32_t model_id, int32_t interrogate, int32_t n_steps, double t_min, double t_max,
                             ffi::ResultBuffer<ffi::F64> loglik) {
  RodeoProblem p{};
  const auto td = theta.dimensions();
  const auto xd = ode_init.dimensions();
  const auto wd = obs_weight.dimensions();
  p.B = td[0]; p.n_theta = (int32_t)td[1]; p.n_block = (int32_t)xd[1]; p.n_bstate = (int32_t)xd[2];
  p.n_bmeas = (int32_t)(W.size() / (xd[1] * xd[2]));
  p.n_steps = n_steps; p.model_id = model_id; p.interrogate = interrogate; p.kalman_type = RODEO_KALMAN_STANDARD;
  p.n_obs = (int32_t)wd[0]; p.n_bobs = (int32_t)wd[2]; p.t_min = t_min; p.t_max = t_max;
  const int rc = rodeo_b200_dalton_f64(&p, W.begin(), Q.begin(), R.begin(), ode_init.typed_data(), theta.typed_data(),
                                       nullptr, obs_ind.typed_data(), obs_data.typed_data(), obs_weight.typed_data(),
                                       obs_var.typed_data(), loglik->typed_data(), nullptr, 0, stream);
  return rc == 0 ? ffi::Error::Success() : ffi::Error::Internal(rodeo_b200_last_error());
}

XLA_FFI_DEFINE_HANDLER_SYMBOL(RodeoB200Dalton, DaltonImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()   // theta
                                  .Arg<ffi::Buffer<ffi::F64>>()   // ode_init
                                  .Arg<ffi::Buffer<ffi::S32>>()   // obs_ind
                                  .Arg<ffi::Buffer<ffi::F64>>()   // obs_data
                                  .Arg<ffi::Buffer<ffi::F64>>()   // obs_weight
                                  .Arg<ffi::Buffer<ffi::F64>>()   // obs_var
                                  .Attr<ffi::Span<const double>>("W")
                                  .Attr<ffi::Span<const double>>("Q")
                                  .Attr<ffi::Span<const double>>("R")
                                  .Attr<int32_t>("model_id")
                                  .Attr<int32_t>("interrogate")
                                  .Attr<int32_t>("n_steps")
                                  .Attr<double>("t_min")
                                  .Attr<double>("t_max")
                                  .Ret<ffi::Buffer<ffi::F64>>());   // loglik (B)

// ---- ops with a history workspace: it is declared as an extra RESULT buffer (uint8, rodeo_b200_workspace_bytes(op, &p,
// 8) bytes, computed by the Python side at trace time from the static shapes), so that XLA owns the allocation and the
// library still never allocates.
static RodeoProblem Describe(const ffi::Buffer<ffi::F64>& theta, const ffi::Buffer<ffi::F64>& ode_init, size_t w_size,
                             int32_t model_id, int32_t interrogate, int32_t n_steps, double t_min, double t_max) {
  RodeoProblem p{};
  const auto td = theta.dimensions();
  const auto xd = ode_init.dimensions();
  p.B = td[0]; p.n_theta = (int32_t)td[1]; p.n_block = (int32_t)xd[1]; p.n_bstate = (int32_t)xd[2];
  p.n_bmeas = (int32_t)(w_size / (xd[1] * xd[2]));
  p.n_steps = n_steps; p.model_id = model_id; p.interrogate = interrogate; p.kalman_type = RODEO_KALMAN_STANDARD;
  p.t_min = t_min; p.t_max = t_max;
  return p;
}
static ffi::Error Status(int rc) {
  return rc == 0 ? ffi::Error::Success() : ffi::Error::Internal(rodeo_b200_last_error());
}

// rodeo.solve_mv (src/rodeo/solve.py:208-302): results mean (B, N+1, nb, p), var (B, N+1, nb, p, p), workspace
static ffi::Error SolveMvImpl(cudaStream_t stream, ffi::Buffer<ffi::F64> theta, ffi::Buffer<ffi::F64> ode_init,
                              ffi::Span<const double> W, ffi::Span<const double> Q, ffi::Span<const double> R,
                              int32_t model_id, int32_t interrogate, int32_t n_steps, double t_min, double t_max,
                              ffi::ResultBuffer<ffi::F64> mean, ffi::ResultBuffer<ffi::F64> var,
                              ffi::ResultBuffer<ffi::U8> workspace) {
  RodeoProblem p = Describe(theta, ode_init, W.size(), model_id, interrogate, n_steps, t_min, t_max);
  return Status(rodeo_b200_solve_mv_f64(&p, W.begin(), Q.begin(), R.begin(), ode_init.typed_data(), theta.typed_data(),
                                        nullptr, mean->typed_data(), var->typed_data(), workspace->typed_data(),
                                        workspace->size_bytes(), stream));
}

// rodeo.solve_sim (src/rodeo/solve.py:125-205): operand key uint32[2] is read on the host side into attributes
static ffi::Error SolveSimImpl(cudaStream_t stream, ffi::Buffer<ffi::F64> theta, ffi::Buffer<ffi::F64> ode_init,
                               ffi::Span<const double> W, ffi::Span<const double> Q, ffi::Span<const double> R,
                               int32_t model_id, int32_t interrogate, int32_t n_steps, double t_min, double t_max,
                               int64_t key0, int64_t key1, int64_t particle_offset,
                               ffi::ResultBuffer<ffi::F64> x, ffi::ResultBuffer<ffi::U8> workspace) {
  RodeoProblem p = Describe(theta, ode_init, W.size(), model_id, interrogate, n_steps, t_min, t_max);
  p.key[0] = (uint32_t)key0; p.key[1] = (uint32_t)key1; p.particle_offset = particle_offset;
  return Status(rodeo_b200_solve_sim_f64(&p, W.begin(), Q.begin(), R.begin(), ode_init.typed_data(), theta.typed_data(),
                                         nullptr, nullptr, x->typed_data(), workspace->typed_data(),
                                         workspace->size_bytes(), stream));
}

// rodeo.inference.fenrir (src/rodeo/inference/fenrir.py:261-328)
static ffi::Error FenrirImpl(cudaStream_t stream, ffi::Buffer<ffi::F64> theta, ffi::Buffer<ffi::F64> ode_init,
                             ffi::Buffer<ffi::S32> obs_ind, ffi::Buffer<ffi::F64> obs_data,
                             ffi::Buffer<ffi::F64> obs_weight, ffi::Buffer<ffi::F64> obs_var,
                             ffi::Span<const double> W, ffi::Span<const double> Q, ffi::Span<const double> R,
                             int32_t model_id, int32_t interrogate, int32_t n_steps, double t_min, double t_max,
                             ffi::ResultBuffer<ffi::F64> loglik, ffi::ResultBuffer<ffi::U8> workspace) {
  RodeoProblem p = Describe(theta, ode_init, W.size(), model_id, interrogate, n_steps, t_min, t_max);
  const auto wd = obs_weight.dimensions();
  p.n_obs = (int32_t)wd[0]; p.n_bobs = (int32_t)wd[2];
  return Status(rodeo_b200_fenrir_f64(&p, W.begin(), Q.begin(), R.begin(), ode_init.typed_data(), theta.typed_data(),
                                      nullptr, obs_ind.typed_data(), obs_data.typed_data(), obs_weight.typed_data(),
                                      obs_var.typed_data(), loglik->typed_data(), workspace->typed_data(),
                                      workspace->size_bytes(), stream));
}

#define RODEO_COMMON_ATTRS()                       \
  .Attr<ffi::Span<const double>>("W")              \
      .Attr<ffi::Span<const double>>("Q")          \
      .Attr<ffi::Span<const double>>("R")          \
      .Attr<int32_t>("model_id")                   \
      .Attr<int32_t>("interrogate")                \
      .Attr<int32_t>("n_steps")                    \
      .Attr<double>("t_min")                       \
      .Attr<double>("t_max")

XLA_FFI_DEFINE_HANDLER_SYMBOL(RodeoB200SolveMv, SolveMvImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()   // theta
                                  .Arg<ffi::Buffer<ffi::F64>>()   // ode_init
                                  RODEO_COMMON_ATTRS()
                                  .Ret<ffi::Buffer<ffi::F64>>()   // mean
                                  .Ret<ffi::Buffer<ffi::F64>>()   // var
                                  .Ret<ffi::Buffer<ffi::U8>>());  // workspace

XLA_FFI_DEFINE_HANDLER_SYMBOL(RodeoB200SolveSim, SolveSimImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  RODEO_COMMON_ATTRS()
                                  .Attr<int64_t>("key0")
                                  .Attr<int64_t>("key1")
                                  .Attr<int64_t>("particle_offset")
                                  .Ret<ffi::Buffer<ffi::F64>>()   // x
                                  .Ret<ffi::Buffer<ffi::U8>>());  // workspace

XLA_FFI_DEFINE_HANDLER_SYMBOL(RodeoB200Fenrir, FenrirImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()   // theta
                                  .Arg<ffi::Buffer<ffi::F64>>()   // ode_init
                                  .Arg<ffi::Buffer<ffi::S32>>()   // obs_ind
                                  .Arg<ffi::Buffer<ffi::F64>>()   // obs_data
                                  .Arg<ffi::Buffer<ffi::F64>>()   // obs_weight
                                  .Arg<ffi::Buffer<ffi::F64>>()   // obs_var
                                  RODEO_COMMON_ATTRS()
                                  .Ret<ffi::Buffer<ffi::F64>>()   // loglik
                                  .Ret<ffi::Buffer<ffi::U8>>());  // workspace
