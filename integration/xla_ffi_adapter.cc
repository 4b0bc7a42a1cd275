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

// solve_mv / solve_sim / fenrir follow the same pattern; those with a history workspace declare it as a second
// result buffer of rodeo_b200_workspace_bytes(op, &p, 8) bytes so that XLA owns the allocation.
