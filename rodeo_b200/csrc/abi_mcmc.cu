// Random-walk Metropolis-Hastings bookkeeping for MANY chains, on the device: the two O(n_chains x dim) pieces of the
// reference's pseudo-marginal step (src/rodeo/inference/pseudo_marginal.py:452-483 `rmh_proposal.generate`; the
// Gaussian proposal of :175-189; blackjax's static_binomial_sampling / compute_asymmetric_acceptance_ratio for the
// accept/reject) around the heavy `logdensity_fn` call (rodeo_b200_solve_sim_loglik_f64).  One thread per chain.
// Random numbers: Philox keyed by (key, chain, stream tag), or injected (testing hook).
#include "rodeo_host.h"

namespace rodeo {
namespace host {

enum : unsigned { TAG_MCMC_PROPOSAL = 0x500u, TAG_MCMC_ACCEPT = 0x600u };

// proposal[c, :] = position[c, :] + sigma[:] * z[c, :]        (pseudo_marginal.py:175-189 with a diagonal sigma)
__global__ void rwmh_propose_kernel(i64 C, int d, const double* __restrict__ pos, const double* __restrict__ sigma,
                                    const double* __restrict__ z, unsigned k0, unsigned k1, i64 chain_offset,
                                    double* __restrict__ prop) {
  const i64 c = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  for (int j0 = 0; j0 < d; j0 += 4) {
    double q[4];
    if (z == nullptr) {
      unsigned r[4];
      philox_call(k0, k1, chain_offset + c, j0 >> 2, TAG_MCMC_PROPOSAL, 0, r);
      normal_pair_exact(r[0], r[1], q[0], q[1]);
      normal_pair_exact(r[2], r[3], q[2], q[3]);
    }
    for (int j = j0; j < d && j < j0 + 4; ++j) {
      const double zz = z != nullptr ? z[c * d + j] : q[j - j0];
      prop[c * d + j] = fma(sigma[j], zz, pos[c * d + j]);
    }
  }
}

// log_p = new_logd - logd (symmetric proposal; NaN -> -inf), p_accept = min(1, exp(log_p)), accept iff u < p_accept;
// accepted chains take the proposal's position and log-density IN PLACE
__global__ void rwmh_accept_kernel(i64 C, int d, double* __restrict__ pos, double* __restrict__ logd,
                                   const double* __restrict__ prop, const double* __restrict__ new_logd,
                                   const double* __restrict__ u_in, unsigned k0, unsigned k1, i64 chain_offset,
                                   int* __restrict__ accepted, double* __restrict__ p_accept) {
  const i64 c = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double lp = new_logd[c] - logd[c];
  if (lp != lp) lp = -Lim<double>::inf();
  const double pa = fmin(exp(lp), 1.0);
  double u;
  if (u_in != nullptr) u = u_in[c];
  else {
    unsigned r[4];
    philox_call(k0, k1, chain_offset + c, 0, TAG_MCMC_ACCEPT, 0, r);
    u = ((double)r[0] * 4294967296.0 + (double)r[1]) * (1.0 / 18446744073709551616.0);      // [0, 1)
  }
  const bool acc = u < pa;
  accepted[c] = acc ? 1 : 0;
  p_accept[c] = pa;
  if (acc) {
    logd[c] = new_logd[c];
    for (int j = 0; j < d; ++j) pos[c * d + j] = prop[c * d + j];
  }
}

}  // namespace host
}  // namespace rodeo

using namespace rodeo;
using namespace rodeo::host;

extern "C" {

int rodeo_b200_rwmh_propose_f64(int64_t n_chains, int dim, const double* position, const double* sigma,
                                const double* z, const uint32_t* key, int64_t chain_offset, double* proposal,
                                void* stream) {
  if (n_chains < 0 || dim < 1 || !position || !sigma || !proposal || (!z && !key)) {
    set_error("rwmh_propose: bad arguments");
    return RODEO_ERR_INVALID;
  }
  if (n_chains == 0) return RODEO_OK;
  rwmh_propose_kernel<<<grid_for(n_chains, 128), 128, 0, (cudaStream_t)stream>>>(
      n_chains, dim, position, sigma, z, key ? key[0] : 0u, key ? key[1] : 0u, chain_offset, proposal);
  g_launches++;
  RODEO_CUDA_OK(cudaGetLastError());
  return RODEO_OK;
}

int rodeo_b200_rwmh_accept_f64(int64_t n_chains, int dim, double* position, double* logdensity, const double* proposal,
                               const double* new_logdensity, const double* u, const uint32_t* key, int64_t chain_offset,
                               int32_t* accepted_out, double* p_accept_out, void* stream) {
  if (n_chains < 0 || dim < 1 || !position || !logdensity || !proposal || !new_logdensity || !accepted_out ||
      !p_accept_out || (!u && !key)) {
    set_error("rwmh_accept: bad arguments");
    return RODEO_ERR_INVALID;
  }
  if (n_chains == 0) return RODEO_OK;
  rwmh_accept_kernel<<<grid_for(n_chains, 128), 128, 0, (cudaStream_t)stream>>>(
      n_chains, dim, position, logdensity, proposal, new_logdensity, u, key ? key[0] : 0u, key ? key[1] : 0u,
      chain_offset, accepted_out, p_accept_out);
  g_launches++;
  RODEO_CUDA_OK(cudaGetLastError());
  return RODEO_OK;
}

}  // extern "C"
