// C-ABI entry points that are not tied to one op: error text, versioning, workspace sizing, the small
// gather / initial-value kernels.  See include/rodeo_b200.h.
#include "rodeo_host.h"

#include <list>
#include <mutex>
#include <string>

namespace rodeo {
namespace host {

static thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// ---- schedule cache (rodeo_sched.cuh) --------------------------------------------------------------------------------
// Covariance schedules depend on (device, instantiation, n_steps, W, Q, R) only.  They are built once on the stream of
// the first call that needs them and kept in library-owned device memory (a few hundred KB each); every later call,
// on any stream, waits on the build's event and reads the table.  Least-recently-used eviction at 16 entries.
namespace {
struct SchedEntry {
  std::string key;
  void* dev = nullptr;
  cudaEvent_t ready = nullptr;
};
std::mutex g_sched_mu;
std::list<SchedEntry> g_sched;      // front = most recently used
constexpr size_t SCHED_MAX = 16;
std::atomic<long long> g_sched_builds{0};
}  // namespace

long long sched_builds() { return g_sched_builds.load(); }

void sched_clear() {
  std::lock_guard<std::mutex> lk(g_sched_mu);
  if (!g_sched.empty()) cudaDeviceSynchronize();
  for (auto& e : g_sched) { cudaFree(e.dev); cudaEventDestroy(e.ready); }
  g_sched.clear();
}

int sched_get(const void* key, size_t key_bytes, size_t bytes, const std::function<int(void*, cudaStream_t)>& build,
              cudaStream_t s, const void** table) {
  int dev = 0;
  RODEO_CUDA_OK(cudaGetDevice(&dev));
  std::string k((const char*)&dev, sizeof(dev));
  k.append((const char*)&bytes, sizeof(bytes));
  k.append((const char*)key, key_bytes);
  std::lock_guard<std::mutex> lk(g_sched_mu);
  for (auto it = g_sched.begin(); it != g_sched.end(); ++it)
    if (it->key == k) {
      g_sched.splice(g_sched.begin(), g_sched, it);
      RODEO_CUDA_OK(cudaStreamWaitEvent(s, it->ready, 0));
      *table = it->dev;
      return RODEO_OK;
    }
  if (g_sched.size() >= SCHED_MAX) {
    // a kernel on another stream may still be reading the victim: drain the device before freeing it (rare path)
    RODEO_CUDA_OK(cudaDeviceSynchronize());
    cudaFree(g_sched.back().dev);
    cudaEventDestroy(g_sched.back().ready);
    g_sched.pop_back();
  }
  SchedEntry e;
  e.key = std::move(k);
  RODEO_CUDA_OK(cudaMalloc(&e.dev, bytes));
  cudaError_t ce = cudaEventCreateWithFlags(&e.ready, cudaEventDisableTiming);
  if (ce != cudaSuccess) { cudaFree(e.dev); return cuda_fail(ce, "cudaEventCreateWithFlags"); }
  int rc = build(e.dev, s);
  if (rc == RODEO_OK) {
    ce = cudaEventRecord(e.ready, s);
    if (ce != cudaSuccess) rc = cuda_fail(ce, "cudaEventRecord");
  }
  if (rc != RODEO_OK) {
    cudaStreamSynchronize(s);
    cudaFree(e.dev); cudaEventDestroy(e.ready);
    return rc;
  }
  g_sched_builds++;
  *table = e.dev;
  g_sched.push_front(std::move(e));
  return RODEO_OK;
}

template <typename T>
__global__ void gather_rows_kernel(i64 B, int n_rows_in, int n_obs, int row, const int* __restrict__ ind,
                                   const T* __restrict__ in, T* __restrict__ out) {
  // out[b, i, :] = in[b, ind[i], :]
  const i64 total = B * n_obs * row;
  for (i64 k = (i64)blockIdx.x * blockDim.x + threadIdx.x; k < total; k += (i64)gridDim.x * blockDim.x) {
    const int c = (int)(k % row);
    const i64 bi = k / row;
    const int i = (int)(bi % n_obs);
    const i64 b = bi / n_obs;
    int src = ind[i];
    src = src < 0 ? 0 : (src >= n_rows_in ? n_rows_in - 1 : src);   // traced gathers clamp
    out[k] = in[(b * n_rows_in + src) * row + c];
  }
}

// FP64 pipe probe: 8 independent DFMA chains per thread, enough warps to saturate every SM sub-partition.
__global__ void __launch_bounds__(256) fp64_probe_kernel(double* out, int iters, double a, double b) {
  double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
      x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
  }
  double s = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
  if (s == 123.456) out[0] = s;   // never true; keeps the chains alive
}

// Gaussian observation log-likelihood on the solver output, fused with the gather Xt[obs_ind]:
//   out[b] = sum_{i, k} log N(obs_data[i, k]; Xt[b, obs_ind[i], k, 0], noise_sd^2)
// (the measurement model of the reference's parameter-inference walkthrough, docs/examples/parameter.md:192-205,
// evaluated inside the pseudo-marginal MCMC step :333-354).  One thread per theta.
template <typename T>
__global__ void gauss_obs_loglik_kernel(i64 B, int n_rows, int n_obs, int nb, int p, const int* __restrict__ ind,
                                        const T* __restrict__ Xt, const T* __restrict__ obs_data, T noise_sd,
                                        T* __restrict__ out) {
  const i64 b = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const T rvar = T(1) / (noise_sd * noise_sd);
  const T cst = T(-0.5) * T(1.8378770664093454836) - log(noise_sd);
  T acc = T(0);
  for (int i = 0; i < n_obs; ++i) {
    int r = ind[i];
    r = r < 0 ? 0 : (r >= n_rows ? n_rows - 1 : r);
    const T* row = Xt + (b * n_rows + r) * (i64)(nb * p);
    for (int k = 0; k < nb; ++k) {
      const T d = obs_data[i * nb + k] - row[k * p];
      acc += cst - T(0.5) * d * d * rvar;
    }
  }
  out[b] = acc;
}

template <class Model, int INTERR, int QK>
struct InitPadRun {
  template <typename T>
  static int run(const RodeoProblem& p, T t, const T* theta, const T* x0, T* X0, cudaStream_t s) {
    if (p.B == 0) return RODEO_OK;
    ode_init_pad_kernel<T, Model><<<grid_for(p.B, 128), 128, 0, s>>>(p.B, t, theta, x0, X0);
    g_launches++;
    RODEO_CUDA_OK(cudaGetLastError());
    return RODEO_OK;
  }
};

}  // namespace host
}  // namespace rodeo

using namespace rodeo;
using namespace rodeo::host;

template <typename T>
static int basic_gather_impl(const RodeoProblem* p, const T* Xt, const int32_t* obs_ind, T* ode_data, void* stream) {
  if (int rc = check_common(p)) return rc;
  if (p->n_obs < 0) { set_error("n_obs < 0"); return RODEO_ERR_INVALID; }
  const long long total = p->B * (long long)p->n_obs * p->n_block * p->n_bstate;
  if (total == 0) return RODEO_OK;
  unsigned grid = (unsigned)((total + 255) / 256);
  if (grid > 148u * 16u) grid = 148u * 16u;
  gather_rows_kernel<T><<<grid, 256, 0, (cudaStream_t)stream>>>(p->B, p->n_steps + 1, p->n_obs, p->n_block * p->n_bstate,
                                                               obs_ind, Xt, ode_data);
  g_launches++;
  RODEO_CUDA_OK(cudaGetLastError());
  return RODEO_OK;
}

template <typename T>
static int ode_init_pad_impl(const RodeoProblem* p, T t, const T* theta, const T* x0, T* X0, void* stream) {
  if (int rc = check_common(p)) return rc;
  if (p->n_bstate < 2) { set_error("first_order_pad needs n_deriv >= 2"); return RODEO_ERR_INVALID; }
  if (p->model_id >= RODEO_MODEL_USER_BASE) {
    if (sizeof(T) != 8) { set_error("user (NVRTC) models are float64 only"); return RODEO_ERR_UNSUPPORTED; }
    long long B = p->B;
    return user_launch_raw(p->model_id, "rodeo::ode_init_pad_kernel<double, UserModel>", B, 128,
                           {&B, &t, &theta, &x0, &X0}, (cudaStream_t)stream);
  }
  RodeoProblem q = *p;
  q.interrogate = RODEO_INTERROGATE_KRAMER;
  return dispatch_model<InitPadRun>(q, (const T*)nullptr, (const T*)nullptr, q, t, theta, x0, X0, (cudaStream_t)stream);
}


extern "C" {

const char* rodeo_b200_last_error(void) { return g_err; }
int rodeo_b200_abi_version(void) { return RODEO_B200_ABI_VERSION; }
size_t rodeo_b200_problem_sizeof(void) { return sizeof(RodeoProblem); }
int64_t rodeo_b200_launch_count(void) { return (int64_t)g_launches.load(); }
int64_t rodeo_b200_schedule_builds(void) { return (int64_t)sched_builds(); }
void rodeo_b200_schedule_clear(void) { sched_clear(); }

size_t rodeo_b200_workspace_bytes(int op, const RodeoProblem* p, int elem_bytes) {
  if (!p || (elem_bytes != 8 && elem_bytes != 4)) return 0;
  switch (op) {
    case RODEO_OP_SOLVE_MV:
    case RODEO_OP_SOLVE_SIM:
    case RODEO_OP_FENRIR: {
      // history of filtered states, theta-innermost: entries filt[KC], filt[2 KC], ... < N.  solve_mv keeps one
      // checkpoint per shared-memory segment (KC = seg_len(nstate)) and recomputes; solve_sim / fenrir keep every
      // state (KC = 1).  See rodeo_kernels.cuh.
      // (solve_sim over a covariance schedule keeps the block means only: n_block * n_bstate values per entry)
      const bool means_only = op == RODEO_OP_SOLVE_SIM && sim_schedule_selected(*p);
      const int nstate = means_only ? p->n_block * p->n_bstate : nstate_of(p->n_block, p->n_bstate);
      // solve_mv of a built-in model with n_block >= 2 runs the (theta, block)-lane kernel, whose segments are longer
      // (a per-theta prior runs the one-lane-per-theta kernel)
      const bool bl = p->n_block >= 2 && p->model_id < RODEO_MODEL_USER_BASE && !p->prior_batched;
      // (solve_sim's one-lane-per-theta kernel keeps only one checkpoint per segment as well, but the (theta, block)-lane
      // kernel the host may pick instead keeps every state: sized for the latter)
      const int KC = (op == RODEO_OP_SOLVE_MV) ? (bl ? seg_len_bl(nstate, p->n_block, elem_bytes) : seg_len(nstate)) : 1;
      const size_t J = ((size_t)p->n_steps + KC - 1) / KC;
      return (J > 0 ? J - 1 : 0) * (size_t)nstate * (size_t)stash_ldb(p->B) * (size_t)elem_bytes;
    }
    default:
      return 0;
  }
}

// Measured FP64 FMA throughput of the current device in TFLOP/s (2 flop per DFMA), best of `reps` timed launches.
// bench.py uses it as the roofline denominator for the FP64-bound kernels (MEASURED_PEAKS.json has no FP64 figure).
namespace {
struct ProbeResources {           // freed on every exit path
  double* d = nullptr;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  cudaStream_t s = nullptr;
  ~ProbeResources() {
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    if (s) cudaStreamDestroy(s);
    if (d) cudaFree(d);
  }
};
}  // namespace

int rodeo_b200_fp64_peak_probe(int reps, double* tflops_out) {
  if (!tflops_out) { set_error("tflops_out is NULL"); return RODEO_ERR_INVALID; }
  int dev = 0, sms = 0;
  RODEO_CUDA_OK(cudaGetDevice(&dev));
  RODEO_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  ProbeResources r;
  RODEO_CUDA_OK(cudaMalloc((void**)&r.d, 8));
  RODEO_CUDA_OK(cudaEventCreate(&r.e0));
  RODEO_CUDA_OK(cudaEventCreate(&r.e1));
  RODEO_CUDA_OK(cudaStreamCreateWithFlags(&r.s, cudaStreamNonBlocking));   // private stream: no implicit joins
  const int iters = 4096, grid = sms * 8, block = 256;
  double best = 0.0;
  for (int k = 0; k < reps + 2; ++k) {
    RODEO_CUDA_OK(cudaEventRecord(r.e0, r.s));
    fp64_probe_kernel<<<grid, block, 0, r.s>>>(r.d, iters, 0.999999, 1e-9);
    RODEO_CUDA_OK(cudaGetLastError());
    RODEO_CUDA_OK(cudaEventRecord(r.e1, r.s));
    RODEO_CUDA_OK(cudaEventSynchronize(r.e1));
    float ms = 0.f;
    RODEO_CUDA_OK(cudaEventElapsedTime(&ms, r.e0, r.e1));
    const double flops = 2.0 * 64.0 * (double)iters * (double)grid * (double)block;
    const double tf = flops / (ms * 1e-3) / 1e12;
    if (k >= 2 && tf > best) best = tf;
  }
  *tflops_out = best;
  return RODEO_OK;
}

int rodeo_b200_gauss_obs_loglik_f64(const RodeoProblem* p, const double* Xt, const int32_t* obs_ind,
                                    const double* obs_data, double noise_sd, double* loglik_out, void* stream) {
  if (int rc = check_common(p)) return rc;
  if (p->n_obs < 1 || !(noise_sd > 0.0)) { set_error("gauss_obs_loglik needs n_obs >= 1 and noise_sd > 0"); return RODEO_ERR_INVALID; }
  if (p->B == 0) return RODEO_OK;
  gauss_obs_loglik_kernel<double><<<grid_for(p->B, 128), 128, 0, (cudaStream_t)stream>>>(
      p->B, p->n_steps + 1, p->n_obs, p->n_block, p->n_bstate, obs_ind, Xt, obs_data, noise_sd, loglik_out);
  g_launches++;
  RODEO_CUDA_OK(cudaGetLastError());
  return RODEO_OK;
}
int rodeo_b200_basic_gather_f64(const RodeoProblem* p, const double* Xt, const int32_t* obs_ind, double* ode_data,
                                void* stream) { return basic_gather_impl<double>(p, Xt, obs_ind, ode_data, stream); }
int rodeo_b200_basic_gather_f32(const RodeoProblem* p, const float* Xt, const int32_t* obs_ind, float* ode_data,
                                void* stream) { return basic_gather_impl<float>(p, Xt, obs_ind, ode_data, stream); }
int rodeo_b200_ode_init_pad_f64(const RodeoProblem* p, double t, const double* theta, const double* x0, double* X0,
                                void* stream) { return ode_init_pad_impl<double>(p, t, theta, x0, X0, stream); }
int rodeo_b200_ode_init_pad_f32(const RodeoProblem* p, float t, const float* theta, const float* x0, float* X0,
                                void* stream) { return ode_init_pad_impl<float>(p, t, theta, x0, X0, stream); }

}  // extern "C"
