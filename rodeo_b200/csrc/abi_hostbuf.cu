// Host-buffer convenience entry points (include/rodeo_b200.h: rodeo_b200_*_host): every pointer is HOST memory.
// They stage inputs into a cached device arena, call the device-pointer entry point, copy the result back and
// synchronise.  The arena only ever grows, so steady-state calls perform no allocation.
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "rodeo_host.h"

using namespace rodeo;
using namespace rodeo::host;

namespace {

struct Arena {
  std::mutex mu;
  char* base = nullptr;
  size_t cap = 0, used = 0;
  cudaStream_t stream = nullptr;
  bool ready_ok = false;      // every resource below exists (set last, so a failed initialisation is retried)
  int device = -1;            // the device the arena, its streams and events belong to
  // chunked pipeline of the log-likelihood entry points: chunk c's inputs travel on `stream`, its kernel runs on
  // lane[c] as soon as they have landed (ready[c]) and `stream` collects the results after done[c]
  static constexpr int NCHUNK = 8;
  // pinned staging for the small per-call arrays (observation tables): one H2D copy instead of one per array
  static constexpr size_t STAGE_BYTES = 256 << 10;
  char* stage = nullptr;
  cudaStream_t lane[NCHUNK] = {};
  cudaEvent_t ready[NCHUNK] = {}, done[NCHUNK] = {};

  int reserve(size_t bytes) {
    used = 0;
    int dev = 0;
    RODEO_CUDA_OK(cudaGetDevice(&dev));
    if (ready_ok && dev != device) release();      // the caller switched devices: rebuild the arena there
    if (!ready_ok) {
      release();                                   // drop the leftovers of a failed initialisation
      device = dev;
      RODEO_CUDA_OK(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
      RODEO_CUDA_OK(cudaHostAlloc((void**)&stage, STAGE_BYTES, cudaHostAllocDefault));
      for (int c = 0; c < NCHUNK; ++c) {
        RODEO_CUDA_OK(cudaStreamCreateWithFlags(&lane[c], cudaStreamNonBlocking));
        RODEO_CUDA_OK(cudaEventCreateWithFlags(&ready[c], cudaEventDisableTiming));
        RODEO_CUDA_OK(cudaEventCreateWithFlags(&done[c], cudaEventDisableTiming));
      }
      ready_ok = true;
    }
    if (bytes <= cap) return RODEO_OK;
    if (base) { RODEO_CUDA_OK(cudaFree(base)); base = nullptr; cap = 0; }
    RODEO_CUDA_OK(cudaMalloc((void**)&base, bytes));
    cap = bytes;
    return RODEO_OK;
  }
  void* take(size_t bytes) {
    void* p = base + used;
    used += round_up(bytes, 256);
    return p;
  }
  // wait for everything this arena has in flight (error paths: the next call reuses the arena and the pinned staging
  // buffer under the assumption that the previous call has completed)
  void drain() {
    if (stream) cudaStreamSynchronize(stream);
    for (int c = 0; c < NCHUNK; ++c)
      if (lane[c]) cudaStreamSynchronize(lane[c]);
  }
  void release() {
    drain();
    if (base) cudaFree(base);
    if (stage) cudaFreeHost(stage);
    if (stream) cudaStreamDestroy(stream);
    for (int c = 0; c < NCHUNK; ++c) {
      if (lane[c]) cudaStreamDestroy(lane[c]);
      if (ready[c]) cudaEventDestroy(ready[c]);
      if (done[c]) cudaEventDestroy(done[c]);
      lane[c] = nullptr; ready[c] = nullptr; done[c] = nullptr;
    }
    base = nullptr; stage = nullptr; cap = used = 0; stream = nullptr; ready_ok = false; device = -1;
  }
};
Arena g_arena;

inline size_t al(size_t b) { return round_up(b, 256); }

}  // namespace

extern "C" void rodeo_b200_host_arena_release(void) {
  std::lock_guard<std::mutex> lk(g_arena.mu);
  g_arena.release();
}

// The bodies return on the first error; the exported wrappers then drain the arena's streams so that nothing is left
// in flight on arena memory or on the pinned staging buffer when the next call starts.
static int dalton_host_body(const RodeoProblem* p, const double* ode_weight, const double* prior_weight,
                            const double* prior_var, const double* ode_init, const double* theta,
                            const int32_t* obs_ind, const double* obs_data, const double* obs_weight,
                            const double* obs_var, double* loglik_out) {
  const size_t B = (size_t)p->B, nb = p->n_block, ps = p->n_bstate, no = p->n_obs, nob = p->n_bobs;
  const size_t b_init = B * nb * ps * 8, b_theta = B * p->n_theta * 8, b_ll = B * 8;
  const size_t b_ind = no * 4, b_y = no * nb * nob * 8, b_D = no * nb * nob * ps * 8, b_Om = no * nb * nob * nob * 8;
  // In the *_host entry points RodeoProblem::prior_var_scale is a HOST array (B, n_block) like every other pointer
  const double* h_scale = (const double*)p->prior_var_scale;
  const size_t b_scale = h_scale ? B * nb * 8 : 0;
  if (int rc = g_arena.reserve(al(b_init) + al(b_theta) + al(b_ll) + al(b_scale) + al(b_ind) + al(b_y) + al(b_D) +
                               al(b_Om) + 256))
    return rc;
  double* d_init = (double*)g_arena.take(b_init);
  double* d_theta = (double*)g_arena.take(b_theta);
  double* d_ll = (double*)g_arena.take(b_ll);
  double* d_scale = h_scale ? (double*)g_arena.take(b_scale) : nullptr;
  int32_t* d_ind = (int32_t*)g_arena.take(b_ind);
  double* d_y = (double*)g_arena.take(b_y);
  double* d_D = (double*)g_arena.take(b_D);
  double* d_Om = (double*)g_arena.take(b_Om);
  cudaStream_t s = g_arena.stream;
  // the four observation tables sit back to back in the arena (d_ind .. d_Om): stage them with the same layout in
  // pinned memory and send them as one copy (the previous call has synchronised, so the staging buffer is free)
  const size_t obs_span = (size_t)((char*)d_Om - (char*)d_ind) + b_Om;
  if (obs_span <= Arena::STAGE_BYTES) {
    char* st = g_arena.stage;
    memcpy(st, obs_ind, b_ind);
    memcpy(st + ((char*)d_y - (char*)d_ind), obs_data, b_y);
    memcpy(st + ((char*)d_D - (char*)d_ind), obs_weight, b_D);
    memcpy(st + ((char*)d_Om - (char*)d_ind), obs_var, b_Om);
    RODEO_CUDA_OK(cudaMemcpyAsync(d_ind, st, obs_span, cudaMemcpyHostToDevice, s));
  } else {
    RODEO_CUDA_OK(cudaMemcpyAsync(d_ind, obs_ind, b_ind, cudaMemcpyHostToDevice, s));
    RODEO_CUDA_OK(cudaMemcpyAsync(d_y, obs_data, b_y, cudaMemcpyHostToDevice, s));
    RODEO_CUDA_OK(cudaMemcpyAsync(d_D, obs_weight, b_D, cudaMemcpyHostToDevice, s));
    RODEO_CUDA_OK(cudaMemcpyAsync(d_Om, obs_var, b_Om, cudaMemcpyHostToDevice, s));
  }
  // The theta batch is independent per theta, so it is cut into two contiguous chunks (up to NCHUNK with
  // RODEO_HOST_CHUNKS): chunk c's kernel starts as soon as its own X0 / theta rows have landed, while the next chunk's
  // rows are still crossing PCIe; the kernels run on separate streams so they share the GPU instead of queueing behind
  // each other's tails.  Results are identical to one launch (per-theta arithmetic; random streams are keyed by the
  // global particle index).  Measured on B200 (65,536 thetas, FN dalton): 1 / 2 / 4 / 8 chunks -> 57.7 / 59.5 / 57.6 /
  // 56.3 G theta*steps/s end to end: the kernel is FP64-bound, so every extra chunk costs a little tail; two hide half
  // of the PCIe transfer
  int nchunk = B >= 32768 ? 2 : 1;
  if (const char* e = getenv("RODEO_HOST_CHUNKS")) {                       // tuning experiments
    const int v = atoi(e);
    if (v >= 1 && v <= Arena::NCHUNK && (size_t)v <= B) nchunk = v;
  }
  // chunk boundaries: equal chunks, or (RODEO_HOST_SPLIT="f0,f1,..": tuning experiments) the given fractions of the batch.
  // Measured (tools/e2e_split.py, B200, 65,536 thetas, pinned buffers): 1 chunk 0.861 ms, 2 equal chunks 0.826,
  // 1/8 + 1/4 + 5/8 0.816, 1/4 + 3/4 0.820, 1/16 + 3/16 + 3/4 0.867, 1/8 + 7/8 0.861 -- nothing beats two equal chunks by
  // more than 1.3 %: 0.73 ms of kernel, the first chunk's rows, the result's way back and two event hand-overs are the floor
  size_t bound[Arena::NCHUNK + 1];
  {
    const size_t per = (B + nchunk - 1) / nchunk;
    for (int c = 0; c <= nchunk; ++c) bound[c] = (size_t)c * per < B ? (size_t)c * per : B;
    if (const char* e = getenv("RODEO_HOST_SPLIT")) {
      double f[Arena::NCHUNK]; int nf = 0; const char* q = e;
      while (nf < Arena::NCHUNK && *q) { char* end; const double v = strtod(q, &end); if (end == q) break; f[nf++] = v; q = *end ? end + 1 : end; }
      if (nf >= 1) {
        nchunk = nf; double acc = 0; bound[0] = 0;
        for (int c = 0; c < nf; ++c) { acc += f[c]; size_t x = (size_t)(acc * (double)B) / 32 * 32; bound[c + 1] = x < B ? x : B; }
        bound[nf] = B;
      }
    }
  }
  const size_t row_init = nb * ps, row_theta = (size_t)p->n_theta;
  bool launched[Arena::NCHUNK] = {};
  for (int c = 0; c < nchunk; ++c) {
    const size_t b0 = bound[c];
    if (b0 >= B || bound[c + 1] <= b0) continue;
    const size_t bn = bound[c + 1] - b0;
    RODEO_CUDA_OK(cudaMemcpyAsync(d_init + b0 * row_init, ode_init + b0 * row_init, bn * row_init * 8,
                                  cudaMemcpyHostToDevice, s));
    RODEO_CUDA_OK(cudaMemcpyAsync(d_theta + b0 * row_theta, theta + b0 * row_theta, bn * row_theta * 8,
                                  cudaMemcpyHostToDevice, s));
    if (h_scale)
      RODEO_CUDA_OK(cudaMemcpyAsync(d_scale + b0 * nb, h_scale + b0 * nb, bn * nb * 8, cudaMemcpyHostToDevice, s));
    RODEO_CUDA_OK(cudaEventRecord(g_arena.ready[c], s));
    RODEO_CUDA_OK(cudaStreamWaitEvent(g_arena.lane[c], g_arena.ready[c], 0));
    RodeoProblem pc = *p;
    pc.B = (int64_t)bn;
    pc.particle_offset = p->particle_offset + (int64_t)b0;
    pc.prior_var_scale = h_scale ? (const void*)(d_scale + b0 * nb) : nullptr;     // this chunk's rows
    if (int rc = rodeo_b200_dalton_f64(&pc, ode_weight, prior_weight, prior_var, d_init + b0 * row_init,
                                       d_theta + b0 * row_theta, nullptr, d_ind, d_y, d_D, d_Om, d_ll + b0, nullptr, 0,
                                       g_arena.lane[c]))
      return rc;
    RODEO_CUDA_OK(cudaEventRecord(g_arena.done[c], g_arena.lane[c]));
    launched[c] = true;
  }
  // only now does the copy stream wait for the kernels: waiting inside the loop would hold back the next chunk's copies
  for (int c = 0; c < nchunk; ++c)
    if (launched[c]) RODEO_CUDA_OK(cudaStreamWaitEvent(s, g_arena.done[c], 0));
  RODEO_CUDA_OK(cudaMemcpyAsync(loglik_out, d_ll, b_ll, cudaMemcpyDeviceToHost, s));
  RODEO_CUDA_OK(cudaStreamSynchronize(s));
  return RODEO_OK;
}

extern "C" int rodeo_b200_dalton_f64_host(const RodeoProblem* p, const double* ode_weight, const double* prior_weight,
                                          const double* prior_var, const double* ode_init, const double* theta,
                                          const int32_t* obs_ind, const double* obs_data, const double* obs_weight,
                                          const double* obs_var, double* loglik_out) {
  if (int rc = check_common(p)) return rc;
  if (p->prior_batched) { set_error("the *_host wrappers take a shared prior (prior_batched must be 0)"); return RODEO_ERR_UNSUPPORTED; }
  if (p->n_obs < 1) { set_error("dalton needs n_obs >= 1"); return RODEO_ERR_INVALID; }
  std::lock_guard<std::mutex> lk(g_arena.mu);
  const int rc = dalton_host_body(p, ode_weight, prior_weight, prior_var, ode_init, theta, obs_ind, obs_data,
                                  obs_weight, obs_var, loglik_out);
  if (rc != RODEO_OK) g_arena.drain();
  return rc;
}

static int solve_mv_host_body(const RodeoProblem* p, const double* ode_weight, const double* prior_weight,
                              const double* prior_var, const double* ode_init, const double* theta,
                              double* mean_out, double* var_out) {
  const size_t B = (size_t)p->B, nb = p->n_block, ps = p->n_bstate, N1 = (size_t)p->n_steps + 1;
  const size_t b_init = B * nb * ps * 8, b_theta = B * p->n_theta * 8;
  const size_t b_mean = B * N1 * nb * ps * 8, b_var = b_mean * ps;
  const size_t b_ws = rodeo_b200_workspace_bytes(RODEO_OP_SOLVE_MV, p, 8);
  const double* h_scale = (const double*)p->prior_var_scale;      // HOST (B, n_block) in the *_host entry points
  const size_t b_scale = h_scale ? B * nb * 8 : 0;
  if (int rc = g_arena.reserve(al(b_init) + al(b_theta) + al(b_scale) + al(b_mean) + al(b_var) + al(b_ws) + 256))
    return rc;
  double* d_init = (double*)g_arena.take(b_init);
  double* d_theta = (double*)g_arena.take(b_theta);
  double* d_scale = h_scale ? (double*)g_arena.take(b_scale) : nullptr;
  double* d_mean = (double*)g_arena.take(b_mean);
  double* d_var = (double*)g_arena.take(b_var);
  void* d_ws = g_arena.take(b_ws);
  cudaStream_t s = g_arena.stream;
  RODEO_CUDA_OK(cudaMemcpyAsync(d_init, ode_init, b_init, cudaMemcpyHostToDevice, s));
  RODEO_CUDA_OK(cudaMemcpyAsync(d_theta, theta, b_theta, cudaMemcpyHostToDevice, s));
  if (h_scale) RODEO_CUDA_OK(cudaMemcpyAsync(d_scale, h_scale, b_scale, cudaMemcpyHostToDevice, s));
  RodeoProblem pc = *p;
  pc.prior_var_scale = d_scale;
  if (int rc = rodeo_b200_solve_mv_f64(&pc, ode_weight, prior_weight, prior_var, d_init, d_theta, nullptr, d_mean,
                                       d_var, d_ws, b_ws, s))
    return rc;
  RODEO_CUDA_OK(cudaMemcpyAsync(mean_out, d_mean, b_mean, cudaMemcpyDeviceToHost, s));
  RODEO_CUDA_OK(cudaMemcpyAsync(var_out, d_var, b_var, cudaMemcpyDeviceToHost, s));
  RODEO_CUDA_OK(cudaStreamSynchronize(s));
  return RODEO_OK;
}

extern "C" int rodeo_b200_solve_mv_f64_host(const RodeoProblem* p, const double* ode_weight, const double* prior_weight,
                                            const double* prior_var, const double* ode_init, const double* theta,
                                            double* mean_out, double* var_out) {
  if (int rc = check_common(p)) return rc;
  if (p->prior_batched) { set_error("the *_host wrappers take a shared prior (prior_batched must be 0)"); return RODEO_ERR_UNSUPPORTED; }
  std::lock_guard<std::mutex> lk(g_arena.mu);
  const int rc = solve_mv_host_body(p, ode_weight, prior_weight, prior_var, ode_init, theta, mean_out, var_out);
  if (rc != RODEO_OK) g_arena.drain();
  return rc;
}
