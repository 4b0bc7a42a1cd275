// Peer-memory all-gather of the per-theta log-likelihoods: ONE kernel per rank stores its shard straight into every
// peer's gathered vector over NVLink / NVSwitch (CUDA IPC mappings, one process per GPU), publishes a flag per peer and
// waits for the peers' flags.  It replaces the NCCL all_gather_into_tensor that follows a log-likelihood kernel when the
// step's latency matters (a 65,536-theta batch over 8 GPUs is a 0.145 ms kernel per rank: a 20-40 us collective is a
// fifth of the step); rodeo_b200/parallel.py: PeerGather, with NCCL as the fallback.
//
// Each rank owns a region [2 slots][n_total] doubles + [2 slots][world] flags.  Call k (epoch k, slot k & 1): every
// thread stores the local shard into all regions' slot; after a CTA barrier, thread r < world issues a system-scope
// fence and a release store of the epoch into region r's flag [slot][rank]; then it acquires its OWN region's flag
// [slot][r] until it reads the epoch -- bounded (spin_limit polls; on expiry *status becomes 1 and the kernel returns, so a
// missing peer can never hang the GPU).  Two slots suffice: a rank can run ahead of a peer by at most one call, because
// call k + 1 cannot complete before every peer has stored call k + 1, which a peer issues only after the consumers of
// call k on its own stream.
#include <cstring>

#include "rodeo_host.h"

namespace rodeo {
namespace {

constexpr int PEER_MAX = 16;
struct PeerPtrs { double* data[PEER_MAX]; unsigned* flags[PEER_MAX]; };

__global__ void __launch_bounds__(1024)
peer_allgather_kernel(const double* __restrict__ local, long long n_local, long long offset, long long n_total,
                      int rank, int world, PeerPtrs pp, unsigned epoch, unsigned* __restrict__ epoch_counter,
                      unsigned long long spin_limit, int* __restrict__ status) {
  if (epoch == 0) {
    // the epoch lives on the device (a launch captured in a CUDA graph cannot take a new argument per replay): every
    // rank's counter advances by one per call, in lockstep
    __shared__ unsigned e_sh;
    if (threadIdx.x == 0) { e_sh = *epoch_counter + 1u; *epoch_counter = e_sh; }
    __syncthreads();
    epoch = e_sh;
  }
  const int slot = (int)(epoch & 1u);
  for (int r = 0; r < world; ++r) {
    double* dst = pp.data[r] + (long long)slot * n_total + offset;
    for (long long i = threadIdx.x; i < n_local; i += blockDim.x) dst[i] = local[i];
  }
  __syncthreads();
  if ((int)threadIdx.x < world) {
    const int r = threadIdx.x;
    __threadfence_system();
    unsigned* theirs = pp.flags[r] + slot * world + rank;
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(theirs), "r"(epoch) : "memory");
    const unsigned* mine = pp.flags[rank] + slot * world + r;
    unsigned v = 0;
    unsigned long long polls = 0;
    for (;;) {
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
      if (v == epoch) break;
      if (++polls >= spin_limit) { atomicExch(status, 1); break; }
      __nanosleep(64);
    }
  }
  __syncthreads();
}

}  // namespace
}  // namespace rodeo

using namespace rodeo;
using namespace rodeo::host;

// layout of a region of n_total elements: data at 0, flags after the two slots (256-byte aligned)
static size_t flags_offset(long long n_total) { return round_up((size_t)2 * (size_t)n_total * sizeof(double), 256); }

extern "C" size_t rodeo_b200_peer_region_bytes(long long n_total, int world) {
  if (n_total < 0 || world < 1 || world > PEER_MAX) return 0;
  return flags_offset(n_total) + round_up((size_t)2 * world * sizeof(unsigned), 256);
}

extern "C" int rodeo_b200_peer_alloc(size_t bytes, void** ptr, unsigned char* handle64) {
  if (!ptr || !handle64 || bytes == 0) { set_error("peer_alloc: bad arguments"); return RODEO_ERR_INVALID; }
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  void* p = nullptr;
  RODEO_CUDA_OK(cudaMalloc(&p, bytes));
  cudaError_t e = cudaMemset(p, 0, bytes);
  cudaIpcMemHandle_t h;
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { cudaFree(p); return cuda_fail(e, "peer_alloc"); }
  memcpy(handle64, &h, 64);
  *ptr = p;
  return RODEO_OK;
}

extern "C" int rodeo_b200_peer_open(const unsigned char* handle64, void** ptr) {
  if (!ptr || !handle64) { set_error("peer_open: bad arguments"); return RODEO_ERR_INVALID; }
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  RODEO_CUDA_OK(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return RODEO_OK;
}

extern "C" int rodeo_b200_peer_close(void* ptr) {
  if (ptr) RODEO_CUDA_OK(cudaIpcCloseMemHandle(ptr));
  return RODEO_OK;
}

extern "C" int rodeo_b200_peer_free(void* ptr) {
  if (ptr) RODEO_CUDA_OK(cudaFree(ptr));
  return RODEO_OK;
}

extern "C" int rodeo_b200_peer_allgather_f64(const double* local, long long n_local, long long offset,
                                             long long n_total, int rank, int world, void* const* regions,
                                             unsigned epoch, unsigned* epoch_counter, unsigned long long spin_limit,
                                             int* status, void* stream) {
  if ((!local && n_local > 0) || !regions || !status || world < 1 || world > PEER_MAX || rank < 0 || rank >= world || n_local < 0 ||
      offset < 0 || offset + n_local > n_total || (epoch == 0 && !epoch_counter)) {
    set_error("peer_allgather: bad arguments (world <= %d, epoch >= 1 or a device counter, offset + n_local <= n_total)",
              PEER_MAX);
    return RODEO_ERR_INVALID;
  }
  PeerPtrs pp;
  memset(&pp, 0, sizeof(pp));
  for (int r = 0; r < world; ++r) {
    if (!regions[r]) { set_error("peer_allgather: region %d is NULL", r); return RODEO_ERR_INVALID; }
    pp.data[r] = (double*)regions[r];
    pp.flags[r] = (unsigned*)((char*)regions[r] + flags_offset(n_total));
  }
  peer_allgather_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(local, n_local, offset, n_total, rank, world, pp, epoch,
                                                              epoch_counter, spin_limit ? spin_limit : (1ull << 22),
                                                              status);
  g_launches++;
  RODEO_CUDA_OK(cudaGetLastError());
  return RODEO_OK;
}
