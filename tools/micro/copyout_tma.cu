// Micro-benchmark (not product code): the solve_mv output pattern written with TMA bulk stores, 16 thetas per warp
// (the block-lane kernel's geometry), K staged rows per theta; compared with lane-per-element runs of the same K.
#include <cstdio>
#include <cuda_runtime.h>
typedef long long i64;
extern __shared__ __align__(16) double dyn[];
template <int K, int TW>
__global__ void __launch_bounds__(32) kb(double* __restrict__ mean, double* __restrict__ var, i64 B, int N) {
  constexpr int PITCH = K * 24 + 2;
  double* buf = dyn;
  const int lane = threadIdx.x;
  const bool act = lane < TW;
  const i64 theta = (i64)blockIdx.x * TW + (act ? lane : 0);
  if (act) for (int i = 0; i < K * 24; ++i) buf[lane * PITCH + i] = i;
  unsigned sm = (unsigned)__cvta_generic_to_shared(buf + (act ? lane : 0) * PITCH);
  unsigned sv = sm + K * 6 * 8;
  for (int j = (N - 1) / K; j >= 0; --j) {
    const int n0 = j * K, cnt = (N - n0) < K ? (N - n0) : K;
    if (act) {
      buf[lane * PITCH + (j % (K * 24))] = j;
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      double* gm = mean + (theta * (N + 1) + n0) * 6;
      double* gv = var + (theta * (N + 1) + n0) * 18;
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gm), "r"(sm), "r"(cnt * 48) : "memory");
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gv), "r"(sv), "r"(cnt * 144) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
    __syncwarp();
  }
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
template <int K, int TW>
__global__ void __launch_bounds__(32) kl(double* __restrict__ mean, double* __restrict__ var, i64 B, int N) {
  double* buf = dyn;                       // [K*18][TW+1]
  const int lane = threadIdx.x;
  const i64 theta0 = (i64)blockIdx.x * TW;
  for (int i = lane; i < K * 18 * (TW + 1); i += 32) buf[i] = i;
  __syncwarp();
  for (int j = (N - 1) / K; j >= 0; --j) {
    const int n0 = j * K, cnt = (N - n0) < K ? (N - n0) : K;
    {
      const int run = cnt * 6; double* dst = mean + (theta0 * (N + 1) + n0) * 6 + lane;
      for (int th = 0; th < TW; ++th) { for (int r = lane; r < run; r += 32) dst[r - lane] = buf[(r % (K * 18)) * (TW + 1) + th]; dst += (i64)(N + 1) * 6; }
    }
    {
      const int run = cnt * 18; double* dst = var + (theta0 * (N + 1) + n0) * 18 + lane;
#pragma unroll 4
      for (int th = 0; th < TW; ++th) { for (int r = lane; r < run; r += 32) dst[r - lane] = buf[(r % (K * 18)) * (TW + 1) + th]; dst += (i64)(N + 1) * 18; }
    }
    __syncwarp();
  }
}
template <int K, int TW, bool TMA> void run(double* mean, double* var, i64 B, int N) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int smem = TMA ? TW * (K * 24 + 2) * 8 : K * 18 * (TW + 1) * 8;
  if (TMA) cudaFuncSetAttribute(kb<K, TW>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  else cudaFuncSetAttribute(kl<K, TW>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  float best = 1e9;
  for (int r = 0; r < 4; ++r) {
    cudaEventRecord(e0);
    if (TMA) kb<K, TW><<<(unsigned)(B / TW), 32, smem>>>(mean, var, B, N); else kl<K, TW><<<(unsigned)(B / TW), 32, smem>>>(mean, var, B, N);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
  }
  double gb = (double)B * (N + 1) * 24 * 8 / 1e9;
  printf("%-18s TW=%d K=%2d smem %5d B  %.3f ms  %.2f TB/s  (%s)\n", TMA ? "TMA bulk per theta" : "lane-per-element", TW, K, smem, best, gb / best, cudaGetErrorString(cudaGetLastError()));
}
int main() {
  const i64 B = 65536; const int N = 800;
  double *mean, *var;
  cudaMalloc(&mean, B * (N + 1) * 6 * 8); cudaMalloc(&var, B * (N + 1) * 18 * 8);
  run<7, 16, false>(mean, var, B, N); run<9, 16, false>(mean, var, B, N); run<12, 16, false>(mean, var, B, N);
  run<5, 16, true>(mean, var, B, N); run<7, 16, true>(mean, var, B, N); run<9, 16, true>(mean, var, B, N); run<12, 16, true>(mean, var, B, N);
  run<7, 32, true>(mean, var, B, N);
  run<9, 8, true>(mean, var, B, N); run<18, 8, true>(mean, var, B, N);
  return 0;
}
