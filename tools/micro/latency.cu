// Dependent-issue latencies of the instructions the filter recursions chain together, measured with clock64() on one
// warp (B200, sm_100a):  nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o latency latency.cu && ./latency
#include <cstdio>
#include <cuda_runtime.h>

template <int CHAINS>
__global__ void dfma_chain(double* out, long long* cyc, int iters, double a, double b) {
  double x[CHAINS];
  for (int c = 0; c < CHAINS; ++c) x[c] = threadIdx.x + c;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 16; ++u)
#pragma unroll
      for (int c = 0; c < CHAINS; ++c) x[c] = fma(x[c], a, b);
  }
  long long t1 = clock64();
  double s = 0;
  for (int c = 0; c < CHAINS; ++c) s += x[c];
  out[threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

__global__ void rcp_chain(double* out, long long* cyc, int iters, int full) {
  double x = 1.0 + threadIdx.x * 1e-3;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      double y;
      asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
      if (full) { const double e = fma(-x, y, 1.0); y = fma(y, fma(e, e, e), y); }
      x = y + 0.5;
    }
  }
  long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

__global__ void div_chain(double* out, long long* cyc, int iters) {
  double x = 1.0 + threadIdx.x * 1e-3;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 16; ++u) x = 1.0 / x + 0.5;
  }
  long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

__global__ void shfl_chain(double* out, long long* cyc, int iters) {
  double x = threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 16; ++u) x = __shfl_xor_sync(0xffffffffu, x, 16) + 1.0;
  }
  long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

__global__ void lds_chain(double* out, long long* cyc, int iters) {
  __shared__ double sm[64];
  sm[threadIdx.x] = threadIdx.x; sm[threadIdx.x + 32] = 1.0;
  __syncwarp();
  double x = threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      sm[threadIdx.x] = x;
      __syncwarp();
      x = sm[threadIdx.x ^ 16] + 1.0;
      __syncwarp();
    }
  }
  long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

__global__ void ffma_chain(float* out, long long* cyc, int iters, float a, float b) {
  float x = threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 16; ++u) x = fmaf(x, a, b);
  }
  long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

int main() {
  double* d; long long* c; float* f;
  cudaMalloc(&d, 1024); cudaMalloc(&c, 8); cudaMalloc(&f, 1024);
  long long h;
  const int iters = 256, n = iters * 16;
#define RUN(name, call, per)                                              \
  call; call; cudaDeviceSynchronize();                                     \
  cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);                            \
  printf("%-34s %8.2f cycles per %s\n", name, (double)h / n, per);
  RUN("DFMA dependent (1 chain)", (dfma_chain<1><<<1, 32>>>(d, c, iters, 0.999, 1e-3)), "DFMA");
  RUN("DFMA 2 chains", (dfma_chain<2><<<1, 32>>>(d, c, iters, 0.999, 1e-3)), "round of 2");
  RUN("DFMA 4 chains", (dfma_chain<4><<<1, 32>>>(d, c, iters, 0.999, 1e-3)), "round of 4");
  RUN("DFMA 8 chains", (dfma_chain<8><<<1, 32>>>(d, c, iters, 0.999, 1e-3)), "round of 8");
  RUN("DFMA 16 chains", (dfma_chain<16><<<1, 32>>>(d, c, iters, 0.999, 1e-3)), "round of 16");
  RUN("MUFU.RCP64H + DADD", (rcp_chain<<<1, 32>>>(d, c, iters, 0)), "link");
  RUN("rcp (MUFU + 3 DFMA) + DADD", (rcp_chain<<<1, 32>>>(d, c, iters, 1)), "link");
  RUN("IEEE 1/x + DADD", (div_chain<<<1, 32>>>(d, c, iters)), "link");
  RUN("SHFL (64-bit) + DADD", (shfl_chain<<<1, 32>>>(d, c, iters)), "link");
  RUN("STS + LDS (64-bit) + DADD", (lds_chain<<<1, 32>>>(d, c, iters)), "link");
  RUN("FFMA dependent", (ffma_chain<<<1, 32>>>(f, c, iters, 0.999f, 1e-3f)), "FFMA");
  // two warps on the same SM sub-partition?  (4 warps of a 128-thread CTA land on the 4 sub-partitions)
  RUN("DFMA 1 chain, 8 warps/CTA", (dfma_chain<1><<<1, 256>>>(d, c, iters, 0.999, 1e-3)), "DFMA (2 warps per sub-partition)");
  RUN("DFMA 1 chain, 16 warps/CTA", (dfma_chain<1><<<1, 512>>>(d, c, iters, 0.999, 1e-3)), "DFMA (4 warps per sub-partition)");
  RUN("DFMA 4 chains, 16 warps/CTA", (dfma_chain<4><<<1, 512>>>(d, c, iters, 0.999, 1e-3)), "round of 4 (4 warps per sub-partition)");
  return 0;
}
