// Micro-benchmark (not product code): which hardware warp slot (%warpid) -- and so which SM sub-partition, slot % 4 --
// do the warps of a 96-thread CTA get when 4 such CTAs are resident per SM (the fenrir_ws_kernel geometry)?
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(96, 4) k(int* out, int spin) {
  extern __shared__ double s[];
  unsigned smid, wid;
  asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
  asm volatile("mov.u32 %0, %%warpid;" : "=r"(wid));
  double a = threadIdx.x;
  for (int i = 0; i < spin; ++i) a = a * 1.0000001 + 1e-9;       // keep every CTA resident for a while
  if ((threadIdx.x & 31) == 0) {
    const int w = blockIdx.x * 3 + (threadIdx.x >> 5);
    out[2 * w] = (int)smid; out[2 * w + 1] = (int)wid;
  }
  if (a == 12345.678) s[0] = a;
}
int main() {
  const int grid = 512;
  int* d; cudaMalloc(&d, grid * 3 * 2 * sizeof(int));
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 40000);
  k<<<grid, 96, 40000>>>(d, 200000);
  static int h[512 * 3 * 2];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  int hist[3][4] = {};
  for (int w = 0; w < grid * 3; ++w) hist[w % 3][h[2 * w + 1] % 4]++;
  for (int r = 0; r < 3; ++r) printf("warp %d of its CTA: hardware slot %% 4 = 0/1/2/3 -> %d %d %d %d\n", r, hist[r][0], hist[r][1], hist[r][2], hist[r][3]);
  printf("first CTAs (smid: slots): ");
  for (int c = 0; c < 12; ++c) printf("[sm %d: %d %d %d] ", h[6 * c], h[6 * c + 1], h[6 * c + 3], h[6 * c + 5]);
  printf("\n%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
