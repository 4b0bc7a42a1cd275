// Micro-benchmark (not product code): HBM write throughput of the solve_mv output pattern.
// 65,536 thetas x 801 rows x (6 + 18) doubles, theta-outermost layout, written backwards in time by one warp per
// 32 thetas, K rows per burst.  Variants: lane-per-element runs (product), 16-byte, per-thread rows.
#include <cstdio>
#include <cuda_runtime.h>
typedef long long i64;
template <int K, int MODE>
__global__ void __launch_bounds__(32) k(double* __restrict__ mean, double* __restrict__ var, i64 B, int N) {
  __shared__ double buf[(K < 6 ? K : 6) * 18 * 33];
  const int lane = threadIdx.x;
  const i64 theta0 = (i64)blockIdx.x * 32;
  for (int i = lane; i < (K < 6 ? K : 6) * 18 * 33; i += 32) buf[i] = i;
  __syncwarp();
  for (int j = (N - 1) / K; j >= 0; --j) {
    const int n0 = j * K, cnt = (N - n0) < K ? (N - n0) : K;
    if (MODE == 0) {   // product pattern: per theta, consecutive lanes store consecutive elements of the run
      {
        const int run = cnt * 6; double* dst = mean + (theta0 * (N + 1) + n0) * 6 + lane;
        for (int th = 0; th < 32; ++th) { for (int r = lane; r < run; r += 32) dst[r - lane] = buf[(r % 18) * 33 + th]; dst += (i64)(N + 1) * 6; }
      }
      {
        const int run = cnt * 18; double* dst = var + (theta0 * (N + 1) + n0) * 18 + lane;
#pragma unroll 4
        for (int th = 0; th < 32; ++th) {
          for (int r = lane; r < run; r += 32) dst[r - lane] = buf[(r % 54) * 33 + th];
          dst += (i64)(N + 1) * 18;
        }
      }
    } else if (MODE == 1) {   // per-thread rows (no staging): each lane writes its own theta's rows, 16 B stores
      for (int s = 0; s < cnt; ++s) {
        double2* m = (double2*)(mean + ((theta0 + lane) * (N + 1) + n0 + s) * 6);
        double2* v = (double2*)(var + ((theta0 + lane) * (N + 1) + n0 + s) * 18);
#pragma unroll
        for (int e = 0; e < 3; ++e) m[e] = make_double2(buf[e * 33 + lane], 1.0);
#pragma unroll
        for (int e = 0; e < 9; ++e) v[e] = make_double2(buf[e * 33 + lane], 2.0);
      }
    } else {   // flattened over (theta, element): 16-byte stores, consecutive lanes -> consecutive 16 B of a run
      const int runm = cnt * 3, runv = cnt * 9;   // in double2 units
      for (int c = lane; c < 32 * runm; c += 32) {
        const int th = c / runm, r = c - th * runm;
        ((double2*)(mean + ((theta0 + th) * (N + 1) + n0) * 6))[r] = make_double2(buf[r * 33 + th], 1.0);
      }
      for (int c = lane; c < 32 * runv; c += 32) {
        const int th = c / runv, r = c - th * runv;
        ((double2*)(var + ((theta0 + th) * (N + 1) + n0) * 18))[r] = make_double2(buf[r * 33 + th], 2.0);
      }
    }
    __syncwarp();
  }
}
// MODE 3: TMA bulk stores.  Staging is theta-major: region[lane] = [mean: K*6][var: K*18] doubles, pitch 74.
template <int K>
__global__ void __launch_bounds__(32) kb(double* __restrict__ mean, double* __restrict__ var, i64 B, int N) {
  constexpr int PITCH = K * 24 + 2;
  __shared__ __align__(16) double buf[32 * PITCH];
  const int lane = threadIdx.x;
  const i64 theta = (i64)blockIdx.x * 32 + lane;
  for (int i = 0; i < K * 24; ++i) buf[lane * PITCH + i] = i;
  unsigned sm = (unsigned)__cvta_generic_to_shared(buf + lane * PITCH);
  unsigned sv = sm + K * 6 * 8;
  for (int j = (N - 1) / K; j >= 0; --j) {
    const int n0 = j * K, cnt = (N - n0) < K ? (N - n0) : K;
    // (product code would write the staged rows here)
    buf[lane * PITCH + (j % (K * 24))] = j;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    double* gm = mean + (theta * (N + 1) + n0) * 6;
    double* gv = var + (theta * (N + 1) + n0) * 18;
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gm), "r"(sm), "r"(cnt * 48) : "memory");
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gv), "r"(sv), "r"(cnt * 144) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  }
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
template <int K> void runb(double* mean, double* var, i64 B, int N) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e9;
  for (int r = 0; r < 4; ++r) {
    cudaEventRecord(e0); kb<K><<<(unsigned)(B / 32), 32>>>(mean, var, B, N); cudaEventRecord(e1);
    cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
  }
  double gb = (double)B * (N + 1) * 24 * 8 / 1e9;
  printf("%-28s K=%d  %.3f ms  %.2f TB/s  (%s)\n", "TMA bulk store per theta", K, best, gb / best, cudaGetErrorString(cudaGetLastError()));
}
template <int K, int MODE> void run(double* mean, double* var, i64 B, int N, const char* name) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e9;
  for (int r = 0; r < 4; ++r) {
    cudaEventRecord(e0); k<K, MODE><<<(unsigned)(B / 32), 32>>>(mean, var, B, N); cudaEventRecord(e1);
    cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
  }
  double gb = (double)B * (N + 1) * 24 * 8 / 1e9;
  printf("%-28s K=%d  %.3f ms  %.2f TB/s  (%s)\n", name, K, best, gb / best, cudaGetErrorString(cudaGetLastError()));
}
int main() {
  const i64 B = 65536; const int N = 800;
  double *mean, *var;
  cudaMalloc(&mean, B * (N + 1) * 6 * 8); cudaMalloc(&var, B * (N + 1) * 18 * 8);
  run<3, 0>(mean, var, B, N, "runs, lane-per-element");
  run<3, 1>(mean, var, B, N, "per-thread rows 16B");
  run<3, 2>(mean, var, B, N, "flattened 16B");
  runb<1>(mean, var, B, N);
  runb<3>(mean, var, B, N);
  runb<6>(mean, var, B, N);

  run<6, 0>(mean, var, B, N, "runs, lane-per-element");
  run<12, 0>(mean, var, B, N, "runs, lane-per-element");
  run<24, 0>(mean, var, B, N, "runs, lane-per-element");
  run<48, 0>(mean, var, B, N, "runs, lane-per-element");
  run<100, 0>(mean, var, B, N, "runs, lane-per-element");
  cudaMemset(var, 0, B * (N + 1) * 18 * 8);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0); cudaMemsetAsync(var, 0, B * (N + 1) * 18 * 8); cudaMemsetAsync(mean, 0, B * (N + 1) * 6 * 8); cudaEventRecord(e1);
  cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1);
  printf("cudaMemset of the same 10.1 GB: %.3f ms  %.2f TB/s\n", ms, (double)B * (N + 1) * 24 * 8 / 1e9 / ms);
  return 0;
}
