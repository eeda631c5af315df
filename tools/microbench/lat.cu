// Dependent-issue latencies of the building blocks of the pivot chain on sm_100a (one warp, or one 128-thread CTA for
// the barrier): cycles per link of a chain of N dependent operations.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ double fast_rcp(double d) {
  double x;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(d));
  double e = fma(-d, x, 1.0);
  x = fma(x, e, x);
  e = fma(-d, x, 1.0);
  x = fma(x, e, x);
  return x;
}
__device__ __forceinline__ double rcp_seed(double d) {
  double x;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(d));
  return x;
}
template <int OP>
__global__ void k(double* out, long long* cyc, int iters, double seed) {
  __shared__ double sh[64];
  double x = seed + threadIdx.x * 1e-9, y = 1.0000001;
  sh[threadIdx.x & 63] = x;
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < iters; i++) {
    if (OP == 0) x = fast_rcp(x);
    if (OP == 1) x = rcp_seed(x) + 0.5;                                    // MUFU.RCP64H + DADD
    if (OP == 2) x = __shfl_sync(0xffffffffu, x, (threadIdx.x + 1) & 31);  // 64-bit shuffle (2 SHFL)
    if (OP == 3) x = fma(x, y, 1e-9);                                      // DFMA
    if (OP == 4) { sh[threadIdx.x & 63] = x; __syncwarp(); x = sh[(threadIdx.x + 1) & 63]; __syncwarp(); }  // STS -> LDS
    if (OP == 5) { sh[threadIdx.x & 63] = x; __syncthreads(); x = sh[(threadIdx.x + 1) & 63]; __syncthreads(); }  // 2 barriers
    if (OP == 6) { double c0 = x, c1 = y; asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(y), "d"(y)); x = c0; y = c1 * 1e-30 + 1.0; }
    if (OP == 7) x = 1.0 / x;                                              // IEEE division
    if (OP == 8) { x = __shfl_sync(0xffffffffu, x, (threadIdx.x + 1) & 31); x = fma(x, y, 1e-9); }  // shuffle + DFMA
  }
  long long t1 = clock64();
  out[threadIdx.x] = x + y;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
template <int OP>
void run(const char* name, int threads) {
  double* out; long long* cyc;
  cudaMalloc(&out, 8 * 1024); cudaMalloc(&cyc, 8);
  const int iters = 4096;
  k<OP><<<1, threads>>>(out, cyc, iters, 1.3);
  k<OP><<<1, threads>>>(out, cyc, iters, 1.3);
  cudaDeviceSynchronize();
  long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("%-44s %3d threads: %.1f cycles per link\n", name, threads, (double)h / iters);
}
int main() {
  run<0>("fast_rcp (MUFU.RCP64H + 2 Newton steps)", 32);
  run<1>("MUFU.RCP64H + DADD", 32);
  run<2>("64-bit shuffle", 32);
  run<3>("DFMA", 32);
  run<8>("64-bit shuffle + DFMA", 32);
  run<4>("STS -> syncwarp -> LDS -> syncwarp", 32);
  run<5>("STS -> bar -> LDS -> bar", 128);
  run<5>("STS -> bar -> LDS -> bar", 256);
  run<6>("DMMA.8x8x4 (+ DFMA)", 32);
  run<7>("1.0 / x (IEEE)", 32);
  return 0;
}
