// Microbenchmark: FP64 FMA issue rate per SM sub-partition on sm_100a as a function of resident warps and
// of the number of independent accumulators per thread (ILP).  Prints cycles per warp-level DFMA.
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void dfma_kernel(double* out, long long* cycles, int iters, double x, double y) {
  double acc[ILP];
#pragma unroll
  for (int i = 0; i < ILP; i++) acc[i] = threadIdx.x + i;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++) acc[i] = fma(acc[i], x, y);
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int ILP>
void run(int threads, int iters) {
  double* out;
  long long* cyc;
  cudaMalloc(&out, sizeof(double) * 148 * threads);
  cudaMalloc(&cyc, sizeof(long long) * 148);
  dfma_kernel<ILP><<<148, threads>>>(out, cyc, iters, 1.0000001, 1e-9);
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  cudaEventRecord(e0);
  dfma_kernel<ILP><<<148, threads>>>(out, cyc, iters, 1.0000001, 1e-9);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  long long h[148];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double c = (double)h[0];
  int warps = threads / 32;
  double per_smsp_warp_instr = (double)iters * ILP * ((warps + 3) / 4);
  double tflops = 2.0 * 148.0 * threads * (double)iters * ILP / (ms * 1e-3) / 1e12;
  printf("threads/SM %4d (warps/SMSP %2d) ILP %2d: %.2f cycles per warp-DFMA per SMSP, per-warp interval %.2f cycles, %.2f TFLOP/s\n",
         threads, (warps + 3) / 4, ILP, c / per_smsp_warp_instr, c / ((double)iters * ILP), tflops);
  cudaFree(out);
  cudaFree(cyc);
}

int main() {
  const int iters = 20000;
  for (int threads : {32, 128, 256, 512, 1024}) {
    run<1>(threads, iters);
    run<4>(threads, iters);
    run<16>(threads, iters);
    run<32>(threads, iters);
  }
  return 0;
}
