// Microbenchmark: FP64 tensor-core MMA (mma.sync.m8n8k4.f64 -> DMMA.8x8x4 in SASS) on sm_100a.
// Prints cycles per warp-level DMMA per SM sub-partition as a function of resident warps and of the number of
// independent accumulator fragments per warp (ILP), the dependent-issue latency (ILP 1, one warp), and the same
// for the loop shape of the grouped sweep (shared-memory fragment loads + DMMAs).  One DMMA.8x8x4 = 256 FMAs,
// i.e. the work of 8 warp-wide DFMAs.
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

template <int ILP>
__global__ void dmma_kernel(double* out, long long* cycles, int iters, double x, double y) {
  double c0[ILP], c1[ILP];
#pragma unroll
  for (int i = 0; i < ILP; i++) { c0[i] = threadIdx.x + i; c1[i] = i; }
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++) dmma(c0[i], c1[i], x, y);
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) s += c0[i] + c1[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

// sweep-shaped loop: per step NA + NB fragment loads from shared memory feed NA*NB DMMAs
template <int NA, int NB>
__global__ void dmma_lds_kernel(double* out, long long* cycles, int iters) {
  __shared__ double S[8 * 68 * 2];
  for (int i = threadIdx.x; i < 8 * 68 * 2; i += blockDim.x) S[i] = 1e-9 * i;
  double c0[NA][NB], c1[NA][NB];
#pragma unroll
  for (int i = 0; i < NA; i++)
#pragma unroll
    for (int j = 0; j < NB; j++) { c0[i][j] = threadIdx.x; c1[i][j] = j; }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const double* base = S + (lane & 3) * 68 + (lane >> 2);
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
    const double* p = base + (it & 1) * 4 * 68;
    double a[NA], b[NB];
#pragma unroll
    for (int i = 0; i < NA; i++) a[i] = p[8 * i];
#pragma unroll
    for (int j = 0; j < NB; j++) b[j] = p[8 * 68 + 8 * (j % 8)];
#pragma unroll
    for (int i = 0; i < NA; i++)
#pragma unroll
      for (int j = 0; j < NB; j++) dmma(c0[i][j], c1[i][j], a[i], b[j]);
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < NA; i++)
#pragma unroll
    for (int j = 0; j < NB; j++) s += c0[i][j] + c1[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <class F>
void time_it(const char* what, int threads, double dmma_per_warp, F launch) {
  double* out;
  long long* cyc;
  cudaMalloc(&out, sizeof(double) * 148 * threads);
  cudaMalloc(&cyc, sizeof(long long) * 148);
  launch(out, cyc);
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  cudaEventRecord(e0);
  launch(out, cyc);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  long long h[148];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  const int warps = threads / 32, wps = (warps + 3) / 4;
  const double c = (double)h[0];
  const double tflops = 2.0 * 256.0 * 148.0 * warps * dmma_per_warp / (ms * 1e-3) / 1e12;
  printf("%s threads/SM %4d (warps/SMSP %2d): %.2f cycles per DMMA per SMSP, per-warp interval %.2f cycles, %.2f TFLOP/s\n",
         what, threads, wps, c / (dmma_per_warp * wps), c / dmma_per_warp, tflops);
  cudaFree(out);
  cudaFree(cyc);
}

template <int ILP>
void run(int threads, int iters) {
  char what[64];
  snprintf(what, sizeof(what), "dmma ILP %2d", ILP);
  time_it(what, threads, (double)iters * ILP,
          [&](double* o, long long* c) { dmma_kernel<ILP><<<148, threads>>>(o, c, iters, 1.0000001, 1e-9); });
}
template <int NA, int NB>
void run_lds(int threads, int iters) {
  char what[64];
  snprintf(what, sizeof(what), "lds %d+%d -> %2d dmma", NA, NB, NA * NB);
  time_it(what, threads, (double)iters * NA * NB,
          [&](double* o, long long* c) { dmma_lds_kernel<NA, NB><<<148, threads>>>(o, c, iters); });
}

int main() {
  const int iters = 20000;
  for (int threads : {32, 128, 256, 512, 1024}) {
    run<1>(threads, iters);
    run<2>(threads, iters);
    run<4>(threads, iters);
    run<9>(threads, iters);
    run<18>(threads, iters);
  }
  for (int threads : {128, 256, 512}) {
    run_lds<1, 9>(threads, iters);
    run_lds<2, 5>(threads, iters);
    run_lds<2, 9>(threads, iters);
    run_lds<3, 6>(threads, iters);
  }
  return 0;
}
