// Microbenchmark of the register-resident sweep loop (invert_spd_rows) in isolation: cycles per pivot
// for ablated variants, at 1..4 CTAs per SM.
#include <cstdio>
#include <cuda_runtime.h>

constexpr int NVP = 64, GC = 2, C = NVP / GC, NT = NVP * GC;

template <int VARIANT>
__global__ void __launch_bounds__(NT, 4) sweep_kernel(const double* H, double* out, long long* cycles, int nv, int reps) {
  __shared__ __align__(16) double ckbuf[2 * (NVP + 2)];
  __shared__ double Hs[NVP * (NVP + 1)];
  const int tid = threadIdx.x, tr = tid / GC, tc = tid % GC, ld = NVP + 1;
  for (int e = tid; e < nv * nv; e += NT) Hs[(e / nv) * ld + e % nv] = H[e];
  __syncthreads();
  long long total = 0;
  double a[C];
  for (int rep = 0; rep < reps; rep++) {
#pragma unroll
    for (int j2 = 0; j2 < C / 2; j2++)
#pragma unroll
      for (int e = 0; e < 2; e++) {
        const int c = 2 * GC * j2 + 2 * tc + e;
        a[2 * j2 + e] = (tr < nv && c < nv) ? Hs[tr * ld + c] : (tr == c ? 1.0 : 0.0);
      }
    double dg = (tr < nv) ? Hs[tr * ld + tr] : 1.0;
    const bool holds_diag = (tc == (tr / 2) % GC);
    double* const buf0 = ckbuf;
    double* const buf1 = ckbuf + (NVP + 2);
    if (tr == 0) {
#pragma unroll
      for (int j2 = 0; j2 < C / 2; j2++)
        *reinterpret_cast<double2*>(buf0 + 2 * GC * j2 + 2 * tc) = make_double2(a[2 * j2], a[2 * j2 + 1]);
      if (holds_diag) { buf0[0] = dg - 1.0; buf0[NVP] = __drcp_rn(dg); }
    }
    if (VARIANT == 2 || VARIANT == 4) {  // no publication in the loop: fill both buffers with something finite
      for (int i = tid; i < 2 * (NVP + 2); i += NT) ckbuf[i] = 1e-3 * (i % 7 + 1);
    }
    __syncthreads();
    long long t0 = clock64();
    bool bad = false;
#pragma unroll 1
    for (int p = 0; p < nv; p++) {
      const double* const cur = (p & 1) ? buf1 : buf0;
      double* const nxt = (p & 1) ? buf0 : buf1;
      if (VARIANT != 3 && VARIANT != 4) __syncthreads();
      const double dinv = cur[NVP];
      const double cu = cur[tr];
      bad = bad || !(dinv > 0.0 && dinv < 1e300);
      const double u = -cu * dinv;
      dg = fma(u, cu, dg);
      const bool own_next = (tr == p + 1);
      double dnext_inv = 0.0;
      if (VARIANT == 0 || VARIANT == 3)
        if (own_next && holds_diag) dnext_inv = __drcp_rn(dg);
#pragma unroll
      for (int j2 = 0; j2 < C / 2; j2++) {
        const double2 v = *reinterpret_cast<const double2*>(cur + 2 * GC * j2 + 2 * tc);
        a[2 * j2] = fma(u, v.x, a[2 * j2]);
        a[2 * j2 + 1] = fma(u, v.y, a[2 * j2 + 1]);
      }
      if (VARIANT == 0 || VARIANT == 1 || VARIANT == 3) {
        if (own_next) {
#pragma unroll
          for (int j2 = 0; j2 < C / 2; j2++)
            *reinterpret_cast<double2*>(nxt + 2 * GC * j2 + 2 * tc) = make_double2(a[2 * j2], a[2 * j2 + 1]);
          if (holds_diag) {
            nxt[p + 1] = dg - 1.0;
            nxt[NVP] = (VARIANT == 1) ? 1.0 / 3.0 : dnext_inv;
          }
        }
      }
    }
    long long t1 = clock64();
    total += t1 - t0;
    double s = bad ? 1.0 : 0.0;
#pragma unroll
    for (int j = 0; j < C; j++) s += a[j];
    out[(blockIdx.x * NT + tid)] = s;
    __syncthreads();
  }
  if (tid == 0) cycles[blockIdx.x] = total;
}

template <int V>
void run(const char* what, const double* dH, int nv, int ctas_per_sm) {
  double* out;
  long long* cyc;
  const int grid = 148 * ctas_per_sm, reps = 50;
  cudaMalloc(&out, sizeof(double) * grid * NT);
  cudaMalloc(&cyc, sizeof(long long) * grid);
  sweep_kernel<V><<<grid, NT>>>(dH, out, cyc, nv, reps);
  cudaDeviceSynchronize();
  sweep_kernel<V><<<grid, NT>>>(dH, out, cyc, nv, reps);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[148 * 4];
  cudaMemcpy(h, cyc, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
  double mean = 0;
  for (int i = 0; i < grid; i++) mean += (double)h[i];
  mean /= grid;
  printf("%-44s ctas/SM %d: %7.1f cycles per pivot (%s)\n", what, ctas_per_sm, mean / reps / nv, cudaGetErrorString(e));
  cudaFree(out);
  cudaFree(cyc);
}

int main() {
  const int nv = 60;
  static double H[60 * 60];
  // SPD: diagonally dominant
  for (int i = 0; i < nv; i++)
    for (int j = 0; j < nv; j++) H[i * nv + j] = (i == j) ? 2.0 + 0.01 * i : 0.01 * ((i * 7 + j * 13) % 11) / 11.0;
  for (int i = 0; i < nv; i++)
    for (int j = 0; j < i; j++) H[i * nv + j] = H[j * nv + i];
  double* dH;
  cudaMalloc(&dH, sizeof(H));
  cudaMemcpy(dH, H, sizeof(H), cudaMemcpyHostToDevice);
  for (int c : {1, 2, 4}) {
    run<0>("V0 full", dH, nv, c);
    run<1>("V1 no reciprocal", dH, nv, c);
    run<2>("V2 no publication (barrier kept)", dH, nv, c);
    run<3>("V3 no barrier (publication kept)", dH, nv, c);
    run<4>("V4 loads + DFMAs only", dH, nv, c);
  }
  return 0;
}
