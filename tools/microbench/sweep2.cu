// Microbenchmark of the tiled register-resident sweep loop (invert_spd_tiles) in isolation, with ablations.
#include <cstdio>
#include <cuda_runtime.h>

constexpr int GR = 16, R = 4, GC = 8, C = 8, NVP = 64, NT = 128;

// VARIANT bits: 1 = no barrier, 2 = no publication, 4 = no reciprocal, 8 = no dg tracking
__device__ __forceinline__ double fast_rcp(double d) {
  double x;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(d));
  double e = fma(-d, x, 1.0);
  x = fma(x, e, x);
  e = fma(-d, x, 1.0);
  x = fma(x, e, x);
  return x;
}

template <int V>
__global__ void __launch_bounds__(NT, 4) sweep_kernel(const double* H, double* out, long long* cycles, int nv, int reps) {
  __shared__ __align__(16) double ckbuf[2 * (NVP + 2)];
  __shared__ double Hs[NVP * (NVP + 1)];
  const int tid = threadIdx.x, tr = tid / GC, tc = tid % GC, ld = NVP + 1;
  for (int e = tid; e < nv * nv; e += NT) Hs[(e / nv) * ld + e % nv] = H[e];
  __syncthreads();
  long long total = 0;
  double a[R][C], dg[R];
  for (int rep = 0; rep < reps; rep++) {
#pragma unroll
    for (int i = 0; i < R; i++) {
      const int r = tr + GR * i;
#pragma unroll
      for (int j = 0; j < C; j++) {
        const int c = 2 * GC * (j / 2) + 2 * tc + (j & 1);
        a[i][j] = (r < nv && c < nv) ? Hs[r * ld + c] : (r == c ? 1.0 : 0.0);
      }
      dg[i] = (r < nv) ? Hs[r * ld + r] : 1.0;
    }
    double* const buf0 = ckbuf;
    double* const buf1 = ckbuf + (NVP + 2);
    if (V & 2) {
      for (int i = tid; i < 2 * (NVP + 2); i += NT) ckbuf[i] = 1e-3 * (i % 7 + 1);
    }
    double dinv_mine = fast_rcp(dg[0]);
    bool bad = false;
    __syncthreads();
    long long t0 = clock64();
#pragma unroll
    for (int i = 0; i < R; i++) {
      if (GR * i >= nv) break;
#pragma unroll 1
      for (int q = 0; q < GR; q++) {
        const int p = GR * i + q;
        if (p >= nv) break;
        double* const cur = (q & 1) ? buf1 : buf0;
        if (!(V & 2)) {
          if (tr == q) {
#pragma unroll
            for (int j2 = 0; j2 < C / 2; j2++)
              *reinterpret_cast<double2*>(cur + 2 * GC * j2 + 2 * tc) = make_double2(a[i][2 * j2], a[i][2 * j2 + 1]);
            if (tc == 0) {
              cur[p] = dg[i] - 1.0;
              cur[NVP] = (V & 4) ? 0.4 : dinv_mine;
            }
          }
        }
        if (!(V & 1)) __syncthreads();
        const double dinv = cur[NVP];
        bad = bad || !(dinv > 0.0 && dinv < 1e300);
        double u[R];
#pragma unroll
        for (int ii = 0; ii < R; ii++) {
          const double c = cur[tr + GR * ii];
          u[ii] = -c * dinv;
          if (!(V & 8)) dg[ii] = fma(u[ii], c, dg[ii]);
        }
        if (!(V & 4)) {
          dinv_mine = fast_rcp((q + 1 < GR) ? dg[i] : dg[(i + 1 < R) ? i + 1 : i]);
        }
#pragma unroll
        for (int j2 = 0; j2 < C / 2; j2++) {
          const double2 v = *reinterpret_cast<const double2*>(cur + 2 * GC * j2 + 2 * tc);
#pragma unroll
          for (int ii = 0; ii < R; ii++) {
            a[ii][2 * j2] = fma(u[ii], v.x, a[ii][2 * j2]);
            a[ii][2 * j2 + 1] = fma(u[ii], v.y, a[ii][2 * j2 + 1]);
          }
        }
      }
    }
    long long t1 = clock64();
    total += t1 - t0;
    double s = bad ? 1.0 : dinv_mine;
#pragma unroll
    for (int i = 0; i < R; i++)
#pragma unroll
      for (int j = 0; j < C; j++) s += a[i][j] + dg[i];
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
  printf("%-52s ctas/SM %d: %7.1f cycles per pivot (%s)\n", what, ctas_per_sm, mean / reps / nv, cudaGetErrorString(e));
  cudaFree(out);
  cudaFree(cyc);
}

int main() {
  const int nv = 60;
  static double H[60 * 60];
  for (int i = 0; i < nv; i++)
    for (int j = 0; j < nv; j++) H[i * nv + j] = (i == j) ? 2.0 + 0.01 * i : 0.01 * ((i * 7 + j * 13) % 11) / 11.0;
  for (int i = 0; i < nv; i++)
    for (int j = 0; j < i; j++) H[i * nv + j] = H[j * nv + i];
  double* dH;
  cudaMalloc(&dH, sizeof(H));
  cudaMemcpy(dH, H, sizeof(H), cudaMemcpyHostToDevice);
  for (int c : {1, 2, 4}) {
    run<0>("full", dH, nv, c);
    run<4>("no reciprocal", dH, nv, c);
    run<8>("no dg tracking (and so garbage reciprocal input)", dH, nv, c);
    run<2>("no publication", dH, nv, c);
    run<1>("no barrier", dH, nv, c);
    run<2 | 4 | 8>("no publication, no rcp, no dg (barrier kept)", dH, nv, c);
    run<1 | 2 | 4 | 8>("loads + DFMAs only", dH, nv, c);
  }
  return 0;
}
