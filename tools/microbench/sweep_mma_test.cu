// Stand-alone check + timing of the two register-resident inversions of mpc_core.h on the GPU:
//   invert_spd_tiles  (one rank-1 update per pivot on the FP64 FMA pipe)
//   invert_spd_mma    (grouped exact sweep, rank-8 updates as DMMA.8x8x4)
// Random SPD matrices with half of the spectrum at the regularisation floor (the MPC Hessian's shape of trouble);
// prints |H Hinv - I|, the distance between the two results and the time per problem at full occupancy.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o sweep_mma_test sweep_mma_test.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>

#define MPC_SWEEP_CLK 1
#include "../../quadruped_ctrl_b200/csrc/mpc_core.h"

struct Params {
  const double* H;   // [n_prob][nv*nv]
  const double* g;   // [n_prob][nv]
  double* Hinv;      // [n_prob][nv*nv]
  double* x;         // [n_prob][nv]
  int* status;
  int nv, n_prob, reps;
  long long* clk;    // [grid][4]: cycles in the phases A, B, C of invert_spd_mma and in the whole inversion, summed
  mpc::Layout L;
};

template <int NT, int GR, int R, int GC, int C, int NWS, int NB, bool PK, int MINB, bool MMA>
__global__ void __launch_bounds__(NT, MINB) sweep_test_kernel(const __grid_constant__ Params P) {
  extern __shared__ __align__(128) char smem[];
  mpc::Work k = mpc::carve(P.L, smem, nullptr);
  const int tid = threadIdx.x, nv = P.nv;
  __shared__ long long clk[40];
  if (tid < 40) clk[tid] = 0;
  k.clk = clk;
  for (int prob = blockIdx.x; prob < P.n_prob; prob += gridDim.x) {
    if (tid == 0) { k.sc->nv = nv; k.sc->status = 0; }
    for (int e = tid; e < nv * nv; e += NT) {
      const int i = e / nv, j = e - i * nv;
      if (!PK || i >= j) k.Hm[mpc::hixT<PK>(k.ld, i, j)] = P.H[(size_t)prob * nv * nv + e];
    }
    for (int e = tid; e < nv; e += NT) k.g[e] = P.g[(size_t)prob * nv + e];
    __syncthreads();
    // reps is odd: H -> Hinv -> H -> ... -> Hinv in place (the inverse of an SPD matrix is SPD), so the timed loop holds
    // nothing but inversions
    for (int rep = 0; rep < P.reps; rep++) {
      const long long t0 = clock64();
      if constexpr (MMA) mpc::invert_spd_mma<NT, NWS, NB, PK>(k, tid, true);
      else mpc::invert_spd_tiles<GR, R, GC, C, PK>(k, tid, true);
      __syncthreads();
      if (tid == 0) clk[3] += clock64() - t0;
    }
    for (int e = tid; e < nv * nv; e += NT) {
      const int i = e / nv, j = e - i * nv;
      P.Hinv[(size_t)prob * nv * nv + e] = k.Hm[mpc::hixT<PK>(k.ld, i, j)];
    }
    for (int e = tid; e < nv; e += NT) P.x[(size_t)prob * nv + e] = k.x[e];
    if (tid == 0) P.status[prob] = k.sc->status;
    __syncthreads();
  }
  if (tid < 40 && P.clk) P.clk[40 * blockIdx.x + tid] = clk[tid];
}

static void make_problem(int nv, unsigned seed, double* H, double* g) {
  srand(seed);
  const int rk = nv / 2;
  std::vector<double> M((size_t)nv * rk);
  for (auto& v : M) v = (rand() / (double)RAND_MAX - 0.5);
  for (int i = 0; i < nv; i++)
    for (int j = 0; j < nv; j++) {
      double s = 0;
      for (int q = 0; q < rk; q++) s += M[(size_t)i * rk + q] * M[(size_t)j * rk + q];
      H[(size_t)i * nv + j] = 2.0 * s + (i == j ? 8e-5 : 0.0);
    }
  for (int i = 0; i < nv; i++) g[i] = rand() / (double)RAND_MAX - 0.5;
}

template <class K>
static float run(K kern, int threads, int grid, const Params& P, size_t smem) {
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  kern<<<grid, threads, smem>>>(P);
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  cudaEventRecord(e0);
  kern<<<grid, threads, smem>>>(P);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) printf("CUDA error: %s\n", cudaGetErrorString(err));
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  return ms;
}

static void check(const char* name, int nv, int n_prob, const std::vector<double>& H, const std::vector<double>& g,
                  const std::vector<double>& Hi, const std::vector<double>& x, const std::vector<int>& st) {
  double worst = 0, worstx = 0;
  int bad = 0;
  for (int p = 0; p < n_prob; p++) {
    if (st[p] != 0) bad++;
    const double* h = &H[(size_t)p * nv * nv];
    const double* hi = &Hi[(size_t)p * nv * nv];
    for (int i = 0; i < nv; i++) {
      double sx = 0;
      for (int j = 0; j < nv; j++) {
        double s = 0;
        for (int q = 0; q < nv; q++) s += h[(size_t)i * nv + q] * hi[(size_t)q * nv + j];
        worst = fmax(worst, fabs(s - (i == j ? 1.0 : 0.0)));
        sx += hi[(size_t)i * nv + j] * g[(size_t)p * nv + j];
      }
      worstx = fmax(worstx, fabs(x[(size_t)p * nv + i] + sx) / (1.0 + fabs(sx)));
    }
  }
  printf("  %-6s nv=%3d: max |H Hinv - I| %.2e, max |x + Hinv g| (rel) %.2e, bad status %d\n", name, nv, worst, worstx, bad);
}

template <int NT, int GR, int R, int GC, int C, int NWS, int NB, bool PK, int MINB>
static void test_class(int nv_cap, int nv, int h, int m_cap) {
  const int npad = 8 * NB;
  const mpc::Layout L = mpc::make_layout(h, nv_cap, m_cap, 1, npad, PK ? 1 : 0, 1);
  const size_t smem = L.fast_bytes;
  const int n_check = 8, n_time = 148 * MINB * 16;
  std::vector<double> H((size_t)n_time * nv * nv), g((size_t)n_time * nv);
  for (int p = 0; p < n_check; p++) make_problem(nv, 100 + p, &H[(size_t)p * nv * nv], &g[(size_t)p * nv]);
  for (int p = n_check; p < n_time; p++) {
    memcpy(&H[(size_t)p * nv * nv], &H[(size_t)(p % n_check) * nv * nv], sizeof(double) * nv * nv);
    memcpy(&g[(size_t)p * nv], &g[(size_t)(p % n_check) * nv], sizeof(double) * nv);
  }
  Params P;
  double *dH, *dg, *dHi, *dx;
  int* dst;
  long long* dclk;
  cudaMalloc(&dclk, sizeof(long long) * 40 * 148 * MINB);
  cudaMalloc(&dH, H.size() * 8);
  cudaMalloc(&dg, g.size() * 8);
  cudaMalloc(&dHi, H.size() * 8);
  cudaMalloc(&dx, g.size() * 8);
  cudaMalloc(&dst, n_time * 4);
  cudaMemcpy(dH, H.data(), H.size() * 8, cudaMemcpyHostToDevice);
  cudaMemcpy(dg, g.data(), g.size() * 8, cudaMemcpyHostToDevice);
  P.H = dH; P.g = dg; P.Hinv = dHi; P.x = dx; P.status = dst; P.nv = nv; P.L = L; P.clk = dclk;
  printf("class nv_cap %d (NVP %d, %d threads, %d CTAs/SM, smem %zu), nv = %d\n", nv_cap, npad, NT, MINB, smem, nv);
  std::vector<double> Hi[2], xs[2];
  for (int mma = 0; mma < 2; mma++) {
    P.n_prob = n_check; P.reps = 1;
    float ms;
    if (mma) ms = run(sweep_test_kernel<NT, GR, R, GC, C, NWS, NB, PK, MINB, true>, NT, n_check, P, smem);
    else ms = run(sweep_test_kernel<NT, GR, R, GC, C, NWS, NB, PK, MINB, false>, NT, n_check, P, smem);
    Hi[mma].resize((size_t)n_check * nv * nv);
    xs[mma].resize((size_t)n_check * nv);
    std::vector<int> st(n_check);
    cudaMemcpy(Hi[mma].data(), dHi, Hi[mma].size() * 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(xs[mma].data(), dx, xs[mma].size() * 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(st.data(), dst, n_check * 4, cudaMemcpyDeviceToHost);
    check(mma ? "mma" : "tiles", nv, n_check, H, g, Hi[mma], xs[mma], st);
    for (int occ = 1; occ <= MINB; occ *= 2) {
    P.n_prob = n_time / 4; P.reps = 17;
    const size_t smem_occ = occ == MINB ? smem : (size_t)(227 * 1024 / occ - 1024);  // padding caps the CTAs per SM
    if (mma) ms = run(sweep_test_kernel<NT, GR, R, GC, C, NWS, NB, PK, MINB, true>, NT, 148 * occ, P, smem_occ);
    else ms = run(sweep_test_kernel<NT, GR, R, GC, C, NWS, NB, PK, MINB, false>, NT, 148 * occ, P, smem_occ);
    {
      const double n_inv = (double)(n_time / 4) * 17;
      long long hc[40];
      cudaMemcpy(hc, dclk, sizeof(hc), cudaMemcpyDeviceToHost);
      const double per_cta = n_inv / (148.0 * occ);
      printf("  %-6s %.3f ms for %.0f inversions: %.0f SM-cycles per inversion at 1.965 GHz (throughput, %d CTAs/SM); "
             "latency per inversion in CTA 0: %.0f cycles", mma ? "mma" : "tiles", ms, n_inv,
             ms * 1e-3 * 1.965e9 / (n_inv / 148), occ, hc[3] / per_cta);
      if (mma)
        for (int w = 0; w < 4; w++) {
          const long long* c = hc + 8 * (1 + w);
          printf("\n      warp %d: gather %.0f (+wait %.0f), pivot block %.0f (+%.0f), panel %.0f (+%.0f), update %.0f", w,
                 c[0] / per_cta, c[1] / per_cta, c[2] / per_cta, c[3] / per_cta, c[4] / per_cta, c[5] / per_cta, c[6] / per_cta);
        }
      printf("\n");
    }
    }
  }
  double dmax = 0, hmax = 0;
  for (size_t e = 0; e < Hi[0].size(); e++) {
    dmax = fmax(dmax, fabs(Hi[0][e] - Hi[1][e]));
    hmax = fmax(hmax, fabs(Hi[0][e]));
  }
  printf("  |Hinv_mma - Hinv_tiles| max %.2e (|Hinv| max %.2e)\n", dmax, hmax);
  cudaFree(dH); cudaFree(dg); cudaFree(dHi); cudaFree(dx); cudaFree(dst);
}

int main() {
  //          NT  GR  R  GC  C NWS NB  PK   MINB
  test_class<128, 16, 4,  8, 8, 4,  8, false, 4>(60, 60, 10, 27);
  test_class<128, 16, 4,  8, 8, 4,  8, false, 4>(60, 36, 10, 27);
  test_class<256, 16, 6, 16, 6, 4, 12, false, 2>(96, 96, 16, 17);
  test_class<256, 16, 6, 16, 6, 4, 12, false, 2>(96, 75, 16, 17);
  test_class<256, 16, 8, 16, 8, 8, 16, true,  2>(120, 120, 10, 31);
  return 0;
}
