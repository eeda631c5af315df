// Batched convex-MPC engine for sm_100a: kernels + the C ABI of include/mpc_batch.h.
//
// Kernels
//   mpc_solve_riccati_kernel  the production kernel of the classes nv <= 60 / 96 (and <= 128 at horizons <= 12): ONE
//                        WARP per problem, the QP solved without its condensed Hessian (csrc/mpc_riccati.h: Riccati
//                        factorisation on the FP64 tensor pipe, H^{-1} products as sweeps over the horizon).
//   mpc_solve_pipe_kernel / mpc_solve_kernel<NT>  the explicit-inverse solver (csrc/mpc_core.h), one problem per CTA
//                        (two in flight in the piped form): long horizons, the single-robot tick, the warm start.
//   mpc_solve_wrench_kernel   nv > 128 through the rank-6h structure of the Hessian.
//   mpc_classify_kernel  one thread per problem: counts stance (step,leg) pairs in the gait table and appends the
//                        problem to its size class (nv = 3 * stance); the host entries classify while they stage a
//                        batch and launch ONE kernel when it is uniform.
//   mpc_build_records_kernel / mpc_gait_state_kernel / mpc_leg_commands_kernel  the callers either side of the solve
//                        (SURVEY 8f N1, N2, N4), one robot per thread.
// All solve kernels are persistent: records are staged global -> shared by TMA bulk copies (cp.async.bulk + mbarrier),
// problems are handed out by a dynamic queue, everything between the record (4*(48+12h)+4h bytes read) and the 12 fp32
// forces (+ optional 12h fp64, status) written lives in shared memory / registers; H and g never touch HBM.
// There is no CPU solver in this library: every entry point either runs these kernels or fails with an error code.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <type_traits>
#include <vector>

#include "mpc_core.h"
#include "mpc_ticks.h"
#include "mpc_legs.h"
#include "mpc_riccati.h"

namespace {

constexpr int kMaxPeers = 8;
constexpr int kSlots = MPC_BATCH_SLOTS;  // scratch slots per engine: that many batches can be in flight
constexpr int kMaxClasses = 6;
constexpr int kRing = 256;
constexpr int kHostClassifyMax = 8192;  // largest batch the host entries classify themselves (see host_classify_records)

#ifndef MPC_SWEEP_DEFAULT
#define MPC_SWEEP_DEFAULT 0
#endif
#ifndef MPC_SOLVER_DEFAULT
#define MPC_SOLVER_DEFAULT 1
#endif

struct SolveParams {
  const char* records;
  unsigned long long stride;
  int h, batch;
  const int* list;    // problem ids of this class (nullptr: identity)
  const int* count;   // number of ids (nullptr: batch)
  float* forces;
  double* solution;
  int32_t* status;
  int* retry_list;    // where to queue a problem whose working set outgrew this class
  int* retry_count;
  char* slab;
  mpc::Layout L;
  mpc::RicLayout RL;  // workspace of the Riccati solver (mpc_solve_riccati_kernel)
  char* ric_slab;     // [grid][RL.slab_bytes] global scratch for working sets that outgrow the Riccati tile
  int* queue;         // {next, done}: dynamic problem queue of the persistent solve kernels (nullptr: static stride)
  int ric_generic;    // development switch (env MPC_RIC_GENERIC): the scalar generic factorisation instead of the DMMA one
  int max_iter;
  int warp_mode;
  float* peers[kMaxPeers];
  int n_peers, rank_offset;
  long long* phase_clk;  // optional [batch][24] SM-clock stamps at phase boundaries (profiling aid)
  int debug_stop;        // profiling aid (env MPC_DEBUG_STOP): 1 stop after assembly, 2 after the inversion; forces are NOT valid
  int* warm_cache;       // [robots][kWarmStride] working-set cache of the warm start (nullptr: cold start)
  const int* warm_ids;   // robot id of every problem (nullptr: problem index)
  int warm_shift;        // horizon steps the gait has advanced since the cached solve
  int32_t* nvar_out;  // assemble-only mode when H_out != nullptr
  double* H_out;
  double* g_out;
};

// ---- TMA bulk copy + mbarrier (PTX) -------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(phase)
      : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// One thread per problem: counts the stance (step,leg) pairs of its gait table (16-byte loads: the table starts
// 16-byte aligned inside the record), picks the size class and appends the problem to that class's list (one
// atomic per warp and class).  Also zeroes the OTHER parity's counters for the next solve, so no memset sits
// between solves.
__global__ void mpc_classify_kernel(const char* records, unsigned long long stride, int h, int batch, int n_classes,
                                    const int* __restrict__ class_cap, int* lists, int* counts, int* counts_next,
                                    int max_batch) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < kMaxClasses) counts_next[b] = 0;
  int c = -1;  // -1: no problem behind this thread
  if (b < batch) {
    const float* rec = (const float*)(records + stride * b);
    const float fmax = rec[MPC_REC_FMAX];
    const uint4* g4 = (const uint4*)((const char*)rec + 4 * (MPC_REC_TRAJ + 12 * h));
    const int nbytes = 4 * h;
    int ns = 0;
    for (int q = 0; q * 16 < nbytes; q++) {
      const uint4 v = g4[q];  // bytes past 4h are the record's zero padding (stride is rounded up to 16)
      const unsigned w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int e = 0; e < 4; e++) {
          if (q * 16 + i * 4 + e < nbytes) {
            const float ub = (float)((w[i] >> (8 * e)) & 0xffu) * fmax;
            ns += !((double)ub < 0.01 && (double)ub > -0.01);
          }
        }
    }
    const int nv = 3 * ns;
    c = 0;
    while (c < n_classes - 1 && nv > class_cap[c]) c++;
  }
  // one atomic per warp and class instead of one per problem (4096 atomics on one counter were most of this kernel)
  const unsigned lane = threadIdx.x & 31u;
  for (int cc = 0; cc < n_classes; cc++) {
    const unsigned m = __ballot_sync(0xffffffffu, c == cc);
    if (m == 0) continue;  // uniform
    int base = 0;
    const int leader = __ffs(m) - 1;
    if ((int)lane == leader) base = atomicAdd(&counts[cc], __popc(m));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (c == cc) lists[cc * max_batch + base + __popc(m & ((1u << lane) - 1u))] = b;
  }
}

// Tick records -> problem records, one robot per thread (SURVEY 8f N1 + N2; body in mpc_ticks.h).
__global__ void mpc_build_records_kernel(const float* ticks, int batch, int h, char* records,
                                         unsigned long long stride, float* state_out) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= batch) return;
  mpc::build_record_from_tick(ticks + (size_t)b * MPC_TICK_WORDS, h, records + stride * b, (size_t)stride,
                              state_out ? state_out + 4 * (size_t)b : nullptr);
}

// SURVEY 8f rows N2 / N4, one robot per thread (bodies in mpc_legs.h).
__global__ void mpc_gait_state_kernel(const int32_t* gait, int batch, float* state_out, unsigned char* table_out,
                                      int table_stride) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= batch) return;
  mpc::gait_state_from_record(gait + (size_t)b * MPC_GAIT_WORDS, state_out + (size_t)b * MPC_GAIT_STATE_WORDS,
                              table_out ? table_out + (size_t)b * table_stride : nullptr);
}
__global__ void mpc_leg_commands_kernel(const float* legs, const float* forces, int batch, float* f_ff, float* tau) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= batch) return;
  mpc::leg_commands_from_record(legs + (size_t)b * MPC_LEG_WORDS, forces + (size_t)12 * b, f_ff + (size_t)12 * b,
                                tau + (size_t)12 * b);
}

// Device-side barrier of the fused gather: after this rank's solve kernels have completed (stream order), tell every
// peer "my rows of epoch E have landed" through its flag word, then wait until every peer has said the same here
// (each lane signals before it waits, so the ranks cannot dead-lock each other).
__global__ void mpc_gather_barrier_kernel(unsigned* const* peer_flags, const unsigned* my_flags, int world, int rank,
                                          unsigned epoch) {
  const int q = threadIdx.x;
  if (q < world) {
    if (peer_flags[q]) {
      __threadfence_system();
      asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(peer_flags[q] + rank), "r"(epoch) : "memory");
    }
    unsigned v;
    do {
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(my_flags + q) : "memory");
    } while ((int)(v - epoch) < 0);
  }
}

// Shard-and-gather epilogue: one problem's 12 forces go straight into every rank's gather buffer over NVLink -- three
// 16-byte stores per peer (lanes 0..2) when the rows are 16-byte aligned, twelve 4-byte ones otherwise.  Called by the
// threads that just wrote forces[12*b ..] (lane i wrote element i, so a lane reads back what lanes of its own warp
// stored: the caller has synchronised them).
__device__ __forceinline__ void peer_store_forces(const SolveParams& P, int b, int tid) {
  const float* src = P.forces + (size_t)12 * b;
  const size_t row = (size_t)12 * (P.rank_offset + b);
  bool wide = (((uintptr_t)P.forces) & 15) == 0;
#pragma unroll
  for (int q = 0; q < kMaxPeers; q++)
    if (q < P.n_peers && P.peers[q]) wide = wide && ((((uintptr_t)P.peers[q]) & 15) == 0);
  if (wide) {
    if (tid < 3) {
      const float4 f = *reinterpret_cast<const float4*>(src + 4 * tid);
#pragma unroll
      for (int q = 0; q < kMaxPeers; q++)
        if (q < P.n_peers && P.peers[q]) *reinterpret_cast<float4*>(P.peers[q] + row + 4 * tid) = f;
    }
  } else if (tid < 12) {
    const float f = src[tid];
#pragma unroll
    for (int q = 0; q < kMaxPeers; q++)
      if (q < P.n_peers && P.peers[q]) P.peers[q][row + tid] = f;
  }
}

// Dynamic problem queue of the persistent solve kernels.  Problems of one class differ widely in work (0 ... 70
// working-set changes), so after its first problem (its block index, no atomic) a CTA takes the next one from an atomic
// counter instead of a fixed stride; the ticket is taken by thread 0 when it prefetches the next record and handed to
// the CTA through shared memory.  The last CTA to leave resets the counters for the next launch on this slot and class
// (those launches are stream-ordered).
__device__ __forceinline__ int queue_take(const SolveParams& P, int item) {
  return P.queue ? (int)gridDim.x + atomicAdd(P.queue, 1) : item + (int)gridDim.x;
}
__device__ __forceinline__ void queue_leave(const SolveParams& P, int count) {
  if (!P.queue) return;
  const int target = min((int)gridDim.x, count);  // the CTAs that took part (the others left at the top)
  __threadfence();
  if (atomicAdd(P.queue + 1, 1) == target - 1) {
    P.queue[0] = 0;
    P.queue[1] = 0;
    __threadfence();
  }
}

// NT threads per CTA.  R > 0: register-resident inversion with R x C tiles on a GR x GC thread grid (NT == GR*GC,
// padded size GR*R == GC*C).  R == 0: the generic shared/global-memory sweep, used by the catch-all class whose
// matrix does not fit in the register file of one SM.
// PROF: the instantiation that serves the profiling / debugging entries (phase clocks, assemble-only output, stage
// stop).  The production instantiation carries none of that code (2% faster: the kernel is instruction-fetch heavy).
// SW: the register-resident inversion -- 0: invert_spd_tiles (rank-1 updates on the FP64 FMA pipe), 1: invert_spd_mma
// (grouped sweep, rank-8 updates as DMMA.8x8x4 on the FP64 tensor pipe; NWS sweeping warps, NB = NVP / 8 block rows).
template <int NT, int GR, int R, int GC, int C, int MINB, bool PK, bool PROF, int SW = 0>
__global__ void __launch_bounds__(NT, MINB) mpc_solve_kernel(const __grid_constant__ SolveParams P) {
  extern __shared__ __align__(128) char smem[];
  const int count = P.count ? *P.count : P.batch;
  if ((int)blockIdx.x >= count) return;
  uint64_t* bar = (uint64_t*)smem;
  char* recbuf = smem + 16;
  char* fast = recbuf + 2 * P.stride;
  const mpc::CtaT<PK, (R == 0 ? 4 : 1)> cx{(int)threadIdx.x, NT};  // PK: H / H^{-1} as a packed lower triangle (see build_classes)
  const mpc::Work k = mpc::carve(P.L, fast, P.slab ? P.slab + (size_t)blockIdx.x * P.L.slab_bytes : nullptr);
  if (threadIdx.x == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const uint32_t rec_bytes = (uint32_t)P.stride;
  int item = blockIdx.x;
  if (threadIdx.x == 0) {
    const int b0 = P.list ? P.list[item] : item;
    mbar_expect_tx(&bar[0], rec_bytes);
    tma_bulk_g2s(recbuf, P.records + P.stride * b0, rec_bytes, &bar[0]);
  }
  int& s_next = k.sc->next_item;  // (dynamic shared memory: a static __shared__ variable would lower the opt-in limit)
  int nxt = 0;
  for (int it = 0; item < count; item = nxt, it++) {
    const int cur = it & 1;
    nxt = item + gridDim.x;  // static stride (the profiling entries, whose early `continue`s skip the hand-over below)
    if (threadIdx.x == 0) {  // buffer cur^1 was released by the barrier that ended the last pass
      const int next = PROF ? nxt : queue_take(P, item);
      s_next = next;
      if (next < count) {
        const int bn = P.list ? P.list[next] : next;
        mbar_expect_tx(&bar[cur ^ 1], rec_bytes);
        tma_bulk_g2s(recbuf + (size_t)(cur ^ 1) * P.stride, P.records + P.stride * bn, rec_bytes, &bar[cur ^ 1]);
      }
    }
    mbar_wait(&bar[cur], (uint32_t)((it >> 1) & 1));
    const int b = P.list ? P.list[item] : item;
    const float* rec = (const float*)(recbuf + (size_t)cur * P.stride);
    const unsigned char* gait = (const unsigned char*)rec + 4 * (MPC_REC_TRAJ + 12 * P.h);

    long long* clk = nullptr;
    if constexpr (PROF) clk = P.phase_clk ? P.phase_clk + (size_t)24 * b : nullptr;
    if constexpr (PROF) const_cast<mpc::Work&>(k).clk = clk;
    if (clk && threadIdx.x == 0) clk[0] = clock64();
    mpc::assemble(cx, rec, gait, k);
    if (clk && threadIdx.x == 0) clk[1] = clock64();
    if constexpr (PROF) {
    if (P.H_out) {  // debug / parity entry: write the reduced QP out and stop
      const int nv = (k.sc->status == MPC_STATUS_OPTIMAL) ? k.sc->nv : 0;
      const int NU = 12 * P.h;
      if (threadIdx.x == 0 && P.nvar_out) P.nvar_out[b] = nv;
      double* Ho = P.H_out + (size_t)b * NU * NU;
      for (int e = threadIdx.x; e < nv * nv; e += NT) {
        const int i = e / nv, j = e - i * nv;
        Ho[(size_t)i * NU + j] = k.Hm[mpc::hixT<PK>(k.ld, i, j)];
      }
      if (P.g_out)
        for (int i = threadIdx.x; i < nv; i += NT) P.g_out[(size_t)b * NU + i] = k.g[i];
      __syncthreads();
      continue;
    }
    if (P.debug_stop == 1) { __syncthreads(); continue; }
    }
    if (k.sc->status == MPC_STATUS_OPTIMAL) {
      if constexpr (R > 0) {
        static_assert(R == 0 || NT == GR * GC, "thread grid");
        if constexpr (SW == 1) mpc::invert_spd_mma<NT, (GR * R == 128 ? 8 : 4), GR * R / 8, PK>(k, (int)threadIdx.x, true);
        else mpc::invert_spd_tiles<GR, R, GC, C, PK>(k, (int)threadIdx.x, true);
      } else {
        mpc::invert_spd(cx, k);
      }
    }
    if (clk && threadIdx.x == 0) clk[2] = clock64();
    if constexpr (PROF) {
      if (P.debug_stop == 2) { __syncthreads(); continue; }
    }
    if (k.sc->status == MPC_STATUS_OPTIMAL) {
      // register-resident classes: x = -H^{-1} g came out of the sweep (the gradient rode along as row nv)
      mpc::active_set_init(cx, rec, gait, k, R > 0 && k.sc->nv < GR * R);
      int* const wc = P.warm_cache ? P.warm_cache + (size_t)(P.warm_ids ? P.warm_ids[b] : b) * mpc::kWarmStride : nullptr;
      if constexpr (R > 0) {
        // shared-memory classes: the active-set loop is a chain of tiny steps, so one warp runs it with
        // __syncwarp / shuffles instead of CTA barriers; the other warps wait at the barrier below
        if (P.warp_mode) {
#ifdef MPC_ROTATE_GI  // measured: rotating the warp over CTAs / problems is 1% slower than always using warp 0
          const int aw = (int)(blockIdx.x + it) & (NT / 32 - 1);
#else
          const int aw = 0;
#endif
          if ((int)(threadIdx.x >> 5) == aw) {
            const mpc::WarpT<PK> wx{(int)(threadIdx.x & 31), 32};
            if (wc) mpc::active_set_warm(wx, rec, k, wc, P.warm_shift);
            mpc::active_set(wx, rec, gait, k, P.max_iter);
          }
        } else {
          if (wc) mpc::active_set_warm(cx, rec, k, wc, P.warm_shift);
          mpc::active_set(cx, rec, gait, k, P.max_iter);
        }
        __syncthreads();
      } else {
        if (wc) mpc::active_set_warm(cx, rec, k, wc, P.warm_shift);
        mpc::active_set(cx, rec, gait, k, P.max_iter);
      }
      if (wc && k.sc->status != mpc::STATUS_RETRY_BIG) mpc::active_set_store(cx, k, wc);
    }
    if (clk && threadIdx.x == 0) clk[3] = clock64();
    const int code = k.sc->status;
    if (code == mpc::STATUS_RETRY_BIG && P.retry_list) {
      if (threadIdx.x == 0) {
        const int slot = atomicAdd(P.retry_count, 1);
        P.retry_list[slot] = b;
      }
    } else {
      if (code == mpc::STATUS_RETRY_BIG && threadIdx.x == 0) k.sc->status = MPC_STATUS_MAX_ITER;
      __syncthreads();
      mpc::scatter(cx, k, P.forces + (size_t)12 * b, P.solution ? P.solution + (size_t)12 * P.h * b : nullptr,
                   P.status ? P.status + b : nullptr);
      if (P.n_peers > 0) {
        __syncwarp();  // lanes 0..11 wrote the forces, lanes 0..2 read them back four at a time
        peer_store_forces(P, b, (int)threadIdx.x);
      }
    }
    if constexpr (!PROF) nxt = s_next;  // (written at the top of this pass, CTA barriers in between; read before the
    __syncthreads();                    //  barrier that lets thread 0 overwrite it)
    if (clk && threadIdx.x == 0) {
      unsigned wid, sid;
      asm volatile("mov.u32 %0, %%warpid;" : "=r"(wid));
      asm volatile("mov.u32 %0, %%smid;" : "=r"(sid));
      clk[4] = clock64(); clk[5] = blockIdx.x; clk[6] = it; clk[7] = (long long)(sid << 8 | wid);
    }
  }
  if constexpr (!PROF) if (threadIdx.x == 0) queue_leave(P, count);
}

// ---- two problems in flight per CTA (the smallest class, production instantiation) --------------------------------
// In mpc_solve_kernel the active set of a problem runs on one warp while the CTA's other warps wait at a barrier (18 %
// of all warp samples of the round-1 capture).  Here the CTA works in rounds: in phase X warp 0 runs the active set and
// the scatter of problem n-1 while warps 1..3 assemble problem n up to the gradient and the M tables (assemble_front,
// on its own hardware barrier); in phase Y all four warps write the H blocks, invert them in registers and set the
// active set up.  The two roles share no shared memory (piped layout, mpc_core.h: disjoint assembly / active-set
// scratch, moment sums outside the H^{-1} tile, scalars and stance lists twice), the record of problem n-1 stays in
// its buffer until its active set is done, and the arithmetic per problem is exactly that of mpc_solve_kernel.
// NG: threads of the active-set role (32: one warp with __syncwarp / shuffles; more: a group on hardware barrier 2);
// the other NT - NG threads assemble on hardware barrier 1.
template <int NT, int GR, int R, int GC, int C, int MINB, bool PK, int NG, int SW = 0>
__global__ void __launch_bounds__(NT, MINB) mpc_solve_pipe_kernel(const __grid_constant__ SolveParams P) {
  extern __shared__ __align__(128) char smem[];
  const int count = P.count ? *P.count : P.batch;
  if ((int)blockIdx.x >= count) return;
  uint64_t* bar = (uint64_t*)smem;
  char* recbuf = smem + 16;
  char* fast = recbuf + 2 * P.stride;
  const int tid = (int)threadIdx.x;
  const mpc::CtaT<PK> cx{tid, NT};
  const mpc::PartT<1, NT - NG> px{tid - NG, NT - NG};
  // the active-set role: one warp, or a group of NG threads with its own barrier
  using GCtx = typename std::conditional<NG == 32, mpc::WarpT<PK>, mpc::PartT<2, NG, PK>>::type;
  const GCtx wx{tid, NG};
  auto gsync = [&]() { if (NG == 32) __syncwarp(); else wx.sync(); };
  if (tid == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const uint32_t rec_bytes = (uint32_t)P.stride;
  int item = blockIdx.x;
  if (tid == 0) {
    const int b0 = P.list ? P.list[item] : item;
    mbar_expect_tx(&bar[0], rec_bytes);
    tma_bulk_g2s(recbuf, P.records + P.stride * b0, rec_bytes, &bar[0]);
  }
  int& s_next = mpc::carve(P.L, fast, nullptr, 0).sc->next_item;
  int prev_b = -1;  // problem whose active set is still to run (its state sits in set (it-1)&1, its record in buffer (it-1)&1)
  for (int it = 0;; it++) {
    const int cur = it & 1;
    const bool have = item < count;
    const float* rec = (const float*)(recbuf + (size_t)cur * P.stride);
    const unsigned char* gait = (const unsigned char*)rec + 4 * (MPC_REC_TRAJ + 12 * P.h);
    if (have) mbar_wait(&bar[cur], (uint32_t)((it >> 1) & 1));
    // ---- phase X: active set + scatter of the previous problem (warp 0) || front half of this problem's assembly ----
    if (tid < NG) {
      if (prev_b >= 0) {
        const mpc::Work kg = mpc::carve(P.L, fast, nullptr, cur ^ 1);
        const float* recp = (const float*)(recbuf + (size_t)(cur ^ 1) * P.stride);
        const unsigned char* gaitp = (const unsigned char*)recp + 4 * (MPC_REC_TRAJ + 12 * P.h);
        int* const wc =
            P.warm_cache ? P.warm_cache + (size_t)(P.warm_ids ? P.warm_ids[prev_b] : prev_b) * mpc::kWarmStride : nullptr;
        if (wc) mpc::active_set_warm(wx, recp, kg, wc, P.warm_shift);
        mpc::active_set(wx, recp, gaitp, kg, P.max_iter);
        gsync();
        const int code = kg.sc->status;
        if (wc && code != mpc::STATUS_RETRY_BIG) mpc::active_set_store(wx, kg, wc);
        if (code == mpc::STATUS_RETRY_BIG && P.retry_list) {
          if (tid == 0) {
            const int slot = atomicAdd(P.retry_count, 1);
            P.retry_list[slot] = prev_b;
          }
        } else {
          gsync();
          if (code == mpc::STATUS_RETRY_BIG && tid == 0) kg.sc->status = MPC_STATUS_MAX_ITER;
          gsync();
          mpc::scatter(wx, kg, P.forces + (size_t)12 * prev_b,
                       P.solution ? P.solution + (size_t)12 * P.h * prev_b : nullptr, P.status ? P.status + prev_b : nullptr);
          if (P.n_peers > 0) {
            __syncwarp();
            peer_store_forces(P, prev_b, tid);
          }
        }
      }
    } else if (have) {
      const mpc::Work ka = mpc::carve(P.L, fast, nullptr, cur);
      mpc::assemble_front(px, rec, gait, ka);
    }
    __syncthreads();
    if (!have) break;
    const int b = P.list ? P.list[item] : item;
    if (tid == 0) {  // buffer cur^1 is free now: the previous problem's active set is done with it
      const int next = queue_take(P, item);
      s_next = next;
      if (next < count) {
        const int bn = P.list ? P.list[next] : next;
        mbar_expect_tx(&bar[cur ^ 1], rec_bytes);
        tma_bulk_g2s(recbuf + (size_t)(cur ^ 1) * P.stride, P.records + P.stride * bn, rec_bytes, &bar[cur ^ 1]);
      }
    }
    // ---- phase Y: H blocks, inversion, active-set set-up (all warps) ----
    const mpc::Work k = mpc::carve(P.L, fast, nullptr, cur);
    if (k.sc->status == MPC_STATUS_OPTIMAL) mpc::assemble_H(cx, rec, k);
    if (k.sc->status == MPC_STATUS_OPTIMAL) {
      if constexpr (SW == 1) mpc::invert_spd_mma<NT, (GR * R == 128 ? 8 : 4), GR * R / 8, PK>(k, tid, true);
      else mpc::invert_spd_tiles<GR, R, GC, C, PK>(k, tid, true);
    }
    if (k.sc->status == MPC_STATUS_OPTIMAL) {
      mpc::active_set_init(cx, rec, gait, k, k.sc->nv < GR * R);  // x = -H^{-1} g came out of the sweep if it had room
      prev_b = b;
    } else {  // bad input / no stance leg / not positive definite: report now, nothing to iterate on
      __syncthreads();
      mpc::scatter(cx, k, P.forces + (size_t)12 * b, P.solution ? P.solution + (size_t)12 * P.h * b : nullptr,
                   P.status ? P.status + b : nullptr);
      if (P.n_peers > 0) {
        __syncwarp();
        peer_store_forces(P, b, tid);
      }
      prev_b = -1;
      __syncthreads();
    }
    item = s_next;  // (handed over through the CTA barriers of phase Y; overwritten only after the next phase-X barrier)
  }
  if (tid == 0) queue_leave(P, count);
}

// ---- wrench-space class: problems with more reduced variables than the register-resident classes hold (nv > 128)
// at horizons with 6h <= 128.  H = 2 (alpha I + G'KG) with K of size 6h (mpc_core.h, "wrench-space class"): two
// register-resident inversions of size 6h instead of one of size nv in an L2 slab, and the active set applies
// H^{-1} = (I - G'MG) / (2 alpha) from shared memory.  One problem per CTA at a time, 256 threads.
// SW as above: which register-resident inversion runs.
// UNR: unroll factor of the active set's loops over the working set (4 for the overflow class, whose T lives in L2).
template <int MINB, bool PROF, int SW, int UNR = 1>
__global__ void __launch_bounds__(256, MINB) mpc_solve_wrench_kernel(const __grid_constant__ SolveParams P) {
  extern __shared__ __align__(128) char smem[];
  const int count = P.count ? *P.count : P.batch;
  if ((int)blockIdx.x >= count) return;
  uint64_t* bar = (uint64_t*)smem;
  char* recbuf = smem + 16;
  char* fast = recbuf + 2 * P.stride;
  const int tid = (int)threadIdx.x;
  const mpc::CtaT<true, UNR, true> cx{tid, 256};
  mpc::Work k = mpc::carve(P.L, fast, P.slab ? P.slab + (size_t)blockIdx.x * P.L.slab_bytes : nullptr);
  if (tid == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const uint32_t rec_bytes = (uint32_t)P.stride;
  int item = blockIdx.x;
  if (tid == 0) {
    const int b0 = P.list ? P.list[item] : item;
    mbar_expect_tx(&bar[0], rec_bytes);
    tma_bulk_g2s(recbuf, P.records + P.stride * b0, rec_bytes, &bar[0]);
  }
  const int n = 6 * P.h;
  int& s_next = k.sc->next_item;
  int nxt = 0;
  for (int it = 0; item < count; item = nxt, it++) {
    const int cur = it & 1;
    nxt = item + gridDim.x;  // static stride (the profiling instantiation)
    if (tid == 0) {
      const int next = PROF ? nxt : queue_take(P, item);
      s_next = next;
      if (next < count) {
        const int bn = P.list ? P.list[next] : next;
        mbar_expect_tx(&bar[cur ^ 1], rec_bytes);
        tma_bulk_g2s(recbuf + (size_t)(cur ^ 1) * P.stride, P.records + P.stride * bn, rec_bytes, &bar[cur ^ 1]);
      }
    }
    mbar_wait(&bar[cur], (uint32_t)((it >> 1) & 1));
    const int b = P.list ? P.list[item] : item;
    const float* rec = (const float*)(recbuf + (size_t)cur * P.stride);
    const unsigned char* gait = (const unsigned char*)rec + 4 * (MPC_REC_TRAJ + 12 * P.h);
    k.i2a = 0.5 / (double)rec[MPC_REC_ALPHA];
    mpc::assemble_front(cx, rec, gait, k);
    auto invert = [&]() {
      if constexpr (SW == 1) mpc::invert_spd_mma<256, 8, 16, true>(k, tid, false, n);
      else mpc::invert_spd_tiles<16, 8, 16, 8, true>(k, tid, false, n);
    };
    if (k.sc->status == MPC_STATUS_OPTIMAL) {
      mpc::assemble_K(cx, rec, k);
      invert();
    }
    if (k.sc->status == MPC_STATUS_OPTIMAL) {
      mpc::wr_form_second(cx, rec, k);
      invert();
    }
    if constexpr (PROF) {
      if (P.debug_stop != 0) { __syncthreads(); continue; }
    }
    if (k.sc->status == MPC_STATUS_OPTIMAL) {
      mpc::active_set_init(cx, rec, gait, k);
      mpc::active_set(cx, rec, gait, k, P.max_iter);
    }
    const int code = k.sc->status;
    if (code == mpc::STATUS_RETRY_BIG && P.retry_list) {
      if (tid == 0) {
        const int slot = atomicAdd(P.retry_count, 1);
        P.retry_list[slot] = b;
      }
    } else {
      if (code == mpc::STATUS_RETRY_BIG && tid == 0) k.sc->status = MPC_STATUS_MAX_ITER;
      __syncthreads();
      mpc::scatter(cx, k, P.forces + (size_t)12 * b, P.solution ? P.solution + (size_t)12 * P.h * b : nullptr,
                   P.status ? P.status + b : nullptr);
      if (P.n_peers > 0) {
        __syncwarp();
        peer_store_forces(P, b, tid);
      }
    }
    if constexpr (!PROF) nxt = s_next;
    __syncthreads();
  }
  if constexpr (!PROF) if (tid == 0) queue_leave(P, count);
}

// ---- Riccati solver (csrc/mpc_riccati.h): ONE WARP PER PROBLEM ------------------------------------------------
// No condensed Hessian and no inversion: the gains of the horizon's Riccati recursion are factored once and every
// H^{-1} product the dual active-set method asks for is a backward + forward sweep over the horizon.  A CTA is one
// warp (__syncwarp only); the persistent grid keeps as many warps per SM as the per-problem workspace allows.  Records
// are staged by TMA bulk copies, double buffered, exactly as in mpc_solve_kernel.
template <bool GENERIC>
__global__ void __launch_bounds__(32, 12) mpc_solve_riccati_kernel(const __grid_constant__ SolveParams P) {
  extern __shared__ __align__(128) char smem[];
  const int count = P.count ? *P.count : P.batch;
  if ((int)blockIdx.x >= count) return;
  uint64_t* bar = (uint64_t*)smem;
  char* recbuf = smem + 16;
  char* fast = recbuf + P.stride;  // ONE record buffer: the co-resident warps hide the copy, shared memory buys residency
  const int lane = (int)threadIdx.x;
  const mpc::WarpT<false> cx{lane, 32};
  mpc::RicWork k = mpc::ric_carve(P.RL, fast);
  k.slab = P.ric_slab ? P.ric_slab + (size_t)blockIdx.x * P.RL.slab_bytes : nullptr;
  if (lane == 0) {
    mbar_init(&bar[0], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  const uint32_t rec_bytes = (uint32_t)P.stride;
  const float* rec = (const float*)recbuf;
  const unsigned char* gait = (const unsigned char*)rec + 4 * (MPC_REC_TRAJ + 12 * P.h);
  // Problems differ by a factor of five in work (0 ... 11 working-set changes on a trot batch), and a 4096-batch is
  // only 2.5 problems per resident warp: after its first problem (its block index, no atomic) a warp takes the next
  // one from a queue instead of a fixed stride.  The last warp to leave resets the queue for the next launch on this
  // slot (launches of one slot and class are stream-ordered).
  int it = 0;
  for (int item = blockIdx.x; item < count; it++) {
    const int b = P.list ? P.list[item] : item;
    if (lane == 0) {  // the buffer was released by the __syncwarp that ended the last pass
      mbar_expect_tx(&bar[0], rec_bytes);
      tma_bulk_g2s(recbuf, P.records + P.stride * b, rec_bytes, &bar[0]);
    }
    mbar_wait(&bar[0], (uint32_t)(it & 1));
    const int code = mpc::ric_solve_problem<GENERIC>(cx, rec, gait, k, P.max_iter);
    __syncwarp();
    if (code == mpc::STATUS_RETRY_BIG && P.retry_list) {
      if (lane == 0) {
        const int slot = atomicAdd(P.retry_count, 1);
        P.retry_list[slot] = b;
      }
    } else {
      if (code == mpc::STATUS_RETRY_BIG && lane == 0) k.sc->status = MPC_STATUS_MAX_ITER;
      __syncwarp();
      mpc::ric_scatter(cx, k, P.forces + (size_t)12 * b, P.solution ? P.solution + (size_t)12 * P.h * b : nullptr,
                       P.status ? P.status + b : nullptr);
      if (P.n_peers > 0) {
        __syncwarp();  // lanes 0..11 wrote the forces, lanes 0..2 read them back four at a time
        peer_store_forces(P, b, lane);
      }
    }
    __syncwarp();
    int nxt = 0;
    if (lane == 0) nxt = queue_take(P, item);
    item = __shfl_sync(0xffffffffu, nxt, 0);
  }
  if (lane == 0) queue_leave(P, count);
}

struct ClassCfg {
  int nv_cap, m_cap, in_fast, threads, grid, variant;
  size_t smem;
  mpc::Layout L;
  // two problems in flight per CTA (mpc_solve_pipe_kernel; the smallest class when its piped layout fits four CTAs per
  // SM with a useful working-set tile): the production launches use these, the profiling entries the ones above
  bool pipe = false;
  int pipe_m_cap = 0, pipe_grid = 0;
  size_t pipe_smem = 0;
  mpc::Layout pipe_L;
  // Riccati solver (mpc_solve_riccati_kernel, one warp per problem): used instead of the class's inverse-based kernel
  // when the engine's solver is 1, except by the profiling / assemble-only / warm-start entries
  bool ric = false;
  int ric_m_cap = 0, ric_grid = 0;
  size_t ric_smem = 0;
  mpc::RicLayout ric_L;
};

thread_local std::string g_err;

}  // namespace

struct mpc_batch {
  int device = 0, h = 0, max_batch = 0, sms = 0;
  size_t stride = 0;
  // Two independent slots (stream + staging + device scratch) so that the host entry can be pipelined:
  // slot k's H2D / kernels / D2H overlap the host's packing of slot 1-k.  Slot 0 serves the synchronous calls.
  struct Slot {
    cudaStream_t stream = nullptr;
    char* rec_dev = nullptr;
    char* out_dev = nullptr;   // forces_dev | status_dev | sol_dev (one allocation)
    char* out_pin = nullptr;   // the same layout in page-locked host memory
    size_t out_bytes = 0;
    float* forces_dev = nullptr;
    double* sol_dev = nullptr;
    int32_t* status_dev = nullptr;
    char* rec_pin = nullptr;
    float* forces_pin = nullptr;
    double* sol_pin = nullptr;
    int32_t* status_pin = nullptr;
    int* lists = nullptr;   // [classes][max_batch] problem ids per size class
    int* counts = nullptr;  // [2][classes], double-buffered by solve parity
    int parity = 0;
    char* slab = nullptr;   // per-CTA global workspace of the catch-all class
    char* ric_slab = nullptr;  // per-warp global workspace of the Riccati classes (working sets beyond the tile)
    int* ric_queue = nullptr;  // [kMaxClasses][2] {next, done} of the Riccati kernel's dynamic problem queue
    int pending_batch = 0;
    bool pending_solution = false;
    int pending_single_class = -1;  // >= 0: the pending solve was host-classified (every problem in that class)
    // tick entry of the host path (allocated on first use): tick records in, controller state out
    char* tick_dev = nullptr;
    char* tick_pin = nullptr;
    float* state_dev = nullptr;
    float* state_pin = nullptr;
    bool pending_state = false;
  } s[kSlots];
  int* caps_dev = nullptr;
  std::vector<ClassCfg> classes;
  ClassCfg dense_big;        // the dense catch-all configuration (last class, unless wrench-space classes replace it)
  bool has_wrench = false;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  // timing ring: event pairs around every class kernel of the last kRing solves (no sync while recording)
  std::vector<cudaEvent_t> ring0, ring1;
  long ring_pos = 0, ring_mark = 0;
  bool timed = false;
  int timed_class = -1;  // -1: events around every class kernel; k: around class k's kernel only
  long launches = 0;
  int max_iter = 4000;
  float* peers[kMaxPeers] = {nullptr};
  int n_peers = 0, rank_offset = 0;
  // fused gather: one region per scratch slot, each [gather_rows*12] fp32 forces followed by kMaxPeers flag words
  // (one per rank); slot q's region starts gather_slot_bytes * q into every rank's buffer
  float* gather_buf = nullptr;
  size_t gather_slot_bytes = 0;
  int gather_rows = 0;
  int gather_world = 0, gather_rank = 0;
  bool gather_fused = true;  // the solve kernels' peer-store epilogue is armed (false: gather by mpc_batch_gather_push_slot)
  unsigned gather_epoch[kSlots] = {0};
  unsigned** peer_flags_dev = nullptr;  // [kSlots][kMaxPeers]
  long long* phase_clk = nullptr;
  int ctas_per_sm_limit = 0;
  int* warm_cache = nullptr;        // warm start: device cache, robot ids, gait shift
  const int* warm_ids = nullptr;
  int warm_shift = 1;
  int sweep = MPC_SWEEP_DEFAULT;  // inversion of the register-resident classes: 0 FMA tiles, 1 DMMA grouped sweep
  char* cur_ric_slab = nullptr;  // the Riccati slab of the slot whose solve is being queued (set by solve_on_stream)
  int* cur_ric_queue = nullptr;  // ... and its queue counters; cur_class: the class being launched
  int cur_class = 0;
  bool ric_dynamic = true;       // env MPC_RIC_STATIC=1: static stride instead of the queue
  int solver = MPC_SOLVER_DEFAULT;  // 0: explicit inverse of the condensed Hessian; 1: Riccati sweeps (mpc_riccati.h)
  int debug_stop = 0;
  int ric_generic = 0;
  bool ric_always = false;  // env MPC_RIC_ALWAYS: small batches too (tests of the kernel under compute-sanitizer)
  bool no_host_classify = false;  // env MPC_NO_HOST_CLASSIFY: batches of one take the general path too
  void* peer_open[kMaxPeers] = {nullptr};
  std::string err;
};

namespace {

// Entry points run on the engine's device and put the caller's current device back afterwards (a process that
// drives several GPUs must not find its thread's device changed by a call on another engine).
struct DeviceGuard {
  int prev = -1;
  bool changed = false;
  cudaError_t err = cudaSuccess;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    if (prev != dev) {
      err = cudaSetDevice(dev);
      changed = err == cudaSuccess;
    }
  }
  ~DeviceGuard() {
    if (changed && prev >= 0) cudaSetDevice(prev);
  }
};
#define ON_DEVICE(eng)                                                     \
  DeviceGuard dev_guard_((eng)->device);                                   \
  if (dev_guard_.err != cudaSuccess) {                                     \
    (eng)->err = std::string("cudaSetDevice: ") + cudaGetErrorString(dev_guard_.err); \
    return MPC_E_CUDA;                                                     \
  }

#define CK(call)                                                                        \
  do {                                                                                  \
    cudaError_t e_ = (call);                                                            \
    if (e_ != cudaSuccess) {                                                            \
      eng->err = std::string(#call) + ": " + cudaGetErrorString(e_);                    \
      return MPC_E_CUDA;                                                                \
    }                                                                                   \
  } while (0)

// kernel variants: padded size -> (threads, thread grid, register tile)
#ifndef MPC_MINB0
#define MPC_MINB0 4  // resident CTAs per SM the smallest class is compiled for (register budget 65536 / (128 * MINB))
#endif
#ifndef MPC_MINB96
#define MPC_MINB96 2
#endif
#ifndef MPC_MINBG  // catch-all class: CTAs per SM it is compiled for
#define MPC_MINBG 2
#endif
#ifndef MPC_MINB128
#define MPC_MINB128 2
#endif
#ifndef MPC_V64_NT  // thread grid of the smallest class: NT = GR*GC threads, R x C register tiles (GR*R = GC*C = 64)
#define MPC_V64_NT 128
#define MPC_V64_GR 16
#define MPC_V64_R 4
#define MPC_V64_GC 8
#define MPC_V64_C 8
#endif
#define MPC_V64_SHAPE MPC_V64_NT, MPC_V64_GR, MPC_V64_R, MPC_V64_GC, MPC_V64_C
enum { V_64 = 0, V_96, V_128, V_GENERIC, V_WRENCH, V_COUNT };
#define MPC_VARIANT_CALL2(v, PROF, SW, EXPR)                                                                  \
  switch (v) {                                                                                               \
    case V_64: { auto kern = mpc_solve_kernel<MPC_V64_SHAPE, MPC_MINB0, false, PROF, SW>; EXPR; } break;      \
    case V_96: { auto kern = mpc_solve_kernel<256, 16, 6, 16, 6, MPC_MINB96, false, PROF, SW>; EXPR; } break; \
    case V_128: { auto kern = mpc_solve_kernel<256, 16, 8, 16, 8, MPC_MINB128, true, PROF, SW>; EXPR; } break; \
    default: { auto kern = mpc_solve_kernel<256, 0, 0, 0, 0, MPC_MINBG, false, PROF, 0>; EXPR; } break;        \
  }
// prof: the profiling instantiation (always the FMA sweep); sw: which inversion the production instantiation runs
#define MPC_VARIANT_CALL(v, prof, sw, EXPR)                                                           \
  if (prof) { MPC_VARIANT_CALL2(v, true, 0, EXPR) }                                                   \
  else if (sw) { MPC_VARIANT_CALL2(v, false, 1, EXPR) } else { MPC_VARIANT_CALL2(v, false, 0, EXPR) }
#ifndef MPC_NG64   // threads of the active-set role in the piped kernels (32: one warp; more: a barrier group)
#define MPC_NG64 32
#endif
#ifndef MPC_NG96
#define MPC_NG96 128
#endif
#ifndef MPC_NG128
#define MPC_NG128 128
#endif
#define MPC_PIPE_CALL2(v, SW, EXPR)                                                                                \
  switch (v) {                                                                                                    \
    case V_64: { auto kern = mpc_solve_pipe_kernel<MPC_V64_SHAPE, MPC_MINB0, false, MPC_NG64, SW>; EXPR; } break;       \
    case V_96: { auto kern = mpc_solve_pipe_kernel<256, 16, 6, 16, 6, MPC_MINB96, false, MPC_NG96, SW>; EXPR; } break; \
    default: { auto kern = mpc_solve_pipe_kernel<256, 16, 8, 16, 8, MPC_MINB128, true, MPC_NG128, SW>; EXPR; } break; \
  }
#define MPC_PIPE_CALL(v, sw, EXPR) \
  if (sw) { MPC_PIPE_CALL2(v, 1, EXPR) } else { MPC_PIPE_CALL2(v, 0, EXPR) }
const int kVariantThreads[V_COUNT] = {MPC_V64_NT, 256, 256, 256, 256};
const int kVariantPad[V_COUNT] = {64, 96, 128, 0, 128};
#ifndef MPC_MINBW  // wrench-space class: CTAs per SM it is compiled for
#define MPC_MINBW 2
#endif
#define MPC_WRENCH_CALL(prof, sw, big, EXPR)                                                     \
  if (prof) { auto kern = mpc_solve_wrench_kernel<MPC_MINBW, true, 0>; EXPR; }                   \
  else if (big) { auto kern = mpc_solve_wrench_kernel<MPC_MINBW, false, 0, 4>; EXPR; }           \
  else if (sw) { auto kern = mpc_solve_wrench_kernel<MPC_MINBW, false, 1>; EXPR; }               \
  else { auto kern = mpc_solve_wrench_kernel<MPC_MINBW, false, 0>; EXPR; }

int configure_kernel(mpc_batch* eng, ClassCfg& c) {
  // the attribute is per kernel instantiation: raise it to the device limit
  int max_smem = 0;
  CK(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, eng->device));
  int occ = 0;
  for (int prof = 1; prof >= 0; prof--) {  // the production instantiation last: its occupancy sizes the grid
    for (int sw = 1; sw >= 0; sw--) {
      if (prof && sw) continue;
      int o = 0;
      if (c.variant == V_WRENCH) {
        MPC_WRENCH_CALL(prof, sw, !c.in_fast, {
          CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
          CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, kern, c.threads, c.smem));
        });
      } else
      MPC_VARIANT_CALL(c.variant, prof, sw, {
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, kern, c.threads, c.smem));
      });
      occ = (prof || occ == 0) ? o : std::min(occ, o);  // one grid serves both production sweeps
    }
  }
  if (occ < 1) {
    eng->err = "solve kernel does not fit on an SM";
    return MPC_E_CUDA;
  }
  c.grid = occ * eng->sms;
  return MPC_OK;
}

// largest working-set capacity whose T tile hides under the assembly temporaries
int free_m_cap(int h, int nv_cap, int npad) {
  int m = 8;
  while (m < nv_cap) {
    const int mm = m + 1;
    const int gi = mm * (mm | 1) + std::max(2 * (npad + 2), nv_cap) + 1 + 4 * h + 6 * (mm + 1);
    if (gi > mpc::kAsmDoubles(h)) break;
    m = mm;
  }
  return m;
}

int build_classes(mpc_batch* eng) {
  const int h = eng->h, nv_max = 12 * h;
  std::vector<int> caps;
  for (int c : {60, 96, 128})
    if (c < nv_max) caps.push_back(c);
  if (nv_max <= 128) caps.push_back(nv_max);
  int max_smem = 0;
  CK(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, eng->device));
  for (int cap : caps) {
    ClassCfg c;
    c.nv_cap = cap;
    c.variant = cap <= 64 ? V_64 : cap <= 96 ? V_96 : V_128;
    c.threads = kVariantThreads[c.variant];
    c.m_cap = free_m_cap(h, cap, kVariantPad[c.variant]);
    c.in_fast = 1;
    // Packed (lower-triangular) storage of H / H^{-1} halves the tile but costs index arithmetic on every access
    // (measured, same box: -11% at nv <= 60, -7% at nv <= 96 with unchanged occupancy).  It pays where it buys
    // residency: the nv <= 128 class goes from one to two CTAs per SM (+33% on four-stance horizon-10 problems).
    const int packed = c.variant == V_128 ? 1 : 0;  // must match the PK argument of the variant's kernel
    c.L = mpc::make_layout(h, c.nv_cap, c.m_cap, 1, kVariantPad[c.variant], packed);
    c.smem = 16 + 2 * eng->stride + c.L.fast_bytes;
    if ((int)c.smem > max_smem) continue;
    int rc = configure_kernel(eng, c);
    if (rc) return rc;
    // Piped: +12 % for nv <= 60 (trot horizon 10, active set on one warp), +13 % for nv <= 96 (gallop horizon 16) and
    // +11 % for nv <= 128 (four-stance horizon 10), the latter two with the active set on a 128-thread group (on one
    // warp the nv <= 128 class LOSES 9 %: nine working-set changes per problem).
    const char* np = getenv("MPC_NO_PIPE");  // development switch: "1" nothing piped, "2" not the nv <= 128 class
    const bool try_pipe = !(np && (np[0] == '1' || (np[0] == '2' && c.variant == V_128)));
    if (try_pipe && (c.variant != V_64 || MPC_V64_NT == 128)) {
      // piped layout: the largest working-set tile that keeps the class's CTAs per SM; not worth it below 16 rows
      int sm_smem = 0;
      CK(cudaDeviceGetAttribute(&sm_smem, cudaDevAttrMaxSharedMemoryPerMultiprocessor, eng->device));
      const int per_sm = c.grid / eng->sms;                     // resident CTAs of the un-piped kernel
      const size_t budget = (size_t)sm_smem / per_sm - 1024;    // 1 KB per CTA is reserved by the system
      for (int m = c.m_cap; m >= 16; m--) {
        const mpc::Layout Lp = mpc::make_layout(h, c.nv_cap, m, 1, kVariantPad[c.variant], packed, 1);
        const size_t need = 16 + 2 * eng->stride + Lp.fast_bytes;
        if (need > budget || (int)need > max_smem) continue;
        int occ = 1 << 30;
        for (int sw = 0; sw < 2; sw++) {
          int o = 0;
          MPC_PIPE_CALL(c.variant, sw, {
            CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
            CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, kern, c.threads, need));
          });
          occ = std::min(occ, o);
        }
        if (occ >= per_sm) {
          c.pipe = true;
          c.pipe_m_cap = m;
          c.pipe_L = Lp;
          c.pipe_smem = need;
          c.pipe_grid = occ * eng->sms;
        }
        break;
      }
    }
    // Riccati variant of the class: working-set tile of MPC_RIC_MCAP columns (default 8 at nv <= 60 -- what keeps 11
    // warps per SM -- and 10 above: measured best or equal on every class; 7 warps per SM at nv <= 96, 6 at nv <= 128)
    // Which classes: measured on the BASELINE workloads (tools/ab_solver.py, same box) the Riccati kernel wins wherever
    // a sweep over the horizon is cheap next to the inversion it replaces -- nv <= 60 (trot h = 10: x1.35), nv <= 96
    // (gallop h = 16: x1.79), nv <= 128 at short horizons (four-stance h = 10: x1.42) -- and loses for the nv <= 128
    // class at long horizons (config 3, h = 20, 12.5 working-set changes of 40 step-phases each: x0.6..0.86), which
    // therefore keeps the inverse-based kernel (MPC_RIC_128=1 / 0 overrides).
    bool ric_ok = c.variant != V_128 || h <= 12;
    if (const char* e = getenv("MPC_RIC_128")) if (c.variant == V_128) ric_ok = atoi(e) != 0;
    if (!getenv("MPC_NO_RICCATI") && ric_ok) {
      // (the tile holds the working sets of all but a fraction of a percent of the BASELINE problems; the rest move
      // into the per-warp global slab and carry on there)
      int m = c.variant == V_64 ? 8 : 10;
      if (const char* e = getenv("MPC_RIC_MCAP")) m = std::max(4, atoi(e));
      m = std::min(m, c.nv_cap);
      const mpc::RicLayout Lr = mpc::make_ric_layout(h, c.nv_cap, m);
      const size_t need = 16 + eng->stride + Lr.bytes;
      if ((int)need <= max_smem) {
        int o = 0;
        CK(cudaFuncSetAttribute(mpc_solve_riccati_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
        CK(cudaFuncSetAttribute(mpc_solve_riccati_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, mpc_solve_riccati_kernel<false>, 32, need));
        if (o >= 1) {
          c.ric = true;
          c.ric_m_cap = m;
          c.ric_L = Lr;
          c.ric_smem = need;
          c.ric_grid = o * eng->sms;
        }
      }
    }
    eng->classes.push_back(c);
  }
  // wrench-space class: nv > 128 at horizons with 6h <= 128 (see mpc_solve_wrench_kernel).  Its working-set tile is
  // the largest that keeps two CTAs per SM; a larger working set is re-queued to the catch-all like everywhere else.
  if (nv_max > 128 && 6 * h <= 128 && !getenv("MPC_NO_WRENCH")) {
    int sm_smem = 0;
    CK(cudaDeviceGetAttribute(&sm_smem, cudaDevAttrMaxSharedMemoryPerMultiprocessor, eng->device));
    const size_t budget = (size_t)sm_smem / MPC_MINBW - 1024;
    ClassCfg c;
    c.nv_cap = nv_max;
    c.variant = V_WRENCH;
    c.threads = 256;
    c.in_fast = 1;
    c.m_cap = 0;
    for (int m = 120; m >= 16; m -= 4) {
      const mpc::Layout Lw = mpc::make_layout(h, nv_max, m, 1, kVariantPad[V_WRENCH], 1, 0, 1);
      const size_t need = 16 + 2 * eng->stride + Lw.fast_bytes;
      if (need > budget || (int)need > max_smem) continue;
      c.m_cap = m;
      c.L = Lw;
      c.smem = need;
      break;
    }
    if (c.m_cap > 0) {
      int rc = configure_kernel(eng, c);
      if (rc) return rc;
      // Riccati variant, OFF unless MPC_RIC_BIG=1: gains of 12h variables and a working-set tile of MPC_RIC_MCAP_BIG
      // columns (default 16) in shared memory, larger working sets in the per-warp slab.  Measured on config 3 (h = 20,
      // nv = 168 / 240, 12.5 working-set changes on average, up to 70): 0.57 M solves/s against 0.80 with the
      // wrench-space kernel -- two warps per SM and 40 step-phases per working-set change lose against one 6h x 6h
      // inversion.  The Riccati solver pays where the working set stays small next to the horizon.
      if (!getenv("MPC_NO_RICCATI") && getenv("MPC_RIC_BIG") && atoi(getenv("MPC_RIC_BIG")) == 1) {
        int m = 16;
        if (const char* e = getenv("MPC_RIC_MCAP_BIG")) m = std::max(4, atoi(e));
        const mpc::RicLayout Lr = mpc::make_ric_layout(h, nv_max, m);
        const size_t need = 16 + eng->stride + Lr.bytes;
        if ((int)need <= max_smem) {
          int o = 0;
          CK(cudaFuncSetAttribute(mpc_solve_riccati_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
          CK(cudaFuncSetAttribute(mpc_solve_riccati_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
          CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, mpc_solve_riccati_kernel<false>, 32, need));
          if (o >= 1) {
            c.ric = true;
            c.ric_m_cap = m;
            c.ric_L = Lr;
            c.ric_smem = need;
            c.ric_grid = o * eng->sms;
          }
        }
      }
      eng->classes.push_back(c);
      // ... and its own catch-all: the same kernel with a working set of any size, the working-set matrix T in a
      // per-CTA global slab (L2 resident).  With it the dense catch-all below (a 12h x 12h inversion in L2) is
      // not needed at this horizon at all.
      ClassCfg o = c;
      o.ric = false;
      o.m_cap = nv_max;
      o.in_fast = 0;
      o.L = mpc::make_layout(h, nv_max, nv_max, 1, kVariantPad[V_WRENCH], 1, 0, 2);
      o.smem = 16 + 2 * eng->stride + o.L.fast_bytes;
      rc = configure_kernel(eng, o);
      if (rc) return rc;
      eng->classes.push_back(o);
      eng->has_wrench = true;
    }
  }
  // catch-all: full-size problem and working set in a per-CTA global slab (L2 resident)
  ClassCfg big;
  big.nv_cap = nv_max;
  big.m_cap = nv_max;
  big.in_fast = 0;
  big.L = mpc::make_layout(h, big.nv_cap, big.m_cap, 0);
  big.smem = 16 + 2 * eng->stride + big.L.fast_bytes;
  big.variant = V_GENERIC;
  big.threads = kVariantThreads[V_GENERIC];
  int rc = configure_kernel(eng, big);
  if (rc) return rc;
  if (const char* e = getenv("MPC_BIG_CTAS")) big.grid = std::min(big.grid, atoi(e) * eng->sms); else
  big.grid = std::min(big.grid, 2 * eng->sms);  // two slabs per SM: more loads in flight, slabs still mostly L2-resident
  eng->dense_big = big;  // with wrench-space classes: only the assemble-only parity entry still uses it
  if (!eng->has_wrench) eng->classes.push_back(big);
  if ((int)eng->classes.size() > kMaxClasses) {
    eng->err = "too many classes";
    return MPC_E_ARG;
  }
  return MPC_OK;
}

void fill_params(const mpc_batch* eng, int slot, SolveParams& P, const void* records, int batch, float* forces,
                 double* solution, int32_t* status) {
  memset(&P, 0, sizeof(P));
  P.records = (const char*)records;
  P.stride = eng->stride;
  P.h = eng->h;
  P.batch = batch;
  P.forces = forces;
  P.solution = solution;
  P.status = status;
  P.max_iter = eng->max_iter;
  P.warp_mode = 1;
  P.phase_clk = eng->phase_clk;
  P.debug_stop = eng->debug_stop;
  P.ric_generic = eng->ric_generic;
  P.warm_cache = eng->warm_cache;
  P.warm_ids = eng->warm_ids;
  P.warm_shift = eng->warm_shift;
  P.n_peers = eng->gather_fused ? eng->n_peers : 0;
  P.rank_offset = eng->rank_offset;
  for (int q = 0; q < kMaxPeers; q++)
    P.peers[q] = eng->peers[q] ? (float*)((char*)eng->peers[q] + eng->gather_slot_bytes * (size_t)slot) : nullptr;
}

// does this launch of class c go through the Riccati kernel?
// batch: problems of the call (-1: not known / do not care).  One warp per problem pays when the warps fill the SMs; a
// batch that does not even fill the inverse-based kernel's grid (the legacy single-robot tick: a batch of ONE) is a
// latency problem, and there 128 threads on one problem finish sooner than 32 (27 us against 37 us per trot tick).
bool use_riccati(const mpc_batch* eng, const ClassCfg& c, bool assemble_only, int batch = -1) {
  if (!(c.ric && eng->solver == 1 && !eng->phase_clk && !eng->debug_stop && !assemble_only && !eng->warm_cache)) return false;
  if (batch >= 0 && c.variant != V_WRENCH && batch <= (c.pipe ? c.pipe_grid : c.grid) && !eng->ric_always) return false;
  return true;
}

int launch_solve(mpc_batch* eng, const ClassCfg& c, const SolveParams& P, int grid, cudaStream_t st) {
  const bool prof = P.phase_clk != nullptr || P.H_out != nullptr || P.debug_stop != 0;
  if (use_riccati(eng, c, P.H_out != nullptr, P.batch)) {
    SolveParams Pr = P;
    Pr.RL = c.ric_L;
    Pr.ric_slab = eng->cur_ric_slab;
    Pr.queue = (eng->ric_dynamic && eng->cur_ric_queue) ? eng->cur_ric_queue + 2 * eng->cur_class : nullptr;
    if (eng->ric_generic) mpc_solve_riccati_kernel<true><<<grid, 32, c.ric_smem, st>>>(Pr);
    else mpc_solve_riccati_kernel<false><<<grid, 32, c.ric_smem, st>>>(Pr);
    eng->launches++;
    CK(cudaGetLastError());
    return MPC_OK;
  }
  int* const queue = (eng->ric_dynamic && eng->cur_ric_queue && !prof) ? eng->cur_ric_queue + 2 * eng->cur_class : nullptr;
  if (c.pipe && !prof) {
    SolveParams Pp = P;
    Pp.L = c.pipe_L;
    Pp.queue = queue;
    MPC_PIPE_CALL(c.variant, eng->sweep, (kern<<<grid, c.threads, c.pipe_smem, st>>>(Pp)));
    eng->launches++;
    CK(cudaGetLastError());
    return MPC_OK;
  }
  SolveParams Pq = P;
  Pq.queue = queue;
  if (c.variant == V_WRENCH) {
    MPC_WRENCH_CALL(prof, eng->sweep, !c.in_fast, (kern<<<grid, c.threads, c.smem, st>>>(Pq)));
  } else {
    MPC_VARIANT_CALL(c.variant, prof, eng->sweep, (kern<<<grid, c.threads, c.smem, st>>>(Pq)));
  }
  eng->launches++;
  CK(cudaGetLastError());
  return MPC_OK;
}

// single_class >= 0: the caller has classified the batch on the host (every problem belongs to that class): no
// classify kernel, no empty-class launches, identity list.  Used by the host entry for a batch of one -- the legacy
// single-robot tick -- where launch overhead, not the solve, is most of the latency.  A working set that outgrows the
// class's tile cannot be re-queued on this path; it comes back as MAX_ITER with few iterations and the host entry
// repeats the solve on the general path.

// Host-side classification (the host entries have the records in host memory): the size class every problem of the
// batch falls into, or -1 when the batch is mixed (or host classification is off).  A uniform batch -- the usual case:
// one gait, one horizon -- then needs no classify kernel, no index lists and no empty-class launches: ONE kernel
// launch per batch.  The rule is mpc_classify_kernel's (the reference's near-zero test on gait * f_max,
// SolverMPC.cpp:441-469); the scan stops at the first problem that disagrees.
int class_of_nv(const mpc_batch* eng, int nv) {
  int c = 0;
  while (c < (int)eng->classes.size() - 1 && nv > eng->classes[c].nv_cap) c++;
  return c;
}
// number of non-zero bytes among the first n bytes at p (eight at a time)
inline int count_nonzero_bytes(const unsigned char* p, int n) {
  int cnt = 0, q = 0;
  for (; q + 8 <= n; q += 8) {
    uint64_t x;
    memcpy(&x, p + q, 8);
    const uint64_t m = ((x & 0x7f7f7f7f7f7f7f7full) + 0x7f7f7f7f7f7f7f7full) | x;  // bit 7 of a byte set <=> byte != 0
    cnt += __builtin_popcountll(m & 0x8080808080808080ull);
  }
  for (; q < n; q++) cnt += p[q] != 0;
  return cnt;
}
inline bool near_zero_ub(float ub) { return (double)ub < 0.01 && (double)ub > -0.01; }
int host_classify_records(const mpc_batch* eng, const void* records_host, int batch) {
  if (eng->no_host_classify || eng->phase_clk || eng->debug_stop || eng->timed) return -1;  // (kernel timing is per class list)
  // Worth it for small and medium batches only (a launch and a few microseconds of device time per batch saved); the
  // scan itself is bound by host-memory latency -- two cache lines per record, prefetched a few records ahead
  if (batch > kHostClassifyMax) return -1;
  const size_t go = mpc_record_gait_offset(eng->h);
  const int nb = 4 * eng->h;
  int single = -1;
  for (int b = 0; b < batch; b++) {
    const char* base = (const char*)records_host + eng->stride * (size_t)b;
    if (b + 12 < batch) {
      const char* ahead = base + eng->stride * 12;
      __builtin_prefetch(ahead + 4 * MPC_REC_FMAX);
      __builtin_prefetch(ahead + go);
      __builtin_prefetch(ahead + go + nb - 1);
    }
    const float fmax = ((const float*)base)[MPC_REC_FMAX];
    const unsigned char* gait = (const unsigned char*)base + go;
    int ns;
    if (!near_zero_ub(1.0f * fmax) && fmax > 0.f) {
      ns = count_nonzero_bytes(gait, nb);  // gait * f_max >= f_max for every non-zero table entry: all of them count
    } else {  // tiny, negative or non-finite f_max: the test entry by entry, as the kernels do it
      ns = 0;
      for (int q = 0; q < nb; q++) ns += !near_zero_ub((float)gait[q] * fmax);
    }
    const int c = class_of_nv(eng, 3 * ns);
    if (single < 0) single = c;
    else if (c != single) return -1;
  }
  return single;
}
// The same from tick records: the contact table is a function of the gait definition (Gait.cpp:142-166, as in
// build_record_from_tick).  With every offset inside [0, h) the h table rows run through every phase of the cycle
// exactly once, so leg j is in stance in min(duration_j, h) of them whatever the iteration is.
int host_classify_ticks(const mpc_batch* eng, const void* ticks_host, int batch) {
  if (eng->no_host_classify || eng->phase_clk || eng->debug_stop || eng->timed) return -1;
  if (batch > kHostClassifyMax) return -1;
  const int h = eng->h;
  int single = -1;
  for (int b = 0; b < batch; b++) {
    const float* tick = (const float*)ticks_host + (size_t)MPC_TICK_WORDS * b;
    const int32_t* ti = (const int32_t*)tick;
    int ns = 0;
    if (!near_zero_ub(1.0f * tick[MPC_TICK_FMAX])) {
      bool regular = ti[MPC_TICK_ITERATION] >= 0;  // (a negative iteration breaks the modulo of the table rows)
      for (int j = 0; j < 4; j++) regular = regular && ti[MPC_TICK_OFFSETS + j] >= 0 && ti[MPC_TICK_OFFSETS + j] < h;
      if (regular) {
        for (int j = 0; j < 4; j++) ns += std::min(std::max(ti[MPC_TICK_DURATIONS + j], 0), h);
      } else {
        const int it0 = ti[MPC_TICK_ITERATION];
        for (int i = 0; i < h; i++) {
          const int iter = (i + it0 + 1) % h;
          for (int j = 0; j < 4; j++) {
            int progress = iter - ti[MPC_TICK_OFFSETS + j];
            if (progress < 0) progress += h;
            ns += progress < ti[MPC_TICK_DURATIONS + j];
          }
        }
      }
    }
    const int c = class_of_nv(eng, 3 * ns);
    if (single < 0) single = c;
    else if (c != single) return -1;
  }
  return single;
}

int ensure_slot(mpc_batch* eng, int q);

int solve_on_stream(mpc_batch* eng, int slot, const void* records, int batch, float* forces, double* solution,
                    int32_t* status, cudaStream_t st, int32_t* nvar_out, double* H_out, double* g_out,
                    int single_class = -1) {
  if (batch == 0) return MPC_OK;
  if (int rc = ensure_slot(eng, slot)) return rc;
  const int nc = (int)eng->classes.size();
  mpc_batch::Slot& S = eng->s[slot];
  eng->cur_ric_slab = S.ric_slab;
  eng->cur_ric_queue = S.ric_queue;
  if (single_class >= 0) {
    eng->cur_class = single_class;
    const ClassCfg& c = eng->classes[single_class];
    SolveParams P;
    fill_params(eng, slot, P, records, batch, forces, solution, status);
    P.L = c.L;
    P.warp_mode = c.variant == V_64 ? 1 : 0;
    P.slab = c.in_fast ? nullptr : S.slab;
    const bool piped = c.pipe && !eng->phase_clk && !eng->debug_stop;
    int grid = std::min(use_riccati(eng, c, false, batch) ? c.ric_grid : piped ? c.pipe_grid : c.grid, batch);
    if (eng->ctas_per_sm_limit > 0) grid = std::min(grid, eng->ctas_per_sm_limit * eng->sms);
    return launch_solve(eng, c, P, grid, st);  // a working-set tile overflow comes back as MAX_ITER (see wait_host)
  }
  // class counters are double-buffered by solve parity: this solve's classify kernel zeroes the other set
  int* counts = S.counts + (S.parity ? kMaxClasses : 0);
  int* counts_next = S.counts + (S.parity ? 0 : kMaxClasses);
  S.parity ^= 1;
  mpc_classify_kernel<<<(batch + 127) / 128, 128, 0, st>>>((const char*)records, eng->stride, eng->h, batch, nc,
                                                           eng->caps_dev, S.lists, counts, counts_next, eng->max_batch);
  eng->launches++;
  CK(cudaGetLastError());
  const bool time_all = eng->timed && eng->timed_class < 0;
  if (time_all) CK(cudaEventRecord(eng->ev0, st));
  for (int ci = 0; ci < nc; ci++) {
    // the assemble-only parity entry writes the reduced QP itself out: the wrench-space class never forms it, so its
    // problems go through the catch-all kernel there
    const ClassCfg& c = (H_out && eng->classes[ci].variant == V_WRENCH) ? eng->dense_big : eng->classes[ci];
    SolveParams P;
    fill_params(eng, slot, P, records, batch, forces, solution, status);
    P.list = S.lists + (size_t)ci * eng->max_batch;
    P.count = counts + ci;
    P.L = c.L;
    // the active-set stage runs on one warp for the smallest class (its per-iteration work fits 32 lanes and
    // __syncwarp beats CTA barriers) and on the whole CTA for the larger ones (measured: four-stance h=10 is
    // 11% faster CTA-wide, gallop h=16 indifferent, trot h=10 4% faster on one warp)
    P.warp_mode = (c.variant == V_64 && !getenv("MPC_BLOCK_GI")) ? 1 : 0;
    P.slab = c.in_fast ? nullptr : S.slab;
    if (ci != nc - 1) {
      P.retry_list = S.lists + (size_t)(nc - 1) * eng->max_batch;
      P.retry_count = counts + (nc - 1);
    }
    P.nvar_out = nvar_out;
    P.H_out = H_out;
    P.g_out = g_out;
    const size_t ring = (size_t)(eng->ring_pos % kRing) * kMaxClasses + ci;
    const bool time_this = eng->timed && (eng->timed_class < 0 || eng->timed_class == ci);
    if (time_this) CK(cudaEventRecord(eng->ring0[ring], st));
    eng->cur_class = ci;
    const bool piped = c.pipe && !eng->phase_clk && !H_out && !eng->debug_stop;
    int grid = std::min(use_riccati(eng, c, H_out != nullptr, batch) ? c.ric_grid : piped ? c.pipe_grid : c.grid, batch);
    if (eng->ctas_per_sm_limit > 0) grid = std::min(grid, eng->ctas_per_sm_limit * eng->sms);
    int rc = launch_solve(eng, c, P, grid, st);
    if (rc) return rc;
    if (time_this) CK(cudaEventRecord(eng->ring1[ring], st));
  }
  if (time_all) CK(cudaEventRecord(eng->ev1, st));
  if (eng->timed) eng->ring_pos++;
  return MPC_OK;
}

// Scratch slot q: stream, device / pinned buffers, class lists, and the catch-all class's per-CTA slab (sized for the
// CTAs a batch of max_batch problems can occupy).  Allocated on first use.
int ensure_slot(mpc_batch* eng, int q) {
  mpc_batch::Slot& S = eng->s[q];
  if (S.stream) return MPC_OK;
#define CKS(call)                                                                        \
  do {                                                                                   \
    cudaError_t e_ = (call);                                                             \
    if (e_ != cudaSuccess) {                                                             \
      eng->err = std::string(#call) + ": " + cudaGetErrorString(e_);                     \
      return e_ == cudaErrorMemoryAllocation ? MPC_E_NOMEM : MPC_E_CUDA;                 \
    }                                                                                    \
  } while (0)
  const size_t B = (size_t)eng->max_batch, NU = 12 * (size_t)eng->h;
  // the slab serves the last class (dense catch-all or wrench-space overflow class) and the assemble-only entry
  const ClassCfg& lastc = eng->classes.back();
  const size_t slab_bytes = std::max(lastc.L.slab_bytes * (size_t)std::min<long long>(lastc.grid, eng->max_batch),
                                     eng->dense_big.L.slab_bytes * (size_t)std::min<long long>(eng->dense_big.grid, eng->max_batch));
  cudaStream_t st = nullptr;
  CKS(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  CKS(cudaMalloc(&S.rec_dev, B * eng->stride));
  // forces | status | solution in ONE block per side, so that a small batch's results come back in one copy
  const size_t off_st = (B * 12 * sizeof(float) + 15) / 16 * 16;
  const size_t off_sol = (off_st + B * sizeof(int32_t) + 15) / 16 * 16;
  S.out_bytes = off_sol + B * NU * sizeof(double);
  CKS(cudaMalloc(&S.out_dev, S.out_bytes));
  S.forces_dev = (float*)S.out_dev;
  S.status_dev = (int32_t*)(S.out_dev + off_st);
  S.sol_dev = (double*)(S.out_dev + off_sol);
  CKS(cudaMallocHost(&S.rec_pin, B * eng->stride));
  CKS(cudaMallocHost(&S.out_pin, S.out_bytes));
  S.forces_pin = (float*)S.out_pin;
  S.status_pin = (int32_t*)(S.out_pin + off_st);
  S.sol_pin = (double*)(S.out_pin + off_sol);
  CKS(cudaMalloc(&S.lists, sizeof(int) * kMaxClasses * B));
  CKS(cudaMalloc(&S.counts, sizeof(int) * 2 * kMaxClasses));
  CKS(cudaMemset(S.counts, 0, sizeof(int) * 2 * kMaxClasses));
  CKS(cudaMalloc(&S.slab, slab_bytes));
  size_t ric_bytes = 0;
  for (const ClassCfg& c : eng->classes)
    if (c.ric) ric_bytes = std::max(ric_bytes, c.ric_L.slab_bytes * (size_t)std::min<long long>(c.ric_grid, eng->max_batch));
  if (ric_bytes) CKS(cudaMalloc(&S.ric_slab, ric_bytes));
  CKS(cudaMalloc(&S.ric_queue, sizeof(int) * 2 * kMaxClasses));
  CKS(cudaMemset(S.ric_queue, 0, sizeof(int) * 2 * kMaxClasses));
  S.stream = st;  // last: marks the slot as complete
#undef CKS
  return MPC_OK;
}

// Event ring of the kernel timing, created when timing is first switched on.
int ensure_timing(mpc_batch* eng) {
  if (!eng->ring0.empty()) return MPC_OK;
  eng->ring0.assign((size_t)kRing * kMaxClasses, nullptr);
  eng->ring1.assign((size_t)kRing * kMaxClasses, nullptr);
  for (size_t i = 0; i < eng->ring0.size(); i++) {
    CK(cudaEventCreate(&eng->ring0[i]));
    CK(cudaEventCreate(&eng->ring1[i]));
  }
  return MPC_OK;
}

}  // namespace

extern "C" {

size_t mpc_record_stride(int h) { return ((size_t)(4 * (MPC_REC_TRAJ + 12 * h) + 4 * h) + 15) / 16 * 16; }
size_t mpc_record_gait_offset(int h) { return (size_t)4 * (MPC_REC_TRAJ + 12 * h); }

const char* mpc_last_error(void) { return g_err.c_str(); }

int mpc_batch_create(mpc_batch_t** out, int device, int horizon, int max_batch) {
  if (!out || horizon < 1 || horizon > MPC_MAX_HORIZON || max_batch < 1) {
    g_err = "mpc_batch_create: bad argument";
    return MPC_E_ARG;
  }
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) {
    g_err = "mpc_batch_create: no such CUDA device (this library has no CPU path)";
    return MPC_E_NODEVICE;
  }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess || prop.major != 10) {
    g_err = "mpc_batch_create: device is not sm_100 (the kernels are built for sm_100a only)";
    return MPC_E_NODEVICE;
  }
  mpc_batch* eng = new mpc_batch();
  auto fail = [&](int rc) {
    g_err = eng->err;
    mpc_batch_destroy(eng);
    return rc;
  };
#define CKC(call)                                                       \
  do {                                                                  \
    cudaError_t e_ = (call);                                            \
    if (e_ != cudaSuccess) {                                            \
      eng->err = std::string(#call) + ": " + cudaGetErrorString(e_);    \
      return fail(e_ == cudaErrorMemoryAllocation ? MPC_E_NOMEM : MPC_E_CUDA); \
    }                                                                   \
  } while (0)
  if (const char* ds = getenv("MPC_DEBUG_STOP")) eng->debug_stop = atoi(ds);
  if (const char* rg = getenv("MPC_RIC_GENERIC")) eng->ric_generic = atoi(rg);
  if (const char* ra = getenv("MPC_RIC_ALWAYS")) eng->ric_always = atoi(ra) != 0;
  if (const char* rs = getenv("MPC_RIC_STATIC")) eng->ric_dynamic = atoi(rs) == 0;
  eng->no_host_classify = getenv("MPC_NO_HOST_CLASSIFY") != nullptr;
  if (const char* sw = getenv("MPC_SWEEP")) eng->sweep = (sw[0] == 'm' || sw[0] == '1') ? 1 : 0;  // "mma" / "fma"
  if (const char* sv = getenv("MPC_SOLVER")) eng->solver = (sv[0] == 'r' || sv[0] == '1') ? 1 : 0;  // "riccati" / "inverse"
  eng->device = device;
  eng->h = horizon;
  eng->max_batch = max_batch;
  eng->sms = prop.multiProcessorCount;
  eng->stride = mpc_record_stride(horizon);
  DeviceGuard guard(device);  // the caller's current device is put back on return
  if (guard.err != cudaSuccess) {
    eng->err = std::string("cudaSetDevice: ") + cudaGetErrorString(guard.err);
    return fail(MPC_E_CUDA);
  }
  CKC(cudaEventCreate(&eng->ev0));
  CKC(cudaEventCreate(&eng->ev1));
  int rc = build_classes(eng);
  if (rc) return fail(rc);
  rc = ensure_slot(eng, 0);  // the other scratch slots are allocated on first use (the legacy single-robot engine never needs them)
  if (rc) return fail(rc);
  CKC(cudaMalloc(&eng->caps_dev, sizeof(int) * kMaxClasses));
  int caps[kMaxClasses] = {0};
  for (size_t i = 0; i < eng->classes.size(); i++) caps[i] = eng->classes[i].nv_cap;
  CKC(cudaMemcpy(eng->caps_dev, caps, sizeof(caps), cudaMemcpyHostToDevice));
#undef CKC
  *out = eng;
  return MPC_OK;
}

void mpc_batch_destroy(mpc_batch_t* eng) {
  if (!eng) return;
  DeviceGuard guard(eng->device);
  for (int q = 0; q < kMaxPeers; q++)
    if (eng->peer_open[q]) cudaIpcCloseMemHandle(eng->peer_open[q]);
  cudaFree(eng->gather_buf);
  cudaFree(eng->peer_flags_dev);
  for (int q = 0; q < kSlots; q++) {
    mpc_batch::Slot& S = eng->s[q];
    cudaFree(S.rec_dev);
    cudaFree(S.out_dev);
    cudaFreeHost(S.rec_pin);
    cudaFreeHost(S.out_pin);
    cudaFree(S.tick_dev);
    cudaFreeHost(S.tick_pin);
    cudaFree(S.state_dev);
    cudaFreeHost(S.state_pin);
    cudaFree(S.lists);
    cudaFree(S.counts);
    cudaFree(S.slab);
    cudaFree(S.ric_slab);
    cudaFree(S.ric_queue);
    if (S.stream) cudaStreamDestroy(S.stream);
  }
  cudaFree(eng->caps_dev);
  if (eng->ev0) cudaEventDestroy(eng->ev0);
  if (eng->ev1) cudaEventDestroy(eng->ev1);
  for (cudaEvent_t e : eng->ring0)
    if (e) cudaEventDestroy(e);
  for (cudaEvent_t e : eng->ring1)
    if (e) cudaEventDestroy(e);
  delete eng;
}

int mpc_batch_solve_device(mpc_batch_t* eng, const void* records_dev, int batch, float* forces_dev,
                           double* solution_dev, int32_t* status_dev, void* cuda_stream) {
  if (!eng) return MPC_E_ARG;
  if (!records_dev || !forces_dev || batch < 0 || batch > eng->max_batch || ((uintptr_t)records_dev & 15)) {
    eng->err = "mpc_batch_solve_device: bad argument (null pointer, batch out of range or records not 16-byte aligned)";
    return MPC_E_ARG;
  }
  ON_DEVICE(eng);
  return solve_on_stream(eng, 0, records_dev, batch, forces_dev, solution_dev, status_dev, (cudaStream_t)cuda_stream,
                         nullptr, nullptr, nullptr);
}

int mpc_batch_solve_device_slot(mpc_batch_t* eng, int slot, const void* records_dev, int batch, float* forces_dev,
                                double* solution_dev, int32_t* status_dev, void* cuda_stream) {
  if (!eng) return MPC_E_ARG;
  if (slot < 0 || slot >= kSlots || !records_dev || !forces_dev || batch < 0 || batch > eng->max_batch ||
      ((uintptr_t)records_dev & 15)) {
    eng->err = "mpc_batch_solve_device_slot: bad argument";
    return MPC_E_ARG;
  }
  ON_DEVICE(eng);
  return solve_on_stream(eng, slot, records_dev, batch, forces_dev, solution_dev, status_dev,
                         (cudaStream_t)cuda_stream, nullptr, nullptr, nullptr);
}

// zero_copy: the caller vouches that records_host is page-locked and stays untouched until wait_host(slot)
static int submit_host_impl(mpc_batch_t* eng, int slot, const void* records_host, int batch, int want_solution,
                            bool zero_copy);

int mpc_batch_submit_host(mpc_batch_t* eng, int slot, const void* records_host, int batch, int want_solution) {
  return submit_host_impl(eng, slot, records_host, batch, want_solution, false);
}

int mpc_batch_submit_host_pinned(mpc_batch_t* eng, int slot, const void* records_pinned, int batch, int want_solution) {
  if (eng && records_pinned) {
    cudaPointerAttributes attr;
    const bool pinned = cudaPointerGetAttributes(&attr, records_pinned) == cudaSuccess && attr.type == cudaMemoryTypeHost;
    if (!pinned) {
      cudaGetLastError();
      eng->err = "mpc_batch_submit_host_pinned: records are not in page-locked host memory";
      return MPC_E_ARG;
    }
  }
  return submit_host_impl(eng, slot, records_pinned, batch, want_solution, true);
}

static int submit_host_impl(mpc_batch_t* eng, int slot, const void* records_host, int batch, int want_solution,
                            bool zero_copy) {
  if (!eng) return MPC_E_ARG;
  if (slot < 0 || slot >= kSlots || !records_host || batch < 0 || batch > eng->max_batch) {
    eng->err = "mpc_batch_submit_host: bad argument";
    return MPC_E_ARG;
  }
  ON_DEVICE(eng);
  if (int rc = ensure_slot(eng, slot)) return rc;
  mpc_batch::Slot& S = eng->s[slot];
  S.pending_batch = batch;
  S.pending_solution = want_solution != 0;
  S.pending_state = false;
  if (batch == 0) return MPC_OK;
  const size_t NU = 12 * (size_t)eng->h;
  // The records are COPIED before this call returns (staged into the slot's pinned buffer chunk by chunk, chunk c's
  // DMA overlapping the host copy of chunk c+1), so the caller may reuse its buffer at once -- whatever kind of
  // memory it is.  The DMA reads a buffer in place only when it is the slot's own pinned buffer or when the caller
  // opted in (mpc_batch_submit_host_pinned: page-locked, untouched until wait_host).
  const size_t bytes = (size_t)batch * eng->stride;
  const bool in_place = zero_copy || records_host == (const void*)S.rec_pin;
  if (in_place) {
    CK(cudaMemcpyAsync(S.rec_dev, records_host, bytes, cudaMemcpyHostToDevice, S.stream));
  } else {
    const size_t chunk = 512u << 10;
    for (size_t off = 0; off < bytes; off += chunk) {
      const size_t n = std::min(chunk, bytes - off);
      memcpy(S.rec_pin + off, (const char*)records_host + off, n);
      CK(cudaMemcpyAsync(S.rec_dev + off, S.rec_pin + off, n, cudaMemcpyHostToDevice, S.stream));
    }
  }
  const int single = host_classify_records(eng, records_host, batch);  // uniform batch: one launch, no classify kernel
  S.pending_single_class = single;
  int rc = solve_on_stream(eng, slot, S.rec_dev, batch, S.forces_dev, want_solution ? S.sol_dev : nullptr, S.status_dev,
                           S.stream, nullptr, nullptr, nullptr, single);
  if (rc) return rc;
  if (S.out_bytes <= 8192) {  // small engine (the legacy single-robot one): one copy brings everything back
    CK(cudaMemcpyAsync(S.out_pin, S.out_dev, S.out_bytes, cudaMemcpyDeviceToHost, S.stream));
    return MPC_OK;
  }
  if (batch == eng->max_batch && !want_solution) {  // forces | status are one contiguous block: one copy
    const size_t bytes_out = (size_t)((char*)S.status_dev - S.out_dev) + (size_t)batch * sizeof(int32_t);
    CK(cudaMemcpyAsync(S.out_pin, S.out_dev, bytes_out, cudaMemcpyDeviceToHost, S.stream));
    return MPC_OK;
  }
  CK(cudaMemcpyAsync(S.forces_pin, S.forces_dev, (size_t)batch * 12 * sizeof(float), cudaMemcpyDeviceToHost, S.stream));
  if (want_solution)
    CK(cudaMemcpyAsync(S.sol_pin, S.sol_dev, (size_t)batch * NU * sizeof(double), cudaMemcpyDeviceToHost, S.stream));
  CK(cudaMemcpyAsync(S.status_pin, S.status_dev, (size_t)batch * sizeof(int32_t), cudaMemcpyDeviceToHost, S.stream));
  return MPC_OK;
}

int mpc_batch_wait_host(mpc_batch_t* eng, int slot, float* forces_host, double* solution_host, int32_t* status_host) {
  if (!eng) return MPC_E_ARG;
  if (slot < 0 || slot >= kSlots) {
    eng->err = "mpc_batch_wait_host: bad slot";
    return MPC_E_ARG;
  }
  mpc_batch::Slot& S = eng->s[slot];
  if (!S.stream) return MPC_OK;  // nothing was ever submitted on this slot
  const int batch = S.pending_batch;
  if (solution_host && !S.pending_solution) {
    eng->err = "mpc_batch_wait_host: the solution was not requested at submit time";
    return MPC_E_ARG;
  }
  ON_DEVICE(eng);
  CK(cudaStreamSynchronize(S.stream));
  const size_t NU = 12 * (size_t)eng->h;
  if (S.pending_single_class >= 0 && S.pending_single_class < (int)eng->classes.size() - 1) {
    // host-classified solve: a problem whose working set outgrew its class's tile could not be re-queued on that path
    // and came back as MAX_ITER short of the iteration cap -- the batch goes through the general path once more
    // (never on the BASELINE workloads; the tiles hold 17..27 rows)
    bool overflow = false;
    for (int b = 0; b < batch && !overflow; b++)
      overflow = (S.status_pin[b] & 0xff) == MPC_STATUS_MAX_ITER && (S.status_pin[b] >> 8) < eng->max_iter;
    if (overflow) {
      S.pending_single_class = -1;
      int rc = solve_on_stream(eng, slot, S.rec_dev, batch, S.forces_dev, S.pending_solution ? S.sol_dev : nullptr,
                               S.status_dev, S.stream, nullptr, nullptr, nullptr);
      if (rc) return rc;
      CK(cudaMemcpyAsync(S.forces_pin, S.forces_dev, (size_t)batch * 12 * sizeof(float), cudaMemcpyDeviceToHost, S.stream));
      if (S.pending_solution)
        CK(cudaMemcpyAsync(S.sol_pin, S.sol_dev, (size_t)batch * NU * sizeof(double), cudaMemcpyDeviceToHost, S.stream));
      CK(cudaMemcpyAsync(S.status_pin, S.status_dev, (size_t)batch * sizeof(int32_t), cudaMemcpyDeviceToHost, S.stream));
      CK(cudaStreamSynchronize(S.stream));
    }
  }
  if (forces_host && forces_host != S.forces_pin) memcpy(forces_host, S.forces_pin, (size_t)batch * 12 * sizeof(float));
  if (solution_host && solution_host != S.sol_pin)
    memcpy(solution_host, S.sol_pin, (size_t)batch * NU * sizeof(double));
  if (status_host && status_host != S.status_pin) memcpy(status_host, S.status_pin, (size_t)batch * sizeof(int32_t));
  S.pending_batch = 0;
  return MPC_OK;
}

// Host entry from TICK records (SURVEY 8f N1 + N2 on the host path): 272 bytes per robot cross the bus instead of the
// 720-byte problem record (h = 10); the records are built on the device (mpc_build_records_kernel) into the slot's
// own buffer and never exist in host memory.  The controller state the reference writes back at this point
// (world_position_desired after the clamp, next x_comp_integral) comes back with the forces: mpc_batch_host_state.
int mpc_batch_submit_host_ticks(mpc_batch_t* eng, int slot, const void* ticks_host, int batch, int want_solution,
                                int zero_copy) {
  if (!eng) return MPC_E_ARG;
  if (slot < 0 || slot >= kSlots || !ticks_host || batch < 0 || batch > eng->max_batch) {
    eng->err = "mpc_batch_submit_host_ticks: bad argument";
    return MPC_E_ARG;
  }
  if (zero_copy) {
    cudaPointerAttributes attr;
    const bool pinned = cudaPointerGetAttributes(&attr, ticks_host) == cudaSuccess && attr.type == cudaMemoryTypeHost;
    if (!pinned) {
      cudaGetLastError();
      eng->err = "mpc_batch_submit_host_ticks: zero_copy asks for page-locked tick records";
      return MPC_E_ARG;
    }
  }
  ON_DEVICE(eng);
  if (int rc = ensure_slot(eng, slot)) return rc;
  mpc_batch::Slot& S = eng->s[slot];
  if (!S.tick_dev) {
    const size_t B = (size_t)eng->max_batch;
    CK(cudaMalloc(&S.tick_dev, B * MPC_TICK_STRIDE));
    CK(cudaMallocHost(&S.tick_pin, B * MPC_TICK_STRIDE));
    CK(cudaMalloc(&S.state_dev, B * 4 * sizeof(float)));
    CK(cudaMallocHost(&S.state_pin, B * 4 * sizeof(float)));
  }
  S.pending_batch = batch;
  S.pending_solution = want_solution != 0;
  S.pending_state = true;
  if (batch == 0) return MPC_OK;
  const size_t bytes = (size_t)batch * MPC_TICK_STRIDE;
  const void* src = ticks_host;
  if (!zero_copy && ticks_host != (const void*)S.tick_pin) {
    memcpy(S.tick_pin, ticks_host, bytes);  // the caller may reuse its buffer at once
    src = S.tick_pin;
  }
  CK(cudaMemcpyAsync(S.tick_dev, src, bytes, cudaMemcpyHostToDevice, S.stream));
  mpc_build_records_kernel<<<(batch + 127) / 128, 128, 0, S.stream>>>((const float*)S.tick_dev, batch, eng->h, S.rec_dev,
                                                                     eng->stride, S.state_dev);
  eng->launches++;
  CK(cudaGetLastError());
  const int single = host_classify_ticks(eng, ticks_host, batch);
  S.pending_single_class = single;
  int rc = solve_on_stream(eng, slot, S.rec_dev, batch, S.forces_dev, want_solution ? S.sol_dev : nullptr, S.status_dev,
                           S.stream, nullptr, nullptr, nullptr, single);
  if (rc) return rc;
  const size_t NU = 12 * (size_t)eng->h;
  if (batch == eng->max_batch && !want_solution) {  // forces | status are one contiguous block: one copy
    const size_t bytes_out = (size_t)((char*)S.status_dev - S.out_dev) + (size_t)batch * sizeof(int32_t);
    CK(cudaMemcpyAsync(S.out_pin, S.out_dev, bytes_out, cudaMemcpyDeviceToHost, S.stream));
  } else {
    CK(cudaMemcpyAsync(S.forces_pin, S.forces_dev, (size_t)batch * 12 * sizeof(float), cudaMemcpyDeviceToHost, S.stream));
    if (want_solution)
      CK(cudaMemcpyAsync(S.sol_pin, S.sol_dev, (size_t)batch * NU * sizeof(double), cudaMemcpyDeviceToHost, S.stream));
    CK(cudaMemcpyAsync(S.status_pin, S.status_dev, (size_t)batch * sizeof(int32_t), cudaMemcpyDeviceToHost, S.stream));
  }
  CK(cudaMemcpyAsync(S.state_pin, S.state_dev, (size_t)batch * 4 * sizeof(float), cudaMemcpyDeviceToHost, S.stream));
  return MPC_OK;
}

int mpc_batch_host_state(mpc_batch_t* eng, int slot, float** state_host, void** ticks_pinned) {
  if (!eng || slot < 0 || slot >= kSlots) return MPC_E_ARG;
  mpc_batch::Slot& S = eng->s[slot];
  if (!S.tick_dev) {  // allocate the tick-side buffers so that a caller can fill the slot's own pinned tick buffer
    ON_DEVICE(eng);
    if (int rc = ensure_slot(eng, slot)) return rc;
    const size_t B = (size_t)eng->max_batch;
    CK(cudaMalloc(&S.tick_dev, B * MPC_TICK_STRIDE));
    CK(cudaMallocHost(&S.tick_pin, B * MPC_TICK_STRIDE));
    CK(cudaMalloc(&S.state_dev, B * 4 * sizeof(float)));
    CK(cudaMallocHost(&S.state_pin, B * 4 * sizeof(float)));
  }
  if (state_host) *state_host = S.state_pin;
  if (ticks_pinned) *ticks_pinned = S.tick_pin;
  return MPC_OK;
}

int mpc_batch_slots(void) { return kSlots; }

int mpc_batch_solve_host(mpc_batch_t* eng, const void* records_host, int batch, float* forces_host,
                         double* solution_host, int32_t* status_host) {
  if (!eng) return MPC_E_ARG;
  if (!records_host || !forces_host || batch < 0 || batch > eng->max_batch) {
    eng->err = "mpc_batch_solve_host: bad argument";
    return MPC_E_ARG;
  }
  int rc = mpc_batch_submit_host(eng, 0, records_host, batch, solution_host != nullptr);
  if (rc) return rc;
  return mpc_batch_wait_host(eng, 0, forces_host, solution_host, status_host);
}

int mpc_batch_build_records_device(mpc_batch_t* eng, const void* ticks_dev, int batch, void* records_dev,
                                   float* state_out_dev, void* cuda_stream) {
  if (!eng) return MPC_E_ARG;
  if (!ticks_dev || !records_dev || batch < 0 || batch > eng->max_batch || ((uintptr_t)records_dev & 15)) {
    eng->err = "mpc_batch_build_records_device: bad argument";
    return MPC_E_ARG;
  }
  if (batch == 0) return MPC_OK;
  ON_DEVICE(eng);
  mpc_build_records_kernel<<<(batch + 127) / 128, 128, 0, (cudaStream_t)cuda_stream>>>(
      (const float*)ticks_dev, batch, eng->h, (char*)records_dev, eng->stride, state_out_dev);
  eng->launches++;
  CK(cudaGetLastError());
  return MPC_OK;
}

int mpc_batch_solve_ticks_device(mpc_batch_t* eng, const void* ticks_dev, int batch, float* forces_dev,
                                 double* solution_dev, int32_t* status_dev, float* state_out_dev, void* cuda_stream) {
  if (!eng) return MPC_E_ARG;
  int rc = mpc_batch_build_records_device(eng, ticks_dev, batch, eng->s[0].rec_dev, state_out_dev, cuda_stream);
  if (rc) return rc;
  return mpc_batch_solve_device(eng, eng->s[0].rec_dev, batch, forces_dev, solution_dev, status_dev, cuda_stream);
}

int mpc_batch_gait_state_device(mpc_batch_t* eng, const void* gait_dev, int batch, void* state_out_dev,
                                unsigned char* table_out_dev, int table_stride, void* cuda_stream) {
  if (!eng) return MPC_E_ARG;
  if (!gait_dev || !state_out_dev || batch < 0 || (table_out_dev && table_stride < 4)) {
    eng->err = "mpc_batch_gait_state_device: bad argument";
    return MPC_E_ARG;
  }
  if (batch == 0) return MPC_OK;
  ON_DEVICE(eng);
  mpc_gait_state_kernel<<<(batch + 127) / 128, 128, 0, (cudaStream_t)cuda_stream>>>(
      (const int32_t*)gait_dev, batch, (float*)state_out_dev, table_out_dev, table_stride);
  eng->launches++;
  CK(cudaGetLastError());
  return MPC_OK;
}

int mpc_batch_leg_commands_device(mpc_batch_t* eng, const void* legs_dev, const float* forces_dev, int batch,
                                  float* f_ff_dev, float* tau_dev, void* cuda_stream) {
  if (!eng) return MPC_E_ARG;
  if (!legs_dev || !forces_dev || !f_ff_dev || !tau_dev || batch < 0) {
    eng->err = "mpc_batch_leg_commands_device: bad argument";
    return MPC_E_ARG;
  }
  if (batch == 0) return MPC_OK;
  ON_DEVICE(eng);
  mpc_leg_commands_kernel<<<(batch + 127) / 128, 128, 0, (cudaStream_t)cuda_stream>>>(
      (const float*)legs_dev, forces_dev, batch, f_ff_dev, tau_dev);
  eng->launches++;
  CK(cudaGetLastError());
  return MPC_OK;
}

int mpc_batch_assemble_device(mpc_batch_t* eng, const void* records_dev, int batch, int32_t* nvar_dev, double* H_dev,
                              double* g_dev, void* cuda_stream) {
  if (!eng) return MPC_E_ARG;
  if (!records_dev || !H_dev || batch < 0 || batch > eng->max_batch) {
    eng->err = "mpc_batch_assemble_device: bad argument";
    return MPC_E_ARG;
  }
  ON_DEVICE(eng);
  return solve_on_stream(eng, 0, records_dev, batch, eng->s[0].forces_dev, nullptr, nullptr, (cudaStream_t)cuda_stream,
                         nvar_dev, H_dev, g_dev);
}

int mpc_batch_set_gather_peers(mpc_batch_t* eng, float* const* peers, int n_peers, int rank_offset) {
  if (!eng || n_peers < 0 || n_peers > kMaxPeers || (n_peers > 0 && !peers)) return MPC_E_ARG;
  for (int q = 0; q < kMaxPeers; q++) eng->peers[q] = q < n_peers ? peers[q] : nullptr;
  eng->n_peers = n_peers;
  eng->rank_offset = rank_offset;
  return MPC_OK;
}

int mpc_batch_gather_alloc(mpc_batch_t* eng, int world_batch, void* ipc_handle_out) {
  if (!eng || world_batch < 1 || !ipc_handle_out) return MPC_E_ARG;
  ON_DEVICE(eng);
  if (eng->gather_buf) CK(cudaFree(eng->gather_buf));
  eng->gather_buf = nullptr;
  eng->gather_slot_bytes = ((size_t)world_batch * 12 * sizeof(float) + kMaxPeers * sizeof(unsigned) + 255) / 256 * 256;
  const size_t bytes = kSlots * eng->gather_slot_bytes;  // one region per scratch slot
  CK(cudaMalloc(&eng->gather_buf, bytes));
  CK(cudaMemset(eng->gather_buf, 0, bytes));
  eng->gather_rows = world_batch;
  cudaIpcMemHandle_t h;
  CK(cudaIpcGetMemHandle(&h, eng->gather_buf));
  static_assert(sizeof(h) == 64, "CUDA IPC handles are 64 bytes");
  memcpy(ipc_handle_out, &h, sizeof(h));
  return MPC_OK;
}

int mpc_batch_gather_connect(mpc_batch_t* eng, const void* ipc_handles, int world, int rank, int rank_offset) {
  if (!eng || !ipc_handles || world < 1 || world > kMaxPeers || rank < 0 || rank >= world || !eng->gather_buf)
    return MPC_E_ARG;
  ON_DEVICE(eng);
  float* peers[kMaxPeers] = {nullptr};
  for (int q = 0; q < world; q++) {
    if (q == rank) {
      peers[q] = eng->gather_buf;
      continue;
    }
    cudaIpcMemHandle_t h;
    memcpy(&h, (const char*)ipc_handles + 64 * (size_t)q, 64);
    void* ptr = nullptr;
    CK(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
    eng->peer_open[q] = ptr;
    peers[q] = (float*)ptr;
  }
  unsigned* flags[kSlots][kMaxPeers] = {{nullptr}};
  for (int sl = 0; sl < kSlots; sl++)
    for (int q = 0; q < world; q++)
      flags[sl][q] = (unsigned*)((char*)peers[q] + eng->gather_slot_bytes * sl) + (size_t)eng->gather_rows * 12;
  if (!eng->peer_flags_dev) CK(cudaMalloc(&eng->peer_flags_dev, sizeof(flags)));
  CK(cudaMemcpy(eng->peer_flags_dev, flags, sizeof(flags), cudaMemcpyHostToDevice));
  eng->gather_world = world;
  eng->gather_rank = rank;
  for (int sl = 0; sl < kSlots; sl++) eng->gather_epoch[sl] = 0;
  return mpc_batch_set_gather_peers(eng, peers, world, rank_offset);
}

int mpc_batch_gather_sync_slot(mpc_batch_t* eng, int slot, void* cuda_stream) {
  if (!eng || slot < 0 || slot >= kSlots || !eng->peer_flags_dev || eng->gather_world < 1) return MPC_E_ARG;
  ON_DEVICE(eng);
  cudaStream_t st = (cudaStream_t)cuda_stream;
  const unsigned epoch = ++eng->gather_epoch[slot];
  const unsigned* my_flags =
      (const unsigned*)((const char*)eng->gather_buf + eng->gather_slot_bytes * slot) + (size_t)eng->gather_rows * 12;
  // one launch: lane q tells rank q "my rows of this epoch have landed", then waits for rank q's word here
  mpc_gather_barrier_kernel<<<1, 32, 0, st>>>(eng->peer_flags_dev + (size_t)slot * kMaxPeers, my_flags,
                                              eng->gather_world, eng->gather_rank, epoch);
  eng->launches += 1;
  CK(cudaGetLastError());
  return MPC_OK;
}

int mpc_batch_gather_sync(mpc_batch_t* eng, void* cuda_stream) { return mpc_batch_gather_sync_slot(eng, 0, cuda_stream); }

int mpc_batch_set_gather_fused(mpc_batch_t* eng, int on) {
  if (!eng) return MPC_E_ARG;
  eng->gather_fused = on != 0;
  return MPC_OK;
}

// Copy-engine gather: this rank's [batch, 12] forces go into every rank's gather region `slot` by DMA over NVLink (no
// SM takes part, so nothing competes with the solve kernels of the batches in flight), then the device-side flag
// barrier of mpc_batch_gather_sync_slot tells the peers and waits for them.
int mpc_batch_gather_push_slot(mpc_batch_t* eng, int slot, const float* forces_dev, int batch, void* cuda_stream) {
  if (!eng || slot < 0 || slot >= kSlots || !forces_dev || batch < 0 || !eng->peer_flags_dev || eng->gather_world < 1 ||
      eng->rank_offset + batch > eng->gather_rows)
    return MPC_E_ARG;
  ON_DEVICE(eng);
  cudaStream_t st = (cudaStream_t)cuda_stream;
  const size_t bytes = (size_t)batch * 12 * sizeof(float);
  const size_t off = eng->gather_slot_bytes * (size_t)slot + (size_t)eng->rank_offset * 12 * sizeof(float);
  for (int q = 0; q < eng->gather_world; q++)
    if (eng->peers[q]) CK(cudaMemcpyAsync((char*)eng->peers[q] + off, forces_dev, bytes, cudaMemcpyDeviceToDevice, st));
  return mpc_batch_gather_sync_slot(eng, slot, st);
}

void* mpc_batch_gather_buffer(mpc_batch_t* eng) { return eng ? eng->gather_buf : nullptr; }
void* mpc_batch_gather_buffer_slot(mpc_batch_t* eng, int slot) {
  if (!eng || !eng->gather_buf || slot < 0 || slot >= kSlots) return nullptr;
  return (char*)eng->gather_buf + eng->gather_slot_bytes * slot;
}

int mpc_batch_set_sweep_variant(mpc_batch_t* eng, int variant) {
  if (!eng || variant < 0 || variant > 1) return MPC_E_ARG;
  eng->sweep = variant;
  return MPC_OK;
}
int mpc_batch_sweep_variant(const mpc_batch_t* eng) { return eng ? eng->sweep : -1; }

int mpc_batch_set_solver(mpc_batch_t* eng, int solver) {
  if (!eng || solver < 0 || solver > 1) return MPC_E_ARG;
  eng->solver = solver;
  return MPC_OK;
}
int mpc_batch_solver(const mpc_batch_t* eng) { return eng ? eng->solver : -1; }

int mpc_batch_warm_stride(void) { return mpc::kWarmStride; }

int mpc_batch_set_warm_start(mpc_batch_t* eng, int* cache_dev, const int* robot_ids_dev, int shift) {
  if (!eng || shift < 0) return MPC_E_ARG;
  eng->warm_cache = cache_dev;
  eng->warm_ids = cache_dev ? robot_ids_dev : nullptr;
  eng->warm_shift = shift;
  return MPC_OK;
}

int mpc_batch_set_max_iterations(mpc_batch_t* eng, int max_iter) {
  if (!eng || max_iter < 1) return MPC_E_ARG;
  eng->max_iter = max_iter;
  return MPC_OK;
}

int mpc_batch_set_phase_clock_buffer(mpc_batch_t* eng, long long* dev_buf) {
  if (!eng) return MPC_E_ARG;
  eng->phase_clk = dev_buf;
  return MPC_OK;
}

int mpc_batch_set_ctas_per_sm_limit(mpc_batch_t* eng, int limit) {
  if (!eng || limit < 0) return MPC_E_ARG;
  eng->ctas_per_sm_limit = limit;
  return MPC_OK;
}

int mpc_batch_set_timing(mpc_batch_t* eng, int enabled) {
  if (!eng) return MPC_E_ARG;
  if (enabled) {
    ON_DEVICE(eng);
    if (int rc = ensure_timing(eng)) return rc;
  }
  eng->timed = enabled != 0;
  eng->timed_class = -1;
  return MPC_OK;
}

int mpc_batch_set_timed_class(mpc_batch_t* eng, int idx) {
  if (!eng || idx < -1 || idx >= (int)eng->classes.size()) return MPC_E_ARG;
  ON_DEVICE(eng);
  if (int rc = ensure_timing(eng)) return rc;
  eng->timed = true;
  eng->timed_class = idx;
  return MPC_OK;
}

long mpc_batch_kernel_launches(const mpc_batch_t* eng) { return eng ? eng->launches : 0; }

float mpc_batch_last_solve_kernel_ms(mpc_batch_t* eng) {
  if (!eng || !eng->timed || eng->timed_class >= 0) return -1.f;
  float ms = -1.f;
  if (cudaEventSynchronize(eng->ev1) != cudaSuccess) return -1.f;
  if (cudaEventElapsedTime(&ms, eng->ev0, eng->ev1) != cudaSuccess) return -1.f;
  return ms;
}

float mpc_batch_last_class_kernel_ms(mpc_batch_t* eng, int idx) {
  if (!eng || !eng->timed || idx < 0 || idx >= (int)eng->classes.size() || eng->ring_pos == 0) return -1.f;
  const size_t slot = (size_t)((eng->ring_pos - 1) % kRing) * kMaxClasses + idx;
  float ms = -1.f;
  if (cudaEventSynchronize(eng->ring1[slot]) != cudaSuccess) return -1.f;
  if (cudaEventElapsedTime(&ms, eng->ring0[slot], eng->ring1[slot]) != cudaSuccess) return -1.f;
  return ms;
}

void mpc_batch_timing_mark(mpc_batch_t* eng) {
  if (eng) eng->ring_mark = eng->ring_pos;
}

int mpc_batch_timing_collect(mpc_batch_t* eng, int idx, float* mean_ms, int* n_solves) {
  if (!eng || !eng->timed || idx < 0 || idx >= (int)eng->classes.size() || !mean_ms) return MPC_E_ARG;
  long first = std::max(eng->ring_mark, eng->ring_pos - kRing);
  double sum = 0;
  int n = 0;
  for (long s = first; s < eng->ring_pos; s++) {
    const size_t slot = (size_t)(s % kRing) * kMaxClasses + idx;
    float ms = 0;
    CK(cudaEventSynchronize(eng->ring1[slot]));
    CK(cudaEventElapsedTime(&ms, eng->ring0[slot], eng->ring1[slot]));
    sum += ms;
    n++;
  }
  *mean_ms = n ? (float)(sum / n) : -1.f;
  if (n_solves) *n_solves = n;
  return MPC_OK;
}

int mpc_batch_host_buffers(mpc_batch_t* eng, int slot, void** records, float** forces, double** solution,
                           int32_t** status) {
  if (!eng || slot < 0 || slot >= kSlots) return MPC_E_ARG;
  {
    ON_DEVICE(eng);
    if (int rc = ensure_slot(eng, slot)) return rc;
  }
  mpc_batch::Slot& S = eng->s[slot];
  if (records) *records = S.rec_pin;
  if (forces) *forces = S.forces_pin;
  if (solution) *solution = S.sol_pin;
  if (status) *status = S.status_pin;
  return MPC_OK;
}

int mpc_batch_device_buffers(mpc_batch_t* eng, int slot, float** forces_dev, int32_t** status_dev) {
  if (!eng || slot < 0 || slot >= kSlots) return MPC_E_ARG;
  {
    ON_DEVICE(eng);
    if (int rc = ensure_slot(eng, slot)) return rc;
  }
  if (forces_dev) *forces_dev = eng->s[slot].forces_dev;
  if (status_dev) *status_dev = eng->s[slot].status_dev;
  return MPC_OK;
}

const char* mpc_batch_last_error(const mpc_batch_t* eng) { return eng ? eng->err.c_str() : ""; }
int mpc_batch_horizon(const mpc_batch_t* eng) { return eng ? eng->h : 0; }

int mpc_batch_num_classes(const mpc_batch_t* eng) { return eng ? (int)eng->classes.size() : 0; }
// info[6]: nv_cap, m_cap, threads, grid, shared bytes, 1 if the tile lives in shared memory
int mpc_batch_class_info(const mpc_batch_t* eng, int idx, int* info) {
  if (!eng || idx < 0 || idx >= (int)eng->classes.size() || !info) return MPC_E_ARG;
  const ClassCfg& c = eng->classes[idx];
  if (use_riccati(eng, c, false)) {  // what the production launches of this class use right now
    info[0] = c.nv_cap; info[1] = c.ric_m_cap; info[2] = 32; info[3] = c.ric_grid; info[4] = (int)c.ric_smem; info[5] = 1;
    return MPC_OK;
  }
  info[0] = c.nv_cap; info[1] = c.pipe ? c.pipe_m_cap : c.m_cap; info[2] = c.threads;
  info[3] = c.pipe ? c.pipe_grid : c.grid; info[4] = (int)(c.pipe ? c.pipe_smem : c.smem); info[5] = c.in_fast;
  return MPC_OK;
}

}  // extern "C"
