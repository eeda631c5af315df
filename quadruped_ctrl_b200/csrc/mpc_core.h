// Per-problem convex-MPC algorithm: record -> reduced condensed QP -> exact optimum.
//
// This header is the body of the sm_100a solve kernel (mpc_kernels.cu includes it
// with one CTA per problem).  It is written against a tiny execution context
// (thread id, thread count, barrier, block reductions) so that the very same
// source also compiles as single-thread host code for the logic tests in
// tests/emu/ (TEST-ONLY build; never linked into libquadruped_mpc_b200.so --
// the product has no CPU path).
//
// What it replaces, in the reference (/root/reference/src/MPC_Ctrl):
//   RobotState::set                      RobotState.cpp:9-43
//   quat_to_rpy, x_0, I_world            SolverMPC.cpp:257-267, 314-319
//   ct_ss_mats, cross_mat                SolverMPC.cpp:226-254
//   c2qp (expm + horizon stacking)       SolverMPC.cpp:87-125
//   weights / X_d / U_b / fmat           SolverMPC.cpp:335-378
//   qH, qg                               SolverMPC.cpp:395-399
//   swing-leg elimination                SolverMPC.cpp:431-525
//   qpOASES QProblem::init + scatter     SolverMPC.cpp:527-557
//
// How (B200-first, nothing of the reference's dense route is materialised):
//   * A_c is nilpotent (A^3 = 0), so exp(k dt A) = I + k dt A + (k dt)^2/2 A^2 and
//     Phi_k := A_d^k B_d = C0 + k C1 + k^2 C2 exactly.  B_qp (13h x 12h), the dense
//     S (13h x 13h) and the 25x25 matrix exponential never exist.
//   * H[(i,.),(j,.)] = 2 sum_{a,b} s_ab(i,j) C_a' Q C_b + 2 alpha I with integer
//     sums s_ab(i,j) = sum_{r>=max(i,j)} (r-i)^a (r-j)^b: 9 FMAs per entry instead of
//     a 13h-long dot product; only stance (step,leg) columns are ever formed.
//   * The reduced H is inverted in place (symmetric sweep) and the QP is solved by
//     the Goldfarb-Idnani dual active-set method on the explicit inverse: every
//     constraint of the friction pyramid touches at most two variables, so the
//     Schur complement entries are O(1) table look-ups and one working-set change
//     costs O(nv*m) fully parallel work.
//   * All arithmetic is fp64 (the reference assembles in fp32 and solves in fp64;
//     fp64 assembly lands inside the reference's own fp32 rounding cloud).
#ifndef QUADRUPED_MPC_CORE_H
#define QUADRUPED_MPC_CORE_H

#include <math.h>
#include <stdint.h>

#include "../../include/mpc_batch.h"

#if defined(__CUDACC__)
#define MPC_HD __device__ __forceinline__
#define MPC_HDN __device__ __noinline__
#else
#define MPC_HD inline
#define MPC_HDN inline
#endif

namespace mpc {

// extra status codes (additive to include/mpc_batch.h) used between kernels
enum { STATUS_RETRY_BIG = 0x40 };  // working set outgrew the fast-memory tile: re-queued to the big class

// ---------------------------------------------------------------------------
// Workspace layout.  `fast` is shared memory on the device; Hm and T live in fast
// memory when they fit and in a per-CTA global (L2-resident) slab otherwise.
// ---------------------------------------------------------------------------
struct Layout {
  int h;        // horizon
  int nv_cap;   // largest reduced variable count this workspace can hold
  int ck_len;   // doubles reserved for the pivot-column broadcast buffer(s)
  int m_cap;    // largest working set (active constraints) it can hold
  int ld;       // leading dimension of Hm (odd -> conflict-free column walks)
  int ldT;      // leading dimension of T
  int big_in_fast;  // 1: Hm and T are carved from `fast`; 0: from the global slab
  int pipe;         // 1: two problems in flight per CTA (see make_layout): disjoint assembly / active-set regions
  int off_scal2, off_ints2, off_gi, off_mom;  // pipe: second per-problem set, active-set region, moment sums
  int off_wr;       // wrench-space class: its scratch vectors (0: none)
  int t_in_slab;    // wrench-space overflow class: T lives in the global slab although Hm is in fast memory
  // byte offsets into `fast`
  int off_scal, off_g, off_x, off_ints, off_union, off_red, off_Hm, off_T;
  int fast_bytes;
  // byte offsets into the global slab (when !big_in_fast)
  size_t slab_Hm, slab_T, slab_bytes;
};

struct Scalars {
  int nv, ns, m, status, iters, p, kdrop, full, done;
  int next_item;  // hand-over slot of the kernels' dynamic problem queue (thread 0 -> CTA); not part of the algorithm
  double sp, t1, t2, t, up, vnp, znp, dreg;
};

constexpr int kSweepGroup = 16;  // pivots per pass of the generic sweep (see invert_spd)
constexpr int kAsmDoubles(int h) { return 3 * 156 + 6 * 144 + 39 + 12 * h + 5 * h; }
constexpr int kRedDoubles = 40;
// DMMA grouped sweep (invert_spd_mma): its panel buffers live in the Hm region, which must be large enough
constexpr int kMmaGroup = 8;                                              // pivots per group (= block size)
constexpr int mma_panel_ld(int nvp) { return nvp + 4; }                   // = 4 mod 16: fragment loads conflict-free
constexpr int mma_panel_doubles(int nvp) { return 3 * kMmaGroup * mma_panel_ld(nvp) + kMmaGroup * kMmaGroup + kMmaGroup; }  // G, F, U, T, 1/d

inline int align_up(int x, int a) { return (x + a - 1) / a * a; }

// The same function sizes the launch (host) and carves the pointers (device).
// packed: Hm as a packed lower triangle (the register-resident classes; the generic in-place sweep of the
// catch-all class and of the host emulation needs full storage)
// pipe: the smallest class keeps two problems in flight per CTA -- one warp runs the active set of problem n while
// the other warps assemble problem n+1 (phases P0..P9) -- so the assembly scratch and the active-set scratch are
// disjoint instead of a union, the moment sums get their own place (H^{-1} of problem n is still being read) and the
// small per-problem state (scalars, stance lists) exists twice.
// wrench: the wrench-space class -- Hm holds a 6h x 6h matrix (not nv_cap x nv_cap) and the workspace carries the
// scratch vectors of wr_apply().
inline Layout make_layout(int h, int nv_cap, int m_cap, int big_in_fast, int npad = 0, int packed = -1, int pipe = 0,
                          int wrench = 0) {
  if (packed < 0) packed = 0;
  Layout L;
  L.pipe = pipe;
  L.off_wr = 0;
  L.off_scal2 = L.off_ints2 = L.off_gi = L.off_mom = 0;
  L.h = h;
  L.nv_cap = nv_cap;
  // register-resident inversion: two (npad + 2)-long buffers (double-buffered pivot row + 1/pivot)
  L.ck_len = npad > 0 ? 2 * (npad + 2) : 2 * kSweepGroup * nv_cap;  // generic sweep: K snapshots + K scaled copies
  L.m_cap = m_cap;
  L.ld = packed ? -1 : (nv_cap | 1);  // packed lower triangle (see hix()) or full storage
  L.ldT = m_cap | 1;
  L.big_in_fast = big_in_fast;
  int o = 0;
  L.off_scal = o;
  o += (int)((sizeof(Scalars) + 15) / 16 * 16);
  L.off_g = o;
  o += 8 * nv_cap;
  L.off_x = o;
  o += 8 * nv_cap;
  // ints: stance[4h] (k = step*4+leg of every stance pair), posk[4h], amask[4h], W / Wia / Wiz [m_cap+1]
  L.off_ints = o;
  o += 4 * (3 * 4 * h + 3 * (m_cap + 1));
  o = (o + 15) / 16 * 16;
  L.off_union = o;
  int gi = L.ck_len + 1 + 4 * h + (m_cap + 1) * 6;  // ck (16-byte aligned), ub, Wca, Wcz, w, r, u, tcol
  int un = kAsmDoubles(h);
  int t_doubles = m_cap * L.ldT;
  const int hm_n = wrench ? 6 * h : nv_cap;  // dimension of the matrix held in Hm
  if (wrench) L.ld = packed ? -1 : (hm_n | 1);
  int hm_doubles = packed ? hm_n * (hm_n + 1) / 2 : hm_n * L.ld;
  if (hm_doubles < 3 * 12 * h) hm_doubles = 3 * 12 * h;  // the assembly parks its moment sums there
  if (npad > 0 && hm_doubles < mma_panel_doubles(npad)) hm_doubles = mma_panel_doubles(npad);  // panel of the DMMA sweep
  const bool t_in_slab = wrench == 2;  // wrench-space overflow class: M in fast memory, the working-set matrix T in the slab
  if (big_in_fast && !t_in_slab) gi += t_doubles;
  if (pipe) {
    o += 8 * un;                 // assembly scratch
    L.off_mom = o;
    o += 8 * 3 * 12 * h;
    L.off_gi = o;
    o += 8 * (gi + 1);           // active-set scratch (T first)
    L.off_scal2 = o;
    o += (int)((sizeof(Scalars) + 15) / 16 * 16);
    L.off_ints2 = o;
    o += 4 * 3 * 4 * h;
    o = (o + 15) / 16 * 16;
  } else {
    if (gi > un) un = gi;
    o += 8 * un;
  }
  if (wrench) {  // wy [6h], wt [2][6h], wv, wz [nv_cap], rcat [24h], rleg [12], dblk [36*h]
    L.off_wr = o;
    o += 8 * (3 * 6 * h + 2 * nv_cap + 24 * h + 12 + 36 * h);
  }
  L.off_red = o;
  o += 8 * kRedDoubles;
  o = (o + 15) / 16 * 16;  // the DMMA sweep stores 16 bytes at a time into its panel (in the Hm region)
  L.off_Hm = o;
  L.off_T = pipe ? L.off_gi : L.off_union;  // T sits at the start of the active-set scratch
  if (big_in_fast) o += 8 * hm_doubles;
  L.fast_bytes = (o + 15) / 16 * 16;
  L.slab_Hm = 0;
  L.slab_T = (size_t)hm_doubles * 8;
  L.slab_bytes = big_in_fast ? 0 : ((size_t)(hm_doubles + t_doubles) * 8 + 255) / 256 * 256;
  if (t_in_slab) {
    L.slab_T = 0;
    L.slab_bytes = ((size_t)t_doubles * 8 + 255) / 256 * 256;
  }
  L.t_in_slab = t_in_slab ? 1 : 0;
  return L;
}

struct Work {
  Scalars* sc;
  double *g, *x, *Hm, *T;
  int *stance, *posk, *amask, *W, *Wia, *Wiz;
  // assembly view of the union
  double *C, *M, *xs, *qe, *psum, *mom;  // mom: [3][h][12] moment sums (parked in Hm unless the layout is piped)
  // active-set view of the union
  double *ck, *ub, *Wca, *Wcz, *w, *r, *u, *tcol;
  double* red;
  // wrench-space class: scratch of wr_apply (wy, wt: 6h; wv, wz: nv), per-row coefficients by catalogue index (rcat),
  // foot positions relative to the COM (rleg[leg][axis]), the 6x6 blocks D_s = sum_legs G G' (dblk), 1/(2 alpha)
  double *wy, *wt, *wv, *wz, *rcat, *rleg, *dblk;
  double i2a;
  int ld, ldT, nv_cap, m_cap, h;
  long long* clk;  // optional per-problem clock stamps (profiling aid), slots 8..23
};

#if defined(__CUDA_ARCH__)
// k.clk is null at compile time in the production instantiation of the kernel: the stamps then vanish
#define MPC_STAMP(k, cx, slot) do { if ((k).clk && (cx).tid == 0) (k).clk[slot] = clock64(); } while (0)
#else
#define MPC_STAMP(k, cx, slot) do { } while (0)
#endif

// set: which of the two per-problem sets (scalars, stance / posk / amask) of a piped layout; 0 otherwise
MPC_HD Work carve(const Layout& L, char* fast, char* slab, int set = 0) {
  Work k;
  k.sc = (Scalars*)(fast + (set ? L.off_scal2 : L.off_scal));
  k.g = (double*)(fast + L.off_g);
  k.x = (double*)(fast + L.off_x);
  int* ip = (int*)(fast + L.off_ints);
  int* ips = set ? (int*)(fast + L.off_ints2) : ip;
  k.stance = ips;
  k.posk = ips + 4 * L.h;
  k.amask = ips + 8 * L.h;
  k.W = ip + 12 * L.h;
  k.Wia = k.W + (L.m_cap + 1);
  k.Wiz = k.Wia + (L.m_cap + 1);
  double* un = (double*)(fast + L.off_union);
  k.C = un;
  k.M = un + 3 * 156;
  k.xs = k.M + 6 * 144;
  k.qe = k.xs + 39;
  k.psum = k.qe + 12 * L.h;
  double* gi = L.pipe ? (double*)(fast + L.off_gi) : un;
  if (L.big_in_fast && L.t_in_slab) {
    k.T = (double*)(slab + L.slab_T);
    k.Hm = (double*)(fast + L.off_Hm);
  } else if (L.big_in_fast) {
    k.T = gi;
    gi += L.m_cap * L.ldT;
    k.Hm = (double*)(fast + L.off_Hm);
  } else {
    k.Hm = (double*)(slab + L.slab_Hm);
    k.T = (double*)(slab + L.slab_T);
  }
  k.ck = gi;
  if (((uintptr_t)k.ck & 15) != 0) k.ck += 1;  // the register-resident inversion moves it 16 bytes at a time
  k.ub = k.ck + L.ck_len;
  k.Wca = k.ub + 4 * L.h;
  k.Wcz = k.Wca + (L.m_cap + 1);
  k.w = k.Wcz + (L.m_cap + 1);
  k.r = k.w + (L.m_cap + 1);
  k.u = k.r + (L.m_cap + 1);
  k.tcol = k.u + (L.m_cap + 1);
  k.red = (double*)(fast + L.off_red);
  k.wy = k.wt = k.wv = k.wz = k.rcat = k.rleg = k.dblk = nullptr;
  k.i2a = 0.0;
  if (L.off_wr) {
    k.wy = (double*)(fast + L.off_wr);
    k.wt = k.wy + 6 * L.h;
    k.wv = k.wt + 12 * L.h;
    k.wz = k.wv + L.nv_cap;
    k.rcat = k.wz + L.nv_cap;
    k.rleg = k.rcat + 24 * L.h;
    k.dblk = k.rleg + 12;
  }
  k.mom = L.pipe ? (double*)(fast + L.off_mom) : k.Hm;
  k.ld = L.ld;
  k.ldT = L.ldT;
  k.nv_cap = L.nv_cap;
  k.m_cap = L.m_cap;
  k.h = L.h;
  k.clk = nullptr;
  return k;
}

// ---------------------------------------------------------------------------
// Execution context.  Device: one CTA.  Host emulation: one thread.
// ---------------------------------------------------------------------------
// kPacked (compile time): Hm is a packed lower triangle (hix() below) instead of full row-major storage.
#if defined(__CUDACC__)
// UNR: unroll factor of the active set's inner loops over the working set.  1 where H^{-1} and T sit in shared memory
// (the kernel is instruction-fetch heavy); 4 for the catch-all class, whose H^{-1} / T live in an L2-resident slab and
// whose loops are chains of dependent-latency loads -- unrolling issues the loads of four steps before the first
// FMA (same summation order, so the same bits).
// kWrench (compile time): H^{-1} is not stored; the active set works through the rank-6h structure of the Hessian
// (see "wrench-space class" below).
template <bool PK, int UNR = 1, bool WR = false>
struct CtaT {
  static constexpr bool kWrench = WR;
  static constexpr bool kOneWarp = false;
  static constexpr bool kPacked = PK;
  static constexpr int kUnroll = UNR;
  int tid, nt;
  __device__ __forceinline__ void sync() const { __syncthreads(); }
};
// One warp of the CTA working alone (the latency-bound active-set stage): barriers are __syncwarp and
// reductions are shuffles only.
template <bool PK>
struct WarpT {
  static constexpr bool kWrench = false;
  static constexpr bool kOneWarp = true;
  static constexpr bool kPacked = PK;
  static constexpr int kUnroll = 1;
  int tid, nt;  // lane, 32
  __device__ __forceinline__ void sync() const { __syncwarp(); }
};
// NTP threads of the CTA (a whole number of warps) working as a group with hardware barrier BAR (not 0, which is
// __syncthreads): the assembly of the next problem while another group (or one warp) runs the active set of the
// current one.
template <int BAR, int NTP, bool PK = false>
struct PartT {
  static constexpr bool kWrench = false;
  static constexpr bool kOneWarp = false;
  static constexpr bool kPacked = PK;
  static constexpr int kUnroll = 1;
  int tid, nt;  // 0..NTP-1, NTP
  __device__ __forceinline__ void sync() const { asm volatile("bar.sync %0, %1;" ::"n"(BAR), "n"(NTP) : "memory"); }
};
using Cta = CtaT<false>;
using Warp = WarpT<false>;
#endif
template <bool PK, bool WR = false>
struct OneThreadT {
  static constexpr bool kWrench = WR;
  static constexpr bool kOneWarp = false;
  static constexpr bool kPacked = PK;
  static constexpr int kUnroll = 1;
  int tid, nt;
  inline void sync() const {}
};
using OneThread = OneThreadT<false>;

// Strided loops are deliberately NOT unrolled: the kernel is instruction-cache bound (32 KB L1.5 I-cache),
// and these loops run a handful of iterations per thread.
#define MPC_FOR(i, n) _Pragma("unroll 1") for (int i = cx.tid; i < (n); i += cx.nt)
#define MPC_ONE if (cx.tid == 0)

// Block-wide argmin of (val, idx) pairs; every thread passes its local best and
// gets the global best back.  Ties resolve to the smaller idx (deterministic).
#if defined(__CUDACC__)
// out of line on purpose (called from many places; see the I-cache note above)
__device__ __noinline__ void warp_argmin(double& val, int& idx) {
#pragma unroll 1
  for (int o = 16; o > 0; o >>= 1) {
    double v2 = __shfl_xor_sync(0xffffffffu, val, o);
    int i2 = __shfl_xor_sync(0xffffffffu, idx, o);
    if (v2 < val || (v2 == val && i2 < idx)) { val = v2; idx = i2; }
  }
}
__device__ __noinline__ double warp_sum(double val) {
#pragma unroll 1
  for (int o = 16; o > 0; o >>= 1) val += __shfl_xor_sync(0xffffffffu, val, o);
  return val;
}
__device__ __noinline__ double atan2_shared(double y, double x) { return atan2(y, x); }
#define MPC_ATAN2 atan2_shared
#else
#define MPC_ATAN2 atan2
#endif

template <class Cx>
MPC_HD void block_argmin(const Cx& cx, double* red, double& val, int& idx) {
#if defined(__CUDA_ARCH__)
  warp_argmin(val, idx);
  if (Cx::kOneWarp) return;
  const int nw = (cx.nt + 31) >> 5;
  int* redi = (int*)(red + 16);
  cx.sync();  // scratch may still be read from a previous reduction
  if ((cx.tid & 31) == 0) { red[cx.tid >> 5] = val; redi[cx.tid >> 5] = idx; }
  cx.sync();
  val = red[0];
  idx = redi[0];
  for (int w = 1; w < nw; w++) {
    double v2 = red[w];
    int i2 = redi[w];
    if (v2 < val || (v2 == val && i2 < idx)) { val = v2; idx = i2; }
  }
#else
  (void)cx; (void)red; (void)val; (void)idx;
#endif
}

template <class Cx>
MPC_HD double block_sum(const Cx& cx, double* red, double val) {
#if defined(__CUDA_ARCH__)
  val = warp_sum(val);
  if (Cx::kOneWarp) return val;
  const int nw = (cx.nt + 31) >> 5;
  cx.sync();
  if ((cx.tid & 31) == 0) red[cx.tid >> 5] = val;
  cx.sync();
  double s = 0;
  for (int w = 0; w < nw; w++) s += red[w];
  return s;
#else
  (void)cx; (void)red;
  return val;
#endif
}

// ---------------------------------------------------------------------------
// Constraint catalogue.  Every stance pair j (reduced variables 3j..3j+2 =
// fx,fy,fz) carries six one-sided rows  n.x >= b  (SolverMPC.cpp:361-378 f_block,
// lower bound 0 :428-429, upper bounds U_b :349-358; the four 5e10 upper bounds
// on the cone rows are not constraints):
//   type 0:  fx/mu + fz >= 0     type 1: -fx/mu + fz >= 0
//   type 2:  fy/mu + fz >= 0     type 3: -fy/mu + fz >= 0
//   type 4:  fz >= 0             type 5: -fz >= -gait*f_max
// A row is (ia, ca, iz=3j+2, cz): n = ca*e_ia + cz*e_iz (ia == iz, ca = 0 for 4,5).
// ---------------------------------------------------------------------------
// Index of entry (i, j) of the symmetric matrix Hm: full row-major storage with leading dimension ld > 0, or
// (ld < 0) a packed lower triangle, (i, j) and (j, i) being one element -- half the footprint for index arithmetic
// on every access.  Hot code takes the choice at compile time (Cx::kPacked).
MPC_HD int tri_index(int i, int j) { return i >= j ? i * (i + 1) / 2 + j : j * (j + 1) / 2 + i; }
template <bool PK>
MPC_HD int hixT(int ld, int i, int j) { return PK ? tri_index(i, j) : i * ld + j; }
MPC_HD int hix(int ld, int i, int j) { return ld > 0 ? i * ld + j : tri_index(i, j); }  // run-time form (cold paths)

struct Row {
  int ia, iz;
  double ca, cz;
};
MPC_HD Row make_row(int c, double mu_inv) {
  const int j = c / 6, t = c - 6 * j;
  Row r;
  r.iz = 3 * j + 2;
  r.cz = (t == 5) ? -1.0 : 1.0;
  if (t < 2) { r.ia = 3 * j; r.ca = (t == 0) ? mu_inv : -mu_inv; }
  else if (t < 4) { r.ia = 3 * j + 1; r.ca = (t == 2) ? mu_inv : -mu_inv; }
  else { r.ia = r.iz; r.ca = 0.0; }
  return r;
}
// n_a' Minv n_b for two catalogue rows: at most four look-ups
template <bool PK>
MPC_HD double row_minv_row(const double* Hm, int ld, const Row& a, const Row& b) {
  double s = a.cz * b.cz * Hm[hixT<PK>(ld, a.iz, b.iz)];
  if (a.ca != 0.0) s += a.ca * b.cz * Hm[hixT<PK>(ld, a.ia, b.iz)];
  if (b.ca != 0.0) s += a.cz * b.ca * Hm[hixT<PK>(ld, a.iz, b.ia)];
  if (a.ca != 0.0 && b.ca != 0.0) s += a.ca * b.ca * Hm[hixT<PK>(ld, a.ia, b.ia)];
  return s;
}

MPC_HD bool finite_f(float v) { return v == v && fabsf(v) <= 3.0e38f; }

// ---------------------------------------------------------------------------
// Stage 1: dynamics, discretisation polynomial, QP assembly (reduced, fp64).
// Leaves nv, ns, stance[], posk[], Hm (= reduced qH), g (= reduced qg) behind.
// ---------------------------------------------------------------------------
// (A X)[i][j] and (A^2 X)[i][j] for the continuous-time A of ct_ss_mats (SolverMPC.cpp:237-244):
// A[0:3,6:9] = R_yaw', A[3,9] = A[4,10] = A[5,11] = 1, A[11,9] = x_drag, A[11,12] = 1, zero elsewhere,
// hence A^2 = (x_drag e9' + e12') on row 5 only and A^3 = 0.  X has leading dimension ldx.
MPC_HD double apply_A(const double* X, int ldx, int i, int j, double yc, double ys, double xd) {
  switch (i) {
    case 0: return yc * X[6 * ldx + j] + ys * X[7 * ldx + j];
    case 1: return -ys * X[6 * ldx + j] + yc * X[7 * ldx + j];
    case 2: return X[8 * ldx + j];
    case 3: return X[9 * ldx + j];
    case 4: return X[10 * ldx + j];
    case 5: return X[11 * ldx + j];
    case 11: return xd * X[9 * ldx + j] + X[12 * ldx + j];
    default: return 0.0;
  }
}
MPC_HD double apply_A2(const double* X, int ldx, int i, int j, double xd) {
  return i == 5 ? xd * X[9 * ldx + j] + X[12 * ldx + j] : 0.0;
}

// assemble = assemble_front (P0..P9: everything up to the gradient and the M tables; touches neither Hm -- when the
// moment sums have a place of their own -- nor anything the active set uses) + assemble_H (P11: the H blocks).
template <class Cx>
MPC_HD void assemble_front(const Cx& cx, const float* rec, const unsigned char* gait, const Work& k) {
  const int h = k.h;
  Scalars* sc = k.sc;
  double* B = k.M;          // 13x12   (the M region is free until the C_a are final)
  double* scr = B + 156;    // cos(yaw), sin(yaw)
  double* C0 = k.C;
  double* C1 = C0 + 156;
  double* C2 = C1 + 156;
  double* x0 = k.xs;        // 13
  double* Ax0 = x0 + 13;
  double* A2x0 = Ax0 + 13;
  double* mom = k.mom;      // [3][h][12] moments of the tracking error (in Hm, free until the H blocks are written,
                            // or in a place of its own when two problems are in flight)
  int* flag = k.amask;      // 0/1 stance flags while the stance list is built (amask[] proper is set up later)
  const float fmax = rec[MPC_REC_FMAX];

  // ---- P0: reset, stance flags (SolverMPC.cpp:441-469: U_b(5k+4) = gait[k]*f_max "near zero" => eliminated) ----
  MPC_ONE {
    sc->status = MPC_STATUS_OPTIMAL;
    sc->m = 0;
    sc->iters = 0;
  }
  MPC_FOR(kk, 4 * h) {
    const float ub = (float)gait[kk] * fmax;
    flag[kk] = ((double)ub < 0.01 && (double)ub > -0.01) ? 0 : 1;
  }
  MPC_FOR(i, 156) B[i] = 0.0;
  cx.sync();
  // ---- P1: input check, stance list, the four transcendental groups on four different warps ----
  MPC_FOR(i, MPC_REC_TRAJ + 12 * h)
    if (!finite_f(rec[i])) sc->status = MPC_STATUS_BAD_INPUT;  // same value from every writer
  MPC_ONE {
    if (!(rec[MPC_REC_MU] > 0.f) || !(rec[MPC_REC_MASS] > 0.f) || !(rec[MPC_REC_DT] > 0.f) ||
        !(rec[MPC_REC_IBODY] > 0.f) || !(rec[MPC_REC_IBODY + 1] > 0.f) || !(rec[MPC_REC_IBODY + 2] > 0.f) ||
        !(fmax >= 0.f))
      sc->status = MPC_STATUS_BAD_INPUT;
  }
  MPC_FOR(kk, 4 * h) {
    int pos = 0;
#pragma unroll 4
    for (int q = 0; q < kk; q++) pos += flag[q];
    if (flag[kk]) { k.stance[pos] = kk; k.posk[kk] = pos; }
    else k.posk[kk] = -1;
    if (kk == 4 * h - 1) { sc->ns = pos + flag[kk]; sc->nv = 3 * (pos + flag[kk]); }
  }
  {
    // quat_to_rpy (SolverMPC.cpp:257-267), q = (w,x,y,z); x_0 = [rpy(2), rpy(1), rpy(0), ...] (:318)
    const double qw = rec[MPC_REC_Q], qx = rec[MPC_REC_Q + 1], qy = rec[MPC_REC_Q + 2], qz = rec[MPC_REC_Q + 3];
    // the transcendental groups on different warps: sincos on thread 0, the two atan2 on neighbouring lanes of the
    // second warp (one instruction stream), asin on the third (three warps suffice: the piped kernel assembles on 96
    // threads); a single thread runs them all
    const int w1 = cx.nt >= 96 ? 32 : 0, w2 = cx.nt >= 96 ? 64 : 0, l1 = cx.nt >= 96 ? 1 : 0;
    if (cx.tid == 0) {
      const double yaw = (double)rec[MPC_REC_YAW];
      sincos(yaw, &scr[1], &scr[0]);  // one range reduction for both
    }
    if (cx.tid == w1) x0[2] = MPC_ATAN2(2. * (qx * qy + qw * qz), qw * qw + qx * qx - qy * qy - qz * qz);
    if (cx.tid == w2) {
      double as = -2. * (qx * qz - qw * qy);
      if (!(as < .99999)) as = .99999;
      x0[1] = asin(as);
    }
    if (cx.tid == w1 + l1) x0[0] = MPC_ATAN2(2. * (qy * qz + qw * qx), qw * qw - qx * qx - qy * qy + qz * qz);
  }
  cx.sync();
  MPC_STAMP(k, cx, 8);
  // ---- P2: x_0 tail, B_c per leg (ct_ss_mats, SolverMPC.cpp:235-254; cross_mat :226-233) ----
  MPC_ONE {
    if (sc->status == MPC_STATUS_OPTIMAL && sc->ns == 0) sc->status = MPC_STATUS_NO_STANCE;
    for (int i = 0; i < 3; i++) {
      x0[3 + i] = rec[MPC_REC_P + i];
      x0[6 + i] = rec[MPC_REC_W + i];
      x0[9 + i] = rec[MPC_REC_V + i];
    }
    x0[12] = (double)-9.8f;
  }
  const double yc = scr[0], ys = scr[1];
  const double xd = (double)rec[MPC_REC_XDRAG];
  MPC_FOR(b, 4) {
    const double R[3][3] = {{yc, -ys, 0}, {ys, yc, 0}, {0, 0, 1}};  // RobotState.cpp:33-35
    const double Ib[3] = {(double)rec[MPC_REC_IBODY], (double)rec[MPC_REC_IBODY + 1], (double)rec[MPC_REC_IBODY + 2]};
    // I_world = R I_body R' and its inverse (SolverMPC.cpp:319, 247)
    double Iw[3][3];
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) {
        double acc = 0;
        for (int q = 0; q < 3; q++) acc += (R[i][q] * Ib[q]) * R[j][q];
        Iw[i][j] = acc;
      }
    const double det = Iw[0][0] * (Iw[1][1] * Iw[2][2] - Iw[1][2] * Iw[2][1]) -
                       Iw[0][1] * (Iw[1][0] * Iw[2][2] - Iw[1][2] * Iw[2][0]) +
                       Iw[0][2] * (Iw[1][0] * Iw[2][1] - Iw[1][1] * Iw[2][0]);
    const double rdet = 1.0 / det;  // one division; the adjugate entries are scaled by it
    double Ii[3][3];
    Ii[0][0] = (Iw[1][1] * Iw[2][2] - Iw[1][2] * Iw[2][1]) * rdet;
    Ii[0][1] = (Iw[0][2] * Iw[2][1] - Iw[0][1] * Iw[2][2]) * rdet;
    Ii[0][2] = (Iw[0][1] * Iw[1][2] - Iw[0][2] * Iw[1][1]) * rdet;
    Ii[1][0] = (Iw[1][2] * Iw[2][0] - Iw[1][0] * Iw[2][2]) * rdet;
    Ii[1][1] = (Iw[0][0] * Iw[2][2] - Iw[0][2] * Iw[2][0]) * rdet;
    Ii[1][2] = (Iw[0][2] * Iw[1][0] - Iw[0][0] * Iw[1][2]) * rdet;
    Ii[2][0] = (Iw[1][0] * Iw[2][1] - Iw[1][1] * Iw[2][0]) * rdet;
    Ii[2][1] = (Iw[0][1] * Iw[2][0] - Iw[0][0] * Iw[2][1]) * rdet;
    Ii[2][2] = (Iw[0][0] * Iw[1][1] - Iw[0][1] * Iw[1][0]) * rdet;
    const double minv = 1.0 / (double)rec[MPC_REC_MASS];
    const double rx = rec[MPC_REC_R + b], ry = rec[MPC_REC_R + 4 + b], rz = rec[MPC_REC_R + 8 + b];
    const double cm[3][3] = {{0, -rz, ry}, {rz, 0, -rx}, {-ry, rx, 0}};
    for (int i = 0; i < 3; i++) {
      for (int j = 0; j < 3; j++) {
        double acc = 0;
        for (int q = 0; q < 3; q++) acc += Ii[i][q] * cm[q][j];
        B[(6 + i) * 12 + b * 3 + j] = acc;
      }
      B[(9 + i) * 12 + b * 3 + i] = minv;
    }
  }
  cx.sync();
  MPC_STAMP(k, cx, 9);
  if (sc->status != MPC_STATUS_OPTIMAL) return;
  const double dt = (double)rec[MPC_REC_DT];
  // ---- P3: exact discretisation B_d = dt B + dt^2/2 AB + dt^3/6 A^2 B (c2qp, SolverMPC.cpp:87-101), straight
  //          from B_c (A B and A^2 B have at most three terms per entry), and A x0, A^2 x0 ----
  MPC_FOR(e, 156) {
    const int i = e / 12, j = e - 12 * i;
    C0[e] = dt * B[e] + (dt * dt / 2.0) * apply_A(B, 12, i, j, yc, ys, xd) +
            (dt * dt * dt / 6.0) * apply_A2(B, 12, i, j, xd);
  }
  MPC_FOR(e, 13) {
    Ax0[e] = apply_A(x0, 1, e, 0, yc, ys, xd);
    A2x0[e] = apply_A2(x0, 1, e, 0, xd);
  }
  cx.sync();
  // ---- P7: Phi_k = A_d^k B_d = C0 + k C1 + k^2 C2 with C1 = dt A B_d, C2 = dt^2/2 A^2 B_d;
  //          weighted tracking error q_e[r] = Q (A_d^{r+1} x0 - x_d[r]) (SolverMPC.cpp:335-347,399) ----
  MPC_FOR(e, 156) {
    const int i = e / 12, j = e - 12 * i;
    C1[e] = dt * apply_A(C0, 12, i, j, yc, ys, xd);
    C2[e] = (dt * dt / 2.0) * apply_A2(C0, 12, i, j, xd);
  }
  MPC_FOR(e, 12 * h) {
    const int r = e / 12, i = e - 12 * r;
    const double tt = (double)(r + 1) * dt;
    const double xr = x0[i] + tt * Ax0[i] + (tt * tt / 2.0) * A2x0[i];
    k.qe[e] = (double)rec[MPC_REC_WEIGHTS + i] * (xr - (double)rec[MPC_REC_TRAJ + e]);
  }
  cx.sync();
  MPC_STAMP(k, cx, 10);
  // ---- P8: moments mom[a][j][row] = sum_{r>=j} (r-j)^a q_e[r][row] ----
  const int na = (xd != 0.0) ? 3 : 2;  // without drag C2 == 0 and every k^2 term drops out
  MPC_FOR(e, na * 12 * h) {
    int a = 0, jr = e;  // e = a*12h + jr without a division by the run-time 12h
    if (jr >= 12 * h) { jr -= 12 * h; a = 1; }
    if (jr >= 12 * h) { jr -= 12 * h; a = 2; }
    const int j = jr / 12, row = jr - 12 * j;
    double acc = 0;
#pragma unroll 2
    for (int r = j; r < h; r++) {
      const double kd = (double)(r - j);
      const double w = a == 0 ? 1.0 : (a == 1 ? kd : kd * kd);
      acc += w * k.qe[12 * r + row];
    }
    mom[e] = acc;
  }
  // power sums P_e(n) = sum_{q=0..n} q^e, e = 0..4, n = 0..h-1: exact integers (< 2^53 for h <= 36)
  MPC_FOR(n, h) {  // closed forms in integer arithmetic (n <= 35: every product below fits in 32 bits)
    const int n1 = n * (n + 1);
    const int p1 = n1 / 2;
    const int p2 = n1 * (2 * n + 1) / 6;
    const int p4 = (n1 * (2 * n + 1)) * (3 * n * n + 3 * n - 1) / 30;  // <= 35*36*71*3779 = 338,069,340
    k.psum[5 * n + 0] = (double)(n + 1);
    k.psum[5 * n + 1] = (double)p1;
    k.psum[5 * n + 2] = (double)p2;
    k.psum[5 * n + 3] = (double)p1 * (double)p1;
    k.psum[5 * n + 4] = (double)p4;
  }
  cx.sync();
  // ---- P9: reduced gradient g_v = 2 sum_a C_a[:,c]' mom[a][j] (SolverMPC.cpp:399), and the tables
  //          M_ab = C_a' Q C_b (12x12).  Row supports: C0 rows 0..11, C1 rows {0..5,11}, C2 row {5}.
  //          M overlays B,t1,t2,scr, which are dead: the barrier above is the last point anything reads them. ----
  const int nv = sc->nv;
  const unsigned rowmask[3] = {0xFFFu, 0x83Fu, 0x020u};
  MPC_FOR(v, nv) {
    const int sidx = v / 3, ax = v - 3 * sidx;
    const int kk = k.stance[sidx], j = kk >> 2, c = (kk & 3) * 3 + ax;
    double acc0 = 0, acc1 = 0;  // two chains: the sum is latency-bound
    for (int a = 0; a < na; a++) {
      const double* Ca = C0 + 156 * a;
      const double* ma = mom + (a * h + j) * 12;
      for (int row = 0; row < 12; row += 2) {
        if ((rowmask[a] >> row) & 1u) acc0 += Ca[row * 12 + c] * ma[row];
        if ((rowmask[a] >> (row + 1)) & 1u) acc1 += Ca[(row + 1) * 12 + c] * ma[row + 1];
      }
    }
    k.g[v] = 2.0 * (acc0 + acc1);
  }
  // six tables M_ab, a <= b (M_ba = M_ab'), of which only those with b < na are needed; one thread per (table,
  // row): twelve independent accumulators, each weighted C_a entry loaded once.  (One thread per ENTRY spreads the
  // work over the whole CTA and was 2 % faster in the one-problem kernel, but it is the longer chain, and in the
  // piped kernel this phase sits on the critical path beside the other problem's active set: -3 %.)
  MPC_FOR(e, 6 * 12) {
    const int tb = e / 12, i = e - 12 * tb;
    const int a = tb < 3 ? 0 : (tb < 5 ? 1 : 2), b = tb < 3 ? tb : (tb < 5 ? tb - 2 : 2);
    if (b < na) {  // a <= b
      const double* Ca = C0 + 156 * a;
      const double* Cb = C0 + 156 * b;
      const unsigned mask = rowmask[a] & rowmask[b];
      double acc[12];
#pragma unroll
      for (int j = 0; j < 12; j++) acc[j] = 0.0;
      for (int q = 0; q < 12; q++) {
        if ((mask >> q) & 1u) {
          const double wa = (double)rec[MPC_REC_WEIGHTS + q] * Ca[q * 12 + i];
#pragma unroll
          for (int j = 0; j < 12; j++) acc[j] += wa * Cb[q * 12 + j];
        }
      }
#pragma unroll
      for (int j = 0; j < 12; j++) k.M[tb * 144 + i * 12 + j] = acc[j];
    }
  }
  cx.sync();
  MPC_STAMP(k, cx, 11);
}

template <class Cx>
MPC_HD void assemble_H(const Cx& cx, const float* rec, const Work& k) {
  const int h = k.h;
  const Scalars* sc = k.sc;
  const int ns = sc->ns;
  const int na = ((double)rec[MPC_REC_XDRAG] != 0.0) ? 3 : 2;  // as in assemble_front
  // ---- P11: reduced Hessian, one 3x3 block per stance pair (a >= b) (SolverMPC.cpp:395):
  //   H[(i,la),(j,lb)] = 2 sum_{pa,pb} s_{pa,pb} M_{pa,pb}[la,lb] + 2 alpha I,
  //   s_{pa,pb} = sum_{q=0..n} q^pa (q+d)^pb, d = i-j >= 0, n = h-1-i, from the power sums P_e(n) = sum q^e
  //   (exact small integers in fp64 for every h <= 36). ----
  const double alpha = (double)rec[MPC_REC_ALPHA];
  const int nblk = ns * (ns + 1) / 2;
  MPC_FOR(e, nblk) {
    // e -> (a, b) with a >= b:  a = floor((sqrt(8e+1)-1)/2), float estimate + exact integer correction
    int a = (int)((sqrtf(8.0f * (float)e + 1.0f) - 1.0f) * 0.5f);
    while ((a + 1) * (a + 2) / 2 <= e) a++;
    while (a * (a + 1) / 2 > e) a--;
    const int b = e - a * (a + 1) / 2;
    const int ka = k.stance[a], kb = k.stance[b];
    const int i = ka >> 2, la = ka & 3, j = kb >> 2, lb = kb & 3;  // i >= j (stance[] is ascending)
    const double d = (double)(i - j);
    const double* P = k.psum + 5 * (h - 1 - i);
    const double P0 = P[0], P1 = P[1], P2 = P[2], P3 = P[3], P4 = P[4];
    // s_{pa,pb}: scalars for the no-drag case (the usual one: four terms, everything in registers); the general
    // case indexes a 3x3 array at run time
    const double s00 = P0, s01 = P1 + d * P0, s10 = P1, s11 = P2 + d * P1;
    for (int ax = 0; ax < 3; ax++)
      for (int bx = 0; bx < 3; bx++) {
        const int ci = la * 3 + ax, cj = lb * 3 + bx;
        double acc = 0;
        if (na == 2) {
          acc += s00 * k.M[0 * 144 + ci * 12 + cj];
          acc += s01 * k.M[1 * 144 + ci * 12 + cj];
          acc += s10 * k.M[1 * 144 + cj * 12 + ci];
          acc += s11 * k.M[3 * 144 + ci * 12 + cj];
        } else {
          double s[3][3];
          s[0][0] = P0;  s[0][1] = P1 + d * P0;  s[0][2] = P2 + 2.0 * d * P1 + d * d * P0;
          s[1][0] = P1;  s[1][1] = P2 + d * P1;  s[1][2] = P3 + 2.0 * d * P2 + d * d * P1;
          s[2][0] = P2;  s[2][1] = P3 + d * P2;  s[2][2] = P4 + 2.0 * d * P3 + d * d * P2;
          for (int pa = 0; pa < na; pa++)
            for (int pb = 0; pb < na; pb++) {
              // table of (min, max): 00 01 02 11 12 22 -> 0..5; the lower-index pair is stored, the other is its
              // transpose
              const int lo = pa < pb ? pa : pb, hi = pa < pb ? pb : pa;
              const int tb = lo * 3 - (lo * (lo - 1)) / 2 + (hi - lo);
              acc += s[pa][pb] * k.M[tb * 144 + (pa <= pb ? ci * 12 + cj : cj * 12 + ci)];
            }
        }
        double val = 2.0 * acc;
        if (a == b && ax == bx) val += 2.0 * alpha;
        k.Hm[hixT<Cx::kPacked>(k.ld, 3 * a + ax, 3 * b + bx)] = val;
        if (!Cx::kPacked) k.Hm[(3 * b + bx) * k.ld + 3 * a + ax] = val;
      }
  }
  cx.sync();
}

template <class Cx>
MPC_HD void assemble(const Cx& cx, const float* rec, const unsigned char* gait, const Work& k) {
  assemble_front(cx, rec, gait, k);
  if (k.sc->status != MPC_STATUS_OPTIMAL) return;
  assemble_H(cx, rec, k);
}

// ---------------------------------------------------------------------------
// Stage 2: Hm <- Hm^{-1} in place by symmetric sweeps (no pivoting: Hm is SPD).
// After sweeping every index the array holds -H^{-1}; the sign is flipped at the
// end.  A non-positive pivot means H is not positive definite.
// ---------------------------------------------------------------------------
// Generic form for a matrix in shared or global memory (the catch-all size class, whose matrix lives in an
// L2-resident global slab).  One pivot per pass: blocking several pivots per pass (rank-k updates) would cut
// the traffic by k, but block sweeps are numerically unusable here -- the k x k diagonal blocks inherit the
// alpha-regularised internal-force null space (cond ~ 1e4) and the one-shot Schur complement then loses
// 3 digits per doubling of k (measured: |Minv H - I| 3e-12 at k=1, 1e-9 at k=2, 9e-6 at k=4, 1e-3 at k=8).
// Only the lower triangle is updated (4x4 register tiles, uniform rank-1 update with the pivot slot holding
// d-1 as in the register-resident version); the pivot row is gathered from row p (j <= p) and column p
// (i > p); the result is mirrored, negated and the 2 taken off the diagonal in one final pass.
template <class Cx>
MPC_HD void invert_spd(const Cx& cx, const Work& k, int n_in = -1) {
  // Pivots are taken in groups of K = kSweepGroup so that the matrix (L2-resident in the catch-all class) is read
  // and written once per GROUP instead of once per pivot.  This is NOT the block sweep ruled out above: nothing is
  // inverted blockwise.  (A) the K pivot rows are gathered into fast memory; (B) they are swept against each other
  // one pivot at a time, and each row is frozen ("snapshot") the moment it becomes the pivot row -- exactly the row
  // the one-pivot-per-pass algorithm would broadcast, with slot p holding d-1 -- together with its scaled copy
  // u = -row/d; (C) one pass over the lower triangle applies the K rank-1 updates in pivot order to every tile,
  // a_ij = fma(u_p[i], row_p[j], a_ij).  Every element sees the same operations in the same order as with one pass
  // per pivot, so the result is bit-identical to it; only the traffic drops by K.
  constexpr int K = kSweepGroup;
  Scalars* sc = k.sc;
  const int nv = n_in >= 0 ? n_in : sc->nv, ld = k.ld;  // n_in: the matrix in Hm is n_in x n_in (wrench-space class)
  double* Hm = k.Hm;
  double* S = k.ck;                       // [K][nv_cap] snapshots of the pivot rows
  double* U = k.ck + K * k.nv_cap;        // [K][nv_cap] -snapshot/d
  const int lds = k.nv_cap;
  const int nt4 = (nv + 3) / 4;
  const int ntiles = nt4 * (nt4 + 1) / 2;
  for (int p0 = 0; p0 < nv; p0 += K) {
    const int kk = (nv - p0 < K) ? nv - p0 : K;
    // (A) gather rows p0..p0+kk-1 (row p = row p for j <= p, column p for i > p)
#pragma unroll 4
    for (int e = cx.tid; e < kk * nv; e += cx.nt) {  // unrolled: independent L2 loads in flight
      const int g = e / nv, i = e - g * nv, p = p0 + g;
      S[g * lds + i] = (i <= p) ? Hm[p * ld + i] : Hm[i * ld + p];
    }
    cx.sync();
    // (B) sweep the group's rows against each other, freezing each at its pivot time
    bool bad = false;
    for (int g = 0; g < kk; g++) {
      const int p = p0 + g;
      const double d = S[g * lds + p];
      if (!(d > 0.0)) { bad = true; break; }  // uniform: every thread reads the same pivot
      const double dinv = 1.0 / d;
      cx.sync();  // everybody has read d before the diagonal slot is rewritten
      MPC_FOR(i, nv) {
        const double c = (i == p) ? d - 1.0 : S[g * lds + i];
        if (i == p) S[g * lds + i] = c;
        U[g * lds + i] = -c * dinv;
      }
      cx.sync();
      // later rows r of the group: entry (r, j) sees what pass (C) applies to the stored element -- (r, j) itself
      // for j <= r, its mirror (j, r) for j > r: u taken at the row index, the pivot-row value at the column index
      MPC_FOR(e, (kk - 1 - g) * nv) {
        const int g2 = g + 1 + e / nv, j = e - (e / nv) * nv, r = p0 + g2;
        const double uv = (j <= r) ? U[g * lds + r] * S[g * lds + j] : U[g * lds + j] * S[g * lds + r];
        S[g2 * lds + j] = S[g2 * lds + j] + uv;
      }
      cx.sync();
    }
    if (bad) {
      cx.sync();
      MPC_ONE sc->status = MPC_STATUS_NOT_PD;
      cx.sync();
      return;
    }
    // (C) one pass over the lower triangle: two tiles per iteration with all loads issued before the arithmetic
    // (the matrix sits in L2, so the pass is bound by outstanding loads per thread)
    for (int t0 = cx.tid; t0 < ntiles; t0 += 2 * cx.nt) {
      double acc[2][4][4];
      int ti0[2], tj0[2];
#pragma unroll
      for (int s2 = 0; s2 < 2; s2++) {
        const int t = t0 + s2 * cx.nt;
        int bi = 0, bj = 0;
        if (t < ntiles) {
          bi = (int)((sqrtf(8.0f * (float)t + 1.0f) - 1.0f) * 0.5f);
          while ((bi + 1) * (bi + 2) / 2 <= t) bi++;
          while (bi * (bi + 1) / 2 > t) bi--;
          bj = t - bi * (bi + 1) / 2;
        }
        ti0[s2] = (t < ntiles) ? 4 * bi : nv;  // nv: nothing in range
        tj0[s2] = 4 * bj;
#pragma unroll
        for (int di = 0; di < 4; di++)
#pragma unroll
          for (int dj = 0; dj < 4; dj++) {
            const int i = ti0[s2] + di, j = tj0[s2] + dj;
            acc[s2][di][dj] = (i < nv && j <= i) ? Hm[i * ld + j] : 0.0;
          }
      }
#pragma unroll
      for (int s2 = 0; s2 < 2; s2++) {
        if (ti0[s2] >= nv) continue;
#pragma unroll 1
        for (int g = 0; g < kk; g++) {
          double u[4], v[4];
#pragma unroll
          for (int e = 0; e < 4; e++) {
            u[e] = (ti0[s2] + e < nv) ? U[g * lds + ti0[s2] + e] : 0.0;
            v[e] = (tj0[s2] + e < nv) ? S[g * lds + tj0[s2] + e] : 0.0;
          }
#pragma unroll
          for (int di = 0; di < 4; di++)
#pragma unroll
            for (int dj = 0; dj < 4; dj++) acc[s2][di][dj] = acc[s2][di][dj] + u[di] * v[dj];
        }
#pragma unroll
        for (int di = 0; di < 4; di++)
#pragma unroll
          for (int dj = 0; dj < 4; dj++) {
            const int i = ti0[s2] + di, j = tj0[s2] + dj;
            if (i < nv && j <= i) Hm[i * ld + j] = acc[s2][di][dj];
          }
      }
    }
    cx.sync();
  }
  // mirror the lower triangle, negate, take the 2 off the (swept) diagonal
#pragma unroll 4
  for (int e = cx.tid; e < nv * nv; e += cx.nt) {  // unrolled: independent L2 loads in flight
    const int i = e / nv, j = e - i * nv;
    if (j < i) {
      const double v = -Hm[i * ld + j];
      Hm[i * ld + j] = v;
      Hm[j * ld + i] = v;
    } else if (j == i) {
      Hm[i * ld + i] = 2.0 - Hm[i * ld + i];
    }
  }
  cx.sync();
}

#if defined(__CUDACC__)
// ---------------------------------------------------------------------------
// Stage 2, register-resident (device only): the same symmetric sweep with the matrix held in
// registers as R x C tiles.  NT = GR*GC threads; thread (tr, tc), tr = tid / GC, tc = tid % GC, owns the
// rows  tr + GR*i (i < R)  and the column pairs  2*GC*j2 + 2*tc + e  (j2 < C/2, e < 2) of the NVP x NVP
// matrix, NVP = GR*R = GC*C >= nv (identity padding).  GC divides 32, so the GC owners of a row are
// lanes of one warp.
//
// Pivots run in natural order.  Pivot p = GR*i + q lives in row slot i of the threads with tr == q, so
// the pivot loop is R copies (i unrolled) of a run-time loop over q and every register index is a
// compile-time constant.  Per pivot the owners publish ROW p (= column p, by symmetry) through shared
// memory, double-buffered (one barrier per pivot), with slot p carrying d-1 instead of d = a_pp and
// slot NVP carrying 1/d; every thread then applies ONE uniform rank-1 update
//     a_rj -= (c_r/d) * c_j          with c = published row, c_p := d-1
// which yields  a_rj - c_r c_j/d  (r,j != p),  c_j/d  (r == p or j == p)  and  2 - 1/d  at (p,p): every
// swept diagonal entry ends exactly 2 above its true value -1/d (later updates are additive), so the fix
// is a single "-2" on the diagonal when the result is stored.  The diagonal entries of a thread's R rows
// are tracked in R extra registers (one more DFMA each per pivot) so that d and 1/d are available without
// a run-time register index, and 1/d of the NEXT pivot is started before the bulk update so its latency
// hides behind the R*C DFMAs.  Shared-memory traffic per thread per pivot: R 8-byte + C/2 16-byte
// broadcast loads for R*C DFMAs (the kernel is otherwise bound by shared-memory load bandwidth, not fp64).
// ---------------------------------------------------------------------------
// 1/d for a positive, finite, normal double: hardware seed (rcp.approx.ftz.f64, ~20 bits) + two Newton steps.
// Branch-free, so every thread can run it next to the bulk update without diverging; non-positive or
// non-finite d gives a result the caller's  dinv > 0 && dinv < 1e300  test rejects (inf / nan / negative).
__device__ __forceinline__ double fast_rcp(double d) {
  double x;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(d));
  double e = fma(-d, x, 1.0);
  x = fma(x, e, x);
  e = fma(-d, x, 1.0);
  x = fma(x, e, x);
  return x;
}

// Symmetry: the tile (row slot i, column-pair slot j2) of every thread lies in the GR x 2GC super-block
// rows [GR*i, GR*i+GR) x columns [2GC*j2, 2GC*j2+2GC).  Super-blocks strictly above the diagonal are never
// loaded, updated or stored (a compile-time predicate, the same for all threads): their entries are the
// transposes of entries in kept blocks.  For R = 4 that is 20 instead of 32 DFMAs per pivot.  The part of pivot
// row p that falls into skipped blocks is published from column p instead (other threads, still compile-time
// register indices).
template <int GR, int GC>
__host__ __device__ constexpr bool sweep_block_kept(int i, int j2) { return 2 * GC * j2 <= GR * i + GR - 1; }

// with_g: row nv of the padded matrix (free when nv < NVP; all of it lies in kept super-blocks) carries the
// gradient g.  It is never pivoted, so the sweep turns it into H^{-1} g and the unconstrained optimum
// x = -H^{-1} g comes out of the inversion for free (no matrix-vector product afterwards).
template <int GR, int R, int GC, int C, bool kPacked>
__device__ __forceinline__ void invert_spd_tiles(const Work& k, int tid, bool with_g, int n_in = -1) {
  constexpr int NVP = GR * R;
  constexpr int BUF = NVP + 2;
  static_assert(GC * C == NVP && C % 2 == 0 && 32 % GC == 0, "tile grid must cover the padded matrix");
  static_assert((2 * GC) % GR == 0 || GR % (2 * GC) == 0, "super-block indices of a row slot must be compile-time");
  Scalars* sc = k.sc;
  const int nv = n_in >= 0 ? n_in : sc->nv, ld = k.ld;  // n_in: the matrix in Hm is n_in x n_in (wrench-space class)
  double* Hm = k.Hm;
  const int tr = tid / GC, tc = tid % GC, lane = tid & 31;
  const bool gaug = with_g && nv < NVP;
  double a[R][C];
#pragma unroll
  for (int i = 0; i < R; i++) {
    const int r = tr + GR * i;
#pragma unroll
    for (int j = 0; j < C; j++) {
      if (!sweep_block_kept<GR, GC>(i, j / 2)) continue;
      const int c = 2 * GC * (j / 2) + 2 * tc + (j & 1);
      double v = (r < nv && c < nv) ? Hm[hixT<kPacked>(ld, r, c)] : (r == c ? 1.0 : 0.0);
      if (gaug && r == nv && c < nv) v = k.g[c];
      if (gaug && c == nv && r < nv) v = k.g[r];
      a[i][j] = v;
    }
  }
  // two broadcast buffers (even / odd pivot): slots 0..NVP-1 the pivot row, slot NVP = 1/d
  double* const buf0 = k.ck;
  double* const buf1 = k.ck + BUF;
  bool bad = false;
#pragma unroll
  for (int i = 0; i < R; i++) {
    if (GR * i >= nv) break;  // uniform
    // diagonal entry of this thread's row in slot i (row r = tr + GR*i, not pivoted yet, so the tile copy is the
    // exact Schur complement): it sits in the lane with tc == (r/2) % GC at local column 2*(r/(2*GC)) + (r&1)
    // (a kept, diagonal super-block); a compile-time select chain + one shuffle inside the row group fetches
    // it.  Tracked from here on with one DFMA per pivot so that d and 1/d of the coming pivots never need a
    // run-time register index.
    double dg;
    {
      const int r = tr + GR * i;
      const int jj = 2 * (r / (2 * GC)) + (r & 1);
      double mine = 0.0;
#pragma unroll
      for (int j = 0; j < C; j++)
        if (sweep_block_kept<GR, GC>(i, j / 2)) mine = (j == jj) ? a[i][j] : mine;
      dg = __shfl_sync(0xffffffffu, mine, (lane & ~(GC - 1)) + ((r / 2) % GC));
    }
    double dinv_mine = fast_rcp(dg);  // 1/d of the pivot this thread's row group publishes next
    constexpr int dummy = 0;
    (void)dummy;
    const int jc = (GR * i) / (2 * GC);  // column super-block of the pivots of this row slot (compile-time)
#pragma unroll 1
    for (int q0 = 0; q0 < GR; q0 += 2) {
#pragma unroll
      for (int e = 0; e < 2; e++) {  // e = parity of q = parity of p: which buffer, which half of a column pair
        const int q = q0 + e;
        const int p = GR * i + q;
        if (p >= nv) break;  // uniform (and p+1 >= nv follows)
        double* const cur = e ? buf1 : buf0;
        if (tr == q) {  // row owners: the part of row p inside kept super-blocks, compile-time registers
#pragma unroll
          for (int j2 = 0; j2 < C / 2; j2++)
            if (sweep_block_kept<GR, GC>(i, j2))
              *reinterpret_cast<double2*>(cur + 2 * GC * j2 + 2 * tc) = make_double2(a[i][2 * j2], a[i][2 * j2 + 1]);
          // slot p was just written (with d) by the lane of this row group that holds column p: the SAME lane
          // overwrites it with d-1 (program order of one thread); every lane of the group tracks the same dg
          if (tc == ((p >> 1) % GC)) {
            cur[p] = dg - 1.0;
            cur[NVP] = dinv_mine;
          }
        }
        if (tc == ((p % (2 * GC)) >> 1)) {  // column holders: the part of row p inside skipped super-blocks
#pragma unroll
          for (int i2 = 0; i2 < R; i2++)
            if (!sweep_block_kept<GR, GC>(i, (GR * i2) / (2 * GC))) cur[tr + GR * i2] = a[i2][2 * jc + e];
        }
        __syncthreads();
        const double dinv = cur[NVP];
        // positive, normal and < 1e300, tested on the high word (two integer instructions instead of two DSETPs
        // on the fp64 pipe): NaN, inf, zero, subnormal and negative values all fall outside the window
        bad = bad || (unsigned)(__double2hiint(dinv) - 0x00100000) >= (unsigned)(0x7E37E43C - 0x00100000);
        double u[R];
#pragma unroll
        for (int ii = 0; ii < R; ii++) {
          const double c = cur[tr + GR * ii];
          u[ii] = -c * dinv;
          if (ii == i) dg = fma(u[ii], c, dg);
        }
        // reciprocal of the next pivot of this block, started before the bulk update so that its latency hides
        // behind the DFMAs.  Every thread computes it for its own row (branch-free: a divergent or
        // warp-selective version puts the reciprocal's latency back on the barrier's critical path).
        dinv_mine = fast_rcp(dg);
#pragma unroll
        for (int j2 = 0; j2 < C / 2; j2++) {
          const double2 v = *reinterpret_cast<const double2*>(cur + 2 * GC * j2 + 2 * tc);
#pragma unroll
          for (int ii = 0; ii < R; ii++) {
            if (!sweep_block_kept<GR, GC>(ii, j2)) continue;
            a[ii][2 * j2] = fma(u[ii], v.x, a[ii][2 * j2]);
            a[ii][2 * j2 + 1] = fma(u[ii], v.y, a[ii][2 * j2 + 1]);
          }
        }
      }
    }
  }
  if (__syncthreads_or(bad)) {
    if (tid == 0) sc->status = MPC_STATUS_NOT_PD;
    __syncthreads();
    return;
  }
  // store H^{-1} = -(swept matrix), taking the 2 off every (swept) diagonal entry; entries whose transpose lies
  // in a skipped super-block are written to both places; row nv (the swept gradient) -> x = -H^{-1} g
#pragma unroll
  for (int i = 0; i < R; i++) {
    const int r = tr + GR * i;
    if (gaug && r == nv) {
#pragma unroll
      for (int j = 0; j < C; j++) {
        if (!sweep_block_kept<GR, GC>(i, j / 2)) continue;
        const int c = 2 * GC * (j / 2) + 2 * tc + (j & 1);
        if (c < nv) k.x[c] = -a[i][j];
      }
    }
    if (r < nv) {
#pragma unroll
      for (int j = 0; j < C; j++) {
        if (!sweep_block_kept<GR, GC>(i, j / 2)) continue;
        const int c = 2 * GC * (j / 2) + 2 * tc + (j & 1);
        if (c < nv) {
          const double val = (c == r) ? (2.0 - a[i][j]) : -a[i][j];
          // packed storage: one writer per unordered pair -- the lower-triangular orientation, which always lies
          // in a kept super-block (the diagonal super-blocks hold both orientations)
          if (!kPacked) {
            Hm[r * ld + c] = val;
            if (!sweep_block_kept<GR, GC>(c / GR, r / (2 * GC))) Hm[c * ld + r] = val;
          } else if (r >= c) {
            Hm[tri_index(r, c)] = val;
          }
        }
      }
    }
  }
  __syncthreads();
}

// ---------------------------------------------------------------------------
// Stage 2 on the FP64 tensor pipe (device only): a grouped symmetric sweep whose rank-8 updates, panel and pivot
// block all run as DMMA.8x8x4 (mma.sync.m8n8k4.f64).
//
// The sweep of invert_spd_tiles applies one rank-1 update per pivot, with a publish -> barrier -> load -> update
// chain per pivot (60 barriers for a trot problem) and 8 issue slots per 256 FMAs.  Here pivots are taken eight at
// a time (a block row of the 8x8 block grid).  Nothing is inverted blockwise -- the group's pivots are still
// eliminated one after the other, through the factored form, so the alpha-regularised null space that rules out
// block sweeps (see invert_spd) does not matter:
//   (A)  gather      the 8 pivot rows of the group (= block row b and, by symmetry, block column b) go from the
//                    accumulator fragments to shared memory (G);
//   (B1) pivot block one warp sweeps the 8x8 block of the pivot columns pivot by pivot (fragment layout, one
//                    rank-1 DMMA per pivot), carrying the row operations along on an identity block.  Row g of
//                    either block is frozen the moment g becomes the pivot: that gives the frozen pivot-column
//                    part of row g (slot p holding d-1, as in invert_spd_tiles), u_g = -row_g/d_g, 1/d_g, and
//                    row g of T, the accumulated row operation (unit lower triangular);
//   (B2) panel       frozen rows of all other columns F = T G, two DMMAs per 8 columns, and U = -D^{-1} F;
//   (C)  update      A += U' F for every stored 8x8 block: two DMMAs (pivots 0-3, 4-7, i.e. pivot order) fed by
//                    one A fragment (u values of the block's rows) and one B fragment (frozen rows of its columns).
// Three CTA barriers per GROUP instead of one per pivot, and 1/8 of the issue slots for the same FMAs.
//
// Storage: the NVP x NVP matrix (NVP = 8*NB, identity padding, row nv = the gradient as in invert_spd_tiles) is
// held as 8x8 accumulator fragments (lane l: row l/4, columns 2*(l%4), 2*(l%4)+1).  Of every pair of mirror
// blocks only one is stored, chosen cyclically so that every warp owns whole block rows with the same number of
// blocks: block row i holds the columns (i+k) mod NB for k = 0 .. NB/2-1, plus k = NB/2 for the rows i with
// (i < NB/2) != (i odd) -- each pair {i, i+NB/2} exactly once.  Warp w owns block rows w*RW .. w*RW+RW-1.
// The panel buffers live in the Hm region (H is in registers for the whole sweep and H^{-1} is only stored at the
// end), so the stage needs no shared memory of its own.
// ---------------------------------------------------------------------------
// shared-window accesses by 32-bit byte address (volatile: ordered with each other and with the DMMAs)
__device__ __forceinline__ double lds_f64(uint32_t a) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ void lds_f64x2(uint32_t a, double& v0, double& v1) {
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v0), "=d"(v1) : "r"(a));
}
__device__ __forceinline__ void sts_f64(uint32_t a, double v) {
  asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory");
}
__device__ __forceinline__ void sts_f64_if(bool p, uint32_t a, double v) {  // predicated: no branch, no divergence
  asm volatile("{\n.reg .pred q;\nsetp.ne.b32 q, %0, 0;\n@q st.shared.f64 [%1], %2;\n}" ::"r"((int)p), "r"(a), "d"(v) : "memory");
}
__device__ __forceinline__ void sts_f64x2(uint32_t a, double v0, double v1) {
  asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(a), "d"(v0), "d"(v1) : "memory");
}
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// NT threads in the CTA, of which the first NWS warps sweep (the others only meet the barriers).
template <int NT, int NWS, int NB, bool kPacked>
__device__ __forceinline__ void invert_spd_mma(const Work& k, int tid, bool with_g, int n_in = -1) {
  constexpr int NVP = 8 * NB, RW = NB / NWS, HB = NB / 2, NSLOT = HB + 1, NCOL = RW - 1 + NSLOT;
  constexpr int LDP = NVP + 4, KG = 8;  // = mma_panel_ld(NVP), kMmaGroup (host-side constexpr functions)
  static_assert(NB % NWS == 0 && HB % 2 == 0 && NWS * 32 <= NT && (NWS & (NWS - 1)) == 0, "block rows per warp");
  Scalars* sc = k.sc;
  const int nv = n_in >= 0 ? n_in : sc->nv, ld = k.ld;  // n_in: the matrix in Hm is n_in x n_in (wrench-space class)
  double* Hm = k.Hm;
  const int warp = tid >> 5, lane = tid & 31, lr = lane >> 2, lc = lane & 3;
  const bool gaug = with_g && nv < NVP;
  const bool active = warp < NWS;
  const int i0 = warp * RW;  // first block row of this warp
  // slot (ri, ks): block row i0 + ri, block column colb[ri + ks] = (i0 + ri + ks) mod NB; ks == HB only for the
  // "extra" rows
  auto extra = [](int i) { return (i < HB) != ((i & 1) != 0); };
  int colb[NCOL];
  bool ext[RW];
#pragma unroll
  for (int t = 0; t < NCOL; t++) colb[t] = (i0 + t >= NB) ? i0 + t - NB : i0 + t;
#pragma unroll
  for (int ri = 0; ri < RW; ri++) ext[ri] = extra(i0 + ri);
  double c[RW][NSLOT][2];
  if (active) {
#pragma unroll
    for (int ri = 0; ri < RW; ri++) {
      const int r = 8 * (i0 + ri) + lr;
#pragma unroll
      for (int ks = 0; ks < NSLOT; ks++) {
#pragma unroll
        for (int e = 0; e < 2; e++) {
          const int cc = 8 * colb[ri + ks] + 2 * lc + e;
          double v = (r == cc) ? 1.0 : 0.0;
          if (ks < HB || ext[ri]) {
            if (r < nv && cc < nv) v = Hm[hixT<kPacked>(ld, r, cc)];
            if (gaug && r == nv && cc < nv) v = k.g[cc];
            if (gaug && cc == nv && r < nv) v = k.g[r];
          }
          c[ri][ks][e] = v;
        }
      }
    }
  }
  __syncthreads();  // H is in registers everywhere: the Hm region becomes the panel buffer
  // carve() picks Hm from shared memory or the global slab at run time, which leaves a generic pointer behind, and
  // generic LD / ST on the panel cost three times what LDS / STS do: the panel is addressed in the shared window
  // explicitly (32-bit byte addresses, ld.shared / st.shared)
  const uint32_t G = (uint32_t)__cvta_generic_to_shared(Hm);  // [KG][LDP]  gathered rows of the current group
  const uint32_t F = G + 8 * KG * LDP;                         // [KG][LDP]  frozen rows
  const uint32_t U = G + 8 * 2 * KG * LDP;                     // [KG][LDP]  -frozen row / pivot
  const uint32_t TF = G + 8 * 3 * KG * LDP;                    // [KG][KG]   frozen rows of the row operation T
  const uint32_t DI = TF + 8 * KG * KG;                        // [KG]       reciprocals of the group's pivots
  const uint32_t frag = 8 * (lc * LDP + lr);                   // fragment position: row lc (+4), column lr
  bool bad = false;
  const int ngroups = (nv + KG - 1) / KG;
#ifdef MPC_SWEEP_CLK  // harness only (tools/microbench/sweep_mma_test.cu): cycles per phase, summed over the groups
  long long tA = 0, tB = 0, tB1 = 0, tC = 0, wA = 0, wB1 = 0, wB = 0, tq = clock64();
#define MPC_SWEEP_TICK(acc) do { const long long now_ = clock64(); acc += now_ - tq; tq = now_; } while (0)
#else
#define MPC_SWEEP_TICK(acc) do { } while (0)
#endif
#pragma unroll 1
  for (int b = 0; b < ngroups; b++) {
    const int p0 = KG * b;
    // ---- (A) gather: block row b as rows, block column b transposed ----
    if (active) {
#pragma unroll
      for (int ri = 0; ri < RW; ri++) {
        if (i0 + ri == b) {  // uniform per warp
#pragma unroll
          for (int ks = 0; ks < NSLOT; ks++)
            if (ks < HB || ext[ri])
              sts_f64x2(G + 8 * (lr * LDP + 2 * lc) + 64 * colb[ri + ks], c[ri][ks][0], c[ri][ks][1]);
        } else {
          const uint32_t gt = G + 8 * ((2 * lc) * LDP + 8 * (i0 + ri) + lr);
#pragma unroll
          for (int ks = 1; ks < NSLOT; ks++)
            if (colb[ri + ks] == b && (ks < HB || ext[ri])) {
              sts_f64(gt, c[ri][ks][0]);
              sts_f64(gt + 8 * LDP, c[ri][ks][1]);
            }
        }
      }
    }
    MPC_SWEEP_TICK(tA);
    __syncthreads();
    MPC_SWEEP_TICK(wA);
    // ---- (B1) pivot block, one warp.  d = the 8x8 block of the pivot columns, t = T' (the transposed row
    //      operation, starts as the identity), both in fragment layout.  Pivot g: the pivot row at index lr comes
    //      from COLUMN g of d (the block is symmetric; the element sits in this lane's own quad), row g of T from
    //      column g of t; both are frozen (stored) and the rank-1 updates d += a v', t += w a' are one DMMA each
    //      with the operands in the k = 0 slot.  Straight-line code: the pivots past the last one of a short last
    //      group are turned into no-ops by selects (d = 1, zero row), not by branches -- one basic block, so the
    //      stores and the T update fill the gaps of the chain. ----
    if (warp == (b & (NWS - 1))) {
      double d0, d1;
      lds_f64x2(G + 8 * (lr * LDP + p0 + 2 * lc), d0, d1);
      double t0 = (lr == 2 * lc) ? 1.0 : 0.0, t1 = (lr == 2 * lc + 1) ? 1.0 : 0.0;
      const bool k0 = lc == 0;
      const uint32_t fo = F + 8 * (p0 + lr), uo = U + 8 * (p0 + lr), to = TF + 8 * lr;
#pragma unroll
      for (int g = 0; g < KG; g++) {
        const int src = (lane & ~3) + (g >> 1);
        const bool live = p0 + g < nv;  // uniform
        double v = __shfl_sync(0xffffffffu, (g & 1) ? d1 : d0, src);                // D(lr, g) = D(g, lr)
        double dp = __shfl_sync(0xffffffffu, (g & 1) ? d1 : d0, 4 * g + (g >> 1));  // D(g, g)
        double w = __shfl_sync(0xffffffffu, (g & 1) ? t1 : t0, src);                // T'(lr, g) = T(g, lr)
        v = live ? v : 0.0;
        dp = live ? dp : 1.0;
        w = live ? w : 0.0;
        const double dinv = fast_rcp(dp);
        // positive, normal and < 1e300, tested on the high word (see invert_spd_tiles)
        bad = bad || (unsigned)(__double2hiint(dinv) - 0x00100000) >= (unsigned)(0x7E37E43C - 0x00100000);
        if (lr == g) v = dp - 1.0;
        const double a = -v * dinv;
        const double az = k0 ? a : 0.0;
        dmma884(d0, d1, az, k0 ? v : 0.0);
        dmma884(t0, t1, k0 ? w : 0.0, az);
        // frozen row g at the pivot columns, its scaled copy, row g of T, 1/d (predicated stores: a branch here
        // would diverge the warp once per pivot, in the middle of the chain)
        sts_f64_if(k0, fo + 8 * g * LDP, v);
        sts_f64_if(k0, uo + 8 * g * LDP, a);
        sts_f64_if(k0, to + 8 * g * KG, w);
        sts_f64_if(lane == 0, DI + 8 * g, dinv);
      }
    }
    MPC_SWEEP_TICK(tB1);
    __syncthreads();
    MPC_SWEEP_TICK(wB1);
    // ---- (B2) frozen rows of the other columns, F = T G (rows past the last pivot come out as zeros), U = -F/d ----
    if (active) {
      const double ta0 = lds_f64(TF + 8 * (lr * KG + lc)), ta1 = lds_f64(TF + 8 * (lr * KG + 4 + lc));
      const double ndi = -lds_f64(DI + 8 * lr);
#pragma unroll
      for (int j = 0; j < RW; j++) {
        const int cbk = i0 + j;
        if (cbk == b) continue;  // uniform: the pivot columns are done
        const double g0 = lds_f64(G + frag + 64 * cbk), g1 = lds_f64(G + frag + 8 * 4 * LDP + 64 * cbk);
        double f0 = 0.0, f1 = 0.0;
        dmma884(f0, f1, ta0, g0);
        dmma884(f0, f1, ta1, g1);
        sts_f64x2(F + 8 * (lr * LDP + 2 * lc) + 64 * cbk, f0, f1);
        sts_f64x2(U + 8 * (lr * LDP + 2 * lc) + 64 * cbk, f0 * ndi, f1 * ndi);
      }
    }
    MPC_SWEEP_TICK(tB);
    __syncthreads();
    MPC_SWEEP_TICK(wB);
    // ---- (C) rank-8 update of every stored block on the tensor pipe ----
    if (active) {
      const uint32_t Fp = F + frag;  // B fragment: F[g = lc (+4)][8*cb + lr]
      const uint32_t Up = U + frag;  // A fragment: U[g = lc (+4)][8*i + lr]
      double a0[RW], a1[RW];
#pragma unroll
      for (int ri = 0; ri < RW; ri++) {
        a0[ri] = lds_f64(Up + 64 * (i0 + ri));
        a1[ri] = lds_f64(Up + 8 * 4 * LDP + 64 * (i0 + ri));
      }
      // t = ri + ks: one B fragment serves every row that has column colb[t].  The fragment loads are volatile asm
      // (ordered with the DMMAs) and pipelined one step ahead by hand: left to itself the compiler hoists all of
      // them above the first DMMA and spills accumulators in the larger classes.
      double b0 = lds_f64(Fp + 64 * colb[0]), b1 = lds_f64(Fp + 8 * 4 * LDP + 64 * colb[0]);
#pragma unroll
      for (int t = 0; t < NCOL; t++) {
        const double b0c = b0, b1c = b1;
        if (t + 1 < NCOL) {
          b0 = lds_f64(Fp + 64 * colb[t + 1]);
          b1 = lds_f64(Fp + 8 * 4 * LDP + 64 * colb[t + 1]);
        }
#pragma unroll
        for (int ri = 0; ri < RW; ri++) {
          const int ks = t - ri;
          if (ks < 0 || ks >= NSLOT) continue;
          if (ks == HB && !ext[ri]) continue;  // uniform per warp
          dmma884(c[ri][ks][0], c[ri][ks][1], a0[ri], b0c);
          dmma884(c[ri][ks][0], c[ri][ks][1], a1[ri], b1c);
        }
      }
    }
    MPC_SWEEP_TICK(tC);
  }
#ifdef MPC_SWEEP_CLK
  if (k.clk && lane == 0 && warp < 4) { long long* c_ = k.clk + 8 * (1 + warp); c_[0] += tA; c_[1] += wA; c_[2] += tB1; c_[3] += wB1; c_[4] += tB; c_[5] += wB; c_[6] += tC; }
#endif
  if (__syncthreads_or(bad)) {  // also: everybody is done with the panel buffers
    if (tid == 0) sc->status = MPC_STATUS_NOT_PD;
    __syncthreads();
    return;
  }
  // store H^{-1} = -(swept matrix), the 2 taken off the swept diagonal; row / column nv -> x = -H^{-1} g
  if (active) {
#pragma unroll
    for (int ri = 0; ri < RW; ri++) {
      const int i = i0 + ri, r = 8 * i + lr;
#pragma unroll
      for (int ks = 0; ks < NSLOT; ks++) {
        if (!(ks < HB || ext[ri])) continue;
        const int cb = colb[ri + ks];
#pragma unroll
        for (int e = 0; e < 2; e++) {
          const int cc = 8 * cb + 2 * lc + e;
          const double v = c[ri][ks][e];
          if (gaug && r == nv && cc < nv) k.x[cc] = -v;
          if (gaug && cc == nv && r < nv && i != cb) k.x[r] = -v;
          if (r < nv && cc < nv) {
            const double val = (cc == r) ? (2.0 - v) : -v;
            if (!kPacked) {
              Hm[r * ld + cc] = val;
              if (i != cb) Hm[cc * ld + r] = val;
            } else if (i != cb || r >= cc) {  // one writer per unordered pair
              Hm[tri_index(r, cc)] = val;
            }
          }
        }
      }
    }
  }
  __syncthreads();
}

#endif

// ---------------------------------------------------------------------------
// Wrench-space class (problems with more reduced variables than 6h, e.g. three or four stance legs).
//
// Every column of B_c factors through the 6-dimensional wrench of its foot: B_c[:, leg] = Bw * G_leg with
// G_leg = [[r_leg]x ; I3] (torque r x f and force f; ct_ss_mats, SolverMPC.cpp:247-253) and Bw = [0; I_world^{-1}; I/m].
// Hence  B_qp(reduced) = Psi * G  with Psi the 13h x 6h horizon response to unit wrenches and G block-diagonal
// (one 6x3 block per stance pair, all pairs of a step in the same block row), and
//     H = 2 (alpha I + G' K G),   K = Psi' S Psi   (6h x 6h, the condensed Hessian of the wrench problem)
// has rank <= 6h above the alpha floor.  By the matrix-inversion lemma
//     H^{-1} = 1/(2 alpha) [ I - G' M G ],   M = (alpha K^{-1} + D)^{-1},   D = G G' (block diagonal, 6x6 per step),
// so two inversions of size 6h (120 at h = 20, where the reduced QP has up to 240 variables) replace one of size nv,
// they fit the register-resident sweep of the nv <= 128 class, and M fits in shared memory where H^{-1} (231 KB
// packed at nv = 240) does not.  H^{-1} is never formed: the active set applies it (wr_apply: two block-sparse
// products with G and one 6h x 6h matrix-vector product) to the dense vectors it needs.
// Verified against the dense route in numpy (|H - 2(alpha I + G'KG)| 6e-16 relative on BASELINE config 3) and
// by the parity tests against the reference solver.
// ---------------------------------------------------------------------------
// K, the condensed Hessian of the wrench problem (before the factor 2), into Hm (6h x 6h); the foot positions into
// rleg; D_s = sum over the stance legs of step s of G G' into dblk; 1/(2 alpha).  Runs after assemble_front.
template <class Cx>
MPC_HD void assemble_K(const Cx& cx, const float* rec, const Work& k) {
  const int h = k.h;
  double* Bw = k.M;            // 13 x 6 (the 12x12 tables of assemble_front are not needed by this class)
  double* Cw = k.M + 96;       // C0w, C1w, C2w: 3 x (13 x 6)
  double* Mw = Cw + 3 * 78;    // six 6 x 6 tables
  const double dt = (double)rec[MPC_REC_DT], xd = (double)rec[MPC_REC_XDRAG];
  const int na = (xd != 0.0) ? 3 : 2;
  MPC_FOR(i, 78) Bw[i] = 0.0;
  MPC_FOR(i, 12) k.rleg[i] = (double)rec[MPC_REC_R + (i % 3) * 4 + i / 3];  // rleg[leg*3 + axis] = r[axis*4 + leg]
  cx.sync();
  MPC_ONE {
    const double yaw = (double)rec[MPC_REC_YAW];
    const double yc = cos(yaw), ys = sin(yaw);
    const double R[3][3] = {{yc, -ys, 0}, {ys, yc, 0}, {0, 0, 1}};
    const double Ib[3] = {(double)rec[MPC_REC_IBODY], (double)rec[MPC_REC_IBODY + 1], (double)rec[MPC_REC_IBODY + 2]};
    double Iw[3][3];
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) {
        double acc = 0;
        for (int q = 0; q < 3; q++) acc += (R[i][q] * Ib[q]) * R[j][q];
        Iw[i][j] = acc;
      }
    const double det = Iw[0][0] * (Iw[1][1] * Iw[2][2] - Iw[1][2] * Iw[2][1]) -
                       Iw[0][1] * (Iw[1][0] * Iw[2][2] - Iw[1][2] * Iw[2][0]) +
                       Iw[0][2] * (Iw[1][0] * Iw[2][1] - Iw[1][1] * Iw[2][0]);
    const double rdet = 1.0 / det;
    Bw[6 * 6 + 0] = (Iw[1][1] * Iw[2][2] - Iw[1][2] * Iw[2][1]) * rdet;
    Bw[6 * 6 + 1] = (Iw[0][2] * Iw[2][1] - Iw[0][1] * Iw[2][2]) * rdet;
    Bw[6 * 6 + 2] = (Iw[0][1] * Iw[1][2] - Iw[0][2] * Iw[1][1]) * rdet;
    Bw[7 * 6 + 0] = (Iw[1][2] * Iw[2][0] - Iw[1][0] * Iw[2][2]) * rdet;
    Bw[7 * 6 + 1] = (Iw[0][0] * Iw[2][2] - Iw[0][2] * Iw[2][0]) * rdet;
    Bw[7 * 6 + 2] = (Iw[0][2] * Iw[1][0] - Iw[0][0] * Iw[1][2]) * rdet;
    Bw[8 * 6 + 0] = (Iw[1][0] * Iw[2][1] - Iw[1][1] * Iw[2][0]) * rdet;
    Bw[8 * 6 + 1] = (Iw[0][1] * Iw[2][0] - Iw[0][0] * Iw[2][1]) * rdet;
    Bw[8 * 6 + 2] = (Iw[0][0] * Iw[1][1] - Iw[0][1] * Iw[1][0]) * rdet;
    const double minv = 1.0 / (double)rec[MPC_REC_MASS];
    for (int i = 0; i < 3; i++) Bw[(9 + i) * 6 + 3 + i] = minv;
    k.red[32] = yc;
    k.red[33] = ys;
  }
  cx.sync();
  const double yc = k.red[32], ys = k.red[33];
  MPC_FOR(e, 78) {  // C0w = B_d for unit wrenches (exact cubic, as in assemble_front)
    const int i = e / 6, j = e - 6 * i;
    Cw[e] = dt * Bw[e] + (dt * dt / 2.0) * apply_A(Bw, 6, i, j, yc, ys, xd) + (dt * dt * dt / 6.0) * apply_A2(Bw, 6, i, j, xd);
  }
  cx.sync();
  MPC_FOR(e, 78) {
    const int i = e / 6, j = e - 6 * i;
    Cw[78 + e] = dt * apply_A(Cw, 6, i, j, yc, ys, xd);
    Cw[156 + e] = (dt * dt / 2.0) * apply_A2(Cw, 6, i, j, xd);
  }
  cx.sync();
  MPC_FOR(e, 6 * 36) {  // tables Mw_ab = C_aw' Q C_bw, a <= b, order 00 01 02 11 12 22
    const int tb = e / 36, ij = e - 36 * tb, i = ij / 6, j = ij - 6 * i;
    const int a = tb < 3 ? 0 : (tb < 5 ? 1 : 2), b = tb < 3 ? tb : (tb < 5 ? tb - 2 : 2);
    double acc = 0.0;
    if (b < na)
      for (int q = 0; q < 12; q++) acc += ((double)rec[MPC_REC_WEIGHTS + q] * Cw[78 * a + q * 6 + i]) * Cw[78 * b + q * 6 + j];
    Mw[e] = acc;
  }
  MPC_FOR(e, 36 * h) {  // D_s = sum_{stance legs of step s} G G',  G = [[r]x ; I]
    const int s = e / 36, ij = e - 36 * s, i = ij / 6, j = ij - 6 * i;
    double acc = 0.0;
    for (int l = 0; l < 4; l++) {
      if (k.posk[4 * s + l] < 0) continue;
      const double rx = k.rleg[3 * l], ry = k.rleg[3 * l + 1], rz = k.rleg[3 * l + 2];
      const double cm[3][3] = {{0, -rz, ry}, {rz, 0, -rx}, {-ry, rx, 0}};
      // G row i: i < 3 -> cm[i][.], else unit vector e_{i-3}
      double d = 0.0;
      for (int q = 0; q < 3; q++) {
        const double gi = i < 3 ? cm[i][q] : (i - 3 == q ? 1.0 : 0.0);
        const double gj = j < 3 ? cm[j][q] : (j - 3 == q ? 1.0 : 0.0);
        d += gi * gj;
      }
      acc += d;
    }
    k.dblk[e] = acc;
  }
  cx.sync();
  // K blocks, i >= j:  K[(i,.),(j,.)] = sum_{pa,pb} s_{pa,pb}(i,j) Mw_{pa,pb}  with the power sums of assemble_front
  const int nblk = h * (h + 1) / 2;
  MPC_FOR(e, nblk * 36) {
    const int blk = e / 36, ij = e - 36 * blk, ci = ij / 6, cj = ij - 6 * ci;
    int i = (int)((sqrtf(8.0f * (float)blk + 1.0f) - 1.0f) * 0.5f);
    while ((i + 1) * (i + 2) / 2 <= blk) i++;
    while (i * (i + 1) / 2 > blk) i--;
    const int j = blk - i * (i + 1) / 2;
    const double d = (double)(i - j);
    const double* P = k.psum + 5 * (h - 1 - i);
    const double P0 = P[0], P1 = P[1], P2 = P[2], P3 = P[3], P4 = P[4];
    double sm[3][3];
    sm[0][0] = P0;  sm[0][1] = P1 + d * P0;  sm[0][2] = P2 + 2.0 * d * P1 + d * d * P0;
    sm[1][0] = P1;  sm[1][1] = P2 + d * P1;  sm[1][2] = P3 + 2.0 * d * P2 + d * d * P1;
    sm[2][0] = P2;  sm[2][1] = P3 + d * P2;  sm[2][2] = P4 + 2.0 * d * P3 + d * d * P2;
    double acc = 0.0;
    for (int pa = 0; pa < na; pa++)
      for (int pb = 0; pb < na; pb++) {
        const int lo = pa < pb ? pa : pb, hi = pa < pb ? pb : pa;
        const int tb = lo * 3 - (lo * (lo - 1)) / 2 + (hi - lo);
        acc += sm[pa][pb] * Mw[tb * 36 + (pa <= pb ? ci * 6 + cj : cj * 6 + ci)];
      }
    const int r_ = 6 * i + ci, c_ = 6 * j + cj;
    if (Cx::kPacked) {
      if (r_ >= c_) k.Hm[tri_index(r_, c_)] = acc;
    } else {
      k.Hm[r_ * k.ld + c_] = acc;
      k.Hm[c_ * k.ld + r_] = acc;
    }
  }
  cx.sync();
}

// Hm <- alpha * Hm + D (block diagonal), between the two inversions: K^{-1} -> alpha K^{-1} + D.
template <class Cx>
MPC_HD void wr_form_second(const Cx& cx, const float* rec, const Work& k) {
  const int n = 6 * k.h;
  const double alpha = (double)rec[MPC_REC_ALPHA];
  MPC_FOR(e, n * n) {
    const int i = e / n, j = e - i * n;
    if (j > i) continue;
    const double dd = (i / 6 == j / 6) ? k.dblk[36 * (i / 6) + 6 * (i % 6) + (j % 6)] : 0.0;
    const double v = alpha * k.Hm[hixT<Cx::kPacked>(k.ld, i, j)] + dd;
    k.Hm[hixT<Cx::kPacked>(k.ld, i, j)] = v;
    if (!Cx::kPacked && i != j) k.Hm[j * k.ld + i] = v;
  }
  cx.sync();
}

// z = H^{-1} v = 1/(2 alpha) (v - G' M G v) for dense v, z of length nv (v == z allowed).  Uses wy, wt.
template <class Cx>
MPC_HD void wr_apply(const Cx& cx, const Work& k, const double* v, double* z) {
  const int h = k.h, n = 6 * h, nv = k.sc->nv;
  MPC_FOR(e, n) {  // wy = G v: per step, sum over its stance legs of [r x v_j ; v_j]
    const int s = e / 6, c = e - 6 * s;
    double acc = 0.0;
    for (int l = 0; l < 4; l++) {
      const int j = k.posk[4 * s + l];
      if (j < 0) continue;
      const double vx = v[3 * j], vy = v[3 * j + 1], vz = v[3 * j + 2];
      const double rx = k.rleg[3 * l], ry = k.rleg[3 * l + 1], rz = k.rleg[3 * l + 2];
      double t;
      switch (c) {
        case 0: t = ry * vz - rz * vy; break;
        case 1: t = rz * vx - rx * vz; break;
        case 2: t = rx * vy - ry * vx; break;
        case 3: t = vx; break;
        case 4: t = vy; break;
        default: t = vz; break;
      }
      acc += t;
    }
    k.wy[e] = acc;
  }
  cx.sync();
  // wt = M wy.  wy is block sparse (only the steps the working set touches carry a wrench), so all-zero blocks are
  // skipped; with enough threads two of them share a row (even / odd steps, partial sums in wt[0..n) and wt[n..2n)).
  // Packed storage: M(i, j) for j <= i is the contiguous run of row i, for j > i a walk down column i.
  const int parts = (cx.nt >= 2 * n) ? 2 : 1;
#pragma unroll 1
  for (int e = cx.tid; e < n * parts; e += cx.nt) {
    const int part = e >= n ? 1 : 0, i = e - part * n;
    const int ti = Cx::kPacked ? i * (i + 1) / 2 : i * k.ld;
    double a0 = 0, a1 = 0;
#pragma unroll 1
    for (int s = part; s < h; s += parts) {
      const double* y = k.wy + 6 * s;
      const double y0 = y[0], y1 = y[1], y2 = y[2], y3 = y[3], y4 = y[4], y5 = y[5];
      if (y0 == 0.0 && y1 == 0.0 && y2 == 0.0 && y3 == 0.0 && y4 == 0.0 && y5 == 0.0) continue;
      const int j0 = 6 * s;
      double m[6];
#pragma unroll
      for (int c = 0; c < 6; c++) {
        const int j = j0 + c;
        m[c] = k.Hm[(!Cx::kPacked || j <= i) ? ti + j : j * (j + 1) / 2 + i];
      }
      a0 += m[0] * y0; a1 += m[1] * y1; a0 += m[2] * y2; a1 += m[3] * y3; a0 += m[4] * y4; a1 += m[5] * y5;
    }
    k.wt[part * n + i] = a0 + a1;
  }
  cx.sync();
  const double i2a = k.i2a;
  MPC_FOR(i, nv) {  // z = (v - G' wt) / (2 alpha);  G_l' t = t_force - r x t_torque
    const int j = i / 3, c = i - 3 * j;
    const int kk = k.stance[j], s = kk >> 2, l = kk & 3;
    double t[6];
#pragma unroll
    for (int q = 0; q < 6; q++) t[q] = k.wt[6 * s + q] + (parts == 2 ? k.wt[n + 6 * s + q] : 0.0);
    const double rx = k.rleg[3 * l], ry = k.rleg[3 * l + 1], rz = k.rleg[3 * l + 2];
    double cr;
    switch (c) {
      case 0: cr = ry * t[2] - rz * t[1]; break;
      case 1: cr = rz * t[0] - rx * t[2]; break;
      default: cr = rx * t[1] - ry * t[0]; break;
    }
    z[i] = i2a * (v[i] - (t[3 + c] - cr));
  }
  cx.sync();
}

// z = H^{-1} n for ONE catalogue row n (the row being added): its wrench image lives in a single step block, so the
// product with M touches six columns and needs no scratch vector for G v.
template <class Cx>
MPC_HD void wr_apply_row(const Cx& cx, const Work& k, const Row& rp, double* z) {
  const int h = k.h, n = 6 * h, nv = k.sc->nv;
  const int jp = rp.iz / 3, kp = k.stance[jp], sp = kp >> 2, lp = kp & 3;
  double nl[3] = {0.0, 0.0, rp.cz};
  if (rp.ca != 0.0) nl[rp.ia - 3 * jp] = rp.ca;
  const double rx = k.rleg[3 * lp], ry = k.rleg[3 * lp + 1], rz = k.rleg[3 * lp + 2];
  const double q[6] = {ry * nl[2] - rz * nl[1], rz * nl[0] - rx * nl[2], rx * nl[1] - ry * nl[0], nl[0], nl[1], nl[2]};
  MPC_FOR(i, n) {
    const int ti = Cx::kPacked ? i * (i + 1) / 2 : i * k.ld;
    double acc = 0.0;
#pragma unroll
    for (int c = 0; c < 6; c++) {
      const int j = 6 * sp + c;
      acc += k.Hm[(!Cx::kPacked || j <= i) ? ti + j : j * (j + 1) / 2 + i] * q[c];
    }
    k.wt[i] = acc;
  }
  cx.sync();
  const double i2a = k.i2a;
  MPC_FOR(i, nv) {
    const int j = i / 3, c = i - 3 * j;
    const int kk = k.stance[j], s = kk >> 2, l = kk & 3;
    const double* t = k.wt + 6 * s;
    const double ax = k.rleg[3 * l], ay = k.rleg[3 * l + 1], az = k.rleg[3 * l + 2];
    double cr;
    switch (c) {
      case 0: cr = ay * t[2] - az * t[1]; break;
      case 1: cr = az * t[0] - ax * t[2]; break;
      default: cr = ax * t[1] - ay * t[0]; break;
    }
    const double vi = (j == jp) ? nl[c] : 0.0;
    z[i] = i2a * (vi - (t[3 + c] - cr));
  }
  cx.sync();
}

// wv = sum of catalogue rows weighted by rcat (dense N c), for wr_apply.  rcat[6j + t] = coefficient of row t of pair j.
template <class Cx>
MPC_HD void wr_rows_to_dense(const Cx& cx, const Work& k, double mu_inv) {
  const int nv = k.sc->nv;
  MPC_FOR(i, nv) {
    const int j = i / 3, c = i - 3 * j;
    const double* rc = k.rcat + 6 * j;
    double v;
    if (c == 0) v = (rc[0] - rc[1]) * mu_inv;
    else if (c == 1) v = (rc[2] - rc[3]) * mu_inv;
    else v = ((rc[0] + rc[1]) + (rc[2] + rc[3])) + (rc[4] - rc[5]);
    k.wv[i] = v;
  }
  cx.sync();
}

// ---------------------------------------------------------------------------
// Stage 3: Goldfarb-Idnani dual active-set iterations on the explicit inverse.
//   x  = -Minv g (unconstrained optimum), working set W empty, duals u = 0;
//   repeat: pick the most violated row p; move along z = Minv(n_p - N r),
//   r = T N' Minv n_p with T = (N' Minv N)^{-1} kept explicitly (bordered on an
//   add, Schur-downdated on a drop), until p is satisfied or blocked duals leave.
// Returns through sc->status.  Exactness: the loop ends only when no row is
// violated by more than `vtol`, every dual is >= 0 by construction and x is the
// stationary point of its working set, i.e. the KKT point of the strictly convex QP.
// ---------------------------------------------------------------------------
// Start of stage 3 (all threads): unconstrained optimum x = -Minv g, empty working set.
template <class Cx>
MPC_HD void active_set_init(const Cx& cx, const float* rec, const unsigned char* gait, const Work& k,
                            bool have_x = false) {
  Scalars* sc = k.sc;
  const int nv = sc->nv, ld = k.ld, ns = sc->ns;
  const double* Hm = k.Hm;
  if constexpr (Cx::kWrench) {  // x = -H^{-1} g through the rank structure
    wr_apply(cx, k, k.g, k.x);
    MPC_FOR(i, nv) k.x[i] = -k.x[i];
    MPC_FOR(c, 6 * ns) k.rcat[c] = 0.0;  // per-row coefficients: all zero between uses
    cx.sync();
  }
  if (!Cx::kWrench && !have_x) MPC_FOR(i, nv) {  // four independent partial sums: the chain is latency-, not throughput-bound
    double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
    int j = 0;
#pragma unroll 1
    for (; j + 3 < nv; j += 4) {
      a0 += Hm[hixT<Cx::kPacked>(ld, j, i)] * k.g[j];
      a1 += Hm[hixT<Cx::kPacked>(ld, (j + 1), i)] * k.g[j + 1];
      a2 += Hm[hixT<Cx::kPacked>(ld, (j + 2), i)] * k.g[j + 2];
      a3 += Hm[hixT<Cx::kPacked>(ld, (j + 3), i)] * k.g[j + 3];
    }
#pragma unroll 1
    for (; j < nv; j++) a0 += Hm[hixT<Cx::kPacked>(ld, j, i)] * k.g[j];
    k.x[i] = -((a0 + a1) + (a2 + a3));
  }
  MPC_FOR(j, ns) {
    k.amask[j] = 0;
    k.ub[j] = (double)((float)gait[k.stance[j]] * rec[MPC_REC_FMAX]);  // U_b(5k+4), a float product upstream
  }
  cx.sync();
  MPC_STAMP(k, cx, 12);
}

template <class Cx>
MPC_HD void active_set(const Cx& cx, const float* rec, const unsigned char* gait, const Work& k, int max_iter) {
  (void)gait;
  Scalars* sc = k.sc;
  const int nv = sc->nv, ns = sc->ns, ld = k.ld, ldT = k.ldT;
  // 1/mu as the reference forms it: a float (f_block is fpt = float, SolverMPC.cpp:361-372, and A_red goes
  // through a float temporary, :516), e.g. exactly 2.5 for mu = 0.4f rather than 2.49999996...
  const double mu_inv = (double)(1.0f / rec[MPC_REC_MU]);
  const double* Hm = k.Hm;
  double* T = k.T;
  const double vtol = 1e-9;

  for (;;) {
    // ---- most violated row: one stance pair per thread, its six slacks from (fx, fy, fz) ----
    double best = -vtol;
    int bidx = 0x7fffffff;
    MPC_FOR(j, ns) {
      const double fx = k.x[3 * j], fy = k.x[3 * j + 1], fz = k.x[3 * j + 2];
      const int mask = k.amask[j];
      const double sl[6] = {fx * mu_inv + fz, fz - fx * mu_inv, fy * mu_inv + fz, fz - fy * mu_inv, fz, k.ub[j] - fz};
#pragma unroll
      for (int t = 0; t < 6; t++)
        if (!((mask >> t) & 1) && sl[t] < best) { best = sl[t]; bidx = 6 * j + t; }
    }
    block_argmin(cx, k.red, best, bidx);
    if (sc->iters == 0) MPC_STAMP(k, cx, 13);
    if (bidx == 0x7fffffff) break;  // uniform
    if (sc->iters >= max_iter) {
      cx.sync();
      MPC_ONE sc->status = MPC_STATUS_MAX_ITER;
      cx.sync();
      break;
    }
    const int p = bidx;
    const Row rp = make_row(p, mu_inv);
    const double bp = (p % 6 == 5) ? -k.ub[p / 6] : 0.0;
    cx.sync();
    MPC_ONE { sc->iters++; sc->up = 0.0; }
    cx.sync();
    if constexpr (Cx::kWrench) wr_apply_row(cx, k, rp, k.g);  // zp = H^{-1} n_p, kept in g (free once x0 exists)
    // ---- inner loop: partial steps drop blocking rows until p can be added ----
    bool fail = false;
    for (;;) {
      const int m = sc->m;
      // w_a = n_a' Minv n_p,  vnp = n_p' Minv n_p
      double vnp;
      if constexpr (Cx::kWrench) {
        MPC_FOR(a, m) k.w[a] = k.Wca[a] * k.g[k.Wia[a]] + k.Wcz[a] * k.g[k.Wiz[a]];
        vnp = rp.ca * k.g[rp.ia] + rp.cz * k.g[rp.iz];
      } else {
        MPC_FOR(a, m) {
          Row ra;
          ra.ia = k.Wia[a]; ra.iz = k.Wiz[a]; ra.ca = k.Wca[a]; ra.cz = k.Wcz[a];
          k.w[a] = row_minv_row<Cx::kPacked>(Hm, ld, ra, rp);
        }
        vnp = row_minv_row<Cx::kPacked>(Hm, ld, rp, rp);
      }
      cx.sync();
      // r = T w   (T symmetric: walk columns for contiguous reads)
      MPC_FOR(a, m) {
        double acc = 0;
#pragma unroll(Cx::kUnroll)
        for (int b = 0; b < m; b++) acc += T[b * ldT + a] * k.w[b];
        k.r[a] = acc;
      }
      cx.sync();
      // znp = vnp - w.r ; t1 = min_{r_a > 0} u_a / r_a
      double part = 0, tbest = 1e300;
      int tidx = 0x7fffffff;
      MPC_FOR(a, m) {
        part += k.w[a] * k.r[a];
        if (k.r[a] > 0.0) {
          const double q = k.u[a] / k.r[a];
          if (q < tbest) { tbest = q; tidx = a; }
        }
      }
      double wr = 0.0;
      if (m > 0) {  // uniform
        wr = block_sum(cx, k.red, part);
        block_argmin(cx, k.red, tbest, tidx);
      }
      const double znp = vnp - wr;
      const bool dependent = !(znp > 1e-11 * vnp);
      const double spc = rp.ca * k.x[rp.ia] + rp.cz * k.x[rp.iz] - bp;  // current slack of p (< 0)
      const double t2 = dependent ? 1e300 : -spc / znp;
      const double t1 = (tidx == 0x7fffffff) ? 1e300 : tbest;
      const double t = t1 < t2 ? t1 : t2;
      if (t >= 1e300) { fail = true; break; }  // infeasible (cannot happen: f = 0 is feasible)
      cx.sync();  // every thread has read x (slack of p) before anybody moves x
      // x += t * Minv (n_p - N r): every row touches <= 2 variables, so z is a combination of at most 2(m+1)
      // rows of Minv.  Applied in the (numerically) dependent case too: there z is only round-off-small, not
      // zero, and x and u must move with the same (r, t) for stationarity x = -Minv (g - N u) to survive.
      if constexpr (Cx::kWrench) {  // z = H^{-1} (n_p - N r) through the rank structure; rcat is all zero between uses
        MPC_FOR(a, m) k.rcat[k.W[a]] = -k.r[a];
        MPC_ONE k.rcat[p] = 1.0;
        cx.sync();
        wr_rows_to_dense(cx, k, mu_inv);
        MPC_FOR(a, m) k.rcat[k.W[a]] = 0.0;
        MPC_ONE k.rcat[p] = 0.0;
        wr_apply(cx, k, k.wv, k.wz);
        MPC_FOR(i, nv) k.x[i] += t * k.wz[i];
      } else {
        MPC_FOR(i, nv) {
          double acc0 = rp.cz * Hm[hixT<Cx::kPacked>(ld, rp.iz, i)], acc1 = rp.ca * Hm[hixT<Cx::kPacked>(ld, rp.ia, i)];
#pragma unroll(Cx::kUnroll)
          for (int a = 0; a < m; a++) {
            const double ra = k.r[a];
            acc0 -= ra * k.Wcz[a] * Hm[hixT<Cx::kPacked>(ld, k.Wiz[a], i)];
            acc1 -= ra * k.Wca[a] * Hm[hixT<Cx::kPacked>(ld, k.Wia[a], i)];
          }
          k.x[i] += t * (acc0 + acc1);
        }
      }
      MPC_FOR(a, m) k.u[a] -= t * k.r[a];
      cx.sync();
      if (t2 <= t1) {
        // ---- full step: p joins the working set; border T ----
        if (m >= k.m_cap) {
          MPC_ONE sc->status = STATUS_RETRY_BIG;
          cx.sync();
          return;
        }
        const double dinv = 1.0 / znp;
#pragma unroll 1
        for (int e = cx.tid; e < m * m; e += cx.nt) {
          const int a = e / m, b = e - a * m;
          T[a * ldT + b] += k.r[a] * k.r[b] * dinv;
        }
        MPC_FOR(a, m) {
          T[a * ldT + m] = -k.r[a] * dinv;
          T[m * ldT + a] = -k.r[a] * dinv;
        }
        MPC_ONE {
          T[m * ldT + m] = dinv;
          k.W[m] = p;
          k.Wia[m] = rp.ia; k.Wiz[m] = rp.iz; k.Wca[m] = rp.ca; k.Wcz[m] = rp.cz;
          k.u[m] = sc->up + t;
          k.amask[p / 6] |= 1 << (p % 6);
          sc->m = m + 1;
        }
        cx.sync();
        break;
      }
      // ---- partial step: row W[tidx] leaves; Schur-downdate T, move the last row into the hole ----
      const int a0 = tidx, last = m - 1;
      MPC_FOR(a, m) k.tcol[a] = T[a0 * ldT + a];
      cx.sync();
      const double taa_inv = 1.0 / k.tcol[a0];
#pragma unroll 1
      for (int e = cx.tid; e < m * m; e += cx.nt) {
        const int a = e / m, b = e - a * m;
        if (a != a0 && b != a0) T[a * ldT + b] -= k.tcol[a] * k.tcol[b] * taa_inv;
      }
      cx.sync();
      if (a0 != last) {
        MPC_FOR(b, last) {
          if (b == a0) continue;
          const double v = T[last * ldT + b];
          T[a0 * ldT + b] = v;
          T[b * ldT + a0] = v;
        }
        MPC_ONE T[a0 * ldT + a0] = T[last * ldT + last];
      }
      cx.sync();
      MPC_ONE {
        sc->up += t;
        const int cdrop = k.W[a0];
        k.amask[cdrop / 6] &= ~(1 << (cdrop % 6));
        if (a0 != last) {
          k.W[a0] = k.W[last];
          k.Wia[a0] = k.Wia[last]; k.Wiz[a0] = k.Wiz[last]; k.Wca[a0] = k.Wca[last]; k.Wcz[a0] = k.Wcz[last];
          k.u[a0] = k.u[last];
        }
        sc->m = last;
      }
      cx.sync();
    }
    if (sc->iters == 1) MPC_STAMP(k, cx, 14);
    if (fail) {
      cx.sync();
      MPC_ONE sc->status = MPC_STATUS_MAX_ITER;
      cx.sync();
      break;
    }
  }
  // ---- polish: x and u are always updated with the same r, so stationarity x = -Minv (g - N u) holds to
  // round-off whatever the accuracy of the explicitly updated T; what drifts with T is the feasibility
  // n_a'x = b_a of the working set.  Two passes of iterative refinement with T as the approximate inverse of
  // S = N'Minv N (du = T (b - N'x); u += du; x += Minv N du) restore it to round-off -- when it is off at all.
  const int m = sc->m;
  if (sc->status == MPC_STATUS_OPTIMAL && m > 0) {
    for (int pass = 0; pass < 2; pass++) {
      double worst = 0.0;
      MPC_FOR(a, m) {
        const double b = (k.W[a] % 6 == 5) ? -k.ub[k.W[a] / 6] : 0.0;
        const double wa = b - (k.Wca[a] * k.x[k.Wia[a]] + k.Wcz[a] * k.x[k.Wiz[a]]);
        k.w[a] = wa;
        worst = fabs(wa) > worst ? fabs(wa) : worst;
      }
      // A pass is only worth its barriers when the working set is off by more than round-off: measured over the five
      // BASELINE workloads the residual is <= 2e-12 N before the first pass and 1.4e-14 after it, and the solution's
      // distance to the reference solver does not change in the third digit with zero, one or two passes.
      {
        double neg = -worst;
        int who = cx.tid;
        block_argmin(cx, k.red, neg, who);
        if (!(-neg > 1e-12)) break;  // uniform
      }
      cx.sync();
      MPC_FOR(a, m) {
        double acc = 0;
#pragma unroll(Cx::kUnroll)
        for (int b = 0; b < m; b++) acc += T[b * ldT + a] * k.w[b];
        k.r[a] = acc;
      }
      cx.sync();
      if constexpr (Cx::kWrench) {
        MPC_FOR(a, m) k.rcat[k.W[a]] = k.r[a];
        cx.sync();
        wr_rows_to_dense(cx, k, mu_inv);
        MPC_FOR(a, m) k.rcat[k.W[a]] = 0.0;
        wr_apply(cx, k, k.wv, k.wz);
        MPC_FOR(i, nv) k.x[i] += k.wz[i];
      } else
      MPC_FOR(i, nv) {
        double acc0 = 0, acc1 = 0;
#pragma unroll(Cx::kUnroll)
        for (int a = 0; a < m; a++) {
          const double ra = k.r[a];
          acc0 += ra * k.Wcz[a] * Hm[hixT<Cx::kPacked>(ld, k.Wiz[a], i)];
          acc1 += ra * k.Wca[a] * Hm[hixT<Cx::kPacked>(ld, k.Wia[a], i)];
        }
        k.x[i] += acc0 + acc1;
      }
      MPC_FOR(a, m) k.u[a] += k.r[a];
      cx.sync();
    }
  }
}

// ---------------------------------------------------------------------------
// Warm start (the reference cold-starts every tick: `QProblem problem_red` is constructed per call,
// SolverMPC.cpp:529, nWSR = 100 :435).  In a closed-loop rollout consecutive problems of a robot differ by one shifted
// horizon step, and so do their optimal working sets.  The engine keeps, per robot, the working set of its last solve
// as (step, leg, row type) codes; the next solve shifts the codes in time, maps them onto its own stance pairs and
// starts the dual active-set method from the KKT point of that working set instead of from the unconstrained optimum:
//   S = N' Minv N (four look-ups per entry) is inverted directly, u = S^{-1}(b - N'x0), x = x0 + Minv N u.
// The Goldfarb-Idnani invariant -- x optimal for its working set, all duals >= 0 -- must hold before the main loop
// may continue from there, so rows whose dual comes out negative are removed and the set is re-solved (a few
// rounds; a singular S or too many rounds fall back to the cold start).  The main loop then adds whatever is still
// violated and ends at the same unique optimum as the cold start.
// ---------------------------------------------------------------------------
constexpr int kWarmStride = 128;  // ints per robot in the cache: count + up to 127 codes ((step*4+leg)*6 + type)

// On entry: x = unconstrained optimum, working set empty (active_set_init).  codes: the robot's cached working set.
template <class Cx>
MPC_HD void active_set_warm(const Cx& cx, const float* rec, const Work& k, const int* codes, int shift) {
  Scalars* sc = k.sc;
  const int nv = sc->nv, ld = k.ld, ldT = k.ldT, h = k.h;
  const double mu_inv = (double)(1.0f / rec[MPC_REC_MU]);
  const double* Hm = k.Hm;
  double* T = k.T;
  MPC_ONE {  // shift the cached rows in time and map them onto this problem's stance pairs
    int m = 0;
    int nc = codes[0];
    if (nc < 0) nc = 0;
    if (nc > kWarmStride - 1) nc = kWarmStride - 1;
    for (int c = 0; c < nc && m < k.m_cap; c++) {
      const int code = codes[1 + c];
      if (code < 0) continue;
      const int kk = code / 6 - 4 * shift, t = code % 6;
      if (kk < 0 || kk >= 4 * h) continue;
      const int j = k.posk[kk];
      if (j < 0 || ((k.amask[j] >> t) & 1)) continue;  // the leg swings at that step now / duplicate
      const Row r = make_row(6 * j + t, mu_inv);
      k.W[m] = 6 * j + t;
      k.Wia[m] = r.ia; k.Wiz[m] = r.iz; k.Wca[m] = r.ca; k.Wcz[m] = r.cz;
      k.amask[j] |= 1 << t;
      m++;
    }
    sc->m = m;
  }
  cx.sync();
  bool accepted = false;
  for (int round = 0; round < 4; round++) {
    const int m = sc->m;
    if (m == 0) break;  // uniform
    // S = N' Minv N into T, the right-hand side b - N'x0 into w, S's diagonal into u (singularity scale)
#pragma unroll 1
    for (int e = cx.tid; e < m * m; e += cx.nt) {
      const int a = e / m, b = e - a * m;
      Row ra, rb;
      ra.ia = k.Wia[a]; ra.iz = k.Wiz[a]; ra.ca = k.Wca[a]; ra.cz = k.Wcz[a];
      rb.ia = k.Wia[b]; rb.iz = k.Wiz[b]; rb.ca = k.Wca[b]; rb.cz = k.Wcz[b];
      const double v = row_minv_row<Cx::kPacked>(Hm, ld, ra, rb);
      T[a * ldT + b] = v;
      if (a == b) k.u[a] = v;
    }
    MPC_FOR(a, m) {
      const double b = (k.W[a] % 6 == 5) ? -k.ub[k.W[a] / 6] : 0.0;
      k.w[a] = b - (k.Wca[a] * k.x[k.Wia[a]] + k.Wcz[a] * k.x[k.Wiz[a]]);
    }
    cx.sync();
    // T <- S^{-1} in place (Gauss-Jordan; S is positive definite for an independent working set, no pivoting)
    bool singular = false;
#pragma unroll 1
    for (int p = 0; p < m; p++) {
      const double d = T[p * ldT + p];
      if (!(d > 1e-11 * k.u[p])) { singular = true; break; }  // uniform: every thread reads the same numbers
      const double dinv = 1.0 / d;
      MPC_FOR(a, m) k.tcol[a] = T[a * ldT + p];                 // old column p
      cx.sync();
      MPC_FOR(b, m) {                                           // new row p
        const double v = (b == p) ? dinv : T[p * ldT + b] * dinv;
        k.r[b] = v;
      }
      cx.sync();
#pragma unroll 1
      for (int e = cx.tid; e < m * m; e += cx.nt) {
        const int a = e / m, b = e - a * m;
        double v;
        if (a == p) v = k.r[b];
        else if (b == p) v = -k.tcol[a] * dinv;
        else v = T[a * ldT + b] - k.tcol[a] * k.r[b];
        T[a * ldT + b] = v;
      }
      cx.sync();
    }
    if (singular) break;
    // u = T (b - N'x0); every dual must be >= 0 for the main loop to take over
    double worst = 0.0;
    int widx = 0x7fffffff;
    MPC_FOR(a, m) {
      double acc = 0;
#pragma unroll(Cx::kUnroll)
      for (int b = 0; b < m; b++) acc += T[a * ldT + b] * k.w[b];
      k.r[a] = acc;
      if (acc < worst) { worst = acc; widx = a; }
    }
    block_argmin(cx, k.red, worst, widx);
    cx.sync();
    if (widx == 0x7fffffff) { accepted = true; break; }
    MPC_ONE {  // remove the rows with negative duals, keep the order of the others
      int mm = 0;
      for (int a = 0; a < m; a++) {
        const int c = k.W[a];
        if (k.r[a] < 0.0) { k.amask[c / 6] &= ~(1 << (c % 6)); continue; }
        if (mm != a) {
          k.W[mm] = c;
          k.Wia[mm] = k.Wia[a]; k.Wiz[mm] = k.Wiz[a]; k.Wca[mm] = k.Wca[a]; k.Wcz[mm] = k.Wcz[a];
        }
        mm++;
      }
      sc->m = mm;
    }
    cx.sync();
  }
  if (!accepted) {  // nothing usable: cold start
    const int m = sc->m;
    cx.sync();
    MPC_ONE {
      for (int a = 0; a < m; a++) k.amask[k.W[a] / 6] &= ~(1 << (k.W[a] % 6));
      sc->m = 0;
    }
    cx.sync();
    return;
  }
  const int m = sc->m;
  MPC_FOR(a, m) k.u[a] = k.r[a];
  MPC_FOR(i, nv) {  // x = x0 + Minv N u
    double acc0 = 0, acc1 = 0;
#pragma unroll(Cx::kUnroll)
    for (int a = 0; a < m; a++) {
      const double ua = k.r[a];
      acc0 += ua * k.Wcz[a] * Hm[hixT<Cx::kPacked>(ld, k.Wiz[a], i)];
      acc1 += ua * k.Wca[a] * Hm[hixT<Cx::kPacked>(ld, k.Wia[a], i)];
    }
    k.x[i] += acc0 + acc1;
  }
  cx.sync();
}

// After the solve: the robot's working set for its next tick, as (step, leg, type) codes.
template <class Cx>
MPC_HD void active_set_store(const Cx& cx, const Work& k, int* codes) {
  MPC_ONE {
    const Scalars* sc = k.sc;
    int m = (sc->status == MPC_STATUS_OPTIMAL) ? sc->m : 0;
    if (m > kWarmStride - 1) m = kWarmStride - 1;
    codes[0] = m;
    for (int a = 0; a < m; a++) codes[1 + a] = k.stance[k.W[a] / 6] * 6 + k.W[a] % 6;
  }
}

// ---------------------------------------------------------------------------
// Stage 4: scatter (SolverMPC.cpp:545-557): eliminated variables are exactly 0.
//   forces   [12] fp32  = q_soln[0..11] (what get_solution(0..11) hands the caller)
//   solution [12h] fp64 = q_soln (optional)
// On any failure status -- the iteration cap included -- the forces are zero (the reference would return stale memory).
// ---------------------------------------------------------------------------
template <class Cx>
MPC_HD void scatter(const Cx& cx, const Work& k, float* forces, double* solution, int32_t* status) {
  const Scalars* sc = k.sc;
  const int code = sc->status;
  // MAX_ITER: a dual active-set method walks through primal-INFEASIBLE points (outside the friction cone, negative
  // fz) until it terminates, so an iterate cut short must not be handed out as forces: zeros, and the flag
  const bool ok = code == MPC_STATUS_OPTIMAL;
  MPC_FOR(i, 12) {
    const int pos = k.posk[i / 3];
    forces[i] = (ok && pos >= 0) ? (float)k.x[3 * pos + (i % 3)] : 0.f;
  }
  if (solution) {
    MPC_FOR(i, 12 * k.h) {
      const int pos = (code == MPC_STATUS_BAD_INPUT) ? -1 : k.posk[i / 3];
      solution[i] = (ok && pos >= 0) ? k.x[3 * pos + (i % 3)] : 0.0;
    }
  }
  MPC_ONE {
    if (status) *status = (code & 0xff) | (sc->iters << 8);
  }
}

// One problem, start to finish.  Returns the status code (uniform across the CTA).
template <class Cx>
MPC_HD int solve_problem(const Cx& cx, const float* rec, const unsigned char* gait, const Work& k, int max_iter) {
  assemble(cx, rec, gait, k);
  if (k.sc->status != MPC_STATUS_OPTIMAL) return k.sc->status;
  invert_spd(cx, k);
  if (k.sc->status != MPC_STATUS_OPTIMAL) return k.sc->status;
  active_set_init(cx, rec, gait, k);
  active_set(cx, rec, gait, k, max_iter);
  return k.sc->status;
}

}  // namespace mpc
#endif
