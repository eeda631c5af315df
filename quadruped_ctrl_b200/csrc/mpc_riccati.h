// Riccati solver of the convex-MPC QP: the same optimum as mpc_core.h's explicit-inverse route, without ever
// forming the condensed Hessian.
//
// What it replaces in the reference is the same as mpc_core.h (SolverMPC.cpp:296-557: c2qp, qH / qg, swing-leg
// elimination, qpOASES); what changes is the algebra.  The condensed QP of solve_mpc,
//     min 1/2 u'Hu + g'u,   H = 2 (B_qp' S B_qp + alpha I),   g = 2 B_qp' S (A_qp x0 - X_d)   (SolverMPC.cpp:395-399)
// is the finite-horizon LQ tracking problem
//     min sum_{k=1..h} (x_k - xd_k)' Q (x_k - xd_k) + alpha sum_{k<h} |u_k|^2,   x_{k+1} = A_d x_k + B_k u_k + a
// over the stance forces u_k of every horizon step (B_k = the stance columns of B_d; a = gravity's column of A_d).
// A product H^{-1} v is therefore one backward and one forward sweep over the horizon with the gains of the
// discrete Riccati recursion (12 x 12 state, at most 12 controls per step):
//     S_k = alpha I + B_k' P_{k+1} B_k,  K_k = S_k^{-1} B_k' P_{k+1} A,  P_k = Q + A' P_{k+1} A - (B_k' P_{k+1} A)' K_k
// -- O(h) work of 12 x 12 matrix products instead of the O((3 * stance)^3) inversion of the condensed Hessian, and
// it only gets cheaper relative to that with longer horizons.  The Goldfarb-Idnani dual active-set method needs
// exactly such products: x = -H^{-1} g once (its backward sweep rides along with the factorisation) and
// d = H^{-1} n_p for every row p that enters the working set (n_p touches one stance foot of one step, so its
// backward sweep starts at that step).  The columns Z = H^{-1} N of the working set are kept, the inverse Schur
// complement T = (N'Z)^{-1} is bordered / down-dated exactly as in mpc_core.h's active_set, and a step direction is
// z = d - Z r.  On the BASELINE workloads a problem changes its working set 2 (trot, horizon 10) to 14 times
// (horizon 20, mixed gaits), so a solve is one factorisation and a handful of sweeps.
//
// Execution model: ONE WARP PER PROBLEM (no CTA barrier anywhere; __syncwarp only).  The algorithm is written once
// against the same tiny execution context as mpc_core.h (ric_setup, ric_factor, ric_forward, ric_hinv_row,
// ric_active_set: strided loops, one barrier per phase), and that generic form is what the single-thread host build of
// tests/emu/ runs and what MPC_RIC_GENERIC=1 runs on the device.  The production kernel replaces the two hot parts by
// device-only forms that compute the same quantities:
//   ric_factor_mma      the factorisation on the FP64 tensor pipe (DMMA.8x8x4 on 8x8 tiles, the tracking sweep riding
//                       along in the tile padding, S swept in registers),
//   ric_forward_fast /  the sweeps with the travelling 12-vector in registers (lane i owns component i), two phases
//   ric_hinv_row_fast   of one __syncwarp per horizon step.
// Working sets that outgrow the fast-memory tile of Z move into a per-warp global slab and carry on (ric_active_set).
// Exactness: A_d = I + dt A + dt^2/2 A^2 and B_d = dt B + dt^2/2 AB + dt^3/6 A^2 B are exact (A^3 = 0, see
// mpc_core.h); results agree with the reference's qpOASES on the fp64-assembled dense QP to ~1e-13 relative
// (tests/test_riccati.py on the host build, tests/test_gpu_parity.py on the kernel, both solvers).
#ifndef QUADRUPED_MPC_RICCATI_H
#define QUADRUPED_MPC_RICCATI_H

#include "mpc_core.h"

namespace mpc {

struct RicLayout {
  int h, nv_cap, m_cap, ldT, ldz, gain_cap;
  int off_sc, off_ints, off_dyn, off_Bd, off_gain, off_x, off_d, off_kap, off_vec, off_red, off_un;
  int bytes;
  size_t slab_bytes;  // per-warp global slab: Z [nv_cap x nv_cap], T [nv_cap x (nv_cap|1)], six vectors, three index lists
};

// doubles a problem's gains can take: per step 12 n_k (K_k) + n_k (n_k + 1) / 2 (S_k^{-1}, packed), n_k <= 12
// (every step's block starts on an even offset: the DMMA factorisation stores two doubles at a time)
constexpr int ric_gain_doubles(int nv_cap, int h) { return (12 * nv_cap + (13 * nv_cap + 1) / 2 + h + 1) / 2 * 2; }
constexpr int kRicDyn = 8 + 12 + 36 + 36 + 12;  // scalars, Q, column form of N, row form of N, x0
constexpr int kRicVec = 6 * 12;                 // pv, pn, xv, xn, wv, pt

inline RicLayout make_ric_layout(int h, int nv_cap, int m_cap) {
  RicLayout L;
  L.h = h;
  L.nv_cap = nv_cap;
  L.m_cap = m_cap;
  L.ldT = m_cap | 1;
  L.ldz = nv_cap;
  L.gain_cap = ric_gain_doubles(nv_cap, h);
  int o = 0;
  L.off_sc = o;
  o += (int)((sizeof(Scalars) + 15) / 16 * 16);
  // ints: step [h][4] (n_k, voff, koff, swing mask: one 16-byte load per step); stance, posk, amask [4h each]; nk [h];
  // voff, koff [h+1 each]; bcol [nv_cap]; W, Wia, Wiz [m_cap+1 each]; NcI, NrI [36 each]
  L.off_ints = o;
  o += 4 * (4 * h + 12 * h + h + 2 * (h + 1) + nv_cap + 3 * (m_cap + 1) + 72);
  o = (o + 15) / 16 * 16;
  L.off_dyn = o;
  o += 8 * kRicDyn;
  L.off_Bd = o;
  o += 8 * 156;
  L.off_gain = o;
  o += 8 * L.gain_cap;
  L.off_x = o;
  o += 8 * nv_cap;
  L.off_d = o;    // d = H^{-1} n_p shares its place with kap: the forward sweep reads kap[v] and writes d[v] in the
  L.off_kap = o;  // same lane, and kap is dead once the sweep has passed
  o += 8 * nv_cap;
  L.off_vec = o;
  o += 8 * kRicVec;
  L.off_red = o;  // (no reduction scratch: one warp reduces by shuffles, the host emulation has one thread)
  L.off_un = o;
  // union: factorisation scratch (P, M, S | Y 13 x 12 each; G 144; scol 2 x 18) | active set (Z, T, six
  // vectors, ub)
  const int fac = 3 * 156 + 144 + 36;  // (Y takes S^{-1}'s place; the continuous-time B of the set-up overlays P and M)
  const int as = nv_cap * m_cap + m_cap * L.ldT + 6 * (m_cap + 1) + 4 * h;
  o += 8 * (fac > as ? fac : as);
  L.bytes = (o + 15) / 16 * 16;
  L.slab_bytes = ((size_t)8 * ((size_t)nv_cap * nv_cap + (size_t)nv_cap * (nv_cap | 1) + 6 * (nv_cap + 1)) +
                  (size_t)4 * 3 * (nv_cap + 1) + 255) / 256 * 256;
  return L;
}

struct RicWork {
  Scalars* sc;
  int *step, *stance, *posk, *amask, *nk, *voff, *koff, *bcol, *W, *Wia, *Wiz, *NcI, *NrI;
  double *dyn, *Q, *NcV, *NrV, *x0, *Bd, *gain, *x, *d, *kap, *pv, *pn, *xv, *xn, *wv, *pt, *red;
  double *P, *Y, *M, *G, *S, *scol, *Bc;              // factorisation view of the union
  double *Z, *T, *w, *r, *u, *tcol, *Wca, *Wcz, *ub;  // active-set view
  int h, nv_cap, m_cap, ldT, ldz;
  char* slab;  // per-warp global scratch for working sets that outgrow the fast-memory tile (nullptr: none)
};
// dyn[]: 0 dt, 1 cos(yaw), 2 sin(yaw), 3 x_drag, 4 alpha, 5 a[5], 6 a[11] (gravity's column of A_d times x0[12]),
// 7 1/mu (float, as the reference forms it)

MPC_HD RicWork ric_carve(const RicLayout& L, char* fast) {
  RicWork k;
  k.sc = (Scalars*)(fast + L.off_sc);
  int* ip = (int*)(fast + L.off_ints);
  k.step = ip;
  ip += 4 * L.h;
  k.stance = ip;
  k.posk = ip + 4 * L.h;
  k.amask = ip + 8 * L.h;
  k.nk = ip + 12 * L.h;
  k.voff = k.nk + L.h;
  k.koff = k.voff + L.h + 1;
  k.bcol = k.koff + L.h + 1;
  k.W = k.bcol + L.nv_cap;
  k.Wia = k.W + L.m_cap + 1;
  k.Wiz = k.Wia + L.m_cap + 1;
  k.NcI = k.Wiz + L.m_cap + 1;
  k.NrI = k.NcI + 36;
  k.dyn = (double*)(fast + L.off_dyn);
  k.Q = k.dyn + 8;
  k.NcV = k.Q + 12;
  k.NrV = k.NcV + 36;
  k.x0 = k.NrV + 36;
  k.Bd = (double*)(fast + L.off_Bd);
  k.gain = (double*)(fast + L.off_gain);
  k.x = (double*)(fast + L.off_x);
  k.d = (double*)(fast + L.off_d);
  k.kap = (double*)(fast + L.off_kap);
  k.pv = (double*)(fast + L.off_vec);
  k.pn = k.pv + 12;
  k.xv = k.pn + 12;
  k.xn = k.xv + 12;
  k.wv = k.xn + 12;
  k.pt = k.wv + 12;
  k.red = nullptr;
  double* un = (double*)(fast + L.off_un);
  k.P = un;
  k.M = k.P + 156;
  k.S = k.M + 156;  // S^{-1} [12 x 12] until K is formed, then Y = P A [13 x 12]
  k.Y = k.S;
  k.G = k.S + 156;
  k.scol = k.G + 144;
  k.Bc = un;  // 13 x 12, dead before the factorisation initialises P
  k.Z = un;
  k.T = k.Z + L.nv_cap * L.m_cap;
  k.w = k.T + L.m_cap * L.ldT;
  k.r = k.w + L.m_cap + 1;
  k.u = k.r + L.m_cap + 1;
  k.tcol = k.u + L.m_cap + 1;
  k.Wca = k.tcol + L.m_cap + 1;
  k.Wcz = k.Wca + L.m_cap + 1;
  k.ub = k.Wcz + L.m_cap + 1;
  k.h = L.h;
  k.nv_cap = L.nv_cap;
  k.m_cap = L.m_cap;
  k.ldT = L.ldT;
  k.ldz = L.ldz;
  k.slab = nullptr;
  return k;
}

// e / n and the remainder for small non-negative e and 1 <= n <= 12 (a float reciprocal instead of an integer
// division on the device; exact in this range)
MPC_HD int ric_div(int e, int n) {
#if defined(__CUDA_ARCH__)
  return __float2int_rz(__fdividef((float)e + 0.5f, (float)n));
#else
  return e / n;
#endif
}

// ---------------------------------------------------------------------------
// Set-up: input check, stance list, x0, exact discretisation (the P0-P3 phases of mpc_core.h's assemble_front with
// the same formulas), the sparse tables of N = A_d - I, per-step offsets.  Returns through sc->status.
// ---------------------------------------------------------------------------
template <class Cx>
MPC_HD void ric_setup(const Cx& cx, const float* rec, const unsigned char* gait, const RicWork& k) {
  const int h = k.h;
  Scalars* sc = k.sc;
  double* B = k.Bc;  // 13 x 12 continuous-time B
  int* flag = k.amask;
  const float fmax = rec[MPC_REC_FMAX];
  MPC_ONE {
    sc->status = MPC_STATUS_OPTIMAL;
    sc->m = 0;
    sc->iters = 0;
  }
  MPC_FOR(kk, 4 * h) {  // SolverMPC.cpp:441-469: U_b(5k+4) = gait[k]*f_max "near zero" => eliminated
    const float ub = (float)gait[kk] * fmax;
    flag[kk] = ((double)ub < 0.01 && (double)ub > -0.01) ? 0 : 1;
  }
  MPC_FOR(i, 156) B[i] = 0.0;
  MPC_FOR(i, 36) {
    k.NcI[i] = 0; k.NrI[i] = 0;
    k.NcV[i] = 0.0; k.NrV[i] = 0.0;
  }
  cx.sync();
  MPC_FOR(i, MPC_REC_TRAJ + 12 * h)
    if (!finite_f(rec[i])) sc->status = MPC_STATUS_BAD_INPUT;  // same value from every writer
  MPC_ONE {
    if (!(rec[MPC_REC_MU] > 0.f) || !(rec[MPC_REC_MASS] > 0.f) || !(rec[MPC_REC_DT] > 0.f) ||
        !(rec[MPC_REC_IBODY] > 0.f) || !(rec[MPC_REC_IBODY + 1] > 0.f) || !(rec[MPC_REC_IBODY + 2] > 0.f) ||
        !(fmax >= 0.f))
      sc->status = MPC_STATUS_BAD_INPUT;
  }
#if defined(__CUDA_ARCH__)
  if (Cx::kOneWarp) {  // prefix count by ballots, 32 table entries per round
    int base = 0;
#pragma unroll 1
    for (int k0 = 0; k0 < 4 * h; k0 += 32) {
      const int kk = k0 + cx.tid;
      const int f = kk < 4 * h ? flag[kk] : 0;
      const unsigned bal = __ballot_sync(0xffffffffu, f != 0);
      const int pos = base + __popc(bal & ((1u << cx.tid) - 1u));
      if (kk < 4 * h) {
        if (f) { k.stance[pos] = kk; k.posk[kk] = pos; }
        else k.posk[kk] = -1;
      }
      base += __popc(bal);
    }
    MPC_ONE { sc->ns = base; sc->nv = 3 * base; }
  } else
#endif
  MPC_FOR(kk, 4 * h) {
    int pos = 0;
#pragma unroll 4
    for (int q = 0; q < kk; q++) pos += flag[q];
    if (flag[kk]) { k.stance[pos] = kk; k.posk[kk] = pos; }
    else k.posk[kk] = -1;
    if (kk == 4 * h - 1) { sc->ns = pos + flag[kk]; sc->nv = 3 * (pos + flag[kk]); }
  }
  {
    // quat_to_rpy (SolverMPC.cpp:257-267), q = (w,x,y,z); x_0 = [rpy(2), rpy(1), rpy(0), ...] (:318); the
    // transcendental groups on four lanes
    const double qw = rec[MPC_REC_Q], qx = rec[MPC_REC_Q + 1], qy = rec[MPC_REC_Q + 2], qz = rec[MPC_REC_Q + 3];
    const int l1 = cx.nt >= 4 ? 1 : 0, l2 = cx.nt >= 4 ? 2 : 0, l3 = cx.nt >= 4 ? 3 : 0;
    if (cx.tid == 0) {
      const double yaw = (double)rec[MPC_REC_YAW];
      double sy, cy;
      sincos(yaw, &sy, &cy);
      k.dyn[1] = cy;
      k.dyn[2] = sy;
    }
#if defined(__CUDA_ARCH__)
    if (Cx::kOneWarp) {
      // the three angles through ONE atan2 call on lanes 1..3 (a warp runs divergent calls one after the other):
      // asin(s) = atan2(s, sqrt(1 - s^2))
      double as = -2. * (qx * qz - qw * qy);
      if (!(as < .99999)) as = .99999;
      const double yy = cx.tid == 1 ? 2. * (qx * qy + qw * qz) : (cx.tid == 2 ? as : 2. * (qy * qz + qw * qx));
      const double xx = cx.tid == 1 ? qw * qw + qx * qx - qy * qy - qz * qz
                                    : (cx.tid == 2 ? sqrt(fma(-as, as, 1.0)) : qw * qw - qx * qx - qy * qy + qz * qz);
      if (cx.tid >= 1 && cx.tid <= 3) k.x0[3 - cx.tid] = MPC_ATAN2(yy, xx);
    } else
#endif
    {
    if (cx.tid == l1) k.x0[2] = MPC_ATAN2(2. * (qx * qy + qw * qz), qw * qw + qx * qx - qy * qy - qz * qz);
    if (cx.tid == l2) {
      double as = -2. * (qx * qz - qw * qy);
      if (!(as < .99999)) as = .99999;
      k.x0[1] = asin(as);
    }
    if (cx.tid == l3) k.x0[0] = MPC_ATAN2(2. * (qy * qz + qw * qx), qw * qw - qx * qx - qy * qy + qz * qz);
    }
  }
  cx.sync();
  const double dt = (double)rec[MPC_REC_DT];
  const double xd = (double)rec[MPC_REC_XDRAG];
  const double yc = k.dyn[1], ys = k.dyn[2];
  MPC_ONE {
    if (sc->status == MPC_STATUS_OPTIMAL && sc->ns == 0) sc->status = MPC_STATUS_NO_STANCE;
    for (int i = 0; i < 3; i++) {
      k.x0[3 + i] = rec[MPC_REC_P + i];
      k.x0[6 + i] = rec[MPC_REC_W + i];
      k.x0[9 + i] = rec[MPC_REC_V + i];
    }
    const double grav = (double)-9.8f;  // x0[12] (SolverMPC.cpp:318)
    k.dyn[0] = dt;
    k.dyn[3] = xd;
    k.dyn[4] = (double)rec[MPC_REC_ALPHA];
    k.dyn[5] = (dt * dt / 2.0) * grav;  // A_d[5][12] = dt^2/2 (row 5 of A^2 carries e12), A_d[11][12] = dt
    k.dyn[6] = dt * grav;
    k.dyn[7] = (double)(1.0f / rec[MPC_REC_MU]);
    // N = A_d - I = dt A + dt^2/2 A^2 (ct_ss_mats, SolverMPC.cpp:237-244; A^2 has row 5 = x_drag e9' + e12' only).
    // Column form: entries (row, value) of column j; row form: entries (column, value) of row i; three slots each.
    const double half = dt * dt / 2.0;
    int* ci = k.NcI; double* cv = k.NcV; int* ri = k.NrI; double* rv = k.NrV;
    ci[3 * 6 + 0] = 0; cv[3 * 6 + 0] = dt * yc;   ci[3 * 6 + 1] = 1; cv[3 * 6 + 1] = -dt * ys;
    ci[3 * 7 + 0] = 0; cv[3 * 7 + 0] = dt * ys;   ci[3 * 7 + 1] = 1; cv[3 * 7 + 1] = dt * yc;
    ci[3 * 8 + 0] = 2; cv[3 * 8 + 0] = dt;
    ci[3 * 9 + 0] = 3; cv[3 * 9 + 0] = dt;        ci[3 * 9 + 1] = 5; cv[3 * 9 + 1] = half * xd;
    ci[3 * 9 + 2] = 11; cv[3 * 9 + 2] = dt * xd;
    ci[3 * 10 + 0] = 4; cv[3 * 10 + 0] = dt;
    ci[3 * 11 + 0] = 5; cv[3 * 11 + 0] = dt;
    ri[3 * 0 + 0] = 6; rv[3 * 0 + 0] = dt * yc;   ri[3 * 0 + 1] = 7; rv[3 * 0 + 1] = dt * ys;
    ri[3 * 1 + 0] = 6; rv[3 * 1 + 0] = -dt * ys;  ri[3 * 1 + 1] = 7; rv[3 * 1 + 1] = dt * yc;
    ri[3 * 2 + 0] = 8; rv[3 * 2 + 0] = dt;
    ri[3 * 3 + 0] = 9; rv[3 * 3 + 0] = dt;
    ri[3 * 4 + 0] = 10; rv[3 * 4 + 0] = dt;
    ri[3 * 5 + 0] = 11; rv[3 * 5 + 0] = dt;       ri[3 * 5 + 1] = 9; rv[3 * 5 + 1] = half * xd;
    ri[3 * 11 + 0] = 9; rv[3 * 11 + 0] = dt * xd;
  }
  MPC_FOR(i, 12) k.Q[i] = (double)rec[MPC_REC_WEIGHTS + i];
  MPC_FOR(b, 4) {  // B_c per leg (ct_ss_mats, SolverMPC.cpp:235-254; cross_mat :226-233), as in mpc_core.h P2
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
  // per-step offsets: n_k = 3 * (stance legs of step k), voff = first reduced variable, koff = first gain double
  MPC_ONE {
    int v = 0, g = 0;
    for (int s = 0; s < h; s++) {
      const int n = 3 * (flag[4 * s] + flag[4 * s + 1] + flag[4 * s + 2] + flag[4 * s + 3]);
      k.nk[s] = n;
      k.voff[s] = v;
      k.koff[s] = g;
      int swm = 0;  // leg columns (of B_d) whose leg is in swing at this step
      for (int l = 0; l < 4; l++)
        if (!flag[4 * s + l]) swm |= 7 << (3 * l);
      k.step[4 * s] = n; k.step[4 * s + 1] = v; k.step[4 * s + 2] = g; k.step[4 * s + 3] = swm;
      v += n;
      g += (12 * n + n * (n + 1) / 2 + 1) / 2 * 2;
    }
    k.voff[h] = v;
    k.koff[h] = g;
  }
  cx.sync();
  if (sc->status != MPC_STATUS_OPTIMAL) return;
  // exact discretisation B_d = dt B + dt^2/2 AB + dt^3/6 A^2 B (c2qp, SolverMPC.cpp:87-101); row 12 is zero
  MPC_FOR(e, 144) {
    const int i = e / 12, j = e - 12 * i;
    k.Bd[e] = dt * B[e] + (dt * dt / 2.0) * apply_A(B, 12, i, j, yc, ys, xd) +
              (dt * dt * dt / 6.0) * apply_A2(B, 12, i, j, xd);
  }
  MPC_FOR(v, sc->nv) k.bcol[v] = (k.stance[v / 3] & 3) * 3 + (v % 3);
  cx.sync();
}

// ---------------------------------------------------------------------------
// Factorisation: backward Riccati recursion, gains K_k and S_k^{-1} of every step into k.gain, and -- riding along --
// the backward sweep of the tracking problem itself (kap = S^{-1}(B'(p + P a)), the feed-forward of x = -H^{-1} g).
// ---------------------------------------------------------------------------
template <class Cx>
MPC_HD void ric_factor(const Cx& cx, const float* rec, const RicWork& k) {
  const int h = k.h;
  Scalars* sc = k.sc;
  double *P = k.P, *Y = k.Y, *M = k.M, *G = k.G, *S = k.S;
  const double* Bd = k.Bd;
  const double* Q = k.Q;
  const double alpha = k.dyn[4], a5 = k.dyn[5], a11 = k.dyn[6];
  // terminal cost of step h: P = Q, p = -Q xd_h
  MPC_FOR(e, 144) {
    const int i = e / 12, j = e - 12 * i;
    P[e] = (i == j) ? Q[i] : 0.0;
  }
  MPC_FOR(i, 12) k.pv[i] = -Q[i] * (double)rec[MPC_REC_TRAJ + 12 * (h - 1) + i];
  cx.sync();
  bool bad = false;
#pragma unroll 1
  for (int s = h - 1; s >= 0; s--) {
    const int n = k.nk[s], v0 = k.voff[s], ntri = n * (n + 1) / 2;
    double* K = k.gain + k.koff[s];  // n x 12, row-major
    double* Si = K + 12 * n;         // S^{-1}, packed lower triangle
    const int* bc = k.bcol + v0;
    // (1) pt = p + P a (a has two entries);  M = P B_k  (12 x n, leading dimension 12)
    MPC_FOR(i, 12) k.pt[i] = k.pv[i] + a5 * P[i * 12 + 5] + a11 * P[i * 12 + 11];
    MPC_FOR(e, 12 * n) {
      const int i = ric_div(e, n), c = e - i * n, col = bc[c];
      double a0 = 0, a1 = 0;
#pragma unroll
      for (int j = 0; j < 12; j += 2) {
        a0 += P[i * 12 + j] * Bd[j * 12 + col];
        a1 += P[i * 12 + j + 1] * Bd[(j + 1) * 12 + col];
      }
      M[i * 12 + c] = a0 + a1;
    }
    cx.sync();
    // (2) S = alpha I + B_k' M (lower triangle, packed);  w = B_k' pt
    MPC_FOR(e, n * n + n) {
      if (e < n * n) {
        const int a = ric_div(e, n), b = e - a * n;
        if (a >= b) {
          const int ca = bc[a];
          double a0 = 0, a1 = 0;
#pragma unroll
          for (int i = 0; i < 12; i += 2) {
            a0 += Bd[i * 12 + ca] * M[i * 12 + b];
            a1 += Bd[(i + 1) * 12 + ca] * M[(i + 1) * 12 + b];
          }
          S[tri_index(a, b)] = (a0 + a1) + (a == b ? alpha : 0.0);
        }
      } else {
        const int c = e - n * n, col = bc[c];
        double a0 = 0, a1 = 0;
#pragma unroll
        for (int i = 0; i < 12; i += 2) {
          a0 += Bd[i * 12 + col] * k.pt[i];
          a1 += Bd[(i + 1) * 12 + col] * k.pt[i + 1];
        }
        k.wv[c] = a0 + a1;
      }
    }
    cx.sync();
    // (3) S <- -S^{-1} by symmetric sweeps (SPD: no pivoting; a non-positive pivot = alpha <= 0 with zero weights)
#pragma unroll 1
    for (int p = 0; p < n; p++) {
      MPC_FOR(a, n) k.scol[a] = S[tri_index(a, p)];
      cx.sync();
      const double dp = k.scol[p];
      if (!(dp > 0.0) || !(dp < 1e300)) bad = true;  // uniform
      const double dinv = 1.0 / dp;
      MPC_FOR(e, n * n) {
        const int a = ric_div(e, n), b = e - a * n;
        if (a >= b) {
          double v;
          if (a == p && b == p) v = -dinv;
          else if (a == p) v = k.scol[b] * dinv;
          else if (b == p) v = k.scol[a] * dinv;
          else v = S[tri_index(a, b)] - k.scol[a] * k.scol[b] * dinv;
          S[tri_index(a, b)] = v;
        }
      }
      cx.sync();
    }
    // (4) G = M' A = M' + M' N  (n x 12);  S^{-1} into the gains
    MPC_FOR(e, ntri) Si[e] = -S[e];
    cx.sync();  // (Y below takes S's place)
    MPC_FOR(e, 12 * n) {
      const int c = e / 12, j = e - 12 * c;
      double acc = M[j * 12 + c];
#pragma unroll
      for (int t = 0; t < 3; t++) acc += k.NcV[3 * j + t] * M[k.NcI[3 * j + t] * 12 + c];
      G[e] = acc;
    }
    // Y = P A = P + P N
    MPC_FOR(e, 144) {
      const int i = e / 12, j = e - 12 * i;
      double acc = P[e];
#pragma unroll
      for (int t = 0; t < 3; t++) acc += k.NcV[3 * j + t] * P[i * 12 + k.NcI[3 * j + t]];
      Y[e] = acc;
    }
    cx.sync();
    // (5) K = S^{-1} G;  kap = S^{-1} w
    MPC_FOR(e, 12 * n + n) {
      if (e < 12 * n) {
        const int c = e / 12, j = e - 12 * c;
        double acc = 0;
#pragma unroll 1
        for (int b = 0; b < n; b++) acc += Si[tri_index(c, b)] * G[b * 12 + j];
        K[e] = acc;
      } else {
        const int c = e - 12 * n;
        double acc = 0;
#pragma unroll 1
        for (int b = 0; b < n; b++) acc += Si[tri_index(c, b)] * k.wv[b];
        k.kap[v0 + c] = acc;
      }
    }
    cx.sync();
    // (6) P <- Q + A'Y - G'K (lower triangle, mirrored);  p <- A' pt - G' kap - Q xd_s.  (Q and xd only for s >= 1:
    //     x_0 is given, it carries no cost.)  Nothing here reads P or p.
    MPC_FOR(e, 144 + 12) {
      if (e < 144) {
        const int i = e / 12, j = e - 12 * i;
        if (i >= j) {
          double acc = Y[e];
#pragma unroll
          for (int t = 0; t < 3; t++) acc += k.NcV[3 * i + t] * Y[k.NcI[3 * i + t] * 12 + j];
          double g0 = 0, g1 = 0;
          int c = 0;
#pragma unroll 1
          for (; c + 1 < n; c += 2) {
            g0 += G[c * 12 + i] * K[c * 12 + j];
            g1 += G[(c + 1) * 12 + i] * K[(c + 1) * 12 + j];
          }
          if (c < n) g0 += G[c * 12 + i] * K[c * 12 + j];
          acc -= g0 + g1;
          if (i == j && s >= 1) acc += Q[i];
          P[i * 12 + j] = acc;
          P[j * 12 + i] = acc;
        }
      } else {
        const int j = e - 144;
        double acc = k.pt[j];
#pragma unroll
        for (int t = 0; t < 3; t++) acc += k.NcV[3 * j + t] * k.pt[k.NcI[3 * j + t]];
#pragma unroll 1
        for (int c = 0; c < n; c++) acc -= G[c * 12 + j] * k.kap[v0 + c];
        if (s >= 1) acc -= Q[j] * (double)rec[MPC_REC_TRAJ + 12 * (s - 1) + j];
        k.pv[j] = acc;
      }
    }
    cx.sync();
  }
  if (bad) {
    MPC_ONE sc->status = MPC_STATUS_NOT_PD;
    cx.sync();
  }
}

#if defined(__CUDACC__)
// ---------------------------------------------------------------------------
// The same factorisation on the FP64 tensor pipe (device only): every product of a Riccati step is a handful of
// DMMA.8x8x4 (mma.sync.m8n8k4.f64) on 8x8 tiles of the 12 x 12 state / 12 x n control blocks,
//     M = P B_k        S = alpha I + B_k' M        G = M' A        K = S^{-1} G        Y = P A
//     P <- Q + A'Y - G'K
// 46 DMMAs per trot step instead of ~2400 scalar FMAs issued through shared-memory operands.  Fragment layout of
// mma.m8n8k4.f64 (lane = 4*lr + lc): A[row lr][k lc], B[k lc][col lr], C[row lr][cols 2lc, 2lc+1].  All operands
// are row-major with leading dimension 12 in shared memory: 12 = -4 (mod 16), so both fragment patterns
// (rows by lr / k by lc, and k by lc / columns by lr) hit every 8-byte slot of a 128-byte wavefront exactly twice --
// the minimum for 32 x 8 bytes.  Tiles are padded by predication: loads outside a matrix return 0, stores outside
// are dropped.  The backward sweep of the tracking problem rides along in the padding: row 12 of P holds
// pt' = (p + P a)', so row 12 of M = P B_k is w' = (B_k' pt)', row 12 of Y = P A is (A' pt)', column 12 of
// K = S^{-1} [G | w] is kap, and row 12 of the new P comes out as (A' pt - K' w)'.
// S (n <= 12, one or four tiles) is swept in registers, the pivot column broadcast through a double-buffered
// 16-entry shared-memory line (one __syncwarp per pivot).
// ---------------------------------------------------------------------------
__device__ __forceinline__ void ric_dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// One backward step with NTU control tiles (0: no stance leg in this step, 1: n <= 6, 2: n = 9 or 12).
template <int NTU>
__device__ __forceinline__ bool ric_step_mma(const RicWork& k, const float* rec, int s, int lane, const double (&AB)[3][2]) {
  constexpr int NT = NTU > 0 ? NTU : 1;  // array extents (unused when NTU == 0)
  constexpr int KS = NTU == 2 ? 3 : 2;   // k-steps over the controls: 8 (n <= 6; rows 6, 7 of G are exact zeros) or 12
  constexpr int KPAD = 4 * KS;
  const int lr = lane >> 2, lc = lane & 3;
  const int n = k.nk[s], v0 = k.voff[s];
  double *P = k.P, *Y = k.Y, *M = k.M, *G = k.G, *SF = k.S;  // SF: S^{-1}, full storage [KPAD x 12]
  double* K = k.gain + k.koff[s];
  double* Si = K + 12 * n;
  const double* Bd = k.Bd;
  bool bad = false;
  int colB[NT];
#pragma unroll
  for (int tb = 0; tb < NT; tb++) {
    const int c = 8 * tb + lr;
    colB[tb] = (NTU > 0 && c < n) ? k.bcol[v0 + c] : -1;
  }
  // A fragments of P (rows 8t + lr <= 12, k = 4 s3 + lc): shared by M = P B and Y = P A
  double pa[2][3];
#pragma unroll
  for (int t = 0; t < 2; t++)
#pragma unroll
    for (int s3 = 0; s3 < 3; s3++) pa[t][s3] = (8 * t + lr < 13) ? P[(8 * t + lr) * 12 + 4 * s3 + lc] : 0.0;
  // ---- Y = P A (rows 0..12): the products now, off the critical path; stored once S^{-1} has left its place ----
  double y[2][2][2];
#pragma unroll
  for (int t = 0; t < 2; t++)
#pragma unroll
    for (int u = 0; u < 2; u++) {
      y[t][u][0] = y[t][u][1] = 0.0;
#pragma unroll
      for (int s3 = 0; s3 < 3; s3++) ric_dmma(y[t][u][0], y[t][u][1], pa[t][s3], AB[s3][u]);
    }
  if constexpr (NTU > 0) {
    // ---- (1) M = P B_k (rows 0..12: row 12 is w') ----
    double m[2][NT][2];
#pragma unroll
    for (int t = 0; t < 2; t++)
#pragma unroll
      for (int tb = 0; tb < NT; tb++) {
        m[t][tb][0] = m[t][tb][1] = 0.0;
#pragma unroll
        for (int s3 = 0; s3 < 3; s3++) {
          const double b = colB[tb] >= 0 ? Bd[(4 * s3 + lc) * 12 + colB[tb]] : 0.0;
          ric_dmma(m[t][tb][0], m[t][tb][1], pa[t][s3], b);
        }
      }
#pragma unroll
    for (int t = 0; t < 2; t++)
#pragma unroll
      for (int tb = 0; tb < NT; tb++) {
        const int r = 8 * t + lr, c = 8 * tb + 2 * lc;
        if (r < 13 && c < 12) *reinterpret_cast<double2*>(M + r * 12 + c) = make_double2(m[t][tb][0], m[t][tb][1]);
      }
    __syncwarp();
    // ---- (4) G = M' A (rows = controls < KPAD, columns = states) ----
    {
      double g[NT][2][2];
#pragma unroll
      for (int tc = 0; tc < NT; tc++) {
        double ma[3];
#pragma unroll
        for (int s3 = 0; s3 < 3; s3++) ma[s3] = (8 * tc + lr < 12) ? M[(4 * s3 + lc) * 12 + 8 * tc + lr] : 0.0;
#pragma unroll
        for (int u = 0; u < 2; u++) {
          g[tc][u][0] = g[tc][u][1] = 0.0;
#pragma unroll
          for (int s3 = 0; s3 < 3; s3++) ric_dmma(g[tc][u][0], g[tc][u][1], ma[s3], AB[s3][u]);
          const int r = 8 * tc + lr, c = 8 * u + 2 * lc;
          if (r < KPAD && c < 12) *reinterpret_cast<double2*>(G + r * 12 + c) = make_double2(g[tc][u][0], g[tc][u][1]);
        }
      }
    }
    // ---- (2) S = alpha I + B_k' M, all NTU x NTU tiles in registers ----
    double sw[NT][NT][2];
    {
      const double alpha = k.dyn[4];
#pragma unroll
      for (int ta = 0; ta < NT; ta++)
#pragma unroll
        for (int tb = 0; tb < NT; tb++) {
          sw[ta][tb][0] = (ta == tb && lr == 2 * lc) ? alpha : 0.0;
          sw[ta][tb][1] = (ta == tb && lr == 2 * lc + 1) ? alpha : 0.0;
#pragma unroll
          for (int s3 = 0; s3 < 3; s3++) {
            const double a = colB[ta] >= 0 ? Bd[(4 * s3 + lc) * 12 + colB[ta]] : 0.0;
            const double b = (8 * tb + lr < 12) ? M[(4 * s3 + lc) * 12 + 8 * tb + lr] : 0.0;
            ric_dmma(sw[ta][tb][0], sw[ta][tb][1], a, b);
          }
        }
    }
    // ---- (3) sweep: S <- -S^{-1}, pivots 0..n-1 (the padding keeps its alpha diagonal; it only ever meets the zero
    //      rows of G) ----
    // Uniform rank-1 update (as invert_spd_tiles): the pivot column is published with slot p holding d - 1, so that
    // a_rc -= (c_r / d) c_c gives a_rc - c_r c_c / d off the pivot, c_c / d on the pivot row / column and 2 - 1/d at
    // (p, p) -- no per-element case distinction; every swept diagonal entry ends exactly 2 above its true value.
    double* scol = k.scol;  // [2][18]: the pivot column (16 slots) and d (slot 16), double buffered
#pragma unroll 1
    for (int p = 0; p < n; p++) {
      double* cur = scol + 18 * (p & 1);
      const int tp = p >> 3, pl = p & 7;
      if (lc == (pl >> 1)) {
#pragma unroll
        for (int t = 0; t < NT; t++)
#pragma unroll
          for (int tb = 0; tb < NT; tb++)
            if (tb == tp) {
              const double v = (pl & 1) ? sw[t][tb][1] : sw[t][tb][0];
              const bool diag = 8 * t + lr == p;
              cur[8 * t + lr] = diag ? v - 1.0 : v;
              if (diag) cur[16] = v;
            }
      }
      __syncwarp();
      const double dinv = fast_rcp(cur[16]);
      bad = bad || (unsigned)(__double2hiint(dinv) - 0x00100000) >= (unsigned)(0x7E37E43C - 0x00100000);
#pragma unroll
      for (int t = 0; t < NT; t++) {
        const double ur = -cur[8 * t + lr] * dinv;
#pragma unroll
        for (int tb = 0; tb < NT; tb++) {
          const double2 cc = *reinterpret_cast<const double2*>(cur + 8 * tb + 2 * lc);
          sw[t][tb][0] = fma(ur, cc.x, sw[t][tb][0]);
          sw[t][tb][1] = fma(ur, cc.y, sw[t][tb][1]);
        }
      }
    }
    // S^{-1} = -(swept): full copy for the K product, packed lower triangle into the gains
#pragma unroll
    for (int t = 0; t < NT; t++)
#pragma unroll
      for (int tb = 0; tb < NT; tb++)
#pragma unroll
        for (int e = 0; e < 2; e++) {
          const int r = 8 * t + lr, c = 8 * tb + 2 * lc + e;
          const double v = (r == c) ? 2.0 - sw[t][tb][e] : -sw[t][tb][e];
          if (r < KPAD && c < KPAD) SF[r * 12 + c] = (r < n && c < n) ? v : 0.0;
          if (r < n && c <= r) Si[r * (r + 1) / 2 + c] = v;
        }
    __syncwarp();
    // ---- (5) K = S^{-1} [G | w]: columns 0..11 the gain, column 12 kap ----
    {
#pragma unroll
      for (int tc = 0; tc < NT; tc++) {
        double sa[KS];
#pragma unroll
        for (int s2 = 0; s2 < KS; s2++) sa[s2] = (8 * tc + lr < KPAD) ? SF[(8 * tc + lr) * 12 + 4 * s2 + lc] : 0.0;
#pragma unroll
        for (int u = 0; u < 2; u++) {
          double k0 = 0.0, k1 = 0.0;
#pragma unroll
          for (int s2 = 0; s2 < KS; s2++) {
            const int cc = 4 * s2 + lc, j = 8 * u + lr;
            const double b = j < 12 ? G[cc * 12 + j] : (j == 12 ? M[12 * 12 + cc] : 0.0);
            ric_dmma(k0, k1, sa[s2], b);
          }
          const int r = 8 * tc + lr, c = 8 * u + 2 * lc;
          if (r < n && c < 12) *reinterpret_cast<double2*>(K + r * 12 + c) = make_double2(k0, k1);
          if (r < n && c == 12) k.kap[v0 + r] = k0;
        }
      }
    }
    __syncwarp();
  }
  // ---- Y goes into the place S^{-1} has just left (everybody is past the K product) ----
#pragma unroll
  for (int t = 0; t < 2; t++)
#pragma unroll
    for (int u = 0; u < 2; u++) {
      const int r = 8 * t + lr, c = 8 * u + 2 * lc;
      if (r < 13 && c < 12) *reinterpret_cast<double2*>(Y + r * 12 + c) = make_double2(y[t][u][0], y[t][u][1]);
    }
  __syncwarp();
  // ---- (6) P <- Q + A'Y - G'K (rows 0..12; row 12 = (A' pt - K' w)') ----
  {
    double pn[2][2][2];
#pragma unroll
    for (int t = 0; t < 2; t++)
#pragma unroll
      for (int u = 0; u < 2; u++) {
        if (t == 0 && u == 1) continue;  // mirror of (1, 0)
        double c0 = 0.0, c1 = 0.0;
#pragma unroll
        for (int s3 = 0; s3 < 3; s3++) {
          const double b = (8 * u + lr < 12) ? Y[(4 * s3 + lc) * 12 + 8 * u + lr] : 0.0;
          ric_dmma(c0, c1, AB[s3][t], b);
        }
        if constexpr (NTU > 0) {
#pragma unroll
          for (int s2 = 0; s2 < KS; s2++) {
            const int cc = 4 * s2 + lc, i = 8 * t + lr;
            const double a = i < 12 ? -G[cc * 12 + i] : (i == 12 ? -M[12 * 12 + cc] : 0.0);
            const double b = (cc < n && 8 * u + lr < 12) ? K[cc * 12 + 8 * u + lr] : 0.0;
            ric_dmma(c0, c1, a, b);
          }
        }
        pn[t][u][0] = c0;
        pn[t][u][1] = c1;
      }
    __syncwarp();  // everybody is done reading the old P (pa was loaded at the top) -- and Y, G, K of this step
#pragma unroll
    for (int t = 0; t < 2; t++)
#pragma unroll
      for (int u = 0; u < 2; u++) {
        if (t == 0 && u == 1) continue;
        const int r = 8 * t + lr, c = 8 * u + 2 * lc;
        if (r < 13 && c < 12) {
          double v0n = pn[t][u][0], v1n = pn[t][u][1];
          if (r == 12) {  // the identity part of A' on the padding row
            v0n += Y[12 * 12 + c];
            v1n += Y[12 * 12 + c + 1];
          }
          if (s >= 1 && r == c) v0n += k.Q[r];
          if (s >= 1 && r == c + 1) v1n += k.Q[r];
          *reinterpret_cast<double2*>(P + r * 12 + c) = make_double2(v0n, v1n);
          if (t == 1 && u == 0 && r < 12) {
            P[c * 12 + r] = v0n;
            P[(c + 1) * 12 + r] = v1n;
          }
        }
      }
    __syncwarp();
    // p = row 12 - Q xd_s;  pt = p + P a for the next step (a has two entries; P is symmetric)
    if (lane < 12) {
      double pv = P[12 * 12 + lane];
      if (s >= 1) pv -= k.Q[lane] * (double)rec[MPC_REC_TRAJ + 12 * (s - 1) + lane];
      P[12 * 12 + lane] = pv + k.dyn[5] * P[5 * 12 + lane] + k.dyn[6] * P[11 * 12 + lane];
    }
    __syncwarp();
  }
  return bad;
}

// requires the factorisation view of the union to hold P [13 x 12], Y [13 x 12], M [13 x 12], G [12 x 12],
// S [12 x 12] and scol [2 x 18] (make_ric_layout)
__device__ __forceinline__ void ric_factor_mma(const RicWork& k, const float* rec, int lane) {
  const int h = k.h, lr = lane >> 2, lc = lane & 3;
  // B fragments of A = I + N for X A (k = 4 s3 + lc, column 8u + lr); the A fragments of A' are the same numbers
  double AB[3][2];
#pragma unroll
  for (int s3 = 0; s3 < 3; s3++)
#pragma unroll
    for (int u = 0; u < 2; u++) {
      const int l = 4 * s3 + lc, j = 8 * u + lr;
      double v = 0.0;
      if (j < 12) {
        v = (l == j) ? 1.0 : 0.0;
#pragma unroll
        for (int t = 0; t < 3; t++) v += (k.NcI[3 * j + t] == l) ? k.NcV[3 * j + t] : 0.0;
      }
      AB[s3][u] = v;
    }
  // terminal cost: P = Q, p = -Q xd_h, pt = p + P a
  for (int e = lane; e < 13 * 12; e += 32) {
    const int i = e / 12, j = e - 12 * i;
    double v = 0.0;
    if (i < 12) v = (i == j) ? k.Q[i] : 0.0;
    else v = -k.Q[j] * (double)rec[MPC_REC_TRAJ + 12 * (h - 1) + j] + (j == 5 ? k.dyn[5] * k.Q[5] : 0.0) +
             (j == 11 ? k.dyn[6] * k.Q[11] : 0.0);
    k.P[e] = v;
  }
  __syncwarp();
  bool bad = false;
#pragma unroll 1
  for (int s = h - 1; s >= 0; s--) {
    const int n = k.nk[s];
    if (n == 0) bad |= ric_step_mma<0>(k, rec, s, lane, AB);
    else if (n <= 6) bad |= ric_step_mma<1>(k, rec, s, lane, AB);
    else bad |= ric_step_mma<2>(k, rec, s, lane, AB);
  }
  if (__any_sync(0xffffffffu, bad)) {
    if (lane == 0) k.sc->status = MPC_STATUS_NOT_PD;
    __syncwarp();
  }
}
#endif  // __CUDACC__

// Forward sweep: u_k = -K_k x_k - kap_k, x_{k+1} = A x_k + B_k u_k (+ a when `affine`: the tracking problem; the
// products H^{-1} v are homogeneous), from x_0 = xstart (nullptr: 0).  kap_k counts as zero for steps > last.
template <class Cx>
MPC_HD void ric_forward(const Cx& cx, const RicWork& k, const double* xstart, bool affine, int last, double* out) {
  const int h = k.h;
  double* xa = k.xv;
  double* xb = k.xn;
  MPC_FOR(i, 12) xa[i] = xstart ? xstart[i] : 0.0;
  cx.sync();
  const double a5 = affine ? k.dyn[5] : 0.0, a11 = affine ? k.dyn[6] : 0.0;
#pragma unroll 1
  for (int s = 0; s < h; s++) {
    const int n = k.nk[s], v0 = k.voff[s];
    const double* K = k.gain + k.koff[s];
    MPC_FOR(c, n) {
      double a0 = (s <= last) ? k.kap[v0 + c] : 0.0, a1 = 0;
#pragma unroll
      for (int j = 0; j < 12; j += 2) {
        a0 += K[c * 12 + j] * xa[j];
        a1 += K[c * 12 + j + 1] * xa[j + 1];
      }
      out[v0 + c] = -(a0 + a1);
    }
    cx.sync();
    if (s == h - 1) break;  // the state after the last step is not needed
    MPC_FOR(i, 12) {
      double acc = xa[i];
#pragma unroll
      for (int t = 0; t < 3; t++) acc += k.NrV[3 * i + t] * xa[k.NrI[3 * i + t]];
#pragma unroll 1
      for (int c = 0; c < n; c++) acc += k.Bd[i * 12 + k.bcol[v0 + c]] * out[v0 + c];
      if (i == 5) acc += a5;
      if (i == 11) acc += a11;
      xb[i] = acc;
    }
    cx.sync();
    double* t = xa; xa = xb; xb = t;
  }
}

// out = H^{-1} n for a catalogue row n = ca e_ia + cz e_iz (both variables belong to one stance foot of one step):
// argmin 1/2 u'Hu - n'u.  The backward sweep starts at that step (the costate is zero behind it).
template <class Cx>
MPC_HD void ric_hinv_row(const Cx& cx, const RicWork& k, const Row& rp, double* out) {
  const int kp = k.stance[rp.iz / 3] >> 2;
  double* pa = k.pv;
  double* pb = k.pn;
  {
    const int n = k.nk[kp], v0 = k.voff[kp];
    const double* K = k.gain + k.koff[kp];
    const double* Si = K + 12 * n;
    // w = -n/2 on the variables of this step
    MPC_FOR(c, n) k.wv[c] = -0.5 * ((v0 + c == rp.ia ? rp.ca : 0.0) + (v0 + c == rp.iz ? rp.cz : 0.0));
    cx.sync();
    MPC_FOR(e, n + 12) {
      if (e < n) {
        double acc = 0;
#pragma unroll 1
        for (int b = 0; b < n; b++) acc += Si[tri_index(e, b)] * k.wv[b];
        k.kap[v0 + e] = acc;
      } else {
        const int j = e - n;
        double acc = 0;
#pragma unroll 1
        for (int c = 0; c < n; c++) acc -= K[c * 12 + j] * k.wv[c];
        pa[j] = acc;
      }
    }
    cx.sync();
  }
#pragma unroll 1
  for (int s = kp - 1; s >= 0; s--) {
    const int n = k.nk[s], v0 = k.voff[s];
    const double* K = k.gain + k.koff[s];
    const double* Si = K + 12 * n;
    MPC_FOR(c, n) {
      const int col = k.bcol[v0 + c];
      double a0 = 0, a1 = 0;
#pragma unroll
      for (int i = 0; i < 12; i += 2) {
        a0 += k.Bd[i * 12 + col] * pa[i];
        a1 += k.Bd[(i + 1) * 12 + col] * pa[i + 1];
      }
      k.wv[c] = a0 + a1;
    }
    cx.sync();
    MPC_FOR(e, n + 12) {
      if (e < n) {
        double acc = 0;
#pragma unroll 1
        for (int b = 0; b < n; b++) acc += Si[tri_index(e, b)] * k.wv[b];
        k.kap[v0 + e] = acc;
      } else {
        const int j = e - n;
        double acc = pa[j];
#pragma unroll
        for (int t = 0; t < 3; t++) acc += k.NcV[3 * j + t] * pa[k.NcI[3 * j + t]];
#pragma unroll 1
        for (int c = 0; c < n; c++) acc -= K[c * 12 + j] * k.wv[c];
        pb[j] = acc;
      }
    }
    cx.sync();
    double* t = pa; pa = pb; pb = t;
  }
  ric_forward(cx, k, nullptr, false, kp, out);
}

#if defined(__CUDACC__)
// ---------------------------------------------------------------------------
// The sweeps on one warp with the state / costate in registers (device only).  Lane i < 12 owns component i of the
// 12-vector that travels along the horizon, row i of B_d and row i of the sparse part of A (or column i, for the
// costate) in registers; a step is two phases with one __syncwarp each:
//   forward   u = -(K x + kap)   (lanes c < n: one 12-term dot, three chains)   |   x' = A x + B_d u_full
//   backward  w = B_k' p         (lanes c < n)                                  |   kap = S^{-1} w  and  p' = A'p - K'w
// u_full is u scattered into the 12 leg columns (zeros for swing legs), so that the B_d product runs over
// compile-time register indices.
// ---------------------------------------------------------------------------
struct RicLaneConst {
  double brow[12];        // row `lane` of B_d
  int ri[3], ci[3];       // sparse part N = A - I: row form / column form entries of row / column `lane`
  double rv[3], cv[3];
};
__device__ __forceinline__ RicLaneConst ric_lane_const(const RicWork& k, int lane) {
  RicLaneConst L;
  const int i = lane < 12 ? lane : 0;
#pragma unroll
  for (int j = 0; j < 12; j += 2) {
    const double2 v = *reinterpret_cast<const double2*>(k.Bd + i * 12 + j);
    L.brow[j] = v.x;
    L.brow[j + 1] = v.y;
  }
#pragma unroll
  for (int t = 0; t < 3; t++) {
    L.ri[t] = k.NrI[3 * i + t]; L.rv[t] = k.NrV[3 * i + t];
    L.ci[t] = k.NcI[3 * i + t]; L.cv[t] = k.NcV[3 * i + t];
  }
  return L;
}

// out[v] = u of every step, from x_0 = xstart (nullptr: 0); kap_k counts as zero for steps > last; `affine` adds a.
__device__ __forceinline__ void ric_forward_fast(const RicWork& k, const RicLaneConst& L, const double* xstart, bool affine,
                                                 int last, double* out, int lane) {
  const int h = k.h;
  double xi = (lane < 12 && xstart) ? xstart[lane] : 0.0;
  const double ai = !affine ? 0.0 : (lane == 5 ? k.dyn[5] : (lane == 11 ? k.dyn[6] : 0.0));
  double* uf = k.pt;  // u scattered to leg columns
#pragma unroll 1
  for (int s = 0; s < h; s++) {
    const int4 si = *reinterpret_cast<const int4*>(k.step + 4 * s);  // n_k, voff, koff, swing mask
    const int n = si.x, v0 = si.y;
    double* xs = (s & 1) ? k.xn : k.xv;
    if (lane < 12) xs[lane] = xi;
    __syncwarp();
    if ((si.w >> lane) & 1) uf[lane] = 0.0;  // swing leg: no force
    if (lane < n) {
      const double* Kr = k.gain + si.z + lane * 12;
      double a0 = (s <= last) ? k.kap[v0 + lane] : 0.0, a1 = 0.0, a2 = 0.0;
#pragma unroll
      for (int j = 0; j < 12; j += 6) {
        const double2 k0 = *reinterpret_cast<const double2*>(Kr + j), k1 = *reinterpret_cast<const double2*>(Kr + j + 2),
                      k2 = *reinterpret_cast<const double2*>(Kr + j + 4);
        const double2 x0 = *reinterpret_cast<const double2*>(xs + j), x1 = *reinterpret_cast<const double2*>(xs + j + 2),
                      x2 = *reinterpret_cast<const double2*>(xs + j + 4);
        a0 = fma(k0.x, x0.x, a0); a1 = fma(k1.x, x1.x, a1); a2 = fma(k2.x, x2.x, a2);
        a0 = fma(k0.y, x0.y, a0); a1 = fma(k1.y, x1.y, a1); a2 = fma(k2.y, x2.y, a2);
      }
      const double u = -((a0 + a1) + a2);
      out[v0 + lane] = u;
      uf[k.bcol[v0 + lane]] = u;
    }
    if (s == h - 1) break;  // the state after the last step is not needed
    __syncwarp();
    if (lane < 12) {
      double a0 = xi + ai, a1 = 0.0, a2 = 0.0;
#pragma unroll
      for (int t = 0; t < 3; t++) a1 = fma(L.rv[t], xs[L.ri[t]], a1);
#pragma unroll
      for (int j = 0; j < 12; j += 4) {
        const double2 u0 = *reinterpret_cast<const double2*>(uf + j), u1 = *reinterpret_cast<const double2*>(uf + j + 2);
        a0 = fma(L.brow[j], u0.x, a0); a2 = fma(L.brow[j + 2], u1.x, a2);
        a0 = fma(L.brow[j + 1], u0.y, a0); a2 = fma(L.brow[j + 3], u1.y, a2);
      }
      xi = (a0 + a1) + a2;
    }
    // (the next step writes the other xs buffer; uf is rewritten only after the next __syncwarp)
  }
  __syncwarp();
}

// out = H^{-1} n for a catalogue row (see ric_hinv_row)
__device__ __forceinline__ void ric_hinv_row_fast(const RicWork& k, const RicLaneConst& L, const Row& rp, double* out, int lane) {
  const int kp = k.stance[rp.iz / 3] >> 2;
  double pj = 0.0;  // lane j < 12: costate component j
  double* ws = k.wv;
#pragma unroll 1
  for (int s = kp; s >= 0; s--) {
    const int4 si = *reinterpret_cast<const int4*>(k.step + 4 * s);
    const int n = si.x, v0 = si.y;
    const double* K = k.gain + si.z;
    const double* Si = K + 12 * n;
    double* ps = (s & 1) ? k.pn : k.pv;
    if (s == kp) {  // w = -n/2 on the variables of this step
      if (lane < n) ws[lane] = -0.5 * ((v0 + lane == rp.ia ? rp.ca : 0.0) + (v0 + lane == rp.iz ? rp.cz : 0.0));
      if (lane < 12) ps[lane] = 0.0;
    } else {
      if (lane < 12) ps[lane] = pj;
      __syncwarp();
      if (lane < n) {
        const double* Bc = k.Bd + k.bcol[v0 + lane];
        double a0 = 0.0, a1 = 0.0, a2 = 0.0;
#pragma unroll
        for (int i = 0; i < 12; i += 6) {
          const double2 p0 = *reinterpret_cast<const double2*>(ps + i), p1 = *reinterpret_cast<const double2*>(ps + i + 2),
                        p2 = *reinterpret_cast<const double2*>(ps + i + 4);
          a0 = fma(Bc[12 * i], p0.x, a0); a1 = fma(Bc[12 * (i + 2)], p1.x, a1); a2 = fma(Bc[12 * (i + 4)], p2.x, a2);
          a0 = fma(Bc[12 * (i + 1)], p0.y, a0); a1 = fma(Bc[12 * (i + 3)], p1.y, a1); a2 = fma(Bc[12 * (i + 5)], p2.y, a2);
        }
        ws[lane] = (a0 + a1) + a2;
      }
    }
    __syncwarp();
    if (lane < n) {  // kap = S^{-1} w (packed lower triangle, row `lane`)
      double a0 = 0.0, a1 = 0.0;
      const int base = lane * (lane + 1) / 2;
#pragma unroll 1
      for (int b = 0; b < n; b += 2) {
        const int i0 = b <= lane ? base + b : b * (b + 1) / 2 + lane;
        const int b1 = b + 1;
        const int i1 = b1 <= lane ? base + b1 : b1 * (b1 + 1) / 2 + lane;
        a0 = fma(Si[i0], ws[b], a0);
        if (b1 < n) a1 = fma(Si[i1], ws[b1], a1);
      }
      k.kap[v0 + lane] = a0 + a1;
    }
    if (lane < 12) {  // p' = A'p - K'w
      double a0 = pj, a1 = 0.0;
      if (s != kp) {
#pragma unroll
        for (int t = 0; t < 3; t++) a1 = fma(L.cv[t], ps[L.ci[t]], a1);
      }
      const double* Kc = K + lane;
#pragma unroll 1
      for (int c = 0; c < n; c += 3) {  // n is a multiple of 3
        a0 = fma(-Kc[12 * c], ws[c], a0);
        a1 = fma(-Kc[12 * (c + 1)], ws[c + 1], a1);
        a0 = fma(-Kc[12 * (c + 2)], ws[c + 2], a0);
      }
      pj = a0 + a1;
    }
    // ws is rewritten only after the next __syncwarp (the ps store of the next step comes first and goes to the
    // other buffer)
  }
  __syncwarp();
  ric_forward_fast(k, L, nullptr, false, kp, out, lane);
}
#endif  // __CUDACC__

// Warp-wide argmin of (val, idx) pairs, ties to the smaller idx (as block_argmin).  One warp on the device: three
// REDUX instructions on an order-preserving 64-bit integer image of the doubles (high word, low word among the lanes
// that hold the minimal high word, index among the winners) instead of five rounds of 64-bit shuffles and compares.
template <class Cx>
MPC_HD void ric_argmin(const Cx& cx, double& val, int& idx) {
#if defined(__CUDA_ARCH__)
  if constexpr (Cx::kOneWarp) {
    const long long bits = __double_as_longlong(val);
    const unsigned long long key = (unsigned long long)bits ^ (unsigned long long)((bits >> 63) | (long long)0x8000000000000000ull);
    const unsigned hi = (unsigned)(key >> 32), lo = (unsigned)key;
    const unsigned mhi = __reduce_min_sync(0xffffffffu, hi);
    const unsigned mlo = __reduce_min_sync(0xffffffffu, hi == mhi ? lo : 0xffffffffu);
    const bool win = hi == mhi && lo == mlo;
    idx = (int)__reduce_min_sync(0xffffffffu, win ? (unsigned)idx : 0x7fffffffu);
    const unsigned long long mk = ((unsigned long long)mhi << 32) | mlo;
    val = __longlong_as_double((long long)((mk >> 63) ? (mk ^ 0x8000000000000000ull) : ~mk));
    return;
  }
#endif
  block_argmin(cx, (double*)nullptr, val, idx);
}

// ---------------------------------------------------------------------------
// Goldfarb-Idnani dual active set with H^{-1} products from the Riccati sweeps.  Same selection rule, tolerances,
// step rules and T updates as mpc_core.h's active_set; rows of H^{-1} are replaced by the stored columns
// Z[:, a] = H^{-1} n_a of the working set and d = H^{-1} n_p of the entering row.
// ---------------------------------------------------------------------------
template <bool GENERIC, class Cx>
MPC_HD void ric_active_set(const Cx& cx, const float* rec, const unsigned char* gait, const RicWork& kin, int max_iter) {
  constexpr bool generic = GENERIC;
  RicWork k = kin;  // (the working-set pointers move to the slab when the fast-memory tile overflows)
  Scalars* sc = k.sc;
#if defined(__CUDA_ARCH__)
  const RicLaneConst LC = ric_lane_const(k, cx.tid & 31);
#endif
  (void)generic;
  const int nv = sc->nv, ns = sc->ns, ldz = k.ldz;
  int ldT = k.ldT;
  const double mu_inv = k.dyn[7];
  double* T = k.T;
  double* Z = k.Z;
  double* d = k.d;
  const double vtol = 1e-9;
  MPC_FOR(j, ns) {
    k.amask[j] = 0;
    k.ub[j] = (double)((float)gait[k.stance[j]] * rec[MPC_REC_FMAX]);  // U_b(5k+4), a float product upstream
  }
  cx.sync();
  for (;;) {
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
    ric_argmin(cx, best, bidx);
    if (bidx == 0x7fffffff) break;  // uniform
    if (sc->iters >= max_iter) {
      cx.sync();
      MPC_ONE sc->status = MPC_STATUS_MAX_ITER;
      cx.sync();
      break;
    }
    if (sc->m >= k.m_cap) {  // no room for another column of Z in the fast-memory tile
      if (k.slab == nullptr || k.m_cap >= k.nv_cap) {  // no slab: the problem goes to a class with a larger tile
        cx.sync();
        MPC_ONE sc->status = STATUS_RETRY_BIG;
        cx.sync();
        return;
      }
      // Move the working set into this warp's global slab (L2 resident) and carry on with room for nv_cap rows: Z keeps
      // its leading dimension, T is re-strided, the per-row vectors and index lists follow.  Rare by construction
      // (the tile holds the working sets of all but a fraction of a percent of the BASELINE problems).
      const int m = sc->m, nb = k.nv_cap + 1, ldTb = k.nv_cap | 1;
      double* Zb = (double*)k.slab;
      double* Tb = Zb + (size_t)k.nv_cap * ldz;
      double* vb = Tb + (size_t)k.nv_cap * ldTb;   // w, r, u, tcol, Wca, Wcz
      int* ib = (int*)(vb + 6 * nb);               // W, Wia, Wiz
      cx.sync();
#pragma unroll 1
      for (int e = cx.tid; e < m * nv; e += cx.nt) {
        const int a = e / nv, i = e - a * nv;
        Zb[a * ldz + i] = Z[a * ldz + i];
      }
#pragma unroll 1
      for (int e = cx.tid; e < m * m; e += cx.nt) {
        const int a = e / m, b = e - a * m;
        Tb[a * ldTb + b] = T[a * ldT + b];
      }
      MPC_FOR(a, m) {
        vb[2 * nb + a] = k.u[a];
        vb[4 * nb + a] = k.Wca[a];
        vb[5 * nb + a] = k.Wcz[a];
        ib[a] = k.W[a];
        ib[nb + a] = k.Wia[a];
        ib[2 * nb + a] = k.Wiz[a];
      }
      cx.sync();
      Z = k.Z = Zb;
      T = k.T = Tb;
      ldT = k.ldT = ldTb;
      k.w = vb; k.r = vb + nb; k.u = vb + 2 * nb; k.tcol = vb + 3 * nb; k.Wca = vb + 4 * nb; k.Wcz = vb + 5 * nb;
      k.W = ib; k.Wia = ib + nb; k.Wiz = ib + 2 * nb;
      k.m_cap = k.nv_cap;
    }
    const int p = bidx;
    const Row rp = make_row(p, mu_inv);
    const double bp = (p % 6 == 5) ? -k.ub[p / 6] : 0.0;
    cx.sync();
    MPC_ONE { sc->iters++; sc->up = 0.0; }
#if defined(__CUDA_ARCH__)
    if constexpr (Cx::kOneWarp && !GENERIC) ric_hinv_row_fast(k, LC, rp, d, cx.tid);
    else
#endif
    ric_hinv_row(cx, k, rp, d);  // (ends with a sync)
    const double vnp = rp.ca * d[rp.ia] + rp.cz * d[rp.iz];
    bool fail = false;
    for (;;) {
      const int m = sc->m;
      MPC_FOR(a, m) k.w[a] = k.Wca[a] * d[k.Wia[a]] + k.Wcz[a] * d[k.Wiz[a]];  // n_a' H^{-1} n_p
      cx.sync();
      MPC_FOR(a, m) {
        double acc = 0;
#pragma unroll 1
        for (int b = 0; b < m; b++) acc += T[b * ldT + a] * k.w[b];
        k.r[a] = acc;
      }
      cx.sync();
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
        ric_argmin(cx, tbest, tidx);
      }
      const double znp = vnp - wr;
      const bool dependent = !(znp > 1e-11 * vnp);
      const double spc = rp.ca * k.x[rp.ia] + rp.cz * k.x[rp.iz] - bp;  // current slack of p (< 0)
      const double t2 = dependent ? 1e300 : -spc / znp;
      const double t1 = (tidx == 0x7fffffff) ? 1e300 : tbest;
      const double t = t1 < t2 ? t1 : t2;
      if (t >= 1e300) { fail = true; break; }
      cx.sync();  // every thread has read x (slack of p) before anybody moves x
      // x += t (d - Z r); applied in the dependent case too (x and u must move with the same (r, t))
      MPC_FOR(i, nv) {
        double acc0 = d[i], acc1 = 0;
        int a = 0;
#pragma unroll 1
        for (; a + 1 < m; a += 2) {
          acc0 -= k.r[a] * Z[a * ldz + i];
          acc1 -= k.r[a + 1] * Z[(a + 1) * ldz + i];
        }
        if (a < m) acc0 -= k.r[a] * Z[a * ldz + i];
        k.x[i] += t * (acc0 + acc1);
      }
      MPC_FOR(a, m) k.u[a] -= t * k.r[a];
      cx.sync();
      if (t2 <= t1) {
        // ---- full step: p joins the working set; border T, keep d as its column of Z ----
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
        MPC_FOR(i, nv) Z[m * ldz + i] = d[i];
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
      // ---- partial step: row W[tidx] leaves; Schur-downdate T, move the last row / column into the hole ----
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
        MPC_FOR(i, nv) Z[a0 * ldz + i] = Z[last * ldz + i];
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
    if (fail) {
      cx.sync();
      MPC_ONE sc->status = MPC_STATUS_MAX_ITER;
      cx.sync();
      break;
    }
  }
  // ---- polish (as in mpc_core.h): restore the feasibility n_a'x = b_a of the working set when T has drifted ----
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
      {
        double neg = -worst;
        int who = cx.tid;
        ric_argmin(cx, neg, who);
        if (!(-neg > 1e-12)) break;  // uniform
      }
      cx.sync();
      MPC_FOR(a, m) {
        double acc = 0;
#pragma unroll 1
        for (int b = 0; b < m; b++) acc += T[b * ldT + a] * k.w[b];
        k.r[a] = acc;
      }
      cx.sync();
      MPC_FOR(i, nv) {
        double acc = 0;
#pragma unroll 1
        for (int a = 0; a < m; a++) acc += k.r[a] * Z[a * ldz + i];
        k.x[i] += acc;
      }
      MPC_FOR(a, m) k.u[a] += k.r[a];
      cx.sync();
    }
  }
}

// Outputs, as mpc_core.h's scatter (SolverMPC.cpp:545-557): eliminated variables are exactly 0; zeros on failure.
template <class Cx>
MPC_HD void ric_scatter(const Cx& cx, const RicWork& k, float* forces, double* solution, int32_t* status) {
  const Scalars* sc = k.sc;
  const int code = sc->status;
  const bool ok = code == MPC_STATUS_OPTIMAL;
  MPC_FOR(i, 12) {
    const int pos = (code == MPC_STATUS_BAD_INPUT) ? -1 : k.posk[i / 3];
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

// One problem, start to finish (everything but the outputs).  Returns the status code (uniform).
// GENERIC: the scalar factorisation and sweeps (the host emulation; on the device a development switch) instead of the
// tensor-pipe factorisation and the register-resident sweeps.
template <bool GENERIC = true, class Cx>
MPC_HD int ric_solve_problem(const Cx& cx, const float* rec, const unsigned char* gait, const RicWork& k, int max_iter) {
  ric_setup(cx, rec, gait, k);
  if (k.sc->status != MPC_STATUS_OPTIMAL) return k.sc->status;
#if defined(__CUDA_ARCH__)
  if constexpr (Cx::kOneWarp && !GENERIC) ric_factor_mma(k, rec, cx.tid);
  else
#endif
  ric_factor(cx, rec, k);
  if (k.sc->status != MPC_STATUS_OPTIMAL) return k.sc->status;
#if defined(__CUDA_ARCH__)
  if constexpr (Cx::kOneWarp && !GENERIC) ric_forward_fast(k, ric_lane_const(k, cx.tid), k.x0, true, k.h, k.x, cx.tid);
  else
#endif
  ric_forward(cx, k, k.x0, true, k.h, k.x);  // x = -H^{-1} g
  ric_active_set<GENERIC>(cx, rec, gait, k, max_iter);
  return k.sc->status;
}

}  // namespace mpc
#endif
