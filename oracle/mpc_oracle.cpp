// TEST INFRASTRUCTURE -- not part of the product path.  Only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
// may load this library.
//
// CPU restatement of the reference hot path solve_mpc()
//   /root/reference/src/MPC_Ctrl/SolverMPC.cpp:296-557
// (Eigen is not installed in this image, so SolverMPC.cpp itself cannot be
// compiled; this file restates its arithmetic with plain loops, and hands the
// reduced QP to the reference's OWN qpOASES, built unmodified into
// oracle/_ref/libqpoases_ref.so by oracle/Makefile).
//
// Parity pin: the reference ships no test, golden vector or fixture for this
// path, so the QP half is pinned by running the reference's
// solver itself and the assembly half by cross-checking three independent
// formulations (this dense fp32/fp64 restatement, the numpy restatement in
// tests/np_reference.py and the closed-form CUDA assembly).
//
// Two precisions, selected per call:
//   32 -> every assembly quantity in float, as the reference (fpt = float,
//         Utilities/common_types.h:14), QP in double (qpOASES real_t);
//   64 -> same algorithm in double: the rounding-free truth used to judge how
//         much of a mismatch is the reference's own fp32 noise.
//
// The third-party arithmetic the reference takes from Eigen 3 (un-vendored) is
// restated from its definition: ABc.exp() (SolverMPC.cpp:93) is evaluated by the
// Taylor series, which terminates exactly because dt*[[A,B],[0,0]] is nilpotent
// of index 4 (rows 12..24 are zero and A^3 = 0); I_world.inverse()
// (SolverMPC.cpp:247) is the 3x3 adjugate formula.
#include <cmath>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <dlfcn.h>
#include <vector>

#include "../include/convexMPC_interface.h"
#include "../include/mpc_batch.h"

extern "C" int qp_port_solve(int nv, int nc, const double* H, const double* g,
                             const double* A, const double* lb, const double* ub,
                             int max_iter, double* x, int* iters);

namespace {

typedef int (*qpoases_fn)(int, int, const double*, const double*, const double*,
                          const double*, const double*, int*, double*, int*, double*);
qpoases_fn g_qpoases = nullptr;
int g_qpoases_ok = 0;
void* g_qpoases_handle = nullptr;

template <class T>
struct Mat {  // tiny row-major dense matrix
  int rows, cols;
  std::vector<T> a;
  Mat(int r, int c) : rows(r), cols(c), a((size_t)r * c, T(0)) {}
  T& operator()(int r, int c) { return a[(size_t)r * cols + c]; }
  T operator()(int r, int c) const { return a[(size_t)r * cols + c]; }
};

template <class T>
Mat<T> matmul(const Mat<T>& x, const Mat<T>& y) {
  Mat<T> z(x.rows, y.cols);
  for (int i = 0; i < x.rows; i++)
    for (int j = 0; j < y.cols; j++) {
      T acc = 0;
      for (int k = 0; k < x.cols; k++) acc += x(i, k) * y(k, j);
      z(i, j) = acc;
    }
  return z;
}

struct Inputs {
  float p[3], v[3], q[4], w[3], r[12], yaw, x_drag, alpha, weights[12];
  float I_body[3], mass, dt, mu, f_max;
  int horizon;
  const float* traj;          // 12*h
  const unsigned char* gait;  // 4*h
};

struct Result {
  int nv, nc, nwsr, rc_init, rc_primal;
  double objective;
};

// near_zero / near_one: SolverMPC.cpp:64-72 (float comparison against 0.01)
inline bool near_zero(float a) { return (a < 0.01 && a > -.01); }
inline bool near_one(float a) { return near_zero(a - 1); }

// The body of solve_mpc.  `sol` receives q_soln[12h] (doubles, zeros for swing
// legs); H_out/g_out (optional, 12h x 12h row-major / 12h) receive the reduced
// problem in its top-left nv x nv corner for assembly parity tests.
template <class T>
void solve_one(const Inputs& in, int backend, double* sol, Result* res,
               double* H_out, double* g_out) {
  const int h = in.horizon;
  const int NX = 13 * h, NU = 12 * h, NC = 20 * h;

  // ---- RobotState::set (RobotState.cpp:9-43) --------------------------------
  T r_feet[3][4];
  for (int rs = 0; rs < 3; rs++)
    for (int c = 0; c < 4; c++) r_feet[rs][c] = in.r[rs * 4 + c];
  T yc = (T)std::cos((T)in.yaw), ys = (T)std::sin((T)in.yaw);
  T R_yaw[3][3] = {{yc, -ys, 0}, {ys, yc, 0}, {0, 0, 1}};
  T I_body[3] = {(T)in.I_body[0], (T)in.I_body[1], (T)in.I_body[2]};
  T m = in.mass;

  // ---- quat_to_rpy (SolverMPC.cpp:257-267); q = (w,x,y,z) -------------------
  T qw = in.q[0], qx = in.q[1], qy = in.q[2], qz = in.q[3];
  // `as` is formed in double (-2. and .99999 are double literals) then narrowed
  double as_d = -2. * (double)(qx * qz - qw * qy);
  if (!(as_d < .99999)) as_d = .99999;
  T as = (T)as_d;
  T rpy[3];
  rpy[0] = std::atan2((T)2 * (qx * qy + qw * qz), qw * qw + qx * qx - qy * qy - qz * qz);
  rpy[1] = std::asin(as);
  rpy[2] = std::atan2((T)2 * (qy * qz + qw * qx), qw * qw - qx * qx - qy * qy + qz * qz);

  // ---- x_0 (SolverMPC.cpp:318) ---------------------------------------------
  T x0[13] = {rpy[2], rpy[1], rpy[0], (T)in.p[0], (T)in.p[1], (T)in.p[2],
              (T)in.w[0], (T)in.w[1], (T)in.w[2], (T)in.v[0], (T)in.v[1], (T)in.v[2],
              (T)-9.8f};

  // ---- I_world = R_yaw I_body R_yaw^T (SolverMPC.cpp:319) and its inverse ---
  T Iw[3][3];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      T acc = 0;
      for (int k = 0; k < 3; k++) acc += (R_yaw[i][k] * I_body[k]) * R_yaw[j][k];
      Iw[i][j] = acc;
    }
  T det = Iw[0][0] * (Iw[1][1] * Iw[2][2] - Iw[1][2] * Iw[2][1]) -
          Iw[0][1] * (Iw[1][0] * Iw[2][2] - Iw[1][2] * Iw[2][0]) +
          Iw[0][2] * (Iw[1][0] * Iw[2][1] - Iw[1][1] * Iw[2][0]);
  T Iinv[3][3];
  Iinv[0][0] = (Iw[1][1] * Iw[2][2] - Iw[1][2] * Iw[2][1]) / det;
  Iinv[0][1] = (Iw[0][2] * Iw[2][1] - Iw[0][1] * Iw[2][2]) / det;
  Iinv[0][2] = (Iw[0][1] * Iw[1][2] - Iw[0][2] * Iw[1][1]) / det;
  Iinv[1][0] = (Iw[1][2] * Iw[2][0] - Iw[1][0] * Iw[2][2]) / det;
  Iinv[1][1] = (Iw[0][0] * Iw[2][2] - Iw[0][2] * Iw[2][0]) / det;
  Iinv[1][2] = (Iw[0][2] * Iw[1][0] - Iw[0][0] * Iw[1][2]) / det;
  Iinv[2][0] = (Iw[1][0] * Iw[2][1] - Iw[1][1] * Iw[2][0]) / det;
  Iinv[2][1] = (Iw[0][1] * Iw[2][0] - Iw[0][0] * Iw[2][1]) / det;
  Iinv[2][2] = (Iw[0][0] * Iw[1][1] - Iw[0][1] * Iw[1][0]) / det;

  // ---- ct_ss_mats (SolverMPC.cpp:235-254), cross_mat (:226-233) --------------
  Mat<T> Ac(13, 13), Bc(13, 12);
  Ac(3, 9) = 1;
  Ac(11, 9) = in.x_drag;
  Ac(4, 10) = 1;
  Ac(5, 11) = 1;
  Ac(11, 12) = 1;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) Ac(i, 6 + j) = R_yaw[j][i];
  for (int b = 0; b < 4; b++) {
    T rx = r_feet[0][b], ry = r_feet[1][b], rz = r_feet[2][b];
    T cm[3][3] = {{0, -rz, ry}, {rz, 0, -rx}, {-ry, rx, 0}};
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) {
        T acc = 0;
        for (int k = 0; k < 3; k++) acc += Iinv[i][k] * cm[k][j];
        Bc(6 + i, b * 3 + j) = acc;
      }
    for (int i = 0; i < 3; i++) Bc(9 + i, b * 3 + i) = (T)1 / m;
  }

  // ---- c2qp (SolverMPC.cpp:87-125) -------------------------------------------
  // expmm = exp(dt*[[A,B],[0,0]]) by its (finite) Taylor series.
  Mat<T> M(25, 25);
  T dt = in.dt;
  for (int i = 0; i < 13; i++) {
    for (int j = 0; j < 13; j++) M(i, j) = dt * Ac(i, j);
    for (int j = 0; j < 12; j++) M(i, 13 + j) = dt * Bc(i, j);
  }
  Mat<T> E(25, 25), term(25, 25);
  for (int i = 0; i < 25; i++) E(i, i) = term(i, i) = 1;
  for (int k = 1; k <= 6; k++) {  // terms beyond k=3 are exactly zero
    term = matmul(term, M);
    for (auto& x : term.a) x /= (T)k;
    for (size_t i = 0; i < E.a.size(); i++) E.a[i] += term.a[i];
  }
  Mat<T> Adt(13, 13), Bdt(13, 12);
  for (int i = 0; i < 13; i++) {
    for (int j = 0; j < 13; j++) Adt(i, j) = E(i, j);
    for (int j = 0; j < 12; j++) Bdt(i, j) = E(i, 13 + j);
  }
  std::vector<Mat<T>> power;  // powerMats[k] = Adt^k (SolverMPC.cpp:103-107)
  power.emplace_back(13, 13);
  for (int i = 0; i < 13; i++) power[0](i, i) = 1;
  for (int i = 1; i < h + 1; i++) power.push_back(matmul(Adt, power[i - 1]));

  Mat<T> A_qp(NX, 13), B_qp(NX, NU);
  // powerMats[r-c]*Bdt depends on r-c only: form each product once and copy it into every block
  // of its diagonal (bit-identical to recomputing it per block as the reference does).
  std::vector<Mat<T>> phi;
  for (int d = 0; d < h; d++) phi.push_back(matmul(power[d], Bdt));
  for (int r = 0; r < h; r++) {
    for (int i = 0; i < 13; i++)
      for (int j = 0; j < 13; j++) A_qp(13 * r + i, j) = power[r + 1](i, j);
    for (int c = 0; c <= r; c++) {
      const Mat<T>& blk = phi[r - c];
      for (int i = 0; i < 13; i++)
        for (int j = 0; j < 12; j++) B_qp(13 * r + i, 12 * c + j) = blk(i, j);
    }
  }

  // ---- weights, trajectory, bounds, friction block (SolverMPC.cpp:335-378) ---
  std::vector<T> Sdiag(NX, 0), X_d(NX, 0), U_b(NC, 0);
  for (int i = 0; i < h; i++)
    for (int j = 0; j < 12; j++) {
      Sdiag[13 * i + j] = in.weights[j];
      X_d[13 * i + j] = in.traj[12 * i + j];
    }
  for (int k = 0; k < 4 * h; k++) {
    for (int c = 0; c < 4; c++) U_b[5 * k + c] = (T)5e10;
    U_b[5 * k + 4] = (T)in.gait[k] * (T)in.f_max;
  }
  T mu_inv = (T)1 / (T)in.mu;
  const T f_block[5][3] = {{mu_inv, 0, 1}, {-mu_inv, 0, 1}, {0, mu_inv, 1}, {0, -mu_inv, 1}, {0, 0, 1}};

  // ---- condensed QP (SolverMPC.cpp:395,399) ---------------------------------
  // qH = 2*(B'SB + alpha*I); qg = 2*B'S(A_qp x0 - X_d).  S is diagonal.
  std::vector<T> e(NX);
  for (int i = 0; i < NX; i++) {
    T acc = 0;
    for (int k = 0; k < 13; k++) acc += A_qp(i, k) * x0[k];
    e[i] = acc - X_d[i];
  }
  Mat<T> qH(NU, NU);
  std::vector<T> qg(NU);
  {
    // qH = 2*(B'(S B) + alpha I) as a rank-1-update sweep over the rows of B_qp (unit-stride inner
    // loop, vectorisable; Eigen's own summation order is unknown, any order restates it).  S is
    // diagonal; the reference stores it dense and pays a 13h x 13h x 12h product for S*B_qp on top.
    Mat<T> SB(NX, NU);
    for (int k = 0; k < NX; k++)
      for (int j = 0; j < NU; j++) SB(k, j) = Sdiag[k] * B_qp(k, j);
    for (int k = 0; k < NX; k++) {
      const T* sb = &SB.a[(size_t)k * NU];
      const int jmax = 12 * (k / 13 + 1);  // B_qp row k is zero beyond its own block column
      for (int i = 0; i < jmax; i++) {
        const T b = B_qp(k, i);
        if (b == (T)0) continue;
        T* row = &qH.a[(size_t)i * NU];
        for (int j = 0; j < jmax; j++) row[j] += b * sb[j];
      }
    }
    for (int i = 0; i < NU; i++) {
      for (int j = 0; j < NU; j++) qH(i, j) = (T)2 * (qH(i, j) + (i == j ? (T)in.alpha : (T)0));
      T acc = 0;
      for (int k = 0; k < NX; k++) acc += ((T)2 * B_qp(k, i)) * Sdiag[k] * e[k];
      qg[i] = acc;
    }
  }

  // ---- float -> double and swing-leg elimination (SolverMPC.cpp:423-525) -----
  std::vector<double> ub(NC), lb(NC, 0.0);
  for (int i = 0; i < NC; i++) ub[i] = (double)U_b[i];
  std::vector<char> var_elim(NU, 0), con_elim(NC, 0);
  int new_vars = NU, new_cons = NC;
  for (int i = 0; i < NC; i++) {
    if (!(near_zero((float)lb[i]) && near_zero((float)ub[i]))) continue;
    // constraint row i of fmat: block k = i/5, local row i%5, columns 3k..3k+2
    int k = i / 5, lr = i % 5;
    for (int c = 0; c < 3; c++) {
      int j = 3 * k + c;
      if (near_one((float)(double)f_block[lr][c])) {
        new_vars -= 3;
        new_cons -= 5;
        int cs = (j * 5) / 3 - 3;
        var_elim[j - 2] = var_elim[j - 1] = var_elim[j] = 1;
        for (int t = 0; t < 5; t++) con_elim[cs + t] = 1;
      }
    }
  }
  std::vector<int> var_ind, con_ind;
  for (int i = 0; i < NU; i++)
    if (!var_elim[i]) var_ind.push_back(i);
  for (int i = 0; i < NC; i++)
    if (!con_elim[i]) con_ind.push_back(i);
  const int nv = (int)var_ind.size(), nc = (int)con_ind.size();
  (void)new_vars;
  (void)new_cons;
  std::vector<double> H_red((size_t)nv * nv), g_red(nv), A_red((size_t)nc * nv, 0.0), lb_red(nc), ub_red(nc);
  for (int i = 0; i < nv; i++) {
    g_red[i] = (double)qg[var_ind[i]];
    for (int j = 0; j < nv; j++) H_red[(size_t)i * nv + j] = (double)qH(var_ind[i], var_ind[j]);
  }
  for (int con = 0; con < nc; con++) {
    int i = con_ind[con], k = i / 5, lr = i % 5;
    for (int st = 0; st < nv; st++) {
      int j = var_ind[st];
      // fmat(i,j) is f_block[lr][j-3k] inside the block, 0 elsewhere; the reference
      // narrows it through a float temporary (SolverMPC.cpp:516)
      float cval = (j / 3 == k) ? (float)f_block[lr][j - 3 * k] : 0.0f;
      A_red[(size_t)con * nv + st] = cval;
    }
    lb_red[con] = lb[i];
    ub_red[con] = ub[i];
  }

  if (H_out) {
    for (int i = 0; i < nv; i++)
      for (int j = 0; j < nv; j++) H_out[(size_t)i * NU + j] = H_red[(size_t)i * nv + j];
  }
  if (g_out)
    for (int i = 0; i < nv; i++) g_out[i] = g_red[i];

  // ---- QP solve (SolverMPC.cpp:527-557) --------------------------------------
  std::vector<double> q_red(nv > 0 ? nv : 1, 0.0);
  res->nv = nv;
  res->nc = nc;
  res->nwsr = 0;
  res->rc_init = 0;
  res->rc_primal = 0;
  res->objective = 0;
  if (nv > 0 && backend >= 0) {
    if (backend == 0) {
      if (!g_qpoases) {
        fprintf(stderr, "[oracle] reference qpOASES (oracle/_ref/libqpoases_ref.so) not loaded\n");
        abort();
      }
      int nwsr = 100;  // SolverMPC.cpp:435
      res->rc_init = g_qpoases(nv, nc, H_red.data(), g_red.data(), A_red.data(), lb_red.data(),
                               ub_red.data(), &nwsr, q_red.data(), &res->rc_primal, &res->objective);
      res->nwsr = nwsr;
    } else {
      int iters = 0;
      res->rc_init = qp_port_solve(nv, nc, H_red.data(), g_red.data(), A_red.data(), lb_red.data(),
                                   ub_red.data(), 10 * (nv + nc), q_red.data(), &iters);
      res->nwsr = iters;
    }
  }
  int vc = 0;
  for (int i = 0; i < NU; i++) {
    if (var_elim[i]) sol[i] = 0.0;
    else sol[i] = q_red[vc++];
  }
}

Inputs unpack(const void* record, int h) {
  const float* f = (const float*)record;
  Inputs in;
  memcpy(in.p, f + MPC_REC_P, 12);
  memcpy(in.v, f + MPC_REC_V, 12);
  memcpy(in.q, f + MPC_REC_Q, 16);
  memcpy(in.w, f + MPC_REC_W, 12);
  memcpy(in.r, f + MPC_REC_R, 48);
  in.yaw = f[MPC_REC_YAW];
  in.x_drag = f[MPC_REC_XDRAG];
  in.alpha = f[MPC_REC_ALPHA];
  memcpy(in.weights, f + MPC_REC_WEIGHTS, 48);
  memcpy(in.I_body, f + MPC_REC_IBODY, 12);
  in.mass = f[MPC_REC_MASS];
  in.dt = f[MPC_REC_DT];
  in.mu = f[MPC_REC_MU];
  in.f_max = f[MPC_REC_FMAX];
  in.horizon = h;
  in.traj = f + MPC_REC_TRAJ;
  in.gait = (const unsigned char*)record + 4 * (MPC_REC_TRAJ + 12 * h);
  return in;
}

// state of the legacy-ABI mirror (convexMPC_interface.cpp:13-20)
problem_setup o_setup;
update_data_t o_update;
std::vector<double> o_soln;
int o_has_solved = 0, o_precision = 32, o_backend = 0;
float o_I_body[3] = {.07f, 0.26f, 0.242f};  // RobotState.cpp:38-40
float o_mass = 9.f;                          // RobotState.h:23
Result o_last;

}  // namespace

extern "C" {

// Loads the reference qpOASES build.  Returns 1 when available.
int oracle_load_qpoases(const char* path) {
  if (g_qpoases) return 1;
  g_qpoases_handle = dlopen(path, RTLD_NOW | RTLD_LOCAL);
  if (!g_qpoases_handle) return 0;
  g_qpoases = (qpoases_fn)dlsym(g_qpoases_handle, "qpoases_ref_solve");
  int (*okfn)(void) = (int (*)(void))dlsym(g_qpoases_handle, "qpoases_ref_successful_return");
  g_qpoases_ok = okfn ? okfn() : 0;
  return g_qpoases ? 1 : 0;
}
int oracle_have_qpoases(void) { return g_qpoases ? 1 : 0; }

size_t oracle_record_stride(int h) { return ((size_t)(4 * (MPC_REC_TRAJ + 12 * h) + 4 * h) + 15) / 16 * 16; }

// Solves `batch` packed records.  precision: 32|64.  backend: 0 = reference
// qpOASES, 1 = oracle's own dense active-set port, -1 = assemble only.
//   sol   [batch*12h] doubles (q_soln);   info [batch*5] ints: nv, nc, nWSR, init rc, primal rc
//   H_out [batch*(12h)^2], g_out [batch*12h] optional
void oracle_solve_batch(const void* records, int batch, int h, int precision, int backend,
                        double* sol, int* info, double* obj, double* H_out, double* g_out) {
  const size_t stride = oracle_record_stride(h);
  const int NU = 12 * h;
  for (int b = 0; b < batch; b++) {
    Inputs in = unpack((const char*)records + stride * b, h);
    Result res;
    double* Hb = H_out ? H_out + (size_t)b * NU * NU : nullptr;
    double* gb = g_out ? g_out + (size_t)b * NU : nullptr;
    if (precision == 64) solve_one<double>(in, backend, sol + (size_t)b * NU, &res, Hb, gb);
    else solve_one<float>(in, backend, sol + (size_t)b * NU, &res, Hb, gb);
    if (info) {
      int* o = info + 5 * b;
      o[0] = res.nv; o[1] = res.nc; o[2] = res.nwsr; o[3] = res.rc_init; o[4] = res.rc_primal;
    }
    if (obj) obj[b] = res.objective;
  }
}

// ---- mirror of the reference C interface (convexMPC_interface.cpp:42-180) ----
void oracle_configure(int precision, int backend) { o_precision = precision; o_backend = backend; }
void oracle_set_robot(const float* I_body, float mass) {
  memcpy(o_I_body, I_body, 12);
  o_mass = mass;
}
void oracle_setup_problem(double dt, int horizon, double mu, double f_max) {
  o_setup.horizon = horizon;
  o_setup.f_max = f_max;
  o_setup.mu = mu;
  o_setup.dt = dt;
  o_soln.assign(12 * horizon, 0.0);
}
void oracle_update_x_drag(float x_drag) { o_update.x_drag = x_drag; }
void oracle_update_problem_data_floats(float* p, float* v, float* q, float* w, float* r, float yaw,
                                       float* weights, float* state_trajectory, float alpha, int* gait) {
  const int h = o_setup.horizon;
  o_update.alpha = alpha;
  o_update.yaw = yaw;
  unsigned char* gait_bytes = reinterpret_cast<unsigned char*>(&o_update) + offsetof(update_data_t, gait);
  for (int i = 0; i < 4 * h; i++) gait_bytes[i] = (unsigned char)gait[i];  // spills into hack_pad like upstream
  memcpy(o_update.p, p, 12);
  memcpy(o_update.v, v, 12);
  memcpy(o_update.q, q, 16);
  memcpy(o_update.w, w, 12);
  memcpy(o_update.r, r, 48);
  memcpy(o_update.weights, weights, 48);
  memcpy(o_update.traj, state_trajectory, sizeof(float) * 12 * h);
  Inputs in;
  memcpy(in.p, o_update.p, 12);
  memcpy(in.v, o_update.v, 12);
  memcpy(in.q, o_update.q, 16);
  memcpy(in.w, o_update.w, 12);
  memcpy(in.r, o_update.r, 48);
  in.yaw = o_update.yaw;
  in.x_drag = o_update.x_drag;
  in.alpha = o_update.alpha;
  memcpy(in.weights, o_update.weights, 48);
  memcpy(in.I_body, o_I_body, 12);
  in.mass = o_mass;
  in.dt = o_setup.dt;
  in.mu = o_setup.mu;
  in.f_max = o_setup.f_max;
  in.horizon = h;
  in.traj = o_update.traj;
  in.gait = gait_bytes;
  if (o_precision == 64) solve_one<double>(in, o_backend, o_soln.data(), &o_last, nullptr, nullptr);
  else solve_one<float>(in, o_backend, o_soln.data(), &o_last, nullptr, nullptr);
  o_has_solved = 1;
}
void oracle_update_problem_data(double* p, double* v, double* q, double* w, double* r, double yaw,
                                double* weights, double* state_trajectory, double alpha, int* gait) {
  const int h = o_setup.horizon;
  float pf[3], vf[3], qf[4], wf[3], rf[12], wt[12];
  std::vector<float> tr(12 * h);
  for (int i = 0; i < 3; i++) { pf[i] = p[i]; vf[i] = v[i]; wf[i] = w[i]; }
  for (int i = 0; i < 4; i++) qf[i] = q[i];
  for (int i = 0; i < 12; i++) { rf[i] = r[i]; wt[i] = weights[i]; }
  for (int i = 0; i < 12 * h; i++) tr[i] = state_trajectory[i];
  oracle_update_problem_data_floats(pf, vf, qf, wf, rf, (float)yaw, wt, tr.data(), (float)alpha, gait);
}
double oracle_get_solution(int index) {
  if (!o_has_solved) return 0.f;
  return o_soln[index];
}
int oracle_last_nwsr(void) { return o_last.nwsr; }
int oracle_last_rc(void) { return o_last.rc_init; }

}  // extern "C"
