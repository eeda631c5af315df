// TEST INFRASTRUCTURE -- not part of the product path.
//
// Independent CPU solver for the reduced MPC QP
//     min 1/2 x'Hx + g'x   s.t.  lb <= A x <= ub        (H symmetric positive definite)
// used (a) to cross-check the reference's qpOASES (oracle/_ref) on every golden
// problem and (b) as the "port" CPU oracle when oracle/_ref is not available.
//
// The reference solves this QP with qpOASES' online active-set homotopy
//   (/root/reference/src/qpOASES/src/QProblem.cpp:316-368 init -> :1301 solveInitialQP
//    -> :1555 solveQP loop), cold-started, nWSR = 100 (SolverMPC.cpp:435,537).
// H is strictly positive definite (alpha > 0, SolverMPC.cpp:395), so the optimum
// is unique and any exact active-set method must return the same point; this file
// uses the Goldfarb-Idnani dual active-set method (Math. Programming 27, 1983) in
// dense fp64 with an explicit inverse Hessian and a re-factorised Schur complement
// per step -- written for clarity, not speed.
#include <cmath>
#include <cstdio>
#include <vector>

namespace {

// in-place lower Cholesky of an n x n row-major matrix; returns false if not PD
bool cholesky(std::vector<double>& a, int n) {
  for (int j = 0; j < n; j++) {
    double d = a[(size_t)j * n + j];
    for (int k = 0; k < j; k++) d -= a[(size_t)j * n + k] * a[(size_t)j * n + k];
    if (!(d > 0)) return false;
    d = std::sqrt(d);
    a[(size_t)j * n + j] = d;
    for (int i = j + 1; i < n; i++) {
      double s = a[(size_t)i * n + j];
      for (int k = 0; k < j; k++) s -= a[(size_t)i * n + k] * a[(size_t)j * n + k];
      a[(size_t)i * n + j] = s / d;
    }
  }
  return true;
}

void chol_solve(const std::vector<double>& L, int n, std::vector<double>& b) {
  for (int i = 0; i < n; i++) {
    double s = b[i];
    for (int k = 0; k < i; k++) s -= L[(size_t)i * n + k] * b[k];
    b[i] = s / L[(size_t)i * n + i];
  }
  for (int i = n - 1; i >= 0; i--) {
    double s = b[i];
    for (int k = i + 1; k < n; k++) s -= L[(size_t)k * n + i] * b[k];
    b[i] = s / L[(size_t)i * n + i];
  }
}

}  // namespace

// returns 0 = optimal, 1 = iteration cap, 2 = H not PD, 3 = infeasible
extern "C" int qp_port_solve(int nv, int nc, const double* H, const double* g,
                             const double* A, const double* lb, const double* ub,
                             int max_iter, double* x, int* iters) {
  const double INF = 1e30, BIG = 1e9, VTOL = 1e-10;
  std::vector<double> L(H, H + (size_t)nv * nv);
  if (!cholesky(L, nv)) return 2;
  // M = H^-1, column by column
  std::vector<double> M((size_t)nv * nv), col(nv);
  for (int j = 0; j < nv; j++) {
    std::fill(col.begin(), col.end(), 0.0);
    col[j] = 1.0;
    chol_solve(L, nv, col);
    for (int i = 0; i < nv; i++) M[(size_t)i * nv + j] = col[i];
  }
  for (int i = 0; i < nv; i++) {
    double s = 0;
    for (int k = 0; k < nv; k++) s -= M[(size_t)i * nv + k] * g[k];
    x[i] = s;
  }
  // working set: constraint index and sign (+1: a'x >= lb, -1: -a'x >= -ub)
  std::vector<int> Wc, Ws;
  std::vector<double> u;
  std::vector<double> np(nv), v(nv), z(nv), w, r, S, Nr(nv);
  int it = 0;
  for (;; ) {
    // most violated one-sided constraint
    int p = -1, psign = 0;
    double worst = -VTOL;
    for (int i = 0; i < nc; i++) {
      double ax = 0;
      for (int k = 0; k < nv; k++) ax += A[(size_t)i * nv + k] * x[k];
      if (lb[i] > -BIG && ax - lb[i] < worst) { worst = ax - lb[i]; p = i; psign = 1; }
      if (ub[i] < BIG && ub[i] - ax < worst) { worst = ub[i] - ax; p = i; psign = -1; }
    }
    if (p < 0) break;
    if (it >= max_iter) { if (iters) *iters = it; return 1; }
    it++;
    for (int k = 0; k < nv; k++) np[k] = psign * A[(size_t)p * nv + k];
    double bp = psign > 0 ? lb[p] : -ub[p];
    double up = 0;
    for (int guard = 0; guard < 4 * (nv + nc) + 16; guard++) {
      const int m = (int)Wc.size();
      for (int i = 0; i < nv; i++) {
        double s = 0;
        for (int k = 0; k < nv; k++) s += M[(size_t)i * nv + k] * np[k];
        v[i] = s;
      }
      double nv_np = 0;
      for (int k = 0; k < nv; k++) nv_np += np[k] * v[k];
      z = v;
      r.assign(m, 0.0);
      if (m > 0) {
        // MN columns, S = N'MN, w = N'v
        std::vector<double> MN((size_t)nv * m);
        for (int j = 0; j < m; j++)
          for (int i = 0; i < nv; i++) {
            double s = 0;
            for (int k = 0; k < nv; k++) s += M[(size_t)i * nv + k] * Ws[j] * A[(size_t)Wc[j] * nv + k];
            MN[(size_t)i * m + j] = s;
          }
        S.assign((size_t)m * m, 0.0);
        w.assign(m, 0.0);
        for (int a = 0; a < m; a++) {
          for (int b = 0; b < m; b++) {
            double s = 0;
            for (int k = 0; k < nv; k++) s += Ws[a] * A[(size_t)Wc[a] * nv + k] * MN[(size_t)k * m + b];
            S[(size_t)a * m + b] = s;
          }
          double s = 0;
          for (int k = 0; k < nv; k++) s += Ws[a] * A[(size_t)Wc[a] * nv + k] * v[k];
          w[a] = s;
        }
        if (!cholesky(S, m)) { if (iters) *iters = it; return 3; }
        r = w;
        chol_solve(S, m, r);
        for (int i = 0; i < nv; i++) {
          double s = 0;
          for (int j = 0; j < m; j++) s += MN[(size_t)i * m + j] * r[j];
          z[i] = v[i] - s;
        }
      }
      double znp = 0;
      for (int k = 0; k < nv; k++) znp += z[k] * np[k];
      double sp = -bp;
      for (int k = 0; k < nv; k++) sp += np[k] * x[k];
      double t1 = INF;
      int kdrop = -1;
      for (int j = 0; j < m; j++)
        if (r[j] > 0 && u[j] / r[j] < t1) { t1 = u[j] / r[j]; kdrop = j; }
      double t2 = (znp > 1e-11 * nv_np) ? -sp / znp : INF;
      double t = t1 < t2 ? t1 : t2;
      if (t >= INF) { if (iters) *iters = it; return 3; }
      for (int j = 0; j < m; j++) u[j] -= t * r[j];
      up += t;
      if (t2 < INF)
        for (int k = 0; k < nv; k++) x[k] += t * z[k];
      if (t2 <= t1) {  // full step: constraint p becomes active
        Wc.push_back(p); Ws.push_back(psign); u.push_back(up);
        break;
      }
      Wc.erase(Wc.begin() + kdrop); Ws.erase(Ws.begin() + kdrop); u.erase(u.begin() + kdrop);
    }
  }
  if (iters) *iters = it;
  return 0;
}
