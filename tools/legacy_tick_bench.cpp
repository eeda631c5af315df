// BASELINE config 1 ("1 robot, trot gait, horizon=10, plane terrain") through the reference's C interface:
// the call sequence of ConvexMPCLocomotion::solveDenseMPC (/root/reference/src/MPC_Ctrl/ConvexMPCLocomotion.cpp:
// 630-674: setup_problem, update_x_drag, update_solver_settings, update_problem_data_floats, 12 x get_solution),
// one MPC tick after the other from C++, timed per tick.  bench.py builds this file twice:
//   * against libquadruped_mpc_b200.so (include/convexMPC_interface.h): the GPU engine as a batch of one,
//   * with -DLEGACY_ORACLE against oracle/liboracle.so: the CPU reference path (fp32 assembly restatement +
//     the reference's qpOASES) behind the same five calls -- SURVEY 8d "CPU path timing (i)".
// Inputs: SURVEY 8d config 1 (nominal state, v = 0.5 m/s, trot offsets (0,5,5,0) / durations 5 of 10), the ten gait
// phases in turn.  Output: "tick_us <median> <p95> <mean>", then one line "phase <k> f <12 forces>" per phase.
//   usage: legacy_tick_bench <ticks> [<horizon>]
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <vector>

#ifdef LEGACY_ORACLE
extern "C" {
void oracle_setup_problem(double dt, int horizon, double mu, double f_max);
void oracle_update_x_drag(float x_drag);
void oracle_update_problem_data_floats(float* p, float* v, float* q, float* w, float* r, float yaw, float* weights,
                                       float* state_trajectory, float alpha, int* gait);
double oracle_get_solution(int index);
int oracle_load_qpoases(const char* path);
void oracle_configure(int precision, int backend);
}
#define setup_problem oracle_setup_problem
#define update_x_drag oracle_update_x_drag
#define update_problem_data_floats oracle_update_problem_data_floats
#define get_solution oracle_get_solution
static void update_solver_settings(int, double, double, double, double, double) {}  // JCQP settings: unused upstream
#else
#include "convexMPC_interface.h"
#endif

int main(int argc, char** argv) {
  const int ticks = argc > 1 ? atoi(argv[1]) : 1000;
  const int h = argc > 2 ? atoi(argv[2]) : 10;
#ifdef LEGACY_ORACLE
  const int have_ref = oracle_load_qpoases(argc > 3 ? argv[3] : "");
  oracle_configure(32, have_ref ? 0 : 1);  // fp32 assembly (the reference's arithmetic); reference qpOASES when built
#endif
  const float dtMPC = 0.002f * 13;
  float Q[12] = {2.5f, 2.5f, 10, 50, 50, 100, 0, 0, 0.5f, 0.2f, 0.2f, 0.1f};  // :598
  float alpha = 4e-5f;                                                          // :604
  float p[3] = {0, 0, 0.29f}, v[3] = {0.5f, 0, 0}, w[3] = {0, 0, 0}, q[4] = {1, 0, 0, 0};
  const float foot[4][3] = {{0.19f, -0.111f, 0}, {0.19f, 0.111f, 0}, {-0.19f, -0.111f, 0}, {-0.19f, 0.111f, 0}};
  float r[12];
  for (int i = 0; i < 12; i++) r[i] = foot[i % 4][i / 4] - p[i / 4];  // :611-613
  float trajAll[12 * 36];
  for (int i = 0; i < h; i++) {  // :547-576, constant-velocity roll-out with running float sums
    float* t = trajAll + 12 * i;
    for (int j = 0; j < 12; j++) t[j] = 0.f;
    t[3] = i == 0 ? p[0] : trajAll[12 * (i - 1) + 3] + dtMPC * v[0];
    t[5] = 0.25f;
    t[9] = v[0];
  }
  const int offsets[4] = {0, h / 2, h / 2, 0}, durations[4] = {h / 2, h / 2, h / 2, h / 2};
  std::vector<double> us(ticks);
  std::vector<double> forces(12 * h, 0.0);
  for (int k = -20; k < ticks; k++) {  // 20 untimed warm-up ticks
    const int phase = ((k % h) + h) % h;
    int mpcTable[4 * 36];
    for (int i = 0; i < h; i++) {  // Gait.cpp:142-166
      const int iter = (i + phase + 1) % h;
      for (int j = 0; j < 4; j++) {
        int progress = iter - offsets[j];
        if (progress < 0) progress += h;
        mpcTable[i * 4 + j] = progress < durations[j] ? 1 : 0;
      }
    }
    const auto t0 = std::chrono::steady_clock::now();
    setup_problem(dtMPC, h, 0.4, 120);                          // :630
    update_x_drag(0.f);                                         // :632
    update_solver_settings(10000, 1e-7, 1e-8, 1.5, 0.1, 0.0);   // :644-651
    update_problem_data_floats(p, v, q, w, r, 0.f, Q, trajAll, alpha, mpcTable);  // :664
    double f[12];
    for (int leg = 0; leg < 4; leg++)
      for (int axis = 0; axis < 3; axis++) f[leg * 3 + axis] = get_solution(leg * 3 + axis);  // :672-674
    const auto t1 = std::chrono::steady_clock::now();
    if (k >= 0) us[k] = std::chrono::duration<double, std::micro>(t1 - t0).count();
    for (int i = 0; i < 12; i++) forces[12 * phase + i] = f[i];
  }
  std::vector<double> s = us;
  std::sort(s.begin(), s.end());
  double mean = 0;
  for (double x : us) mean += x;
  std::printf("tick_us %.3f %.3f %.3f\n", s[ticks / 2], s[(int)(0.95 * (ticks - 1))], mean / ticks);
  for (int ph = 0; ph < h; ph++) {
    std::printf("phase %d f", ph);
    for (int i = 0; i < 12; i++) std::printf(" %.9g", forces[12 * ph + i]);
    std::printf("\n");
  }
  return 0;
}
