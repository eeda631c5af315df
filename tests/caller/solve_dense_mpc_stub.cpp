// TEST-ONLY link check of the drop-in boundary.
//
// A C++ caller that issues exactly the call sequence of ConvexMPCLocomotion::solveDenseMPC
// (/root/reference/src/MPC_Ctrl/ConvexMPCLocomotion.cpp:592-687: setup_problem :630, update_x_drag :632 -- C++
// linkage --, update_solver_settings :651, update_problem_data_floats :664, get_solution(leg*3+axis) :672-674),
// compiled against include/convexMPC_interface.h only and linked against libquadruped_mpc_b200.so the way the
// reference's GaitCtrller library would be.  Inputs: the config-1 nominal state of
// SURVEY 8d (trot, horizon 10) for the two horizons the reference's mode-1 gaits switch between.
// Output: one line per solve, "h <horizon> status <code> f <12 numbers>", parsed by tests/.
#include <cmath>
#include <cstdio>

#include "convexMPC_interface.h"

extern "C" int mpc_last_status(void);  // additive getter of this build (not in the reference)

static void solve_once(int horizonLength, float vx) {
  const float dtMPC = 0.002f * 13;
  // :598-604
  float Q[12] = {2.5f, 2.5f, 10, 50, 50, 100, 0, 0, 0.5f, 0.2f, 0.2f, 0.1f};
  float alpha = 4e-5f;
  float p[3] = {0, 0, 0.29f}, v[3] = {vx, 0, 0}, w[3] = {0, 0, 0}, q[4] = {1, 0, 0, 0};
  float yaw = 0.f;
  // :611-613  r[i] = pFoot[i%4][i/4] - position[i/4]   (row-major 3x4: r[axis*4+leg])
  const float foot[4][3] = {{0.19f, -0.111f, 0}, {0.19f, 0.111f, 0}, {-0.19f, -0.111f, 0}, {-0.19f, 0.111f, 0}};
  float r[12];
  for (int i = 0; i < 12; i++) r[i] = foot[i % 4][i / 4] - p[i / 4];
  // :514-577 reference trajectory (constant velocity roll-out), :386 contact table of a trot
  float trajAll[12 * 36];
  int mpcTable[4 * 36];
  for (int i = 0; i < horizonLength; i++) {
    float row[12] = {0, 0, 0, p[0] + dtMPC * i * vx, 0, 0.25f, 0, 0, 0, vx, 0, 0};
    for (int j = 0; j < 12; j++) trajAll[12 * i + j] = row[j];
    const int a = ((i % horizonLength) < horizonLength / 2) ? 1 : 0;
    mpcTable[4 * i + 0] = a; mpcTable[4 * i + 1] = 1 - a; mpcTable[4 * i + 2] = 1 - a; mpcTable[4 * i + 3] = a;
  }
  setup_problem(dtMPC, horizonLength, 0.4, 120);                 // :630
  update_x_drag(0.f);                                            // :632
  update_solver_settings(10000, 1e-7, 1e-8, 1.5, 0.1, 0.0);      // :644-651
  update_problem_data_floats(p, v, q, w, r, yaw, Q, trajAll, alpha, mpcTable);  // :664
  std::printf("h %d status %d f", horizonLength, mpc_last_status());
  for (int leg = 0; leg < 4; leg++)
    for (int axis = 0; axis < 3; axis++) std::printf(" %.9g", get_solution(leg * 3 + axis));  // :672-674
  std::printf("\n");
}

int main() {
  solve_once(10, 0.5f);
  solve_once(14, 0.5f);   // horizon switch between ticks (mode-1 gaits, :173-233)
  solve_once(10, 0.3f);
  return 0;
}
