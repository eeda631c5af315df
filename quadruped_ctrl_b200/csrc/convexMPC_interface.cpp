// Legacy single-robot C interface of the reference, served by the batched sm_100a engine.
//
// Mirrors /root/reference/src/MPC_Ctrl/convexMPC_interface.cpp:13-180 symbol for symbol
// (same names, argument meaning, global-singleton semantics, void returns) so that
// ConvexMPCLocomotion::solveDenseMPC (ConvexMPCLocomotion.cpp:630-674) links against this
// library unchanged.  Every solve is a batch of one on CUDA device 0; there is no CPU solver
// here: without a usable sm_100 device the call reports the failure on stderr, leaves
// get_solution() at 0 and mpc_last_status() negative.
#include "../../include/convexMPC_interface.h"

#include <cstddef>
#include <cstdio>
#include <cstring>
#include <map>
#include <vector>

#include "../../include/mpc_batch.h"

namespace {

// reference: convexMPC_interface.cpp:13-20 (process-global singletons, non re-entrant)
problem_setup problem_configuration;
update_data_t update;
int has_solved = 0;
int first_run = 1;

std::map<int, mpc_batch_t*> engines;  // one engine per horizon seen (mode-1 gaits switch h between ticks)
std::vector<char> record;
std::vector<double> q_soln;
int last_status = -1, last_iters = 0;
float robot_I_body[3] = {.07f, 0.26f, 0.242f};  // RobotState.cpp:38-40
float robot_mass = 9.f;                           // RobotState.h:23

mpc_batch_t* engine_for(int h) {
  auto it = engines.find(h);
  if (it != engines.end()) return it->second;
  mpc_batch_t* e = nullptr;
  int rc = mpc_batch_create(&e, 0, h, 1);
  if (rc != MPC_OK) {
    fprintf(stderr, "[quadruped_mpc_b200] cannot create the GPU MPC engine (rc=%d): %s\n", rc, mpc_last_error());
    last_status = rc;
    return nullptr;
  }
  engines[h] = e;
  return e;
}

// reference: mint_to_u8, convexMPC_interface.cpp:75-79
inline unsigned char mint_to_u8(int i) { return (unsigned char)i; }

// the contents of problem_configuration / update that solve_mpc reads, as one batch record (mpc_batch.h)
void pack_record(char* out, int h) {
  float* f = (float*)out;
  memcpy(f + MPC_REC_P, update.p, 12);
  memcpy(f + MPC_REC_V, update.v, 12);
  memcpy(f + MPC_REC_Q, update.q, 16);
  memcpy(f + MPC_REC_W, update.w, 12);
  memcpy(f + MPC_REC_R, update.r, 48);
  f[MPC_REC_YAW] = update.yaw;
  f[MPC_REC_XDRAG] = update.x_drag;
  f[MPC_REC_ALPHA] = update.alpha;
  memcpy(f + MPC_REC_WEIGHTS, update.weights, 48);
  memcpy(f + MPC_REC_IBODY, robot_I_body, 12);
  f[MPC_REC_MASS] = robot_mass;
  f[MPC_REC_DT] = problem_configuration.dt;
  f[MPC_REC_MU] = problem_configuration.mu;
  f[MPC_REC_FMAX] = problem_configuration.f_max;
  memcpy(f + MPC_REC_TRAJ, update.traj, sizeof(float) * 12 * h);
  memcpy(out + mpc_record_gait_offset(h), reinterpret_cast<const unsigned char*>(&update) + offsetof(update_data_t, gait),
         4 * h);  // gait[] runs on into hack_pad[] as upstream
}

void solve_now() {
  const int h = problem_configuration.horizon;
  has_solved = 0;
  if (h < 1 || h > MPC_MAX_HORIZON) {
    fprintf(stderr, "[quadruped_mpc_b200] horizon %d outside 1..%d\n", h, MPC_MAX_HORIZON);
    last_status = MPC_E_ARG;
    return;
  }
  mpc_batch_t* eng = engine_for(h);
  if (!eng) return;
  const size_t stride = mpc_record_stride(h);
  record.assign(stride, 0);
  pack_record(record.data(), h);
  q_soln.assign(12 * h, 0.0);
  float forces[12];
  int32_t status = 0;
  int rc = mpc_batch_solve_host(eng, record.data(), 1, forces, q_soln.data(), &status);
  if (rc != MPC_OK) {
    fprintf(stderr, "[quadruped_mpc_b200] failed to solve! (rc=%d: %s)\n", rc, mpc_batch_last_error(eng));
    last_status = rc;
    return;
  }
  last_status = MPC_STATUS_CODE(status);
  last_iters = (int)MPC_STATUS_ITERS(status);
  if (last_status != MPC_STATUS_OPTIMAL && last_status != MPC_STATUS_NO_STANCE)
    printf("failed to solve!\n");  // the reference's only failure report (SolverMPC.cpp:540-541)
  has_solved = 1;
}

}  // namespace

// reference: convexMPC_interface.cpp:42-66
void setup_problem(double dt, int horizon, double mu, double f_max) {
  if (first_run) first_run = 0;
  problem_configuration.horizon = horizon;
  problem_configuration.f_max = f_max;
  problem_configuration.mu = mu;
  problem_configuration.dt = dt;
}

// reference: convexMPC_interface.cpp:107-119
void update_solver_settings(int max_iter, double rho, double sigma, double solver_alpha, double terminate,
                            double use_jcqp) {
  update.max_iterations = max_iter;
  update.rho = rho;
  update.sigma = sigma;
  update.solver_alpha = solver_alpha;
  update.terminate = terminate;
  if (use_jcqp > 1.5) update.use_jcqp = 2;
  else if (use_jcqp > 0.5) update.use_jcqp = 1;
  else update.use_jcqp = 0;
  // Upstream, use_jcqp = 1 / 2 hands the same QP (full / reduced) to the JCQP ADMM solver instead of qpOASES
  // (SolverMPC.cpp:406-420, 558-619) and stops at the residual tolerance `terminate`; the shipped caller hard-wires
  // 0.0 (ConvexMPCLocomotion.cpp:649).  This engine has one QP kernel: it returns the KKT-exact optimum that ADMM
  // iterates towards, so rho / sigma / solver_alpha / terminate / max_iter have nothing to act on.  Said once, loudly,
  // rather than silently ignored; mpc_jcqp_requested() lets a caller check.
  static int told = 0;
  if (update.use_jcqp != 0 && !told) {
    told = 1;
    fprintf(stderr, "[quadruped_mpc_b200] update_solver_settings: use_jcqp=%d requested -- there is no ADMM path in this "
                    "library; the QP is solved to its exact optimum by the active-set kernel and rho/sigma/alpha/"
                    "terminate/max_iter are ignored\n", update.use_jcqp);
  }
}

// Additive: the use_jcqp mode last requested through update_solver_settings (0, 1 or 2); see the note there.
extern "C" int mpc_jcqp_requested(void) { return update.use_jcqp; }

// reference: convexMPC_interface.cpp:121-169
void update_problem_data_floats(float* p, float* v, float* q, float* w, float* r, float yaw, float* weights,
                                float* state_trajectory, float alpha, int* gait) {
  const int h = problem_configuration.horizon;
  if (h < 1 || h > MPC_MAX_HORIZON) {
    fprintf(stderr, "[quadruped_mpc_b200] update_problem_data_floats before a valid setup_problem (horizon %d)\n", h);
    last_status = MPC_E_ARG;
    has_solved = 0;
    return;
  }
  update.alpha = alpha;
  update.yaw = yaw;
  // 4h bytes starting at update.gait: for h > 9 this runs on into hack_pad, exactly where upstream's
  // out-of-bounds loop puts them.  Addressed from the struct base so the compiler cannot assume i < 36.
  unsigned char* gait_bytes = reinterpret_cast<unsigned char*>(&update) + offsetof(update_data_t, gait);
  for (int i = 0; i < 4 * h; i++) gait_bytes[i] = mint_to_u8(gait[i]);
  memcpy((void*)update.p, (void*)p, sizeof(float) * 3);
  memcpy((void*)update.v, (void*)v, sizeof(float) * 3);
  memcpy((void*)update.q, (void*)q, sizeof(float) * 4);
  memcpy((void*)update.w, (void*)w, sizeof(float) * 3);
  memcpy((void*)update.r, (void*)r, sizeof(float) * 12);
  memcpy((void*)update.weights, (void*)weights, sizeof(float) * 12);
  memcpy((void*)update.traj, (void*)state_trajectory, sizeof(float) * 12 * h);
  solve_now();
}

// reference: convexMPC_interface.cpp:88-105 (doubles narrowed to float, then the same path)
void update_problem_data(double* p, double* v, double* q, double* w, double* r, double yaw, double* weights,
                         double* state_trajectory, double alpha, int* gait) {
  const int h = problem_configuration.horizon;
  if (h < 1 || h > MPC_MAX_HORIZON) {
    fprintf(stderr, "[quadruped_mpc_b200] update_problem_data before a valid setup_problem (horizon %d)\n", h);
    last_status = MPC_E_ARG;
    has_solved = 0;
    return;
  }
  float pf[3], vf[3], qf[4], wf[3], rf[12], wt[12];
  std::vector<float> tr(12 * h);
  for (int i = 0; i < 3; i++) { pf[i] = (float)p[i]; vf[i] = (float)v[i]; wf[i] = (float)w[i]; }
  for (int i = 0; i < 4; i++) qf[i] = (float)q[i];
  for (int i = 0; i < 12; i++) { rf[i] = (float)r[i]; wt[i] = (float)weights[i]; }
  for (int i = 0; i < 12 * h; i++) tr[i] = (float)state_trajectory[i];
  update_problem_data_floats(pf, vf, qf, wf, rf, (float)yaw, wt, tr.data(), (float)alpha, gait);
}

// reference: convexMPC_interface.cpp:171-173 (C++ linkage, as upstream)
void update_x_drag(float x_drag) { update.x_drag = x_drag; }

// reference: convexMPC_interface.cpp:175-180
double get_solution(int index) {
  if (!has_solved) return 0.f;
  if (index < 0 || index >= (int)q_soln.size()) return 0.f;  // upstream reads out of bounds here
  return q_soln[index];
}

int mpc_last_status(void) { return last_status; }
int mpc_last_iterations(void) { return last_iters; }

// Additive: the reference hard-codes the Mini-Cheetah inertia and mass (RobotState.cpp:38-40, RobotState.h:23).
extern "C" void mpc_set_robot(const float* I_body_diag, float mass) {
  if (I_body_diag) memcpy(robot_I_body, I_body_diag, 12);
  robot_mass = mass;
}

// Additive: writes the batch record (include/mpc_batch.h) of the inputs last handed to setup_problem /
// update_x_drag / update_problem_data*, i.e. the bridge from the legacy calls to the batched API.
// `out` must hold mpc_record_stride(horizon) bytes.  Returns the horizon, or MPC_E_ARG.
extern "C" int mpc_legacy_record(void* out) {
  const int h = problem_configuration.horizon;
  if (!out || h < 1 || h > MPC_MAX_HORIZON) return MPC_E_ARG;
  memset(out, 0, mpc_record_stride(h));
  pack_record((char*)out, h);
  return h;
}

// Additive: releases the cached engines (e.g. before the process unloads the library).
extern "C" void mpc_shutdown(void) {
  for (auto& kv : engines) mpc_batch_destroy(kv.second);
  engines.clear();
  has_solved = 0;
}
