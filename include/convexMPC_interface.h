/*
 * Drop-in boundary of the convex-MPC hot path (legacy, single robot).
 *
 * Binary-compatible restatement of the reference interface
 *   /root/reference/src/MPC_Ctrl/convexMPC_interface.h:1-49
 * so that ConvexMPCLocomotion::solveDenseMPC
 *   (/root/reference/src/MPC_Ctrl/ConvexMPCLocomotion.cpp:630-674)
 * links against libquadruped_mpc_b200.so without source changes.
 *
 * Every call below is served by the batched sm_100a engine declared in
 * mpc_batch.h with a batch of one; there is no CPU solver behind it.
 *
 * Struct layout must stay byte-identical to the reference (field order and
 * sizes at convexMPC_interface.h:13-38): tests/test_abi.py checks the offsets.
 */
#ifndef _convexmpc_interface
#define _convexmpc_interface
#define K_MAX_GAIT_SEGMENTS 36

#ifdef __cplusplus
#define EXTERNC extern "C"
#else
#define EXTERNC
#endif

/* reference: convexMPC_interface.h:13-19 */
struct problem_setup
{
  float dt;
  float mu;
  float f_max;
  int horizon;
};

/* reference: convexMPC_interface.h:21-38.  gait[] is deliberately overrun into
 * hack_pad[] for horizons above 9 (4*h > 36 bytes), as upstream does. */
struct update_data_t
{
  float p[3];
  float v[3];
  float q[4];
  float w[3];
  float r[12];
  float yaw;
  float weights[12];
  float traj[12*K_MAX_GAIT_SEGMENTS];
  float alpha;
  unsigned char gait[K_MAX_GAIT_SEGMENTS];
  unsigned char hack_pad[1000];
  int max_iterations;
  double rho, sigma, solver_alpha, terminate;
  int use_jcqp;
  float x_drag;
};

/* reference: convexMPC_interface.cpp:42-66.  Stores dt/horizon/mu/f_max.  The
 * reference re-mallocs and zeroes every QP matrix here on every MPC tick
 * (SolverMPC.cpp:127-224); this build only (re)sizes the pinned staging record
 * when the horizon changes. */
EXTERNC void setup_problem(double dt, int horizon, double mu, double f_max);

/* reference: convexMPC_interface.cpp:88-105 (double inputs, narrowed to float). */
EXTERNC void update_problem_data(double* p, double* v, double* q, double* w, double* r, double yaw, double* weights, double* state_trajectory, double alpha, int* gait);

/* reference: convexMPC_interface.cpp:175-180.  0 before the first solve. */
EXTERNC double get_solution(int index);

/* reference: convexMPC_interface.cpp:107-119.  The JCQP settings are stored and
 * otherwise unused: the GPU engine always returns the active-set-exact optimum
 * the reference's qpOASES path (use_jcqp == 0, the only live one) returns. */
EXTERNC void update_solver_settings(int max_iter, double rho, double sigma, double solver_alpha, double terminate, double use_jcqp);

/* reference: convexMPC_interface.cpp:121-169.  Copies the inputs, solves
 * synchronously on cuda device 0, and returns once the 12*h solution is on the host. */
EXTERNC void update_problem_data_floats(float* p, float* v, float* q, float* w,
                                        float* r, float yaw, float* weights,
                                        float* state_trajectory, float alpha, int* gait);

/* reference: convexMPC_interface.h:48 / .cpp:171-173 -- declared OUTSIDE EXTERNC
 * upstream, so it carries C++ linkage (_Z13update_x_dragf); kept that way. */
void update_x_drag(float x_drag);

/* Additive (not in the reference): status of the last legacy solve.
 * 0 = optimal, see MPC_STATUS_* in mpc_batch.h; -1 = never solved. */
EXTERNC int mpc_last_status(void);
/* Additive: working-set iterations the last legacy solve took. */
EXTERNC int mpc_last_iterations(void);
/* Additive: body inertia diagonal [3] and mass, which the reference hard-codes
 * (RobotState.cpp:38-40, RobotState.h:23); defaults are those constants. */
EXTERNC void mpc_set_robot(const float* I_body_diag, float mass);
/* Additive: the inputs last given to setup_problem / update_x_drag / update_problem_data* as one batch
 * record of include/mpc_batch.h (out holds mpc_record_stride(horizon) bytes); returns the horizon. */
EXTERNC int mpc_legacy_record(void* out);
/* Additive: the use_jcqp mode last requested through update_solver_settings (0, 1, 2).  The reference's JCQP ADMM
 * alternative (SolverMPC.cpp:406-420, 558-619; never selected upstream, ConvexMPCLocomotion.cpp:649) has no
 * counterpart here: the active-set kernel returns the exact optimum of the same QP, the request is reported on
 * stderr once, and rho / sigma / solver_alpha / terminate / max_iter are ignored. */
EXTERNC int mpc_jcqp_requested(void);
/* Additive: destroys the cached GPU engines. */
EXTERNC void mpc_shutdown(void);
#endif
