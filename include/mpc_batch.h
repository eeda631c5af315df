/*
 * Batched convex-MPC engine for sm_100a -- C ABI (additive; the reference has
 * no batched entry point).  The legacy single-robot interface in
 * convexMPC_interface.h is a batch-of-one client of this API.
 *
 * One "problem" is one robot x horizon instance of the reference's
 * solve_mpc(update_data_t*, problem_setup*)
 *   (/root/reference/src/MPC_Ctrl/SolverMPC.cpp:296-639):
 * the condensed single-rigid-body QP  min 1/2 u'Hu + g'u  over the 12*h contact
 * forces, subject to the friction pyramid and 0 <= fz <= gait*f_max per
 * (step, leg), with swing legs eliminated (SolverMPC.cpp:441-525) and the
 * optimum returned as 12*h numbers, zeros for swing legs (SolverMPC.cpp:545-557).
 *
 * Plain pointers and sizes only; no torch / Eigen types cross this boundary.
 * Functions return 0 on success and a negative MPC_E_* code otherwise; they
 * never throw and never fall back to a CPU solver.
 */
#ifndef QUADRUPED_MPC_BATCH_H
#define QUADRUPED_MPC_BATCH_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------------
 * Problem record: one contiguous, 16-byte aligned block per problem (array of
 * records), so that a warp stages its whole problem with a single bulk copy.
 * All scalars are IEEE fp32, exactly the types update_data_t/problem_setup
 * carry (convexMPC_interface.h:13-38); I_body and mass are the constants the
 * reference hard-codes in RobotState (RobotState.cpp:38-40, RobotState.h:23),
 * promoted to per-problem inputs.
 *
 *   float index   field
 *   0..2          p[3]      world position                (update_data_t::p)
 *   3..5          v[3]      world velocity                (::v)
 *   6..9          q[4]      orientation (w,x,y,z)         (::q)
 *   10..12        w[3]      world angular velocity        (::w)
 *   13..24        r[12]     foot - COM, r[axis*4+leg]     (::r)
 *   25            yaw                                     (::yaw)
 *   26            x_drag                                  (::x_drag)
 *   27            alpha                                   (::alpha)
 *   28..39        weights[12]                             (::weights)
 *   40..42        I_body diagonal                         (RobotState.cpp:38)
 *   43            mass                                    (RobotState.h:23)
 *   44            dt                                      (problem_setup::dt)
 *   45            mu                                      (problem_setup::mu)
 *   46            f_max                                   (problem_setup::f_max)
 *   47            reserved (0)
 *   48..48+12h-1  traj[12*h], row-major per step          (::traj)
 *   then          gait[4*h] bytes, gait[step*4+leg] 0/1   (::gait)
 *   zero padding up to mpc_record_stride(h)
 * ---------------------------------------------------------------------- */
enum {
  MPC_REC_P = 0, MPC_REC_V = 3, MPC_REC_Q = 6, MPC_REC_W = 10, MPC_REC_R = 13,
  MPC_REC_YAW = 25, MPC_REC_XDRAG = 26, MPC_REC_ALPHA = 27, MPC_REC_WEIGHTS = 28,
  MPC_REC_IBODY = 40, MPC_REC_MASS = 43, MPC_REC_DT = 44, MPC_REC_MU = 45,
  MPC_REC_FMAX = 46, MPC_REC_RESERVED = 47, MPC_REC_TRAJ = 48
};
#define MPC_MAX_HORIZON 36 /* traj[12*36] in update_data_t caps h (convexMPC_interface.h:3,30) */

/* Bytes between consecutive records for horizon h: 4*(48+12h)+4h rounded up to 16. */
size_t mpc_record_stride(int horizon);
/* Byte offset of the gait table inside a record. */
size_t mpc_record_gait_offset(int horizon);

/* ------------------------------------------------------------------------
 * Tick record (SURVEY 8f, rows N1 + N2): the compact per-robot input of one MPC
 * tick, from which the engine builds the problem record ON THE DEVICE -- the
 * reference does this on the host in ConvexMPCLocomotion::updateMPCIfNeeded
 * (ConvexMPCLocomotion.cpp:498-577: reference trajectory), solveDenseMPC
 * (:592-621: r = pFoot - position, weights, alpha) and
 * OffsetDurationGait::getMpcTable (Gait.cpp:142-166: contact table).
 * 68 32-bit words (272 bytes, 16-byte aligned), fp32 unless noted:
 *   0..2 p   3..5 vWorld   6..9 q(w,x,y,z)   10..12 omegaWorld
 *   13..24  pFoot[leg*3+axis], WORLD foot positions (not COM-relative)
 *   25 yaw (seResult.rpy[2])   26 x_drag (x_comp_integral)   27 alpha
 *   28..39 weights   40..42 I_body   43 mass   44 dtMPC   45 mu   46 f_max
 *   47 body_height
 *   48,49 rpy_comp[0..1]           (standing: _roll_des, _pitch_des)
 *   50    _yaw_des_true            (standing: stand_traj[5])
 *   51,52 world_position_desired   (standing: stand_traj[0], stand_traj[1])
 *   53    _yaw_turn_rate   54,55 v_des_world x,y
 *   56 (int32) 1 when current_gait == 4 (standing trajectory, :515-531)
 *   57 (int32) gait iteration (_iteration after Gait::setIterations)
 *   58..61 (int32) offsets[4]   62..65 (int32) durations[4]   66,67 reserved
 * The gait has `horizon` segments (nIterations == horizonLength upstream).
 * ---------------------------------------------------------------------- */
enum {
  MPC_TICK_P = 0, MPC_TICK_V = 3, MPC_TICK_Q = 6, MPC_TICK_W = 10, MPC_TICK_PFOOT = 13,
  MPC_TICK_YAW = 25, MPC_TICK_XDRAG = 26, MPC_TICK_ALPHA = 27, MPC_TICK_WEIGHTS = 28,
  MPC_TICK_IBODY = 40, MPC_TICK_MASS = 43, MPC_TICK_DT = 44, MPC_TICK_MU = 45, MPC_TICK_FMAX = 46,
  MPC_TICK_HEIGHT = 47, MPC_TICK_RPY_COMP = 48, MPC_TICK_YAW_DES = 50, MPC_TICK_POS_DES = 51,
  MPC_TICK_YAW_RATE = 53, MPC_TICK_VDES = 54, MPC_TICK_STANDING = 56, MPC_TICK_ITERATION = 57,
  MPC_TICK_OFFSETS = 58, MPC_TICK_DURATIONS = 62, MPC_TICK_WORDS = 68
};
#define MPC_TICK_STRIDE (4 * MPC_TICK_WORDS)

/* ------------------------------------------------------------------------
 * Gait record / gait state: OffsetDurationGait on the device.
 * Input, 12 int32 words per robot: iterationsPerMPC, currentIteration, nIterations (segments of the gait cycle),
 * reserved, offsets[4], durations[4] -- the arguments of Gait::setIterations (Gait.cpp:187-193) and the gait
 * definition (Gait.cpp:23-41).  Output, 10 32-bit words per robot: _iteration (int32), _phase (fp32),
 * getContactState()[4] (Gait.cpp:61-80), getSwingState()[4] (Gait.cpp:97-123); optionally getMpcTable()
 * (Gait.cpp:142-166) as 4*nIterations bytes.
 * ---------------------------------------------------------------------- */
enum {
  MPC_GAIT_ITERATIONS_PER_MPC = 0, MPC_GAIT_CURRENT_ITERATION = 1, MPC_GAIT_SEGMENTS = 2,
  MPC_GAIT_OFFSETS = 4, MPC_GAIT_DURATIONS = 8, MPC_GAIT_WORDS = 12
};
enum {
  MPC_GAIT_STATE_ITERATION = 0, MPC_GAIT_STATE_PHASE = 1, MPC_GAIT_STATE_CONTACT = 2, MPC_GAIT_STATE_SWING = 6,
  MPC_GAIT_STATE_WORDS = 10
};

/* ------------------------------------------------------------------------
 * Leg record: what the reference does with the 12 solved forces, per robot, on the device --
 * f_ff = -rBody * f (ConvexMPCLocomotion.cpp:672-685, rBody from the orientation quaternion,
 * orientation_tools.h:170-188) and LegController::updateCommand (LegController.cpp:114-155: Cartesian PD on top
 * of the feed-forward force, torque = tauFeedForward + J' force with J from computeLegJacobianAndPosition
 * (:204-240), joint-space damping).  100 fp32 words (400 bytes):
 *   0..3   orientation (w,x,y,z)            4..15  joint angles q[leg*3+joint]     16..27 joint rates qd
 *   28..39 pDes[leg*3+axis]                 40..51 vDes                             52..63 kpCartesian diagonal
 *   64..75 kdCartesian diagonal             76..87 tauFeedForward[leg*3+joint]
 *   88,89  crtlParam(2), crtlParam(3) (joint-space kp, kd)
 *   90..93 (int32) 1 where the leg takes the MPC force as forceFeedForward (stance), 0 = none (swing)
 *   94..97 link lengths: abad, hip, knee, knee y-offset (MiniCheetah.h:31-37: 0.062, 0.209, 0.195, 0.004)
 *   98,99  reserved
 * ---------------------------------------------------------------------- */
enum {
  MPC_LEG_Q = 0, MPC_LEG_JOINT_Q = 4, MPC_LEG_JOINT_QD = 16, MPC_LEG_PDES = 28, MPC_LEG_VDES = 40, MPC_LEG_KP = 52,
  MPC_LEG_KD = 64, MPC_LEG_TAU_FF = 76, MPC_LEG_JOINT_GAINS = 88, MPC_LEG_USE_FF = 90, MPC_LEG_LINKS = 94,
  MPC_LEG_WORDS = 100
};

/* Per-problem status word written next to the forces:
 *   bits 0..7   code (MPC_STATUS_*), bits 8..31 working-set iterations taken. */
enum {
  MPC_STATUS_OPTIMAL = 0,      /* KKT-exact optimum of the reduced QP                      */
  MPC_STATUS_MAX_ITER = 1,     /* iteration cap hit; forces = 0 (a dual active-set iterate is not primal feasible) */
  MPC_STATUS_BAD_INPUT = 2,    /* non-finite input or mu/mass/inertia/dt <= 0; forces = 0  */
  MPC_STATUS_NOT_PD = 3,       /* Hessian not positive definite (alpha <= 0 with zero weights) */
  MPC_STATUS_NO_STANCE = 4     /* every leg in swing over the whole horizon: all-zero optimum */
};
#define MPC_STATUS_CODE(s) ((s) & 0xff)
#define MPC_STATUS_ITERS(s) (((unsigned)(s)) >> 8)

/* error codes */
enum {
  MPC_OK = 0, MPC_E_ARG = -1, MPC_E_CUDA = -2, MPC_E_NOMEM = -3, MPC_E_NODEVICE = -4
};

typedef struct mpc_batch mpc_batch_t; /* opaque engine handle, one per (device, horizon) */

/* Creates an engine on CUDA device `device` for horizon `horizon`
 * (1..MPC_MAX_HORIZON) able to solve up to `max_batch` problems per call.
 * Fails with MPC_E_NODEVICE when no sm_100 device is present. */
int mpc_batch_create(mpc_batch_t** out, int device, int horizon, int max_batch);
void mpc_batch_destroy(mpc_batch_t* eng);

/* Device-resident solve, asynchronous on `cuda_stream` (a cudaStream_t; NULL =
 * legacy default stream).  Uses slot 0's device scratch: do not overlap two
 * device-resident solves of one engine on different streams.
 *   records_dev  [batch * mpc_record_stride(h)] bytes, 16-byte aligned
 *   forces_dev   [batch * 12] fp32: first-horizon-step forces, force[leg*3+axis],
 *                world frame, exactly what get_solution(0..11) returns upstream
 *   solution_dev optional [batch * 12*h] fp64: the whole q_soln vector
 *                (SolverMPC.cpp:545-557); NULL to skip
 *   status_dev   optional [batch] int32 status words; NULL to skip            */
int mpc_batch_solve_device(mpc_batch_t* eng, const void* records_dev, int batch,
                           float* forces_dev, double* solution_dev,
                           int32_t* status_dev, void* cuda_stream);

#define MPC_BATCH_SLOTS 6
/* The same on scratch slot `slot` (0 .. MPC_BATCH_SLOTS-1).  Device-resident solves of one engine may overlap when they
 * use different slots and different streams: the tail of one batch then shares the GPU with the head of the
 * next (independent batches; nothing is exchanged between them). */
int mpc_batch_solve_device_slot(mpc_batch_t* eng, int slot, const void* records_dev, int batch,
                           float* forces_dev, double* solution_dev,
                           int32_t* status_dev, void* cuda_stream);

/* Host-resident solve: stages records through pinned memory, runs the device
 * solve, copies forces/solution/status back and synchronises. */
int mpc_batch_solve_host(mpc_batch_t* eng, const void* records_host, int batch,
                         float* forces_host, double* solution_host,
                         int32_t* status_host);

/* Pipelined host-resident solve.  The engine has MPC_BATCH_SLOTS slots, each with its own stream, pinned
 * staging and device buffers (allocated on first use).  submit COPIES `records_host` into the slot's pinned
 * staging buffer before it returns -- the caller may reuse its buffer at once, whatever kind of memory it is --
 * and queues H2D, kernels and D2H on the slot's stream without waiting for the GPU; wait blocks until that slot
 * is done and copies the results out (pass the slot's own pinned pointers from mpc_batch_host_buffers to skip
 * the copies: records written straight into the slot's buffer are not copied again).
 * Alternating slots overlaps one batch's transfers and host-side packing with the other's kernels.
 * mpc_batch_solve_host == submit(slot 0) + wait(slot 0). */
int mpc_batch_submit_host(mpc_batch_t* eng, int slot, const void* records_host, int batch,
                          int want_solution);
/* Zero-copy variant, opt-in: `records_pinned` must be page-locked (cudaHostAlloc / cudaHostRegister; checked,
 * MPC_E_ARG otherwise) and MUST NOT BE MODIFIED until mpc_batch_wait_host(slot) returns -- the DMA engine reads
 * it in place after this call has returned. */
int mpc_batch_submit_host_pinned(mpc_batch_t* eng, int slot, const void* records_pinned, int batch,
                                 int want_solution);
int mpc_batch_wait_host(mpc_batch_t* eng, int slot, float* forces_host, double* solution_host,
                        int32_t* status_host);
/* Both submit entries classify the batch on the host while they stage it: when every problem falls into one size
 * class (one gait, one horizon -- the usual batch) the slot runs ONE kernel launch (no classify kernel, no index
 * lists, no empty-class launches).  Mixed batches take the general path.  Results are identical either way. */

/* The same from TICK records (layout above; SURVEY 8f N1 + N2): MPC_TICK_STRIDE = 272 bytes per robot cross the
 * bus instead of the problem record (720 bytes at h = 10); the problem records are built on the device into the
 * slot's own buffer.  zero_copy != 0: `ticks_host` is page-locked (checked) and stays untouched until
 * mpc_batch_wait_host(slot); otherwise it is copied before the call returns.  Collect with mpc_batch_wait_host.
 * mpc_batch_host_state: the slot's pinned [max_batch*4] fp32 controller state written back by the tick builder
 * (world_position_desired x, y after the clamp, next x_comp_integral, 0 -- see mpc_batch_build_records_device),
 * valid after wait_host, and the slot's own pinned tick buffer (filling it in place skips the staging copy). */
int mpc_batch_submit_host_ticks(mpc_batch_t* eng, int slot, const void* ticks_host, int batch,
                                int want_solution, int zero_copy);
int mpc_batch_host_state(mpc_batch_t* eng, int slot, float** state_host, void** ticks_pinned);
/* MPC_BATCH_SLOTS of the library that is loaded. */
int mpc_batch_slots(void);

/* Builds problem records from tick records on the device (one thread per robot), asynchronously on
 * `cuda_stream`.  records_dev [batch * mpc_record_stride(h)].  state_out_dev (optional, [batch*4] fp32)
 * receives what the reference writes back into its controller state at this point:
 * world_position_desired x, y after the 0.1 m clamp (ConvexMPCLocomotion.cpp:536-545), the next
 * x_comp_integral (:636-640) and 0. */
int mpc_batch_build_records_device(mpc_batch_t* eng, const void* ticks_dev, int batch,
                                   void* records_dev, float* state_out_dev, void* cuda_stream);
/* Build + solve in one call: records go to the engine's own device buffer (slot 0) and never
 * leave the GPU. */
int mpc_batch_solve_ticks_device(mpc_batch_t* eng, const void* ticks_dev, int batch,
                                 float* forces_dev, double* solution_dev, int32_t* status_dev,
                                 float* state_out_dev, void* cuda_stream);

/* SURVEY 8f row N2 on the device, one robot per thread: gait_dev [batch][MPC_GAIT_WORDS] int32 ->
 * state_out_dev [batch][MPC_GAIT_STATE_WORDS]; table_out_dev (optional) receives each robot's contact table at
 * table_out_dev + i * table_stride (4 * nIterations bytes each; table_stride >= 4 * the largest nIterations). */
int mpc_batch_gait_state_device(mpc_batch_t* eng, const void* gait_dev, int batch, void* state_out_dev,
                                unsigned char* table_out_dev, int table_stride, void* cuda_stream);
/* SURVEY 8f row N4 on the device, one robot per thread: legs_dev [batch][MPC_LEG_WORDS], forces_dev [batch*12]
 * (what a solve returned) -> f_ff_dev [batch*12] body-frame feed-forward forces, tau_dev [batch*12] joint torques
 * (tau_abad/hip/knee_ff per leg, [leg*3+joint]). */
int mpc_batch_leg_commands_device(mpc_batch_t* eng, const void* legs_dev, const float* forces_dev, int batch,
                                  float* f_ff_dev, float* tau_dev, void* cuda_stream);

/* Debug / parity entry: assembles the reduced QP only and writes it out.
 *   nvar_dev [batch] int32: reduced variable count nv = 3 * (#stance (step,leg))
 *   H_dev    [batch * (12h)*(12h)] fp64 row-major, leading dimension 12h; top-left nv x nv used
 *   g_dev    [batch * 12h] fp64                                             */
int mpc_batch_assemble_device(mpc_batch_t* eng, const void* records_dev, int batch,
                              int32_t* nvar_dev, double* H_dev, double* g_dev,
                              void* cuda_stream);

/* Shard-and-gather epilogue for a batch split over several GPUs: when set, the
 * solve kernel also stores each problem's 12 forces straight into every peer's
 * gather buffer at row (rank_offset + i) through NVLink peer mappings, which
 * replaces the separate all-gather.  peers[k] is the base of rank k's
 * [world_batch * 12] fp32 gather buffer as mapped into THIS process (or NULL to
 * skip rank k); pass n_peers = 0 to turn the epilogue off. */
int mpc_batch_set_gather_peers(mpc_batch_t* eng, float* const* peers, int n_peers,
                               int rank_offset);

/* The same over CUDA IPC for one-process-per-GPU jobs:
 *   1. every rank: gather_alloc -> its gather buffer [world_batch*12] fp32 + the 64-byte IPC handle;
 *   2. the ranks exchange the handles (any transport, e.g. torch.distributed.all_gather);
 *   3. every rank: gather_connect(handles[world][64], world, rank, rank_offset) opens the peers'
 *      buffers and arms the epilogue (rank_offset = first global row of this rank's shard).
 * After a solve on every rank and one cross-rank barrier, every rank's buffer (gather_buffer)
 * holds the forces of the whole batch; no separate all-gather runs. */
int mpc_batch_gather_alloc(mpc_batch_t* eng, int world_batch, void* ipc_handle_out);
int mpc_batch_gather_connect(mpc_batch_t* eng, const void* ipc_handles, int world, int rank,
                             int rank_offset);
void* mpc_batch_gather_buffer(mpc_batch_t* eng);
/* The gather buffer exists once per scratch slot (mpc_batch_solve_device_slot): a solve on slot q stores into
 * every rank's slot-q region and is followed by mpc_batch_gather_sync_slot(q), so two batches can be in flight. */
void* mpc_batch_gather_buffer_slot(mpc_batch_t* eng, int slot);
/* The cross-rank barrier of the fused gather, on the device: queued on `cuda_stream` after a solve,
 * it tells every peer (a release store into its flag word over NVLink) that this rank's rows have
 * landed and waits until every peer has said the same here.  Work queued behind it on the stream
 * sees the complete gather buffer.  Every rank must call it once per solve.
 * The hand-shake orders the stores BEFORE the reads.  The other direction is the caller's: a rank must not start
 * its next solve on a slot while some rank still reads that slot's region -- any cross-rank barrier after the reads
 * does (one more gather_sync on the slot is such a barrier), or rotate over the slots and consume a region before
 * the solve after next is queued. */
int mpc_batch_gather_sync(mpc_batch_t* eng, void* cuda_stream);
int mpc_batch_gather_sync_slot(mpc_batch_t* eng, int slot, void* cuda_stream);
/* The same gather by the COPY ENGINES instead of the solve kernel's epilogue: after gather_connect, turn the epilogue
 * off (set_gather_fused 0) and call gather_push_slot once per solve, on any stream ordered after the solve (a
 * communication stream of its own, so that the next batch's kernels never wait for it): this rank's [batch, 12] forces
 * at forces_dev are DMA-copied over NVLink into region `slot` of every rank's gather buffer at this rank's row offset
 * -- no SM takes part, nothing competes with the solve kernels in flight -- and the flag barrier of gather_sync_slot
 * follows.  Work queued behind it on the stream sees the whole batch in mpc_batch_gather_buffer_slot(slot). */
int mpc_batch_set_gather_fused(mpc_batch_t* eng, int on);
int mpc_batch_gather_push_slot(mpc_batch_t* eng, int slot, const float* forces_dev, int batch, void* cuda_stream);

/* Inversion of the reduced Hessian in the register-resident size classes (selectable per engine, any time):
 *   0  symmetric sweep with one rank-1 update per pivot on the FP64 FMA pipe (invert_spd_tiles)
 *   1  grouped symmetric sweep, rank-8 updates / panel / pivot block as DMMA.8x8x4 on the FP64 tensor pipe
 *      (invert_spd_mma) -- the "tensor-core path" of BASELINE config 5
 * Both hold 1e-9 against the reference solver on the fp64-assembled QP; the results differ in the last bits.  Environment MPC_SWEEP=fma|mma sets
 * the default of new engines. */
int mpc_batch_set_sweep_variant(mpc_batch_t* eng, int variant);
int mpc_batch_sweep_variant(const mpc_batch_t* eng);

/* Solver of the size classes with nv <= 128 (selectable per engine, any time):
 *   0  explicit inverse of the reduced condensed Hessian (symmetric sweep in registers) + dual active set on it
 *   1  Riccati sweeps (default): the condensed Hessian is never formed; the gains of the horizon's Riccati recursion
 *      are factored once per problem and every H^{-1} product of the dual active-set method is one backward and one
 *      forward sweep over the horizon (csrc/mpc_riccati.h), one warp per problem
 * Both end at the same KKT point (1e-9 against the reference solver on the fp64-assembled QP is asserted for
 * either; they differ in the last bits).  The warm start, the phase-clock and the assemble-only entries always use
 * solver 0.  Environment MPC_SOLVER=riccati|inverse sets the default of new engines. */
int mpc_batch_set_solver(mpc_batch_t* eng, int solver);
int mpc_batch_solver(const mpc_batch_t* eng);

/* Warm start across MPC ticks (the reference cold-starts every solve, SolverMPC.cpp:529).
 * cache_dev: device memory, [robots][mpc_batch_warm_stride()] int32, zero-initialised by the caller once; the engine
 * reads a robot's entry before its solve and rewrites it afterwards (the optimal working set as (step, leg, row)
 * codes).  robot_ids_dev: [batch] int32 robot id of every problem of the coming solves (NULL: the problem's index).
 * shift: horizon steps the gait table has advanced since the cached solve (1 for consecutive MPC ticks, 0 to
 * re-solve the same tick).  The result is the cold-start optimum (the QP is strictly convex; the cache only decides
 * where the dual active-set method starts), in fewer working-set changes.  cache_dev = NULL turns it off.
 * A robot must not appear twice in one batch. */
int mpc_batch_set_warm_start(mpc_batch_t* eng, int* cache_dev, const int* robot_ids_dev, int shift);
int mpc_batch_warm_stride(void);

/* Iteration cap of the active-set loop (working-set additions); default 4000.
 * The reference caps qpOASES at nWSR = 100 (SolverMPC.cpp:435) and returns stale
 * memory beyond it; this engine reports MPC_STATUS_MAX_ITER instead. */
int mpc_batch_set_max_iterations(mpc_batch_t* eng, int max_iter);

/* Turns CUDA-event timing of the solve kernels on or off (off by default). */
int mpc_batch_set_timing(mpc_batch_t* eng, int enabled);

/* Profiling aid: when dev_buf (device memory, [max_batch][24] int64) is non-NULL the solve
 * kernel stamps clock64() per problem at: record landed, assembled, inverted, active set
 * done, outputs written; slots 5,6 = CTA index and its pass number, 8..12 = stamps inside the assembly / active-set stages.  NULL turns it off. */
int mpc_batch_set_phase_clock_buffer(mpc_batch_t* eng, long long* dev_buf);

/* Tuning aid: caps the persistent grid of every class at limit * (number of SMs) CTAs;
 * 0 restores the occupancy-derived default. */
int mpc_batch_set_ctas_per_sm_limit(mpc_batch_t* eng, int limit);

/* Turns timing on for ONE size class only (two events per solve instead of two per class);
 * idx = -1 times every class again. */
int mpc_batch_set_timed_class(mpc_batch_t* eng, int idx);

/* Size classes the engine sorts problems into (by reduced variable count).
 * info[6] = { nv_cap, m_cap, threads per CTA, grid, shared-memory bytes per CTA,
 *             1 if the QP tile lives in shared memory (0: per-CTA global slab) }. */
int mpc_batch_num_classes(const mpc_batch_t* eng);
int mpc_batch_class_info(const mpc_batch_t* eng, int idx, int* info);

/* Number of kernels the engine has launched since creation (for accounting). */
long mpc_batch_kernel_launches(const mpc_batch_t* eng);
/* Device time in ms of the solve kernels of the most recent solve call, measured
 * with CUDA events on the call's stream (synchronises that stream). */
float mpc_batch_last_solve_kernel_ms(mpc_batch_t* eng);
/* Same, for the solve kernel of size class `idx` alone. */
float mpc_batch_last_class_kernel_ms(mpc_batch_t* eng, int idx);
/* Timing over a region without synchronising inside it: mark, run up to 256 solves,
 * then collect the mean duration of size class `idx`'s solve kernel over the solves
 * since the mark (CUDA events recorded on each call's own stream). */
void mpc_batch_timing_mark(mpc_batch_t* eng);
int mpc_batch_timing_collect(mpc_batch_t* eng, int idx, float* mean_ms, int* n_solves);
/* Slot `slot`'s pinned host staging buffers ([max_batch] records / forces / solution /
 * status).  A caller that fills *records and passes these same pointers to
 * mpc_batch_solve_host (slot 0) / submit_host / wait_host skips the pageable->pinned copies. */
int mpc_batch_host_buffers(mpc_batch_t* eng, int slot, void** records, float** forces,
                           double** solution, int32_t** status);
/* Slot `slot`'s DEVICE result buffers of the host entry ([max_batch*12] fp32 forces, [max_batch] int32 status):
 * valid after mpc_batch_wait_host(slot) until the next submit on that slot.  Lets a sharded job hand the forces
 * to its gather without uploading again what was just downloaded. */
int mpc_batch_device_buffers(mpc_batch_t* eng, int slot, float** forces_dev, int32_t** status_dev);
/* Human-readable description of the last error on this engine ("" if none). */
const char* mpc_batch_last_error(const mpc_batch_t* eng);
/* Library-level: text of the last error when no engine exists (create failed). */
const char* mpc_last_error(void);
/* Horizon the engine was built for / its record stride. */
int mpc_batch_horizon(const mpc_batch_t* eng);

#ifdef __cplusplus
}
#endif
#endif
