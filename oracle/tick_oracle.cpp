// TEST INFRASTRUCTURE -- not part of the product path.
//
// CPU restatement of what the reference does on the host between the state estimate and the MPC C interface:
//   ConvexMPCLocomotion::updateMPCIfNeeded   /root/reference/src/MPC_Ctrl/ConvexMPCLocomotion.cpp:498-577
//   ConvexMPCLocomotion::solveDenseMPC       /root/reference/src/MPC_Ctrl/ConvexMPCLocomotion.cpp:592-664
//   OffsetDurationGait::getMpcTable          /root/reference/src/MPC_Ctrl/Gait.cpp:142-166
//   update_problem_data_floats (gait int -> u8, convexMPC_interface.cpp:75-79,121-169)
// written the way the reference writes it (trajInitial[12] copied into every step, then the running sums),
// in float, compiled with -ffp-contract=off like the reference's non-FMA x86-64 build.  It is the oracle for the
// device-side record builder (csrc/mpc_ticks.h); the two share no code.
#include <cstdint>
#include <cstring>

#include "../include/mpc_batch.h"

extern "C" {

size_t oracle_record_stride(int h);

// ticks: [batch][MPC_TICK_WORDS] 32-bit words; records: [batch][stride] bytes; state_out: [batch][4] floats or NULL
void oracle_build_records(const float* ticks, int batch, int horizonLength, unsigned char* records, float* state_out) {
  const size_t stride = ((size_t)(4 * (MPC_REC_TRAJ + 12 * horizonLength) + 4 * horizonLength) + 15) / 16 * 16;
  for (int b = 0; b < batch; b++) {
    const float* tk = ticks + (size_t)b * MPC_TICK_WORDS;
    const int32_t* ti = (const int32_t*)tk;
    unsigned char* rec_b = records + stride * b;
    memset(rec_b, 0, stride);
    float* rec = (float*)rec_b;

    // ---- updateMPCIfNeeded (:498-577) ----
    const float* p = tk + MPC_TICK_P;  // seResult.position
    float trajAll[12 * 36];
    const float dtMPC = tk[MPC_TICK_DT];
    float world_position_desired[2] = {tk[MPC_TICK_POS_DES], tk[MPC_TICK_POS_DES + 1]};
    const float _body_height = tk[MPC_TICK_HEIGHT];
    const float _yaw_turn_rate = tk[MPC_TICK_YAW_RATE];
    const float v_des_world[2] = {tk[MPC_TICK_VDES], tk[MPC_TICK_VDES + 1]};
    if (ti[MPC_TICK_STANDING]) {  // current_gait == 4 (:515-531)
      float trajInitial[12] = {tk[MPC_TICK_RPY_COMP] /*_roll_des*/, tk[MPC_TICK_RPY_COMP + 1] /*_pitch_des*/,
                               tk[MPC_TICK_YAW_DES] /*stand_traj[5]*/, tk[MPC_TICK_POS_DES] /*stand_traj[0]*/,
                               tk[MPC_TICK_POS_DES + 1] /*stand_traj[1]*/, _body_height, 0, 0, 0, 0, 0, 0};
      for (int i = 0; i < horizonLength; i++)
        for (int j = 0; j < 12; j++) trajAll[12 * i + j] = trajInitial[j];
    } else {  // (:533-577)
      const float max_pos_error = .1;
      float xStart = world_position_desired[0];
      float yStart = world_position_desired[1];
      if (xStart - p[0] > max_pos_error) xStart = p[0] + 0.1;
      if (p[0] - xStart > max_pos_error) xStart = p[0] - 0.1;
      if (yStart - p[1] > max_pos_error) yStart = p[1] + 0.1;
      if (p[1] - yStart > max_pos_error) yStart = p[1] - 0.1;
      world_position_desired[0] = xStart;
      world_position_desired[1] = yStart;
      float trajInitial[12] = {tk[MPC_TICK_RPY_COMP], tk[MPC_TICK_RPY_COMP + 1], tk[MPC_TICK_YAW_DES] /*_yaw_des_true*/,
                               xStart, yStart, _body_height, 0, 0, _yaw_turn_rate, v_des_world[0], v_des_world[1], 0};
      for (int i = 0; i < horizonLength; i++) {
        for (int j = 0; j < 12; j++) trajAll[12 * i + j] = trajInitial[j];
        if (i == 0) {
          trajAll[2] = tk[MPC_TICK_YAW_DES];
        } else {
          trajAll[12 * i + 3] = trajAll[12 * (i - 1) + 3] + dtMPC * v_des_world[0];
          trajAll[12 * i + 4] = trajAll[12 * (i - 1) + 4] + dtMPC * v_des_world[1];
          trajAll[12 * i + 2] = trajAll[12 * (i - 1) + 2] + dtMPC * _yaw_turn_rate;
        }
      }
    }

    // ---- getMpcTable (Gait.cpp:142-166), _nIterations == horizonLength ----
    int mpcTable[4 * 36];
    const int _iteration = ti[MPC_TICK_ITERATION];
    for (int i = 0; i < horizonLength; i++) {
      int iter = (i + _iteration + 1) % horizonLength;
      for (int j = 0; j < 4; j++) {
        int progress = iter - ti[MPC_TICK_OFFSETS + j];
        if (progress < 0) progress += horizonLength;
        mpcTable[i * 4 + j] = (progress < ti[MPC_TICK_DURATIONS + j]) ? 1 : 0;
      }
    }

    // ---- solveDenseMPC (:592-664) ----
    float r[12];
    for (int i = 0; i < 12; i++) r[i] = tk[MPC_TICK_PFOOT + (i % 4) * 3 + i / 4] - p[i / 4];  // pFoot[i%4][i/4] - position[i/4]
    const float pz_err = p[2] - _body_height;
    const float vxy0 = tk[MPC_TICK_V];
    float x_comp_integral = tk[MPC_TICK_XDRAG];
    const float x_drag_used = x_comp_integral;  // update_x_drag(x_comp_integral) precedes the integrator update
    const float cmpc_x_drag = 3.0;
    if (vxy0 > 0.3 || vxy0 < -0.3) x_comp_integral += cmpc_x_drag * pz_err * dtMPC / vxy0;

    // ---- what update_problem_data_floats / setup_problem / update_x_drag hand to solve_mpc, as one record ----
    memcpy(rec + MPC_REC_P, p, 12);
    memcpy(rec + MPC_REC_V, tk + MPC_TICK_V, 12);
    memcpy(rec + MPC_REC_Q, tk + MPC_TICK_Q, 16);
    memcpy(rec + MPC_REC_W, tk + MPC_TICK_W, 12);
    memcpy(rec + MPC_REC_R, r, 48);
    rec[MPC_REC_YAW] = tk[MPC_TICK_YAW];
    rec[MPC_REC_XDRAG] = x_drag_used;
    rec[MPC_REC_ALPHA] = tk[MPC_TICK_ALPHA];
    memcpy(rec + MPC_REC_WEIGHTS, tk + MPC_TICK_WEIGHTS, 48);
    memcpy(rec + MPC_REC_IBODY, tk + MPC_TICK_IBODY, 12);
    rec[MPC_REC_MASS] = tk[MPC_TICK_MASS];
    rec[MPC_REC_DT] = dtMPC;
    rec[MPC_REC_MU] = tk[MPC_TICK_MU];
    rec[MPC_REC_FMAX] = tk[MPC_TICK_FMAX];
    memcpy(rec + MPC_REC_TRAJ, trajAll, sizeof(float) * 12 * horizonLength);
    unsigned char* gait = rec_b + 4 * (MPC_REC_TRAJ + 12 * horizonLength);
    for (int i = 0; i < 4 * horizonLength; i++) gait[i] = (unsigned char)mpcTable[i];  // mint_to_u8
    if (state_out) {
      state_out[4 * b + 0] = world_position_desired[0];
      state_out[4 * b + 1] = world_position_desired[1];
      state_out[4 * b + 2] = x_comp_integral;
      state_out[4 * b + 3] = 0.f;
    }
  }
}

}  // extern "C"
