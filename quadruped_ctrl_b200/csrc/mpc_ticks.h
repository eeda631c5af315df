// Tick record -> problem record (SURVEY 8f rows N1 + N2), one robot per thread.
//
// Restates, in the reference's own fp32 arithmetic and operation order,
//   ConvexMPCLocomotion::updateMPCIfNeeded   ConvexMPCLocomotion.cpp:498-577  (reference trajectory)
//   ConvexMPCLocomotion::solveDenseMPC       ConvexMPCLocomotion.cpp:592-640  (r, x_comp_integral)
//   OffsetDurationGait::getMpcTable          Gait.cpp:142-166                 (contact table)
// so that the record built on the device is byte-identical to the one the host path packs.  Products and sums
// are kept un-fused (__fmul_rn / __fadd_rn): the reference is compiled without FMA contraction.
#ifndef QUADRUPED_MPC_TICKS_H
#define QUADRUPED_MPC_TICKS_H

#include <stdint.h>
#include <string.h>

#include "../../include/mpc_batch.h"

#if defined(__CUDACC__)
#define MPC_TK_HD __host__ __device__ inline
#else
#define MPC_TK_HD inline
#endif

namespace mpc {

MPC_TK_HD float tk_mul(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fmul_rn(a, b);
#else
  volatile float r = a * b;  // volatile: keep the product rounded to float before it is added
  return r;
#endif
}
MPC_TK_HD float tk_add(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fadd_rn(a, b);
#else
  volatile float r = a + b;
  return r;
#endif
}

// tick: MPC_TICK_WORDS 32-bit words.  rec: mpc_record_stride(h) bytes (fully written, padding zeroed).
// state_out (optional): [4] floats.
MPC_TK_HD void build_record_from_tick(const float* tick, int h, char* rec_bytes, size_t stride, float* state_out) {
  float* rec = (float*)rec_bytes;
  const int32_t* ti = (const int32_t*)tick;
  // ---- state, passed through (solveDenseMPC :606-609) ----
  for (int i = 0; i < 3; i++) {
    rec[MPC_REC_P + i] = tick[MPC_TICK_P + i];
    rec[MPC_REC_V + i] = tick[MPC_TICK_V + i];
    rec[MPC_REC_W + i] = tick[MPC_TICK_W + i];
    rec[MPC_REC_IBODY + i] = tick[MPC_TICK_IBODY + i];
  }
  for (int i = 0; i < 4; i++) rec[MPC_REC_Q + i] = tick[MPC_TICK_Q + i];
  // r[i] = pFoot[i % 4][i / 4] - position[i / 4]   (:611-613)
  for (int i = 0; i < 12; i++)
    rec[MPC_REC_R + i] = tk_add(tick[MPC_TICK_PFOOT + (i % 4) * 3 + i / 4], -tick[MPC_TICK_P + i / 4]);
  rec[MPC_REC_YAW] = tick[MPC_TICK_YAW];
  rec[MPC_REC_XDRAG] = tick[MPC_TICK_XDRAG];
  rec[MPC_REC_ALPHA] = tick[MPC_TICK_ALPHA];
  for (int i = 0; i < 12; i++) rec[MPC_REC_WEIGHTS + i] = tick[MPC_TICK_WEIGHTS + i];
  rec[MPC_REC_MASS] = tick[MPC_TICK_MASS];
  rec[MPC_REC_DT] = tick[MPC_TICK_DT];
  rec[MPC_REC_MU] = tick[MPC_TICK_MU];
  rec[MPC_REC_FMAX] = tick[MPC_TICK_FMAX];
  rec[MPC_REC_RESERVED] = 0.f;
  // ---- reference trajectory (updateMPCIfNeeded :514-577) ----
  float* traj = rec + MPC_REC_TRAJ;
  const float dt = tick[MPC_TICK_DT];
  float xs = tick[MPC_TICK_POS_DES], ys = tick[MPC_TICK_POS_DES + 1];
  if (ti[MPC_TICK_STANDING]) {
    // trajInitial = {_roll_des, _pitch_des, stand_traj[5], stand_traj[0], stand_traj[1], _body_height, 0...}
    for (int i = 0; i < h; i++) {
      float* t = traj + 12 * i;
      t[0] = tick[MPC_TICK_RPY_COMP];
      t[1] = tick[MPC_TICK_RPY_COMP + 1];
      t[2] = tick[MPC_TICK_YAW_DES];
      t[3] = xs;
      t[4] = ys;
      t[5] = tick[MPC_TICK_HEIGHT];
      for (int j = 6; j < 12; j++) t[j] = 0.f;
    }
  } else {
    const float px = tick[MPC_TICK_P], py = tick[MPC_TICK_P + 1];
    const float max_pos_error = .1f;  // :534 (a float: "const float max_pos_error = .1")
    // the comparisons are float; the corrections add the DOUBLE literal 0.1 and narrow (:539-543)
    if (tk_add(xs, -px) > max_pos_error) xs = (float)((double)px + 0.1);
    if (tk_add(px, -xs) > max_pos_error) xs = (float)((double)px - 0.1);
    if (tk_add(ys, -py) > max_pos_error) ys = (float)((double)py + 0.1);
    if (tk_add(py, -ys) > max_pos_error) ys = (float)((double)py - 0.1);
    const float vx = tick[MPC_TICK_VDES], vy = tick[MPC_TICK_VDES + 1], yr = tick[MPC_TICK_YAW_RATE];
    float x = xs, y = ys, yaw = tick[MPC_TICK_YAW_DES];
    for (int i = 0; i < h; i++) {
      float* t = traj + 12 * i;
      if (i > 0) {  // running float sums, one product and one sum per step (:566-572)
        x = tk_add(x, tk_mul(dt, vx));
        y = tk_add(y, tk_mul(dt, vy));
        yaw = tk_add(yaw, tk_mul(dt, yr));
      }
      t[0] = tick[MPC_TICK_RPY_COMP];
      t[1] = tick[MPC_TICK_RPY_COMP + 1];
      t[2] = yaw;
      t[3] = x;
      t[4] = y;
      t[5] = tick[MPC_TICK_HEIGHT];
      t[6] = 0.f;
      t[7] = 0.f;
      t[8] = yr;
      t[9] = vx;
      t[10] = vy;
      t[11] = 0.f;
    }
  }
  // ---- contact table (Gait.cpp:142-166), then zero padding up to the stride ----
  unsigned char* gait = (unsigned char*)rec_bytes + 4 * (MPC_REC_TRAJ + 12 * h);
  const int it0 = ti[MPC_TICK_ITERATION];
  for (int i = 0; i < h; i++) {
    const int iter = (i + it0 + 1) % h;
    for (int j = 0; j < 4; j++) {
      int progress = iter - ti[MPC_TICK_OFFSETS + j];
      if (progress < 0) progress += h;
      gait[i * 4 + j] = (progress < ti[MPC_TICK_DURATIONS + j]) ? 1 : 0;
    }
  }
  for (size_t o = (size_t)4 * (MPC_REC_TRAJ + 12 * h) + 4 * h; o < stride; o++) rec_bytes[o] = 0;
  // ---- controller state the reference writes back at this point ----
  if (state_out) {
    state_out[0] = xs;  // world_position_desired (:545-546); unchanged for the standing trajectory
    state_out[1] = ys;
    // x_comp_integral += cmpc_x_drag * pz_err * dtMPC / vxy[0] when |v_x| > 0.3 (:634-640), float, left to right
    float xi = tick[MPC_TICK_XDRAG];
    const float vwx = tick[MPC_TICK_V];
    // the reference compares the float with the DOUBLE literal 0.3 (:636): v_x == 0.3f (0.300000012) integrates
    if ((double)vwx > 0.3 || (double)vwx < -0.3) {
      const float pz_err = tk_add(tick[MPC_TICK_P + 2], -tick[MPC_TICK_HEIGHT]);
      xi = tk_add(xi, tk_mul(tk_mul(3.0f, pz_err), dt) / vwx);
    }
    state_out[2] = xi;
    state_out[3] = 0.f;
  }
}

}  // namespace mpc
#endif
