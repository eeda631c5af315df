// The callers either side of the MPC solve that SURVEY 8f lists as rows N2 and N4, one robot per thread, in the
// reference's own fp32 arithmetic and operation order (un-fused products and sums: the reference build has no FMA):
//
//   gait_state_from_record   OffsetDurationGait::setIterations     Gait.cpp:187-193   (iteration, phase)
//                            OffsetDurationGait::getContactState   Gait.cpp:61-80
//                            OffsetDurationGait::getSwingState     Gait.cpp:97-123
//                            OffsetDurationGait::getMpcTable       Gait.cpp:142-166   (optional)
//   leg_commands_from_record f_ff = -rBody * f                     ConvexMPCLocomotion.cpp:672-685
//                            quaternionToRotationMatrix            Utilities/orientation_tools.h:170-188
//                            computeLegJacobianAndPosition         Controllers/LegController.cpp:204-240
//                            LegController::updateData (v = J qd)  Controllers/LegController.cpp:89-108
//                            LegController::updateCommand          Controllers/LegController.cpp:114-155
//
// Eigen evaluates a fixed-size 3-vector dot product as  a0*b0 + (a1*b1 + a2*b2)  (its reduction unroller halves the
// range: redux_novec_unroller<.., 0, 3> = func(<0,1>, <1,2>)); dot3() below follows that.  Eigen itself is not
// available in this build, so that order is taken from its source as of 3.3, the version the reference targets.
#ifndef QUADRUPED_MPC_LEGS_H
#define QUADRUPED_MPC_LEGS_H

#include <math.h>
#include <stdint.h>

#include "../../include/mpc_batch.h"
#include "mpc_ticks.h"

namespace mpc {

MPC_TK_HD float tk_div(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fdiv_rn(a, b);
#else
  volatile float r = a / b;
  return r;
#endif
}
MPC_TK_HD float dot3(float a0, float b0, float a1, float b1, float a2, float b2) {
  return tk_add(tk_mul(a0, b0), tk_add(tk_mul(a1, b1), tk_mul(a2, b2)));
}

// rec: MPC_GAIT_WORDS int32 words.  out: MPC_GAIT_STATE_WORDS 32-bit words.  table (optional): 4 * nIterations bytes.
MPC_TK_HD void gait_state_from_record(const int32_t* rec, float* out, unsigned char* table) {
  const int ipm = rec[MPC_GAIT_ITERATIONS_PER_MPC], cur = rec[MPC_GAIT_CURRENT_ITERATION], n = rec[MPC_GAIT_SEGMENTS];
  int32_t* outi = (int32_t*)out;
  // setIterations (Gait.cpp:187-193)
  const int iteration = (cur / ipm) % n;
  const float phase = tk_div((float)(cur % (ipm * n)), (float)(ipm * n));
  outi[MPC_GAIT_STATE_ITERATION] = iteration;
  out[MPC_GAIT_STATE_PHASE] = phase;
  for (int i = 0; i < 4; i++) {
    // setGaitParam (Gait.cpp:36-37): offsets / durations as fractions of the gait cycle
    const float off = tk_div((float)rec[MPC_GAIT_OFFSETS + i], (float)n);
    const float dur = tk_div((float)rec[MPC_GAIT_DURATIONS + i], (float)n);
    // getContactState (Gait.cpp:61-80)
    float progress = tk_add(phase, -off);
    if (progress < 0) progress = tk_add(progress, 1.f);
    if (progress > dur) progress = 0.f;
    else progress = tk_div(progress, dur);
    out[MPC_GAIT_STATE_CONTACT + i] = progress;
    // getSwingState (Gait.cpp:97-123)
    float swing_offset = tk_add(off, dur);
    if (swing_offset > 1) swing_offset = tk_add(swing_offset, -1.f);
    const float swing_duration = tk_add(1.f, -dur);
    float sp = tk_add(phase, -swing_offset);
    if (sp < 0) sp = tk_add(sp, 1.f);
    if (sp > swing_duration) sp = 0.f;
    else if ((double)swing_duration < 0.0000000001) sp = 0.f;
    else sp = tk_div(sp, swing_duration);
    out[MPC_GAIT_STATE_SWING + i] = sp;
  }
  if (table) {  // getMpcTable (Gait.cpp:142-166)
    for (int i = 0; i < n; i++) {
      const int iter = (i + iteration + 1) % n;
      for (int j = 0; j < 4; j++) {
        int progress = iter - rec[MPC_GAIT_OFFSETS + j];
        if (progress < 0) progress += n;
        table[i * 4 + j] = (progress < rec[MPC_GAIT_DURATIONS + j]) ? 1 : 0;
      }
    }
  }
}

// leg: MPC_LEG_WORDS 32-bit words.  forces: the 12 first-step forces of the solve (world frame, [leg*3+axis]).
// f_ff, tau: [12] each, [leg*3+axis] / [leg*3+joint].
MPC_TK_HD void leg_commands_from_record(const float* leg, const float* forces, float* f_ff, float* tau) {
  const int32_t* li = (const int32_t*)leg;
  // quaternionToRotationMatrix (orientation_tools.h:170-188): R as written there, then transposed in place
  const float e0 = leg[MPC_LEG_Q], e1 = leg[MPC_LEG_Q + 1], e2 = leg[MPC_LEG_Q + 2], e3 = leg[MPC_LEG_Q + 3];
  float Rt[3][3];  // the matrix BEFORE transposeInPlace; rBody(i, j) = Rt[j][i]
  Rt[0][0] = tk_add(1.f, -tk_mul(2.f, tk_add(tk_mul(e2, e2), tk_mul(e3, e3))));
  Rt[0][1] = tk_mul(2.f, tk_add(tk_mul(e1, e2), -tk_mul(e0, e3)));
  Rt[0][2] = tk_mul(2.f, tk_add(tk_mul(e1, e3), tk_mul(e0, e2)));
  Rt[1][0] = tk_mul(2.f, tk_add(tk_mul(e1, e2), tk_mul(e0, e3)));
  Rt[1][1] = tk_add(1.f, -tk_mul(2.f, tk_add(tk_mul(e1, e1), tk_mul(e3, e3))));
  Rt[1][2] = tk_mul(2.f, tk_add(tk_mul(e2, e3), -tk_mul(e0, e1)));
  Rt[2][0] = tk_mul(2.f, tk_add(tk_mul(e1, e3), -tk_mul(e0, e2)));
  Rt[2][1] = tk_mul(2.f, tk_add(tk_mul(e2, e3), tk_mul(e0, e1)));
  Rt[2][2] = tk_add(1.f, -tk_mul(2.f, tk_add(tk_mul(e1, e1), tk_mul(e2, e2))));
  const float l1 = leg[MPC_LEG_LINKS], l2 = leg[MPC_LEG_LINKS + 1], l3 = leg[MPC_LEG_LINKS + 2], l4 = leg[MPC_LEG_LINKS + 3];
  const float kpj = leg[MPC_LEG_JOINT_GAINS], kdj = leg[MPC_LEG_JOINT_GAINS + 1];
  for (int b = 0; b < 4; b++) {
    // f_ff[leg] = -rBody * f (ConvexMPCLocomotion.cpp:680); legs without a feed-forward force (swing) get zero
    const float fx = forces[3 * b], fy = forces[3 * b + 1], fz = forces[3 * b + 2];
    float ff[3];
    for (int i = 0; i < 3; i++) ff[i] = li[MPC_LEG_USE_FF + b] ? dot3(-Rt[0][i], fx, -Rt[1][i], fy, -Rt[2][i], fz) : 0.f;
    for (int i = 0; i < 3; i++) f_ff[3 * b + i] = ff[i];
    // computeLegJacobianAndPosition (LegController.cpp:204-240), float
    const float sideSign = (b == 0 || b == 2) ? -1.f : 1.f;  // Quadruped.h:85-89
    const float q0 = leg[MPC_LEG_JOINT_Q + 3 * b], q1 = leg[MPC_LEG_JOINT_Q + 3 * b + 1], q2 = leg[MPC_LEG_JOINT_Q + 3 * b + 2];
    // std::sin(float) upstream (glibc sinf, correctly rounded in practice); here the double routine rounded to float,
    // which is the same number on the host and on the device (device sinf alone is 1-2 ulp off)
    const float s1 = (float)sin((double)q0), s2 = (float)sin((double)q1), s3 = (float)sin((double)q2);
    const float c1 = (float)cos((double)q0), c2 = (float)cos((double)q1), c3 = (float)cos((double)q2);
    const float c23 = tk_add(tk_mul(c2, c3), -tk_mul(s2, s3));
    const float s23 = tk_add(tk_mul(s2, c3), tk_mul(c2, s3));
    const float l14s = tk_mul(tk_add(l1, l4), sideSign);
    float J[3][3], p[3];
    J[0][0] = 0.f;
    J[0][1] = tk_add(tk_mul(l3, c23), tk_mul(l2, c2));
    J[0][2] = tk_mul(l3, c23);
    J[1][0] = tk_add(tk_add(tk_mul(tk_mul(l3, c1), c23), tk_mul(tk_mul(l2, c1), c2)), -tk_mul(l14s, s1));
    J[1][1] = tk_add(tk_mul(tk_mul(-l3, s1), s23), -tk_mul(tk_mul(l2, s1), s2));
    J[1][2] = tk_mul(tk_mul(-l3, s1), s23);
    J[2][0] = tk_add(tk_add(tk_mul(tk_mul(l3, s1), c23), tk_mul(tk_mul(l2, c2), s1)), tk_mul(l14s, c1));
    J[2][1] = tk_add(tk_mul(tk_mul(l3, c1), s23), tk_mul(tk_mul(l2, c1), s2));
    J[2][2] = tk_mul(tk_mul(l3, c1), s23);
    p[0] = tk_add(tk_mul(l3, s23), tk_mul(l2, s2));
    p[1] = tk_add(tk_add(tk_mul(l14s, c1), tk_mul(l3, tk_mul(s1, c23))), tk_mul(tk_mul(l2, c2), s1));
    p[2] = tk_add(tk_add(tk_mul(l14s, s1), -tk_mul(l3, tk_mul(c1, c23))), -tk_mul(tk_mul(l2, c1), c2));
    // updateData: v = J * qd (LegController.cpp:106)
    const float* qd = leg + MPC_LEG_JOINT_QD + 3 * b;
    float v[3];
    for (int i = 0; i < 3; i++) v[i] = dot3(J[i][0], qd[0], J[i][1], qd[1], J[i][2], qd[2]);
    // updateCommand (LegController.cpp:118-131): Cartesian PD on top of the feed-forward force, then J' * force.
    // kpCartesian / kdCartesian are diagonal upstream (ConvexMPCLocomotion.cpp:245-254); a diagonal matrix times a
    // vector gives K_ii * d_i exactly (the other products are +-0).
    float force[3];
    for (int i = 0; i < 3; i++) {
      force[i] = ff[i];
      force[i] = tk_add(force[i], tk_mul(leg[MPC_LEG_KP + 3 * b + i], tk_add(leg[MPC_LEG_PDES + 3 * b + i], -p[i])));
      force[i] = tk_add(force[i], tk_mul(leg[MPC_LEG_KD + 3 * b + i], tk_add(leg[MPC_LEG_VDES + 3 * b + i], -v[i])));
    }
    for (int j = 0; j < 3; j++) {
      const float legTorque = tk_add(leg[MPC_LEG_TAU_FF + 3 * b + j], dot3(J[0][j], force[0], J[1][j], force[1], J[2][j], force[2]));
      // :136-154  crtlParam(2) * (0.0 - q) - crtlParam(3) * qd + legTorque, left to right.  The literal 0.0 makes
      // the first product and both sums DOUBLE operations (the second product stays float); the result is narrowed
      // when it is stored into LegCommand's float fields (RobotLegState.h:37-39).
      const float qj = leg[MPC_LEG_JOINT_Q + 3 * b + j], qdj = qd[j];
#if defined(__CUDA_ARCH__)
      const double t1 = __dmul_rn((double)kpj, 0.0 - (double)qj);
      tau[3 * b + j] = (float)__dadd_rn(__dadd_rn(t1, -(double)tk_mul(kdj, qdj)), (double)legTorque);
#else
      volatile double t1 = (double)kpj * (0.0 - (double)qj);
      volatile double t2 = t1 - (double)tk_mul(kdj, qdj);
      volatile double t3 = t2 + (double)legTorque;
      tau[3 * b + j] = (float)t3;
#endif
    }
  }
}

}  // namespace mpc
#endif
