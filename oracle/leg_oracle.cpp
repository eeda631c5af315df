// CPU ORACLE -- TEST INFRASTRUCTURE ONLY (never linked into or loaded by the product).
//
// Literal restatement, with small matrix helpers that evaluate the way Eigen evaluates the reference's
// expressions, of the host code around the MPC solve that SURVEY 8f lists as rows N2 and N4:
//   OffsetDurationGait::setIterations / getContactState / getSwingState / getMpcTable
//       /root/reference/src/MPC_Ctrl/Gait.cpp:187-193, 61-80, 97-123, 142-166 (setGaitParam :23-41)
//   f_ff = -seResult.rBody * f            /root/reference/src/MPC_Ctrl/ConvexMPCLocomotion.cpp:672-685
//   ori::quaternionToRotationMatrix       /root/reference/src/Utilities/orientation_tools.h:170-188
//   computeLegJacobianAndPosition         /root/reference/src/Controllers/LegController.cpp:204-240
//   LegController::updateData/updateCommand  /root/reference/src/Controllers/LegController.cpp:89-155
// Parity pin: unpinned by the reference (it has no tests, SURVEY section 4, and needs Eigen to compile); the
// product's device code (csrc/mpc_legs.h) is checked against this file bit for bit.
// Built with -ffp-contract=off: the reference's build has no FMA.
#include <cmath>
#include <cstdint>
#include <cstring>

#include "../include/mpc_batch.h"

namespace {

struct V3 { float v[3]; };
struct M3 { float m[3][3]; };  // m[row][col]

// Eigen, fixed size 3: (lhs.row(i).transpose().cwiseProduct(rhs)).sum() with the reduction unrolled by halving,
// i.e. p0 + (p1 + p2)
inline float sum3(float p0, float p1, float p2) { return p0 + (p1 + p2); }
inline V3 mul(const M3& A, const V3& x) {
  V3 y;
  for (int i = 0; i < 3; i++) y.v[i] = sum3(A.m[i][0] * x.v[0], A.m[i][1] * x.v[1], A.m[i][2] * x.v[2]);
  return y;
}
inline M3 transpose(const M3& A) {
  M3 T;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) T.m[i][j] = A.m[j][i];
  return T;
}
inline M3 neg(const M3& A) {
  M3 T;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) T.m[i][j] = -A.m[i][j];
  return T;
}

}  // namespace

extern "C" {

// gait [B][MPC_GAIT_WORDS] int32 -> state [B][MPC_GAIT_STATE_WORDS], tables (optional) [B][table_stride] bytes
void oracle_gait_state(const int32_t* gait, int batch, float* state, unsigned char* tables, int table_stride) {
  for (int b = 0; b < batch; b++) {
    const int32_t* g = gait + (size_t)b * MPC_GAIT_WORDS;
    const int iterationsPerMPC = g[MPC_GAIT_ITERATIONS_PER_MPC], currentIteration = g[MPC_GAIT_CURRENT_ITERATION];
    const int _nIterations = g[MPC_GAIT_SEGMENTS];
    float _offsetsFloat[4], _durationsFloat[4];
    for (int i = 0; i < 4; i++) {  // setGaitParam: offsets.cast<float>() / (float) nSegment
      _offsetsFloat[i] = (float)g[MPC_GAIT_OFFSETS + i] / (float)_nIterations;
      _durationsFloat[i] = (float)g[MPC_GAIT_DURATIONS + i] / (float)_nIterations;
    }
    // setIterations
    const int _iteration = (currentIteration / iterationsPerMPC) % _nIterations;
    const float _phase = (float)(currentIteration % (iterationsPerMPC * _nIterations)) /
                         (float)(iterationsPerMPC * _nIterations);
    float* out = state + (size_t)b * MPC_GAIT_STATE_WORDS;
    memcpy(out + MPC_GAIT_STATE_ITERATION, &_iteration, 4);
    out[MPC_GAIT_STATE_PHASE] = _phase;
    {  // getContactState
      float progress[4];
      for (int i = 0; i < 4; i++) progress[i] = _phase - _offsetsFloat[i];
      for (int i = 0; i < 4; i++) {
        if (progress[i] < 0) progress[i] += 1.;
        if (progress[i] > _durationsFloat[i]) {
          progress[i] = 0.;
        } else {
          progress[i] = progress[i] / _durationsFloat[i];
        }
      }
      for (int i = 0; i < 4; i++) out[MPC_GAIT_STATE_CONTACT + i] = progress[i];
    }
    {  // getSwingState
      float swing_offset[4], swing_duration[4], progress[4];
      for (int i = 0; i < 4; i++) swing_offset[i] = _offsetsFloat[i] + _durationsFloat[i];
      for (int i = 0; i < 4; i++)
        if (swing_offset[i] > 1) swing_offset[i] -= 1.;
      for (int i = 0; i < 4; i++) swing_duration[i] = 1.f - _durationsFloat[i];  // Eigen casts the scalar to float
      for (int i = 0; i < 4; i++) progress[i] = _phase - swing_offset[i];
      for (int i = 0; i < 4; i++) {
        if (progress[i] < 0) progress[i] += 1.f;
        if (progress[i] > swing_duration[i]) {
          progress[i] = 0.;
        } else {
          if (swing_duration[i] < 0.0000000001) {
            progress[i] = 0.0;
          } else {
            progress[i] = progress[i] / swing_duration[i];
          }
        }
      }
      for (int i = 0; i < 4; i++) out[MPC_GAIT_STATE_SWING + i] = progress[i];
    }
    if (tables) {  // getMpcTable
      unsigned char* t = tables + (size_t)b * table_stride;
      for (int i = 0; i < _nIterations; i++) {
        int iter = (i + _iteration + 1) % _nIterations;
        for (int j = 0; j < 4; j++) {
          int progress = iter - g[MPC_GAIT_OFFSETS + j];
          if (progress < 0) progress += _nIterations;
          t[i * 4 + j] = progress < g[MPC_GAIT_DURATIONS + j] ? 1 : 0;
        }
      }
    }
  }
}

// legs [B][MPC_LEG_WORDS], forces [B][12] -> f_ff [B][12], tau [B][12]
void oracle_leg_commands(const float* legs, const float* forces, int batch, float* f_ff_out, float* tau_out) {
  for (int b = 0; b < batch; b++) {
    const float* L = legs + (size_t)b * MPC_LEG_WORDS;
    int32_t use_ff[4];
    memcpy(use_ff, L + MPC_LEG_USE_FF, 16);
    // quaternionToRotationMatrix
    const float e0 = L[MPC_LEG_Q], e1 = L[MPC_LEG_Q + 1], e2 = L[MPC_LEG_Q + 2], e3 = L[MPC_LEG_Q + 3];
    M3 R;
    R.m[0][0] = 1 - 2 * (e2 * e2 + e3 * e3); R.m[0][1] = 2 * (e1 * e2 - e0 * e3); R.m[0][2] = 2 * (e1 * e3 + e0 * e2);
    R.m[1][0] = 2 * (e1 * e2 + e0 * e3); R.m[1][1] = 1 - 2 * (e1 * e1 + e3 * e3); R.m[1][2] = 2 * (e2 * e3 - e0 * e1);
    R.m[2][0] = 2 * (e1 * e3 - e0 * e2); R.m[2][1] = 2 * (e2 * e3 + e0 * e1); R.m[2][2] = 1 - 2 * (e1 * e1 + e2 * e2);
    const M3 rBody = transpose(R);  // R.transposeInPlace()
    const float l1 = L[MPC_LEG_LINKS], l2 = L[MPC_LEG_LINKS + 1], l3 = L[MPC_LEG_LINKS + 2], l4 = L[MPC_LEG_LINKS + 3];
    const float kpJoint = L[MPC_LEG_JOINT_GAINS], kdJoint = L[MPC_LEG_JOINT_GAINS + 1];
    for (int leg = 0; leg < 4; leg++) {
      V3 f = {{forces[(size_t)12 * b + 3 * leg], forces[(size_t)12 * b + 3 * leg + 1], forces[(size_t)12 * b + 3 * leg + 2]}};
      V3 forceFeedForward = {{0, 0, 0}};
      if (use_ff[leg]) forceFeedForward = mul(neg(rBody), f);  // f_ff[leg] = -seResult.rBody * f
      for (int i = 0; i < 3; i++) f_ff_out[(size_t)12 * b + 3 * leg + i] = forceFeedForward.v[i];
      // computeLegJacobianAndPosition
      const float sideSigns[4] = {-1, 1, -1, 1};
      const float sideSign = sideSigns[leg];
      const float* q = L + MPC_LEG_JOINT_Q + 3 * leg;
      const float* qd = L + MPC_LEG_JOINT_QD + 3 * leg;
      const float s1 = (float)std::sin((double)q[0]), s2 = (float)std::sin((double)q[1]), s3 = (float)std::sin((double)q[2]);
      const float c1 = (float)std::cos((double)q[0]), c2 = (float)std::cos((double)q[1]), c3 = (float)std::cos((double)q[2]);
      const float c23 = c2 * c3 - s2 * s3;
      const float s23 = s2 * c3 + c2 * s3;
      M3 J;
      J.m[0][0] = 0;
      J.m[0][1] = l3 * c23 + l2 * c2;
      J.m[0][2] = l3 * c23;
      J.m[1][0] = l3 * c1 * c23 + l2 * c1 * c2 - (l1 + l4) * sideSign * s1;
      J.m[1][1] = -l3 * s1 * s23 - l2 * s1 * s2;
      J.m[1][2] = -l3 * s1 * s23;
      J.m[2][0] = l3 * s1 * c23 + l2 * c2 * s1 + (l1 + l4) * sideSign * c1;
      J.m[2][1] = l3 * c1 * s23 + l2 * c1 * s2;
      J.m[2][2] = l3 * c1 * s23;
      V3 p;
      p.v[0] = l3 * s23 + l2 * s2;
      p.v[1] = (l1 + l4) * sideSign * c1 + l3 * (s1 * c23) + l2 * c2 * s1;
      p.v[2] = (l1 + l4) * sideSign * s1 - l3 * (c1 * c23) - l2 * c1 * c2;
      V3 qdv = {{qd[0], qd[1], qd[2]}};
      const V3 v = mul(J, qdv);  // datas[leg].v = datas[leg].J * datas[leg].qd
      // updateCommand
      V3 legTorque = {{L[MPC_LEG_TAU_FF + 3 * leg], L[MPC_LEG_TAU_FF + 3 * leg + 1], L[MPC_LEG_TAU_FF + 3 * leg + 2]}};
      V3 footForce = forceFeedForward;
      M3 kp = {{{0}}}, kd = {{{0}}};
      for (int i = 0; i < 3; i++) {
        kp.m[i][i] = L[MPC_LEG_KP + 3 * leg + i];
        kd.m[i][i] = L[MPC_LEG_KD + 3 * leg + i];
      }
      V3 dp, dv;
      for (int i = 0; i < 3; i++) {
        dp.v[i] = L[MPC_LEG_PDES + 3 * leg + i] - p.v[i];
        dv.v[i] = L[MPC_LEG_VDES + 3 * leg + i] - v.v[i];
      }
      const V3 a = mul(kp, dp), c = mul(kd, dv);
      for (int i = 0; i < 3; i++) footForce.v[i] += a.v[i];
      for (int i = 0; i < 3; i++) footForce.v[i] += c.v[i];
      const V3 jt = mul(transpose(J), footForce);
      for (int i = 0; i < 3; i++) legTorque.v[i] += jt.v[i];
      for (int j = 0; j < 3; j++) {
        // crtlParam(2) * (0.0 - q) - crtlParam(3) * qd + legTorque: double arithmetic except the second product
        const double t = kpJoint * (0.0 - q[j]) - kdJoint * qd[j] + legTorque.v[j];
        tau_out[(size_t)12 * b + 3 * leg + j] = (float)t;
      }
    }
  }
}

}  // extern "C"
