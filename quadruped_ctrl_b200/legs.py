"""Gait records and leg records (host side): the inputs of the two device-side callers around the MPC solve that
SURVEY 8f lists as rows N2 and N4 (layouts in include/mpc_batch.h).  Pure numpy.

Gait record  -> OffsetDurationGait::setIterations / getContactState / getSwingState / getMpcTable (Gait.cpp).
Leg record   -> f_ff = -rBody f (ConvexMPCLocomotion.cpp:672-685) and LegController::updateCommand
                (LegController.cpp:114-155) for the four legs of one robot.
"""
import numpy as np

GAIT_WORDS, GAIT_STATE_WORDS, LEG_WORDS = 12, 10, 100
LEG_Q, LEG_JOINT_Q, LEG_JOINT_QD, LEG_PDES, LEG_VDES, LEG_KP, LEG_KD = 0, 4, 16, 28, 40, 52, 64
LEG_TAU_FF, LEG_JOINT_GAINS, LEG_USE_FF, LEG_LINKS = 76, 88, 90, 94
MINI_CHEETAH_LINKS = np.array([0.062, 0.209, 0.195, 0.004], np.float32)   # Dynamics/MiniCheetah.h:31-37


def pack_gait_records(iterations_per_mpc, current_iteration, n_iterations, offsets, durations):
    """int32 [B, 12]: iterationsPerMPC, currentIteration, nIterations, 0, offsets[4], durations[4]."""
    cur = np.asarray(current_iteration, np.int32).reshape(-1)
    B = cur.shape[0]
    g = np.zeros((B, GAIT_WORDS), np.int32)
    g[:, 0] = np.broadcast_to(np.asarray(iterations_per_mpc, np.int32), (B,))
    g[:, 1] = cur
    g[:, 2] = np.broadcast_to(np.asarray(n_iterations, np.int32), (B,))
    g[:, 4:8] = np.broadcast_to(np.asarray(offsets, np.int32), (B, 4))
    g[:, 8:12] = np.broadcast_to(np.asarray(durations, np.int32), (B, 4))
    return g


def pack_leg_records(quat, q, qd, p_des, v_des, kp, kd, tau_ff=None, joint_gains=(0.0, 0.0), use_ff=None,
                     links=MINI_CHEETAH_LINKS):
    """float32 [B, 100] leg records.  quat [B,4] (w,x,y,z); q, qd, p_des, v_des [B,4,3]; kp, kd [B,4,3] or [3]
    (diagonals of kpCartesian / kdCartesian); tau_ff [B,4,3]; joint_gains (crtlParam(2), crtlParam(3)); use_ff
    [B,4] 0/1 (stance legs take the MPC force as forceFeedForward)."""
    quat = np.asarray(quat, np.float32)
    B = quat.shape[0]
    r = np.zeros((B, LEG_WORDS), np.float32)
    r[:, LEG_Q:LEG_Q + 4] = quat
    for off, val in ((LEG_JOINT_Q, q), (LEG_JOINT_QD, qd), (LEG_PDES, p_des), (LEG_VDES, v_des)):
        r[:, off:off + 12] = np.asarray(val, np.float32).reshape(B, 12)
    for off, val in ((LEG_KP, kp), (LEG_KD, kd)):
        v = np.asarray(val, np.float32)
        r[:, off:off + 12] = np.broadcast_to(v if v.ndim == 3 else v.reshape(1, 1, 3), (B, 4, 3)).reshape(B, 12)
    if tau_ff is not None:
        r[:, LEG_TAU_FF:LEG_TAU_FF + 12] = np.asarray(tau_ff, np.float32).reshape(B, 12)
    r[:, LEG_JOINT_GAINS:LEG_JOINT_GAINS + 2] = np.broadcast_to(np.asarray(joint_gains, np.float32), (B, 2))
    r.view(np.int32)[:, LEG_USE_FF:LEG_USE_FF + 4] = 1 if use_ff is None else np.asarray(use_ff, np.int32).reshape(B, 4)
    r[:, LEG_LINKS:LEG_LINKS + 4] = np.broadcast_to(np.asarray(links, np.float32), (B, 4))
    return r


def synth_leg_records(batch, seed=7):
    """Seeded leg records around the Mini-Cheetah standing pose, with the controller's gains
    (ConvexMPCLocomotion.cpp:378-381, 448-463)."""
    from . import workloads as W
    rng = np.random.default_rng(seed)
    roll, pitch, yaw = rng.normal(0, 0.1, batch), rng.normal(0, 0.1, batch), rng.normal(0, 1.0, batch)
    quat = W.rpy_to_quat(roll, pitch, yaw)
    q = np.array([0.0, -0.8, 1.6]) + rng.normal(0, 0.2, (batch, 4, 3))
    qd = rng.normal(0, 1.0, (batch, 4, 3))
    p_des = np.array([0.0, 0.0, -0.29]) + rng.normal(0, 0.05, (batch, 4, 3))
    v_des = rng.normal(0, 0.3, (batch, 4, 3))
    stance = rng.random((batch, 4)) < 0.6
    kp = np.where(stance[..., None], 0.0, np.array([700.0, 700.0, 200.0]))      # stance: 0 * Kp_stance
    kd = np.where(stance[..., None], np.array([7.0, 7.0, 7.0]), np.array([10.0, 10.0, 10.0]))
    tau_ff = np.where(rng.random((batch, 4, 1)) < 0.2, rng.normal(0, 0.5, (batch, 4, 3)), 0.0)
    return pack_leg_records(quat, q, qd, p_des, v_des, kp, kd, tau_ff, joint_gains=(0.3, 0.05),
                            use_ff=stance.astype(np.int32))
