"""Tick records (host side): the compact per-robot input of one MPC tick (layout in include/mpc_batch.h).

A tick carries what ConvexMPCLocomotion::updateMPCIfNeeded / solveDenseMPC read at an MPC tick -- the state
estimate, the WORLD foot positions, the command state and the gait definition -- and the engine builds the
problem record (reference trajectory, COM-relative feet, contact table) from it on the device.  Pure numpy here.
"""
import numpy as np

from . import records as R

TICK_P, TICK_V, TICK_Q, TICK_W, TICK_PFOOT = 0, 3, 6, 10, 13
TICK_YAW, TICK_XDRAG, TICK_ALPHA, TICK_WEIGHTS = 25, 26, 27, 28
TICK_IBODY, TICK_MASS, TICK_DT, TICK_MU, TICK_FMAX, TICK_HEIGHT = 40, 43, 44, 45, 46, 47
TICK_RPY_COMP, TICK_YAW_DES, TICK_POS_DES, TICK_YAW_RATE, TICK_VDES = 48, 50, 51, 53, 54
TICK_STANDING, TICK_ITERATION, TICK_OFFSETS, TICK_DURATIONS, TICK_WORDS = 56, 57, 58, 62, 68
TICK_STRIDE = 4 * TICK_WORDS


def pack_ticks(p, v, q, w, p_foot, yaw, pos_des, yaw_des, yaw_rate, v_des_world, offsets, durations, iteration,
               standing=None, rpy_comp=None, body_height=0.25, x_drag=None, alpha=None, weights=None, I_body=None,
               mass=None, dt=None, mu=None, f_max=None):
    """Builds a float32 [B, 68] array of tick records (the int fields are stored bit-exactly in their slots).

    p, v, w [B,3]; q [B,4] (w,x,y,z); p_foot [B,4,3] world foot positions; yaw [B]; pos_des [B,2]
    (world_position_desired, or stand_traj[0:2] when standing); yaw_des [B]; yaw_rate [B]; v_des_world [B,2];
    offsets, durations [B,4] or [4] ints; iteration [B] ints; standing [B] bool."""
    p = np.asarray(p, np.float32)
    B = p.shape[0]
    t = np.zeros((B, TICK_WORDS), np.float32)
    ti = t.view(np.int32)

    def put(off, val, n, default=None):
        if val is None:
            val = default
        t[:, off:off + n] = np.broadcast_to(np.asarray(val, np.float32).reshape(-1, n) if np.ndim(val) else
                                            np.float32(val), (B, n))

    col = lambda a: None if a is None else np.asarray(a, np.float32).reshape(-1, 1)  # noqa: E731
    put(TICK_P, p, 3)
    put(TICK_V, v, 3)
    put(TICK_Q, q, 4)
    put(TICK_W, w, 3)
    put(TICK_PFOOT, np.asarray(p_foot, np.float32).reshape(B, 12), 12)
    put(TICK_YAW, col(yaw), 1)
    put(TICK_XDRAG, col(x_drag), 1, 0.0)
    put(TICK_ALPHA, col(alpha), 1, R.DEFAULT_ALPHA)
    put(TICK_WEIGHTS, weights, 12, R.DEFAULT_WEIGHTS)
    put(TICK_IBODY, I_body, 3, R.DEFAULT_IBODY)
    put(TICK_MASS, col(mass), 1, R.DEFAULT_MASS)
    put(TICK_DT, col(dt), 1, R.DEFAULT_DT)
    put(TICK_MU, col(mu), 1, R.DEFAULT_MU)
    put(TICK_FMAX, col(f_max), 1, R.DEFAULT_FMAX)
    put(TICK_HEIGHT, col(body_height) if np.ndim(body_height) else body_height, 1)
    put(TICK_RPY_COMP, rpy_comp, 2, np.zeros(2, np.float32))
    put(TICK_YAW_DES, col(yaw_des), 1)
    put(TICK_POS_DES, np.asarray(pos_des, np.float32).reshape(B, 2), 2)
    put(TICK_YAW_RATE, col(yaw_rate), 1)
    put(TICK_VDES, np.asarray(v_des_world, np.float32).reshape(B, 2), 2)
    ti[:, TICK_STANDING] = 0 if standing is None else np.asarray(standing, np.int32)
    ti[:, TICK_ITERATION] = np.asarray(iteration, np.int32)
    ti[:, TICK_OFFSETS:TICK_OFFSETS + 4] = np.broadcast_to(np.asarray(offsets, np.int32), (B, 4))
    ti[:, TICK_DURATIONS:TICK_DURATIONS + 4] = np.broadcast_to(np.asarray(durations, np.int32), (B, 4))
    return t


def synth_ticks(batch, horizon, seed=1234, mixed_gaits=False):
    """Seeded synthetic ticks in the spirit of workloads.config2 / config3: perturbed trotting (or mixed-gait)
    robots with WORLD foot positions, a command velocity along the body yaw and a position target that is
    sometimes further than 0.1 m away (to exercise the clamp of ConvexMPCLocomotion.cpp:536-543)."""
    from . import gait as G
    from . import workloads as W
    rng = np.random.default_rng(seed)
    h = horizon
    roll, pitch, yaw, p, v, w, feet, v_nom = W._states(rng, batch)
    q = W.rpy_to_quat(roll, pitch, yaw)
    p_foot = feet + p[:, None, :]
    pos_des = p[:, :2] + rng.normal(0, 0.08, (batch, 2))
    if mixed_gaits:
        ids = rng.integers(0, 12, batch)
        off = np.zeros((batch, 4), np.int32)
        dur = np.zeros((batch, 4), np.int32)
        standing = np.zeros(batch, bool)
        for b in range(batch):
            name = G.gait_by_number(ids[b])
            o, d = G.rescale(*G.GAITS_14[name], h)
            off[b], dur[b] = o, d
            standing[b] = name == "standing"
    else:
        off = np.broadcast_to(np.array([0, h // 2, h // 2, 0], np.int32), (batch, 4))
        dur = np.full((batch, 4), h // 2, np.int32)
        standing = np.zeros(batch, bool)
    x_drag = np.where(rng.random(batch) < 0.3, rng.normal(0, 0.2, batch), 0.0)
    return pack_ticks(p, v, q, w, p_foot, yaw, pos_des, yaw, rng.normal(0, 0.3, batch), v_nom[:, :2], off, dur,
                      rng.integers(0, h, batch), standing=standing, rpy_comp=rng.normal(0, 0.01, (batch, 2)),
                      body_height=0.25, x_drag=x_drag)
