"""Seeded synthetic workloads for the five BASELINE.json configs.

The trajectory builder restates the reference's reference-trajectory rollout
(/root/reference/src/MPC_Ctrl/ConvexMPCLocomotion.cpp:514-577) and the r / weights /
alpha packing of solveDenseMPC (:592-621) for a whole batch at once with numpy.
Everything here is host-side input synthesis; nothing is read from /root/reference.
"""
import numpy as np

from . import gait as G
from . import records as R

BODY_HEIGHT = 0.25  # _body_height, ConvexMPCLocomotion.cpp:79
# nominal stance feet relative to the COM in the yaw frame, legs FR, FL, RR, RL
# (hip +-0.19 / +-0.049 + abad 0.062: Dynamics/MiniCheetah.h:25-31,105)
NOMINAL_FEET = np.array([[0.19, -0.111, -0.29], [0.19, 0.111, -0.29], [-0.19, -0.111, -0.29], [-0.19, 0.111, -0.29]])


def rpy_to_quat(roll, pitch, yaw):
    """ZYX Euler -> (w,x,y,z), the inverse of the reference's quat_to_rpy (SolverMPC.cpp:257-267)."""
    cr, sr = np.cos(roll / 2), np.sin(roll / 2)
    cp, sp = np.cos(pitch / 2), np.sin(pitch / 2)
    cy, sy = np.cos(yaw / 2), np.sin(yaw / 2)
    return np.stack([cr * cp * cy + sr * sp * sy, sr * cp * cy - cr * sp * sy,
                     cr * sp * cy + sr * cp * sy, cr * cp * sy - sr * sp * cy], -1)


def build_trajectory(horizon, dt_mpc, rpy_comp, yaw_des, x_start, y_start, yaw_rate, v_des_world,
                     body_height=BODY_HEIGHT):
    """trajAll [B,12h] for a moving gait (ConvexMPCLocomotion.cpp:547-576)."""
    B = np.shape(yaw_des)[0]
    t = np.zeros((B, horizon, 12), np.float32)
    t[:, :, 0] = np.asarray(rpy_comp[0], np.float32).reshape(-1, 1)
    t[:, :, 1] = np.asarray(rpy_comp[1], np.float32).reshape(-1, 1)
    t[:, :, 5] = body_height
    t[:, :, 8] = np.asarray(yaw_rate, np.float32).reshape(-1, 1)
    t[:, :, 9] = np.asarray(v_des_world[0], np.float32).reshape(-1, 1)
    t[:, :, 10] = np.asarray(v_des_world[1], np.float32).reshape(-1, 1)
    dt = np.float32(dt_mpc)
    x = np.asarray(x_start, np.float32).copy()
    y = np.asarray(y_start, np.float32).copy()
    yw = np.asarray(yaw_des, np.float32).copy()
    vx = np.asarray(v_des_world[0], np.float32)
    vy = np.asarray(v_des_world[1], np.float32)
    yr = np.asarray(yaw_rate, np.float32)
    for i in range(horizon):  # float32 running sums, as the reference accumulates them
        if i > 0:
            x = (x + dt * vx).astype(np.float32)
            y = (y + dt * vy).astype(np.float32)
            yw = (yw + dt * yr).astype(np.float32)
        t[:, i, 3], t[:, i, 4], t[:, i, 2] = x, y, yw
    return t.reshape(B, 12 * horizon)


def build_stand_trajectory(horizon, roll_des, pitch_des, yaw, x, y, body_height=BODY_HEIGHT):
    """trajAll for current_gait == 4 (standing), ConvexMPCLocomotion.cpp:515-531."""
    B = np.shape(yaw)[0]
    t = np.zeros((B, horizon, 12), np.float32)
    t[:, :, 0] = np.asarray(roll_des, np.float32).reshape(-1, 1)
    t[:, :, 1] = np.asarray(pitch_des, np.float32).reshape(-1, 1)
    t[:, :, 2] = np.asarray(yaw, np.float32).reshape(-1, 1)
    t[:, :, 3] = np.asarray(x, np.float32).reshape(-1, 1)
    t[:, :, 4] = np.asarray(y, np.float32).reshape(-1, 1)
    t[:, :, 5] = body_height
    return t.reshape(B, 12 * horizon)


def _states(rng, B, sigma_scale=1.0, v_nominal=0.5):
    """Perturbed robot states around the config-1 nominal."""
    s = sigma_scale
    roll = rng.normal(0, 0.05 * s, B)
    pitch = rng.normal(0, 0.05 * s, B)
    yaw = rng.normal(0, 0.3 * s, B)
    p = np.stack([rng.normal(0, 0.1 * s, B), rng.normal(0, 0.1 * s, B), 0.29 + rng.normal(0, 0.01 * s, B)], -1)
    cy, sy = np.cos(yaw), np.sin(yaw)
    v_nom = np.stack([v_nominal * cy, v_nominal * sy, np.zeros(B)], -1)
    v = v_nom + np.stack([rng.normal(0, 0.1 * s, B), rng.normal(0, 0.05 * s, B), rng.normal(0, 0.02 * s, B)], -1)
    w = rng.normal(0, 0.1 * s, (B, 3))
    feet = np.broadcast_to(NOMINAL_FEET, (B, 4, 3)).copy()
    # rotate the nominal stance into the world by yaw, then perturb
    fx = cy[:, None] * feet[:, :, 0] - sy[:, None] * feet[:, :, 1]
    fy = sy[:, None] * feet[:, :, 0] + cy[:, None] * feet[:, :, 1]
    feet = np.stack([fx, fy, feet[:, :, 2]], -1)
    feet = feet + np.stack([rng.normal(0, 0.03 * s, (B, 4)), rng.normal(0, 0.02 * s, (B, 4)),
                            rng.normal(0, 0.005 * s, (B, 4))], -1)
    return roll, pitch, yaw, p, v, w, feet, v_nom


def _pack(horizon, roll, pitch, yaw, p, v, w, feet, traj, gait, **kw):
    q = rpy_to_quat(roll, pitch, yaw)
    # r[axis*4+leg] = pFoot[leg][axis] - position[axis] (ConvexMPCLocomotion.cpp:611-613); feet are already COM-relative
    r = np.transpose(feet, (0, 2, 1)).reshape(len(yaw), 12)
    return R.pack_records(horizon, p, v, q, w, r, yaw, traj, gait, **kw)


def config1(horizon=10):
    """1 robot, trot, h=10, nominal state, all 10 gait phases -> [10, stride] records."""
    h = horizon
    B = h
    z = np.zeros(B)
    p = np.tile(np.array([0, 0, 0.29]), (B, 1))
    v = np.tile(np.array([0.5, 0, 0]), (B, 1))
    w = np.zeros((B, 3))
    feet = np.broadcast_to(NOMINAL_FEET, (B, 4, 3)).copy()
    traj = build_trajectory(h, R.DEFAULT_DT, (z, z), z, p[:, 0], p[:, 1], z, (v[:, 0], v[:, 1]))
    gait = G.mpc_tables(h, (0, h // 2, h // 2, 0), (h // 2,) * 4, np.arange(B))
    return _pack(h, z, z, z, p, v, w, feet, traj, gait)


def config2(batch=4096, horizon=10, seed=1234):
    """batch independent robots, trot, h=10, Gaussian state/foothold perturbations."""
    rng = np.random.default_rng(seed)
    h = horizon
    roll, pitch, yaw, p, v, w, feet, v_nom = _states(rng, batch)
    z = np.zeros(batch)
    traj = build_trajectory(h, R.DEFAULT_DT, (z, z), yaw, p[:, 0], p[:, 1], z, (v_nom[:, 0], v_nom[:, 1]))
    phase = rng.integers(0, h, batch)
    gait = G.mpc_tables(h, (0, h // 2, h // 2, 0), (h // 2,) * 4, phase)
    return _pack(h, roll, pitch, yaw, p, v, w, feet, traj, gait)


def config3(batch=4096, horizon=20, seed=2345):
    """h=20, gait ids sampled from 0..11 (undefined ids -> trot), randomised inertia and mass."""
    rng = np.random.default_rng(seed)
    h = horizon
    roll, pitch, yaw, p, v, w, feet, v_nom = _states(rng, batch)
    ids = rng.integers(0, 12, batch)
    off = np.zeros((batch, 4), np.int64)
    dur = np.zeros((batch, 4), np.int64)
    standing = np.zeros(batch, bool)
    for b in range(batch):
        name = G.gait_by_number(ids[b])
        o, d = G.rescale(*G.GAITS_14[name], h)
        off[b], dur[b] = o, d
        standing[b] = name == "standing"
    phase = rng.integers(0, h, batch)
    gait = G.mpc_tables(h, off, dur, phase)
    z = np.zeros(batch)
    traj = build_trajectory(h, R.DEFAULT_DT, (z, z), yaw, p[:, 0], p[:, 1], z, (v_nom[:, 0], v_nom[:, 1]))
    stand_traj = build_stand_trajectory(h, z, z, yaw, p[:, 0], p[:, 1])
    traj[standing] = stand_traj[standing]
    I_body = R.DEFAULT_IBODY[None, :] * rng.uniform(0.7, 1.3, (batch, 3))
    mass = R.DEFAULT_MASS * rng.uniform(0.8, 1.2, batch)
    return _pack(h, roll, pitch, yaw, p, v, w, feet, traj, gait, I_body=I_body, mass=mass)


def config4(batch=65536, horizon=10, seed=3456):
    """config 2 plus stairs foothold perturbations: per-foot z from the stair box heights and x offsets."""
    rng = np.random.default_rng(seed)
    h = horizon
    roll, pitch, yaw, p, v, w, feet, v_nom = _states(rng, batch)
    feet[:, :, 2] += rng.choice(np.array([0.0, 0.02, 0.04, 0.06, 0.08]), (batch, 4))
    feet[:, :, 0] += rng.uniform(-0.05, 0.05, (batch, 4))
    z = np.zeros(batch)
    traj = build_trajectory(h, R.DEFAULT_DT, (z, z), yaw, p[:, 0], p[:, 1], z, (v_nom[:, 0], v_nom[:, 1]))
    phase = rng.integers(0, h, batch)
    gait = G.mpc_tables(h, (0, h // 2, h // 2, 0), (h // 2,) * 4, phase)
    return _pack(h, roll, pitch, yaw, p, v, w, feet, traj, gait)


def config5(batch=65536, horizon=16, seed=4567):
    """h=16 galloping: offsets (0,4,7,11)/durations 7 of 14 rescaled to 16 segments."""
    rng = np.random.default_rng(seed)
    h = horizon
    roll, pitch, yaw, p, v, w, feet, v_nom = _states(rng, batch, v_nominal=1.5)
    off, dur = G.rescale(*G.GAITS_14["galloping"], h)
    z = np.zeros(batch)
    traj = build_trajectory(h, R.DEFAULT_DT, (z, z), yaw, p[:, 0], p[:, 1], z, (v_nom[:, 0], v_nom[:, 1]))
    phase = rng.integers(0, h, batch)
    gait = G.mpc_tables(h, off, dur, phase)
    return _pack(h, roll, pitch, yaw, p, v, w, feet, traj, gait)


def four_stance(batch=256, horizon=10, seed=777):
    """Standing gait (every leg in stance over the horizon): nv = 12h, the largest reduced QP."""
    rng = np.random.default_rng(seed)
    h = horizon
    roll, pitch, yaw, p, v, w, feet, v_nom = _states(rng, batch, v_nominal=0.0)
    z = np.zeros(batch)
    traj = build_stand_trajectory(h, z, z, yaw, p[:, 0], p[:, 1])
    gait = np.ones((batch, 4 * h), np.int32)
    return _pack(h, roll, pitch, yaw, p, v, w, feet, traj, gait)


def _trot_ticks(batch, horizon, seed, stairs):
    """config2 / config4 as TICK records (include/mpc_batch.h): the same seeded states, footholds, command and gait
    phase, with the feet given in the WORLD frame -- the records the engine builds from them on the device equal
    config2(...) / config4(...) up to the fp32 rounding of (foot + p) - p, gait tables byte for byte."""
    from . import ticks as T
    rng = np.random.default_rng(seed)
    h = horizon
    roll, pitch, yaw, p, v, w, feet, v_nom = _states(rng, batch)
    if stairs:
        feet[:, :, 2] += rng.choice(np.array([0.0, 0.02, 0.04, 0.06, 0.08]), (batch, 4))
        feet[:, :, 0] += rng.uniform(-0.05, 0.05, (batch, 4))
    phase = rng.integers(0, h, batch)
    q = rpy_to_quat(roll, pitch, yaw)
    return T.pack_ticks(p, v, q, w, feet + p[:, None, :], yaw, p[:, :2], yaw, np.zeros(batch), v_nom[:, :2],
                        (0, h // 2, h // 2, 0), (h // 2,) * 4, phase, body_height=BODY_HEIGHT)


def config2_ticks(batch=4096, horizon=10, seed=1234):
    return _trot_ticks(batch, horizon, seed, False)


def config4_ticks(batch=65536, horizon=10, seed=3456):
    return _trot_ticks(batch, horizon, seed, True)


CONFIGS = {"config1": config1, "config2": config2, "config3": config3, "config4": config4, "config5": config5,
           "four_stance": four_stance}
HORIZONS = {"config1": 10, "config2": 10, "config3": 20, "config4": 10, "config5": 16, "four_stance": 10}
