"""Packed problem records (host side), layout defined in include/mpc_batch.h.

One record = one robot x horizon instance = the contents of the reference's
``update_data_t`` + ``problem_setup`` (convexMPC_interface.h:13-38) that solve_mpc
reads, plus the body inertia / mass the reference hard-codes (RobotState.cpp:38-40,
RobotState.h:23).  Pure numpy; no device code here.
"""
import numpy as np

REC_P, REC_V, REC_Q, REC_W, REC_R = 0, 3, 6, 10, 13
REC_YAW, REC_XDRAG, REC_ALPHA, REC_WEIGHTS = 25, 26, 27, 28
REC_IBODY, REC_MASS, REC_DT, REC_MU, REC_FMAX, REC_RESERVED, REC_TRAJ = 40, 43, 44, 45, 46, 47, 48
MAX_HORIZON = 36

# constants the reference's caller passes (ConvexMPCLocomotion.cpp:598,604,630; GaitCtrller.cpp:6)
DEFAULT_WEIGHTS = np.array([2.5, 2.5, 10, 50, 50, 100, 0, 0, 0.5, 0.2, 0.2, 0.1], np.float32)
DEFAULT_ALPHA = np.float32(4e-5)
DEFAULT_MU = np.float32(0.4)
DEFAULT_FMAX = np.float32(120.0)
DEFAULT_DT = np.float32(0.002 * 13)
DEFAULT_IBODY = np.array([0.07, 0.26, 0.242], np.float32)
DEFAULT_MASS = np.float32(9.0)


def record_stride(horizon):
    return (4 * (REC_TRAJ + 12 * horizon) + 4 * horizon + 15) // 16 * 16


def gait_offset(horizon):
    return 4 * (REC_TRAJ + 12 * horizon)


def pack_records(horizon, p, v, q, w, r, yaw, traj, gait, weights=None, alpha=None, x_drag=None,
                 I_body=None, mass=None, dt=None, mu=None, f_max=None):
    """Builds a uint8 [B, stride] array of records from per-field arrays.

    p,v,w [B,3]; q [B,4] (w,x,y,z); r [B,12] as r[axis*4+leg]; yaw [B]; traj [B,12h];
    gait [B,4h] 0/1 as gait[step*4+leg].  Scalars / None broadcast the reference defaults.
    """
    p = np.asarray(p, np.float32)
    B = p.shape[0]
    h = horizon
    if not 1 <= h <= MAX_HORIZON:
        raise ValueError("horizon must be in 1..%d" % MAX_HORIZON)
    rec = np.zeros((B, record_stride(h)), np.uint8)
    f = rec.view(np.float32)

    def put(off, val, n, default=None):
        if val is None:
            val = default
        f[:, off:off + n] = np.broadcast_to(np.asarray(val, np.float32).reshape(-1, n) if np.ndim(val) else
                                            np.float32(val), (B, n))

    put(REC_P, p, 3)
    put(REC_V, v, 3)
    put(REC_Q, q, 4)
    put(REC_W, w, 3)
    put(REC_R, r, 12)
    put(REC_YAW, np.asarray(yaw, np.float32).reshape(-1, 1), 1)
    put(REC_XDRAG, None if x_drag is None else np.asarray(x_drag, np.float32).reshape(-1, 1), 1, 0.0)
    put(REC_ALPHA, None if alpha is None else np.asarray(alpha, np.float32).reshape(-1, 1), 1, DEFAULT_ALPHA)
    put(REC_WEIGHTS, weights, 12, DEFAULT_WEIGHTS)
    put(REC_IBODY, I_body, 3, DEFAULT_IBODY)
    put(REC_MASS, None if mass is None else np.asarray(mass, np.float32).reshape(-1, 1), 1, DEFAULT_MASS)
    put(REC_DT, None if dt is None else np.asarray(dt, np.float32).reshape(-1, 1), 1, DEFAULT_DT)
    put(REC_MU, None if mu is None else np.asarray(mu, np.float32).reshape(-1, 1), 1, DEFAULT_MU)
    put(REC_FMAX, None if f_max is None else np.asarray(f_max, np.float32).reshape(-1, 1), 1, DEFAULT_FMAX)
    put(REC_TRAJ, np.asarray(traj, np.float32).reshape(B, 12 * h), 12 * h)
    go = gait_offset(h)
    rec[:, go:go + 4 * h] = np.asarray(gait).reshape(B, 4 * h).astype(np.uint8)
    return rec


def unpack_records(rec, horizon):
    """Inverse of pack_records: dict of per-field arrays."""
    rec = np.ascontiguousarray(rec, np.uint8)
    f = rec.view(np.float32)
    h = horizon
    go = gait_offset(h)
    return dict(p=f[:, REC_P:REC_P + 3], v=f[:, REC_V:REC_V + 3], q=f[:, REC_Q:REC_Q + 4], w=f[:, REC_W:REC_W + 3],
                r=f[:, REC_R:REC_R + 12], yaw=f[:, REC_YAW], x_drag=f[:, REC_XDRAG], alpha=f[:, REC_ALPHA],
                weights=f[:, REC_WEIGHTS:REC_WEIGHTS + 12], I_body=f[:, REC_IBODY:REC_IBODY + 3],
                mass=f[:, REC_MASS], dt=f[:, REC_DT], mu=f[:, REC_MU], f_max=f[:, REC_FMAX],
                traj=f[:, REC_TRAJ:REC_TRAJ + 12 * h], gait=rec[:, go:go + 4 * h])


def algorithmic_bytes(horizon):
    """Compulsory HBM bytes per solve: inputs 4*(47+12h)+4h, output 48."""
    return 4 * (47 + 12 * horizon) + 4 * horizon + 48


def algorithmic_flops(horizon, nv):
    """fp64 operations per solve of the algorithm the engine runs, for the secondary on-chip
    figure of the bench line: symmetric sweep inversion nv^3 (nv^3/2 FMAs), closed-form assembly (9 FMAs per entry of
    the lower triangle + the six 12x12 tables + gradient moments), x = -H^{-1} g riding along (nv^2).  The active-set
    iterations are data dependent and left out.  The reference's dense route (SolverMPC.cpp:395) would be
    2*(12h)^2*(13h) = 3744 h^3 for the Hessian alone."""
    sweep = nv ** 3 + 2 * nv * nv
    assembly = 18 * (nv * (nv + 1) // 2) + 6 * 2 * 12 * 12 * 12 + 2 * 3 * 12 * horizon * horizon
    return sweep + assembly



def algorithmic_flops_riccati(horizon, nv):
    """fp64 operations per solve of the Riccati solver (csrc/mpc_riccati.h), counted on the algorithm (not on the padded
    8x8 tiles the tensor pipe executes), for n = nv / horizon controls per step: the factorisation
    (M = PB 288n, S = B'M 24n^2, sweep n^3, G = M'A and the sparse products with A ~ 72n + 1728, K = S^-1 G 24n^2,
    G'K 288n per step) plus ONE backward / forward sweep pair (x = -H^-1 g: 96n + 2n^2 + 100 per step).  The sweeps of the
    active-set iterations are data dependent and left out, like the iterations of the inverse-based solver."""
    n = nv / float(horizon)
    factor = n ** 3 + 48 * n * n + 648 * n + 1728
    sweep = 96 * n + 2 * n * n + 100
    return int(horizon * (factor + sweep))
