"""Closed-loop rollouts of a batch of robots on the MPC's own model (host side, numpy): the single-rigid-body dynamics
the reference linearises (SolverMPC.cpp:235-254) integrated over one MPC period with the first-step forces of every
solve, the gait advancing one segment per tick, swing feet re-placed under the hips at touch-down.  It produces the
sequence of related problems a controller sees tick after tick -- the setting of the warm start --
without any of the reference's host-side estimators or leg controllers.  Input synthesis only; no device code here.
"""
import numpy as np

from . import gait as G
from . import records as R
from . import workloads as W


class Rollout:
    def __init__(self, batch, horizon=10, gait="trotting", seed=0, v_cmd=0.5, mu=None, f_max=None):
        rng = np.random.default_rng(seed)
        self.B, self.h = batch, horizon
        self.off, self.dur = G.rescale(*G.GAITS_14[gait], horizon)
        roll, pitch, yaw, p, v, w, feet, _ = W._states(rng, batch, sigma_scale=0.5, v_nominal=v_cmd)
        self.rpy = np.stack([roll, pitch, yaw], -1)
        self.p, self.v, self.w = p.copy(), v.copy(), w.copy()
        self.feet = feet + p[:, None, :]            # world foot positions [B, 4, 3]
        self.feet[:, :, 2] = 0.0                    # on the ground plane
        self.v_cmd = v_cmd
        self.mu, self.f_max = mu, f_max   # None: the reference's 0.4 / 120 N (ConvexMPCLocomotion.cpp:630)
        self.yaw_cmd = yaw.copy()
        self.tick = rng.integers(0, horizon, batch)  # every robot at its own gait phase
        self.mass = np.full(batch, float(R.DEFAULT_MASS))
        self.I_body = np.tile(R.DEFAULT_IBODY.astype(np.float64), (batch, 1))

    def records(self):
        """The packed problem records of the current tick (what solveDenseMPC would hand over)."""
        B, h = self.B, self.h
        cy, sy = np.cos(self.yaw_cmd), np.sin(self.yaw_cmd)
        vd = (self.v_cmd * cy, self.v_cmd * sy)
        z = np.zeros(B)
        traj = W.build_trajectory(h, R.DEFAULT_DT, (z, z), self.yaw_cmd, self.p[:, 0], self.p[:, 1], z, vd)
        gait = G.mpc_tables(h, self.off, self.dur, self.tick % h)
        q = W.rpy_to_quat(self.rpy[:, 0], self.rpy[:, 1], self.rpy[:, 2])
        r = np.transpose(self.feet - self.p[:, None, :], (0, 2, 1)).reshape(B, 12)
        kw = {}
        if self.mu is not None:
            kw["mu"] = np.full(B, self.mu)
        if self.f_max is not None:
            kw["f_max"] = np.full(B, self.f_max)
        return R.pack_records(h, self.p, self.v, q, self.w, r, self.rpy[:, 2], traj, gait, I_body=self.I_body,
                              mass=self.mass, **kw)

    def advance(self, forces):
        """Integrates one MPC period with the first-step forces [B, 12] (world frame, force[leg*3+axis])."""
        dt = float(R.DEFAULT_DT)
        f = np.asarray(forces, np.float64).reshape(self.B, 4, 3)
        F = f.sum(1)
        acc = F / self.mass[:, None] + np.array([0.0, 0.0, -9.8])
        rr = self.feet - self.p[:, None, :]
        tau = np.cross(rr, f).sum(1)
        cy, sy = np.cos(self.rpy[:, 2]), np.sin(self.rpy[:, 2])
        Rz = np.zeros((self.B, 3, 3))
        Rz[:, 0, 0], Rz[:, 0, 1], Rz[:, 1, 0], Rz[:, 1, 1], Rz[:, 2, 2] = cy, -sy, sy, cy, 1.0
        Iw = np.einsum("bij,bj,bkj->bik", Rz, self.I_body, Rz)
        wd = np.linalg.solve(Iw, tau[..., None])[..., 0]
        self.p = self.p + self.v * dt + 0.5 * acc * dt * dt
        self.v = self.v + acc * dt
        # small-angle Euler-rate update, as the linearised model has it (rpy' = R_yaw' w)
        self.rpy = self.rpy + np.einsum("bji,bj->bi", Rz, self.w) * dt
        self.w = self.w + wd * dt
        # gait: next segment; legs that touch down now are placed under their hips, a half stance ahead
        before = G.mpc_tables(self.h, self.off, self.dur, self.tick % self.h)[:, :4]
        self.tick = self.tick + 1
        after = G.mpc_tables(self.h, self.off, self.dur, self.tick % self.h)[:, :4]
        land = (before == 0) & (after == 1)
        cy, sy = np.cos(self.rpy[:, 2]), np.sin(self.rpy[:, 2])
        hip = W.NOMINAL_FEET[None, :, :2]
        hx = cy[:, None] * hip[..., 0] - sy[:, None] * hip[..., 1] + self.p[:, None, 0]
        hy = sy[:, None] * hip[..., 0] + cy[:, None] * hip[..., 1] + self.p[:, None, 1]
        stance_t = 0.5 * dt * np.asarray(self.dur, np.float64)[None, :]
        tx, ty = hx + self.v[:, None, 0] * stance_t, hy + self.v[:, None, 1] * stance_t
        self.feet[..., 0] = np.where(land, tx, self.feet[..., 0])
        self.feet[..., 1] = np.where(land, ty, self.feet[..., 1])
