"""SURVEY 8f rows N2 (gait state) and N4 (forces -> leg commands) on the CPU tier: the device bodies
(csrc/mpc_legs.h, host build) against the oracle's literal restatement (oracle/leg_oracle.cpp), bit for bit, and the
oracle against independent numpy statements of the same formulas."""
import numpy as np

from quadruped_ctrl_b200 import gait as G
from quadruped_ctrl_b200 import legs as LG

from common import emu_gait_state, emu_leg_commands


def _gait_records(seed, n=600):
    rng = np.random.default_rng(seed)
    recs = []
    for b in range(n):
        nseg = int(rng.choice([10, 14, 16, 20, 36, 1, 2]))
        name = list(G.GAITS_14)[int(rng.integers(0, len(G.GAITS_14)))]
        off, dur = G.rescale(*G.GAITS_14[name], nseg)
        recs.append(LG.pack_gait_records(int(rng.choice([13, 10, 27])), int(rng.integers(0, 100000)), nseg, off, dur))
    return np.concatenate(recs)


def test_gait_state_matches_oracle_bit_for_bit(oracle):
    g = _gait_records(3)
    so, to = oracle.gait_state(g, want_table=True)
    se, te = emu_gait_state(g, want_table=True)
    assert np.array_equal(so.view(np.int32), se.view(np.int32))
    assert np.array_equal(to, te)


def test_gait_state_oracle_against_numpy(oracle):
    """setIterations, the contact table and the stance / swing progress, restated with numpy."""
    g = _gait_records(4, 300)
    so, to = oracle.gait_state(g, want_table=True)
    for b in range(g.shape[0]):
        ipm, cur, n = int(g[b, 0]), int(g[b, 1]), int(g[b, 2])
        off, dur = g[b, 4:8], g[b, 8:12]
        it, ph = G.set_iterations(n, ipm, cur)
        assert so[b].view(np.int32)[0] == it
        assert so[b, 1] == np.float32(np.float32(cur % (ipm * n)) / np.float32(ipm * n))
        assert np.array_equal(to[b, :4 * n], G.mpc_table(n, off, dur, it).reshape(-1).astype(np.uint8))
        of, df = (off.astype(np.float32) / np.float32(n)), (dur.astype(np.float32) / np.float32(n))
        pr = np.float32(ph) - of
        pr = np.where(pr < 0, pr + np.float32(1), pr).astype(np.float32)
        with np.errstate(divide="ignore", invalid="ignore"):
            contact = np.where(pr > df, np.float32(0), pr / df).astype(np.float32)
        ok = df > 0
        assert np.array_equal(so[b, 2:6][ok], contact[ok])
        # (a zero-length stance gives 0/0 = NaN upstream as well: Gait.cpp:73)
        assert ((so[b, 2:6][ok] >= 0) & (so[b, 2:6][ok] <= 1)).all()
        assert ((so[b, 6:10][ok] >= 0) & (so[b, 6:10][ok] <= 1)).all()
        # a leg is never in stance and in swing at once (both progress values positive)
        assert not ((so[b, 2:6] > 0) & (so[b, 6:10] > 0))[ok].any()


def test_leg_commands_match_oracle_bit_for_bit(oracle):
    legs = LG.synth_leg_records(800, 11)
    forces = np.random.default_rng(12).normal(0, 30, (800, 12)).astype(np.float32)
    fo, to = oracle.leg_commands(legs, forces)
    fe, te = emu_leg_commands(legs, forces)
    assert np.array_equal(fo.view(np.int32), fe.view(np.int32))
    assert np.array_equal(to.view(np.int32), te.view(np.int32))


def test_leg_commands_oracle_against_numpy(oracle):
    """The same formulas in float64 numpy: rotation of the forces into the body frame, leg Jacobian, Cartesian PD,
    J' f, joint damping -- the fp32 oracle must sit within float rounding of it."""
    B = 300
    legs = LG.synth_leg_records(B, 21)
    forces = np.random.default_rng(22).normal(0, 30, (B, 12)).astype(np.float32)
    fo, to = oracle.leg_commands(legs, forces)
    L = legs.astype(np.float64)
    use = legs.view(np.int32)[:, LG.LEG_USE_FF:LG.LEG_USE_FF + 4]
    w, x, y, z = (L[:, i] for i in range(4))
    R = np.stack([np.stack([1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)], -1),
                  np.stack([2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)], -1),
                  np.stack([2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)], -1)], 1)
    rBody = np.transpose(R, (0, 2, 1))          # body <- world
    l1, l2, l3, l4 = LG.MINI_CHEETAH_LINKS.astype(np.float64)
    for leg in range(4):
        f = forces[:, 3 * leg:3 * leg + 3].astype(np.float64)
        ff = -np.einsum("bij,bj->bi", rBody, f) * use[:, leg:leg + 1]
        assert np.abs(fo[:, 3 * leg:3 * leg + 3] - ff).max() < 1e-4
        q = L[:, LG.LEG_JOINT_Q + 3 * leg:LG.LEG_JOINT_Q + 3 * leg + 3]
        qd = L[:, LG.LEG_JOINT_QD + 3 * leg:LG.LEG_JOINT_QD + 3 * leg + 3]
        s1, s2, s3, c1, c2, c3 = (np.sin(q[:, 0]), np.sin(q[:, 1]), np.sin(q[:, 2]), np.cos(q[:, 0]), np.cos(q[:, 1]),
                                  np.cos(q[:, 2]))
        c23, s23 = c2 * c3 - s2 * s3, s2 * c3 + c2 * s3
        sg = -1.0 if leg in (0, 2) else 1.0
        J = np.zeros((B, 3, 3))
        J[:, 0, 1], J[:, 0, 2] = l3 * c23 + l2 * c2, l3 * c23
        J[:, 1, 0] = l3 * c1 * c23 + l2 * c1 * c2 - (l1 + l4) * sg * s1
        J[:, 1, 1], J[:, 1, 2] = -l3 * s1 * s23 - l2 * s1 * s2, -l3 * s1 * s23
        J[:, 2, 0] = l3 * s1 * c23 + l2 * c2 * s1 + (l1 + l4) * sg * c1
        J[:, 2, 1], J[:, 2, 2] = l3 * c1 * s23 + l2 * c1 * s2, l3 * c1 * s23
        p = np.stack([l3 * s23 + l2 * s2, (l1 + l4) * sg * c1 + l3 * s1 * c23 + l2 * c2 * s1,
                      (l1 + l4) * sg * s1 - l3 * c1 * c23 - l2 * c1 * c2], -1)
        v = np.einsum("bij,bj->bi", J, qd)
        sl = slice(3 * leg, 3 * leg + 3)
        force = ff + L[:, LG.LEG_KP:LG.LEG_KP + 12][:, sl] * (L[:, LG.LEG_PDES:LG.LEG_PDES + 12][:, sl] - p) \
            + L[:, LG.LEG_KD:LG.LEG_KD + 12][:, sl] * (L[:, LG.LEG_VDES:LG.LEG_VDES + 12][:, sl] - v)
        tq = L[:, LG.LEG_TAU_FF:LG.LEG_TAU_FF + 12][:, sl] + np.einsum("bji,bj->bi", J, force)
        tau = L[:, LG.LEG_JOINT_GAINS:LG.LEG_JOINT_GAINS + 1] * (0.0 - q) - L[:, LG.LEG_JOINT_GAINS + 1:LG.LEG_JOINT_GAINS + 2] * qd + tq
        assert np.abs(to[:, sl] - tau).max() < 2e-3 * max(1.0, np.abs(tau).max())
