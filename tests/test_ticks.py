"""Device-side record builder (SURVEY 8f rows N1 + N2) on the CPU tier: the kernel body (csrc/mpc_ticks.h, host
build) against the oracle's literal restatement of the reference's host code (oracle/tick_oracle.cpp), byte for
byte, and the oracle against the independent numpy builders used for workload synthesis."""
import numpy as np

from quadruped_ctrl_b200 import gait as G
from quadruped_ctrl_b200 import records as R
from quadruped_ctrl_b200 import ticks as T
from quadruped_ctrl_b200 import workloads as W

from common import emu_build_records


def test_builder_matches_oracle_byte_for_byte(oracle):
    for h, mixed, seed in ((10, False, 1), (10, True, 2), (16, True, 3), (20, True, 4), (36, False, 5), (1, False, 6)):
        tk = T.synth_ticks(300, h, seed, mixed_gaits=mixed)
        rec_o, st_o = oracle.build_records(tk, h)
        rec_e, st_e = emu_build_records(tk, h)
        assert rec_o.shape == (300, R.record_stride(h))
        assert np.array_equal(rec_o, rec_e), (h, mixed)
        assert np.array_equal(st_o, st_e)


def test_oracle_builder_against_numpy_restatement(oracle):
    """The same tick built by the oracle and by the independent numpy pieces (gait.mpc_tables,
    workloads.build_trajectory / build_stand_trajectory, records.pack_records)."""
    h = 10
    tk = T.synth_ticks(400, h, 9, mixed_gaits=True)
    ti = tk.view(np.int32)
    rec_o, st_o = oracle.build_records(tk, h)
    u = R.unpack_records(rec_o, h)
    # contact tables
    tab = G.mpc_tables(h, ti[:, T.TICK_OFFSETS:T.TICK_OFFSETS + 4], ti[:, T.TICK_DURATIONS:T.TICK_DURATIONS + 4],
                       ti[:, T.TICK_ITERATION])
    assert np.array_equal(u["gait"], tab.astype(np.uint8))
    # COM-relative feet r[axis*4+leg]
    pf = tk[:, T.TICK_PFOOT:T.TICK_PFOOT + 12].reshape(-1, 4, 3)
    r = np.transpose(pf - tk[:, None, T.TICK_P:T.TICK_P + 3], (0, 2, 1)).reshape(-1, 12)
    assert np.array_equal(u["r"], r.astype(np.float32))
    # trajectories: clamp of the position target, running float sums / constant stand trajectory
    p = tk[:, T.TICK_P:T.TICK_P + 2]
    des = tk[:, T.TICK_POS_DES:T.TICK_POS_DES + 2].copy()
    start = des.copy()
    hi = (des - p) > np.float32(0.1)
    lo = (p - des) > np.float32(0.1)
    start[hi] = (p.astype(np.float64) + 0.1).astype(np.float32)[hi]
    start[lo] = (p.astype(np.float64) - 0.1).astype(np.float32)[lo]
    standing = ti[:, T.TICK_STANDING] != 0
    assert hi.any() and lo.any() and standing.any() and (~standing).any()
    rc = (tk[:, T.TICK_RPY_COMP], tk[:, T.TICK_RPY_COMP + 1])
    mov = W.build_trajectory(h, tk[:, T.TICK_DT], rc, tk[:, T.TICK_YAW_DES], start[:, 0], start[:, 1],
                             tk[:, T.TICK_YAW_RATE], (tk[:, T.TICK_VDES], tk[:, T.TICK_VDES + 1]))
    stand = W.build_stand_trajectory(h, rc[0], rc[1], tk[:, T.TICK_YAW_DES], des[:, 0], des[:, 1])
    want = np.where(standing[:, None], stand, mov)
    assert np.array_equal(u["traj"], want)
    assert np.array_equal(st_o[~standing, :2], start[~standing]) and np.array_equal(st_o[standing, :2], des[standing])
    # pass-through fields and the x_drag integrator
    for name, off, n in (("p", T.TICK_P, 3), ("v", T.TICK_V, 3), ("q", T.TICK_Q, 4), ("w", T.TICK_W, 3),
                         ("weights", T.TICK_WEIGHTS, 12), ("I_body", T.TICK_IBODY, 3)):
        assert np.array_equal(u[name], tk[:, off:off + n])
    assert np.array_equal(u["x_drag"], tk[:, T.TICK_XDRAG])
    vx = tk[:, T.TICK_V]
    moving = (vx > 0.3) | (vx < -0.3)
    pz_err = tk[:, T.TICK_P + 2] - tk[:, T.TICK_HEIGHT]
    nxt = tk[:, T.TICK_XDRAG] + (np.float32(3.0) * pz_err * tk[:, T.TICK_DT] / np.where(moving, vx, 1)).astype(np.float32)
    assert np.array_equal(st_o[moving, 2], nxt[moving]) and np.array_equal(st_o[~moving, 2], tk[~moving, T.TICK_XDRAG])


def test_ticks_reproduce_config1(oracle):
    """Config 1 of BASELINE.json through the tick route equals workloads.config1 byte for byte."""
    h = 10
    rec = W.config1(h)
    B = rec.shape[0]
    p = np.tile(np.array([0, 0, 0.29], np.float32), (B, 1))
    v = np.tile(np.array([0.5, 0, 0], np.float32), (B, 1))
    q = np.tile(np.array([1, 0, 0, 0], np.float32), (B, 1))
    feet = np.broadcast_to(W.NOMINAL_FEET, (B, 4, 3)).astype(np.float32) + p[:, None, :]
    tk = T.pack_ticks(p, v, q, np.zeros((B, 3)), feet, np.zeros(B), p[:, :2], np.zeros(B), np.zeros(B), v[:, :2],
                      (0, 5, 5, 0), (5, 5, 5, 5), np.arange(B), body_height=0.25)
    rec_t, _ = oracle.build_records(tk, h)
    assert np.array_equal(rec_t, rec)


def test_x_drag_integrator_threshold_is_the_reference_double_compare(oracle):
    """ConvexMPCLocomotion.cpp:636 compares the float v_x with the DOUBLE literal 0.3: v_x == 0.3f
    (0.300000012 > 0.3) integrates, v_x just below does not.  Kernel body == oracle on exactly those values."""
    h = 10
    tk = T.synth_ticks(8, h, 21)
    edge = np.float32(0.3)
    below = np.nextafter(edge, np.float32(0))
    tk[:, T.TICK_V] = np.array([edge, -edge, below, -below, np.nextafter(edge, np.float32(1)), 0.0, 0.31, -0.31],
                               np.float32)
    tk[:, T.TICK_XDRAG] = 0.125
    tk[:, T.TICK_P + 2] = 0.29
    rec_o, st_o = oracle.build_records(tk, h)
    rec_e, st_e = emu_build_records(tk, h)
    assert np.array_equal(rec_o, rec_e) and np.array_equal(st_o, st_e)
    moved = st_o[:, 2] != np.float32(0.125)
    assert moved.tolist() == [True, True, False, False, True, False, True, True]


def test_trot_tick_workloads_rebuild_the_record_workloads():
    """workloads.config2_ticks / config4_ticks (the bench's e2e_ticks leg) describe the problems of config2 / config4:
    records built from them (oracle restatement of the reference's host code) equal the packed records up to the fp32
    rounding of the COM-relative feet, the gait tables byte for byte."""
    from quadruped_ctrl_b200 import records as R
    from quadruped_ctrl_b200 import workloads as W
    from oracle import oracle as O
    h = 10
    for ticks_fn, rec_fn, seed in ((W.config2_ticks, W.config2, 1234), (W.config4_ticks, W.config4, 3456)):
        rec_t, _ = O.build_records(ticks_fn(256, h, seed), h)
        rec = rec_fn(256, h, seed)
        go = R.gait_offset(h) if hasattr(R, "gait_offset") else 4 * (48 + 12 * h)
        assert np.array_equal(rec_t[:, go:go + 4 * h], rec[:, go:go + 4 * h])
        a = rec_t[:, :go].copy().view(np.float32)
        b = rec[:, :go].copy().view(np.float32)
        assert np.abs(a - b).max() < 1e-6
