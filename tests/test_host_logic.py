"""Host-side logic (CPU tier): gait tables, record packing, workload synthesis, sharding arithmetic."""
import numpy as np

from quadruped_ctrl_b200 import gait as G
from quadruped_ctrl_b200 import records as R
from quadruped_ctrl_b200 import workloads as W
from quadruped_ctrl_b200.sharding import shard_bounds, shard_sizes


def naive_mpc_table(n, offsets, durations, iteration):
    """Literal restatement of OffsetDurationGait::getMpcTable (Gait.cpp:142-166)."""
    t = np.zeros(n * 4, np.int32)
    for i in range(n):
        it = (i + iteration + 1) % n
        for j in range(4):
            progress = it - offsets[j]
            if progress < 0:
                progress += n
            t[i * 4 + j] = 1 if progress < durations[j] else 0
    return t


def test_mpc_table_matches_literal_loop():
    for name, (off, dur) in G.GAITS_14.items():
        for it in range(14):
            assert (G.mpc_table(14, off, dur, it).reshape(-1) == naive_mpc_table(14, off, dur, it)).all(), name
    off, dur = (0, 5, 5, 0), (5, 5, 5, 5)   # the mode-1 h=10 trot (ConvexMPCLocomotion.cpp:191-192)
    its = np.arange(10)
    batch = G.mpc_tables(10, off, dur, its)
    for it in its:
        assert (batch[it] == naive_mpc_table(10, off, dur, int(it))).all()


def test_trot_has_two_feet_down_every_step():
    t = G.mpc_tables(10, (0, 5, 5, 0), (5, 5, 5, 5), np.arange(10)).reshape(10, 10, 4)
    assert (t.sum(-1) == 2).all()
    assert (t[..., 0] == t[..., 3]).all() and (t[..., 1] == t[..., 2]).all() and (t[..., 0] != t[..., 1]).all()


def test_gait_number_mapping_and_rescale():
    assert G.gait_by_number(0) == "trotting" and G.gait_by_number(3) == "trotting" and G.gait_by_number(6) == "trotting"
    assert G.gait_by_number(4) == "standing" and G.gait_by_number(7) == "galloping"
    assert G.rescale(*G.GAITS_14["galloping"], 16) == ((0, 5, 8, 13), (8, 8, 8, 8))
    assert G.rescale(*G.GAITS_14["standing"], 20) == ((0, 0, 0, 0), (20, 20, 20, 20))
    assert G.set_iterations(10, 13, 13 * 23 + 5) == (3, (13 * 23 + 5) % 130 / 130.0)


def test_record_round_trip_and_layout():
    rng = np.random.default_rng(0)
    for h in (1, 10, 16, 36):
        B = 5
        fields = dict(p=rng.normal(size=(B, 3)), v=rng.normal(size=(B, 3)), q=rng.normal(size=(B, 4)),
                      w=rng.normal(size=(B, 3)), r=rng.normal(size=(B, 12)), yaw=rng.normal(size=B),
                      traj=rng.normal(size=(B, 12 * h)), gait=rng.integers(0, 2, (B, 4 * h)))
        rec = R.pack_records(h, **fields, x_drag=rng.normal(size=B), mass=rng.uniform(5, 12, B),
                             I_body=rng.uniform(0.05, 0.3, (B, 3)))
        assert rec.shape == (B, R.record_stride(h)) and rec.dtype == np.uint8
        u = R.unpack_records(rec, h)
        for k in ("p", "v", "q", "w", "r", "traj"):
            assert (u[k] == fields[k].astype(np.float32)).all(), k
        assert (u["gait"] == fields["gait"]).all()
        assert (u["weights"] == R.DEFAULT_WEIGHTS).all() and (u["alpha"] == R.DEFAULT_ALPHA).all()
        assert (u["mu"] == np.float32(0.4)).all() and (u["f_max"] == 120).all() and (u["dt"] == np.float32(0.026)).all()
    assert R.algorithmic_bytes(10) == 756 and R.algorithmic_bytes(16) == 1068 and R.algorithmic_bytes(20) == 1276


def test_workloads_are_seeded_and_shaped():
    a, b = W.config2(64), W.config2(64)
    assert (a == b).all() and not (a == W.config2(64, seed=1)).all()
    for name, B in (("config2", 32), ("config3", 32), ("config4", 32), ("config5", 32), ("four_stance", 8)):
        h = W.HORIZONS[name]
        rec = W.CONFIGS[name](B)
        assert rec.shape == (B, R.record_stride(h))
        u = R.unpack_records(rec, h)
        assert np.isfinite(rec.view(np.float32)[:, :R.REC_TRAJ + 12 * h]).all()
        assert np.allclose(np.linalg.norm(u["q"], axis=1), 1.0, atol=1e-6)
        assert set(np.unique(u["gait"])) <= {0, 1}
    assert (R.unpack_records(W.config2(16), 10)["gait"].reshape(16, 10, 4).sum(-1) == 2).all()
    u3 = R.unpack_records(W.config3(256), 20)
    assert u3["mass"].min() >= 9 * 0.8 - 1e-5 and u3["mass"].max() <= 9 * 1.2 + 1e-5
    assert len(np.unique(u3["gait"].reshape(256, -1).sum(1))) > 3     # mixed gaits -> mixed problem sizes
    assert W.config1().shape[0] == 10


def test_trajectory_builder_matches_the_reference_rollout():
    """trajAll of ConvexMPCLocomotion.cpp:547-576: constant rpy/height/velocity, position and yaw integrated."""
    t = W.build_trajectory(4, 0.026, (np.array([0.01]), np.array([0.02])), np.array([0.3]), np.array([1.0]),
                           np.array([2.0]), np.array([0.5]), (np.array([0.4]), np.array([-0.2]))).reshape(4, 12)
    dt = np.float32(0.026)
    for i in range(4):
        assert np.allclose(t[i, [0, 1, 5, 8, 9, 10]], [0.01, 0.02, 0.25, 0.5, 0.4, -0.2])
        assert np.isclose(t[i, 3], 1.0 + i * dt * 0.4, atol=1e-6) and np.isclose(t[i, 4], 2.0 - i * dt * 0.2, atol=1e-6)
        assert np.isclose(t[i, 2], 0.3 + i * dt * 0.5, atol=1e-6)
        assert (t[i, [6, 7, 11]] == 0).all()


def test_shard_bounds_tile_the_batch():
    for total in (0, 1, 7, 4096, 65536, 65537):
        for world in (1, 2, 3, 4, 8):
            sizes = shard_sizes(total, world)
            assert sum(sizes) == total and max(sizes) - min(sizes) <= 1
            pos = 0
            for r in range(world):
                lo, hi = shard_bounds(total, world, r)
                assert lo == pos
                pos = hi
            assert pos == total
