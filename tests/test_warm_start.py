"""Warm start on the CPU tier, through the host build of the kernel body: a closed-loop rollout of
robots on the MPC's own model, every tick solved twice -- cold, and warm from the working set of the previous tick
shifted by one horizon step.  Same optimum (the QP is strictly convex), fewer working-set changes."""
import numpy as np

from quadruped_ctrl_b200 import rollout as RO

from common import emu_solve, emu_solve_warm, rel


def _run(gait, h, B, ticks, seed, **kw):
    ro = RO.Rollout(B, h, gait, seed, **kw)
    cache = np.zeros((B, 128), np.int32)
    it_cold, it_warm, worst = [], [], 0.0
    for t in range(ticks):
        rec = ro.records()
        cold = emu_solve(rec, h)
        warm = emu_solve_warm(rec, h, cache, shift=1)
        assert (cold["status"] == 0).all() and (warm["status"] == 0).all(), (t, np.bincount(cold["status"]))
        worst = max(worst, rel(warm["sol"], cold["sol"]).max())
        it_cold.append(cold["iters"].mean())
        it_warm.append(warm["iters"].mean())
        assert np.abs(ro.p[:, 2] - 0.29).max() < 0.2, "rollout left the neighbourhood of the nominal height"
        ro.advance(cold["forces"])
    return np.array(it_cold), np.array(it_warm), worst


def test_warm_start_reaches_the_cold_optimum_in_a_trot_rollout():
    # a slippery floor and weak legs: friction-cone and force-limit rows are active at almost every tick
    it_cold, it_warm, worst = _run("trotting", 10, 24, 120, 1, mu=0.15, f_max=52.0)
    print("\ntrot h=10, 120 ticks: working-set additions per tick cold %.2f, warm %.2f; max |warm - cold| %.1e"
          % (it_cold[5:].mean(), it_warm[5:].mean(), worst))
    assert worst < 1e-9
    assert it_cold[5:].mean() > 1.0          # the scenario does exercise the active set
    assert it_warm[5:].mean() < 0.85 * it_cold[5:].mean()


def test_warm_start_with_many_active_rows():
    """Standing (four stance legs over the whole horizon) and walking: larger working sets, rows entering and
    leaving every tick; a stale or partly wrong cache must never change the answer."""
    for gait, h, kw in (("standing", 10, dict(v_cmd=0.0, f_max=30.0, mu=0.2)), ("walking", 10, dict(v_cmd=0.3, f_max=34.0, mu=0.2))):
        it_cold, it_warm, worst = _run(gait, h, 8, 40, 2, **kw)
        print("\n%s h=%d: additions per tick cold %.2f, warm %.2f; max |warm - cold| %.1e"
              % (gait, h, it_cold[3:].mean(), it_warm[3:].mean(), worst))
        assert worst < 1e-9
        assert it_warm[3:].mean() <= it_cold[3:].mean() + 0.5


def test_warm_start_from_garbage_is_harmless():
    """A cache full of rows that are not active at all (every row of every stance pair!) only costs time."""
    h, B = 10, 6
    ro = RO.Rollout(B, h, "trotting", 3)
    rec = ro.records()
    cold = emu_solve(rec, h)
    cache = np.zeros((B, 128), np.int32)
    cache[:, 0] = 120
    cache[:, 1:121] = np.random.default_rng(0).integers(0, 4 * h * 6, (B, 120))
    warm = emu_solve_warm(rec, h, cache, shift=0)
    assert (warm["status"] == 0).all()
    assert rel(warm["sol"], cold["sol"]).max() < 1e-9
    cache[:, 0] = 10**6          # a corrupt count is clamped
    cache[:, 1:] = -5
    warm = emu_solve_warm(rec, h, cache, shift=3)
    assert rel(warm["sol"], cold["sol"]).max() < 1e-9
