"""Logic of the Riccati solver (csrc/mpc_riccati.h) on the CPU tier.

The solver never forms the condensed Hessian: every H^{-1} product of the dual active-set method is a backward and a
forward sweep over the horizon with the gains of the Riccati recursion.  Its source is compiled here as single-thread
host code (tests/emu/emu.cpp -- test-only, never part of the product library) and checked against the golden fixture
(reference qpOASES on the fp64-assembled dense QP), the oracle and the explicit-inverse path of csrc/mpc_core.h.
The -m gpu tests run the same algorithm as the real kernel (tensor-pipe factorisation, register-resident sweeps).
"""
import numpy as np
import pytest

from quadruped_ctrl_b200 import records as R
from quadruped_ctrl_b200 import workloads as W

from common import GOLDEN_CASES, emu_solve, emu_solve_riccati, load_golden, rel

ST_OPT, ST_MAXIT, ST_BAD, ST_NOTPD, ST_NOSTANCE, ST_RETRY = 0, 1, 2, 3, 4, 0x40


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_matches_golden(name):
    G = load_golden()
    rec, h = G[name + "_records"], int(G[name + "_h"])
    e = emu_solve_riccati(rec, h)
    assert (e["status"] == ST_OPT).all()
    ok = G[name + "_o64_rc"] == 0
    assert (e["nv"] == G[name + "_o64_nv"]).all()
    # whole 12h solution vs the reference solver on the fp64-assembled dense QP: agreement to round-off
    assert rel(e["sol"], G[name + "_o64_sol"])[ok].max() < 1e-10
    # first-step forces vs the reference-faithful fp32 path: never further than that path's own rounding cloud
    cloud = rel(G[name + "_o32_sol"][:, :12], G[name + "_o64_sol"][:, :12])
    e32 = rel(e["forces"], G[name + "_o32_sol"][:, :12])
    ok32 = ok & (G[name + "_o32_rc"] == 0)
    assert (e32[ok32] <= cloud[ok32] + 1e-5).all()
    well = ok32 & (cloud <= 2e-5)
    if well.any():
        assert e32[well].max() <= 1e-4


@pytest.mark.parametrize("name,batch", [("config2", 200), ("config4", 200), ("four_stance", 40), ("config5", 48),
                                        ("config3", 64)])
def test_same_optimum_and_same_iterations_as_the_inverse_path(name, batch, oracle):
    """Both solvers run the same dual active-set method (same selection rule, same step rules); only the way H^{-1}
    products are formed differs.  They must visit the same working sets and end at the same point."""
    h = W.HORIZONS[name]
    rec = W.CONFIGS[name](batch, seed=4242)
    r = emu_solve_riccati(rec, h)
    e = emu_solve(rec, h)
    o = oracle.solve_batch(rec, h, 64)
    assert (r["status"] == ST_OPT).all()
    assert (r["nv"] == e["nv"]).all()
    ok = o["rc"] == 0
    assert rel(r["sol"], o["sol"])[ok].max() < 1e-10
    assert rel(r["sol"], e["sol"]).max() < 1e-10
    # ties between equally violated rows may resolve differently in the last bits; the counts agree almost everywhere
    assert (r["iters"] == e["iters"]).mean() > 0.9
    assert (r["m"] == e["m"]).mean() > 0.9


def test_problems_the_reference_gives_up_on_are_still_solved(oracle):
    rec = W.config3(96)
    r = emu_solve_riccati(rec, 20)
    p = oracle.solve_batch(rec, 20, 64, "port")
    assert (r["status"] == ST_OPT).all()
    assert rel(r["sol"], p["sol"]).max() < 1e-10


@pytest.mark.parametrize("h", [1, 2, 9, 36])
def test_extreme_horizons(h, oracle):
    rec = np.concatenate([W.config2(6, h, 5), W.four_stance(4, h, 6)])
    r = emu_solve_riccati(rec, h)
    o = oracle.solve_batch(rec, h, 64)
    assert ((r["status"] == ST_OPT) | (r["status"] == ST_NOSTANCE)).all()  # (a trot table of one row can be all swing)
    ok = o["rc"] == 0
    assert ok.sum() >= 3 and rel(r["sol"], o["sol"])[ok].max() < 1e-9


def test_status_codes():
    h = 10
    rec = W.config2(6, h, 99)
    f = rec.view(np.float32)
    go = R.gait_offset(h)
    rec[0, go:go + 4 * h] = 0            # no stance anywhere
    f[1, R.REC_P] = np.nan               # non-finite input
    f[2, R.REC_MU] = 0.0                 # invalid friction coefficient
    f[3, R.REC_MASS] = -1.0
    f[4, R.REC_FMAX] = 0.001             # every row "near zero": everything eliminated (SolverMPC.cpp:448-452)
    e = emu_solve_riccati(rec, h)
    assert e["status"].tolist() == [ST_NOSTANCE, ST_BAD, ST_BAD, ST_BAD, ST_NOSTANCE, ST_OPT]
    assert (e["forces"][:5] == 0).all() and (e["sol"][:5] == 0).all()
    assert np.abs(e["forces"][5]).max() > 1.0
    # alpha = 0 with all-zero weights: S_k = B'PB + alpha I is singular -> not positive definite
    rec2 = W.config2(2, h, 5)
    f2 = rec2.view(np.float32)
    f2[:, R.REC_ALPHA] = 0.0
    f2[:, R.REC_WEIGHTS:R.REC_WEIGHTS + 12] = 0.0
    assert (emu_solve_riccati(rec2, h)["status"] == ST_NOTPD).all()


def test_iteration_cap_and_column_tile_overflow():
    rec = W.four_stance(8, 10, 3)
    full = emu_solve_riccati(rec, 10)
    assert (full["status"] == ST_OPT).all() and full["m"].max() > 4
    capped = emu_solve_riccati(rec, 10, max_iter=2)
    assert (capped["status"] == ST_MAXIT).any()
    # a column tile (Z = H^{-1} N) smaller than the optimum's active set: reported for a retry in another class
    small = emu_solve_riccati(rec, 10, nv_cap=120, m_cap=4)
    over = full["m"] > 4
    assert (small["status"][over] == ST_RETRY).all()
    same = ~over
    assert (small["sol"][same] == full["sol"][same]).all()
    # with a global slab behind the tile the working set moves there and the solve carries on: same bits as with a
    # tile that was large enough from the start
    moved = emu_solve_riccati(rec, 10, nv_cap=120, m_cap=4, with_slab=True)
    assert (moved["status"] == ST_OPT).all()
    assert (moved["sol"] == full["sol"]).all() and (moved["iters"] == full["iters"]).all()


def test_flight_phases_and_mixed_stance_counts(oracle):
    """Steps without any stance leg (gallop, pronk) carry no control: the recursion passes through them with
    P <- Q + A'PA; steps with one to four stance legs give 3 to 12 controls."""
    h = 12
    rec = W.config2(16, h, 21)
    go = R.gait_offset(h)
    rng = np.random.default_rng(3)
    gait = (rng.random((16, h, 4)) < 0.45).astype(np.uint8)
    gait[:, 3] = 0            # a flight phase in every problem
    gait[:, 7] = 1            # and a four-stance step
    rec[:, go:go + 4 * h] = gait.reshape(16, -1)
    r = emu_solve_riccati(rec, h)
    o = oracle.solve_batch(rec, h, 64)
    assert (r["status"] == ST_OPT).all()
    ok = o["rc"] == 0
    assert ok.any() and rel(r["sol"], o["sol"])[ok].max() < 1e-9


def test_x_drag_couples_through_the_sparse_part_of_A(oracle):
    """x_drag adds A[11][9] and the dt^2/2 term in row 5 (SolverMPC.cpp:241): the sparse tables of N carry it."""
    h = 10
    rec = W.config2(12, h, 8)
    rec.view(np.float32)[:, R.REC_XDRAG] = np.linspace(-0.8, 0.9, 12, dtype=np.float32)
    r = emu_solve_riccati(rec, h)
    o = oracle.solve_batch(rec, h, 64)
    ok = o["rc"] == 0
    assert (r["status"] == ST_OPT).all() and rel(r["sol"], o["sol"])[ok].max() < 1e-9


def test_randomised_parameters_far_from_the_reference_defaults(oracle):
    """Horizons 3 ... 14, random contact tables, alpha over 3.5 decades (down to 3e-7), weights over four decades with
    a fifth of them zero, random dt / mu / f_max / mass / inertia / x_drag: both solvers stay optimal and agree with the
    independent dense active-set port (the conditioning of these QPs reaches 1e9+: 1e-7 is what is asserted, ~1e-9 is
    what is measured, the Riccati route usually the closer of the two)."""
    rng = np.random.default_rng(123)
    for trial in range(8):
        h, B = int(rng.integers(3, 15)), 16
        rec = W.config2(B, h, 1000 + trial)
        f = rec.view(np.float32)
        go = R.gait_offset(h)
        gait = (rng.random((B, h, 4)) < rng.uniform(0.3, 0.9)).astype(np.uint8)
        rec[:, go:go + 4 * h] = gait.reshape(B, -1)
        f[:, R.REC_ALPHA] = 10 ** rng.uniform(-6.5, -3, B).astype(np.float32)
        w = (10 ** rng.uniform(-2, 2.3, (B, 12))).astype(np.float32)
        w[rng.random((B, 12)) < 0.2] = 0
        f[:, R.REC_WEIGHTS:R.REC_WEIGHTS + 12] = w
        f[:, R.REC_DT] = rng.uniform(0.01, 0.05, B).astype(np.float32)
        f[:, R.REC_MU] = rng.uniform(0.2, 1.0, B).astype(np.float32)
        f[:, R.REC_FMAX] = rng.uniform(40, 300, B).astype(np.float32)
        f[:, R.REC_MASS] = rng.uniform(5, 40, B).astype(np.float32)
        f[:, R.REC_IBODY:R.REC_IBODY + 3] = rng.uniform(0.03, 1.5, (B, 3)).astype(np.float32)
        f[:, R.REC_XDRAG] = rng.uniform(-0.5, 0.5, B).astype(np.float32)
        r = emu_solve_riccati(rec, h)
        e = emu_solve(rec, h)
        o = oracle.solve_batch(rec, h, 64, "port")
        ok = o["rc"] == 0
        some = ((r["status"] == ST_OPT) | (r["status"] == ST_NOSTANCE))
        assert some.all() and (r["status"] == e["status"]).all()
        assert ok.any()
        assert rel(r["sol"], o["sol"])[ok].max() < 1e-7
        assert rel(e["sol"], o["sol"])[ok].max() < 1e-6   # (the explicit inverse loses a digit more on the worst of them)
