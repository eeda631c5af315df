"""Logic of the kernel source (csrc/mpc_core.h) on the CPU tier.

The CUDA kernel body is compiled here as single-thread host code (tests/emu/emu.cpp -- test-only, never
part of the product library) and checked against the golden fixture and the oracle, so that algorithmic
regressions are caught without a GPU.  The -m gpu tests run the same source as the real kernel.
"""
import numpy as np
import pytest

from quadruped_ctrl_b200 import records as R
from quadruped_ctrl_b200 import workloads as W

from common import GOLDEN_CASES, emu_solve, load_golden, rel

ST_OPT, ST_MAXIT, ST_BAD, ST_NOTPD, ST_NOSTANCE, ST_RETRY = 0, 1, 2, 3, 4, 0x40


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_matches_golden(name):
    G = load_golden()
    rec, h = G[name + "_records"], int(G[name + "_h"])
    e = emu_solve(rec, h)
    assert (e["status"] == ST_OPT).all()
    ok = G[name + "_o64_rc"] == 0
    assert (e["nv"] == G[name + "_o64_nv"]).all()
    # whole 12h solution vs the reference solver on the fp64-assembled QP: agreement to round-off
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
def test_matches_oracle_on_seeded_batches(name, batch, oracle):
    h = W.HORIZONS[name]
    rec = W.CONFIGS[name](batch, seed=4242)
    e = emu_solve(rec, h, want_qp=True)
    o = oracle.solve_batch(rec, h, 64, want_qp=True)
    assert (e["status"] == ST_OPT).all()
    for b in range(batch):
        nv = int(o["nv"][b])
        assert e["nv"][b] == nv
        assert np.abs(e["H"][b][:nv, :nv] - o["H"][b][:nv, :nv]).max() <= 1e-12 * np.abs(o["H"][b]).max()
        assert np.abs(e["g"][b][:nv] - o["g"][b][:nv]).max() <= 1e-12 * np.abs(o["g"][b]).max()
    ok = o["rc"] == 0
    assert rel(e["sol"], o["sol"])[ok].max() < 1e-10


def test_problems_the_reference_gives_up_on_are_still_solved(oracle):
    """qpOASES stops at nWSR = 100 (SolverMPC.cpp:435) and the reference returns stale memory; the engine returns
    the optimum (checked against the oracle's own exact solver)."""
    rec = W.config3(128)
    e = emu_solve(rec, 20)
    p = oracle.solve_batch(rec, 20, 64, "port")
    assert (e["status"] == ST_OPT).all()
    assert rel(e["sol"], p["sol"]).max() < 1e-10


def test_kkt_conditions_hold():
    """Size-independent property: the returned point is the KKT point of the reduced QP."""
    for name, B in (("config2", 64), ("config3", 24), ("edge", None)):
        if name == "edge":
            rec, h = load_golden()["edge_records"], 10
        else:
            h = W.HORIZONS[name]
            rec = W.CONFIGS[name](B, seed=7)
        e = emu_solve(rec, h, want_qp=True)
        f = R.unpack_records(rec, h)
        for b in range(rec.shape[0]):
            nv = int(e["nv"][b])
            H, g = e["H"][b][:nv, :nv], e["g"][b][:nv]
            keep = np.repeat(f["gait"][b] != 0, 3)
            x = e["sol"][b][keep]
            assert (e["sol"][b][~keep] == 0).all()
            fmax = float(f["f_max"][b])
            # the cone slope the reference uses is the FLOAT quotient 1/mu (f_block is float, SolverMPC.cpp:361-372)
            mui = float(np.float32(1.0) / np.float32(f["mu"][b]))
            mu = 1.0 / mui
            X = x.reshape(-1, 3)
            # primal feasibility
            assert (X[:, 2] >= -1e-7).all() and (X[:, 2] <= fmax + 1e-7).all()
            assert (np.abs(X[:, 0]) <= mu * X[:, 2] + 1e-7).all() and (np.abs(X[:, 1]) <= mu * X[:, 2] + 1e-7).all()
            # stationarity: the gradient lies in the cone of active constraint normals -> solve the small NNLS
            from scipy.optimize import nnls
            grad = H @ x + g
            cols = []
            for j in range(X.shape[0]):
                fx, fy, fz = X[j]
                rows = [((fx * mui + fz), [mui, 0, 1]), ((-fx * mui + fz), [-mui, 0, 1]),
                        ((fy * mui + fz), [0, mui, 1]), ((-fy * mui + fz), [0, -mui, 1]), (fz, [0, 0, 1]),
                        (fmax - fz, [0, 0, -1])]
                for slack, n in rows:
                    if slack < 1e-6:
                        v = np.zeros(nv)
                        v[3 * j:3 * j + 3] = n
                        cols.append(v)
            if cols:
                N = np.stack(cols, 1)
                lam, res = nnls(N, grad)
            else:
                res = np.linalg.norm(grad)
            assert res <= 1e-6 * max(1.0, np.linalg.norm(g)), (name, b, res)


def test_status_codes():
    h = 10
    rec = W.config2(6, h, 5)
    f = rec.view(np.float32)
    go = R.gait_offset(h)
    rec[0, go:go + 4 * h] = 0            # no stance anywhere
    f[1, R.REC_P] = np.nan               # non-finite input
    f[2, R.REC_MU] = 0.0                 # invalid friction coefficient
    f[3, R.REC_MASS] = -1.0
    f[4, R.REC_FMAX] = 0.001             # every row "near zero": everything eliminated (SolverMPC.cpp:448-452)
    e = emu_solve(rec, h)
    assert e["status"].tolist() == [ST_NOSTANCE, ST_BAD, ST_BAD, ST_BAD, ST_NOSTANCE, ST_OPT]
    assert (e["forces"][:5] == 0).all() and (e["sol"][:5] == 0).all()
    assert np.abs(e["forces"][5]).max() > 1.0


def test_iteration_cap_and_working_set_overflow():
    rec = W.four_stance(8, 10, 3)
    full = emu_solve(rec, 10)
    assert (full["status"] == ST_OPT).all() and full["m"].max() > 4
    capped = emu_solve(rec, 10, max_iter=2)
    assert (capped["status"] == ST_MAXIT).any()
    # a working-set tile smaller than the optimum's active set is reported for a retry in the big class
    small = emu_solve(rec, 10, nv_cap=120, m_cap=4)
    over = full["m"] > 4
    assert (small["status"][over] == ST_RETRY).all()
    same = ~over
    assert (small["sol"][same] == full["sol"][same]).all()


def test_gait_values_above_one_scale_the_force_limit():
    """U_b = gait * f_max (SolverMPC.cpp:349-358): a table entry of 2 doubles the bound instead of being a flag."""
    h = 10
    rec = W.config2(4, h, 11)
    f = rec.view(np.float32)
    f[:, R.REC_FMAX] = 10.0
    lo = emu_solve(rec, h)
    go = R.gait_offset(h)
    rec2 = rec.copy()
    rec2[:, go:go + 4 * h] *= 2
    hi = emu_solve(rec2, h)
    assert lo["sol"].max() <= 10.0 + 1e-9
    assert hi["sol"].max() > 10.0 + 1e-3 and hi["sol"].max() <= 20.0 + 1e-9


def test_packed_layout_gives_the_same_answers(monkeypatch):
    """The register-resident classes keep H / H^{-1} as a packed lower triangle (hix() with ld < 0).  The host
    emulation of that layout (assembly stores and active-set reads go through the packed index) must agree with
    the full-storage layout to round-off."""
    rec = W.CONFIGS["config2"](48, seed=77)
    rec4 = W.CONFIGS["four_stance"](16, seed=78)
    for r, h in ((rec, 10), (rec4, 10)):
        monkeypatch.setenv("MPC_EMU_PACKED", "0")
        full = emu_solve(r, h, want_qp=True)
        monkeypatch.setenv("MPC_EMU_PACKED", "1")
        packed = emu_solve(r, h, want_qp=True)
        assert (packed["status"] == ST_OPT).all() and (full["status"] == ST_OPT).all()
        assert np.array_equal(packed["H"], full["H"]) and np.array_equal(packed["g"], full["g"])
        assert rel(packed["sol"], full["sol"]).max() < 1e-10
        assert (packed["iters"] == full["iters"]).all()


REGIONS = ["sc", "g", "x", "stance", "posk", "amask", "W", "Wia", "Wiz", "C", "M", "xs", "qe", "psum", "mom", "T", "ck",
           "ub", "Wca", "Wcz", "w", "r", "u", "tcol", "red", "Hm"]


def _regions(h, nv_cap, m_cap, npad, packed, pipe, which):
    import ctypes
    from common import emu_lib
    L = emu_lib()
    out = (ctypes.c_long * (2 * len(REGIONS)))()
    fb = ctypes.c_long()
    n = L.emu_layout_regions(h, nv_cap, m_cap, npad, packed, pipe, which, out, ctypes.byref(fb))
    assert n == len(REGIONS)
    return {REGIONS[i]: (out[2 * i], out[2 * i + 1]) for i in range(n)}, fb.value


@pytest.mark.parametrize("h,nv_cap,m_cap,npad,packed", [(10, 60, 27, 64, 0), (14, 60, 21, 64, 0), (16, 96, 33, 96, 0),
                                                         (10, 120, 31, 128, 1), (20, 128, 33, 128, 1)])
def test_piped_layout_keeps_the_two_roles_apart(h, nv_cap, m_cap, npad, packed):
    """mpc_solve_pipe_kernel: while one role runs the active set of problem n-1 (set A) the other assembles problem n
    (set B).  Everything the assembly front writes must be disjoint from everything the active set touches, every
    region must lie inside the workspace, and the two per-problem sets must not share their scalars / stance lists."""
    A, fb = _regions(h, nv_cap, m_cap, npad, packed, 1, 0)
    B, fb2 = _regions(h, nv_cap, m_cap, npad, packed, 1, 1)
    assert fb == fb2
    for R in (A, B):
        for name, (lo, hi) in R.items():
            assert 0 <= lo <= hi <= fb, (name, lo, hi, fb)
    front_writes = ["sc", "stance", "posk", "amask", "C", "M", "xs", "qe", "psum", "mom", "g"]
    active_set = ["sc", "x", "posk", "amask", "W", "Wia", "Wiz", "T", "ub", "Wca", "Wcz", "w", "r", "u", "tcol", "Hm"]

    def overlap(a, b):
        return a[0] < b[1] and b[0] < a[1]
    for f in front_writes:
        for g in active_set:
            assert not overlap(B[f], A[g]), (f, g, B[f], A[g])
            assert not overlap(A[f], B[g]), (f, g)
    # within one set no two distinct regions overlap either, except the deliberate aliases of the un-piped layout
    names = list(A)
    for i, a in enumerate(names):
        for b in names[i + 1:]:
            assert not overlap(A[a], A[b]), (a, b, A[a], A[b])


def test_unpiped_layout_regions_are_inside_the_workspace():
    for packed in (0, 1):
        R, fb = _regions(10, 60, 33, 64, packed, 0, 0)
        for name, (lo, hi) in R.items():
            assert 0 <= lo <= hi <= fb, (name, lo, hi, fb)


def test_wrench_space_class_matches_the_dense_route(oracle):
    """The wrench-space class (H^{-1} = (I - G'MG)/(2 alpha), two inversions of size 6h instead of one of size nv;
    host build of the same source the CUDA kernel runs) against the reference solver and the dense route:
    same optimum for three- and four-stance problems at horizons 10, 16, 20, hard (many active rows) included."""
    from quadruped_ctrl_b200 import records as R
    from common import emu_solve_wrench
    for name, h, B in (("four_stance", 10, 6), ("four_stance", 16, 4), ("config3", 20, 40)):
        rec = W.four_stance(B, h, 3) if name == "four_stance" else W.config3(B, h, 7)
        o = oracle.solve_batch(rec, h, 64)
        wr = emu_solve_wrench(rec, h)
        ok = o["rc"] == 0
        assert (wr["status"] == 0).all()
        assert rel(wr["sol"], o["sol"])[ok].max() < 1e-9
    rec = W.four_stance(6, 20, 5)
    rec.view(np.float32)[:, R.REC_FMAX] = 7.0       # most fz rows saturate: > 100 working-set changes
    port = oracle.solve_batch(rec, 20, 64, "port")
    wr = emu_solve_wrench(rec, 20)
    assert (wr["status"] == 0).all() and wr["iters"].mean() > 100
    assert rel(wr["sol"], port["sol"]).max() < 1e-8
    # a working-set tile that is too small is reported for re-queueing, not silently truncated
    wr = emu_solve_wrench(rec, 20, m_cap=16)
    assert (wr["status"] == 0x40).all()
