"""Parity of the CUDA path (through the C ABI) against the CPU oracle on the BASELINE configs.

Tolerances (north_star): first-step 12 forces within 1e-4 relative of the reference CPU path.
The reference assembles the QP in fp32; its own rounding cloud around the exact (fp64) answer is
measured here as |oracle32 - oracle64| and reported beside the GPU numbers.  The GPU assembles
in fp64, so it is compared
  * with oracle64 (reference qpOASES on the fp64-assembled QP) at 1e-9 on the whole 12h solution,
  * with oracle32 (the reference-faithful path) at 1e-4 on the first-step forces for every problem
    whose oracle32 answer is itself within 2e-5 of oracle64 (the well-conditioned set).
"""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from quadruped_ctrl_b200 import engine as E
from quadruped_ctrl_b200 import workloads as W

pytestmark = pytest.mark.gpu


from common import GOLDEN_CASES, load_golden, rel  # noqa: E402


CASES = [("config1", None), ("config2", 512), ("config4", 512), ("four_stance", 128), ("config5", 192),
         ("config3", 192)]


@pytest.mark.parametrize("solver", ["riccati", "inverse"])
@pytest.mark.parametrize("name,batch", CASES)
def test_forces_match_oracle(name, batch, solver, oracle, cuda_engine_factory):
    """Both solvers: Riccati sweeps (csrc/mpc_riccati.h, the default) and the explicit inverse of the condensed Hessian
    (csrc/mpc_core.h)."""
    h = W.HORIZONS[name]
    rec = W.CONFIGS[name]() if batch is None else W.CONFIGS[name](batch)
    B = rec.shape[0]
    eng = cuda_engine_factory(h, B, solver)
    assert eng.solver() == solver
    forces, sol, status = eng.solve_host(rec, want_solution=True)
    code = E.status_code(status)
    assert (code == E.STATUS_OPTIMAL).all(), np.bincount(code)
    backend = oracle.default_backend()
    o64 = oracle.solve_batch(rec, h, 64, backend)
    o32 = oracle.solve_batch(rec, h, 32, backend)
    ok64 = o64["rc"] == 0
    # whole 12h solution against the fp64 truth
    e64 = rel(sol, o64["sol"])
    print("\n[%s/%s] B=%d backend=%s  |gpu-o64| max %.2e  med %.2e" % (name, solver, B, backend, e64[ok64].max(), np.median(e64)))
    assert e64[ok64].max() < 1e-9   # agreement with the reference solver to round-off (fp64 QP)
    # first-step forces against the reference-faithful fp32 path, on its well-conditioned set
    cloud = rel(o32["forces"], o64["forces"])
    e32 = rel(forces.astype(np.float64), o32["forces"])
    well = ok64 & (o32["rc"] == 0) & (cloud <= 2e-5)
    print("[%s] |gpu-o32| max on well-conditioned set (%d/%d) %.2e ; reference's own fp32 cloud max %.2e" %
          (name, well.sum(), B, e32[well].max() if well.any() else 0.0, cloud.max()))
    if well.any():
        assert e32[well].max() <= 1e-4
    # triangle inequality: the GPU is never further from the reference path than that path is from the exact answer
    assert (e32[ok64] <= cloud[ok64] + 1e-5).all()
    # swing legs are exactly zero
    gait = rec[:, 4 * (48 + 12 * h):4 * (48 + 12 * h) + 4 * h].reshape(B, h, 4)
    swing = np.repeat(gait == 0, 3, axis=2).reshape(B, 12 * h)
    assert (sol[swing] == 0.0).all()
    assert (forces[swing[:, :12]] == 0.0).all()


def test_assembly_matches_oracle(oracle, cuda_engine_factory):
    for name, B in (("config2", 64), ("config3", 48)):
        h = W.HORIZONS[name]
        rec = W.CONFIGS[name](B)
        eng = cuda_engine_factory(h, B)
        nv, H, g = eng.assemble_device(torch.from_numpy(rec).cuda())
        torch.cuda.synchronize()
        o = oracle.solve_batch(rec, h, 64, "assemble", want_qp=True)
        assert (nv.cpu().numpy() == o["nv"]).all()
        Hn, gn = H.cpu().numpy(), g.cpu().numpy()
        assert np.abs(Hn - o["H"]).max() <= 1e-12 * np.abs(o["H"]).max()
        assert np.abs(gn - o["g"]).max() <= 1e-12 * np.abs(o["g"]).max()


def test_device_entry_matches_host_entry(cuda_engine_factory):
    rec = W.config2(300)
    eng = cuda_engine_factory(10, 300)
    f_host, s_host, st_host = eng.solve_host(rec, want_solution=True)
    f_dev, s_dev, st_dev = eng.solve_device(torch.from_numpy(rec).cuda(), want_solution=True)
    torch.cuda.synchronize()
    assert (f_dev.cpu().numpy() == f_host).all()
    assert (s_dev.cpu().numpy() == s_host).all()
    assert (st_dev.cpu().numpy() == st_host).all()


@pytest.mark.parametrize("solver", ["riccati", "inverse"])
@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_golden_fixture(name, solver, cuda_engine_factory):
    """Committed fixture (reference qpOASES outputs generated in the build container): no oracle needed."""
    G = load_golden()
    rec, h = G[name + "_records"], int(G[name + "_h"])
    eng = cuda_engine_factory(h, rec.shape[0], solver)
    forces, sol, status = eng.solve_host(rec, want_solution=True)
    assert (E.status_code(status) == E.STATUS_OPTIMAL).all()
    ok = G[name + "_o64_rc"] == 0
    assert rel(sol, G[name + "_o64_sol"])[ok].max() < 1e-9
    cloud = rel(G[name + "_o32_sol"][:, :12], G[name + "_o64_sol"][:, :12])
    e32 = rel(forces, G[name + "_o32_sol"][:, :12])
    ok32 = ok & (G[name + "_o32_rc"] == 0)
    assert (e32[ok32] <= cloud[ok32] + 1e-5).all()
    well = ok32 & (cloud <= 2e-5)
    if well.any():
        assert e32[well].max() <= 1e-4


def test_legacy_interface_matches_oracle(oracle):
    """setup_problem -> update_x_drag -> update_solver_settings -> update_problem_data_floats -> get_solution,
    the call sequence of ConvexMPCLocomotion.cpp:630-674, for every gait phase of config 1 and a horizon switch."""
    from quadruped_ctrl_b200 import interface as I
    from quadruped_ctrl_b200 import records as R
    for rec, h in ((W.config1(), 10), (W.config5(4), 16), (W.config1(), 10)):
        f = R.unpack_records(rec, h)
        ref = oracle.solve_batch(rec, h, 64)
        for b in range(rec.shape[0]):
            I.setup_problem(float(f["dt"][b]), h, float(f["mu"][b]), float(f["f_max"][b]))
            I.update_x_drag(float(f["x_drag"][b]))
            I.update_solver_settings(10000, 1e-7, 1e-8, 1.5, 0.1, 0.0)
            I.update_problem_data_floats(f["p"][b], f["v"][b], f["q"][b], f["w"][b], f["r"][b], float(f["yaw"][b]),
                                         f["weights"][b], f["traj"][b], float(f["alpha"][b]), f["gait"][b].astype(np.int32))
            assert I.last_status() == 0
            got = np.array([I.get_solution(i) for i in range(12 * h)])
            assert rel(got[None], ref["sol"][b][None])[0] < 1e-6
    # the double-precision entry narrows to float and takes the same path
    b = 3
    rec, h = W.config1(), 10
    f = R.unpack_records(rec, h)
    I.setup_problem(float(f["dt"][b]), h, float(f["mu"][b]), float(f["f_max"][b]))
    I.update_problem_data(f["p"][b].astype(np.float64), f["v"][b].astype(np.float64), f["q"][b].astype(np.float64),
                          f["w"][b].astype(np.float64), f["r"][b].astype(np.float64), float(f["yaw"][b]),
                          f["weights"][b].astype(np.float64), f["traj"][b].astype(np.float64), float(f["alpha"][b]),
                          f["gait"][b].astype(np.int32))
    ref = oracle.solve_batch(rec, h, 64)
    got = np.array([I.get_solution(i) for i in range(12)])
    assert rel(got[None], ref["forces"][b][None])[0] < 1e-6
    I.shutdown()


def test_cpp_caller_stub_matches_oracle(oracle):
    """SURVEY 8b link test on the GPU: the C++ caller with the reference's call sequence
    (tests/caller/solve_dense_mpc_stub.cpp, linked against the library) gets the oracle's forces."""
    import subprocess
    from quadruped_ctrl_b200 import records as R
    from test_abi import _build_caller_stub
    exe = _build_caller_stub()
    try:
        out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    finally:
        os.unlink(exe)
    assert out.returncode == 0, out.stderr
    lines = [l.split() for l in out.stdout.splitlines() if l.startswith("h ")]
    assert len(lines) == 3
    foot = np.array([[0.19, -0.111, 0], [0.19, 0.111, 0], [-0.19, -0.111, 0], [-0.19, 0.111, 0]], np.float32)
    for l, (h, vx) in zip(lines, ((10, 0.5), (14, 0.5), (10, 0.3))):
        assert int(l[1]) == h and int(l[3]) == 0
        got = np.array([float(x) for x in l[5:17]])
        p = np.array([0, 0, 0.29], np.float32)
        traj = np.zeros((h, 12), np.float32)
        dt = np.float32(0.002) * np.float32(13)
        for i in range(h):
            traj[i, 3] = p[0] + dt * np.float32(i) * np.float32(vx)
            traj[i, 5], traj[i, 9] = 0.25, vx
        gait = np.zeros((h, 4), np.uint8)
        for i in range(h):
            a = 1 if i < h // 2 else 0
            gait[i] = (a, 1 - a, 1 - a, a)
        r = (foot - p[None, :]).T.reshape(-1)   # r[axis*4+leg]
        rec = R.pack_records(h, p=p[None], v=np.array([[vx, 0, 0]], np.float32), q=np.array([[1, 0, 0, 0]], np.float32),
                             w=np.zeros((1, 3), np.float32), r=r[None].astype(np.float32), yaw=np.zeros(1, np.float32),
                             traj=traj.reshape(1, -1), gait=gait.reshape(1, -1), dt=np.array([dt], np.float32))
        ref = oracle.solve_batch(rec, h, 64)
        assert rel(got[None], ref["forces"][:1].astype(np.float64))[0] < 1e-6


@pytest.mark.parametrize("solver", ["riccati", "inverse"])
@pytest.mark.parametrize("h", [1, 2, 9, 36])
def test_extreme_horizons(h, solver, oracle, cuda_engine_factory):
    """Horizon 1 and 2 (smallest problems), 9 (the last horizon whose gait table fits update_data_t.gait without
    running on into hack_pad) and 36 (K_MAX_GAIT_SEGMENTS, the largest the reference's structs can carry: trot
    nv = 216, four-stance nv = 432 -- the catch-all class with its grouped sweep)."""
    rec = np.concatenate([W.config2(6, h, 70 + h), W.four_stance(3, h, 80 + h)])
    eng = cuda_engine_factory(h, rec.shape[0], solver)
    f, s, st = eng.solve_device(torch.from_numpy(rec).cuda(), want_solution=True)
    torch.cuda.synchronize()
    o = oracle.solve_batch(rec, h, 64)
    ok = o["rc"] == 0
    code = E.status_code(st.cpu().numpy())
    assert ((code == E.STATUS_OPTIMAL) | (code == E.STATUS_NO_STANCE)).all()
    assert rel(s.cpu().numpy(), o["sol"])[ok].max() < 1e-9
    assert ok.sum() >= 3


@pytest.mark.parametrize("solver", ["riccati", "inverse"])
def test_status_codes_and_failure_outputs(solver, cuda_engine_factory):
    from quadruped_ctrl_b200 import records as R
    h = 10
    rec = W.config2(6, h, 5)
    f = rec.view(np.float32)
    go = R.gait_offset(h)
    rec[0, go:go + 4 * h] = 0
    f[1, R.REC_P] = np.nan
    f[2, R.REC_MU] = 0.0
    f[3, R.REC_MASS] = -1.0
    f[4, R.REC_FMAX] = 0.001
    eng = cuda_engine_factory(h, 6, solver)
    forces, sol, status = eng.solve_host(rec, want_solution=True)
    assert E.status_code(status).tolist() == [E.STATUS_NO_STANCE, E.STATUS_BAD_INPUT, E.STATUS_BAD_INPUT,
                                              E.STATUS_BAD_INPUT, E.STATUS_NO_STANCE, E.STATUS_OPTIMAL]
    assert (forces[:5] == 0).all() and (sol[:5] == 0).all() and np.abs(forces[5]).max() > 1.0
    eng.set_max_iterations(1)
    rec = W.four_stance(16, h, 3)
    eng2 = cuda_engine_factory(h, 16, solver)
    eng2.set_max_iterations(1)
    f2, s2, st = eng2.solve_host(rec, want_solution=True)
    capped = E.status_code(st) == E.STATUS_MAX_ITER
    assert capped.any()
    # a dual active-set iterate cut short is primal infeasible: it is never handed out as forces
    assert (f2[capped] == 0).all() and (s2[capped] == 0).all()


@pytest.mark.parametrize("solver", ["riccati", "inverse"])
def test_mixed_stance_counts_flight_phases_x_drag_and_singular_hessian(solver, oracle, cuda_engine_factory):
    """Random contact tables (0 to 4 stance legs per step: 0, 3, 6, 9 or 12 controls -- every tile shape of the Riccati
    factorisation, flight phases included), x_drag != 0 (the sparse part of A_d, SolverMPC.cpp:241) and a singular
    Hessian (alpha = 0 with zero weights -> NOT_PD), in batches large enough to run the production kernel of each
    solver."""
    from quadruped_ctrl_b200 import records as R
    h, B = 10, 1536
    rec = W.config2(B, h, 31)
    go = R.gait_offset(h)
    rng = np.random.default_rng(7)
    gait = (rng.random((B, h, 4)) < 0.4).astype(np.uint8)
    gait[:, 2] = 0                      # a flight phase in every problem
    gait[: B // 2, 6] = 1               # a four-stance step in half of them
    rec[:, go:go + 4 * h] = gait.reshape(B, -1)
    f = rec.view(np.float32)
    f[:, R.REC_XDRAG] = rng.uniform(-0.8, 0.9, B).astype(np.float32)
    eng = cuda_engine_factory(h, B, solver)
    forces, sol, status = eng.solve_host(rec, want_solution=True)
    assert (E.status_code(status) == E.STATUS_OPTIMAL).all(), np.bincount(E.status_code(status))
    idx = np.sort(rng.choice(B, 256, replace=False))
    o = oracle.solve_batch(rec[idx], h, 64)
    ok = o["rc"] == 0
    assert ok.sum() > 200 and rel(sol[idx], o["sol"])[ok].max() < 1e-9
    bad = W.config2(1024, h, 5)
    fb = bad.view(np.float32)
    fb[:7, R.REC_ALPHA] = 0.0
    fb[:7, R.REC_WEIGHTS:R.REC_WEIGHTS + 12] = 0.0
    eng2 = cuda_engine_factory(h, 1024, solver)
    f2, s2, st2 = eng2.solve_host(bad, want_solution=True)
    code = E.status_code(st2)
    assert (code[:7] == E.STATUS_NOT_PD).all() and (code[7:] == E.STATUS_OPTIMAL).all()
    assert (f2[:7] == 0).all() and (s2[:7] == 0).all()


@pytest.mark.parametrize("solver", ["riccati", "inverse"])
def test_full_size_properties(solver, cuda_engine_factory):
    """BASELINE sizes (B=4096 config 2; B=65536 config 4 shard-free) through size-independent properties:
    every problem optimal, swing legs exactly zero, friction pyramid and force limits hold, a permuted batch
    gives the permuted answer bit for bit, and repeated solves are bitwise reproducible."""
    for name, B in (("config2", 4096), ("config4", 65536)):
        h = 10
        rec = W.CONFIGS[name](B)
        eng = cuda_engine_factory(h, B, solver)
        d = torch.from_numpy(rec).cuda()
        forces, sol, status = eng.solve_device(d, want_solution=True)
        torch.cuda.synchronize()
        F, S, st = forces.cpu().numpy(), sol.cpu().numpy(), status.cpu().numpy()
        assert (E.status_code(st) == E.STATUS_OPTIMAL).all()
        gait = rec[:, 4 * (48 + 12 * h):4 * (48 + 12 * h) + 4 * h].reshape(B, h * 4)
        X = S.reshape(B, 4 * h, 3)
        assert (X[gait == 0] == 0).all()
        fz = X[..., 2]
        assert (fz >= -1e-7).all() and (fz <= 120 + 1e-7).all()
        assert (np.abs(X[..., 0]) <= 0.4 * fz + 1e-6).all() and (np.abs(X[..., 1]) <= 0.4 * fz + 1e-6).all()
        assert np.array_equal(F, S[:, :12].astype(np.float32))
        perm = np.random.default_rng(0).permutation(B)
        f2, s2, _ = eng.solve_device(torch.from_numpy(rec[perm]).cuda(), want_solution=True)
        torch.cuda.synchronize()
        assert np.array_equal(s2.cpu().numpy(), S[perm])
        f3, _, _ = eng.solve_device(d)
        torch.cuda.synchronize()
        assert np.array_equal(f3.cpu().numpy(), F)


def _reduced_constraints(rec_row, h):
    """(C, lo): the one-sided rows C x >= lo of the reduced QP of one record (stance pairs in ascending (step, leg)
    order, six rows per pair as in csrc/mpc_core.h; 1/mu as the reference's float)."""
    from quadruped_ctrl_b200 import records as R
    f = R.unpack_records(rec_row[None], h)
    gait = f["gait"][0].astype(np.float32)
    fmax = np.float32(f["f_max"][0])
    ub = gait * fmax
    stance = np.nonzero(~((ub < 0.01) & (ub > -0.01)))[0]
    mu_inv = float(np.float32(1.0) / np.float32(f["mu"][0]))
    nv = 3 * len(stance)
    C, lo = [], []
    for j, k in enumerate(stance):
        for ax, sg in ((0, 1), (0, -1), (1, 1), (1, -1)):
            row = np.zeros(nv)
            row[3 * j + ax] = sg * mu_inv
            row[3 * j + 2] = 1.0
            C.append(row)
            lo.append(0.0)
        row = np.zeros(nv)
        row[3 * j + 2] = 1.0
        C.append(row)
        lo.append(0.0)
        row = np.zeros(nv)
        row[3 * j + 2] = -1.0
        C.append(row)
        lo.append(-float(ub[k]))
    return np.array(C), np.array(lo), stance


def _kkt_report(H, g, C, lo, x):
    """Relative KKT residuals of x for min 1/2 x'Hx + g'x, Cx >= lo: (primal violation, stationarity with the best
    non-negative multipliers on the active rows, found by NNLS)."""
    from scipy.optimize import nnls
    slack = C @ x - lo
    primal = max(0.0, -slack.min())
    act = slack < 1e-7 * max(1.0, np.abs(x).max())
    r = H @ x + g
    if act.any():
        lam, res = nnls(C[act].T, r, maxiter=50 * int(act.sum()) + 100)
    else:
        res = np.linalg.norm(r)
    return primal, res / max(1.0, np.linalg.norm(g))


@pytest.mark.parametrize("sweep", ["riccati", "fma", "mma"])
def test_full_size_config3_and_config5(sweep, oracle, cuda_engine_factory):
    """BASELINE sizes of the two configs that were never solved above B=192: config 3 (B=4096, h=20, mixed gaits --
    every size class incl. the catch-all) and config 5 (B=65536, h=16 gallop), with either inversion: every problem
    optimal, size-independent properties on the whole batch, bitwise reproducibility, and a seeded sample against
    the oracle at the tolerances of test_forces_match_oracle."""
    for name, B, n_sample in (("config3", 4096, 384), ("config5", 65536, 1024)):
        h = W.HORIZONS[name]
        rec = W.CONFIGS[name](B)
        eng = cuda_engine_factory(h, B, "riccati" if sweep == "riccati" else "inverse")
        if sweep != "riccati":
            eng.set_sweep_variant(sweep)   # the inverse solver with either register-resident inversion
        d = torch.from_numpy(rec).cuda()
        forces, sol, status = eng.solve_device(d, want_solution=True)
        torch.cuda.synchronize()
        F, S, st = forces.cpu().numpy(), sol.cpu().numpy(), status.cpu().numpy()
        assert (E.status_code(st) == E.STATUS_OPTIMAL).all(), np.bincount(E.status_code(st))
        gait = rec[:, 4 * (48 + 12 * h):4 * (48 + 12 * h) + 4 * h].reshape(B, h * 4)
        X = S.reshape(B, 4 * h, 3)
        assert (X[gait == 0] == 0).all()
        fz = X[..., 2]
        assert (fz >= -1e-7).all() and (fz <= 120 + 1e-7).all()
        assert (np.abs(X[..., 0]) <= 0.4 * fz + 1e-6).all() and (np.abs(X[..., 1]) <= 0.4 * fz + 1e-6).all()
        assert np.array_equal(F, S[:, :12].astype(np.float32))
        f3, _, _ = eng.solve_device(d)
        torch.cuda.synchronize()
        assert np.array_equal(f3.cpu().numpy(), F)
        idx = np.sort(np.random.default_rng(5).choice(B, n_sample, replace=False))
        o64 = oracle.solve_batch(rec[idx], h, 64)
        o32 = oracle.solve_batch(rec[idx], h, 32)
        ok64 = o64["rc"] == 0
        e64 = rel(S[idx], o64["sol"])
        assert e64[ok64].max() < 1e-9
        cloud = rel(o32["forces"], o64["forces"])
        e32 = rel(F[idx].astype(np.float64), o32["forces"])
        well = ok64 & (o32["rc"] == 0) & (cloud <= 2e-5)
        print("\n[%s/%s] B=%d sample %d: |gpu-o64| max %.2e; well-conditioned %d, |gpu-o32| max there %.2e; cloud max %.2e; "
              "reference failures %d" % (name, sweep, B, n_sample, e64[ok64].max(), well.sum(),
                                         e32[well].max() if well.any() else 0, cloud[ok64].max(), (~ok64).sum()))
        if well.any():
            assert e32[well].max() <= 1e-4
        assert (e32[ok64] <= cloud[ok64] + 1e-5).all()
        eng.set_sweep_variant("fma")


@pytest.mark.parametrize("solver", ["riccati", "inverse"])
def test_kkt_where_the_reference_gives_up(solver, oracle, cuda_engine_factory):
    """Problems on which reference qpOASES hits nWSR = 100 and returns an error (SolverMPC.cpp:435, 537-557) are
    masked out of every oracle comparison -- so they are judged here on their own: the GPU answer must satisfy the
    KKT conditions of the fp64-assembled QP (primal feasibility, stationarity with non-negative multipliers on the
    active rows) and agree with the independent active-set port.  Forced with a low f_max on four-stance h=20
    problems (most fz rows saturate: more than 100 working-set changes)."""
    from quadruped_ctrl_b200 import records as R
    h = 20
    rec = W.four_stance(48, h, 21)
    rec.view(np.float32)[:, R.REC_FMAX] = 7.0
    rec = np.concatenate([rec, W.config3(64, h, 22)])
    eng = cuda_engine_factory(h, rec.shape[0], solver)
    forces, sol, status = eng.solve_host(rec, want_solution=True)
    assert (E.status_code(status) == E.STATUS_OPTIMAL).all()
    backend = oracle.default_backend()
    o = oracle.solve_batch(rec, h, 64, backend, want_qp=True)
    failed = np.nonzero(o["rc"] != 0)[0]
    print("\nreference backend %s: %d of %d problems returned an error; GPU iterations on those: %s" %
          (backend, len(failed), rec.shape[0], E.status_iterations(status)[failed][:8]))
    if backend == "reference":
        assert len(failed) >= 8          # the scenario really is one the reference cannot finish
    port = oracle.solve_batch(rec, h, 64, "port")
    check = failed if len(failed) else np.arange(8)
    for b in check[:24]:
        C, lo, stance = _reduced_constraints(rec[b], h)
        nv = C.shape[1]
        assert nv == o["nv"][b]
        H, g = o["H"][b][:nv, :nv], o["g"][b][:nv]
        x = sol[b].reshape(4 * h, 3)[stance].reshape(-1)
        primal, stat = _kkt_report(H, g, C, lo, x)
        assert primal <= 1e-8 and stat <= 1e-8, (b, primal, stat)
        obj = 0.5 * x @ H @ x + g @ x
        xp = port["sol"][b].reshape(4 * h, 3)[stance].reshape(-1)
        objp = 0.5 * xp @ H @ xp + g @ xp
        assert obj <= objp + 1e-9 * max(1.0, abs(objp))
    assert rel(sol, port["sol"]).max() < 1e-7


@pytest.mark.parametrize("solver", ["riccati", "inverse"])
def test_working_set_overflow_is_requeued_not_dropped(solver, oracle, cuda_engine_factory):
    """Problems whose active set outgrows the shared-memory tile of their size class are re-solved by the
    catch-all class inside the same call.  Forced here with f_max so low that most fz rows saturate."""
    from quadruped_ctrl_b200 import records as R
    h = 10
    rec = W.four_stance(64, h, 9)
    rec.view(np.float32)[:, R.REC_FMAX] = 6.0
    eng = cuda_engine_factory(h, 64, solver)
    m_cap = eng.classes()[-2]["m_cap"]
    forces, sol, status = eng.solve_host(rec, want_solution=True)
    assert (E.status_code(status) == E.STATUS_OPTIMAL).all()
    o = oracle.solve_batch(rec, h, 64, "port")
    assert rel(sol, o["sol"]).max() < 1e-6
    assert E.status_iterations(status).max() > m_cap   # at least one problem really did overflow the tile
    # batches of one are classified on the host and launch their class's kernel alone (the legacy single-robot
    # tick); an overflow there is repeated on the general path: same bits as inside the batch
    worst = int(np.argmax(E.status_iterations(status)))
    for b in (worst, 0, 1):
        f1, s1, st1 = eng.solve_host(rec[b:b + 1], want_solution=True)
        assert np.array_equal(f1[0], forces[b]) and np.array_equal(s1[0], sol[b]) and st1[0] == status[b]


def test_pipelined_host_entry_matches_synchronous(cuda_engine_factory):
    """submit_host / wait_host on alternating slots return exactly what solve_host returns."""
    eng = cuda_engine_factory(10, 512)
    batches = [W.config2(512, 10, 100 + i) for i in range(5)]
    ref = [eng.solve_host(b, want_solution=True) for b in batches]
    got = []
    for i, b in enumerate(batches):
        eng.submit_host(i & 1, b, want_solution=True)
        if i > 0:
            f = np.empty((512, 12), np.float32)
            s = np.empty((512, 120), np.float64)
            st = np.empty(512, np.int32)
            eng.wait_host((i - 1) & 1, f, s, st)
            got.append((f, s, st))
    f = np.empty((512, 12), np.float32)
    s = np.empty((512, 120), np.float64)
    st = np.empty(512, np.int32)
    eng.wait_host((len(batches) - 1) & 1, f, s, st)
    got.append((f, s, st))
    for (rf, rs, rst), (gf, gs, gst) in zip(ref, got):
        assert np.array_equal(rf, gf) and np.array_equal(rs, gs) and np.array_equal(rst, gst)


def test_two_device_slots_overlap_without_interference(cuda_engine_factory):
    """Device-resident solves on the engine's two scratch slots / two streams (the bench's pipelined steps):
    interleaved, overlapping launches give bit for bit what one-at-a-time solves give, for mixed size classes."""
    B = 1024
    eng = cuda_engine_factory(10, B)
    batches = [torch.from_numpy(np.concatenate([W.config2(B // 2, 10, 300 + i), W.CONFIGS["four_stance"](B // 2, seed=400 + i)]))
               .cuda() for i in range(6)]
    ref = []
    for b in batches:
        f, s, st = eng.solve_device(b, want_solution=True)
        torch.cuda.synchronize()
        ref.append((f.cpu().numpy(), s.cpu().numpy(), st.cpu().numpy()))
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    outs = []
    for i, b in enumerate(batches):
        with torch.cuda.stream(streams[i & 1]):
            outs.append(eng.solve_device(b, want_solution=True, stream=streams[i & 1], slot=i & 1))
    torch.cuda.synchronize()
    for (rf, rs, rst), (f, s, st) in zip(ref, outs):
        assert np.array_equal(rf, f.cpu().numpy())
        assert np.array_equal(rs, s.cpu().numpy())
        assert np.array_equal(rst, st.cpu().numpy())
    assert (E.status_code(ref[0][2]) == E.STATUS_OPTIMAL).all()


def test_host_entry_copies_before_returning_unless_zero_copy_is_requested(cuda_engine_factory):
    """submit_host copies the records before it returns, whatever memory they are in (the batch here spans several
    512 KB staging chunks): a caller that refills its (page-locked!) buffer right after submit gets the answers of
    the records it submitted.  The zero-copy variant (submit_host_pinned) reads page-locked memory in place and
    refuses pageable memory.  All three give the same bytes as the device entry."""
    B = 2048
    eng = cuda_engine_factory(10, B)
    rec = W.config2(B, 10, 555)
    assert rec.nbytes > 2 * (512 << 10)
    f1, s1, st1 = eng.solve_host(rec, want_solution=True)             # pageable numpy array, staged
    pinned = torch.from_numpy(rec.copy()).pin_memory()
    eng.submit_host(1, pinned.numpy(), want_solution=True)            # page-locked, still staged ...
    pinned.numpy()[:] = 0xFF                                          # ... so scribbling over it at once is harmless
    f2, s2, st2 = np.empty_like(f1), np.empty_like(s1), np.empty_like(st1)
    eng.wait_host(1, f2, s2, st2)
    assert np.array_equal(f1, f2) and np.array_equal(s1, s2) and np.array_equal(st1, st2)
    pinned.numpy()[:] = rec
    eng.submit_host(2, pinned.numpy(), want_solution=True, zero_copy=True)   # read in place by the DMA engine
    f3, s3, st3 = np.empty_like(f1), np.empty_like(s1), np.empty_like(st1)
    eng.wait_host(2, f3, s3, st3)
    assert np.array_equal(f1, f3) and np.array_equal(s1, s3) and np.array_equal(st1, st3)
    with pytest.raises(E.MpcError):
        eng.submit_host(2, rec, zero_copy=True)                       # pageable memory cannot be read in place
    fd, sd, std = eng.solve_device(torch.from_numpy(rec).cuda(), want_solution=True)
    torch.cuda.synchronize()
    assert np.array_equal(f1, fd.cpu().numpy()) and np.array_equal(s1, sd.cpu().numpy())
    assert (E.status_code(st1) == E.STATUS_OPTIMAL).all()


def test_uniform_host_batch_is_one_launch_and_identical(cuda_engine_factory):
    """The host entries classify the batch while they stage it: a uniform batch (one size class) runs ONE kernel
    launch -- no classify kernel, no empty-class launches -- and gives the bytes of the general path; a mixed batch
    takes the general path (classify + one launch per class)."""
    h, B = 10, 1024
    eng = cuda_engine_factory(h, B)
    nc = len(eng.classes())
    rec = W.config2(B, h, 777)
    fd, sd, std = eng.solve_device(torch.from_numpy(rec).cuda(), want_solution=True)   # general path
    torch.cuda.synchronize()
    l0 = eng.kernel_launches()
    f1, s1, st1 = eng.solve_host(rec, want_solution=True)
    assert eng.kernel_launches() - l0 == 1
    assert np.array_equal(f1, fd.cpu().numpy()) and np.array_equal(s1, sd.cpu().numpy())
    assert np.array_equal(st1, std.cpu().numpy())
    mixed = np.concatenate([W.config2(B // 2, h, 778), W.CONFIGS["four_stance"](B // 2, seed=779)])
    fd, sd, std = eng.solve_device(torch.from_numpy(mixed).cuda(), want_solution=True)
    torch.cuda.synchronize()
    l0 = eng.kernel_launches()
    f2, s2, st2 = eng.solve_host(mixed, want_solution=True)
    assert eng.kernel_launches() - l0 == 1 + nc
    assert np.array_equal(f2, fd.cpu().numpy()) and np.array_equal(s2, sd.cpu().numpy())
    assert (E.status_code(st1) == E.STATUS_OPTIMAL).all() and (E.status_code(st2) == E.STATUS_OPTIMAL).all()


def test_host_tick_entry_matches_device_tick_entry(oracle, cuda_engine_factory):
    """mpc_batch_submit_host_ticks: 272-byte tick records from host memory (staged or zero-copy), records built on the
    device -- the forces, solution, status and the controller state written back equal the device tick entry's and the
    oracle's record builder, bit for bit; uniform batches are two launches (builder + one solve kernel)."""
    from quadruped_ctrl_b200 import ticks as T
    for h, mixed, B in ((10, False, 1000), (16, True, 300)):
        tk = T.synth_ticks(B, h, 21 + h, mixed_gaits=mixed)
        _, st_o = oracle.build_records(tk, h)
        eng = cuda_engine_factory(h, B)
        f1, s1, c1, st1 = eng.solve_ticks_device(torch.from_numpy(tk).cuda(), want_solution=True)
        torch.cuda.synchronize()
        pinned = torch.from_numpy(tk.copy()).pin_memory()
        for slot, zero_copy, src in ((0, False, tk), (E.SLOTS - 1, True, pinned.numpy())):
            l0 = eng.kernel_launches()
            eng.submit_host_ticks(slot, src, want_solution=True, zero_copy=zero_copy)
            f = np.empty((B, 12), np.float32)
            s = np.empty((B, 12 * h), np.float64)
            c = np.empty(B, np.int32)
            eng.wait_host(slot, f, s, c)
            if not mixed:
                assert eng.kernel_launches() - l0 == 2
            assert np.array_equal(f, f1.cpu().numpy()) and np.array_equal(s, s1.cpu().numpy())
            assert np.array_equal(c, c1.cpu().numpy())
            assert np.array_equal(eng.host_state(slot)[0][:B], st_o)
        with pytest.raises(E.MpcError):
            eng.submit_host_ticks(0, tk, zero_copy=True)   # pageable memory cannot be read in place
        assert (E.status_code(c1.cpu().numpy()) == E.STATUS_OPTIMAL).all()


def test_device_record_builder_is_byte_exact(oracle, cuda_engine_factory):
    """SURVEY 8f N1 + N2: records built on the device from tick records equal the oracle's restatement of the
    reference's host code (ConvexMPCLocomotion.cpp:498-640, Gait.cpp:142-166) byte for byte, and solving the
    ticks equals solving those records."""
    from quadruped_ctrl_b200 import ticks as T
    for h, mixed, B in ((10, False, 1000), (20, True, 300), (16, True, 300)):
        tk = T.synth_ticks(B, h, 11 + h, mixed_gaits=mixed)
        rec_o, st_o = oracle.build_records(tk, h)
        eng = cuda_engine_factory(h, B)
        d_tk = torch.from_numpy(tk).cuda()
        rec_d, st_d = eng.build_records_device(d_tk)
        torch.cuda.synchronize()
        assert np.array_equal(rec_d.cpu().numpy(), rec_o)
        assert np.array_equal(st_d.cpu().numpy(), st_o)
        f1, s1, c1, st2 = eng.solve_ticks_device(d_tk, want_solution=True)
        f2, s2, c2 = eng.solve_device(torch.from_numpy(rec_o).cuda(), want_solution=True)
        torch.cuda.synchronize()
        assert np.array_equal(f1.cpu().numpy(), f2.cpu().numpy())
        assert np.array_equal(s1.cpu().numpy(), s2.cpu().numpy())
        assert np.array_equal(c1.cpu().numpy(), c2.cpu().numpy())
        assert np.array_equal(st2.cpu().numpy(), st_o)
        assert (E.status_code(c1.cpu().numpy()) == E.STATUS_OPTIMAL).all()


def test_fused_peer_gather_two_gpus():
    """N>1 on real GPUs: the solve kernel's peer-store epilogue + device-side flag barrier fill every rank's gather
    buffer with exactly what an NCCL all-gather returns (tests/gpu_peer_gather_check.py under torchrun)."""
    import subprocess
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    script = os.path.join(os.path.dirname(os.path.abspath(__file__)), "gpu_peer_gather_check.py")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29531", script], capture_output=True,
                         text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]


def test_gait_state_and_leg_commands_on_device(oracle, cuda_engine_factory):
    """SURVEY 8f rows N2 and N4 on the GPU: OffsetDurationGait's iteration / phase / contact and swing progress /
    contact table, and the reference's use of the solved forces (f_ff = -rBody f, LegController::updateCommand),
    one robot per thread -- bit for bit what the oracle's restatement of the reference's host code gives, also when
    the forces come straight out of a solve on the device."""
    from quadruped_ctrl_b200 import gait as G
    from quadruped_ctrl_b200 import legs as LG
    rng = np.random.default_rng(31)
    recs = []
    for b in range(2000):
        nseg = int(rng.choice([10, 14, 16, 20, 36]))
        name = list(G.GAITS_14)[int(rng.integers(0, len(G.GAITS_14)))]
        off, dur = G.rescale(*G.GAITS_14[name], nseg)
        recs.append(LG.pack_gait_records(13, int(rng.integers(0, 200000)), nseg, off, dur))
    g = np.concatenate(recs)
    eng = cuda_engine_factory(10, 2048)
    st_d, tb_d = eng.gait_state_device(torch.from_numpy(g).cuda(), want_table=True)
    st_o, tb_o = oracle.gait_state(g, want_table=True)
    torch.cuda.synchronize()
    assert np.array_equal(st_d.cpu().numpy().view(np.int32), st_o.view(np.int32))
    assert np.array_equal(tb_d.cpu().numpy(), tb_o)
    # forces out of a solve -> leg commands, everything on the device
    B = 2048
    rec = W.config2(B, 10, 41)
    forces, _, status = eng.solve_device(torch.from_numpy(rec).cuda())
    legs = LG.synth_leg_records(B, 42)
    f_ff, tau = eng.leg_commands_device(torch.from_numpy(legs).cuda(), forces)
    torch.cuda.synchronize()
    assert (E.status_code(status.cpu().numpy()) == E.STATUS_OPTIMAL).all()
    fo, to = oracle.leg_commands(legs, forces.cpu().numpy())
    assert np.array_equal(f_ff.cpu().numpy().view(np.int32), fo.view(np.int32))
    assert np.array_equal(tau.cpu().numpy().view(np.int32), to.view(np.int32))
    assert np.abs(fo).max() > 1.0


def test_warm_start_in_a_closed_loop_rollout(cuda_engine_factory):
    """SURVEY 8f row N3 on the GPU: robots rolled out in closed loop on the MPC's own model (rollout.Rollout), every
    tick solved cold and warm (device-resident per-robot working-set cache, robots permuted inside the batch through
    robot ids).  The warm solve returns the cold optimum to 1e-9 at every tick and needs fewer working-set changes."""
    from quadruped_ctrl_b200 import rollout as RO
    for gait, h, B, ticks, kw in (("trotting", 10, 384, 100, dict(mu=0.15, f_max=52.0)),
                                  ("walking", 10, 192, 40, dict(v_cmd=0.3, f_max=34.0, mu=0.2))):
        ro = RO.Rollout(B, h, gait, 5, **kw)
        eng = cuda_engine_factory(h, B)
        cache = eng.new_warm_cache(B)
        perm = torch.from_numpy(np.random.default_rng(1).permutation(B).astype(np.int32)).cuda()
        it_c, it_w, worst = [], [], 0.0
        for t in range(ticks):
            rec = ro.records()
            d = torch.from_numpy(rec).cuda()
            eng.set_warm_start(None)
            fc, sc_, stc = eng.solve_device(d, want_solution=True)
            # the batch in another order: robot ids tie every problem to its own cache entry
            eng.set_warm_start(cache, robot_ids=perm, shift=1)
            fw, sw, stw = eng.solve_device(d[perm.long()].contiguous(), want_solution=True)
            torch.cuda.synchronize()
            stc, stw = stc.cpu().numpy(), stw.cpu().numpy()
            assert (E.status_code(stc) == 0).all() and (E.status_code(stw) == 0).all()
            worst = max(worst, rel(sw.cpu().numpy(), sc_.cpu().numpy()[perm.cpu().numpy()]).max())
            it_c.append(E.status_iterations(stc).mean())
            it_w.append(E.status_iterations(stw).mean())
            ro.advance(fc.cpu().numpy())
        eng.set_warm_start(None)
        print("\n[%s h=%d, %d robots x %d ticks] working-set additions per solve: cold %.2f, warm %.2f; max |warm - cold| %.1e"
              % (gait, h, B, ticks, np.mean(it_c[5:]), np.mean(it_w[5:]), worst))
        assert worst < 1e-9
        assert np.mean(it_w[5:]) < 0.9 * np.mean(it_c[5:])
