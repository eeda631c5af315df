"""The oracle against its pins (CPU tier).

* golden fixture: outputs of the reference's own qpOASES run in the build container
  (tests/golden/make_golden.py); re-running the oracle must reproduce them;
* the oracle's independent active-set port against the reference qpOASES on the same QPs;
* the oracle's assembly (Taylor series of the nilpotent block matrix) against an independent dense
  numpy/scipy restatement that uses a real matrix exponential.
"""
import numpy as np
import pytest

from quadruped_ctrl_b200 import records as R
from quadruped_ctrl_b200 import workloads as W

from common import GOLDEN_CASES, load_golden, rel
import np_reference as NP


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_port_backend_reproduces_golden(name, oracle):
    """The 'port' backend (oracle's own solver, available everywhere) lands on the reference's answers."""
    G = load_golden()
    rec, h = G[name + "_records"], int(G[name + "_h"])
    for prec, tag in ((32, "o32"), (64, "o64")):
        o = oracle.solve_batch(rec, h, prec, "port")
        ok = G["%s_%s_rc" % (name, tag)] == 0  # problems the reference itself solved (nWSR <= 100)
        assert (o["nv"] == G["%s_%s_nv" % (name, tag)]).all()
        e = rel(o["sol"], G["%s_%s_sol" % (name, tag)])
        # fp64-assembled QPs: two exact solvers agree to round-off.  fp32-assembled QPs are not exactly
        # symmetric (qH(i,j) and qH(j,i) round differently) and the two solvers read different triangles,
        # so they agree only to the fp32 noise amplified by the conditioning of H.
        assert e[ok].max() < (2e-4 if prec == 32 else 1e-9), (name, tag, e.max())


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_reference_backend_reproduces_golden(name, oracle):
    if not oracle.have_reference_qpoases():
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    G = load_golden()
    rec, h = G[name + "_records"], int(G[name + "_h"])
    for prec, tag in ((32, "o32"), (64, "o64")):
        o = oracle.solve_batch(rec, h, prec, "reference")
        assert (o["rc"] == G["%s_%s_rc" % (name, tag)]).all()
        assert (o["nwsr"] == G["%s_%s_nwsr" % (name, tag)]).all()
        assert rel(o["sol"], G["%s_%s_sol" % (name, tag)]).max() < 1e-12


def test_assembly_against_dense_numpy_expm(oracle):
    """Three-way pin of the assembly: C oracle (fp64) == dense numpy with scipy.linalg.expm."""
    for name, B in (("config2", 6), ("config3", 4), ("edge", None)):
        if name == "edge":
            G = load_golden()
            rec, h = G["edge_records"], 10
        else:
            h = W.HORIZONS[name]
            rec = W.CONFIGS[name](B)
        o = oracle.solve_batch(rec, h, 64, "assemble", want_qp=True)
        f = R.unpack_records(rec, h)
        for b in range(rec.shape[0]):
            one = {k: v[b] for k, v in f.items()}
            qH, qg, _, _, _ = NP.assemble_dense(one, h)
            Hr, gr, idx = NP.reduce_qp(qH, qg, one["gait"], one["f_max"])
            nv = len(idx)
            assert nv == o["nv"][b]
            assert np.abs(o["H"][b][:nv, :nv] - Hr).max() <= 1e-11 * np.abs(Hr).max()
            assert np.abs(o["g"][b][:nv] - gr).max() <= 1e-11 * np.abs(gr).max()


def test_legacy_abi_mirror_matches_batch_path(oracle):
    """oracle_setup_problem / update_problem_data_floats / get_solution == the batched entry on config 1."""
    import ctypes
    L = oracle.lib()
    rec = W.config1()
    h = 10
    f = R.unpack_records(rec, h)
    ref = oracle.solve_batch(rec, h, 32)
    L.oracle_configure(32, 0 if oracle.have_reference_qpoases() else 1)
    fp = ctypes.POINTER(ctypes.c_float)
    for b in range(rec.shape[0]):
        L.oracle_setup_problem(float(f["dt"][b]), h, float(f["mu"][b]), float(f["f_max"][b]))
        L.oracle_update_x_drag(float(f["x_drag"][b]))
        arrs = [np.ascontiguousarray(f[k][b], np.float32) for k in ("p", "v", "q", "w", "r", "weights", "traj")]
        gait = np.ascontiguousarray(f["gait"][b], np.int32)
        L.oracle_update_problem_data_floats(arrs[0].ctypes.data_as(fp), arrs[1].ctypes.data_as(fp),
                                            arrs[2].ctypes.data_as(fp), arrs[3].ctypes.data_as(fp),
                                            arrs[4].ctypes.data_as(fp), float(f["yaw"][b]), arrs[5].ctypes.data_as(fp),
                                            arrs[6].ctypes.data_as(fp), float(f["alpha"][b]),
                                            gait.ctypes.data_as(ctypes.POINTER(ctypes.c_int)))
        got = np.array([L.oracle_get_solution(i) for i in range(12)])
        assert np.allclose(got, ref["forces"][b], rtol=0, atol=1e-12)


def test_fp32_cloud_is_what_the_survey_measured(oracle):
    """The reference's own fp32 rounding cloud on config 2 stays below the 1e-4 parity target."""
    rec = W.config2(256)
    o32 = oracle.solve_batch(rec, 10, 32)
    o64 = oracle.solve_batch(rec, 10, 64)
    assert rel(o32["forces"], o64["forces"]).max() < 1e-4
