"""The C-ABI boundary (CPU tier): struct layout, exported symbols, loud failure without a GPU."""
import ctypes
import os
import re
import subprocess
import tempfile

import numpy as np
import pytest

from quadruped_ctrl_b200 import engine as E
from quadruped_ctrl_b200 import interface as I

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INC = os.path.join(ROOT, "include")


@pytest.fixture(scope="module")
def lib():
    E.build()
    return E.lib()


def declared_functions(header):
    """Names of every function a header declares (prototype lines ending in ';' with a parenthesis)."""
    txt = open(os.path.join(INC, header)).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    txt = re.sub(r"//.*", "", txt)
    names = []
    for m in re.finditer(r"([A-Za-z_][A-Za-z_0-9]*)\s*\(([^;{}()]|\([^()]*\))*\)\s*;", txt):
        n = m.group(1)
        if n not in ("defined", "sizeof", "MPC_STATUS_CODE", "MPC_STATUS_ITERS"):
            names.append(n)
    return names


def test_every_declared_symbol_is_exported(lib):
    out = subprocess.check_output(["nm", "-D", "--defined-only", E.LIB_PATH], text=True)
    exported = {line.split()[-1] for line in out.splitlines() if line.strip()}
    batch = declared_functions("mpc_batch.h")
    legacy = declared_functions("convexMPC_interface.h")
    assert len(batch) >= 15 and len(legacy) >= 8
    for n in batch:
        assert n in exported, n
    for n in legacy:
        # update_x_drag is declared outside EXTERNC upstream (convexMPC_interface.h:48): C++ linkage
        assert (n in exported) or (n == "update_x_drag" and "_Z13update_x_dragf" in exported), n
    assert set(E.BATCH_SYMBOLS) <= exported and set(E.LEGACY_SYMBOLS) <= exported


def test_struct_layout_is_the_reference_layout():
    """Field offsets of problem_setup / update_data_t as a C compiler sees OUR header must equal the
    layout of the reference header (convexMPC_interface.h:13-38), written out here by hand."""
    src = r'''
#include <stdio.h>
#include <stddef.h>
#include "convexMPC_interface.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu\n", sizeof(struct problem_setup), offsetof(struct problem_setup, dt),
         offsetof(struct problem_setup, mu), offsetof(struct problem_setup, f_max), offsetof(struct problem_setup, horizon));
  printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(struct update_data_t),
         offsetof(struct update_data_t, p), offsetof(struct update_data_t, v), offsetof(struct update_data_t, q),
         offsetof(struct update_data_t, w), offsetof(struct update_data_t, r), offsetof(struct update_data_t, yaw),
         offsetof(struct update_data_t, weights), offsetof(struct update_data_t, traj), offsetof(struct update_data_t, alpha),
         offsetof(struct update_data_t, gait), offsetof(struct update_data_t, hack_pad),
         offsetof(struct update_data_t, max_iterations), offsetof(struct update_data_t, rho),
         offsetof(struct update_data_t, sigma), offsetof(struct update_data_t, solver_alpha),
         offsetof(struct update_data_t, terminate), offsetof(struct update_data_t, use_jcqp),
         offsetof(struct update_data_t, x_drag));
  return 0;
}'''
    ref_dir = "/root/reference/src/MPC_Ctrl"
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "t.c")
        open(c, "w").write(src)
        exe = os.path.join(d, "t")
        subprocess.check_call(["gcc", "-I", INC, c, "-o", exe])
        l1, l2 = subprocess.check_output([exe], text=True).splitlines()
        if os.path.exists(os.path.join(ref_dir, "convexMPC_interface.h")):
            # the reference is mounted (build container): compile the SAME probe against ITS header (no Eigen in
            # that file; read in place, nothing copied) and compare the two compilers' views line for line
            exe_ref = os.path.join(d, "t_ref")
            subprocess.check_call(["g++", "-x", "c++", "-I", ref_dir, c, "-o", exe_ref])
            r1, r2 = subprocess.check_output([exe_ref], text=True).splitlines()
            assert (l1, l2) == (r1, r2)
    assert [int(x) for x in l1.split()] == [16, 0, 4, 8, 12]
    # floats: p0 v12 q24 w40 r52 yaw100 weights104 traj152 (+12*36*4=1728) alpha1880; u8 gait1884 (+36) pad1920 (+1000)
    # -> 2920; int max_iterations 2920; doubles rho 2928, sigma 2936, solver_alpha 2944, terminate 2952;
    # int use_jcqp 2960; float x_drag 2964; sizeof 2968
    assert [int(x) for x in l2.split()] == [2968, 0, 12, 24, 40, 52, 100, 104, 152, 1880, 1884, 1920, 2920, 2928,
                                            2936, 2944, 2952, 2960, 2964]


def test_record_layout_helpers(lib):
    from quadruped_ctrl_b200 import records as R
    for h in (1, 9, 10, 14, 16, 20, 36):
        assert lib.mpc_record_stride(h) == R.record_stride(h)
        assert lib.mpc_record_gait_offset(h) == R.gait_offset(h)
        assert R.record_stride(h) % 16 == 0 and R.record_stride(h) >= R.gait_offset(h) + 4 * h


def test_no_gpu_means_a_loud_failure_not_a_cpu_answer(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(E.MpcError):
        E.MpcBatch(10, 8)
    # the legacy entry reports on stderr, keeps get_solution() at 0 and the status negative
    I.setup_problem(0.026, 10, 0.4, 120.0)
    I.update_solver_settings(10000, 1e-7, 1e-8, 1.5, 0.1, 0.0)
    I.update_x_drag(0.0)
    I.update_problem_data_floats(np.zeros(3), np.zeros(3), [1, 0, 0, 0], np.zeros(3), np.zeros(12), 0.0, np.ones(12),
                                 np.zeros(120), 4e-5, np.ones(40))
    assert I.last_status() < 0
    assert all(I.get_solution(i) == 0.0 for i in range(12))


def test_bad_arguments_are_rejected(lib):
    h = ctypes.c_void_p()
    assert lib.mpc_batch_create(ctypes.byref(h), 0, 0, 16) == E.MPC_E_ARG
    assert lib.mpc_batch_create(ctypes.byref(h), 0, 37, 16) == E.MPC_E_ARG
    assert lib.mpc_batch_create(ctypes.byref(h), 0, 10, 0) == E.MPC_E_ARG
    assert lib.mpc_batch_create(None, 0, 10, 16) == E.MPC_E_ARG
    assert b"bad argument" in lib.mpc_last_error()


def test_product_never_touches_the_oracle():
    """No file of the product package imports, links or executes anything under oracle/."""
    pkg = os.path.join(ROOT, "quadruped_ctrl_b200")
    for dp, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cpp", ".h", "Makefile")):
                txt = open(os.path.join(dp, fn), errors="ignore").read()
                assert "oracle" not in txt.lower().replace("tests/emu", ""), os.path.join(dp, fn)
    out = subprocess.check_output(["ldd", E.LIB_PATH], text=True)
    assert "oracle" not in out and "qpoases" not in out.lower()


def test_legacy_calls_pack_the_same_record_as_the_batch_packer(lib):
    """setup_problem / update_x_drag / update_problem_data_floats -> mpc_legacy_record must reproduce
    records.pack_records byte for byte, including horizons whose 4h gait bytes run on into hack_pad
    (h > 9, convexMPC_interface.h:32-33) -- needs no GPU."""
    import contextlib
    import io
    from quadruped_ctrl_b200 import records as R
    from quadruped_ctrl_b200 import workloads as W
    for rec, h in ((W.config1(), 10), (W.config5(4), 16), (W.config3(4), 20), (W.config2(3, 36, 1), 36)):
        f = R.unpack_records(rec, h)
        for b in range(rec.shape[0]):
            I.set_robot(f["I_body"][b], float(f["mass"][b]))
            I.setup_problem(float(f["dt"][b]), h, float(f["mu"][b]), float(f["f_max"][b]))
            I.update_x_drag(float(f["x_drag"][b]))
            I.update_problem_data_floats(f["p"][b], f["v"][b], f["q"][b], f["w"][b], f["r"][b], float(f["yaw"][b]),
                                         f["weights"][b], f["traj"][b], float(f["alpha"][b]),
                                         f["gait"][b].astype(np.int32))
            assert (I.legacy_record(h) == rec[b]).all(), (h, b)
    I.set_robot([0.07, 0.26, 0.242], 9.0)


def _build_caller_stub():
    """g++-compiles tests/caller/solve_dense_mpc_stub.cpp against include/ and links it to the product library."""
    src = os.path.join(ROOT, "tests", "caller", "solve_dense_mpc_stub.cpp")
    exe = os.path.join(tempfile.gettempdir(), "solve_dense_mpc_stub_%d" % os.getpid())
    libdir = os.path.dirname(E.LIB_PATH)
    subprocess.check_call(["g++", "-std=c++14", "-Wall", "-Wextra", "-Werror", "-I", INC, src, "-o", exe,
                           "-L", libdir, "-l:" + os.path.basename(E.LIB_PATH), "-Wl,-rpath," + libdir])
    return exe


def test_cpp_caller_with_the_reference_call_sequence_links_and_runs(lib):
    """SURVEY 8b link test: a C++14 caller (the reference's flags: -Wall -Wextra -Werror) issuing the call sequence
    of ConvexMPCLocomotion::solveDenseMPC compiles against our header, links against the library -- including the
    C++-linkage update_x_drag -- and runs.  Without a GPU every solve must fail loudly (status < 0, forces 0)."""
    import torch
    exe = _build_caller_stub()
    try:
        r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    finally:
        os.unlink(exe)
    assert r.returncode == 0, r.stderr
    lines = [l.split() for l in r.stdout.splitlines() if l.startswith("h ")]
    assert [int(l[1]) for l in lines] == [10, 14, 10]
    if not torch.cuda.is_available():
        for l in lines:
            assert int(l[3]) < 0 and all(float(x) == 0.0 for x in l[5:17])
        assert "no" in r.stderr.lower() or "cuda" in r.stderr.lower()
