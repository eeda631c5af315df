"""ctypes front-end of the CPU oracle (TEST INFRASTRUCTURE ONLY).

liboracle.so  = oracle/mpc_oracle.cpp (restatement of SolverMPC.cpp:296-557) +
                oracle/qp_port.cpp (independent dense active-set QP solver)
_ref/libqpoases_ref.so = the reference's own qpOASES, built from /root/reference.

backend "reference" -> reference qpOASES solves the reduced QP (kind "reference");
backend "port"      -> qp_port.cpp solves it (kind "port").
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    """Compiles liboracle.so and, when /root/reference is present, _ref/libqpoases_ref.so."""
    need = force or not os.path.exists(os.path.join(_HERE, "liboracle.so"))
    src_t = max(os.path.getmtime(os.path.join(_HERE, f)) for f in ("mpc_oracle.cpp", "qp_port.cpp", "tick_oracle.cpp", "leg_oracle.cpp",
                                                                   "Makefile"))
    if not need and os.path.getmtime(os.path.join(_HERE, "liboracle.so")) < src_t:
        need = True
    ref_missing = os.path.isdir("/root/reference/src/qpOASES") and not os.path.exists(
        os.path.join(_HERE, "_ref", "libqpoases_ref.so"))
    if need or ref_missing:
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))


def lib():
    global _LIB
    if _LIB is None:
        build()
        L = ctypes.CDLL(os.path.join(_HERE, "liboracle.so"))
        L.oracle_load_qpoases.argtypes = [ctypes.c_char_p]
        L.oracle_load_qpoases.restype = ctypes.c_int
        L.oracle_record_stride.argtypes = [ctypes.c_int]
        L.oracle_record_stride.restype = ctypes.c_size_t
        L.oracle_solve_batch.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                         ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                         ctypes.c_void_p]
        L.oracle_solve_batch.restype = None
        L.oracle_setup_problem.argtypes = [ctypes.c_double, ctypes.c_int, ctypes.c_double, ctypes.c_double]
        L.oracle_update_x_drag.argtypes = [ctypes.c_float]
        fp = ctypes.POINTER(ctypes.c_float)
        L.oracle_update_problem_data_floats.argtypes = [fp, fp, fp, fp, fp, ctypes.c_float, fp, fp, ctypes.c_float,
                                                        ctypes.POINTER(ctypes.c_int)]
        L.oracle_get_solution.argtypes = [ctypes.c_int]
        L.oracle_get_solution.restype = ctypes.c_double
        L.oracle_configure.argtypes = [ctypes.c_int, ctypes.c_int]
        L.oracle_build_records.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                           ctypes.c_void_p]
        L.oracle_build_records.restype = None
        L.oracle_gait_state.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
        L.oracle_gait_state.restype = None
        L.oracle_leg_commands.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p,
                                          ctypes.c_void_p]
        L.oracle_leg_commands.restype = None
        L.oracle_load_qpoases(os.path.join(_HERE, "_ref", "libqpoases_ref.so").encode())
        _LIB = L
    return _LIB


def have_reference_qpoases():
    return bool(lib().oracle_have_qpoases())


def default_backend():
    return "reference" if have_reference_qpoases() else "port"


_BACKENDS = {"reference": 0, "port": 1, "assemble": -1}


def solve_batch(records, horizon, precision=32, backend=None, want_qp=False):
    """records: uint8 [B, stride] packed problem records (include/mpc_batch.h).

    Returns dict(sol [B,12h] f64, forces [B,12] f64, nv, nc, nwsr, rc, rc_primal, obj
    [, H [B,12h,12h], g [B,12h]])."""
    L = lib()
    backend = backend or default_backend()
    if backend == "reference" and not have_reference_qpoases():
        raise RuntimeError("oracle/_ref/libqpoases_ref.so is not built (needs /root/reference)")
    records = np.ascontiguousarray(records, dtype=np.uint8)
    B = records.shape[0]
    assert records.shape[1] == L.oracle_record_stride(horizon), (records.shape, L.oracle_record_stride(horizon))
    NU = 12 * horizon
    sol = np.zeros((B, NU), np.float64)
    info = np.zeros((B, 5), np.int32)
    obj = np.zeros(B, np.float64)
    H = np.zeros((B, NU, NU), np.float64) if want_qp else None
    g = np.zeros((B, NU), np.float64) if want_qp else None
    L.oracle_solve_batch(records.ctypes.data, B, horizon, precision, _BACKENDS[backend], sol.ctypes.data,
                         info.ctypes.data, obj.ctypes.data, H.ctypes.data if want_qp else None,
                         g.ctypes.data if want_qp else None)
    out = dict(sol=sol, forces=sol[:, :12].copy(), nv=info[:, 0], nc=info[:, 1], nwsr=info[:, 2], rc=info[:, 3],
               rc_primal=info[:, 4], obj=obj)
    if want_qp:
        out["H"], out["g"] = H, g
    return out


def _worker(args):
    records, horizon, precision, backend = args
    return solve_batch(records, horizon, precision, backend)["forces"]


def solve_batch_parallel(records, horizon, precision=32, backend=None, workers=None):
    """Same as solve_batch()['forces'] over `workers` forked processes (the reference solver
    keeps process-global state -- convexMPC_interface.cpp:13-20, qpOASES' global message
    handler -- so parallelism is per process, one solver instance each)."""
    import multiprocessing as mp
    workers = workers or os.cpu_count() or 1
    lib()
    chunks = [c for c in np.array_split(np.ascontiguousarray(records), workers) if len(c)]
    if len(chunks) <= 1:
        return solve_batch(records, horizon, precision, backend)["forces"]
    with mp.get_context("fork").Pool(len(chunks)) as pool:
        parts = pool.map(_worker, [(c, horizon, precision, backend) for c in chunks])
    return np.concatenate(parts, 0)


def build_records(ticks, horizon):
    """ticks: float32-viewable [B, 68] tick records (include/mpc_batch.h).  Returns (records uint8 [B, stride],
    state_out float32 [B, 4]) as the reference's host code would produce them."""
    L = lib()
    ticks = np.ascontiguousarray(ticks).view(np.float32).reshape(-1, 68)
    B = ticks.shape[0]
    stride = L.oracle_record_stride(horizon)
    rec = np.zeros((B, stride), np.uint8)
    st = np.zeros((B, 4), np.float32)
    L.oracle_build_records(ticks.ctypes.data, B, horizon, rec.ctypes.data, st.ctypes.data)
    return rec, st


def gait_state(gait, want_table=False):
    """gait: int32 [B, 12] gait records (include/mpc_batch.h).  Returns (state float32 [B, 10] (word 0 an int32),
    tables uint8 [B, 4*max nIterations] | None) as the reference's OffsetDurationGait would produce them."""
    L = lib()
    gait = np.ascontiguousarray(gait, np.int32).reshape(-1, 12)
    B = gait.shape[0]
    state = np.zeros((B, 10), np.float32)
    stride = 4 * int(gait[:, 2].max()) if want_table else 0
    tables = np.zeros((B, stride), np.uint8) if want_table else None
    L.oracle_gait_state(gait.ctypes.data, B, state.ctypes.data, tables.ctypes.data if want_table else None, stride)
    return state, tables


def leg_commands(legs, forces):
    """legs: float32-viewable [B, 100] leg records, forces float32 [B, 12].  Returns (f_ff [B,12], tau [B,12])."""
    L = lib()
    legs = np.ascontiguousarray(legs).view(np.float32).reshape(-1, 100)
    forces = np.ascontiguousarray(forces, np.float32).reshape(-1, 12)
    B = legs.shape[0]
    f_ff = np.zeros((B, 12), np.float32)
    tau = np.zeros((B, 12), np.float32)
    L.oracle_leg_commands(legs.ctypes.data, forces.ctypes.data, B, f_ff.ctypes.data, tau.ctypes.data)
    return f_ff, tau
