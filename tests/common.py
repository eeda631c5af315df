"""Shared helpers for the test suite."""
import ctypes
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden", "mpc_golden.npz")
GOLDEN_CASES = ["config1", "config2", "config3", "config4", "config5", "four_stance", "edge"]


def rel(a, b, floor=1.0):
    """Per-problem relative L2 error with an absolute floor of `floor` newtons on the norm (problems whose
    optimum is ~0 N -- every stance leg at the apex of its cone -- have no meaningful relative error)."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return np.linalg.norm(a - b, axis=1) / np.maximum(np.linalg.norm(b, axis=1), floor)


def load_golden():
    return np.load(GOLDEN)


_EMU = None


def emu_lib():
    """TEST-ONLY single-thread host build of the kernel source (tests/emu/emu.cpp)."""
    global _EMU
    if _EMU is None:
        src = os.path.join(ROOT, "tests", "emu", "emu.cpp")
        so = os.path.join(ROOT, "tests", "emu", "libmpc_emu.so")
        core = os.path.join(ROOT, "quadruped_ctrl_b200", "csrc", "mpc_core.h")
        ticks = os.path.join(ROOT, "quadruped_ctrl_b200", "csrc", "mpc_ticks.h")
        legs = os.path.join(ROOT, "quadruped_ctrl_b200", "csrc", "mpc_legs.h")
        ric = os.path.join(ROOT, "quadruped_ctrl_b200", "csrc", "mpc_riccati.h")
        if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(core),
                                                                 os.path.getmtime(ticks), os.path.getmtime(legs),
                                                                 os.path.getmtime(ric)):
            subprocess.check_call(["g++", "-O2", "-std=c++17", "-Wall", "-Wno-unknown-pragmas", "-ffp-contract=off", "-fPIC", "-shared", src, "-o", so])
        _EMU = ctypes.CDLL(so)
    return _EMU


def emu_solve(rec, h, nv_cap=0, m_cap=0, want_qp=False, max_iter=100000):
    L = emu_lib()
    rec = np.ascontiguousarray(rec, np.uint8)
    B = rec.shape[0]
    NU = 12 * h
    f = np.zeros((B, 12), np.float32)
    sol = np.zeros((B, NU))
    info = np.zeros((B, 4), np.int32)
    H = np.zeros((B, NU, NU)) if want_qp else None
    g = np.zeros((B, NU)) if want_qp else None
    vp = ctypes.c_void_p
    rc = L.emu_solve_batch(vp(rec.ctypes.data), B, h, nv_cap, m_cap, max_iter, vp(f.ctypes.data), vp(sol.ctypes.data),
                           vp(info.ctypes.data), vp(H.ctypes.data) if want_qp else None,
                           vp(g.ctypes.data) if want_qp else None)
    assert rc == 0, rc
    return dict(forces=f, sol=sol, nv=info[:, 0], m=info[:, 1], iters=info[:, 2], status=info[:, 3], H=H, g=g)


def emu_build_records(ticks, h):
    """Host build of the device-side tick -> record builder."""
    from quadruped_ctrl_b200 import records as R
    L = emu_lib()
    ticks = np.ascontiguousarray(ticks).view(np.float32).reshape(-1, 68)
    B = ticks.shape[0]
    rec = np.full((B, R.record_stride(h)), 0xCD, np.uint8)
    st = np.zeros((B, 4), np.float32)
    L.emu_build_records(ctypes.c_void_p(ticks.ctypes.data), B, h, ctypes.c_void_p(rec.ctypes.data),
                        ctypes.c_void_p(st.ctypes.data))
    return rec, st


def emu_gait_state(gait, want_table=False):
    """Host build of the device-side gait-state body (csrc/mpc_legs.h)."""
    L = emu_lib()
    gait = np.ascontiguousarray(gait, np.int32).reshape(-1, 12)
    B = gait.shape[0]
    state = np.full((B, 10), np.nan, np.float32)
    stride = 4 * int(gait[:, 2].max()) if want_table else 0
    tables = np.zeros((B, stride), np.uint8) if want_table else None
    L.emu_gait_state(ctypes.c_void_p(gait.ctypes.data), B, ctypes.c_void_p(state.ctypes.data),
                     ctypes.c_void_p(tables.ctypes.data) if want_table else None, stride)
    return state, tables


def emu_leg_commands(legs, forces):
    """Host build of the device-side leg-command body (csrc/mpc_legs.h)."""
    L = emu_lib()
    legs = np.ascontiguousarray(legs).view(np.float32).reshape(-1, 100)
    forces = np.ascontiguousarray(forces, np.float32).reshape(-1, 12)
    B = legs.shape[0]
    f_ff = np.full((B, 12), np.nan, np.float32)
    tau = np.full((B, 12), np.nan, np.float32)
    L.emu_leg_commands(ctypes.c_void_p(legs.ctypes.data), ctypes.c_void_p(forces.ctypes.data), B,
                       ctypes.c_void_p(f_ff.ctypes.data), ctypes.c_void_p(tau.ctypes.data))
    return f_ff, tau


def emu_solve_warm(rec, h, cache, shift=1, nv_cap=0, m_cap=0, max_iter=100000):
    """Host build of the kernel body with the warm start: cache int32 [B, 128] is read and rewritten in place."""
    L = emu_lib()
    rec = np.ascontiguousarray(rec, np.uint8)
    B = rec.shape[0]
    assert cache.dtype == np.int32 and cache.shape == (B, L.emu_warm_stride()) and cache.flags.c_contiguous
    f = np.zeros((B, 12), np.float32)
    sol = np.zeros((B, 12 * h))
    info = np.zeros((B, 4), np.int32)
    vp = ctypes.c_void_p
    rc = L.emu_solve_batch_warm(vp(rec.ctypes.data), B, h, nv_cap, m_cap, max_iter, vp(f.ctypes.data),
                                vp(sol.ctypes.data), vp(info.ctypes.data), vp(cache.ctypes.data), int(shift))
    assert rc == 0, rc
    return dict(forces=f, sol=sol, nv=info[:, 0], m=info[:, 1], iters=info[:, 2], status=info[:, 3])


def emu_solve_wrench(rec, h, m_cap=0, max_iter=100000):
    """Host build of the wrench-space class (H^{-1} through the rank-6h structure of the Hessian)."""
    L = emu_lib()
    rec = np.ascontiguousarray(rec, np.uint8)
    B = rec.shape[0]
    f = np.zeros((B, 12), np.float32)
    sol = np.zeros((B, 12 * h))
    info = np.zeros((B, 4), np.int32)
    vp = ctypes.c_void_p
    rc = L.emu_solve_batch_wrench(vp(rec.ctypes.data), B, h, m_cap, max_iter, vp(f.ctypes.data), vp(sol.ctypes.data),
                                  vp(info.ctypes.data))
    assert rc == 0, rc
    return dict(forces=f, sol=sol, nv=info[:, 0], m=info[:, 1], iters=info[:, 2], status=info[:, 3])


def emu_solve_riccati(rec, h, nv_cap=0, m_cap=0, max_iter=100000, with_slab=False):
    """Host build of the Riccati solver (csrc/mpc_riccati.h): no condensed Hessian, H^{-1} products by sweeps."""
    L = emu_lib()
    rec = np.ascontiguousarray(rec, np.uint8)
    B = rec.shape[0]
    f = np.zeros((B, 12), np.float32)
    sol = np.zeros((B, 12 * h))
    info = np.zeros((B, 4), np.int32)
    vp = ctypes.c_void_p
    rc = L.emu_solve_batch_riccati(vp(rec.ctypes.data), B, h, nv_cap, m_cap, max_iter, vp(f.ctypes.data),
                                   vp(sol.ctypes.data), vp(info.ctypes.data), int(bool(with_slab)))
    assert rc == 0, rc
    return dict(forces=f, sol=sol, nv=info[:, 0], m=info[:, 1], iters=info[:, 2], status=info[:, 3])
