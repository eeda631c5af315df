"""Python mirror of the reference's MPC C interface (src/MPC_Ctrl/convexMPC_interface.h:40-48).

Same function names, argument order and meaning as the reference; each call goes straight
through the C ABI of libquadruped_mpc_b200.so (include/convexMPC_interface.h), i.e. the
same symbols ConvexMPCLocomotion::solveDenseMPC binds (ConvexMPCLocomotion.cpp:630-674).
"""
import ctypes

import numpy as np

from . import engine as _E


def _fp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def _dp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def setup_problem(dt, horizon, mu, f_max):
    _E.lib().setup_problem(float(dt), int(horizon), float(mu), float(f_max))


def update_solver_settings(max_iter, rho, sigma, solver_alpha, terminate, use_jcqp):
    _E.lib().update_solver_settings(int(max_iter), float(rho), float(sigma), float(solver_alpha), float(terminate),
                                    float(use_jcqp))


def update_x_drag(x_drag):
    _E.lib()._Z13update_x_dragf(float(x_drag))


def update_problem_data_floats(p, v, q, w, r, yaw, weights, state_trajectory, alpha, gait):
    a = [np.ascontiguousarray(x, np.float32) for x in (p, v, q, w, r, weights, state_trajectory)]
    g = np.ascontiguousarray(gait, np.int32)
    _E.lib().update_problem_data_floats(_fp(a[0]), _fp(a[1]), _fp(a[2]), _fp(a[3]), _fp(a[4]), float(yaw), _fp(a[5]),
                                        _fp(a[6]), float(alpha), g.ctypes.data_as(ctypes.POINTER(ctypes.c_int)))


def update_problem_data(p, v, q, w, r, yaw, weights, state_trajectory, alpha, gait):
    a = [np.ascontiguousarray(x, np.float64) for x in (p, v, q, w, r, weights, state_trajectory)]
    g = np.ascontiguousarray(gait, np.int32)
    _E.lib().update_problem_data(_dp(a[0]), _dp(a[1]), _dp(a[2]), _dp(a[3]), _dp(a[4]), float(yaw), _dp(a[5]),
                                 _dp(a[6]), float(alpha), g.ctypes.data_as(ctypes.POINTER(ctypes.c_int)))


def get_solution(index):
    return float(_E.lib().get_solution(int(index)))


def last_status():
    return int(_E.lib().mpc_last_status())


def last_iterations():
    return int(_E.lib().mpc_last_iterations())


def set_robot(I_body_diag, mass):
    a = np.ascontiguousarray(I_body_diag, np.float32)
    _E.lib().mpc_set_robot(_fp(a), float(mass))


def shutdown():
    _E.lib().mpc_shutdown()


def legacy_record(horizon):
    """The inputs last handed to the legacy calls, as one packed batch record (uint8 [stride])."""
    from . import records as R
    out = np.zeros(R.record_stride(horizon), np.uint8)
    h = _E.lib().mpc_legacy_record(out.ctypes.data)
    if h != horizon:
        raise _E.MpcError("mpc_legacy_record: horizon is %d, expected %d" % (h, horizon))
    return out
