"""Independent numpy/scipy restatement of the reference's dense QP assembly (test infrastructure).

Follows /root/reference/src/MPC_Ctrl/SolverMPC.cpp:87-125 (c2qp), :226-254 (ct_ss_mats), :257-267
(quat_to_rpy), :296-399 (solve_mpc up to qH/qg) and RobotState.cpp:9-43 literally and densely in fp64:
a REAL matrix exponential (scipy.linalg.expm of the 25x25 block matrix, as Eigen's .exp() at :93),
explicit powerMats, dense B_qp, dense S.  It shares no code or algebra with the C oracle (Taylor series)
or the CUDA kernel (closed-form polynomial), so agreement of the three pins the assembly.
"""
import numpy as np
from scipy.linalg import expm


def quat_to_rpy(q):
    w, x, y, z = [float(v) for v in q]
    a = min(-2.0 * (x * z - w * y), 0.99999)
    return np.array([np.arctan2(2 * (x * y + w * z), w * w + x * x - y * y - z * z), np.arcsin(a),
                     np.arctan2(2 * (y * z + w * x), w * w - x * x - y * y + z * z)])


def cross_mat(I_inv, r):
    cm = np.array([[0, -r[2], r[1]], [r[2], 0, -r[0]], [-r[1], r[0], 0]], float)
    return I_inv @ cm


def assemble_dense(f, horizon):
    """f: dict from records.unpack_records for ONE problem.  Returns (qH [12h,12h], qg [12h], x0, A_qp, B_qp)."""
    h = horizon
    yaw = float(f["yaw"])
    c, s = np.cos(yaw), np.sin(yaw)
    R_yaw = np.array([[c, -s, 0], [s, c, 0], [0, 0, 1.0]])
    I_body = np.diag(np.asarray(f["I_body"], float))
    r_feet = np.asarray(f["r"], float).reshape(3, 4)
    rpy = quat_to_rpy(f["q"])
    x0 = np.concatenate([[rpy[2], rpy[1], rpy[0]], np.asarray(f["p"], float), np.asarray(f["w"], float),
                         np.asarray(f["v"], float), [float(np.float32(-9.8))]])
    I_world = R_yaw @ I_body @ R_yaw.T
    A = np.zeros((13, 13))
    A[3, 9] = A[4, 10] = A[5, 11] = 1.0
    A[11, 9] = float(f["x_drag"])
    A[11, 12] = 1.0
    A[0:3, 6:9] = R_yaw.T
    B = np.zeros((13, 12))
    I_inv = np.linalg.inv(I_world)
    for b in range(4):
        B[6:9, 3 * b:3 * b + 3] = cross_mat(I_inv, r_feet[:, b])
        B[9:12, 3 * b:3 * b + 3] = np.eye(3) / float(f["mass"])
    ABc = np.zeros((25, 25))
    ABc[:13, :13] = A
    ABc[:13, 13:] = B
    E = expm(ABc * float(f["dt"]))
    Adt, Bdt = E[:13, :13], E[:13, 13:]
    power = [np.eye(13)]
    for i in range(1, h + 1):
        power.append(Adt @ power[i - 1])
    A_qp = np.zeros((13 * h, 13))
    B_qp = np.zeros((13 * h, 12 * h))
    for r in range(h):
        A_qp[13 * r:13 * r + 13] = power[r + 1]
        for cc in range(r + 1):
            B_qp[13 * r:13 * r + 13, 12 * cc:12 * cc + 12] = power[r - cc] @ Bdt
    full_weight = np.concatenate([np.asarray(f["weights"], float), [0.0]])
    S = np.diag(np.tile(full_weight, h))
    X_d = np.zeros(13 * h)
    traj = np.asarray(f["traj"], float)
    for i in range(h):
        X_d[13 * i:13 * i + 12] = traj[12 * i:12 * i + 12]
    qH = 2 * (B_qp.T @ S @ B_qp + float(f["alpha"]) * np.eye(12 * h))
    qg = 2 * B_qp.T @ S @ (A_qp @ x0 - X_d)
    return qH, qg, x0, A_qp, B_qp


def reduce_qp(qH, qg, gait, f_max):
    """Swing-leg elimination (SolverMPC.cpp:441-525): keep variables of (step,leg) pairs with gait*f_max != ~0."""
    ub = (np.asarray(gait, np.float32) * np.float32(f_max)).astype(np.float64)
    keep_pair = ~((ub < 0.01) & (ub > -0.01))
    keep = np.repeat(keep_pair, 3)
    idx = np.nonzero(keep)[0]
    return qH[np.ix_(idx, idx)], qg[idx], idx
