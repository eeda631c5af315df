"""Feasibility study (numpy, CPU, test infrastructure): Goldfarb-Idnani dual active set in which every product
H^{-1} v comes from a Riccati recursion over the horizon instead of an explicit inverse of the condensed Hessian.
Compared against the oracle's reference-qpOASES solution of the fp64-assembled dense QP.

  python tests/riccati_study.py config2 64
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle as O  # noqa: E402
from quadruped_ctrl_b200 import records as R  # noqa: E402
from quadruped_ctrl_b200 import workloads as W  # noqa: E402
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import np_reference as NP  # noqa: E402


def dynamics(f):
    """A (13x13), B (13x12) discretised, Q (13), x0 (13) of one problem (closed form: A_c is nilpotent)."""
    yaw = float(f["yaw"])
    c, s = np.cos(yaw), np.sin(yaw)
    Ry = np.array([[c, -s, 0], [s, c, 0], [0, 0, 1.0]])
    Ib = np.diag(np.asarray(f["I_body"], float))
    r = np.asarray(f["r"], float).reshape(3, 4)
    rpy = NP.quat_to_rpy(f["q"])
    x0 = np.concatenate([[rpy[2], rpy[1], rpy[0]], np.asarray(f["p"], float), np.asarray(f["w"], float),
                         np.asarray(f["v"], float), [float(np.float32(-9.8))]])
    Iw = Ry @ Ib @ Ry.T
    A = np.zeros((13, 13))
    A[3, 9] = A[4, 10] = A[5, 11] = 1.0
    A[11, 9] = float(f["x_drag"])
    A[11, 12] = 1.0
    A[0:3, 6:9] = Ry.T
    B = np.zeros((13, 12))
    Ii = np.linalg.inv(Iw)
    for b in range(4):
        B[6:9, 3 * b:3 * b + 3] = NP.cross_mat(Ii, r[:, b])
        B[9:12, 3 * b:3 * b + 3] = np.eye(3) / float(f["mass"])
    dt = float(f["dt"])
    Ad = np.eye(13) + dt * A + dt * dt / 2 * A @ A
    Bd = dt * B + dt * dt / 2 * A @ B + dt ** 3 / 6 * A @ A @ B
    Q = np.concatenate([np.asarray(f["weights"], float), [0.0]])
    return Ad, Bd, Q, x0


class Riccati:
    def __init__(self, Ad, Bd, Q, alpha, stance, h):
        """stance[k] = list of stance legs at step k."""
        self.A, self.h, self.Q = Ad, h, Q
        self.cols = [np.concatenate([[3 * l, 3 * l + 1, 3 * l + 2] for l in st]).astype(int) if len(st) else
                     np.zeros(0, int) for st in stance]
        self.B = [Bd[:, c] for c in self.cols]
        self.off = np.concatenate([[0], np.cumsum([len(c) for c in self.cols])])
        self.nv = int(self.off[-1])
        P = np.diag(Q)
        self.K, self.Sinv = [None] * h, [None] * h
        for k in range(h - 1, -1, -1):
            Bk = self.B[k]
            S = alpha * np.eye(Bk.shape[1]) + Bk.T @ P @ Bk
            Si = np.linalg.inv(S)
            K = Si @ Bk.T @ P @ Ad
            self.K[k], self.Sinv[k] = K, Si
            P = Ad.T @ P @ Ad - Ad.T @ P @ Bk @ K
            P = 0.5 * (P + P.T)
            if k >= 1:
                P = P + np.diag(Q)

    def solve(self, x0, xd, lin):
        """argmin u'(B'SB + alpha I)u + 2u'B'S(A x0 - xd) + lin'u ; xd [h,13] (row k-1 = x_d of step k) or None."""
        h, A = self.h, self.A
        p = np.zeros(13) if xd is None else -self.Q * xd[h - 1]
        kap = [None] * h
        for k in range(h - 1, -1, -1):
            Bk = self.B[k]
            w = Bk.T @ p + 0.5 * lin[self.off[k]:self.off[k + 1]]
            kap[k] = self.Sinv[k] @ w
            p = A.T @ p - self.K[k].T @ w
            if k >= 1 and xd is not None:
                p = p - self.Q * xd[k - 1]
        x = x0.copy()
        u = np.zeros(self.nv)
        for k in range(h):
            uk = -self.K[k] @ x - kap[k]
            u[self.off[k]:self.off[k + 1]] = uk
            x = A @ x + self.B[k] @ uk
        return u

    def hinv(self, v):
        return self.solve(np.zeros(13), None, -v)


def solve_one(f, h, tol=1e-9, max_iter=400):
    gait = np.asarray(f["gait"]).reshape(h, 4)
    ub = (gait.astype(np.float32) * np.float32(f["f_max"])).astype(np.float64)
    keep = ~((ub < 0.01) & (ub > -0.01))
    stance = [list(np.nonzero(keep[k])[0]) for k in range(h)]
    Ad, Bd, Q, x0 = dynamics(f)
    ric = Riccati(Ad, Bd, Q, float(f["alpha"]), stance, h)
    nv = ric.nv
    xd = np.zeros((h, 13))
    xd[:, :12] = np.asarray(f["traj"], float).reshape(h, 12)
    x = ric.solve(x0, xd, np.zeros(nv))
    # constraints C u >= b
    mu_inv = float(np.float32(1.0) / np.float32(f["mu"]))
    rows, rhs = [], []
    full_idx = []
    j = 0
    for k in range(h):
        for l in stance[k]:
            fmax = ub[k, l]
            for (cx, cy, cz, bb) in ((mu_inv, 0, 1, 0), (-mu_inv, 0, 1, 0), (0, mu_inv, 1, 0), (0, -mu_inv, 1, 0),
                                     (0, 0, 1, 0), (0, 0, -1, -fmax)):
                rows.append((j, cx, cy, cz))
                rhs.append(bb)
            full_idx += [12 * k + 3 * l, 12 * k + 3 * l + 1, 12 * k + 3 * l + 2]
            j += 3
    rhs = np.array(rhs, float)
    nc = len(rows)

    def normal(c):
        j, cx, cy, cz = rows[c]
        n = np.zeros(nv)
        n[j:j + 3] = (cx, cy, cz)
        return n

    Cm = np.stack([normal(c) for c in range(nc)]) if nc else np.zeros((0, nv))
    Wset, lam, Z = [], [], []
    it = 0
    solves = 1
    while True:
        s = Cm @ x - rhs
        s[Wset] = 0.0
        p = int(np.argmin(s)) if nc else -1
        if p < 0 or s[p] >= -tol:
            break
        n = Cm[p]
        d = ric.hinv(n)
        solves += 1
        up = 0.0
        while True:
            it += 1
            if it > max_iter:
                return None, it, solves
            if Wset:
                Zm = np.stack(Z, 1)
                N = Cm[Wset].T
                T = np.linalg.inv(N.T @ Zm)
                r = T @ (Zm.T @ n)
                z = d - Zm @ r
            else:
                r = np.zeros(0)
                z = d
            zn = z @ n
            sp = Cm[p] @ x - rhs[p]
            t2 = -sp / zn if zn > 1e-14 else np.inf
            t1, l = np.inf, -1
            for jj in range(len(Wset)):
                if r[jj] > 1e-14 and lam[jj] / r[jj] < t1:
                    t1, l = lam[jj] / r[jj], jj
            t = min(t1, t2)
            if not np.isfinite(t):
                return None, it, solves
            if np.isfinite(t2):
                x = x + t * z
            lam = [lam[jj] - t * r[jj] for jj in range(len(Wset))]
            up += t
            if t == t2:
                Wset.append(p)
                lam.append(up)
                Z.append(d)
                break
            del Wset[l], lam[l], Z[l]
    sol = np.zeros(12 * h)
    sol[full_idx] = x
    return sol, it, solves


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "config2"
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 32
    h = {"config2": 10, "config3": 20, "config4": 10, "config5": 16}[name]
    rec = W.CONFIGS[name](n, h, 1234)
    o64 = O.solve_batch(rec, h, 64)
    F = R.unpack_records(rec, h)
    errs, its, sv = [], [], []
    for b in range(n):
        f = {k: v[b] for k, v in F.items()}
        sol, it, solves = solve_one(f, h)
        if sol is None:
            print(b, "FAILED", it)
            continue
        if o64["rc"][b] != 0:
            continue
        e = np.linalg.norm(sol - o64["sol"][b]) / max(np.linalg.norm(o64["sol"][b]), 1.0)
        errs.append(e)
        its.append(it)
        sv.append(solves)
    print("%s: n=%d  max rel err vs oracle64 %.3e  median %.3e ; active-set iterations mean %.2f max %d ; Riccati solves "
          "mean %.2f max %d" % (name, len(errs), max(errs), np.median(errs), np.mean(its), max(its), np.mean(sv), max(sv)))


if __name__ == "__main__":
    main()
