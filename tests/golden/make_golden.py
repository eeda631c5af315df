"""Generates tests/golden/mpc_golden.npz -- run in the BUILD container (needs /root/reference).

The reference ships no golden vectors for this path, so the fixture is made from
outputs of the reference's own solver run here: oracle/_ref/libqpoases_ref.so is the reference's
qpOASES 3.2 compiled unmodified from /root/reference/src/qpOASES (oracle/Makefile), driven with the call
sequence of SolverMPC.cpp:529-539 on QPs assembled by the oracle restatement in fp32 (reference-faithful)
and fp64 (rounding-free).  Small on purpose: a handful of problems per BASELINE config + edge cases.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import oracle as O  # noqa: E402
from quadruped_ctrl_b200 import records as R  # noqa: E402
from quadruped_ctrl_b200 import workloads as W  # noqa: E402


def edge_cases(h=10):
    """Edge cases of the domain: f_max saturation, x_drag != 0, a flight phase, one-leg stance, heavy payload."""
    rec = W.config2(8, h, 99)
    f = rec.view(np.float32)
    f[0, R.REC_FMAX] = 12.0          # f_max so low that the upper bound is active
    f[1, R.REC_XDRAG] = 0.35         # drag integrator engaged (ConvexMPCLocomotion.cpp:632-640)
    f[2, R.REC_MU] = 0.1             # slippery: cone rows active
    f[3, R.REC_MASS] = 20.0          # heavy payload
    go = R.gait_offset(h)
    rec[4, go:go + 4 * h] = 0
    rec[4, go + 4 * 3:go + 4 * 7] = 1          # flight, stance for 4 steps, flight
    rec[5, go:go + 4 * h] = 0
    rec[5, go:go + 4 * h:4] = 1                # a single leg in stance over the horizon
    f[6, R.REC_P + 2] = 0.15                   # far below the commanded height: large forces
    f[7, R.REC_V] = 2.0                        # fast
    return rec


def main():
    assert O.have_reference_qpoases(), "needs oracle/_ref (built from /root/reference)"
    out = {}
    cases = {"config1": (W.config1(), 10), "config2": (W.config2(12), 10), "config3": (W.config3(8), 20),
             "config4": (W.config4(12), 10), "config5": (W.config5(8), 16), "four_stance": (W.four_stance(6), 10),
             "edge": (edge_cases(), 10)}
    for name, (rec, h) in cases.items():
        o32 = O.solve_batch(rec, h, 32, "reference")
        o64 = O.solve_batch(rec, h, 64, "reference")
        out[name + "_records"] = rec
        out[name + "_h"] = np.int32(h)
        for tag, o in (("o32", o32), ("o64", o64)):
            out["%s_%s_sol" % (name, tag)] = o["sol"]
            out["%s_%s_nwsr" % (name, tag)] = o["nwsr"]
            out["%s_%s_rc" % (name, tag)] = o["rc"]
            out["%s_%s_nv" % (name, tag)] = o["nv"]
    path = os.path.join(ROOT, "tests", "golden", "mpc_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
