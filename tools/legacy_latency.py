#!/usr/bin/env python
"""Latency of the legacy single-robot entry (setup_problem ... get_solution), the call sequence the reference issues
once per MPC tick (development aid; run on the GPU box)."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from quadruped_ctrl_b200 import interface as I, records as R, workloads as W  # noqa: E402

for h, rec in ((10, W.config1()), (14, W.CONFIGS["config5"](4)[:0] if False else None)):
    if rec is None:
        continue
    f = R.unpack_records(rec, h)
    n = rec.shape[0]

    def tick(b):
        I.setup_problem(float(f["dt"][b]), h, float(f["mu"][b]), float(f["f_max"][b]))
        I.update_x_drag(float(f["x_drag"][b]))
        I.update_solver_settings(10000, 1e-7, 1e-8, 1.5, 0.1, 0.0)
        I.update_problem_data_floats(f["p"][b], f["v"][b], f["q"][b], f["w"][b], f["r"][b], float(f["yaw"][b]),
                                     f["weights"][b], f["traj"][b], float(f["alpha"][b]), f["gait"][b].astype(np.int32))
        return [I.get_solution(i) for i in range(12)]

    for _ in range(20):
        tick(0)
    ts = []
    for k in range(300):
        t0 = time.perf_counter()
        tick(k % n)
        ts.append(time.perf_counter() - t0)
    ts = np.array(ts) * 1e6
    print("legacy tick h=%d: median %.1f us, p95 %.1f us (python ctypes overhead included), status %d"
          % (h, np.median(ts), np.percentile(ts, 95), I.last_status()))
I.shutdown()
