#!/usr/bin/env python
"""Same-process A/B of the two solvers (explicit inverse of the condensed Hessian vs Riccati sweeps) on the BASELINE
workloads: solves/s of each (one batch at a time, rounds interleaved), the distance between their solutions, the class
configuration each one runs with, and the problems the Riccati tile re-queued."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from quadruped_ctrl_b200 import engine as E, workloads as W  # noqa: E402

cases = [("config2", 4096), ("config2", 65536), ("four_stance", 4096), ("config5", 4096), ("config5", 65536), ("config3", 4096)]
if len(sys.argv) > 1:
    cases = [(a.split(":")[0], int(a.split(":")[1])) for a in sys.argv[1:]]
for name, B in cases:
    h = W.HORIZONS[name]
    rec = torch.from_numpy(W.CONFIGS[name](B)).cuda()
    eng = E.MpcBatch(h, B)
    sols, rates, cls = {}, {"inverse": [], "riccati": []}, {}
    for rnd in range(3):
        for v in ("inverse", "riccati"):
            eng.set_solver(v)
            cls[v] = [(c["nv_cap"], c["m_cap"], c["threads"], c["grid"], c["smem"]) for c in eng.classes()]
            f, s, st = eng.solve_device(rec, want_solution=True)
            for _ in range(2):
                eng.solve_device(rec, forces=f, status=st)
            torch.cuda.synchronize()
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n = 10 if B > 4096 else 30
            ev0.record()
            for _ in range(n):
                eng.solve_device(rec, forces=f, status=st)
            ev1.record()
            torch.cuda.synchronize()
            rates[v].append(B / (ev0.elapsed_time(ev1) / n) / 1e3)
            sols[v] = (s.cpu().numpy(), st.cpu().numpy())
    d = np.linalg.norm(sols["inverse"][0] - sols["riccati"][0], axis=1) / np.maximum(np.linalg.norm(sols["inverse"][0], axis=1), 1.0)
    ok = [(sols[v][1] & 0xff == 0).mean() for v in ("inverse", "riccati")]
    it = [(sols[v][1] >> 8).mean() for v in ("inverse", "riccati")]
    print("%-12s B=%-6d inverse %.2f M/s  riccati %.2f M/s  (x%.3f)  |sol_ric - sol_inv| max %.1e  optimal %.4f / %.4f  "
          "iterations %.2f / %.2f" % (name, B, max(rates["inverse"]), max(rates["riccati"]),
                                      max(rates["riccati"]) / max(rates["inverse"]), d.max(), ok[0], ok[1], it[0], it[1]), flush=True)
    print("    classes (nv_cap, m_cap, threads, grid, smem): inverse %s" % cls["inverse"])
    print("                                                  riccati %s" % cls["riccati"], flush=True)
    eng.close()
