#!/usr/bin/env python
"""Same-process A/B of the two register-resident inversions (FMA rank-1 sweep vs DMMA grouped sweep) on the BASELINE
workloads: solves/s of each (one batch at a time, rounds interleaved) and the distance between their solutions."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from quadruped_ctrl_b200 import engine as E, workloads as W  # noqa: E402

cases = [("config2", 65536), ("config2", 4096), ("four_stance", 4096), ("config5", 4096), ("config3", 4096)]
for name, B in cases:
    h = W.HORIZONS[name]
    rec = torch.from_numpy(W.CONFIGS[name](B)).cuda()
    eng = E.MpcBatch(h, B)
    sols, rates = {}, {"fma": [], "mma": []}
    for rnd in range(3):
        for v in ("fma", "mma"):
            eng.set_sweep_variant(v)
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
    d = np.linalg.norm(sols["fma"][0] - sols["mma"][0], axis=1) / np.maximum(np.linalg.norm(sols["fma"][0], axis=1), 1.0)
    ok = [(sols[v][1] & 0xff == 0).mean() for v in ("fma", "mma")]
    print("%-12s B=%-6d fma %.2f M/s  mma %.2f M/s  (x%.3f)  |sol_mma - sol_fma| max %.1e  optimal %.3f / %.3f" %
          (name, B, max(rates["fma"]), max(rates["mma"]), max(rates["mma"]) / max(rates["fma"]), d.max(), ok[0], ok[1]),
          flush=True)
    eng.close()
