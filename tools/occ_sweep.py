#!/usr/bin/env python
"""Throughput of the dominant solve kernel against resident CTAs per SM (development aid; run on the GPU box).
A latency-bound kernel scales with the limit; a saturated unit shows as a plateau."""
import sys

import torch

sys.path.insert(0, ".")
from quadruped_ctrl_b200 import engine as E, workloads as W  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "config2"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
h = W.HORIZONS[name]
rec = torch.from_numpy(W.CONFIGS[name](B)).cuda()
eng = E.MpcBatch(h, B)
f, _, st = eng.solve_device(rec)
for lim in (1, 2, 3, 4, 5, 6, 7, 8):
    eng.set_ctas_per_sm_limit(lim)
    for _ in range(2):
        eng.solve_device(rec, forces=f, status=st)
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 5
    ev0.record()
    for _ in range(n):
        eng.solve_device(rec, forces=f, status=st)
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / n
    print("%s B=%d CTAs/SM limit %d: %.3f ms -> %.2f M solves/s (%.0f SM-cycles per problem at 1.965 GHz)"
          % (name, B, lim, ms, B / ms / 1e3, ms * 1e-3 * 1.965e9 * 148 / B))
eng.close()
