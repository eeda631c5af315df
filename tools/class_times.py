#!/usr/bin/env python
"""Per-class kernel time of one workload (development aid).  python tools/class_times.py config3 4096"""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from quadruped_ctrl_b200 import engine as E, records as R, workloads as W  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "config3"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
h = W.HORIZONS[name]
rec_h = W.CONFIGS[name](B)
go = R.gait_offset(h)
nv = 3 * (rec_h[:, go:go + 4 * h] > 0).sum(1)
rec = torch.from_numpy(rec_h).cuda()
eng = E.MpcBatch(h, B)
eng.set_timing(True)
f, _, st = eng.solve_device(rec)
for _ in range(3):
    eng.solve_device(rec, forces=f, status=st)
torch.cuda.synchronize()
its = st.cpu().numpy() >> 8
lo = 0
for ci, c in enumerate(eng.classes()):
    sel = (nv > lo) & (nv <= c["nv_cap"]) if ci < len(eng.classes()) - 1 else nv > lo
    ms = eng.last_class_kernel_ms(ci) if hasattr(eng, "last_class_kernel_ms") else float("nan")
    print("class %d nv<=%d grid %d: %d problems, kernel %.3f ms, iterations mean %.1f max %d" % (
        ci, c["nv_cap"], c["grid"], sel.sum(), ms, its[sel].mean() if sel.any() else 0, its[sel].max() if sel.any() else 0))
    if ci < len(eng.classes()) - 1:
        lo = c["nv_cap"]
