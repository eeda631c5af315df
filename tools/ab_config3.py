#!/usr/bin/env python
"""Same-box timing of the mixed-gait horizon-20 workload (config 3) for several builds / catch-all grid sizes."""
import os
import subprocess
import sys

CHILD = r'''
import sys, numpy as np, torch
sys.path.insert(0, ".")
from quadruped_ctrl_b200 import engine as E, workloads as W
from oracle import oracle as O
name, B = "config3", 4096
h = W.HORIZONS[name]
rec_h = W.CONFIGS[name](B)
rec = torch.from_numpy(rec_h).cuda()
eng = E.MpcBatch(h, B)
f, s, st = eng.solve_device(rec, want_solution=True)
torch.cuda.synchronize()
o = O.solve_batch(rec_h[:96], h, 64)
den = np.maximum(np.linalg.norm(o["sol"], axis=1), 1.0)
err = (np.linalg.norm(s.cpu().numpy()[:96] - o["sol"], axis=1) / den).max()
for _ in range(2): eng.solve_device(rec, forces=f, status=st)
torch.cuda.synchronize()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record()
for _ in range(5): eng.solve_device(rec, forces=f, status=st)
ev1.record(); torch.cuda.synchronize()
ms = ev0.elapsed_time(ev1) / 5
print("config3 B=%d: %.3f ms -> %.0f solves/s ; status %s ; max rel err vs oracle64 %.1e ; classes %s" % (
    B, ms, B / ms * 1e3, np.bincount(st.cpu().numpy() & 0xff), err, [(c["nv_cap"], c["grid"]) for c in eng.classes()]))
'''
for spec in sys.argv[1:]:
    lib, _, ctas = spec.partition(":")
    env = dict(os.environ, MPC_LIB_PATH=os.path.abspath(lib))
    if ctas:
        env["MPC_BIG_CTAS"] = ctas
    r = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True)
    print("%-24s %s" % (spec, r.stdout.strip() or r.stderr.strip()[-400:]), flush=True)
