#!/usr/bin/env python
"""Same-box A/B of engine builds (development aid): interleaved timing of config 2 for each library given.
  python tools/ab_quick.py lib_a.so lib_b.so ...   (each library is loaded in its own subprocess, rounds interleaved)"""
import os
import subprocess
import sys

CHILD = r'''
import sys, torch
sys.path.insert(0, ".")
from quadruped_ctrl_b200 import engine as E, workloads as W
out = []
for name, B in (("config2", 65536), ("config2", 4096), ("four_stance", 4096), ("config5", 4096)):
    h = W.HORIZONS[name]
    rec = torch.from_numpy(W.CONFIGS[name](B)).cuda()
    eng = E.MpcBatch(h, B)
    f, _, st = eng.solve_device(rec)
    for _ in range(3): eng.solve_device(rec, forces=f, status=st)
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 10 if B > 4096 else 30
    ev0.record()
    for _ in range(n): eng.solve_device(rec, forces=f, status=st)
    ev1.record(); torch.cuda.synchronize()
    out.append("%s/%d %.2fM" % (name, B, B / (ev0.elapsed_time(ev1) / n) / 1e3))
    eng.close()
print("  ".join(out))
'''
libs = sys.argv[1:]
for rnd in range(2):
    for lib in libs:
        env = dict(os.environ, MPC_LIB_PATH=os.path.abspath(lib))
        r = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True)
        print("%-28s %s" % (os.path.basename(lib), r.stdout.strip() or r.stderr.strip()[-300:]), flush=True)
