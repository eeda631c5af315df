#!/usr/bin/env python
"""Where the tick host entry's time goes (development aid; run on the GPU box)."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from quadruped_ctrl_b200 import engine as E, workloads as W  # noqa: E402

B, h = 4096, 10
eng = E.MpcBatch(h, B)
tk = torch.from_numpy(W.config2_ticks(B, h, 1234)).pin_memory()
rec = torch.from_numpy(W.config2(B, h, 1234)).pin_memory()
d_tk = tk.cuda()
d_rec = torch.empty((B, eng.stride), dtype=torch.uint8, device="cuda")


def ev_time(fn, n=50):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / n


print("build_records kernel: %.1f us" % ev_time(lambda: eng.build_records_device(d_tk, d_rec)))
print("solve_ticks_device:   %.1f us" % ev_time(lambda: eng.solve_ticks_device(d_tk)))
d_rec2 = rec.cuda()
f = torch.empty((B, 12), dtype=torch.float32, device="cuda")
st = torch.empty((B,), dtype=torch.int32, device="cuda")
print("solve_device:         %.1f us" % ev_time(lambda: eng.solve_device(d_rec2, forces=f, status=st)))


def host_loop(submit, n=200, depth=E.SLOTS - 1):
    t_sub = 0.0
    t0 = time.perf_counter()
    for i in range(n):
        a = time.perf_counter()
        submit(i % E.SLOTS)
        t_sub += time.perf_counter() - a
        if i >= depth:
            eng.wait_host((i - depth) % E.SLOTS)
    for j in range(max(0, n - depth), n):
        eng.wait_host(j % E.SLOTS)
    tot = time.perf_counter() - t0
    return tot / n * 1e6, t_sub / n * 1e6


for name, fn in (("records zero-copy", lambda q: eng.submit_host(q, rec.numpy(), zero_copy=True)),
                 ("ticks zero-copy", lambda q: eng.submit_host_ticks(q, tk.numpy(), zero_copy=True)),
                 ("records staged", lambda q: eng.submit_host(q, rec.numpy())),
                 ("ticks staged", lambda q: eng.submit_host_ticks(q, tk.numpy()))):
    host_loop(fn, 30)
    us, sub = host_loop(fn)
    print("%-18s %.1f us/step (%.2f M solves/s), host time in submit %.1f us" % (name, us, B / us, sub))
