#!/usr/bin/env python
"""Where the end-to-end (host buffers) step time goes (development aid; run on the GPU box)."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from quadruped_ctrl_b200 import engine as E, workloads as W  # noqa: E402

B, h = 4096, 10
eng = E.MpcBatch(h, B)
sets = [torch.from_numpy(W.config2(B, h, 1234 + i)).pin_memory() for i in range(8)]
arrs = [s.numpy() for s in sets]
n = 200


def run(depth, nslots=E.SLOTS):
    t_sub = t_wait = 0.0
    t0 = time.perf_counter()
    for i in range(n):
        a = time.perf_counter()
        eng.submit_host(i % nslots, arrs[i % 8])
        b = time.perf_counter()
        if i >= depth:
            eng.wait_host((i - depth) % nslots)
        c = time.perf_counter()
        t_sub += b - a
        t_wait += c - b
    for j in range(max(0, n - depth), n):
        eng.wait_host(j % nslots)
    torch.cuda.synchronize()
    tot = time.perf_counter() - t0
    print("depth %d: %.1f us/step (%.2f M solves/s); host time in submit %.1f us, in wait %.1f us per step"
          % (depth, tot / n * 1e6, B * n / tot / 1e6, t_sub / n * 1e6, t_wait / n * 1e6))


run(1)
run(1)
run(2)
run(2)
# raw copies for scale
dev = torch.empty((B, eng.stride), dtype=torch.uint8, device="cuda")
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(50):
    dev.copy_(sets[i % 8], non_blocking=True)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 50
print("H2D of one batch (%.2f MB, pinned): %.1f us = %.1f GB/s" % (B * eng.stride / 1e6, dt * 1e6, B * eng.stride / dt / 1e9))
# device-resident serial step for reference
rec = sets[0].cuda()
f, _, st = eng.solve_device(rec)
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(100):
    eng.solve_device(rec, forces=f, status=st)
torch.cuda.synchronize()
print("device-resident serial step: %.1f us" % ((time.perf_counter() - t0) / 100 * 1e6))
# synchronous host solve (submit + wait, no overlap)
t0 = time.perf_counter()
for i in range(50):
    eng.submit_host(0, arrs[i % 8])
    eng.wait_host(0)
print("synchronous host step: %.1f us" % ((time.perf_counter() - t0) / 50 * 1e6))
