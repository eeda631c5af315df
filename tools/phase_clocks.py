#!/usr/bin/env python
"""Per-phase SM-clock breakdown of the solve kernel (profiling aid; run on the GPU box)."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from quadruped_ctrl_b200 import engine as E, workloads as W  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "config2"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
h = W.HORIZONS[name]
rec = torch.from_numpy(W.CONFIGS[name](B)).cuda()
eng = E.MpcBatch(h, B)
if len(sys.argv) > 3:
    eng.set_ctas_per_sm_limit(int(sys.argv[3]))
for _ in range(3):
    eng.solve_device(rec)
buf = torch.zeros((B, 24), dtype=torch.int64, device="cuda")
eng.set_phase_clock_buffer(buf)
eng.solve_device(rec)
torch.cuda.synchronize()
c = buf.cpu().numpy()
d = np.diff(c[:, :5], axis=1)
names = ["assemble", "invert", "active_set", "scatter+sync"]
# optional 4th argument: only problems with more than that many reduced variables (e.g. the catch-all class)
if len(sys.argv) > 4:
    from quadruped_ctrl_b200 import records as R
    rec_h = W.CONFIGS[name](B)
    go = R.gait_offset(h)
    nvv = 3 * (rec_h[:, go:go + 4 * h] > 0).sum(1)
    keep = nvv > int(sys.argv[4])
    c = c[keep]
    d = d[keep]
    print("filtered to %d problems with nv > %s" % (keep.sum(), sys.argv[4]))
print("%s B=%d: cycles per problem per CTA (median / mean / p95)" % (name, B))
for i, n in enumerate(names):
    print("  %-14s %8.0f %8.0f %8.0f" % (n, np.median(d[:, i]), d[:, i].mean(), np.percentile(d[:, i], 95)))
sub = [("asm: P0-P1 (flags, checks, trig)", 0, 8), ("asm: P2 (B_c per leg)", 8, 9), ("asm: P3-P7 (C0,C1,C2,qe)", 9, 10),
       ("asm: P8-P9 (moments, g, M_ab)", 10, 11), ("asm: P11 (H blocks)", 11, 1), ("gi: x = -Minv g", 2, 12),
       ("gi: iterations", 12, 3), ("gi: first search", 12, 13), ("gi: first working-set change", 13, 14)]
iters = None
for n, a, b in sub:
    sel = (c[:, a] > 0) & (c[:, b] > 0)
    dd = (c[:, b] - c[:, a])[sel]
    print("    %-36s %8.0f %8.0f %8.0f" % (n, np.median(dd), dd.mean(), np.percentile(dd, 95)))
its = None
tot = c[:, 4] - c[:, 0]
print("  %-14s %8.0f %8.0f %8.0f" % ("total", np.median(tot), tot.mean(), np.percentile(tot, 95)))
# gap between consecutive problems of the same CTA (record wait + loop overhead)
order = np.lexsort((c[:, 6], c[:, 5]))
cs = c[order]
same = cs[1:, 5] == cs[:-1, 5]
gap = (cs[1:, 0] - cs[:-1, 4])[same]
print("  inter-problem gap median %.0f mean %.0f" % (np.median(gap), gap.mean()))
first = cs[np.r_[True, ~same]]
print("  kernel span (max end - min start) %.0f cycles; first-problem start spread %.0f" %
      (c[:, 4].max() - c[:, 0].min(), first[:, 0].max() - first[:, 0].min()))

# hardware warp slot of thread 0 (%warpid) per SM: which SM sub-partitions (slot % 4) the CTAs' first warps sit on
wid = c[:, 7] & 0xff
sm = c[:, 7] >> 8
print("  %%warpid of warp 0, histogram of (slot %% 4): %s ; distinct slots %s" % (np.bincount(wid % 4, minlength=4), np.unique(wid)))
one = sm == sm[0]
print("  SM %d: slots used %s" % (sm[0], np.unique(wid[one])))

# Are the co-resident CTAs phase-locked?  For every SM: at the start of each sweep, how many OTHER CTAs of that SM are
# inside their sweep (clk[1]..clk[2]) at that instant, against the expectation if the CTAs ran independently.
starts, ends, sms = c[:, 1], c[:, 2], sm
frac_in = ((ends - starts).sum() / max(1, (c[:, 4] - c[:, 0]).sum()))
cnt = []
for s_id in np.unique(sms)[:40]:
    sel = sms == s_id
    a, b_, t0, t1 = starts[sel], ends[sel], c[sel, 0], c[sel, 4]
    lo, hi = np.percentile(t0, 20), np.percentile(t1, 80)   # steady part of the launch
    for t in a[(a > lo) & (a < hi)]:
        others = ((a < t) & (b_ > t)).sum()
        alive = ((t0 < t) & (t1 > t)).sum() - 1
        cnt.append((others, alive))
cnt = np.array(cnt)
if len(cnt):
    print("  sweep overlap: at a sweep start %.2f other CTAs of the SM are sweeping (mean; %.2f other CTAs mid-problem; "
          "independent phases would give %.2f); histogram %s"
          % (cnt[:, 0].mean(), cnt[:, 1].mean(), cnt[:, 1].mean() * frac_in, np.bincount(cnt[:, 0])))
