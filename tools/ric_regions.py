#!/usr/bin/env python
"""Aggregates the per-line table of tools/ncu_lines.py (all lines) by function / factorisation stage of
csrc/mpc_riccati.h.   python tools/ric_regions.py gpurun_out/r2_ric_c2_lines.txt [problems]"""
import re
import sys
from collections import defaultdict

src = open('quadruped_ctrl_b200/csrc/mpc_riccati.h').read().splitlines()
marks = []
for i, l in enumerate(src, 1):
    m = re.match(r"(?:MPC_HD|__device__ __forceinline__) \S+ (\w+)\(", l) or re.match(r"template <int NTU>", l) and None
    if m:
        marks.append((i, m.group(1)))
    m = re.match(r"\s*// ---- \((\d)\)", l) or re.match(r"\s*// ---- (Y) = P A", l)
    if m:
        marks.append((i, "step_mma(" + m.group(1) + ")"))
    if re.match(r"__device__ __forceinline__ bool ric_step_mma", l):
        marks.append((i, "step_mma(head)"))


def region(f, ln):
    if f != 'mpc_riccati.h':
        return f
    name = '?'
    for i, n in marks:
        if i <= ln:
            name = n
    return name


inst = defaultdict(float)
samp = defaultdict(float)
tot = None
for l in open(sys.argv[1]):
    m = re.match(r"total warp instructions ([\d.e+]+), samples (\d+)", l)
    if m:
        tot = float(m.group(1))
    m = re.match(r"\s*([\d.]+)%\s+([\d.]+)%\s+(\S+):(\d+|\?)", l)
    if not m:
        continue
    f = m.group(3)
    ln = int(m.group(4)) if m.group(4) != '?' else 0
    r = region(f, ln)
    inst[r] += float(m.group(1))
    samp[r] += float(m.group(2))
nprob = float(sys.argv[2]) if len(sys.argv) > 2 else 4096
if tot:
    print("warp instructions per problem: %.0f" % (tot / nprob))
for r in sorted(inst, key=lambda r: -samp[r]):
    print("%6.2f%% inst %6.2f%% samples  %s" % (inst[r], samp[r], r))
