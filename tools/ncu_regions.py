#!/usr/bin/env python
"""Executed warp instructions and stall samples of an ncu capture aggregated by source region of mpc_core.h
(function / assembly phase), per problem.   python tools/ncu_regions.py <report.ncu-rep> <lib.so> <kernel substring> <problems>"""
import os
import re
import sys
from collections import defaultdict

sys.path.insert(0, "tools")
from ncu_lines import line_table, sass_rows  # noqa: E402

rep, lib, sub, nprob = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
hdr, rows = sass_rows(rep)
ix = {h: i for i, h in enumerate(hdr)}
table = line_table(lib, sub)
src = open(os.path.join("quadruped_ctrl_b200", "csrc", "mpc_core.h")).read().splitlines()
marks = []
for i, l in enumerate(src, 1):
    m = re.match(r"(?:MPC_HDN?|__device__ (?:__forceinline__|__noinline__)) \S+ (\w+)\(", l) or re.match(r"\s*// ---- (P\d+)", l)
    if m:
        marks.append((i, m.group(1)))


def region(key):
    if key is None:
        return "(none)"
    if key[0] != "mpc_core.h":
        return key[0]
    name = "?"
    for i, n in marks:
        if i <= key[1]:
            name = n
    return "core:" + name


n = min(len(table), len(rows))
if len(table) != len(rows):
    print("warning: %d rows in report, %d in nvdisasm (library must be the profiled build)" % (len(rows), len(table)))
inst, samp = defaultdict(float), defaultdict(float)
for i in range(n):
    r = region(table[i][0])
    inst[r] += float(rows[i][ix["Instructions Executed"]] or 0)
    samp[r] += float(rows[i][ix["# Samples"]] or 0)
ti, ts = sum(inst.values()), sum(samp.values())
print("%-34s %12s %8s %8s" % ("region", "inst/problem", "inst%", "samples%"))
for r in sorted(inst, key=lambda k: -samp[k]):
    if inst[r] or samp[r]:
        print("%-34s %12.0f %7.1f%% %7.1f%%" % (r, inst[r] / nprob, 100 * inst[r] / ti, 100 * samp[r] / max(ts, 1)))
