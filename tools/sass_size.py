#!/usr/bin/env python
"""Static SASS footprint of a kernel by source region (development aid): instruction bytes per function / line range.
  python tools/sass_size.py <lib.so> <kernel substring>"""
import sys
from collections import defaultdict

sys.path.insert(0, "tools")
from ncu_lines import line_table  # noqa: E402

lib, sub = sys.argv[1:3]
table = line_table(lib, sub)
print("kernel %s: %d SASS instructions = %.1f KB" % (sub, len(table), len(table) * 16 / 1024))
# regions of mpc_core.h by line number (function starts), found by scanning the source for the stage markers
import re, os
src = open(os.path.join("quadruped_ctrl_b200", "csrc", "mpc_core.h")).read().splitlines()
marks = []
for i, l in enumerate(src, 1):
    m = re.match(r"(?:MPC_HDN?|__device__ (?:__forceinline__|__noinline__)) \S+ (\w+)\(", l) or re.match(r"\s*// ---- (P\d+)", l)
    if m:
        marks.append((i, m.group(1)))
def region(line):
    name = "?"
    for i, n in marks:
        if i <= line:
            name = n
    return name
cnt = defaultdict(int)
for key, _ in table:
    if key is None:
        cnt["(none)"] += 1
    elif key[0] == "mpc_core.h":
        cnt["core:" + region(key[1])] += 1
    else:
        cnt[key[0]] += 1
for k, v in sorted(cnt.items(), key=lambda kv: -kv[1]):
    print("  %-40s %6d instr %7.1f KB" % (k, v, v * 16 / 1024))
