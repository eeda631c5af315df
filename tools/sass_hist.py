#!/usr/bin/env python
"""SASS opcode histogram of the shipped library's kernels (code-object evidence for profiles/):
  python tools/sass_hist.py quadruped_ctrl_b200/libquadruped_mpc_b200.so > profiles/r2_sass_opcodes.txt
Per kernel: instruction count, SASS bytes, and the counts of the opcodes that identify which units the kernel uses
(DFMA/DMUL/DADD: FP64 FMA pipe; DMMA: FP64 tensor pipe; UBLKCP / SYNCS: TMA bulk copy + mbarrier; LDS/STS: shared memory;
SHFL; CREDUX: warp-wide integer reductions; BAR / WARPSYNC; MUFU.RCP64H)."""
import re
import subprocess
import sys
from collections import Counter, OrderedDict

lib = sys.argv[1]
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
kern = OrderedDict()
cur = None
for ln in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", ln)
    if m:
        cur = m.group(1)
        kern[cur] = Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P[T\d]\s+)?([A-Z][A-Z0-9_.]+)", ln)
    if m and cur:
        kern[cur][m.group(1)] += 1
dem = subprocess.run(["cu++filt"] + list(kern), capture_output=True, text=True).stdout.splitlines()
KEYS = ["DFMA", "DMUL", "DADD", "DMMA", "UBLKCP", "SYNCS", "LDS", "STS", "LDG", "STG", "SHFL", "BAR", "WARPSYNC", "MUFU",
        "CREDUX", "REDUX", "ATOMG", "HMMA", "UTCHMMA", "UTMALDG"]
for (name, c), d in zip(kern.items(), dem if len(dem) == len(kern) else list(kern)):
    n = sum(c.values())
    fam = Counter()
    for op, v in c.items():
        for k in KEYS:
            if op.startswith(k):
                fam[k] += v
    short = re.sub(r"\(anonymous namespace\)::|<unnamed>::", "", d)
    short = re.sub(r"\((anonymous namespace|<unnamed>)::SolveParams\)", "", short)
    print("%s\n    %d instructions (%.1f KB): %s" % (short[:150], n, n * 16 / 1024.0,
                                                      "  ".join("%s %d" % (k, fam[k]) for k in KEYS if fam[k])))
