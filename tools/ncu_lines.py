#!/usr/bin/env python
"""Attributes an ncu SASS source page to CUDA source lines.

ncu's CSV export of the source page is SASS-only; `nvdisasm -g` of the same cubin carries the
file:line markers.  Both list the kernel's instructions in address order, so they are joined by index.

  python tools/ncu_lines.py <report.ncu-rep> <lib.so> <kernel substring, e.g. "ILi128E"> [top N]
"""
import csv
import io
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict


def sass_rows(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    # one section (header row + SASS rows) per captured launch; NCU_LINES_KERNEL picks the section (default: the first)
    his = [i for i, r in enumerate(rows) if "Source" in r and "Address" in r]
    which = int(os.environ.get("NCU_LINES_KERNEL", "0"))
    hi = his[which]
    end = his[which + 1] if which + 1 < len(his) else len(rows)
    hdr = rows[hi]
    return hdr, [r for r in rows[hi + 1:end] if len(r) == len(hdr)]


def line_table(lib, kernel_sub):
    d = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=d, capture_output=True)
    lines = []
    for f in sorted(os.listdir(d)):
        if not f.endswith(".cubin"):
            continue
        txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(d, f)], capture_output=True, text=True).stdout
        cur = None
        inside = False
        for ln in txt.splitlines():
            if ln.startswith(".text."):
                inside = kernel_sub in ln
                continue
            if not inside:
                continue
            m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
            if m:
                cur = (os.path.basename(m.group(1)), int(m.group(2)))
                continue
            if re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln):
                lines.append((cur, ln.strip()))
        if lines:
            break
    return lines


def main():
    rep, lib, sub = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    hdr, rows = sass_rows(rep)
    ix = {h: i for i, h in enumerate(hdr)}
    table = line_table(lib, sub)
    if len(table) != len(rows):
        print("warning: %d SASS rows in the report vs %d in nvdisasm" % (len(rows), len(table)))
    n = min(len(table), len(rows))
    inst = defaultdict(float)
    samp = defaultdict(float)
    for i in range(n):
        key = table[i][0]
        inst[key] += float(rows[i][ix["Instructions Executed"]] or 0)
        samp[key] += float(rows[i][ix["# Samples"]] or 0)
    ti, ts = sum(inst.values()), sum(samp.values())
    src = {}
    print("total warp instructions %.3g, samples %d" % (ti, ts))
    print("%7s %7s  %s" % ("inst%", "samp%", "file:line"))
    for key in sorted(inst, key=lambda k: -samp[k])[:top]:
        text = ""
        if key:
            p = os.path.join(os.path.dirname(os.path.abspath(lib)), "csrc", key[0])
            if p not in src and os.path.exists(p):
                src[p] = open(p).read().splitlines()
            if p in src and key[1] - 1 < len(src[p]):
                text = src[p][key[1] - 1].strip()[:90]
        print("%6.2f%% %6.2f%%  %s:%s  %s" % (100 * inst[key] / ti, 100 * samp[key] / max(ts, 1), key[0] if key else "?",
                                             key[1] if key else "?", text))


if __name__ == "__main__":
    main()
