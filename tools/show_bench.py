#!/usr/bin/env python
"""Compact view of bench.py JSON lines (one per file argument or per line of a file)."""
import json
import sys

for path in sys.argv[1:]:
    for line in open(path):
        line = line.strip()
        if not line.startswith("{"):
            continue
        d = json.loads(line)
        if "unavailable" in d:
            print(path, d)
            continue
        r = d.get("roofline", {})
        print("%s | %s" % (path, d["config"]["workload"][:60]))
        print("   solver %s  collective: %s  host queue %s ms/step" % (
            d["config"].get("solver"), str(d["config"].get("collective"))[:60],
            "%.4f" % d["host_queue_ms_per_step"] if "host_queue_ms_per_step" in d else "-"))
        print("   value %.3f M/s  serial %s  e2e %.3f M/s  ms/step %.3f  n_gpus %d  launches %s  ok %s gather %s" % (
            d["value"] / 1e6, "%.3f" % (d["serial"]["value"] / 1e6) if "serial" in d else "-", d["e2e"]["value"] / 1e6,
            d["ms_per_step"], d["n_gpus"], d.get("gpu_launches"), d.get("results_ok"), d.get("gather_ok")))
        if "e2e_ticks" in d:
            print("   e2e from tick records %.3f M/s (%d B/robot up)" % (d["e2e_ticks"]["value"] / 1e6, 272))
        if r:
            print("   kernel: %s  %.3f ms  frac %.2e  classes alone %s  per class %s" % (
                r.get("kernel"), r.get("kernel_ms", 0), r.get("frac", 0),
                ["%.3f" % x for x in r.get("class_kernel_ms_alone", [])], d["config"].get("problems_per_class")))
        p = d.get("parity")
        if p:
            print("   parity: %s" % {k: v for k, v in p.items() if k != "criterion"})
        c = d.get("cpu_baseline")
        if c:
            print("   cpu: %.1f solves/s on %s cores (%s)" % (c["value"], c["cores"], c["kind"]))
