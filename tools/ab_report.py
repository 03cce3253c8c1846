#!/usr/bin/env python
"""Table of tools/ab.py results: python tools/ab_report.py gpurun_out/<tag>"""
import glob, json, os, sys
out = sys.argv[1]
tab = {}
for f in sorted(glob.glob(os.path.join(out, "ab_*.jsonl"))):
    v = os.path.basename(f)[3:-6]
    tab[v] = {r["workload"]: r for r in map(json.loads, open(f)) }
vs = list(tab)
wl = []
for v in vs:
    for w in tab[v]:
        if w not in wl:
            wl.append(w)
for w in wl:
    print("%-22s" % w, "  ".join("%s %.3f/%.3f %s" % (v, tab[v].get(w, {}).get("span_kernel_ms", -1), tab[v].get(w, {}).get("call_ms", -1), tab[v].get(w, {}).get("sha", "-")[:6]) for v in vs))
