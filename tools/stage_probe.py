#!/usr/bin/env python
"""Per-stage device times of upb_call_regions on config 3 for a list of bandwidths (upb_set_profiling)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import unipeak_b200
from unipeak_b200.synth import HG19_SIZES
from unipeak_b200.synth_det import make_synthetic_det

n, tags = 16, int(os.environ.get("TAGS", 30_000_000))
data = make_synthetic_det(HG19_SIZES, n, tags, directional=True, seed=20261018, device="cuda:0")
G = int(sum(HG19_SIZES))
ctx = unipeak_b200.Context(0)
ctx.load_packed2(data)
ctx.set_profiling(True)
for bw in [int(x) for x in (sys.argv[1:] or ["50", "200"])]:
    st = dict(bandwidth=bw, region_threshold=25.0, kurtosis_threshold=50.0, corr_threshold=-1.0, hit_threshold=10.0 * n, directional=True,
              background=sum(data.tags(s) for s in range(n)) / G / 2)
    for _ in range(3):
        v = ctx.call_regions_view(**st)
    acc = {}
    for _ in range(5):
        v = ctx.call_regions_view(**st)
        for name, ms in ctx.stage_times():
            acc.setdefault(name, []).append(ms)
    s = ctx.stats()
    print("bw=%d rows=%d cand=%d call_ms=%.3f device_ms=%.3f" % (bw, v.n_rows, s["n_candidates"], s["ms_call"], s["ms_total"]))
    for k, vv in acc.items():
        print("   %-45s %.3f ms" % (k, float(np.mean(vv))))
    c = v.columns()
    print("   longest region %d, mean %.1f, total cells %d" % (int(c["len"].max()), float(c["len"].mean()), int(c["len"].sum())))
