#!/usr/bin/env python
"""Groups-per-call sweep of the two-lane region call (UPB_CALL_GROUPS, read by the library at every call): C3 resident, call time
(wall inside the library, device first-to-last event) per setting and the table fingerprint, which must not move."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import unipeak_b200  # noqa: E402
from ab import fingerprint, G  # noqa: E402
from unipeak_b200.synth import HG19_SIZES  # noqa: E402
from unipeak_b200.synth_det import make_synthetic_det  # noqa: E402


def main():
    groups = [int(x) for x in (sys.argv[1:] or ["1", "2", "3", "4", "6", "8", "1", "2"])]
    steps = int(os.environ.get("AB_STEPS", "20"))
    ctx = unipeak_b200.Context(0)
    data = make_synthetic_det(HG19_SIZES, 16, 30_000_000, directional=True, seed=20261018, bandwidth=200, device="cuda:0")
    ctx.load(data)
    bg = sum(data.tags(s) for s in range(16)) / G / 2
    for bw in (50, 10, 200):
        st = dict(bandwidth=bw, region_threshold=25.0, kurtosis_threshold=50.0, corr_threshold=-1.0, hit_threshold=160.0, directional=True, background=bg)
        for g in groups:
            os.environ["UPB_CALL_GROUPS"] = str(g)
            for _ in range(3):
                ctx.call_regions_view(**st)
            call, dev, kde = [], [], []
            for _ in range(steps):
                v = ctx.call_regions_view(**st)
                s = ctx.stats()
                call.append(s["ms_call"]); dev.append(s["ms_total"]); kde.append(s["ms_kde"])
            print(json.dumps({"bandwidth": bw, "groups": g, "call_ms": float(np.mean(call)), "call_ms_min": float(np.min(call)), "device_ms": float(np.mean(dev)),
                              "span_brackets_ms": float(np.mean(kde)), "launches": s["launches"], "regions": v.n_rows, "sha": fingerprint(v)}), flush=True)
            ctx.release_view()
    ctx.close()


if __name__ == "__main__":
    main()
