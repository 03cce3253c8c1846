#!/usr/bin/env python
"""Groups-per-call sweep of the two-lane region call (UPB_CALL_GROUPS, read by the library at every call): C3 resident — whole, and the
shard rank 0 of a contig-sharded job of 2 / 4 / 8 ranks would hold — call time (wall inside the library, device first-to-last event)
per setting and the table fingerprint, which must not move within a workload."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import unipeak_b200  # noqa: E402
from ab import fingerprint, G  # noqa: E402
from unipeak_b200.multi import shard_contigs  # noqa: E402
from unipeak_b200.synth import HG19_SIZES  # noqa: E402
from unipeak_b200.synth_det import make_synthetic_det  # noqa: E402


def main():
    groups = [int(x) for x in (sys.argv[1:] or ["1", "0", "2", "1", "0"])]
    steps = int(os.environ.get("AB_STEPS", "15"))
    worlds = [int(x) for x in os.environ.get("AB_WORLDS", "1,2,4,8").split(",")]
    ctx = unipeak_b200.Context(0)
    bg = 16 * 30_000_000 / G / 2
    for world in worlds:
        owner = shard_contigs(HG19_SIZES, world)
        mine = [c for c in range(len(HG19_SIZES)) if owner[c] == 0]
        data = make_synthetic_det(HG19_SIZES, 16, 30_000_000, directional=True, seed=20261018, bandwidth=200, device="cuda:0", contigs=mine if world > 1 else None)
        ctx.load(data)
        for bw in ((50, 10, 200) if world == 1 else (50,)):
            st = dict(bandwidth=bw, region_threshold=25.0, kurtosis_threshold=50.0, corr_threshold=-1.0, hit_threshold=160.0, directional=True, background=bg)
            for g in (groups if world == 1 else [x for x in (1, 0, 2) if x in groups]):
                # g = 0: one group, its statistics + emit tail in two halves (UPB_SPLIT_TAIL); g >= 1: that many groups of the whole call, tail not split
                os.environ["UPB_CALL_GROUPS"] = str(max(g, 1))
                os.environ["UPB_SPLIT_TAIL"] = "2" if g == 0 else "0"
                for _ in range(3):
                    ctx.call_regions_view(**st)
                call, dev, kde = [], [], []
                for _ in range(steps):
                    v = ctx.call_regions_view(**st)
                    s = ctx.stats()
                    call.append(s["ms_call"]); dev.append(s["ms_total"]); kde.append(s["ms_kde"])
                print(json.dumps({"shard_of": world, "bandwidth": bw, "groups": g, "call_ms": round(float(np.mean(call)), 4), "call_ms_min": round(float(np.min(call)), 4),
                                  "device_ms": round(float(np.mean(dev)), 4), "span_brackets_ms": round(float(np.mean(kde)), 4), "launches": s["launches"], "regions": v.n_rows,
                                  "sha": fingerprint(v)}), flush=True)
                ctx.release_view()
        del data
    ctx.close()


if __name__ == "__main__":
    main()
