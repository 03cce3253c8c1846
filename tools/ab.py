#!/usr/bin/env python
"""A/B of experiment builds of the library (UPB_LIB=build_exp/<variant>.so python tools/ab.py [c3] [c2] [edge]).

One process per library (the C ABI is bound at import).  For every workload: the span kernel's CUDA-event time and the call time
averaged over the steps, and a SHA-256 over every column of the region table (integers and the raw bit patterns of the doubles),
so that two builds can be compared bit for bit at full size without the CPU oracle.  Prints one JSON line per workload."""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import unipeak_b200  # noqa: E402
from unipeak_b200.synth import HG19_SIZES  # noqa: E402
from unipeak_b200.synth_det import make_synthetic_det as make_synthetic_gpu  # noqa: E402

G = int(sum(HG19_SIZES))
COLS = ("buffer", "forward_label", "dropped", "contig", "left", "len", "peak", "hit_sum", "peak_score", "kurtosis", "corr", "counts",
        "tags_in_regions", "unit_candidates", "unit_rows")


def fingerprint(view) -> str:
    h = hashlib.sha256()
    cols = view.columns()
    for k in COLS:
        h.update(np.ascontiguousarray(cols[k]).view(np.uint8).tobytes())
    return h.hexdigest()[:16]


def run(ctx, name, st, steps):
    for _ in range(2):
        ctx.call_regions_view(**st)
    kde, call = [], []
    for _ in range(steps):
        v = ctx.call_regions_view(**st)
        s = ctx.stats()
        kde.append(s["ms_kde"]); call.append(s["ms_call"])
    print(json.dumps({"lib": os.path.basename(unipeak_b200.capi.LIB_PATH), "workload": name, "span_kernel_ms": float(np.mean(kde)), "span_kernel_ms_min": float(np.min(kde)),
                      "call_ms": float(np.mean(call)), "call_ms_min": float(np.min(call)), "regions": v.n_rows, "candidates": s["n_candidates"],
                      "exact_evals": s["exact_evals"], "sha": fingerprint(v)}), flush=True)
    ctx.release_view()


def main():
    which = sys.argv[1:] or ["c3"]
    steps = int(os.environ.get("AB_STEPS", "20"))
    ctx = unipeak_b200.Context(0)
    if "c3" in which:
        data = make_synthetic_gpu(HG19_SIZES, 16, 30_000_000, directional=True, seed=20261018, bandwidth=200, device="cuda:0")
        ctx.load(data)
        bg = sum(data.tags(s) for s in range(16)) / G / 2
        for bw in ((50, 10, 20, 100, 150, 200) if "bw" in which else (50,)):
            run(ctx, "c3 bw=%d" % bw, dict(bandwidth=bw, region_threshold=25.0, kurtosis_threshold=50.0, corr_threshold=-1.0, hit_threshold=160.0, directional=True, background=bg), steps)
        if "ctl" in which:                                            # two of the samples as negative controls: the isCtl branch of B1
            ctl = np.zeros(16, np.uint8); ctl[[3, 11]] = 1
            run(ctx, "c3 bw=50 controls", dict(bandwidth=50, region_threshold=25.0, kurtosis_threshold=50.0, corr_threshold=-1.0, hit_threshold=140.0, directional=True,
                                               control=ctl, background=bg * 14 / 16), max(3, steps // 4))
        del data
    if "c2" in which:
        data = make_synthetic_gpu(HG19_SIZES, 2, 20_000_000, directional=False, seed=20261019, device="cuda:0")
        ctx.load(data)
        total = sum(data.tags(s) for s in range(2))
        run(ctx, "c2 regions -D", dict(bandwidth=50, region_threshold=25.0, kurtosis_threshold=50.0, corr_threshold=0.3, hit_threshold=20.0, directional=False, background=total / G), max(3, steps // 4))
    ctx.close()


if __name__ == "__main__":
    main()
