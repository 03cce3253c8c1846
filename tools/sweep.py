#!/usr/bin/env python
"""BASELINE.json configs 5 (bandwidth sweep on the C3 pooled profile) and 2 (2 x 20M non-directional: strand_shift per
sample, then regions -D) on one GPU: one data set per workload, timings from CUDA events / wall clock around the C ABI
calls.  Prints one JSON line per measurement (kept under profiles/)."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import unipeak_b200  # noqa: E402
from unipeak_b200.synth import HG19_SIZES, MM9_SIZES  # noqa: E402
from unipeak_b200.synth_det import make_synthetic_det as make_synthetic_gpu  # noqa: E402

PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6650.0) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
G = int(sum(HG19_SIZES))


def timed(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        out = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3, out


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    ctx = unipeak_b200.Context(0)
    if which in ("all", "c5"):
        data = make_synthetic_gpu(HG19_SIZES, 16, 30_000_000, directional=True, seed=20261018, bandwidth=200, device="cuda:0")
        ctx.load(data)
        E = int(data.pos.shape[0])
        total = sum(data.tags(s) for s in range(16))
        bg = total / G / 2
        for bw in (10, 20, 50, 100, 150, 200):
            st = dict(bandwidth=bw, region_threshold=25.0, kurtosis_threshold=50.0, corr_threshold=-1.0, hit_threshold=160.0, directional=True,
                      background=bg)
            ms, v = timed(lambda: ctx.call_regions_view(**st))
            s = ctx.stats()
            print(json.dumps({"config": "c5 bandwidth sweep (C3 pooled profile, 16 x 30M, hg19)", "bandwidth": bw, "ms_per_step": ms, "genome_bp_per_s": G / ms * 1e3,
                              "span_kernel_ms": s["ms_kde"], "span_kernel_GBps_algorithmic": 8 * E / s["ms_kde"] / 1e6, "frac_of_measured_hbm": 8 * E / s["ms_kde"] / 1e6 / PEAK,
                              "regions": v.n_rows, "candidates": s["n_candidates"], "warm_spans": s["warm_spans"], "hot_blocks": s["hot_blocks"], "exact_evals": s["exact_evals"]}))
        ctx.release_view()
        del data
    if which in ("all", "c2"):
        data = make_synthetic_gpu(HG19_SIZES, 2, 20_000_000, directional=False, seed=20261019, device="cuda:0")
        shifts = []
        for smp in range(2):
            one = data.subset_samples([smp])
            ctx.load(one)
            bg = one.tags(0) / G
            ms, rep = timed(lambda: ctx.strand_shift(background=bg), reps=3, warm=1)
            shifts.append(rep["best_shift"])
            print(json.dumps({"config": "c2 strand_shift, one 20M-tag ChIP sample, hg19", "sample": smp, "ms": ms, "best_shift": rep["best_shift"],
                              "regions_found": rep["n_regions_found"], "regions_tested": rep["n_regions_tested"]}))
        ctx.load(data)
        total = sum(data.tags(s) for s in range(2))
        st = dict(bandwidth=50, region_threshold=25.0, kurtosis_threshold=50.0, corr_threshold=0.3, hit_threshold=20.0, directional=False, background=total / G)
        ms, v = timed(lambda: ctx.call_regions_view(**st), reps=3, warm=1)
        s = ctx.stats()
        print(json.dumps({"config": "c2 regions -D, 2 x 20M tags, hg19 (no -s offsets applied)", "ms_per_step": ms, "genome_bp_per_s": G / ms * 1e3,
                          "span_kernel_ms": s["ms_kde"], "regions": v.n_rows, "candidates": s["n_candidates"], "estimated_shifts": shifts}))
    if which in ("all", "c4"):
        # config 4 scaled to one GPU's comfortable size: 96 samples x 5 M tags (the full 96 x 50 M is ~77 GB resident), mm9, then
        # tags_in_regions of 32 held-out samples x 5 M tags over the resulting table
        Gm = int(sum(MM9_SIZES))
        data = make_synthetic_gpu(MM9_SIZES, 96, 5_000_000, directional=True, seed=20261020, device="cuda:0")
        ctx.load(data)
        total = sum(data.tags(s) for s in range(96))
        st = dict(bandwidth=50, region_threshold=25.0, kurtosis_threshold=50.0, corr_threshold=-1.0, hit_threshold=960.0, directional=True, background=total / Gm / 2)
        ms, v = timed(lambda: ctx.call_regions_view(**st), reps=3, warm=1)
        s = ctx.stats()
        cols = v.columns()
        fwd = cols["forward_label"].astype(bool) & (cols["dropped"] == 0)
        right = cols["left"] + cols["len"] - 1
        keys = np.stack([cols["contig"], np.where(cols["forward_label"] != 0, cols["left"], right), np.where(cols["forward_label"] != 0, right, cols["left"])], axis=1)[cols["dropped"] == 0].astype(np.uint32)
        print(json.dumps({"config": "c4 scaled: regions, 96 samples x 5M tags, mm9 (full config: 96 x 50M)", "ms_per_step": ms, "genome_bp_per_s": Gm / ms * 1e3,
                          "span_kernel_ms": s["ms_kde"], "tag_entries": int(data.pos.shape[0]), "regions": v.n_rows, "candidates": s["n_candidates"]}))
        ctx.release_view()
        del data
        held = make_synthetic_gpu(MM9_SIZES, 32, 5_000_000, directional=True, seed=20261021, device="cuda:0")
        ctx.load(held)
        ms, counts = timed(lambda: ctx.count_tags_in_regions_arr(keys, directional=True, extension=0), reps=3, warm=1)
        print(json.dumps({"config": "c4 scaled: tags_in_regions, 32 held-out samples x 5M tags over the table above", "ms": ms, "rows": int(keys.shape[0]),
                          "tag_entries": int(held.pos.shape[0]), "tags_counted": int(counts.astype(np.uint64).sum())}))
    ctx.close()


if __name__ == "__main__":
    main()
