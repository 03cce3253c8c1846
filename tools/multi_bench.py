#!/usr/bin/env python
"""upb_multi (the GPUs of one box behind one handle INSIDE the library: one process, one worker thread and context per GPU) on the
headline workload — config 3 of BASELINE.json — next to one context: the tables must be identical, the time per call is the wall
clock of upb_multi_call_regions (host arrays resident on the devices).

    python tools/multi_bench.py --gpus N [--tags T]
"""
import argparse
import hashlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import unipeak_b200  # noqa: E402
from unipeak_b200.capi import MultiContext  # noqa: E402
from unipeak_b200.synth import HG19_SIZES  # noqa: E402
from unipeak_b200.synth_det import make_synthetic_det  # noqa: E402

COLS = ("buffer", "forward_label", "dropped", "contig", "left", "len", "peak", "hit_sum", "peak_score", "kurtosis", "corr", "counts", "tags_in_regions", "unit_candidates")


def sha(cols):
    h = hashlib.sha256()
    for k in COLS:
        h.update(np.ascontiguousarray(cols[k]).view(np.uint8).tobytes())
    return h.hexdigest()[:16]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=2)
    ap.add_argument("--samples", type=int, default=16)
    ap.add_argument("--tags", type=int, default=30_000_000)
    ap.add_argument("--steps", type=int, default=10)
    args = ap.parse_args()
    G = int(sum(HG19_SIZES))
    data = make_synthetic_det(HG19_SIZES, args.samples, args.tags, directional=True, seed=20261018, device="cuda:0")
    st = dict(bandwidth=50, region_threshold=25.0, kurtosis_threshold=50.0, corr_threshold=-1.0, hit_threshold=10.0 * args.samples, directional=True,
              background=sum(data.tags(s) for s in range(args.samples)) / G / 2)
    one = unipeak_b200.Context(0)
    one.load(data)
    one.call_regions(**st, rows=False)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        want = one.call_regions(**st, rows=False)
    one_ms = (time.perf_counter() - t0) / args.steps * 1e3
    want_sha = sha(want["columns"])
    one.close()
    m = MultiContext(args.gpus)
    t0 = time.perf_counter()
    m.load(data, packed=True)
    load_ms = (time.perf_counter() - t0) * 1e3
    m.call_regions(**st, rows=False)
    m.call_regions(**st, rows=False)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        got = m.call_regions(**st, rows=False)
    multi_ms = (time.perf_counter() - t0) / args.steps * 1e3
    same = sha(got["columns"]) == want_sha
    print(json.dumps({"config": "c3 through upb_multi (one process, %d GPUs): %d samples x %d tags, hg19" % (args.gpus, args.samples, args.tags), "n_gpus": args.gpus,
                      "ms_per_call_multi": multi_ms, "ms_per_call_one_gpu": one_ms, "load_ms_multi_incl_pack": load_ms, "rows": int(got["columns"]["left"].shape[0]),
                      "identical_to_one_gpu": bool(same), "note": "wall clock around the Python binding (includes copying the table's columns out of the shared segment)"}))
    m.close()
    if not same:
        sys.exit(1)


if __name__ == "__main__":
    main()
