#!/usr/bin/env python
"""bin/tags_in_regions on N GPUs (the second half of BASELINE.json config 4: 32 held-out RNA-seq samples x 50 M tags, mm9), samples
sharded over the ranks — src/tags_in_regions.cpp:175-199 walks its samples one after the other and nothing couples them, so every
rank counts its own samples against the same region table and the columns are all-gathered (NCCL).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/tir_sharded_bench.py [--samples 32] [--tags 50000000]

The region table comes from upb_call_regions on two further samples (every rank derives the same table from the same seed).  Prints
one JSON line on rank 0: ms per count pass (device time, max over ranks), regions, and whether the columns of a sample that rank 0
counts again on its own GPU — next to other samples, as a one-GPU run would hold it — are identical."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import unipeak_b200  # noqa: E402
from unipeak_b200.synth import MM9_SIZES, TagData  # noqa: E402
from unipeak_b200.synth_det import make_synthetic_det  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--samples", type=int, default=32)
    ap.add_argument("--tags", type=int, default=50_000_000)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--seed", type=int, default=20261020)
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    sizes = MM9_SIZES
    G = int(sum(sizes))
    ctx = unipeak_b200.Context(local)
    # the region table: two samples pooled (same seed on every rank -> same table everywhere)
    base = make_synthetic_det(sizes, 2, args.tags, directional=True, seed=args.seed, device=str(dev))
    ctx.load(base)
    st = dict(bandwidth=50, region_threshold=25.0, kurtosis_threshold=50.0, corr_threshold=-1.0, hit_threshold=20.0, directional=True,
              background=sum(base.tags(s) for s in range(2)) / G / 2)
    cols = ctx.call_regions(**st, rows=False)["columns"]
    keep = cols["dropped"] == 0
    left = cols["left"][keep].astype(np.int64); right = left + cols["len"][keep].astype(np.int64) - 1
    fwd = cols["forward_label"][keep] != 0
    keys = np.stack([cols["contig"][keep].astype(np.int64), np.where(fwd, left, right), np.where(fwd, right, left)], axis=1).astype(np.uint32)
    del base
    # this rank's held-out samples, by global sample number: the set does not depend on the number of ranks
    lo, hi = args.samples * rank // world, args.samples * (rank + 1) // world

    def sample(k):                                                   # sample 2 + k of the series whose samples 0 and 1 made the table: same peaks, own tags
        return make_synthetic_det(sizes, 1, args.tags, directional=True, seed=args.seed, device=str(dev), first_sample=2 + k)

    def concat(parts):
        segs = [[[p.seg(0, d, c) for c in range(len(sizes))] for d in range(2)] for p in parts]
        return TagData.from_segments(sizes, segs)

    mine = concat([sample(k) for k in range(lo, hi)]) if hi > lo else None
    counts = np.zeros((keys.shape[0], max(hi - lo, 1)), dtype=np.uint32)
    ms = 0.0
    if mine is not None:
        ctx.load(mine)
        ctx.count_tags_in_regions_arr(keys, directional=True)        # warm-up
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.reps):
            counts = ctx.count_tags_in_regions_arr(keys, directional=True)
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) / args.reps * 1e3
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    per = args.samples // world + 1
    padded = np.zeros((keys.shape[0], per), dtype=np.int64)
    padded[:, :hi - lo] = counts[:, :hi - lo]
    mat = torch.from_numpy(padded).to(dev)
    parts = [torch.empty_like(mat) for _ in range(world)]
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_gather(parts, mat)                                   # the counts matrix, column blocks by rank
    else:
        parts = [mat]
    if rank == 0:
        full = np.concatenate([parts[r].cpu().numpy()[:, :args.samples * (r + 1) // world - args.samples * r // world] for r in range(world)], axis=1)
        # cross-check: the last sample (another rank's when N > 1) counted here, loaded next to sample 0 as a one-GPU run would hold them
        k = args.samples - 1
        pair = concat([sample(0), sample(k)]) if k > 0 else sample(0)
        ctx.load(pair)
        chk = ctx.count_tags_in_regions_arr(keys, directional=True)
        same = bool(np.array_equal(chk[:, 0], full[:, 0]) and np.array_equal(chk[:, -1], full[:, k]))
        print(json.dumps({"config": "c4 tags_in_regions: %d held-out samples x %d tags, mm9, samples sharded over the ranks" % (args.samples, args.tags),
                          "n_gpus": world, "ms_per_pass": float(t.item()), "regions": int(keys.shape[0]), "samples": args.samples,
                          "tags_counted": int(full.sum()), "columns_identical_to_one_gpu": same}))
        if not same:
            sys.exit(1)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
