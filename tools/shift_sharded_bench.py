#!/usr/bin/env python
"""bin/strand_shift on N GPUs (BASELINE.json config 2: one non-directional ChIP-seq sample, 20 M tags, full hg19), contigs sharded
over the ranks (unipeak_b200.multi.strand_shift_sharded over NCCL), checked against upb_strand_shift on one GPU holding everything.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/shift_sharded_bench.py [--tags T]

Prints one JSON line on rank 0: ms per calibration (wall clock around the protocol, max over ranks, after a warm-up call), the
single-GPU time of the same call on rank 0, and whether the two reports are identical (they must be: same regions tested in the
same order, same doubles).  Written in round 1 after the GPU budget was spent: not yet run on hardware."""
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
from unipeak_b200.multi import shard_contigs, strand_shift_sharded  # noqa: E402
from unipeak_b200.synth import HG19_SIZES  # noqa: E402
from unipeak_b200.synth_det import make_synthetic_det as make_synthetic_gpu  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tags", type=int, default=20_000_000)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--seed", type=int, default=20261019)
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    sizes = HG19_SIZES
    G = int(sum(sizes))
    owner = shard_contigs(sizes, world)
    mine = [c for c in range(len(sizes)) if owner[c] == rank]
    data = make_synthetic_gpu(sizes, 1, args.tags, directional=False, seed=args.seed, contigs=mine, device=str(dev))
    tags = torch.tensor([data.tags(0)], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(tags)                                        # genome-wide tag total -> background (src/strand_shift.cpp:133)
    bg = float(tags.item()) / G
    ctx = unipeak_b200.Context(local)
    ctx.load(data)

    def sharded():
        return strand_shift_sharded(ctx, dist=dist if world > 1 else None, rank=rank, world=world, device=dev, background=bg)

    rep = sharded()                                                   # warm-up: scratch buffers, tile index
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.reps):
        rep = sharded()
    torch.cuda.synchronize()
    dt = torch.tensor([(time.perf_counter() - t0) / args.reps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)

    line = None
    if rank == 0:
        # the same sample whole, on this rank's GPU: what one GPU reports
        whole = make_synthetic_gpu(sizes, 1, args.tags, directional=False, seed=args.seed, device=str(dev))
        one = unipeak_b200.Context(local)
        one.load(whole)
        one.strand_shift(background=bg)
        t1 = time.perf_counter()
        for _ in range(args.reps):
            want = one.strand_shift(background=bg)
        single_ms = (time.perf_counter() - t1) / args.reps * 1e3
        one.close()
        same = all(rep[k] == want[k] for k in ("status", "n_regions_found", "n_regions_tested", "n_tags_in_regions", "best_shift", "shift_freq", "tested"))
        line = {"config": "c2 strand_shift, one non-directional sample, %d tags, hg19" % args.tags, "n_gpus": world, "ms_sharded": float(dt.item()) * 1e3,
                "ms_one_gpu": single_ms, "identical_to_one_gpu": bool(same), "best_shift": rep["best_shift"], "regions_found": rep["n_regions_found"],
                "regions_tested": rep["n_regions_tested"]}
        print(json.dumps(line))
    ctx.close()
    if world > 1:
        dist.destroy_process_group()
    if line is not None and not line["identical_to_one_gpu"]:
        sys.exit(1)


if __name__ == "__main__":
    main()
