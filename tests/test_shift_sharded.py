"""bin/strand_shift over contig-sharded ranks: the host-only C entry points, and the protocol of
unipeak_b200.multi.strand_shift_sharded over a world_size-2 gloo group against the same protocol on one rank.
The device side (upb_shift_regions / upb_shift_sweep against upb_strand_shift) is tests/test_zz_shift_sharded_gpu.py."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from unipeak_b200 import capi
from unipeak_b200.multi import allgather_rows, shard_contigs, strand_shift_sharded

SIZES = [900_000, 500_000, 400_000, 300_000, 100_000]
PARAMS = dict(bandwidth=50, region_threshold=25.0, kurtosis_threshold=50.0, corr_threshold=0.3, hit_threshold=10, min_shift=25, max_shift=150,
              n_regions_to_test=40, background=0.01)


class FakeCtx:
    """Stands in for capi.Context on machines without a GPU: a seeded Region vector (95 % of the regions too short for the sweep, sums
    with many ties) and a sweep whose answer is a pure function of the region, so that any sharding must reproduce the one-rank report."""

    def __init__(self, contigs, seed=5, per_contig=1500):
        self.contigs = list(contigs)
        self.seed, self.per_contig = seed, per_contig

    def shift_regions(self, **p):
        cols = {k: [] for k in ("contig", "left", "len", "sum", "replayed")}
        for c in self.contigs:
            rng = np.random.default_rng(self.seed * 1000 + c)
            left = np.sort(rng.choice(SIZES[c] - 2000, size=self.per_contig, replace=False)).astype(np.uint32) + 200
            long_ = rng.random(self.per_contig) < 0.05
            length = np.where(long_, rng.integers(2 * p["max_shift"] + 4, 900, self.per_contig), rng.integers(1, 2 * p["max_shift"] + 4, self.per_contig)).astype(np.uint32)
            cols["contig"].append(np.full(self.per_contig, c, np.uint32)); cols["left"].append(left); cols["len"].append(length)
            cols["sum"].append(rng.integers(10, 60, self.per_contig).astype(np.uint32)); cols["replayed"].append(np.zeros(self.per_contig, np.uint8))
        return {k: (np.concatenate(v) if v else np.zeros(0, np.uint8 if k == "replayed" else np.uint32)) for k, v in cols.items()}

    def shift_sweep(self, contig, left, length, **p):
        assert all(int(c) in self.contigs for c in contig), "asked to sweep a region of another rank's contig"
        assert np.all(np.asarray(length) > 2 * p["max_shift"] + 3)
        h = (np.asarray(contig, np.uint64) * np.uint64(2654435761) + np.asarray(left, np.uint64) * np.uint64(40503)) % np.uint64(1 << 31)
        return (h % np.uint64(p["max_shift"] + 1)).astype(np.uint32), ((h >> np.uint64(8)) % np.uint64(1000)).astype(np.float64) / 1000.0


def test_examine_order_prefix_is_a_prefix_of_the_full_sort():
    rng = np.random.default_rng(3)
    for n, hi in ((0, 5), (1, 5), (17, 3), (5000, 40), (20000, 7)):
        sums = rng.integers(0, hi, n).astype(np.uint32)
        full, end_full = capi.shift_examine_order(sums, n)
        assert end_full == n and sorted(full.tolist()) == list(range(n))
        assert np.all(np.diff(sums[full].astype(np.int64)) <= 0)
        for want in (0, 1, 16, 17, 300, n):
            order, end = capi.shift_examine_order(sums, want)
            assert end >= min(want, n) and sorted(order.tolist()) == list(range(n))
            assert np.array_equal(order[:end], full[:end]), (n, hi, want)


def test_shift_mode_matches_a_restatement():
    # src/strand_shift.cpp:239-258 with Kernel(5) (misc/kernel.cpp:16-35), restated in numpy with the same summation order
    rng = np.random.default_rng(9)
    taps = capi.kernel_taps(5, 1.0)
    for _ in range(20):
        m = int(rng.integers(12, 200))
        freq = rng.integers(0, 30, m).astype(np.uint64)
        min_shift = int(rng.integers(0, m))
        dens = np.zeros(m)
        for i in range(m):
            for j in range(11):
                k = i - 5 + j
                if 0 <= k < m:
                    dens[k] = dens[k] + float(freq[i]) * taps[j]
        best, bd = 0, 0.0
        for i in range(min_shift, m - 5):
            if dens[i] > bd:
                best, bd = i, dens[i]
        assert capi.shift_mode(freq, min_shift) == best


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        owner = shard_contigs(SIZES, world)
        ctx = FakeCtx([c for c in range(len(SIZES)) if owner[c] == rank])
        parts = allgather_rows(np.full((rank + 1, 2), rank, np.int64), dist=dist, world=world, device=torch.device("cpu"))
        assert [p.shape for p in parts] == [(r + 1, 2) for r in range(world)] and all(np.all(p == r) for r, p in enumerate(parts))
        q.put((rank, strand_shift_sharded(ctx, dist=dist, rank=rank, world=world, device=torch.device("cpu"), **PARAMS)))
    finally:
        dist.destroy_process_group()


def test_sharded_report_equals_one_rank_report_gloo():
    want = strand_shift_sharded(FakeCtx(range(len(SIZES))), dist=None, rank=0, world=1, device=None, **PARAMS)
    assert want["status"] == 0 and want["n_regions_tested"] == 40 and want["n_regions_found"] == 5 * 1500
    assert len({(t["contig"]) for t in want["tested"]}) > 1          # the tested regions come from several contigs
    world = 2
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mpctx = mp.get_context("spawn")
    q = mpctx.Queue()
    procs = [mpctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=180) for _ in range(world))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for r in range(world):
        assert got[r] == want, "rank %d reports something else than one rank would" % r


def test_sharded_no_region_qualifies():
    p = dict(PARAMS, corr_threshold=2.0)                              # nothing reaches a correlation of 2
    out = strand_shift_sharded(FakeCtx(range(len(SIZES))), dist=None, rank=0, world=1, device=None, **p)
    assert out["status"] == -8 and out["n_regions_tested"] == 0 and out["tested"] == [] and sum(out["shift_freq"]) == 0
