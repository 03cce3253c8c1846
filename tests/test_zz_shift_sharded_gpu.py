"""bin/strand_shift over contig-sharded contexts on the device: upb_shift_regions / upb_shift_sweep and the protocol
of unipeak_b200.multi.strand_shift_sharded against upb_strand_shift (itself pinned to the oracle in test_gpu_parity.py) — with one
rank, and with the contigs split over two contexts that exchange their lists the way two ranks would."""
import threading

import numpy as np
import pytest
import torch

from refutil import oracle_strand_shift
from unipeak_b200.multi import shard_contigs, strand_shift_sharded
from unipeak_b200.synth import TagData, make_synthetic

pytestmark = pytest.mark.gpu
KW = dict(n_regions_to_test=40, region_threshold=8.0, corr_threshold=0.1)
SIZES = [1500000, 900000, 600000, 300000]


class ThreadGroup:
    """all_gather between the threads of one process (each thread is a rank with its own Context)."""

    def __init__(self, world):
        self.world = world
        self.barrier = threading.Barrier(world)
        self.slots = [None] * world
        self.rank = threading.local()

    def all_gather(self, out_list, t):
        self.slots[self.rank.value] = t.clone()
        self.barrier.wait()
        for r in range(self.world):
            out_list[r].copy_(self.slots[r])
        self.barrier.wait()


def _data():
    return make_synthetic(SIZES, 1, 500000, directional=False, seed=21, peak_spacing=10000)


def _same(got, want):
    assert got["status"] == want["status"] == 0
    for k in ("n_regions_found", "n_regions_tested", "n_tags_in_regions", "best_shift", "shift_freq"):
        assert got[k] == want[k], k
    assert len(got["tested"]) == len(want["tested"])
    for g, w in zip(got["tested"], want["tested"]):
        assert (g["contig"], g["left"], g["len"], g["best_shift"]) == (w["contig"], w["left"], w["len"], w["best_shift"])
        assert g["best_corr"] == w["best_corr"]                       # bit-identical doubles


def test_one_rank_equals_strand_shift(gpu_ctx):
    d = _data()
    ref = oracle_strand_shift(d, **KW)
    gpu_ctx.load(d)
    want = gpu_ctx.strand_shift(background=ref["background"], **KW)
    _same(want, ref)
    assert len({t["contig"] for t in want["tested"]}) > 1
    got = strand_shift_sharded(gpu_ctx, dist=None, rank=0, world=1, device=None, background=ref["background"], **KW)
    _same(got, want)
    # the pieces themselves: the region list is the reference's vector, the sweep reproduces the tested regions' maxima
    regs = gpu_ctx.shift_regions(background=ref["background"], **KW)
    assert regs["left"].shape[0] == want["n_regions_found"]
    key = regs["contig"].astype(np.int64) * (1 << 32) + regs["left"]
    assert np.all(np.diff(key) > 0)
    t = want["tested"]
    bs, bc = gpu_ctx.shift_sweep([x["contig"] for x in t], [x["left"] for x in t], [x["len"] for x in t], background=ref["background"], **KW)
    assert bs.tolist() == [x["best_shift"] for x in t] and bc.tolist() == [x["best_corr"] for x in t]


def test_two_shards_equal_one(gpu_ctx):
    import unipeak_b200
    d = _data()
    bg = d.tags(0) / d.genome_size()                                   # src/strand_shift.cpp:133 (genome-wide total: one all-reduce in a real job)
    gpu_ctx.load(d)
    want = gpu_ctx.strand_shift(background=bg, **KW)
    world = 2
    owner = shard_contigs(SIZES, world)
    empty = (np.zeros(0, np.uint32), np.zeros(0, np.uint32))
    group = ThreadGroup(world)
    results, errors = [None] * world, []

    def rank_main(rank):
        try:
            group.rank.value = rank
            segs = [[[d.seg(0, dd, c) if owner[c] == rank else empty for c in range(len(SIZES))] for dd in range(2)]]
            ctx = unipeak_b200.Context(0)
            try:
                ctx.load(TagData.from_segments(SIZES, segs))
                results[rank] = strand_shift_sharded(ctx, dist=group, rank=rank, world=world, device=torch.device("cpu"), background=bg, **KW)
            finally:
                ctx.close()
        except BaseException as e:                                     # a dead rank must not leave the other one at the barrier
            errors.append(e)
            group.barrier.abort()

    threads = [threading.Thread(target=rank_main, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=300)
    assert not errors, errors
    for r in range(world):
        _same(results[r], want)
