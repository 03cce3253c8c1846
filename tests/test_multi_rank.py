"""The N > 1 host logic (contig sharding, row packing, the gather to rank 0) on CPU with the gloo backend,
world_size 2 — no GPU involved: every rank fabricates the rows its shard would have produced."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from unipeak_b200.multi import gather_region_rows, pack_rows, shard_contigs, unpack_rows
from unipeak_b200.synth import HG19_SIZES


def fake_rows(contigs, n_samples, seed):
    rng = np.random.default_rng(seed)
    rows = []
    for b in range(2):
        for c in contigs:
            left = np.sort(rng.choice(np.arange(100, 100000), size=int(rng.integers(0, 40)), replace=False))
            for i, l in enumerate(left):
                rows.append((b, c, int(l), int(rng.integers(1, 90)), int(l) + 3, i, int(rng.integers(10, 500))))
    a = np.array(rows, dtype=np.int64).reshape(len(rows), 7)
    R = a.shape[0]
    cols = dict(buffer=a[:, 0].astype(np.uint8), contig=a[:, 1].astype(np.uint32), left=a[:, 2].astype(np.uint32), len=a[:, 3].astype(np.uint32),
                peak=a[:, 4].astype(np.uint32), cand_index=a[:, 5].astype(np.uint64), hit_sum=a[:, 6].astype(np.uint32),
                kurtosis=rng.random(R), corr=np.full(R, np.nan), peak_score=rng.random(R) * 100,
                counts=rng.integers(0, 50, size=(R, n_samples)).astype(np.uint32))
    return cols


def test_shard_contigs_balances_hg19():
    for world in (1, 2, 4, 8):
        owner = shard_contigs(HG19_SIZES, world)
        load = [sum(s for s, o in zip(HG19_SIZES, owner) if o == r) for r in range(world)]
        assert sum(load) == sum(HG19_SIZES)
        assert max(load) <= 1.12 * sum(HG19_SIZES) / world + max(HG19_SIZES) * (world >= 8)
        assert sorted(set(owner)) == list(range(world))


def test_pack_roundtrip():
    cols = fake_rows([0, 2], 5, 1)
    back = unpack_rows(pack_rows(cols, 5), 5)
    for k in ("buffer", "contig", "left", "len", "peak", "cand_index", "hit_sum"):
        assert np.array_equal(back[k], cols[k].astype(np.int64))
    assert np.array_equal(back["kurtosis"].view(np.uint64), cols["kurtosis"].view(np.uint64))
    assert np.all(np.isnan(back["corr"]))
    assert np.array_equal(back["counts"], cols["counts"])


def _worker(rank, world, port, n_samples, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        owner = shard_contigs([900, 500, 400, 300, 100], world)
        mine = [c for c in range(5) if owner[c] == rank]
        cols = fake_rows(mine, n_samples, 100 + rank)
        tir = cols["counts"].sum(axis=0).astype(np.int64)
        uc = np.zeros(10, np.int64)
        for b, c in zip(cols["buffer"], cols["contig"]):
            uc[int(b) * 5 + int(c)] += 1
        out = gather_region_rows(cols, n_samples, tir, uc, dist=dist, rank=rank, world=world, device=torch.device("cpu"))
        if rank == 0:
            q.put({k: (v if not isinstance(v, np.ndarray) else v.copy()) for k, v in out.items()})
        else:
            assert out is None
    finally:
        dist.destroy_process_group()


def test_gather_two_ranks_gloo():
    world, n_samples = 2, 3
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_samples, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    # what a single process would have produced: both shards' rows, canonical order
    owner = shard_contigs([900, 500, 400, 300, 100], world)
    parts = [fake_rows([c for c in range(5) if owner[c] == r], n_samples, 100 + r) for r in range(world)]
    allc = {k: np.concatenate([p[k] for p in parts]) for k in parts[0]}
    order = np.lexsort((allc["left"], allc["contig"], allc["buffer"]))
    assert out["n_rows"] == order.shape[0]
    for k in ("buffer", "contig", "left", "len", "peak", "hit_sum"):
        assert np.array_equal(out[k], allc[k][order].astype(np.int64)), k
    assert np.array_equal(out["counts"], allc["counts"][order])
    assert np.array_equal(out["kurtosis"].view(np.uint64), allc["kurtosis"][order].view(np.uint64))
    assert out["tags_in_regions"].tolist() == allc["counts"].sum(axis=0).tolist()
    key = out["buffer"] * 10 ** 12 + out["contig"] * 10 ** 8 + out["left"]
    assert np.all(np.diff(key) > 0)
