"""Multi-GPU plumbing of the region-calling path (one process per GPU, torch.distributed).

The path shards by (ProfileBuffer, contig) with no data-path exchange: every rank loads and
calls only its own contigs.  Three small, latency-bound collectives remain:
  * per-sample tag totals -> background (src/regions.cpp:190-210), all-reduce before the call;
  * per-sample tags-in-regions totals and per-(buffer, contig) candidate counts, all-reduce after it;
  * the gather of the region rows to rank 0, which puts them in canonical order (buffer, contig, left).
Works on any backend (NCCL on the GPUs; gloo in the CPU tests).

The gather on GPUs is SharedRegionTable: the table lives in ONE host segment shared by the ranks and mapped into every GPU; each
rank's emit kernel writes its rows there directly over its own PCIe link after a device-side exchange of the per-unit row counts
(upb_call_regions_shared) — the destination is host memory either way, eight links beat one, and no collective runs in a step.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import numpy as np


def shard_contigs(sizes: Sequence[int], world: int) -> List[int]:
    """Longest-processing-time assignment of whole contigs to ranks (tag density is roughly uniform).
    Returns owner[c]."""
    order = sorted(range(len(sizes)), key=lambda c: (-int(sizes[c]), c))
    load = [0] * world
    owner = [0] * len(sizes)
    for c in order:
        r = min(range(world), key=lambda k: (load[k], k))
        owner[c] = r
        load[r] += int(sizes[c])
    return owner


ROW_FIELDS = ("buffer", "contig", "left", "len", "peak", "cand_index", "hit_sum")
FP_FIELDS = ("kurtosis", "corr", "peak_score")


def pack_rows(cols, n_samples: int) -> np.ndarray:
    """One int64 matrix per rank: integer fields, FP64 fields as bit patterns, then the per-sample counts."""
    R = int(cols["left"].shape[0])
    width = len(ROW_FIELDS) + len(FP_FIELDS) + n_samples
    pack = np.zeros((R, width), dtype=np.int64)
    for j, k in enumerate(ROW_FIELDS):
        pack[:, j] = cols[k].astype(np.int64)
    for j, k in enumerate(FP_FIELDS):
        pack[:, len(ROW_FIELDS) + j] = np.ascontiguousarray(cols[k], dtype=np.float64).view(np.int64)
    pack[:, len(ROW_FIELDS) + len(FP_FIELDS):] = cols["counts"]
    return pack


def unpack_rows(pack: np.ndarray, n_samples: int) -> Dict[str, np.ndarray]:
    out = {}
    for j, k in enumerate(ROW_FIELDS):
        out[k] = pack[:, j].copy()
    for j, k in enumerate(FP_FIELDS):
        out[k] = pack[:, len(ROW_FIELDS) + j].copy().view(np.float64)
    out["counts"] = pack[:, len(ROW_FIELDS) + len(FP_FIELDS):].astype(np.uint32)
    return out


def gather_region_rows(cols, n_samples: int, tags_in_regions, unit_candidates, *, dist, rank: int, world: int, device) -> Optional[Dict]:
    """Collects every rank's canonical rows on rank 0.  Returns None on the other ranks."""
    import torch
    pack = pack_rows(cols, n_samples)
    R = pack.shape[0]
    tir = torch.as_tensor(np.asarray(tags_in_regions, dtype=np.int64)).to(device)
    uc = torch.as_tensor(np.asarray(unit_candidates, dtype=np.int64)).to(device)
    nrows = torch.tensor([R], dtype=torch.int64, device=device)
    allrows = [torch.zeros(1, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_reduce(tir)
    dist.all_reduce(uc)
    dist.all_gather(allrows, nrows)
    counts = [int(x.item()) for x in allrows]
    rmax = max(counts + [1])
    padded = np.zeros((rmax, pack.shape[1]), dtype=np.int64)
    padded[:R] = pack
    t = torch.from_numpy(padded).to(device)
    gathered = [torch.empty_like(t) for _ in range(world)] if rank == 0 else None
    dist.gather(t, gathered, dst=0)
    if rank != 0:
        return None
    rows = np.concatenate([g.cpu().numpy()[:k] for g, k in zip(gathered, counts)], axis=0)
    order = np.lexsort((rows[:, 2], rows[:, 1], rows[:, 0]))       # canonical: buffer, contig, left
    out = unpack_rows(rows[order], n_samples)
    out.update(n_rows=int(rows.shape[0]), tags_in_regions=tir.cpu().numpy(), unit_candidates=uc.cpu().numpy())
    return out


class SharedRegionTable:
    """The table-wide host segment of upb_call_regions_shared, for ranks that are PROCESSES (torchrun): a POSIX shared-memory file
    created exclusively (O_CREAT | O_EXCL, mode 0600) under a random name that rank 0 broadcasts, mapped and page-locked by every rank
    (upb_host_register) and unlinked as soon as all ranks hold it — nothing is left behind when a job dies, and two jobs on one host
    cannot meet.  Each rank's GPU writes its rows straight to their place in table order; there is no collective in a step.
    `dist` is only used here, at set-up.  Ranks that are threads of one process can pass `local=True` (pinned memory from
    upb_host_alloc, shared by address)."""

    def __init__(self, ctx, n_samples: int, n_contigs: int, *, rank: int, world: int, row_capacity: int, dist=None, local_ptr: int = 0):
        import mmap
        import os
        import uuid
        from .capi import SharedTable, UpbError, host_register, lib, shared_table_bytes
        self.ctx, self.rank, self.world, self.dist = ctx, rank, world, dist
        self.nbytes = shared_table_bytes(n_contigs, n_samples, world, row_capacity)
        self._mm = None
        self._registered = False
        if local_ptr:
            self.host_ptr = self.dev_ptr = int(local_ptr)
        else:
            names = [None]
            if rank == 0:
                names[0] = "/dev/shm/upb_%s" % uuid.uuid4().hex
                fd = os.open(names[0], os.O_CREAT | os.O_EXCL | os.O_RDWR, 0o600)
                os.ftruncate(fd, self.nbytes)
            if world > 1:
                dist.broadcast_object_list(names, src=0)
            if rank != 0:
                fd = os.open(names[0], os.O_RDWR)
            self._mm = mmap.mmap(fd, self.nbytes)
            os.close(fd)
            self._arr = np.frombuffer(self._mm, dtype=np.uint8)
            self.host_ptr = self._arr.ctypes.data
            if world > 1:
                dist.barrier()
            if rank == 0:
                os.unlink(names[0])                                    # the mappings keep the memory alive
            self.dev_ptr = host_register(self.host_ptr, self.nbytes)
            self._registered = True
        self.table = SharedTable(self.host_ptr, self.dev_ptr, self.nbytes, rank, world, n_contigs, n_samples, row_capacity)
        if rank == 0 and not local_ptr:
            rc = lib().upb_shared_table_init(self.table)
            if rc:
                raise UpbError(rc, "upb_shared_table_init")
        if world > 1 and dist is not None and not local_ptr:
            dist.barrier()
        self.step = 0

    def __call__(self, **settings):
        """One table-wide call; every rank calls with the same settings.  Rank 0 gets the whole table, the others their row count."""
        self.step += 1
        return self.ctx.call_regions_shared(self.table, self.step, **settings)

    def close(self):
        from .capi import host_unregister
        if self._registered:
            host_unregister(self.host_ptr)
            self._registered = False
        self._arr = None
        if self._mm is not None:
            try:
                self._mm.close()
            except BufferError:
                pass
            self._mm = None


# ------------------------------------------------------------------------------------------------
# bin/strand_shift over contig-sharded ranks (include/unipeak_b200.h "upb_shift_*")
# ------------------------------------------------------------------------------------------------
def allgather_rows(mat: np.ndarray, *, dist, world: int, device) -> List[np.ndarray]:
    """All-gather of int64 matrices with the same number of columns and any number of rows: [rank 0's, rank 1's, ...]."""
    import torch
    mat = np.ascontiguousarray(mat, dtype=np.int64)
    if world == 1:
        return [mat]
    n = torch.tensor([mat.shape[0]], dtype=torch.int64, device=device)
    alln = [torch.zeros(1, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(alln, n)
    counts = [int(x.item()) for x in alln]
    padded = np.zeros((max(counts + [1]), mat.shape[1]), dtype=np.int64)
    padded[:mat.shape[0]] = mat
    t = torch.from_numpy(padded).to(device)
    parts = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(parts, t)
    return [p.cpu().numpy()[:k] for p, k in zip(parts, counts)]


def strand_shift_sharded(ctx, *, dist, rank: int, world: int, device, bandwidth=50, region_threshold=25.0, kurtosis_threshold=50.0,
                         corr_threshold=0.3, hit_threshold=10, min_shift=25, max_shift=150, n_regions_to_test=1000, background: float) -> Dict:
    """src/strand_shift.cpp:139-259 with the contigs spread over the ranks; `background` comes from the genome-wide tag total
    (one all-reduce by the caller, as for bin/regions).  Every rank returns the same report, in the format of Context.strand_shift;
    with world == 1 it is that report.  `ctx` needs shift_regions() and shift_sweep() (capi.Context)."""
    from .capi import UpbError, shift_examine_order, shift_mode
    params = dict(bandwidth=bandwidth, region_threshold=region_threshold, kurtosis_threshold=kurtosis_threshold, corr_threshold=corr_threshold,
                  hit_threshold=hit_threshold, min_shift=min_shift, max_shift=max_shift, n_regions_to_test=n_regions_to_test, background=background)
    # 1 + 2: the reference's Region vector (one non-directional buffer: contig order, then position), known to every rank
    loc = ctx.shift_regions(**params)
    mat = np.stack([loc[k].astype(np.int64) for k in ("contig", "left", "len", "sum", "replayed")], axis=1) if loc["left"].shape[0] else np.zeros((0, 5), np.int64)
    parts = allgather_rows(mat, dist=dist, world=world, device=device)
    owner = np.concatenate([np.full(p.shape[0], r, dtype=np.int64) for r, p in enumerate(parts)])
    glob = np.concatenate(parts, axis=0)
    o = np.lexsort((glob[:, 1], glob[:, 0]))
    glob, owner = glob[o], owner[o]
    contig, left, length, rsum, replayed = (glob[:, j] for j in range(5))
    R = int(glob.shape[0])
    n_freq = max_shift + 1
    freq = np.zeros(n_freq, dtype=np.uint64)
    to_test = int(n_regions_to_test) & 0xFFFF                       # UShort
    out = dict(status=0, n_regions_found=R, n_regions_tested=0, n_tags_in_regions=0, best_shift=0, shift_freq=[], tested=[])

    # 3: examination order = the head of the reference's std::sort by Region::sum(), grown on demand exactly as upb_strand_shift grows it
    order, sorted_end = shift_examine_order(rsum.astype(np.uint32), min(R, 2 * to_test + 256))
    tested, cursor, tags = 0, 0, 0
    while tested < to_test and cursor < R:
        batch = []
        want = (to_test - tested) + (to_test - tested) // 4 + 8
        while cursor < R and len(batch) < want:
            if cursor >= sorted_end:
                order, sorted_end = shift_examine_order(rsum.astype(np.uint32), min(R, max(2 * sorted_end, cursor + 1024)))
            i = int(order[cursor])
            if length[i] > 2 * max_shift + 3:                         # :206 large enough to try all shifts
                batch.append(i)                                       # (a region of the left-edge replay is swept over the replay's own score vectors)
            cursor += 1
        if not batch:
            break
        # 4: every rank sweeps the regions of the batch that lie on its contigs; (batch slot, shift, corr bits) go round
        b = np.asarray(batch, dtype=np.int64)
        slots = np.nonzero(owner[b] == rank)[0]
        mine = b[slots]
        bs_m, bc_m = ctx.shift_sweep(contig[mine], left[mine], length[mine], **params)
        res = np.stack([slots.astype(np.int64), bs_m.astype(np.int64), bc_m.view(np.int64)], axis=1) if slots.shape[0] else np.zeros((0, 3), np.int64)
        allres = np.concatenate(allgather_rows(res, dist=dist, world=world, device=device), axis=0)
        bs = np.zeros(b.shape[0], dtype=np.int64)
        bc = np.zeros(b.shape[0], dtype=np.float64)
        bs[allres[:, 0]] = allres[:, 1]
        bc[allres[:, 0]] = allres[:, 2].copy().view(np.float64)
        # 5: :220-225 in examination order
        for k in range(b.shape[0]):
            if tested >= to_test:
                break
            if bc[k] >= corr_threshold:                               # NaN fails, as in the reference
                i = int(b[k])
                freq[int(bs[k])] += 1
                tags += int(rsum[i])
                out["tested"].append(dict(contig=int(contig[i]), left=int(left[i]), len=int(length[i]), best_shift=int(bs[k]), best_corr=float(bc[k])))
                tested += 1
    out["n_regions_tested"] = tested
    out["n_tags_in_regions"] = tags
    out["shift_freq"] = [int(x) for x in freq]
    if tested == 0:                                                   # :228-230 "no regions qualified with given settings"
        out["status"] = -8
        out["shift_freq"] = [0] * n_freq
        return out
    out["best_shift"] = shift_mode(freq, min_shift)
    return out
