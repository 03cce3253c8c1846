#!/usr/bin/env python
"""bench.py — genome bp/s of pooled KDE + region calling + count matrix (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c3|c2|c1|tiny]

A step is one pass of the hot path over the resident synthetic input: upb_call_regions() — pooled
Epanechnikov KDE, thresholding, region extraction, hit/kurtosis(/correlation) filters, regions x samples
count matrix — ending with the region table in host memory in the reference's emission order.
N > 1 (torchrun, one process per GPU): contigs are sharded over the ranks (strong scaling: the genome is
fixed).  NCCL carries the per-sample tag totals (background) at set-up; inside a step there is no collective:
every rank's GPU writes its rows straight into ONE host table shared by the ranks after a device-side exchange
of the per-unit row counts (upb_call_regions_shared).

`value`   : genome bp / step time, inputs already resident in HBM (CUDA events on the library's stream around the K steps,
            bracketed by barrier + synchronize, max over ranks; the wall clock of the same region is reported beside it).
`e2e`     : the same through the C-ABI from pinned HOST buffers every step: upb_load_tags_async (H2D copy, decode, counting
            sort, zone maps — pipelined by (strand, contig range)) + upb_call_regions, which starts on the groups that have
            arrived; ends with the table on the host.  Host form (--host-form): the variable-length byte one (~1.5 B per
            entry + count exceptions), which is what the shipped front ends load; `pack_ms` is what its host-side encoder
            took (outside the timed region: a front end encodes while it parses).  `e2e_alt2` is the other compact form
            (u16 per entry + exception list), `e2e_alt` the plain one (u32 pos + u32 cnt, 8 B per entry).
`t_hot_cold`: sample-major arrays resident -> index build (tile-major copy + zone maps: the k-way merge's replacement)
            -> call -> table on the host (SURVEY 8d's t_hot taken literally), with its own bp/s and roofline fraction.
`roofline`: the fused span kernel, algorithmic bytes (8 B per tag entry read once + 40+4n B per candidate written)
            over its CUDA-event duration, against MEASURED_PEAKS.json hbm_gbs; `kernels` = every stage of the call the
            same way; `sweep` = the span kernel at bandwidth 10 ... 200 (BASELINE config 5).
`cpu_baseline` / --impl reference: the reference's own Kernel/ProfileBuffer/Region objects
            (oracle/_ref/ref_replay, kind "reference") — or the C restatement (kind "port") when the
            compiled reference is absent — single-threaded like the reference, on a bounded sample
            (two contigs of the same workload, same thresholds); the whole table is additionally compared with
            the fingerprints of the reference's own full-size run (tests/golden/fullsize_*.json).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from unipeak_b200.synth import HG19_NAMES, HG19_SIZES, MM9_NAMES, MM9_SIZES, TagData  # noqa: E402

WORKLOADS = {
    # name: (description, n_samples, tags_per_sample, directional, contig sizes, names, mappable)
    "c3": ("3SEQ pooled region calling, 16 samples x 30M tags, full hg19, with regions x experiments count matrix", 16, 30_000_000, True, HG19_SIZES, HG19_NAMES, 0),
    "c2": ("regions -D on 2 non-directional ChIP-seq samples, full hg19, 20M tags each", 2, 20_000_000, False, HG19_SIZES, HG19_NAMES, 0),
    "c4": ("RNA-seq mm9, 96 samples x 50M tags (regions part of config 4; sized for contig-sharding over 8 GPUs: ~77 GB resident on one)", 96, 50_000_000, True, MM9_SIZES, MM9_NAMES, 0),
    "c1": ("regions on one directional 3SEQ sample, chr21 only, 1M tags, -m whole-genome", 1, 1_000_000, True, [48129895], ["chr21"], 3095693983),
    "tiny": ("smoke-sized: 4 samples x 200k tags on 3 contigs", 4, 200_000, True, [5_000_000, 3_000_000, 1_000_000], ["cA", "cB", "cC"], 0),
}
# bounded CPU sample: chr21 + chr22 for the hg19 workloads.  Two contigs, not one: several directional samples on
# a single-contig table make the reference's own merge loop assert.
CPU_SAMPLE_CONTIGS = {"c3": [20, 21], "c2": [20, 21], "c4": [18, 20], "c1": [0], "tiny": [1, 2]}


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


from unipeak_b200.multi import SharedRegionTable, shard_contigs  # noqa: E402


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region, in-process through NVML (pynvml) every 10 ms — spawning
    nvidia-smi five times a second from every rank, as this class did before, perturbed the launch-bound multi-GPU steps.
    Falls back to nvidia-smi when pynvml is missing."""
    QUERY = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    REASONS = ((0x8, "hw_slowdown"), (0x40, "hw_thermal_slowdown"), (0x20, "sw_thermal_slowdown"), (0x4, "sw_power_cap"))

    def __init__(self, index, uuid=None):
        self.index = index
        self.samples = []          # (sm_mhz, max_mhz, set of reasons)
        self._stop = threading.Event()
        self._t = None
        self._h = None
        try:
            import pynvml
            pynvml.nvmlInit()
            h = None
            if uuid:
                try:
                    h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + str(uuid)) if not str(uuid).startswith("GPU-") else str(uuid))
                except Exception:
                    h = None
            self._h = h or pynvml.nvmlDeviceGetHandleByIndex(index)
            self._nvml = pynvml
            self._max = float(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self._h = None

    def _sample_nvml(self):
        n = self._nvml
        sm = float(n.nvmlDeviceGetClockInfo(self._h, n.NVML_CLOCK_SM))
        try:
            mask = int(n.nvmlDeviceGetCurrentClocksEventReasons(self._h))
        except Exception:
            mask = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self._h))
        self.samples.append((sm, self._max, {name for bit, name in self.REASONS if mask & bit}))

    def _sample_smi(self):
        out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.QUERY, "--format=csv,noheader,nounits"],
                             stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, timeout=5).stdout.strip()
        if out:
            s = [x.strip() for x in out.split(",")]
            self.samples.append((float(s[0]), float(s[1]), {name for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[3:7])
                                                             if v.lower().startswith("active")}))

    def _run(self):
        while not self._stop.is_set():
            try:
                self._sample_nvml() if self._h is not None else self._sample_smi()
            except Exception:
                pass
            self._stop.wait(0.01 if self._h is not None else 0.5)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        sm = [s[0] for s in self.samples]
        mx = max([s[1] for s in self.samples] + [0])
        reasons = set().union(*[s[2] for s in self.samples]) if self.samples else set()
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm),
                "source": "nvml" if self._h is not None else "nvidia-smi"}


class NoSampler:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        pass

    def summary(self):
        return None


def settings_from_totals(tags, n_samples, directional, mappable, genome, bandwidth=50):
    """What src/regions.cpp:152-232 derives with default flags (no controls, no coefficients)."""
    total = int(sum(tags))
    background = float(total) / float(mappable or genome)
    if directional:
        background /= 2
    return dict(bandwidth=bandwidth, region_threshold=25.0, kurtosis_threshold=50.0, corr_threshold=(-1.0 if directional else 0.3),
                hit_threshold=10.0 * n_samples, directional=directional, control=np.zeros(n_samples, np.uint8), coeffs=[],
                background=background)


def cpu_sample(data: TagData, contigs):
    """Bounded CPU sample: a few contigs of the same workload (all samples, both strands)."""
    return data.subset_contigs(list(contigs))


def run_cpu_reference(sample: TagData, st: dict, steps: int, workdir: str):
    """Time the reference's hot path (or the C port) on the bounded sample; returns (seconds per step, kind, rows)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import refutil
    if refutil.have_ref():
        best, rows = None, None
        for _ in range(max(1, steps)):
            r = refutil.run_ref_replay_regions(sample, st, workdir)
            best = r["seconds"] if best is None else min(best, r["seconds"])
            rows = r
        return best, "reference", rows
    best, rows = None, None
    for _ in range(max(1, steps)):
        t0 = time.perf_counter()
        r = refutil.oracle_regions(sample, st)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
        rows = r
    return best, "port", rows


def pin_check(view, workload, args, directional):
    """Compare the whole table with the fingerprints of the reference's own full-size run (tools/pin_fullsize.py)."""
    if args.samples or args.tags or args.bandwidth != 50 or args.seed != SEEDS.get(workload):
        return None
    path = os.path.join(ROOT, "tests", "golden", "fullsize_%s.json" % workload)
    if not os.path.exists(path):
        return None
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    from pin_fullsize import column_hashes
    pin = json.load(open(path))["regions"]
    c = view.columns()
    keep = np.concatenate([np.nonzero(c["dropped"] == 0)[0], np.nonzero(c["dropped"] != 0)[0]])
    cols = {k: np.ascontiguousarray(np.asarray(c[k])[keep]) for k in ("forward_label", "contig", "left", "len", "peak", "peak_score", "kurtosis", "corr", "counts")}
    h = column_hashes(cols)
    bad = [k for k, v in pin["columns"].items() if h[k] != v and not (k == "corr" and directional)]
    ok = (not bad and view.n_rows == pin["n_rows"] + pin["n_dropped"] and view.n_rejected == pin["n_rejected"]
          and [int(x) for x in c["tags_in_regions"]] == pin["tags_in_regions"])
    return {"identical_to_reference": bool(ok), "columns_checked": sorted(k for k in pin["columns"] if not (k == "corr" and directional)),
            "rows": pin["n_rows"] + pin["n_dropped"], "differing": bad,
            "what": "every column of the whole table (SHA-256 of the raw values, doubles as bit patterns) against the compiled reference's own run over this whole workload (tests/golden/fullsize_%s.json)" % workload,
            "reference_seconds_whole_workload_cpu_container": pin.get("reference_seconds")}


SEEDS = {"c3": 20261018, "c2": 20261019, "c1": 20261020, "c4": 20261021, "tiny": 7}


FORM_TEXT = {
    "vbyte": "vbyte (variable-length byte code, ~1.5 B per entry: marker bit + 6 + 7k-bit position delta, optional count byte, exception list for counts > 127; "
             "upb_load_tags_async — what bin/regions loads)",
    "packed2": "packed2 (u16 per entry: 12-bit position delta + 4-bit count, exception list; upb_load_tags_async)",
    "plain": "plain (u32 pos + u32 cnt: upb_load_tags_async)",
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--samples", type=int, default=0)
    ap.add_argument("--tags", type=int, default=0)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-sweep", action="store_true", help="skip roofline.sweep (bandwidth 10 ... 200)")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--bandwidth", type=int, default=50, help="-b (config 5 sweeps 10..200)")
    ap.add_argument("--host-form", default=os.environ.get("UPB_HOST_FORM", "vbyte"), choices=["vbyte", "packed2", "plain"],
                    help="host form of the tag arrays in the headline e2e leg (the other packed form is reported as e2e_alt2, the plain one as e2e_alt)")
    args = ap.parse_args()
    args.seed = args.seed or SEEDS[args.workload]

    rank, world, local_rank = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    desc, n_samples, tags_per_sample, directional, sizes, names, mappable = WORKLOADS[args.workload]
    n_samples = args.samples or n_samples
    tags_per_sample = args.tags or tags_per_sample
    genome = int(sum(sizes))
    metric = "genome bp/s for pooled KDE+region calling"
    config = {"workload": desc if not (args.samples or args.tags) else "%s [overridden: %d samples x %d tags]" % (desc, n_samples, tags_per_sample),
              "samples": n_samples, "tags_per_sample": tags_per_sample, "genome_bp": genome, "bandwidth": args.bandwidth,
              "l2": "inputs (GBs of tag entries) are far larger than the 126 MB L2; no explicit flush",
              "generator": "unipeak_b200.synth_det (integer hash; bit-identical on CPU and GPU), seed %d" % args.seed,
              "parallelism": "contig-sharded x%d" % world if world > 1 else "single GPU"}

    import torch
    from unipeak_b200.synth_det import make_synthetic_det
    if args.impl == "reference":
        # the reference's own CPU implementation of the path, rank 0 only
        if rank != 0:
            return
        cs = CPU_SAMPLE_CONTIGS[args.workload]
        dev = "cuda:%d" % local_rank if torch.cuda.is_available() else "cpu"
        # only the sample's contigs are generated (they are exactly those of the whole workload); the thresholds need the whole
        # workload's tag totals, which the generator gives without the arrays: every tag lands somewhere unless it falls off an edge
        full = make_synthetic_det(sizes, n_samples, tags_per_sample, directional=directional, seed=args.seed, device=dev, bandwidth=max(50, args.bandwidth))
        st = settings_from_totals([full.tags(s) for s in range(n_samples)], n_samples, directional, mappable, genome, args.bandwidth)
        sample = cpu_sample(full, cs)
        del full
        with tempfile.TemporaryDirectory() as wd:
            if args.warmup:
                run_cpu_reference(sample, st, 1, wd)
            secs, kind, _rows = run_cpu_reference(sample, st, args.steps, wd)
        bp = int(sum(sizes[c] for c in cs))
        val = bp / secs
        line = {"impl": "reference", "metric": metric, "value": val, "unit": "bp/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": secs * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": config,
                "cpu_baseline": {"value": val, "unit": "bp/s", "cores": 1, "kind": kind,
                                 "sample": "contigs %s (%d bp = %.1f %% of the genome, %d tag entries) of the workload, all samples, workload-wide thresholds; best of %d runs; "
                                           "bp/s of this slice stands for the whole workload (EXTRAPOLATED: tag density is uniform across contigs in the synthetic data)"
                                           % ("+".join(names[c] for c in cs), bp, 100.0 * bp / genome, int(sample.pos.shape[0]), max(1, args.steps))},
                "e2e": {"value": val, "unit": "bp/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    # ------------------------------------------------------------------------------------------------
    # our arm
    # ------------------------------------------------------------------------------------------------
    import torch.distributed as dist
    import unipeak_b200
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU path)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)

    owner = shard_contigs(sizes, world)
    mine = [c for c in range(len(sizes)) if owner[c] == rank]
    data = make_synthetic_det(sizes, n_samples, tags_per_sample, directional=directional, seed=args.seed, contigs=mine, device=str(dev), bandwidth=max(50, args.bandwidth))

    # pinned host copies (the e2e legs copy from these every step)
    h_off = torch.from_numpy(data.off.view(np.int64)).pin_memory()
    h_pos = torch.from_numpy(data.pos.view(np.int32)).pin_memory()
    h_cnt = torch.from_numpy(data.cnt.view(np.int32)).pin_memory()
    E = int(data.pos.shape[0])
    C = len(sizes)

    # genome-wide tag totals -> background (src/regions.cpp:190-210): the one NCCL all-reduce the path needs, at set-up
    tags = torch.tensor([data.tags(s) for s in range(n_samples)], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(tags)
    st = settings_from_totals(tags.cpu().tolist(), n_samples, directional, mappable, genome, args.bandwidth)

    ctx = unipeak_b200.Context(local_rank)
    ctx.set_contigs(sizes)

    # the same data in the 2-byte host form; the encoder's time is reported (pack_ms), outside every timed region — a front end
    # encodes while it parses (bin/regions does: unipeak_b200/host/upb_host.cpp)
    h_w2 = torch.empty(max(E, 1), dtype=torch.int16).pin_memory()
    t0 = time.perf_counter()
    _, exc2 = unipeak_b200.pack_tags2(n_samples, C, data.off, data.pos, data.cnt, out=h_w2.numpy().view(np.uint16))
    pack_ms = (time.perf_counter() - t0) * 1e3
    h_exc2 = torch.from_numpy(exc2.view(np.uint8).copy()).pin_memory()
    n_exc2 = int(exc2.shape[0])
    # ... and in the variable-length byte form (upb_vbyte.hpp: ~1.5 bytes per entry)
    h_vb = torch.empty(max(6 * E, 1) if E < (1 << 24) else max(3 * E, 1), dtype=torch.uint8).pin_memory()
    t0 = time.perf_counter()
    try:
        vb, vb_off, vb_tiles, exc_v = unipeak_b200.pack_tags_vbyte(n_samples, C, data.off, data.pos, data.cnt, out=h_vb.numpy())
    except ValueError:                                                # a track so sparse that it needs more than 3 bytes per entry: the worst case is 6
        h_vb = torch.empty(6 * E, dtype=torch.uint8).pin_memory()
        t0 = time.perf_counter()
        vb, vb_off, vb_tiles, exc_v = unipeak_b200.pack_tags_vbyte(n_samples, C, data.off, data.pos, data.cnt, out=h_vb.numpy())
    pack_vbyte_ms = (time.perf_counter() - t0) * 1e3
    n_vb = int(vb.shape[0])
    h_excv = torch.from_numpy(exc_v.view(np.uint8).copy()).pin_memory()
    n_excv = int(exc_v.shape[0])
    h2d_bytes = {"packed2": 2 * E + 16 * n_exc2 + 8 * int(data.off.shape[0]), "plain": 8 * E + 8 * int(data.off.shape[0]),
                 "vbyte": n_vb + 16 * n_excv + 8 * int(vb_tiles.shape[0]) + 16 * int(data.off.shape[0])}
    if world > 1:                                                     # whole job: the ranks' loads together
        hb = torch.tensor([h2d_bytes["packed2"], h2d_bytes["plain"], h2d_bytes["vbyte"]], dtype=torch.int64, device=dev)
        dist.all_reduce(hb)
        h2d_bytes = {"packed2": int(hb[0].item()), "plain": int(hb[1].item()), "vbyte": int(hb[2].item())}

    def load_plain():
        ctx.load_async_ptr(n_samples, h_off.data_ptr(), pos_ptr=h_pos.data_ptr(), cnt_ptr=h_cnt.data_ptr(), directional_hint=directional)

    def load_packed2():
        ctx.load_async_ptr(n_samples, h_off.data_ptr(), words_ptr=h_w2.data_ptr(), n_exc=n_exc2, exc_ptr=h_exc2.data_ptr(), directional_hint=directional)

    def load_vbyte():
        ctx.load_async_ptr(n_samples, h_off.data_ptr(), vbytes_ptr=h_vb.data_ptr(), vbyte_off_ptr=vb_off.ctypes.data, vtile_off_ptr=vb_tiles.ctypes.data,
                           n_vtiles=int(vb_tiles.shape[0]) - 1, n_exc=n_excv, exc_ptr=h_excv.data_ptr(), directional_hint=directional)

    loaders = {"vbyte": load_vbyte, "packed2": load_packed2, "plain": load_plain}
    loaders[args.host_form]()
    ctx.load_wait()
    first = ctx.call_regions_view(**dict(st, reference_order=True))       # sizes the context's arrays; its row count sizes the shared table
    shared = None
    if world > 1:
        rows = torch.tensor([first.n_rows], dtype=torch.int64, device=dev)
        dist.all_reduce(rows)
        ctx.release_view()
        shared = SharedRegionTable(ctx, n_samples, C, rank=rank, world=world, row_capacity=int(rows.item()) * 3 // 2 + 4096, dist=dist)
        config["parallelism"] += ", region-list gather: every rank's GPU writes its rows into one shared mapped host table after a device-side exchange of the per-unit row counts; no collective inside a step"

    def step():
        if world == 1:
            return ctx.call_regions_view(**dict(st, reference_order=True))
        return shared(**dict(st, reference_order=True))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # device timing: CUDA events on the stream the library launches on (torch's current stream never sees those kernels).
    # Every call returns only when its table is in host memory, so the events bracket kernels, copies and the host gaps between
    # steps; the wall clock around the same region is reported next to it as a cross-check.
    lib_stream = torch.cuda.ExternalStream(ctx.stream_ptr(), device=dev)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    for _ in range(args.warmup):
        last = step()
    kde_ms, launches = [], 0
    call_ms, host_ms, dev_ms, step_wall = [], [], [], []
    try:
        gpu_uuid = torch.cuda.get_device_properties(dev).uuid
    except Exception:
        gpu_uuid = None
    # the sampler is built (NVML initialised: ~0.1 s) BEFORE the barrier: built after it, rank 0 entered the timed region that much later than
    # the other ranks, whose first step then waited for its row counts on the device — 20 steps read 9.3 ms instead of 2.1 (r02, N = 2)
    sampler = ClockSampler(local_rank, gpu_uuid) if rank == 0 else NoSampler()      # rank 0's GPU stands for the box
    barrier()
    with sampler as clocks:
        t0 = time.perf_counter()
        ev0.record(lib_stream)
        for _ in range(args.steps):
            ts = time.perf_counter()
            last = step()
            step_wall.append((time.perf_counter() - ts) * 1e3)
            s = ctx.stats()
            kde_ms.append(s["ms_kde"]); launches += s["launches"]; call_ms.append(s["ms_call"]); host_ms.append(s["host_ms"]); dev_ms.append(s["ms_total"])
        ev1.record(lib_stream)
        barrier()
        dt_wall = time.perf_counter() - t0
    stats = ctx.stats()
    ev1.synchronize()
    dt = ev0.elapsed_time(ev1) * 1e-3
    tmax = torch.tensor([dt, dt_wall], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    dt, dt_wall = float(tmax[0].item()), float(tmax[1].item())
    ms_per_step = dt / args.steps * 1e3
    wall_ms_per_step = dt_wall / args.steps * 1e3
    n_rows = int(last.n_rows) if rank == 0 else 0
    pin = pin_check(last, args.workload, args, directional) if rank == 0 else None

    def timed(fn, reps):
        """reps x fn() between two events on the library's stream; (device ms, wall ms) per repetition, max over ranks."""
        barrier()
        t0 = time.perf_counter()
        ev0.record(lib_stream)
        for _ in range(reps):
            res = fn()
        ev1.record(lib_stream)
        barrier()
        wall = time.perf_counter() - t0
        ev1.synchronize()
        t = torch.tensor([ev0.elapsed_time(ev1) * 1e-3, wall], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0].item()) / reps * 1e3, float(t[1].item()) / reps * 1e3, res

    # e2e: pinned host arrays -> device every step; the call starts on the load groups that have arrived
    e2e_steps = max(1, min(args.steps, 5))

    load_phases = []                                                 # (copied_ms, indexed_ms) of this rank's loads in the e2e legs

    def e2e_leg(loader):
        def one():
            loader()
            res = step()
            load_phases.append(ctx.last_load_ms())
            return res
        one()                                                        # untimed: buffers of this path reach their size
        del load_phases[:]
        return timed(one, e2e_steps)

    alt_ms, alt_wall_ms, _ = e2e_leg(load_plain)
    other_form = "packed2" if args.host_form == "vbyte" else "vbyte"
    alt2_ms, alt2_wall_ms, alt2_last = e2e_leg(loaders[other_form])
    alt2_rows = int(alt2_last.n_rows) if rank == 0 else 0
    e2e_ms, e2e_wall_ms, last = e2e_leg(loaders[args.host_form])
    # per-rank phases of the headline leg: bytes this rank moves, when its last byte was on the device, when its index was complete (library events)
    ph = np.asarray(load_phases, dtype=np.float64).reshape(-1, 2) if load_phases else np.full((1, 2), -1.0)
    e2e_phase = torch.tensor([float(ph[:, 0].mean()), float(ph[:, 1].mean()), float(E)], dtype=torch.float64, device=dev)
    e2e_phase_all = [torch.zeros_like(e2e_phase) for _ in range(world)]
    if world > 1:
        dist.all_gather(e2e_phase_all, e2e_phase)
    else:
        e2e_phase_all = [e2e_phase]
    e2e_per_rank = {"columns": ["h2d_done_ms", "index_done_ms", "tag_entries"], "rows": [[round(float(x), 3) for x in t.tolist()] for t in e2e_phase_all]}
    e2e_rows = int(last.n_rows) if rank == 0 else 0
    E_all = int(sum(float(t[2].item()) for t in e2e_phase_all))
    d2h = n_rows * (41 + 4 * n_samples)

    # t_hot_cold: from sample-major arrays resident in HBM (SURVEY 8d's "sorted arrays resident") to the table on the host: the
    # index build (tile-major counting sort + both zone maps — the device-side replacement of the reference's k-way merge, row a3)
    # is inside
    def cold():
        ctx.reindex()
        return step()
    cold()
    cold_ms, cold_wall_ms, _ = timed(cold, e2e_steps)

    # roofline of the fused KDE kernel (rank-local numbers; every rank runs the same kernel on its shard)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_gbs, peak_src = (peaks.get("hbm_gbs"), "measured") if peaks.get("hbm_gbs") else (6650.0, "fallback")
    traffic = None
    tj = {}
    try:    # dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture of this kernel on this workload (profiles/)
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        key = "%s_bw%d" % (args.workload, args.bandwidth)
        if world == 1 and not (args.samples or args.tags) and key in tj:
            traffic = tj[key]["span_kernel_dram_bytes"]
    except Exception:
        pass
    kde_avg_ms = float(np.mean(kde_ms)) if kde_ms else float("nan")
    n_cand = stats["n_candidates"]
    alg_bytes = 8 * E + (40 + 4 * n_samples) * n_cand                # every entry read once + run record, kurtosis, per-sample hits per candidate
    achieved = alg_bytes / (kde_avg_ms * 1e-3) / 1e9 if kde_avg_ms > 0 else 0.0
    roofline = {"bound": "hbm", "kernel": "span_kernel<directional>" if directional else "span_kernel<nondirectional>", "achieved": achieved, "peak": peak_gbs,
                "peak_source": peak_src, "unit": "GB/s", "frac": achieved / peak_gbs, "traffic": traffic,
                "traffic_source": tj.get("_source") if traffic else None,
                "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": kde_avg_ms, "kernel_share_of_step": kde_avg_ms / ms_per_step}

    # every stage of the call the same way (CUDA events between the stages, upb_set_profiling; separate steps, not the timed ones)
    kernels, sweep = None, None
    if True:                                                            # (N > 1: every rank steps, rank 0's stage times are reported, without byte counts)
        ctx.set_profiling(True)
        acc = {}
        for _ in range(5):
            step()
            for name, ms in ctx.stage_times():
                acc.setdefault(name, []).append(ms)
        ctx.set_profiling(False)
        R1 = max(n_rows, 1)
        # algorithmic bytes per stage: zone maps for the two prefilters, the entries for the span kernel, the entries
        # of the accepted candidates' mini-tiles for the pooling, pooled cells for the moments, rows for the emit
        stage_bytes = {
            "warm_span_kernel+hot_span_kernel": 8 * 2 * (genome >> 9) + 4 * stats["warm_spans"] * 160,
            "span_kernel": alg_bytes,
            "order+stitch+hit filter": 32 * n_cand * 4,
            "region_pool_kernel": None,
            "region_moment_kernel": None,
            "place+emit rows (to host)+totals": R1 * (41 + 4 * n_samples),
        }
        if world > 1:
            stage_bytes = {}
        kernels = []
        for name, v in acc.items():
            ms = float(np.mean(v))
            b = stage_bytes.get(name)
            kernels.append({"name": name, "ms": ms, "algorithmic_bytes": b, "frac": (b / (ms * 1e-3) / 1e9 / peak_gbs) if (b and ms > 0) else None})
        if world == 1 and args.workload in ("c3", "tiny") and not (args.samples or args.tags) and not args.no_sweep:
            sweep = []
            for bw in (10, 20, 50, 100, 150, 200):
                stb = dict(st, bandwidth=bw, reference_order=True)
                ctx.call_regions_view(**stb)
                ms_l = []
                for _ in range(5):
                    ctx.call_regions_view(**stb)
                    s5 = ctx.stats()
                    ms_l.append(s5["ms_kde"])
                ab = 8 * E + (40 + 4 * n_samples) * s5["n_candidates"]
                m = float(np.mean(ms_l))
                sweep.append({"bandwidth": bw, "kernel_ms": m, "achieved": ab / (m * 1e-3) / 1e9, "frac": ab / (m * 1e-3) / 1e9 / peak_gbs, "candidates": s5["n_candidates"], "call_ms": s5["ms_call"]})
            ctx.call_regions_view(**dict(st, reference_order=True))
        roofline["sweep"] = sweep

    per_rank = None
    if world > 1:
        hm = np.asarray(host_ms, dtype=np.float64).reshape(-1, 5).mean(axis=0) if host_ms else np.zeros(5)
        minev = torch.tensor([float(np.mean(step_wall)) if step_wall else 0.0, float(np.mean(call_ms)) if call_ms else 0.0, float(np.mean(dev_ms)) if dev_ms else 0.0,
                              float(np.mean(kde_ms)) if kde_ms else 0.0] + [float(x) for x in hm] + [float(E)] +
                             [float(np.median(np.asarray(host_ms)[:, 1])) if host_ms else 0.0, float(np.max(np.asarray(host_ms)[:, 1])) if host_ms else 0.0], dtype=torch.float64, device=dev)
        allr = [torch.zeros_like(minev) for _ in range(world)]
        dist.all_gather(allr, minev)
        per_rank = {"columns": ["python_step_ms", "call_ms", "device_ms", "span_kernel_ms", "host_prepare_ms", "host_enqueue_ms", "host_wait_ms", "host_post_ms", "host_assemble_ms", "tag_entries", "host_enqueue_median_ms", "host_enqueue_max_ms"],
                    "note": "means over the timed steps; python_step_ms: wall clock around the binding's call; call_ms: inside the library; device_ms: first event to last event on the call's stream; host_*: where the library's host thread was",
                    "rows": [[round(float(x), 4) for x in t.cpu().tolist()] for t in allr]}

    if rank == 0:
        cold_bytes = alg_bytes + 16 * E + 12 * 2 * (genome >> 9) + 4 * 2 * (genome >> 5)   # + sort: 8E read + 8E written, + both zone maps written
        line = {"metric": metric, "value": genome / (ms_per_step * 1e-3), "unit": "bp/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_per_step, "wall_ms_per_step": wall_ms_per_step, "timing": "CUDA events on the library's stream around the K steps (max over ranks); wall clock alongside",
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": config, "clocks": clocks.summary(),
                "e2e": {"value": genome / (e2e_ms * 1e-3), "unit": "bp/s", "ms_per_step": e2e_ms, "wall_ms_per_step": e2e_wall_ms, "h2d_bytes_per_step": h2d_bytes[args.host_form], "d2h_bytes_per_step": d2h,
                        "host_format": FORM_TEXT[args.host_form],
                        "bytes_per_entry": h2d_bytes[args.host_form] / max(E_all, 1),
                        "pack_ms": {"vbyte": pack_vbyte_ms, "packed2": pack_ms, "plain": 0.0}[args.host_form],
                        "pack_note": "host-side encoder (upb_pack_tags_vbyte / upb_pack_tags2, one thread per (sample, strand), from the plain arrays), outside the timed region: a front end encodes while it parses (bin/regions: TagLoader)",
                        "per_rank_ms": e2e_per_rank, "per_rank_note": "from the first queued copy of a rank's load: last byte on the device, index complete (library CUDA events)",
                        "rows_equal_resident_leg": bool(e2e_rows == n_rows)},
                "e2e_alt": {"host_format": "plain (u32 pos + u32 cnt: upb_load_tags_async)", "value": genome / (alt_ms * 1e-3), "unit": "bp/s", "ms_per_step": alt_ms, "wall_ms_per_step": alt_wall_ms, "h2d_bytes_per_step": h2d_bytes["plain"]},
                "e2e_alt2": {"host_format": FORM_TEXT[other_form], "value": genome / (alt2_ms * 1e-3), "unit": "bp/s", "ms_per_step": alt2_ms, "wall_ms_per_step": alt2_wall_ms,
                             "h2d_bytes_per_step": h2d_bytes[other_form], "pack_ms": {"vbyte": pack_vbyte_ms, "packed2": pack_ms}[other_form], "rows_equal_resident_leg": bool(alt2_rows == n_rows)},
                "t_hot_cold": {"value": genome / (cold_ms * 1e-3), "unit": "bp/s", "ms_per_step": cold_ms, "wall_ms_per_step": cold_wall_ms,
                               "what": "sample-major pos/cnt resident in HBM -> upb_reindex (tile-major counting sort + zone maps) -> upb_call_regions -> table on the host",
                               "algorithmic_bytes": cold_bytes, "roofline_frac": cold_bytes / (cold_ms * 1e-3) / 1e9 / peak_gbs / world},
                "gpu_launches": launches, "roofline": roofline, "kernels": kernels,
                "hot_path_roofline_frac": alg_bytes / (ms_per_step * 1e-3) / 1e9 / peak_gbs,
                "result": {"regions": n_rows, "candidates": n_cand, "tag_entries_rank0": E, "spans": stats["kde_tiles"], "warm_spans": stats["warm_spans"], "pieces": stats["kde_hot_tiles"],
                           "hot_blocks": stats["hot_blocks"], "exact_evals": stats["exact_evals"], "stitched": stats["n_stitched"],
                           "device_ms_per_step_rank0": stats["ms_total"], "call_ms_rank0": stats["ms_call"], "launches_per_call": stats["launches"]}}
        if pin is not None:
            line["parity_full_size"] = pin
        if world == 1 and not args.no_cpu:
            cs = CPU_SAMPLE_CONTIGS[args.workload]
            sample = cpu_sample(data, cs)
            with tempfile.TemporaryDirectory() as wd:
                secs, kind, ref = run_cpu_reference(sample, st, 1, wd)
            bp = int(sum(sizes[c] for c in cs))
            # parity spot-check on the sample: every region the CPU side finds on these contigs must equal ours, doubles bit for bit
            ctx.release_view()
            full = ctx.call_regions(**dict(st, reference_order=True, rows=True))
            ours = [r for r in full["rows"] + full["dropped"] if r["contig"] in cs]
            theirs = ref["rows"] + ref["dropped"]
            bits = lambda x: np.float64(x).view(np.uint64).item() if x == x else -1
            key_o = lambda r: (cs.index(r["contig"]), r["left"], r["len"], r["peak"], tuple(r["counts"]), bits(r["kurtosis"]), bits(r["peak_score"]))
            key_t = lambda r: (r["contig"], r["left"], r["len"], r["peak"], tuple(r["counts"]), bits(r["kurtosis"]), bits(r["peak_score"]))
            same = sorted(key_o(r) for r in ours) == sorted(key_t(r) for r in theirs)
            line["cpu_baseline"] = {"value": bp / secs, "unit": "bp/s", "cores": 1, "kind": kind, "seconds": secs,
                                    "sample": "contigs %s (%d bp = %.1f %% of the genome, %d tag entries) of the workload, all samples, workload-wide thresholds; bp/s of this slice stands for the whole workload (EXTRAPOLATED)"
                                              % ("+".join(names[c] for c in cs), bp, 100.0 * bp / genome, int(sample.pos.shape[0])),
                                    "host_cores_available": os.cpu_count(), "rows_match_gpu": bool(same), "rows": len(theirs),
                                    "rows_key": "contig, left, len, peak, counts[], kurtosis bits, peak_score bits"}
        if per_rank is not None:
            line["result"]["per_rank_ms"] = per_rank
        print(json.dumps(line))
    if shared is not None:
        barrier()
        shared.close()
    try:
        ctx.close()
    except Exception:                                                 # noqa: BLE001
        pass
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
