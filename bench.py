#!/usr/bin/env python
"""bench.py — genome bp/s of pooled KDE + region calling + count matrix (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c3|c2|c1|tiny]

A step is one pass of the hot path over the resident synthetic input: upb_call_regions() — pooled
Epanechnikov KDE, thresholding, region extraction, hit/kurtosis(/correlation) filters, regions x samples
count matrix — ending with the region table in host memory in the reference's emission order.
N > 1 (torchrun, one process per GPU): contigs are sharded over the ranks (strong scaling: the genome is
fixed); NCCL carries only the per-sample tag totals (background), the tags-in-regions totals and the
gather of the region rows to rank 0.

`value`  : genome bp / step time, inputs already resident in HBM (CUDA events on the library's stream around the K steps,
           bracketed by barrier + synchronize, max over ranks; the wall clock of the same region is reported beside it).
`e2e`    : the same through the C-ABI from pinned HOST buffers every step: H2D copy + decode + counting sort + zone maps +
           call + table D2H.  Default host form: the compact one (upb_load_tags_packed: u16 position deltas + u8 counts +
           exception list, 3 B per entry); `e2e_alt` is the plain form (upb_load_tags: u32 + u32, 8 B per entry); `e2e_packed2`
           (N = 1 only, measured after everything else, guarded) is the 2-byte form (upb_load_tags_packed2), a side leg until its
           device path has been through the GPU tests.
`roofline`: the fused span kernel, algorithmic bytes (8 B per tag entry read once + 40+4n B per candidate written)
           over its CUDA-event duration, against MEASURED_PEAKS.json hbm_gbs.
`cpu_baseline` / --impl reference: the reference's own Kernel/ProfileBuffer/Region objects
           (oracle/_ref/ref_replay, kind "reference") — or the C restatement (kind "port") when the
           compiled reference is absent — single-threaded like the reference, on a bounded sample
           (one contig of the same workload, same thresholds).
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


from unipeak_b200.multi import DeviceGather, HostGather, shard_contigs  # noqa: E402


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
    ap.add_argument("--seed", type=int, default=20261018)
    ap.add_argument("--bandwidth", type=int, default=50, help="-b (config 5 sweeps 10..200)")
    ap.add_argument("--host-format", default="packed", choices=["packed", "plain"],
                    help="host buffers of the e2e leg: the compact form (u16 position deltas + u8 counts + exception list, upb_load_tags_packed) or plain u32 pos / u32 cnt (upb_load_tags); the other one is reported as e2e_alt")
    ap.add_argument("--gather", default="host", choices=["host", "nccl"],
                    help="N > 1 region-list gather: every rank writes its rows into one shared, mapped host table (host), or one NCCL reduce of a device slab + D2H on rank 0 (nccl)")
    args = ap.parse_args()

    rank, world, local_rank = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    desc, n_samples, tags_per_sample, directional, sizes, names, mappable = WORKLOADS[args.workload]
    n_samples = args.samples or n_samples
    tags_per_sample = args.tags or tags_per_sample
    genome = int(sum(sizes))
    metric = "genome bp/s for pooled KDE+region calling"
    config = {"workload": desc if not (args.samples or args.tags) else "%s [overridden: %d samples x %d tags]" % (desc, n_samples, tags_per_sample),
              "samples": n_samples, "tags_per_sample": tags_per_sample, "genome_bp": genome, "bandwidth": args.bandwidth,
              "l2": "inputs (GBs of tag entries) are far larger than the 126 MB L2; no explicit flush",
              "parallelism": "contig-sharded x%d" % world if world > 1 else "single GPU"}

    import torch
    if args.impl == "reference":
        # the reference's own CPU implementation of the path, rank 0 only
        if rank != 0:
            return
        cs = CPU_SAMPLE_CONTIGS[args.workload]
        if torch.cuda.is_available():
            from unipeak_b200.synth_gpu import make_synthetic_gpu
            data = make_synthetic_gpu(sizes, n_samples, tags_per_sample, directional=directional, seed=args.seed, device="cuda:%d" % local_rank, bandwidth=max(50, args.bandwidth))
        else:
            from unipeak_b200.synth import make_synthetic
            data = make_synthetic(sizes, n_samples, tags_per_sample, directional=directional, seed=args.seed, bandwidth=max(50, args.bandwidth))
        # thresholds come from the whole workload's tag totals, exactly as in our arm
        st = settings_from_totals([data.tags(s) for s in range(n_samples)], n_samples, directional, mappable, genome, args.bandwidth)
        sample = cpu_sample(data, cs)
        del data
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
                                 "sample": "contigs %s (%d bp, %d tag entries) of the workload, all samples, workload-wide thresholds; best of %d runs"
                                           % ("+".join(names[c] for c in cs), bp, int(sample.pos.shape[0]), max(1, args.steps))},
                "e2e": {"value": val, "unit": "bp/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    # ------------------------------------------------------------------------------------------------
    # our arm
    # ------------------------------------------------------------------------------------------------
    import torch.distributed as dist
    import unipeak_b200
    from unipeak_b200.synth_gpu import make_synthetic_gpu
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU path)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)

    owner = shard_contigs(sizes, world)
    mine = [c for c in range(len(sizes)) if owner[c] == rank]
    data = make_synthetic_gpu(sizes, n_samples, tags_per_sample, directional=directional, seed=args.seed, contigs=mine, device=str(dev), bandwidth=max(50, args.bandwidth))

    # pinned host copies (the e2e leg copies from these every step)
    h_off = torch.from_numpy(data.off.view(np.int64)).pin_memory()
    h_pos = torch.from_numpy(data.pos.view(np.int32)).pin_memory()
    h_cnt = torch.from_numpy(data.cnt.view(np.int32)).pin_memory()
    E = int(data.pos.shape[0])
    import unipeak_b200

    # genome-wide tag totals -> background (src/regions.cpp:190-210): the one NCCL all-reduce the path needs up front
    tags = torch.tensor([data.tags(s) for s in range(n_samples)], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(tags)
    st = settings_from_totals(tags.cpu().tolist(), n_samples, directional, mappable, genome, args.bandwidth)

    ctx = unipeak_b200.Context(local_rank)
    ctx.set_contigs(sizes)

    # the same data in the compact host form (encoded once, outside every timed region — a front end produces it while parsing)
    h_dpos = torch.empty(max(E, 1), dtype=torch.int16).pin_memory()
    h_cnt8 = torch.empty(max(E, 1), dtype=torch.uint8).pin_memory()
    _, _, exc = unipeak_b200.pack_tags(n_samples, len(sizes), data.off, data.pos, data.cnt, out=(h_dpos.numpy().view(np.uint16), h_cnt8.numpy()))
    h_exc = torch.from_numpy(exc.view(np.uint8).copy()).pin_memory()
    n_exc = int(exc.shape[0])

    def load_plain():
        ctx.load_tags_ptr(n_samples, h_off.data_ptr(), h_pos.data_ptr(), h_cnt.data_ptr())

    def load_packed():
        ctx.load_tags_packed_ptr(n_samples, h_off.data_ptr(), h_dpos.data_ptr(), h_cnt8.data_ptr(), n_exc, h_exc.data_ptr())

    load, load_alt = (load_packed, load_plain) if args.host_format == "packed" else (load_plain, load_packed)
    h2d_bytes = {"packed": 3 * E + 16 * n_exc + 8 * int(data.off.shape[0]), "plain": 8 * E + 8 * int(data.off.shape[0])}

    # the 2-byte host form (12-bit position delta + 4-bit count per entry: upb_load_tags_packed2) — reported beside the others as
    # `e2e_packed2`, never as the headline, until its device side has been through the GPU tests (include/unipeak_b200.h)
    h_w2 = torch.empty(max(E, 1), dtype=torch.int16).pin_memory()
    _, exc2 = unipeak_b200.pack_tags2(n_samples, len(sizes), data.off, data.pos, data.cnt, out=h_w2.numpy().view(np.uint16))
    h_exc2 = torch.from_numpy(exc2.view(np.uint8).copy()).pin_memory()
    n_exc2 = int(exc2.shape[0])
    h2d_bytes["packed2"] = 2 * E + 16 * n_exc2 + 8 * int(data.off.shape[0])

    def load_packed2():
        ctx.load_tags_packed2_ptr(n_samples, h_off.data_ptr(), h_w2.data_ptr(), n_exc2, h_exc2.data_ptr())

    C = len(sizes)
    gather = None
    if world > 1:
        gather = (HostGather if args.gather == "host" else DeviceGather)(ctx, n_samples, C, dist=dist, rank=rank, world=world, device=dev)
        config["parallelism"] += ", region-list gather: " + ("rows written by every rank into one shared mapped host table, two small NCCL all-reduces" if args.gather == "host" else "one NCCL reduce of a device slab + D2H on rank 0")

    def step():
        res = ctx.call_regions_view(**dict(st, reference_order=(1 if world == 1 else 2)))   # ranks of a multi-GPU job keep their rows on the device
        if world == 1:
            return res
        # NCCL: one small all-reduce (rows / candidates per unit, tags in regions) + one reduce of the table-wide slab to rank 0
        return gather(res)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # device timing: CUDA events on the stream the library launches on (torch's current stream never sees those kernels).
    # Every call returns only when its table is in host memory, so the events bracket kernels, copies and the host gaps between
    # steps; the wall clock around the same region is reported next to it as a cross-check.
    lib_stream = torch.cuda.ExternalStream(ctx.stream_ptr(), device=dev)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    load()
    for _ in range(args.warmup):
        last = step()
    kde_ms, launches = [], 0
    barrier()
    if gather is not None:
        gather.times = [0.0, 0.0, 0.0, 0]                             # per-phase host clocks of the timed steps only
    call_ms = []
    try:
        gpu_uuid = torch.cuda.get_device_properties(dev).uuid
    except Exception:
        gpu_uuid = None
    with (ClockSampler(local_rank, gpu_uuid) if rank == 0 else NoSampler()) as clocks:   # rank 0's GPU stands for the box
        t0 = time.perf_counter()
        ev0.record(lib_stream)
        for _ in range(args.steps):
            last = step()
            s = ctx.stats()
            kde_ms.append(s["ms_kde"]); launches += s["launches"] + (2 if world > 1 else 0); call_ms.append(s["ms_call"])   # + the two scatter kernels of the gather
        if world > 1:
            lib_stream.wait_stream(torch.cuda.current_stream())      # the NCCL gather of the last step runs on torch's stream
        ev1.record(lib_stream)
        barrier()
        dt_wall = time.perf_counter() - t0
    stats = ctx.stats()
    gather_times = list(gather.times) if gather is not None and getattr(gather, "times", None) else [0.0, 0.0, 0.0, 1]
    ev1.synchronize()
    dt = ev0.elapsed_time(ev1) * 1e-3
    tmax = torch.tensor([dt, dt_wall], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    dt, dt_wall = float(tmax[0].item()), float(tmax[1].item())
    ms_per_step = dt / args.steps * 1e3
    wall_ms_per_step = dt_wall / args.steps * 1e3

    # e2e: pinned host arrays -> device every step (same two clocks); once per host format
    e2e_steps = max(1, min(args.steps, 3))

    def time_e2e(loader):
        loader(); step()                                             # untimed: buffers of this path reach their size
        barrier()
        t0 = time.perf_counter()
        ev0.record(lib_stream)
        for _ in range(e2e_steps):
            loader()
            res = step()
        if world > 1:
            lib_stream.wait_stream(torch.cuda.current_stream())
        ev1.record(lib_stream)
        barrier()
        wall = time.perf_counter() - t0
        ev1.synchronize()
        e2e = torch.tensor([ev0.elapsed_time(ev1) * 1e-3, wall], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(e2e, op=dist.ReduceOp.MAX)
        return float(e2e[0].item()) / e2e_steps * 1e3, float(e2e[1].item()) / e2e_steps * 1e3, res

    alt_ms, alt_wall_ms, _ = time_e2e(load_alt)
    e2e_ms, e2e_wall_ms, last = time_e2e(load)
    alt_format = "plain" if args.host_format == "packed" else "packed"
    n_rows = int(last["n_rows"]) if last else 0
    d2h = n_rows * (41 + 4 * n_samples)

    # roofline of the fused KDE kernel (rank-local numbers; every rank runs the same kernel on its shard)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_gbs, peak_src = (peaks.get("hbm_gbs"), "measured") if peaks.get("hbm_gbs") else (6650.0, "fallback")
    traffic = None
    try:    # dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture of this kernel on this workload (profiles/)
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        key = "%s_bw%d" % (args.workload, args.bandwidth)
        if world == 1 and not (args.samples or args.tags) and key in tj:
            traffic = tj[key]["span_kernel_dram_bytes"]
    except Exception:
        pass
    kde_avg_ms = float(np.mean(kde_ms)) if kde_ms else float("nan")
    n_runs_bytes = (40 + 4 * n_samples) * stats["n_candidates"]     # run record + kurtosis + per-sample hits
    alg_bytes = 8 * E + n_runs_bytes
    achieved = alg_bytes / (kde_avg_ms * 1e-3) / 1e9 if kde_avg_ms > 0 else 0.0
    roofline = {"bound": "hbm", "kernel": "span_kernel<directional>" if directional else "span_kernel<nondirectional>", "achieved": achieved, "peak": peak_gbs,
                "peak_source": peak_src, "unit": "GB/s", "frac": achieved / peak_gbs, "traffic": traffic,
                "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": kde_avg_ms, "kernel_share_of_step": kde_avg_ms / ms_per_step}

    # where the step goes on every rank (host clocks of the last call and of the gather phases, averaged over the gather's calls)
    per_rank = None
    if world > 1:
        gt = gather_times
        k = max(1, gt[3])
        mine = torch.tensor([float(np.mean(call_ms)) if call_ms else 0.0, stats["ms_total"], gt[0] / k * 1e3, gt[1] / k * 1e3, gt[2] / k * 1e3, float(E)], dtype=torch.float64, device=dev)
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        per_rank = {"columns": ["call_ms", "device_ms", "allreduce_counts_ms", "scatter_ms", "completion_ms", "tag_entries"],
                    "rows": [[round(float(x), 4) for x in t.cpu().tolist()] for t in allr]}

    if rank == 0:
        line = {"metric": metric, "value": genome / (ms_per_step * 1e-3), "unit": "bp/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_per_step, "wall_ms_per_step": wall_ms_per_step, "timing": "CUDA events on the library's stream around the K steps (max over ranks); wall clock alongside",
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": config, "clocks": clocks.summary(),
                "e2e": {"value": genome / (e2e_ms * 1e-3), "unit": "bp/s", "ms_per_step": e2e_ms, "wall_ms_per_step": e2e_wall_ms, "h2d_bytes_per_step": h2d_bytes[args.host_format], "d2h_bytes_per_step": d2h,
                        "host_format": args.host_format + (" (u16 position deltas + u8 counts + exception list, decoded on the device: upb_load_tags_packed)" if args.host_format == "packed" else " (u32 pos + u32 cnt: upb_load_tags)")},
                "e2e_alt": {"host_format": alt_format, "value": genome / (alt_ms * 1e-3), "unit": "bp/s", "ms_per_step": alt_ms, "wall_ms_per_step": alt_wall_ms, "h2d_bytes_per_step": h2d_bytes[alt_format]},
                "gpu_launches": launches, "roofline": roofline,
                "result": {"regions": n_rows, "candidates": stats["n_candidates"], "tag_entries_rank0": E, "spans": stats["kde_tiles"], "warm_spans": stats["warm_spans"], "pieces": stats["kde_hot_tiles"],
                           "hot_blocks": stats["hot_blocks"], "exact_evals": stats["exact_evals"], "stitched": stats["n_stitched"],
                           "device_ms_per_step_rank0": stats["ms_total"], "call_ms_rank0": stats["ms_call"]}}
        if world == 1 and not args.no_cpu:
            cs = CPU_SAMPLE_CONTIGS[args.workload]
            sample = cpu_sample(data, cs)
            with tempfile.TemporaryDirectory() as wd:
                secs, kind, ref = run_cpu_reference(sample, st, 1, wd)
            bp = int(sum(sizes[c] for c in cs))
            # parity spot-check on the sample: every region the CPU side finds on these contigs must equal ours
            ctx.release_view()
            full = ctx.call_regions(**dict(st, reference_order=True, rows=True))
            ours = [r for r in full["rows"] + full["dropped"] if r["contig"] in cs]
            theirs = ref["rows"] + ref["dropped"]
            key_o = lambda r: (cs.index(r["contig"]), r["left"], r["len"], r["peak"], tuple(r["counts"]), r["kurtosis"])
            key_t = lambda r: (r["contig"], r["left"], r["len"], r["peak"], tuple(r["counts"]), r["kurtosis"])
            same = sorted(key_o(r) for r in ours) == sorted(key_t(r) for r in theirs)
            line["cpu_baseline"] = {"value": bp / secs, "unit": "bp/s", "cores": 1, "kind": kind, "seconds": secs,
                                    "sample": "contigs %s (%d bp, %d tag entries) of the workload, all samples, workload-wide thresholds"
                                              % ("+".join(names[c] for c in cs), bp, int(sample.pos.shape[0])),
                                    "host_cores_available": os.cpu_count(), "rows_match_gpu": bool(same), "rows": len(theirs)}
        if per_rank is not None:
            line["result"]["per_rank_ms"] = per_rank
        if world == 1:
            # last of all, so that nothing above depends on it: the 2-byte host form as a side leg (see load_packed2)
            try:
                ctx.release_view()
                p2_ms, p2_wall_ms, res2 = time_e2e(load_packed2)
                line["e2e_packed2"] = {"host_format": "packed2 (u16 per entry: 12-bit position delta + 4-bit count, exception list; upb_load_tags_packed2)",
                                       "value": genome / (p2_ms * 1e-3), "unit": "bp/s", "ms_per_step": p2_ms, "wall_ms_per_step": p2_wall_ms,
                                       "h2d_bytes_per_step": h2d_bytes["packed2"], "rows_equal_headline_leg": bool(int(res2["n_rows"]) == n_rows)}
            except Exception as e:                                    # noqa: BLE001 — must never take the line down with it
                line["e2e_packed2"] = {"host_format": "packed2", "error": str(e)[:200]}
        print(json.dumps(line))
    if gather is not None and hasattr(gather, "close"):
        gather.close()
    try:
        ctx.close()
    except Exception:                                                 # noqa: BLE001
        pass
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
