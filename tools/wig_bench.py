#!/usr/bin/env python
"""Throughput of the wiggle front end (misc/format.cpp:503-565): bin/host_selftest wigtime on one 30 M-tag whole-genome
sample, sequential feeder (UPB_WIG_THREADS=1) against the parallel one (byte ranges tokenised by threads) with 2, 4, 8 and all
threads; the arrays read must be identical (checksum).  Host only — prints one JSON line."""
import json, os, subprocess, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from unipeak_b200.synth import HG19_NAMES, HG19_SIZES, make_synthetic  # noqa: E402

tags = int(sys.argv[1]) if len(sys.argv) > 1 else 30_000_000
with tempfile.TemporaryDirectory() as wd:
    d = make_synthetic(HG19_SIZES, 1, tags, directional=True, seed=7, contig_names=HG19_NAMES)
    d.write_contig_table(os.path.join(wd, "contigs.txt"))
    d.write_wig(os.path.join(wd, "s.wig"), 0, "s")
    size = os.path.getsize(os.path.join(wd, "s.wig"))
    out = {"file_bytes": size, "entries": None, "cores": os.cpu_count(), "runs": []}
    sums = set()
    for th in ("1", "2", "4", "8", str(os.cpu_count())):
        best = None
        for _ in range(2):
            p = subprocess.run([os.path.join(ROOT, "bin", "host_selftest"), "wigtime", "contigs.txt", "s.wig"], cwd=wd, stdout=subprocess.PIPE,
                               env=dict(os.environ, UPB_WIG_THREADS=th))
            kv = dict(l.split("\t") for l in p.stdout.decode().splitlines() if "\t" in l)
            sec = float(kv["seconds"])
            best = sec if best is None else min(best, sec)
            sums.add(kv["checksum"]); out["entries"] = int(kv["entries"])
        out["runs"].append({"threads": int(th), "seconds": best, "MB_per_s": size / best / 1e6, "entries_per_s": out["entries"] / best})
    out["identical_arrays"] = len(sums) == 1
    out["speedup_all_threads"] = out["runs"][0]["seconds"] / out["runs"][-1]["seconds"]
    print(json.dumps(out))
