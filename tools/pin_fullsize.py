#!/usr/bin/env python
"""Full-size parity pins (VERDICT r01 "missing #5"): run the compiled, unmodified reference (oracle/_ref/ref_replay — its own
Kernel / ProfileBuffer / Region objects) over WHOLE BASELINE configurations in the CPU container and commit what it produced as
per-column SHA-256 fingerprints, tests/golden/fullsize_<config>.json.  No GPU is involved: the inputs come from the
deterministic generator (unipeak_b200.synth_det), which yields the same arrays on the GPU box, where
tests/test_fullsize_pins.py regenerates them, checks the input fingerprint, runs the CUDA path and compares every column.

    python tools/pin_fullsize.py c3 c2 [c1] [--scale 0.1] [--workdir /tmp/pin]

c3: regions, 16 x 30 M tags, hg19, directional (the bench workload, same seed and settings as bench.py)
c2: strand_shift per sample + regions -D on 2 x 20 M tags, hg19
c1: regions, 1 x 1 M tags, chr21, -m whole genome
Needs /root/reference only through `make -C oracle ref` having been run (oracle/_ref/ref_replay).
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from unipeak_b200.synth import HG19_SIZES  # noqa: E402
from unipeak_b200.synth_det import fingerprint, make_synthetic_det  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
CANON_NAN = np.uint64(0x7FF8000000000000)

CONFIGS = {
    # name: (n_samples, tags, directional, sizes, mappable, seed)
    "c3": (16, 30_000_000, True, HG19_SIZES, 0, 20261018),
    "c2": (2, 20_000_000, False, HG19_SIZES, 0, 20261019),
    "c1": (1, 1_000_000, True, [48129895], 3095693983, 20261020),
}


def region_settings(tags, n, directional, mappable, genome, bandwidth=50):
    """bench.settings_from_totals: what src/regions.cpp:152-232 derives with default flags."""
    total = int(sum(tags))
    background = float(total) / float(mappable or genome)
    if directional:
        background /= 2
    return dict(bandwidth=bandwidth, region_threshold=25.0, kurtosis_threshold=50.0, corr_threshold=(-1.0 if directional else 0.3),
                hit_threshold=10.0 * n, directional=directional, control=np.zeros(n, np.uint8), coeffs=[], background=background)


def canon_bits(x: np.ndarray) -> np.ndarray:
    """float64 -> uint64 bit patterns with every NaN mapped to one pattern (the sign / payload of a NaN is not a result)."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    b = x.view(np.uint64).copy()
    b[np.isnan(x)] = CANON_NAN
    return b


def column_hashes(cols: dict) -> dict:
    """cols: forward_label u8, contig/left/len/peak u32, peak_score/kurtosis/corr f64, counts u32 [R, n] — rows in emission order,
    dropped rows appended.  The same function runs on the GPU box over the CUDA path's table."""
    out = {}
    for k, dt in (("forward_label", np.uint8), ("contig", "<u4"), ("left", "<u4"), ("len", "<u4"), ("peak", "<u4")):
        out[k] = hashlib.sha256(np.ascontiguousarray(cols[k]).astype(dt).tobytes()).hexdigest()
    for k in ("peak_score", "kurtosis", "corr"):
        out[k] = hashlib.sha256(canon_bits(cols[k]).astype("<u8").tobytes()).hexdigest()
    out["counts"] = hashlib.sha256(np.ascontiguousarray(cols["counts"]).astype("<u4").tobytes()).hexdigest()
    return out


def parse_replay_regions(path: str, n: int):
    with open(path, "rb") as f:
        buf = f.read()
    n_rows, n_dropped, n_passed, n_rejected = np.frombuffer(buf, "<u8", 4, 0).tolist()
    tags = np.frombuffer(buf, "<u8", n, 32)
    rec = np.dtype([("forward_label", "u1"), ("contig", "<u4"), ("left", "<u4"), ("len", "<u4"), ("peak", "<u4"),
                    ("peak_score", "<f8"), ("kurtosis", "<f8"), ("corr", "<f8"), ("counts", "<u4", (n,))])
    rows = np.frombuffer(buf, rec, n_rows + n_dropped, 32 + 8 * n)
    cols = {k: np.ascontiguousarray(rows[k]) for k in rec.names}
    return dict(n_rows=int(n_rows), n_dropped=int(n_dropped), n_passed=int(n_passed), n_rejected=int(n_rejected),
                tags_in_regions=[int(x) for x in tags]), cols


def hexf(x: float) -> str:
    return float(x).hex()


def replay_regions(data, st, workdir):
    inp, outp = os.path.join(workdir, "in.upbi"), os.path.join(workdir, "out.bin")
    data.write_upbi(inp)
    args = [os.path.join(ROOT, "oracle", "_ref", "ref_replay"), "regions", inp, outp, "bw=%d" % st["bandwidth"], "r=" + hexf(st["region_threshold"]),
            "k=" + hexf(st["kurtosis_threshold"]), "u=" + hexf(st["corr_threshold"]), "t=" + hexf(st["hit_threshold"]),
            "background=" + hexf(st["background"]), "directional=%d" % (1 if st["directional"] else 0), "reps=1"]
    t0 = time.time()
    p = subprocess.run(args, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    if p.returncode != 0:
        raise RuntimeError("ref_replay failed: " + p.stderr[-1000:])
    secs = float(p.stdout.strip().split("=")[1])
    head, cols = parse_replay_regions(outp, data.n_samples)
    os.unlink(inp)
    os.unlink(outp)
    return head, cols, secs, time.time() - t0


def pin(name: str, scale: float, workdir: str):
    import refutil
    n, tags, directional, sizes, mappable, seed = CONFIGS[name]
    tags = int(round(tags * scale))
    genome = int(sum(sizes))
    t0 = time.time()
    data = make_synthetic_det(sizes, n, tags, directional=directional, seed=seed, bandwidth=50, device="cpu")
    t_gen = time.time() - t0
    totals = [data.tags(s) for s in range(n)]
    st = region_settings(totals, n, directional, mappable, genome)
    out = dict(config=name, scale=scale, n_samples=n, tags_per_sample=tags, directional=directional, seed=seed, genome_bp=genome,
               input_sha256=fingerprint(data), tag_entries=int(data.pos.shape[0]), tag_totals=totals,
               settings={k: (hexf(v) if isinstance(v, float) else v) for k, v in st.items() if k not in ("control", "coeffs")},
               generated_by="tools/pin_fullsize.py on the compiled reference (oracle/_ref/ref_replay), CPU container", generator_seconds=round(t_gen, 1))
    head, cols, secs, wall = replay_regions(data, st, workdir)
    out["regions"] = dict(head, reference_seconds=secs, reference_bp_per_s=genome / secs, columns=column_hashes(cols),
                          rows_per_contig=np.bincount(cols["contig"], minlength=len(sizes)).tolist())
    print("%s regions: %d rows + %d dropped, reference %.1f s (%.1f Mbp/s)" % (name, head["n_rows"], head["n_dropped"], secs, genome / secs / 1e6), flush=True)
    if name == "c2":
        out["strand_shift"] = []
        for s in range(n):
            one = data.subset_samples([s])
            r = refutil.run_ref_replay_shift(one, workdir)
            t = r["tested"]
            out["strand_shift"].append(dict(
                sample=s, n_regions_found=r["n_regions_found"], n_regions_tested=r["n_regions_tested"], n_tags_in_regions=r["n_tags_in_regions"],
                best_shift=r["best_shift"], shift_freq=r["shift_freq"], reference_seconds=r["seconds"],
                tested_sha256=hashlib.sha256(b"".join(np.asarray([x[k] for x in t], dtype=dt).tobytes() for k, dt in
                                                      (("contig", "<u4"), ("left", "<u4"), ("len", "<u4"), ("best_shift", "<u4"))) +
                                             canon_bits(np.asarray([x["best_corr"] for x in t])).tobytes()).hexdigest()))
            print("%s strand_shift sample %d: best_shift %d, %d tested of %d, reference %.1f s" % (name, s, r["best_shift"], r["n_regions_tested"], r["n_regions_found"], r["seconds"]), flush=True)
            os.unlink(os.path.join(workdir, "in.upbi"))
    suffix = "" if scale == 1.0 else "_s%g" % scale
    path = os.path.join(GOLDEN, "fullsize_%s%s.json" % (name, suffix))
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
        f.write("\n")
    print("wrote", path, flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("configs", nargs="+", choices=sorted(CONFIGS))
    ap.add_argument("--scale", type=float, default=1.0, help="fraction of the tags per sample (1.0 = the BASELINE configuration)")
    ap.add_argument("--workdir", default="/tmp/upb_pin")
    a = ap.parse_args()
    os.makedirs(a.workdir, exist_ok=True)
    for c in a.configs:
        pin(c, a.scale, a.workdir)


if __name__ == "__main__":
    main()
