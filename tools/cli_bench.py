#!/usr/bin/env python
"""End-to-end timing of the drop-in command-line drivers beside the reference's binaries (oracle/_ref), files in, files out:
config 1 of BASELINE.json (bin/regions on one directional 3SEQ sample, chr21 only, 1 M tags, -m whole-genome — the reference's own
CPU-runnable case) and a mid-size pooled call (4 samples x 3 M tags on chr1..chr5).  Parsing and writing dominate here; the
hot path is timed by bench.py.  Prints one JSON line per case; the outputs of the two programs are compared byte for byte.

    python tools/cli_bench.py [--workdir DIR]
"""
import argparse, json, os, subprocess, sys, tempfile, time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from unipeak_b200.synth import HG19_NAMES, HG19_SIZES, make_synthetic  # noqa: E402

REF, NEW = os.path.join(ROOT, "oracle", "_ref"), os.path.join(ROOT, "bin")


def timed(cmd, cwd, reps):
    best = None
    for _ in range(reps):
        t0 = time.perf_counter()
        p = subprocess.run(cmd, cwd=cwd, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE)
        dt = time.perf_counter() - t0
        if p.returncode != 0:
            raise RuntimeError("%s failed: %s" % (cmd[0], p.stderr.decode()[-500:]))
        best = dt if best is None else min(best, dt)
    return best


def case(name, sizes, names, n_samples, tags, extra_args, wd, reps=3):
    d = make_synthetic(sizes, n_samples, tags, directional=True, seed=2026, contig_names=names)
    d.write_contig_table(os.path.join(wd, "contigs.txt"))
    wigs = []
    for s in range(n_samples):
        d.write_wig(os.path.join(wd, "s%d.wig" % s), s, "s%d" % s)
        wigs.append("s%d.wig" % s)
    wig_bytes = sum(os.path.getsize(os.path.join(wd, w)) for w in wigs)
    args = ["-q", "-f", "-c", "contigs.txt"] + extra_args
    out = {}
    for which, bindir in (("reference", REF), ("ours", NEW)):
        if not os.access(os.path.join(bindir, "regions"), os.X_OK):
            out[which] = None
            continue
        out[which] = timed([os.path.join(bindir, "regions")] + args + ["-o", which + ".txt"] + wigs, wd, reps)
    same = None
    if out["reference"] is not None and out["ours"] is not None:
        same = open(os.path.join(wd, "reference.txt"), "rb").read() == open(os.path.join(wd, "ours.txt"), "rb").read()
    rows = sum(1 for l in open(os.path.join(wd, "ours.txt")) if l[:1] not in "#\t\n")
    trace = None
    if out["ours"] is not None:                                        # where our program's time goes: the library's and the driver's own timeline
        t0 = time.perf_counter()
        p = subprocess.run([os.path.join(NEW, "regions")] + args + ["-o", "ours_trace.txt"] + wigs, cwd=wd, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE,
                           env=dict(os.environ, UPB_TRACE="1"))
        trace = {"seconds": time.perf_counter() - t0, "lines": [l for l in p.stderr.decode().splitlines() if l.startswith("[upb]")][:40]}
    print(json.dumps({"case": name, "samples": n_samples, "tags_per_sample": tags, "genome_bp": int(sum(sizes)), "wig_bytes": wig_bytes,
                      "seconds_reference_cli": out["reference"], "seconds_ours_cli": out["ours"], "regions": rows, "identical_output": same, "trace_ours": trace,
                      "note": "wall clock of the whole program (process start, CUDA context creation ~0.3 s, parse, call, write); best of %d" % reps}))
    sys.stdout.flush()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workdir", default=None)
    a = ap.parse_args()
    with tempfile.TemporaryDirectory(dir=a.workdir) as wd:
        case("c1: regions, 1 sample x 1M tags, chr21, -m 3095693983", [48129895], ["chr21"], 1, 1_000_000, ["-m", "3095693983"], wd)
        case("pooled: regions, 4 samples x 3M tags, chr1..chr5", HG19_SIZES[:5], HG19_NAMES[:5], 4, 3_000_000, [], wd)


if __name__ == "__main__":
    main()
