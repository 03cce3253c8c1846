#!/usr/bin/env python
"""Generates the golden fixtures in this directory by running the UNMODIFIED reference (oracle/_ref, built
by `make -C oracle ref` from /root/reference) on small seeded inputs.

    python tests/golden/make_golden.py

Each fixture is one .npz: the input arrays (contig sizes, offsets, positions, counts), the derived settings
(src/regions.cpp:104-232) and the reference's results with every double stored as its IEEE-754 bit pattern
(uint64), so comparisons are bit-exact.  The reference ships no vectors of its own; these
are outputs of the reference itself, which is what pins the oracle and the CUDA path on the GPU box, where
/root/reference does not exist.
"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import refutil  # noqa: E402
from unipeak_b200.synth import TagData, make_synthetic  # noqa: E402


def bits(xs):
    return np.asarray(xs, dtype=np.float64).view(np.uint64)


def rows_to_arrays(rows, n):
    if not rows:
        z = np.zeros(0, np.uint32)
        return dict(forward_label=z.astype(np.uint8), contig=z, left=z, len=z, peak=z, peak_score=bits([]), kurtosis=bits([]), corr=bits([]),
                    counts=np.zeros((0, n), np.uint32))
    return dict(forward_label=np.array([r["forward_label"] for r in rows], np.uint8), contig=np.array([r["contig"] for r in rows], np.uint32),
                left=np.array([r["left"] for r in rows], np.uint32), len=np.array([r["len"] for r in rows], np.uint32),
                peak=np.array([r["peak"] for r in rows], np.uint32), peak_score=bits([r["peak_score"] for r in rows]),
                kurtosis=bits([r["kurtosis"] for r in rows]), corr=bits([r["corr"] for r in rows]),
                counts=np.array([r["counts"] for r in rows], np.uint32).reshape(len(rows), n))


def save_regions(name, data: TagData, **kw):
    st = refutil.region_settings(data, **kw)
    with tempfile.TemporaryDirectory() as wd:
        ref = refutil.run_ref_replay_regions(data, st, wd)
    out = dict(kind=np.array("regions"), contig_sizes=data.contig_sizes, n_samples=np.array(data.n_samples), off=data.off, pos=data.pos, cnt=data.cnt,
               bandwidth=np.array(st["bandwidth"]), region_threshold=bits([st["region_threshold"]]), kurtosis_threshold=bits([st["kurtosis_threshold"]]),
               corr_threshold=bits([st["corr_threshold"]]), hit_threshold=bits([st["hit_threshold"]]), directional=np.array(int(st["directional"])),
               control=np.asarray(st["control"], np.uint8), coeffs=bits(st["coeffs"]), background=bits([st["background"]]),
               n_passed=np.array(ref["n_passed"]), n_rejected=np.array(ref["n_rejected"]), tags_in_regions=np.array(ref["tags_in_regions"], np.uint64))
    for prefix, rows in (("row_", ref["rows"]), ("drop_", ref["dropped"])):
        for k, v in rows_to_arrays(rows, data.n_samples).items():
            out[prefix + k] = v
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print("%-28s %6d entries, %4d rows + %d dropped" % (name, data.pos.shape[0], len(ref["rows"]), len(ref["dropped"])))


def save_shift(name, data: TagData, **kw):
    with tempfile.TemporaryDirectory() as wd:
        ref = refutil.run_ref_replay_shift(data, wd, **kw)
    total = sum(data.tags(s) for s in range(data.n_samples))
    p = dict(bandwidth=50, region_threshold=25.0, kurtosis_threshold=50.0, corr_threshold=0.3, hit_threshold=10, min_shift=25, max_shift=150,
             n_regions_to_test=1000)
    p.update(kw)
    out = dict(kind=np.array("shift"), contig_sizes=data.contig_sizes, n_samples=np.array(data.n_samples), off=data.off, pos=data.pos, cnt=data.cnt,
               background=bits([float(total) / float(data.genome_size())]), n_regions_found=np.array(ref["n_regions_found"]),
               n_regions_tested=np.array(ref["n_regions_tested"]), n_tags_in_regions=np.array(ref["n_tags_in_regions"]), best_shift=np.array(ref["best_shift"]),
               shift_freq=np.array(ref["shift_freq"], np.uint64), tested_contig=np.array([t["contig"] for t in ref["tested"]], np.uint32),
               tested_left=np.array([t["left"] for t in ref["tested"]], np.uint32), tested_len=np.array([t["len"] for t in ref["tested"]], np.uint32),
               tested_best_shift=np.array([t["best_shift"] for t in ref["tested"]], np.uint32), tested_best_corr=bits([t["best_corr"] for t in ref["tested"]]))
    for k, v in p.items():
        out["p_" + k] = bits([v]) if isinstance(v, float) else np.array(v)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print("%-28s %6d entries, %d regions tested, best shift %d" % (name, data.pos.shape[0], ref["n_regions_tested"], ref["best_shift"]))


def save_tags_in_regions(name, data: TagData, base_samples, extra_samples, directional=True, extension=0):
    """bin/regions on the base samples, then bin/tags_in_regions (the real reference CLIs) for the extra samples."""
    import re
    base = data.subset_samples(base_samples)
    extra = data.subset_samples(extra_samples)
    with tempfile.TemporaryDirectory() as wd:
        f0 = refutil.write_inputs(base, os.path.join(wd, "base"))
        f1 = refutil.write_inputs(extra, os.path.join(wd, "extra"), prefix="x")
        table = os.path.join(wd, "regions.txt")
        refutil.run_ref_cli("regions", ["-q", "-o", table, "-c", f0["contigs"]] + (["-D"] if not directional else []) + f0["wigs"])
        out = os.path.join(wd, "more.txt")
        refutil.run_ref_cli("tags_in_regions", ["-f", table, "-o", out, "-c", f0["contigs"], "-e", str(extension)] + (["-D"] if not directional else []) + f1["wigs"])
        names = data.names()
        keys = []
        for row in refutil.read_table_rows(table):
            m = re.match(r"(\w+):(\d+)-(\d+)", row)
            keys.append((names.index(m.group(1)), int(m.group(2)), int(m.group(3))))
        n = len(extra_samples)
        got = [[int(x) for x in row.split("\t")[-n:]] for row in refutil.read_table_rows(out)]
    np.savez_compressed(os.path.join(HERE, name + ".npz"), kind=np.array("tir"), contig_sizes=extra.contig_sizes, n_samples=np.array(extra.n_samples),
                        off=extra.off, pos=extra.pos, cnt=extra.cnt, directional=np.array(int(directional)), extension=np.array(extension),
                        keys=np.array(keys, np.uint32).reshape(len(keys), 3), counts=np.array(got, np.uint32).reshape(len(keys), n))
    print("%-28s %6d entries, %4d region rows x %d extra samples" % (name, extra.pos.shape[0], len(keys), n))


def dense(sizes, n_samples, seed, fill, max_count):
    rng = np.random.default_rng(seed)
    segs = []
    for _ in range(n_samples):
        per = [[None] * len(sizes) for _ in range(2)]
        for d in range(2):
            for c, size in enumerate(sizes):
                pos = np.arange(51, size - 49, dtype=np.uint32)
                pos = pos[rng.random(pos.shape[0]) < fill]
                per[d][c] = (pos, rng.integers(1, max_count + 1, size=pos.shape[0]).astype(np.uint32))
        segs.append(per)
    return TagData.from_segments(sizes, segs)


def main():
    if not refutil.have_ref():
        raise SystemExit("oracle/_ref is missing: run `make -C oracle ref` where /root/reference is mounted")
    save_regions("regions_directional", make_synthetic([120000, 60000, 30000], 3, 12000, directional=True, seed=101, peak_spacing=4000))
    save_regions("regions_nondirectional", make_synthetic([120000, 60000], 2, 16000, directional=False, seed=102, peak_spacing=4000), directional=False)
    save_regions("regions_controls_q4", make_synthetic([120000, 60000], 4, 10000, directional=True, seed=103, peak_spacing=4000), control=[1, 3])
    save_regions("regions_coeffs_q3", make_synthetic([120000, 60000], 3, 10000, directional=True, seed=104, peak_spacing=4000), coeffs=[1.0, 2.0, 0.5])
    save_regions("regions_left_edge_q2", make_synthetic([40000, 20000], 2, 16000, directional=True, seed=105, peak_spacing=1500, keep_clear_of_edges=False))
    save_regions("regions_one_contig_q7", make_synthetic([600000], 1, 20000, directional=True, seed=106, peak_spacing=8000), mappable=100000000)
    save_regions("regions_bw200", make_synthetic([150000, 50000], 2, 16000, directional=True, seed=107, peak_spacing=6000, bandwidth=200), bandwidth=200)
    save_regions("regions_bw10", make_synthetic([150000, 50000], 2, 16000, directional=True, seed=108, peak_spacing=6000), bandwidth=10)
    d = dense([70000, 3000], 1, 109, 0.9, 3)
    save_regions("regions_long_q8", d, mappable=100 * d.genome_size())
    save_regions("regions_long_q8_nondir", d, mappable=100 * d.genome_size(), directional=False, corr_threshold=-0.5)
    d3 = make_synthetic([120000, 60000], 4, 12000, directional=True, seed=111, peak_spacing=4000, contig_names=["chrA", "chrB"])
    save_tags_in_regions("tir_directional", d3, [0, 1], [2, 3])
    save_tags_in_regions("tir_directional_ext", d3, [0], [1, 2, 3], extension=75)
    d4 = make_synthetic([120000, 60000], 3, 16000, directional=False, seed=112, peak_spacing=4000, contig_names=["chrA", "chrB"])
    save_tags_in_regions("tir_nondirectional", d4, [0], [1, 2], directional=False, extension=20)
    d5 = make_synthetic([300000], 3, 16000, directional=True, seed=113, peak_spacing=6000, contig_names=["chrA"])
    save_tags_in_regions("tir_one_contig_q13", d5.subset_samples([0, 1, 2]), [0], [1, 2])
    save_shift("shift_chip", make_synthetic([1500000], 1, 150000, directional=False, seed=110, peak_spacing=10000),
               n_regions_to_test=25, region_threshold=8.0, corr_threshold=0.1)


if __name__ == "__main__":
    main()
