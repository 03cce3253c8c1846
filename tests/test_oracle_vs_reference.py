"""Pins the CPU oracle (oracle/unipeak_oracle.c) to the UNMODIFIED reference compiled into oracle/_ref.

Two layers:
  * ref_replay (the reference's own Kernel / ProfileBuffer / Region objects fed from arrays): every raw
    double compared bit for bit;
  * the reference's real command-line binaries (bin/regions, bin/strand_shift, bin/tags_in_regions) fed
    canonical wiggle text: their output rows compared as text with the oracle rows formatted the way
    misc/format.cpp:1141-1162 formats them — this also validates the replay harness' merge loop, which is
    ours, against the reference's own main().

The reference ships no tests or golden vectors; this differential test IS the pin.  It runs
wherever oracle/_ref exists (build container: `make -C oracle ref`; the GPU box gets the built files).
"""
import math
import os
import re

import numpy as np
import pytest

import refutil
from refutil import (format_region_row, oracle_regions, oracle_strand_shift, oracle_tags_in_regions, read_table_rows,
                     region_settings, run_ref_cli, run_ref_replay_regions, run_ref_replay_shift, write_inputs)
from unipeak_b200.synth import TagData, make_synthetic

pytestmark = pytest.mark.skipif(not refutil.have_ref(), reason="compiled reference (oracle/_ref) not present")

ROW_KEYS = ("forward_label", "contig", "left", "len", "peak", "counts")
FP_KEYS = ("peak_score", "kurtosis", "corr")


def same_fp(a, b):
    return a == b or (math.isnan(a) and math.isnan(b))


def assert_same_rows(got, want, what):
    assert len(got) == len(want), "%s: oracle %d rows, reference %d" % (what, len(got), len(want))
    for i, (g, w) in enumerate(zip(got, want)):
        for k in ROW_KEYS:
            assert g[k] == w[k], "%s row %d %s: oracle %r reference %r" % (what, i, k, g[k], w[k])
        for k in FP_KEYS:
            assert same_fp(g[k], w[k]), "%s row %d %s: oracle %r reference %r" % (what, i, k, g[k], w[k])


def check_replay(data, tmp_path, **kw):
    st = region_settings(data, **kw)
    ours = oracle_regions(data, st)
    assert ours["status"] == 0
    ref = run_ref_replay_regions(data, st, str(tmp_path))
    assert_same_rows(ours["rows"], ref["rows"], "written")
    assert_same_rows(ours["dropped"], ref["dropped"], "dropped")
    assert ours["n_passed"] == ref["n_passed"] and ours["n_rejected"] == ref["n_rejected"]
    assert ours["tags_in_regions"] == ref["tags_in_regions"]
    return ours


def test_replay_directional(tmp_path):
    d = make_synthetic([200000, 100000, 50000], 3, 20000, directional=True, seed=1, peak_spacing=5000)
    assert len(check_replay(d, tmp_path)["rows"]) > 10


def test_replay_nondirectional_corr(tmp_path):
    d = make_synthetic([200000, 100000, 50000], 2, 30000, directional=False, seed=2, peak_spacing=5000)
    check_replay(d, tmp_path, directional=False)


def test_replay_controls_q4(tmp_path):
    d = make_synthetic([300000, 150000], 4, 25000, directional=True, seed=5, peak_spacing=5000)
    check_replay(d, tmp_path, control=[1, 3])
    check_replay(d, tmp_path, directional=False, control=[2])


def test_replay_coefficients_q3(tmp_path):
    d = make_synthetic([300000, 150000], 3, 25000, directional=True, seed=15, peak_spacing=5000)
    check_replay(d, tmp_path, coeffs=[1.0, 2.0, 0.5])
    check_replay(d, tmp_path, proportional=True)
    check_replay(d, tmp_path, directional=False, coeffs=[0.7, 1.3], control=[1])


def test_replay_left_edge_q2(tmp_path):
    # tags inside the first `bandwidth` positions of a contig: the reference's window does not slide there
    d = make_synthetic([60000, 30000], 2, 30000, directional=True, seed=21, peak_spacing=2000, keep_clear_of_edges=False)
    assert min(int(d.seg(s, dd, c)[0][0]) for s in range(2) for dd in range(2) for c in range(2)) <= 50
    check_replay(d, tmp_path)
    check_replay(d, tmp_path, directional=False)


@pytest.mark.parametrize("variant", [0, 1, 2, 3])
def test_replay_left_edge_window_inherited_by_the_next_contig(tmp_path, variant):
    # every tag of a contig within `bandwidth` of its start: the flush leaves stale cells (and sometimes an open region) for the
    # next contig of the same buffer - also when that contig is touched by the control sample only (variant 1)
    d, ctrl = refutil.chain_leak_data(variant)
    kw = dict(region_threshold=3.0, hit_threshold=2.0, kurtosis_threshold=0.0, mappable=20 * d.genome_size())
    if ctrl:
        kw["control"] = ctrl
    ours = check_replay(d, tmp_path, **kw)
    if variant in (0, 1, 2):
        assert any(r["contig"] == 2 and r["left"] == 2 for r in ours["rows"])       # opened on contig 1, written with contig 2
    check_replay(d, tmp_path, directional=False, corr_threshold=-1.0, **kw)
    check_replay(d, tmp_path, directional=False, corr_threshold=0.0, **kw)


def test_replay_strand_shift_long_region_at_contig_start(tmp_path):
    d = refutil.long_start_region_data()
    kw = dict(n_regions_to_test=30, region_threshold=4.0, corr_threshold=0.05, hit_threshold=5)
    ours = oracle_strand_shift(d, **kw)
    ref = run_ref_replay_shift(d, str(tmp_path), **kw)
    assert ours["status"] == 0
    assert any(t["contig"] == 0 and t["left"] <= 50 for t in ours["tested"])       # the region at the contig start is swept
    for k in ("n_regions_found", "n_regions_tested", "n_tags_in_regions", "best_shift", "shift_freq"):
        assert ours[k] == ref[k], k
    for g, w in zip(ours["tested"], ref["tested"]):
        assert (g["contig"], g["left"], g["len"], g["best_shift"]) == (w["contig"], w["left"], w["len"], w["best_shift"])
        assert g["best_corr"] == w["best_corr"]


def test_replay_single_sample_single_contig_q7(tmp_path):
    d = make_synthetic([2000000], 1, 60000, directional=True, seed=3, peak_spacing=10000)
    check_replay(d, tmp_path, mappable=300000000)


@pytest.mark.parametrize("bw", [10, 100, 200])
def test_replay_bandwidths(tmp_path, bw):
    d = make_synthetic([400000, 100000], 2, 40000, directional=True, seed=7, peak_spacing=8000, bandwidth=200)
    check_replay(d, tmp_path, bandwidth=bw)


def test_replay_long_region_q8(tmp_path):
    rng = np.random.default_rng(13)
    sizes = [100000, 5000]
    per = [[None, None], [None, None]]
    for dd in range(2):
        for c, size in enumerate(sizes):
            pos = np.arange(51, size - 50, dtype=np.uint32)
            pos = pos[rng.random(pos.shape[0]) < 0.9]
            per[dd][c] = (pos, rng.integers(1, 4, size=pos.shape[0]).astype(np.uint32))
    d = TagData.from_segments(sizes, [per])
    ours = check_replay(d, tmp_path, mappable=100 * d.genome_size())
    assert max(r["len"] for r in ours["rows"] + ours["dropped"]) > 65536
    check_replay(d, tmp_path, mappable=100 * d.genome_size(), directional=False, corr_threshold=-0.5)


def test_replay_reference_assert_is_status_1(tmp_path):
    # several directional samples on a one-contig table: the reference's own merge loop trips its assert
    d = make_synthetic([500000], 3, 20000, directional=True, seed=3, peak_spacing=5000)
    st = region_settings(d)
    assert oracle_regions(d, st)["status"] == 1
    with pytest.raises(RuntimeError):
        run_ref_replay_regions(d, st, str(tmp_path))


def test_replay_strand_shift(tmp_path):
    d = make_synthetic([3000000], 1, 300000, directional=False, seed=11, peak_spacing=10000)
    kw = dict(n_regions_to_test=30, region_threshold=8.0, corr_threshold=0.1)
    ours = oracle_strand_shift(d, **kw)
    ref = run_ref_replay_shift(d, str(tmp_path), **kw)
    for k in ("n_regions_found", "n_regions_tested", "n_tags_in_regions", "best_shift", "shift_freq"):
        assert ours[k] == ref[k], k
    for g, w in zip(ours["tested"], ref["tested"]):
        assert (g["contig"], g["left"], g["len"], g["best_shift"]) == (w["contig"], w["left"], w["len"], w["best_shift"])
        assert g["best_corr"] == w["best_corr"]


# ---- the reference's real binaries ----------------------------------------------------------------------
def test_cli_regions_text(tmp_path):
    d = make_synthetic([200000, 100000, 50000], 3, 20000, directional=True, seed=1, peak_spacing=5000,
                       contig_names=["chrA", "chrB", "chrC"])
    files = write_inputs(d, str(tmp_path))
    out = str(tmp_path / "regions.txt")
    run_ref_cli("regions", ["-q", "-f", "-o", out, "-c", files["contigs"]] + files["wigs"])
    ours = oracle_regions(d, region_settings(d))
    want = [format_region_row(r, d.names(), peaks=True) for r in ours["rows"]]
    assert read_table_rows(out) == want
    # header lines the drop-in CLI must reproduce
    head = [l.rstrip("\n") for l in open(out) if l.startswith("#")]
    assert any(l.startswith("# background=") for l in head) and any(l.startswith("# hit_threshold=30") for l in head)


def test_cli_regions_nondirectional_text(tmp_path):
    d = make_synthetic([200000, 100000], 2, 30000, directional=False, seed=2, peak_spacing=5000, contig_names=["chrA", "chrB"])
    files = write_inputs(d, str(tmp_path))
    out = str(tmp_path / "regions.txt")
    run_ref_cli("regions", ["-q", "-D", "-y", "-f", "-o", out, "-c", files["contigs"]] + files["wigs"])
    ours = oracle_regions(d, region_settings(d, directional=False))
    want = [format_region_row(r, d.names(), peaks=True, corrs=True) for r in ours["rows"]]
    assert read_table_rows(out) == want


def test_cli_strand_shift_text(tmp_path):
    d = make_synthetic([3000000], 1, 300000, directional=False, seed=11, peak_spacing=10000, contig_names=["chrA"])
    files = write_inputs(d, str(tmp_path))
    out = str(tmp_path / "shift.txt")
    run_ref_cli("strand_shift", ["-o", out, "-c", files["contigs"], "-g", "30", "-r", "8", "-u", "0.1"] + files["wigs"])
    ours = oracle_strand_shift(d, n_regions_to_test=30, region_threshold=8.0, corr_threshold=0.1)
    text = open(out).read()
    m = re.search(r"# best_shift=(\d+)", text)
    assert m and int(m.group(1)) == ours["best_shift"]
    hist = {int(a): int(b) for a, b in re.findall(r"^(\d+)\t(\d+)$", text, flags=re.M)}
    assert hist == {i: f for i, f in enumerate(ours["shift_freq"]) if f}


def test_cli_tags_in_regions_text(tmp_path):
    d = make_synthetic([200000, 100000], 3, 20000, directional=True, seed=31, peak_spacing=5000, contig_names=["chrA", "chrB"])
    base = d.subset_samples([0])
    extra = d.subset_samples([1, 2])
    f0 = write_inputs(base, str(tmp_path / "base"))
    f1 = write_inputs(extra, str(tmp_path / "extra"), prefix="x")
    table = str(tmp_path / "regions.txt")
    run_ref_cli("regions", ["-q", "-o", table, "-c", f0["contigs"]] + f0["wigs"])
    out = str(tmp_path / "more.txt")
    run_ref_cli("tags_in_regions", ["-f", table, "-o", out, "-c", f0["contigs"]] + f1["wigs"])
    names = d.names()
    keys = []
    for row in read_table_rows(table):
        m = re.match(r"(\w+):(\d+)-(\d+)", row)
        keys.append((names.index(m.group(1)), int(m.group(2)), int(m.group(3))))
    status, counts = oracle_tags_in_regions(extra, keys, directional=True, extension=0)
    assert status == 0
    got = [[int(x) for x in row.split("\t")[-2:]] for row in read_table_rows(out)]
    assert got == counts.tolist()
