"""Randomised parity: seeded random workloads x random settings through the C ABI against the CPU oracle, bit for bit (regions,
labels, counts matrix, every double, filter counters; the -w profile and tags_in_regions on a subset).  The fixed cases of
test_gpu_parity.py cover the quirks one at a time; this sweeps their combinations — contig tables with tiny and empty contigs,
1 to 6 samples, controls, -z coefficients, directional / -D, bandwidths 5 .. 300, loose and tight thresholds, tags at the contig
starts (left-edge replay) and contigs whose every tag lies within the bandwidth of the start (inherited window).
The seeds are fixed: a failure names the case that reproduces it."""
import os

import numpy as np
import pytest

from refutil import oracle_regions, oracle_tags_in_regions, region_settings
from test_gpu_parity import assert_rows_equal
from unipeak_b200.synth import TagData, make_synthetic

pytestmark = pytest.mark.gpu


def _case(seed):
    rng = np.random.default_rng(1000 + seed)
    n_contigs = int(rng.integers(1, 6))
    sizes = [int(rng.choice([700, 3000, 20000, 90000, 250000])) for _ in range(n_contigs)]
    n = int(rng.integers(1, 7))
    directional = bool(rng.integers(0, 2))
    bw = int(rng.choice([5, 10, 25, 50, 50, 80, 120, 300]))
    tags = int(rng.choice([300, 3000, 30000]))
    d = make_synthetic(sizes, n, tags, directional=directional, seed=int(rng.integers(1, 10 ** 6)), peak_spacing=int(rng.choice([500, 2000, 8000])),
                       bandwidth=bw, keep_clear_of_edges=bool(rng.integers(0, 3) == 0))
    # now and then: empty a contig for some samples, or keep only the tags within the bandwidth of a contig's start
    segs = [[[d.seg(s, dd, c) for c in range(n_contigs)] for dd in range(2)] for s in range(n)]
    for c in range(n_contigs):
        mode = int(rng.integers(0, 8))
        for s in range(n):
            for dd in range(2):
                p, k = segs[s][dd][c]
                if mode == 0 and rng.integers(0, 2):
                    segs[s][dd][c] = (p[:0], k[:0])
                elif mode == 1:
                    keep = p <= bw
                    if not keep.any() and rng.integers(0, 2):          # plant a few early tags
                        q = np.sort(rng.choice(np.arange(1, min(bw, sizes[c]) + 1), size=min(5, min(bw, sizes[c])), replace=False)).astype(np.uint32)
                        segs[s][dd][c] = (q, rng.integers(1, 9, size=q.shape[0]).astype(np.uint32))
                    else:
                        segs[s][dd][c] = (p[keep], k[keep])
    d = TagData.from_segments(sizes, segs)
    kw = dict(directional=directional, bandwidth=bw, region_threshold=float(rng.choice([2.0, 5.0, 25.0])), hit_threshold=float(rng.choice([1.0, 3.0, 10.0])),
              kurtosis_threshold=float(rng.choice([0.0, 3.0, 50.0])), mappable=int(rng.choice([1, 10, 200])) * d.genome_size())
    if not directional:
        kw["corr_threshold"] = float(rng.choice([-1.0, -0.5, 0.2]))
    if n > 1 and rng.integers(0, 3) == 0:
        kw["control"] = sorted(set(int(x) for x in rng.choice(np.arange(n), size=int(rng.integers(1, n)), replace=False)))
    n_src = n - len(kw.get("control", []))
    if rng.integers(0, 4) == 0 and n_src > 0:
        kw["coeffs"] = [float(x) for x in np.round(rng.uniform(0.25, 3.0, size=n_src), 3)]
    return d, kw


def _pass1_reads_a_minus_track(d, st):
    """src/regions.cpp:341 consumes whatever sits at the head of a sample's stream for the current contig, whatever its strand: the '-'
    track of a sample is reached in pass 1 when it starts at or after the last contig of the sample's '+' track."""
    if not st["directional"]:
        return False
    with_tags, early = 0, False
    for s in range(d.n_samples):
        plus = [c for c in range(d.n_contigs) if d.seg(s, 0, c)[0].shape[0]]
        minus = [c for c in range(d.n_contigs) if d.seg(s, 1, c)[0].shape[0]]
        if plus or minus:
            with_tags += 1
        if minus and (not plus or minus[0] >= plus[-1]):
            early = True
    return early and with_tags > 1


@pytest.mark.parametrize("seed", list(range(int(os.environ.get("UPB_FUZZ_SEEDS", "96")))))
def test_random_workload_matches_the_oracle(gpu_ctx, seed):
    d, kw = _case(seed)
    if sum(d.tags(s) for s in range(d.n_samples) if s not in kw.get("control", [])) == 0:
        pytest.skip("no non-control tags left (the reference divides by their number)")
    st = region_settings(d, **kw)
    want = oracle_regions(d, st, want_profile=True)
    if want["status"] != 0:                                             # the reference itself aborts on this input (Q7): the library must say so
        gpu_ctx.load(d)
        with pytest.raises(Exception):
            gpu_ctx.call_regions(**st)
        return
    # the same arrays through the three host forms in turn: plain, the variable-length byte form (what the drivers load), the 2-byte form
    (gpu_ctx.load, gpu_ctx.load_vbyte, gpu_ctx.load_packed2)[seed % 3](d)
    what = "seed %d %r" % (seed, {k: v for k, v in kw.items() if k != "mappable"})
    if _pass1_reads_a_minus_track(d, st):
        # several directional samples and a '-' track the reference's merge loop reads in pass 1: it aborts or, as here, pools that sample's
        # '-' tags apart from the others' (SURVEY Q7).  The library refuses the emission-order emulation by name and still returns the
        # canonical table — which holds what the reference would have found had it pooled all samples
        from unipeak_b200.capi import UpbError
        with pytest.raises(UpbError) as e:
            gpu_ctx.call_regions(**st)
        assert e.value.code == -7, what
        assert gpu_ctx.call_regions(**dict(st, reference_order=False))["n_passed"] >= 0
        return
    got = gpu_ctx.call_regions(**st)
    assert_rows_equal(got["rows"], want["rows"], what + " written")
    assert_rows_equal(got["dropped"], want["dropped"], what + " dropped")
    assert (got["n_passed"], got["n_rejected"], got["tags_in_regions"]) == (want["n_passed"], want["n_rejected"], want["tags_in_regions"]), what
    if seed % 3 == 0 and not kw.get("coeffs"):                          # the -w profile of every (buffer, contig)
        pts = []
        for b in range(2 if st["directional"] else 1):
            for c in range(d.n_contigs):
                pos, score = gpu_ctx.profile_contig(b, c, **st)
                pts += [(b, c, int(p), float(s)) for p, s in zip(pos, score)]
        key = lambda t: (t[0], t[1], t[2])
        assert sorted(pts, key=key) == sorted(want["profile"], key=key), what + " profile"
    if seed % 4 == 0 and want["rows"]:                                  # tags_in_regions of the same samples over the table just written
        keys = [(r["contig"], r["left"], r["left"] + r["len"] - 1) if r["forward_label"] else (r["contig"], r["left"] + r["len"] - 1, r["left"]) for r in want["rows"]]
        status, counts = oracle_tags_in_regions(d, keys, directional=st["directional"], extension=0)
        if status == 0:
            assert np.array_equal(gpu_ctx.count_tags_in_regions(keys, directional=st["directional"], extension=0), counts), what + " tags_in_regions"


def _shift_case(seed):
    rng = np.random.default_rng(5000 + seed)
    n_contigs = int(rng.integers(1, 5))
    sizes = [int(rng.choice([20000, 150000, 600000, 1500000])) for _ in range(n_contigs)]
    tags = int(rng.choice([20000, 100000, 300000]))
    d = make_synthetic(sizes, 1, tags, directional=False, seed=int(rng.integers(1, 10 ** 6)), peak_spacing=int(rng.choice([3000, 8000, 20000])),
                       keep_clear_of_edges=bool(rng.integers(0, 2)))
    kw = dict(n_regions_to_test=int(rng.choice([5, 40, 300, 70000])), region_threshold=float(rng.choice([4.0, 8.0, 25.0])), corr_threshold=float(rng.choice([-1.0, 0.1, 0.3])),
              hit_threshold=int(rng.choice([1, 10])), kurtosis_threshold=float(rng.choice([0.0, 50.0])), bandwidth=int(rng.choice([25, 50, 100])),
              min_shift=int(rng.choice([0, 25])), max_shift=int(rng.choice([60, 150])))
    return d, kw


@pytest.mark.parametrize("seed", list(range(int(os.environ.get("UPB_FUZZ_SHIFT_SEEDS", "32")))))
def test_random_strand_shift_matches_the_oracle(gpu_ctx, seed):
    from refutil import oracle_strand_shift
    d, kw = _shift_case(seed)
    want = oracle_strand_shift(d, **kw)
    (gpu_ctx.load_vbyte if seed % 2 else gpu_ctx.load)(d)
    got = gpu_ctx.strand_shift(background=want["background"], **kw)
    what = "seed %d %r" % (seed, kw)
    assert (got["status"] != 0) == (want["status"] != 0), what
    for k in ("n_regions_found", "n_regions_tested", "n_tags_in_regions", "best_shift", "shift_freq"):
        if want["status"] == 0 or k == "n_regions_found":
            assert got[k] == want[k], what + " " + k
    for g, w in zip(got["tested"], want["tested"]):
        assert (g["contig"], g["left"], g["len"], g["best_shift"]) == (w["contig"], w["left"], w["len"], w["best_shift"]), what
        assert g["best_corr"] == w["best_corr"] or (g["best_corr"] != g["best_corr"] and w["best_corr"] != w["best_corr"]), what
