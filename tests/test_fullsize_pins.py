"""Full-size parity pins (tests/golden/fullsize_*.json, written by tools/pin_fullsize.py in the CPU container): the compiled,
unmodified reference was run over WHOLE BASELINE configurations — config 1 (1 x 1 M tags, chr21), config 2 (2 x 20 M tags, hg19:
strand_shift per sample + regions -D) and config 3 (16 x 30 M tags, hg19, 179 574 rows) — and every column of what it produced
was fingerprinted (SHA-256 over the raw values; doubles as bit patterns).  The inputs come from the deterministic generator,
which yields the same arrays on the CPU and on the GPU, so the B200 box regenerates them, proves it with the input fingerprint,
runs the CUDA path and compares every column: bit-exact parity at full size, not on a slice.
CPU side (no GPU): the generator is pinned, and the C restatement (oracle/) reproduces the reference's config-1 table."""
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))

from pin_fullsize import CONFIGS, canon_bits, column_hashes, region_settings  # noqa: E402
from unipeak_b200.synth_det import fingerprint, make_synthetic_det  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pin(name):
    with open(os.path.join(GOLDEN, "fullsize_%s.json" % name)) as f:
        return json.load(f)


def make(name, scale, device):
    n, tags, directional, sizes, mappable, seed = CONFIGS[name]
    data = make_synthetic_det(sizes, n, int(round(tags * scale)), directional=directional, seed=seed, bandwidth=50, device=device)
    st = region_settings([data.tags(s) for s in range(n)], n, directional, mappable, int(sum(sizes)))
    return data, st


def test_generator_is_pinned():
    d = make_synthetic_det([5_000_000, 3_000_000, 1_000_000], 4, 200_000, seed=7)
    assert fingerprint(d) .startswith("da60e35daf213f1f")
    d2 = make_synthetic_det([5_000_000, 3_000_000, 1_000_000], 4, 200_000, seed=7, contigs=[1])
    for s in range(4):
        for dd in range(2):
            assert np.array_equal(d.seg(s, dd, 1)[0], d2.seg(s, dd, 1)[0]) and d2.seg(s, dd, 0)[0].shape[0] == 0
    dn = make_synthetic_det([5_000_000, 3_000_000], 2, 300_000, directional=False, seed=9)
    assert fingerprint(dn).startswith("b3b22273752a0fa0")


def test_oracle_reproduces_the_reference_on_config1():
    from refutil import oracle_regions
    p = pin("c1")
    data, st = make("c1", 1.0, "cpu")
    assert fingerprint(data) == p["input_sha256"]
    got = oracle_regions(data, st)
    rows = got["rows"] + got["dropped"]
    cols = {k: np.asarray([r[k] for r in rows]) for k in ("forward_label", "contig", "left", "len", "peak", "peak_score", "kurtosis", "corr")}
    cols["counts"] = np.asarray([r["counts"] for r in rows], dtype=np.uint32).reshape(len(rows), data.n_samples)
    h = column_hashes(cols)
    want = p["regions"]
    assert len(got["rows"]) == want["n_rows"] and len(got["dropped"]) == want["n_dropped"]
    assert got["n_passed"] == want["n_passed"] and got["n_rejected"] == want["n_rejected"] and got["tags_in_regions"] == want["tags_in_regions"]
    for k, v in want["columns"].items():
        if k != "corr":                                                # directional: the reference never computes it; ref_replay asks anyway
            assert h[k] == v, k


def _gpu_columns(view):
    c = view.columns()
    keep = np.concatenate([np.nonzero(c["dropped"] == 0)[0], np.nonzero(c["dropped"] != 0)[0]])   # the pin lists the written rows, then the dropped ones
    return {k: np.ascontiguousarray(np.asarray(c[k])[keep]) for k in ("forward_label", "contig", "left", "len", "peak", "peak_score", "kurtosis", "corr", "counts")}, c


@pytest.mark.gpu
@pytest.mark.parametrize("name,scale", [("c1", 1.0), ("c3_s0.05", 0.05), ("c2_s0.05", 0.05), ("c2", 1.0), ("c3", 1.0)])
def test_gpu_matches_the_reference_at_full_size(name, scale):
    import torch
    import unipeak_b200
    p = pin(name)
    cfg = name.split("_")[0]
    if scale == 1.0 and cfg != "c1" and torch.cuda.mem_get_info()[0] < 40e9:
        pytest.skip("needs ~40 GB of free HBM")
    data, st = make(cfg, scale, "cuda:0")
    assert fingerprint(data) == p["input_sha256"], "the GPU box generated other inputs than the CPU container"
    ctx = unipeak_b200.Context(0)
    try:
        ctx.load_packed2(data)
        v = ctx.call_regions_view(**dict(st, reference_order=True))
        cols, raw = _gpu_columns(v)
        want = p["regions"]
        assert v.n_rows == want["n_rows"] + want["n_dropped"] and int((raw["dropped"] != 0).sum()) == want["n_dropped"]
        assert v.n_passed == want["n_passed"] and v.n_rejected == want["n_rejected"]
        assert [int(x) for x in raw["tags_in_regions"]] == want["tags_in_regions"]
        assert np.bincount(cols["contig"], minlength=len(want["rows_per_contig"])).tolist() == want["rows_per_contig"]
        h = column_hashes(cols)
        for k, val in want["columns"].items():
            if k == "corr" and st["directional"]:
                continue
            assert h[k] == val, "column %s differs from the reference's" % k
        ctx.release_view()
        for ss in p.get("strand_shift", []):                             # config 2: the calibration of every sample on its own
            import hashlib
            one = data.subset_samples([ss["sample"]])
            ctx.load(one)
            total = sum(one.tags(s) for s in range(one.n_samples))
            got = ctx.strand_shift(background=float(total) / float(one.genome_size()))
            assert got["status"] == 0
            for k in ("n_regions_found", "n_regions_tested", "n_tags_in_regions", "best_shift", "shift_freq"):
                assert got[k] == ss[k], (ss["sample"], k)
            t = got["tested"]
            sha = hashlib.sha256(b"".join(np.asarray([x[k] for x in t], dtype=dt).tobytes() for k, dt in
                                          (("contig", "<u4"), ("left", "<u4"), ("len", "<u4"), ("best_shift", "<u4"))) +
                                 canon_bits(np.asarray([x["best_corr"] for x in t])).tobytes()).hexdigest()
            assert sha == ss["tested_sha256"], ss["sample"]
    finally:
        ctx.close()
