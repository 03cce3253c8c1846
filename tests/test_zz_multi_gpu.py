"""upb_multi (include/unipeak_b200.h): the GPUs of one box behind one handle — contigs spread over the GPUs for regions and
strand_shift, samples for tags_in_regions — against the one-context entry points, which test_gpu_parity.py pins to the oracle.
Every comparison is bit for bit.  devices=[0, 0] puts both ranks on one GPU (runs anywhere); [0, 1] needs two GPUs."""
import math
import os

import numpy as np
import pytest

import refutil
from refutil import region_settings
from unipeak_b200.capi import MultiContext, lib
from unipeak_b200.synth import make_synthetic

pytestmark = pytest.mark.gpu


def _devices():
    out = [pytest.param([0, 0], id="two-ranks-one-gpu"), pytest.param([0, 0, 0], id="three-ranks-one-gpu")]
    try:
        n = lib().upb_device_count()
    except Exception:
        n = 0
    out.append(pytest.param([0, 1], id="two-gpus", marks=pytest.mark.skipif(n < 2, reason="needs two GPUs")))
    return out


def _same_rows(a, b):
    assert len(a) == len(b)
    for x, y in zip(a, b):
        for k in ("forward_label", "buffer", "contig", "left", "len", "peak", "counts"):
            assert x[k] == y[k], (k, x, y)
        for k in ("peak_score", "kurtosis", "corr"):
            assert x[k] == y[k] or (math.isnan(x[k]) and math.isnan(y[k])), (k, x, y)


def _check_regions(ctx, multi, data, packed=False, **kw):
    st = region_settings(data, **kw)
    ctx.load(data)
    want = ctx.call_regions(**st)
    multi.load(data, packed=packed)
    for _ in range(2):                                                # the first call sizes the shared table, the second reuses it
        got = multi.call_regions(**st)
        _same_rows(got["rows"], want["rows"])
        _same_rows(got["dropped"], want["dropped"])
        for k in ("n_passed", "n_rejected", "tags_in_regions"):
            assert got[k] == want[k], k
        assert np.array_equal(got["unit_candidates"], want["unit_candidates"])
    return want


@pytest.mark.parametrize("devices", _devices())
def test_multi_regions_equal_one_context(gpu_ctx, devices):
    m = MultiContext(len(devices), devices)
    try:
        d = make_synthetic([400000, 250000, 150000, 90000, 60000], 3, 40000, directional=True, seed=41, peak_spacing=5000)
        m.set_contigs(d.contig_sizes)
        assert len(set(m.owner().tolist())) == min(len(devices), 5)
        assert len(_check_regions(gpu_ctx, m, d)["rows"]) > 50
        _check_regions(gpu_ctx, m, d, packed=True, control=[1], region_threshold=10.0)
        _check_regions(gpu_ctx, m, d, directional=False)
        # other settings on the same load: more rows than the table was sized for -> the table grows
        _check_regions(gpu_ctx, m, d, region_threshold=2.0, hit_threshold=1.0, kurtosis_threshold=0.0)
        # tags at the contig starts: the left-edge replay runs on the rank that owns the contig
        e = make_synthetic([60000, 30000], 2, 30000, directional=True, seed=21, peak_spacing=2000, keep_clear_of_edges=False)
        _check_regions(gpu_ctx, m, e)
    finally:
        m.close()


@pytest.mark.parametrize("devices", _devices())
def test_multi_strand_shift_equals_one_context(gpu_ctx, devices):
    m = MultiContext(len(devices), devices)
    try:
        kw = dict(n_regions_to_test=40, region_threshold=8.0, corr_threshold=0.1)
        for d in (make_synthetic([1500000, 900000, 600000, 300000], 1, 500000, directional=False, seed=21, peak_spacing=10000),
                  refutil.long_start_region_data()):
            bg = refutil.oracle_strand_shift(d, **kw)["background"]
            gpu_ctx.load(d)
            want = gpu_ctx.strand_shift(background=bg, **kw)
            m.load(d)
            got = m.strand_shift(background=bg, **kw)
            assert got == want
            assert got["n_regions_tested"] > 5
    finally:
        m.close()


@pytest.mark.parametrize("devices", _devices())
def test_multi_tags_in_regions_equals_one_context(gpu_ctx, devices):
    m = MultiContext(len(devices), devices)
    try:
        d = make_synthetic([300000, 200000, 100000], 5, 30000, directional=True, seed=33, peak_spacing=5000)
        st = region_settings(d)
        gpu_ctx.load(d)
        rows = gpu_ctx.call_regions(**st)["rows"]
        keys = [(r["contig"], r["left"], r["left"] + r["len"] - 1) if r["forward_label"] else (r["contig"], r["left"] + r["len"] - 1, r["left"]) for r in rows]
        assert len(keys) > 30
        for directional, ext in ((True, 0), (True, 25)):
            want = gpu_ctx.count_tags_in_regions(keys, directional=directional, extension=ext)
            m.load_by_sample(d)
            got = m.count_tags_in_regions(keys, directional=directional, extension=ext)
            assert np.array_equal(got, want)
    finally:
        m.close()


def test_multi_profile_comes_from_the_owner(gpu_ctx):
    m = MultiContext(2, [0, 0])
    try:
        d = make_synthetic([60000, 30000], 2, 30000, directional=True, seed=21, peak_spacing=2000, keep_clear_of_edges=False)
        st = region_settings(d)
        gpu_ctx.load(d)
        m.load(d)
        for b in range(2):
            for c in range(2):
                p0, s0 = gpu_ctx.profile_contig(b, c, **st)
                p1, s1 = m.profile_contig(b, c, **st)
                assert np.array_equal(p0, p1) and np.array_equal(s0, s1)
    finally:
        m.close()


@pytest.mark.parametrize("seed", list(range(int(os.environ.get("UPB_FUZZ_MULTI_SEEDS", "8")))))   # ~4 s each: ranks that share a GPU stall each other when scratch grows (upb_multi.cu); 24 were green in r02
def test_multi_random_workloads_equal_one_context(gpu_ctx, seed):
    # the randomised workloads of test_zz_fuzz_gpu.py (controls, -z, -D, left-edge tags, tiny and empty contigs), two and three ranks on one GPU
    from test_zz_fuzz_gpu import _case, _pass1_reads_a_minus_track
    from unipeak_b200.capi import UpbError
    d, kw = _case(seed)
    if sum(d.tags(s) for s in range(d.n_samples) if s not in kw.get("control", [])) == 0:
        pytest.skip("no non-control tags left")
    st = region_settings(d, **kw)
    if _pass1_reads_a_minus_track(d, st):
        st = dict(st, reference_order=False)                           # (the emission-order emulation refuses these by name, Q7)
    gpu_ctx.load(d)
    try:
        want = gpu_ctx.call_regions(**st)
    except UpbError as e:
        pytest.skip("one context refuses this input too: %s" % e)
    m = MultiContext(2 + seed % 2, [0] * (2 + seed % 2))
    try:
        m.load(d, packed=bool(seed & 2))
        try:
            got = m.call_regions(**st)
        except UpbError as e:
            assert "another rank" in str(e), e                         # a window inherited across ranks: refused, not guessed
            return
        _same_rows(got["rows"], want["rows"])
        _same_rows(got["dropped"], want["dropped"])
        for k in ("n_passed", "n_rejected", "tags_in_regions"):
            assert got[k] == want[k], k
    finally:
        m.close()
