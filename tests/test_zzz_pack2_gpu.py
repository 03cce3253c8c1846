"""The 2-byte host form (upb_pack_tags2 / upb_load_tags_packed2: 12-bit position delta + 4-bit count per entry) decodes on the device
to the same arrays as the plain load: same region table, same tags_in_regions counts (those read the decoded pos[] / cnt[] directly).
Added after the GPU budget of round 1 was spent; kept in a file of its own, last in collection order, until it has run once."""
import numpy as np
import pytest

import unipeak_b200
from refutil import region_settings
from test_gpu_parity import assert_rows_equal
from unipeak_b200.synth import TagData, make_synthetic

pytestmark = pytest.mark.gpu


def _data():
    rng = np.random.default_rng(8)
    sizes = [3000000, 500000, 200000]
    d = make_synthetic(sizes, 3, 40000, directional=True, seed=78, peak_spacing=6000)
    pos, cnt = d.pos.copy(), d.cnt.copy()
    cnt[rng.choice(pos.shape[0], size=60, replace=False)] = rng.integers(14, 18, size=60).astype(np.uint32)      # around the 4-bit escape
    cnt[rng.choice(pos.shape[0], size=20, replace=False)] = rng.integers(255, 100000, size=20).astype(np.uint32)
    segs = [[[None] * 3 for _ in range(2)] for _ in range(3)]
    for s in range(3):
        for dd in range(2):
            for c in range(3):
                a, b = int(d.off[(s * 2 + dd) * 4 + c]), int(d.off[(s * 2 + dd) * 4 + c + 1])
                p, k = pos[a:b], cnt[a:b]
                if c == 0:                                            # thin chr0 out: gaps far beyond 4095
                    keep = (p < 400000) | (p > 2500000)
                    p, k = p[keep], k[keep]
                if s == 1 and c == 2:
                    p, k = p[:0], k[:0]                               # an empty segment
                segs[s][dd][c] = (p, k)
    return TagData.from_segments(sizes, segs)


def test_load_packed2_equals_plain_load(gpu_ctx):
    data = _data()
    words, exc = unipeak_b200.pack_tags2(3, 3, data.off, data.pos, data.cnt)
    assert ((words & 0xFFF) == 0xFFF).sum() >= 16 + 2 and ((words >> 12) == 0xF).sum() >= 20
    st = region_settings(data)
    gpu_ctx.load(data)
    want = gpu_ctx.call_regions(**st)
    keys = [(r["contig"], r["left"], r["left"] + r["len"] - 1) for r in want["rows"] if r["forward_label"]][:50]
    tir_want = gpu_ctx.count_tags_in_regions(keys, directional=True, extension=0)
    gpu_ctx.load_packed2(data)
    got = gpu_ctx.call_regions(**st)
    assert_rows_equal(got["rows"], want["rows"], "2-byte load")
    assert_rows_equal(got["dropped"], want["dropped"], "2-byte load (dropped)")
    assert got["tags_in_regions"] == want["tags_in_regions"] and len(got["rows"]) > 3
    assert np.array_equal(gpu_ctx.count_tags_in_regions(keys, directional=True, extension=0), tir_want)
    # a broken stream is refused: an escape without its exception entry
    off = np.ascontiguousarray(data.off, dtype=np.uint64)
    with pytest.raises(unipeak_b200.UpbError):
        gpu_ctx.load_tags_packed2_ptr(3, off.ctypes.data, words.ctypes.data, exc.shape[0] - 1, exc.ctypes.data)
    # and the 3-byte form still decodes to the same table (the two share their decoder)
    gpu_ctx.load_packed(data)
    again = gpu_ctx.call_regions(**st)
    assert_rows_equal(again["rows"], want["rows"], "3-byte load")
