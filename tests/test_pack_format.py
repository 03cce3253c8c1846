"""The compact host form of the tag arrays (include/unipeak_b200.h: upb_pack_tags / upb_load_tags_packed): the host-only
encoder against a numpy decoder — 16-bit position deltas inside a segment, 8-bit counts, index-sorted exception list."""
import numpy as np

import unipeak_b200
from unipeak_b200.synth import TagData, make_synthetic


def decode(n_samples, n_contigs, off, dpos, cnt8, exc):
    E = int(off[-1])
    pos = np.zeros(E, np.uint32)
    cnt = cnt8[:E].astype(np.uint32)
    xi = {int(e["index"]): (int(e["pos"]), int(e["cnt"])) for e in exc}
    for sd in range(n_samples * 2):
        for c in range(n_contigs):
            a, b = int(off[sd * (n_contigs + 1) + c]), int(off[sd * (n_contigs + 1) + c + 1])
            run = 0
            for e in range(a, b):
                if dpos[e] == 0xFFFF:
                    run = xi[e][0]
                else:
                    run += int(dpos[e])
                pos[e] = run
                if cnt8[e] == 0xFF:
                    cnt[e] = xi[e][1]
    return pos, cnt


def test_pack_roundtrip_and_size():
    d = make_synthetic([400000, 90000], 2, 6000, directional=True, seed=3, peak_spacing=3000)
    cnt = d.cnt.copy()
    cnt[::97] = 255
    cnt[5::211] = 70000
    data = TagData(d.contig_sizes, d.n_samples, d.off, d.pos, cnt)
    dpos, cnt8, exc = unipeak_b200.pack_tags(2, 2, data.off, data.pos, data.cnt)
    E = int(data.pos.shape[0])
    assert dpos.dtype == np.uint16 and cnt8.dtype == np.uint8 and dpos.shape[0] == E and cnt8.shape[0] == E
    assert np.all(np.diff(exc["index"].astype(np.int64)) > 0)                  # sorted by entry index
    starts = [int(data.off[sd * 3 + c]) for sd in range(4) for c in range(2) if data.off[sd * 3 + c + 1] > data.off[sd * 3 + c]]
    assert all(dpos[s] == 0xFFFF for s in starts)                              # the first entry of every segment is absolute
    pos, cnt2 = decode(2, 2, data.off, dpos, cnt8, exc)
    assert np.array_equal(pos, data.pos) and np.array_equal(cnt2, data.cnt)
    assert 3 * E + 16 * exc.shape[0] < 0.5 * 8 * E                             # well under half the plain form


def decode2(n_samples, n_contigs, off, words, exc):
    E = int(off[-1])
    pos = np.zeros(E, np.uint32)
    cnt = (words[:E] >> 12).astype(np.uint32)
    xi = {int(e["index"]): (int(e["pos"]), int(e["cnt"])) for e in exc}
    for sd in range(n_samples * 2):
        for c in range(n_contigs):
            a, b = int(off[sd * (n_contigs + 1) + c]), int(off[sd * (n_contigs + 1) + c + 1])
            run = 0
            for e in range(a, b):
                d = int(words[e]) & 0xFFF
                if d == 0xFFF:
                    run = xi[e][0]
                else:
                    run += d
                pos[e] = run
                if cnt[e] == 0xF:
                    cnt[e] = xi[e][1]
    return pos, cnt


def test_pack2_roundtrip_and_size():
    # the 2-byte form: 12-bit deltas, 4-bit counts; gaps of exactly 4094 / 4095 / 4096 and counts of 14 / 15 / 16 sit on the escape boundary
    d = make_synthetic([400000, 90000], 2, 6000, directional=True, seed=3, peak_spacing=3000)
    segs = [[[d.seg(s, dd, c) for c in range(2)] for dd in range(2)] for s in range(2)]
    segs[0][0][0] = (np.array([5, 5 + 4094, 5 + 4094 + 4095, 5 + 4094 + 4095 + 4096, 300000], np.uint32), np.array([14, 15, 16, 1, 70000], np.uint32))
    data = TagData.from_segments(d.contig_sizes, segs)
    words, exc = unipeak_b200.pack_tags2(2, 2, data.off, data.pos, data.cnt)
    E = int(data.pos.shape[0])
    assert words.dtype == np.uint16 and words.shape[0] == E
    assert np.all(np.diff(exc["index"].astype(np.int64)) > 0)
    assert (int(words[1]) & 0xFFF) == 4094 and (int(words[2]) & 0xFFF) == 0xFFF and (int(words[3]) & 0xFFF) == 0xFFF
    assert (int(words[0]) >> 12) == 14 and (int(words[1]) >> 12) == 0xF and (int(words[2]) >> 12) == 0xF and (int(words[3]) >> 12) == 1
    pos, cnt = decode2(2, 2, data.off, words, exc)
    assert np.array_equal(pos, data.pos) and np.array_equal(cnt, data.cnt)
    assert 2 * E + 16 * exc.shape[0] < 0.32 * 8 * E                            # under a third of the plain form on this data
