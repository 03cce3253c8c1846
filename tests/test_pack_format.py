"""The compact host form of the tag arrays (include/unipeak_b200.h: upb_pack_tags2 / upb_load_tags_packed2): the host-only
encoder against a numpy decoder — 12-bit position deltas inside a segment, 4-bit counts, index-sorted exception list, and an
absolute position at the start of every segment and every 2048th entry of it (self-contained tiles for the device decoder)."""
import numpy as np

import unipeak_b200
from unipeak_b200.synth import TagData, make_synthetic


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
    for sd in range(4):                                                        # tiles decode on their own: entry 0, 2048, 4096 ... of a segment is absolute
        for c in range(2):
            a, b = int(data.off[sd * 3 + c]), int(data.off[sd * 3 + c + 1])
            assert all((int(words[e]) & 0xFFF) == 0xFFF for e in range(a, b, 2048))
    assert max(int(data.off[k + 1] - data.off[k]) for k in range(len(data.off) - 1)) > 2048
    assert 2 * E + 16 * exc.shape[0] < 0.32 * 8 * E                            # under a third of the plain form on this data
