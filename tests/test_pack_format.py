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
