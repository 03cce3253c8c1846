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


# ---- the variable-length byte form (upb_pack_tags_vbyte / upb_load_tags_vbyte, unipeak_b200/csrc/upb_vbyte.hpp) ----------------------
def decode_vbyte(n_samples, n_contigs, off, vb, byte_off, tile_off, exc):
    """Plain sequential reading of the byte code: first byte 1 c v5..v0, 0-4 value bytes (7 bits each, low to high), then (c) the count."""
    E = int(off[-1])
    pos, cnt = np.zeros(E, np.uint32), np.zeros(E, np.uint32)
    xi = {int(e["index"]): (int(e["pos"]), int(e["cnt"])) for e in exc}
    tile = 0
    for sd in range(n_samples * 2):
        for c in range(n_contigs):
            k = sd * (n_contigs + 1) + c
            a, b = int(off[k]), int(off[k + 1])
            at, end = int(byte_off[k]), int(byte_off[k + 1])
            run = 0
            for e in range(a, b):
                head = (e - a) % 2048 == 0
                if head:
                    assert int(tile_off[tile]) == at
                    tile += 1
                first = int(vb[at])
                assert first & 0x80
                n = 1
                while at + n < end and not (int(vb[at + n]) & 0x80):
                    n += 1
                cont = [int(x) for x in vb[at + 1:at + n]]
                has_count = bool(first & 0x40)
                kk = 1
                if has_count:
                    kk = cont.pop()
                    assert kk != 1
                assert len(cont) <= 4 and (not cont or cont[-1] != 0)
                v = first & 63
                for j, x in enumerate(cont):
                    v |= x << (6 + 7 * j)
                run = v if head else run + v
                pos[e] = run
                if kk == 0:
                    assert xi[e][0] == 0
                    kk = xi[e][1]
                cnt[e] = kk
                at += n
            assert at == end
    assert tile == tile_off.shape[0] - 1 and int(tile_off[-1]) == int(byte_off[-1])
    return pos, cnt


def test_vbyte_roundtrip_and_size():
    d = make_synthetic([400000, 90000], 2, 6000, directional=True, seed=3, peak_spacing=3000)
    segs = [[[d.seg(s, dd, c) for c in range(2)] for dd in range(2)] for s in range(2)]
    # value boundaries of the code (63 | 64, 8191 | 8192, 2^20 - 1 | 2^20 ...) and count boundaries (1, 2, 127, 128, 70000)
    p = np.cumsum(np.array([5, 63, 64, 8191, 8192, 100000, 1, 1, 1], np.int64)).astype(np.uint32)
    segs[0][0][0] = (p, np.array([1, 2, 127, 128, 70000, 1, 3, 1, 126], np.uint32))
    data = TagData.from_segments(d.contig_sizes, segs)
    vb, boff, toff, exc = unipeak_b200.pack_tags_vbyte(2, 2, data.off, data.pos, data.cnt)
    E = int(data.pos.shape[0])
    assert vb.dtype == np.uint8 and boff.shape[0] == data.off.shape[0]
    assert np.all(np.diff(exc["index"].astype(np.int64)) > 0) and np.all(exc["pos"] == 0)
    assert set(int(i) for i in exc["index"]) >= {3, 4}                           # 128 and 70000 do not fit the count byte
    # the nine planted entries byte by byte: value 5 -> 1 byte; +63 -> 1 (count byte: +1); +64 -> 2; +8191 -> 2; +8192 -> 3; ...
    sizes = [1, 1 + 1, 2 + 1, 2 + 1, 3 + 1, 3, 1 + 1, 1, 1 + 1]
    assert int(boff[1]) - int(boff[0]) == sum(sizes)
    assert int(vb[0]) == 0x80 | 5 and int(vb[1]) == (0x80 | 0x40 | 63) and int(vb[2]) == 2
    pos, cnt = decode_vbyte(2, 2, data.off, vb, boff, toff, exc)
    assert np.array_equal(pos, data.pos) and np.array_equal(cnt, data.cnt)
    assert max(int(data.off[k + 1] - data.off[k]) for k in range(len(data.off) - 1)) > 2048
    assert vb.shape[0] + 16 * exc.shape[0] + 8 * toff.shape[0] < 2 * E          # denser than the 2-byte form on this data


def test_vbyte_sizes_first_then_buffers_and_empty_input():
    import ctypes as C
    L = unipeak_b200.lib()
    d = make_synthetic([5000], 1, 300, directional=True, seed=11)
    nb, nt, nx = C.c_uint64(0), C.c_uint64(0), C.c_uint64(0)
    off = np.ascontiguousarray(d.off, np.uint64)
    rc = L.upb_pack_tags_vbyte(1, 1, off.ctypes.data, d.pos.ctypes.data, d.cnt.ctypes.data, None, 0, C.byref(nb), None, None, 0, C.byref(nt), None, 0, C.byref(nx))
    assert rc != 0 and nb.value >= d.pos.shape[0] and nt.value == 2          # one tile per non-empty segment: '+' and '-'
    empty = TagData.from_segments([5000], [[[(np.zeros(0, np.uint32), np.zeros(0, np.uint32))], [(np.zeros(0, np.uint32), np.zeros(0, np.uint32))]]])
    vb, boff, toff, exc = unipeak_b200.pack_tags_vbyte(1, 1, empty.off, empty.pos, empty.cnt)
    assert vb.shape[0] == 0 and toff.tolist() == [0] and exc.shape[0] == 0 and not boff.any()


def test_vbyte_device_decoder_logic_on_the_host(tmp_path):
    """The tile stage of the device decoder calls the functions of upb_vbyte.hpp; tests/native/vbyte_tile_emul.cpp runs the same functions
    for 256 'threads' in turn on random tiles (entries of every length, entries that straddle the threads' byte ranges, malformed codes)."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "vbyte_tile_emul")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-o", exe, os.path.join(root, "tests", "native", "vbyte_tile_emul.cpp")])
    for seed in (1, 2, 3):
        out = subprocess.run([exe, str(seed), "120"], capture_output=True, text=True)
        assert out.returncode == 0 and out.stdout.startswith("ok"), out.stdout + out.stderr


def test_vbyte_encoder_into_a_caller_buffer():
    # with `out` the binding encodes in one call (tile count from `off`, exception list sized by guess, repeated when the guess was short)
    d = make_synthetic([900000, 200000], 3, 40000, directional=True, seed=21, peak_spacing=3000)
    want = unipeak_b200.pack_tags_vbyte(3, 2, d.off, d.pos, d.cnt)
    got = unipeak_b200.pack_tags_vbyte(3, 2, d.off, d.pos, d.cnt, out=np.empty(3 * d.pos.shape[0], np.uint8))
    assert all(np.array_equal(a, b) for a, b in zip(want, got))
    big = d.cnt.copy()
    big[:] = 500                                                               # every count escapes: far more exceptions than the guess
    want = unipeak_b200.pack_tags_vbyte(3, 2, d.off, d.pos, big)
    got = unipeak_b200.pack_tags_vbyte(3, 2, d.off, d.pos, big, out=np.empty(6 * d.pos.shape[0], np.uint8))
    assert got[3].shape[0] == d.pos.shape[0] and all(np.array_equal(a, b) for a, b in zip(want, got))
    import pytest
    with pytest.raises(ValueError):
        unipeak_b200.pack_tags_vbyte(3, 2, d.off, d.pos, d.cnt, out=np.empty(d.pos.shape[0] // 2, np.uint8))
