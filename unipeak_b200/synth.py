"""Seeded synthetic tag-frequency data in the shape SURVEY.md §8(d) describes, plus the flat
in-memory layout every layer of this repo exchanges (the C-ABI, the oracle, the replay harness).

Layout (`TagData`): for sample s, strand d (0 '+', 1 '-') and contig c the entries are
``pos[off[(s*2+d)*(C+1)+c] : off[(s*2+d)*(C+1)+c+1]]`` — strictly ascending 1-based positions with
non-zero counts.  This is what the reference's wiggle front end (misc/format.cpp:503-565) delivers
to its drivers one line at a time.
"""
from __future__ import annotations

import dataclasses
import struct
from typing import List, Sequence, Tuple

import numpy as np

HG19_SIZES = [
    249250621, 243199373, 198022430, 191154276, 180915260, 171115067, 159138663, 146364022,
    141213431, 135534747, 135006516, 133851895, 115169878, 107349540, 102531392, 90354753,
    81195210, 78077248, 59128983, 63025520, 48129895, 51304566, 155270560, 59373566, 16571,
]
HG19_NAMES = ["chr%s" % c for c in list(range(1, 23)) + ["X", "Y", "M"]]
# UCSC mm9 chromosome sizes (the reference ships the same table as contig/mm9.txt): 22 contigs, 2 654 911 517 bp
MM9_SIZES = [197195432, 181748087, 159599783, 155630120, 152537259, 149517037, 152524553, 131738871, 124076172, 129993255, 121843856,
             121257530, 120284312, 125194864, 103494974, 98319150, 95272651, 90772031, 61342430, 166650296, 15902555, 16299]
MM9_NAMES = ["chr%s" % c for c in list(range(1, 20)) + ["X", "Y", "M"]]


@dataclasses.dataclass
class TagData:
    contig_sizes: np.ndarray  # uint32 [C]
    n_samples: int
    off: np.ndarray           # uint64 [n_samples*2*(C+1)]
    pos: np.ndarray           # uint32 [E]
    cnt: np.ndarray           # uint32 [E]
    contig_names: List[str] = dataclasses.field(default_factory=list)

    @property
    def n_contigs(self) -> int:
        return int(self.contig_sizes.shape[0])

    def names(self) -> List[str]:
        return self.contig_names or ["c%d" % i for i in range(self.n_contigs)]

    def seg(self, s: int, d: int, c: int) -> Tuple[np.ndarray, np.ndarray]:
        base = (s * 2 + d) * (self.n_contigs + 1)
        a, b = int(self.off[base + c]), int(self.off[base + c + 1])
        return self.pos[a:b], self.cnt[a:b]

    def tags(self, s: int) -> int:
        base = s * 2 * (self.n_contigs + 1)
        a = int(self.off[base])
        b = int(self.off[base + 2 * (self.n_contigs + 1) - 1])
        return int(self.cnt[a:b].astype(np.uint64).sum())

    def genome_size(self) -> int:
        return int(self.contig_sizes.astype(np.uint64).sum())

    # ---- construction -------------------------------------------------------------------
    @staticmethod
    def from_segments(contig_sizes: Sequence[int], segments, contig_names=None) -> "TagData":
        """segments[s][d][c] = (pos array, cnt array), already sorted/unique."""
        C = len(contig_sizes)
        n = len(segments)
        off = np.zeros(n * 2 * (C + 1), dtype=np.uint64)
        pos_parts, cnt_parts = [], []
        total = 0
        for s in range(n):
            for d in range(2):
                base = (s * 2 + d) * (C + 1)
                for c in range(C):
                    p, k = segments[s][d][c]
                    p = np.asarray(p, dtype=np.uint32)
                    k = np.asarray(k, dtype=np.uint32)
                    assert p.shape == k.shape
                    off[base + c] = total
                    pos_parts.append(p)
                    cnt_parts.append(k)
                    total += p.shape[0]
                off[base + C] = total
        pos = np.concatenate(pos_parts) if pos_parts else np.zeros(0, np.uint32)
        cnt = np.concatenate(cnt_parts) if cnt_parts else np.zeros(0, np.uint32)
        return TagData(np.asarray(contig_sizes, dtype=np.uint32), n, off, pos.astype(np.uint32), cnt.astype(np.uint32),
                       list(contig_names) if contig_names else [])

    def subset_samples(self, samples: Sequence[int]) -> "TagData":
        segs = [[[self.seg(s, d, c) for c in range(self.n_contigs)] for d in range(2)] for s in samples]
        return TagData.from_segments(self.contig_sizes, segs, self.contig_names)

    def subset_contigs(self, contigs: Sequence[int]) -> "TagData":
        segs = [[[self.seg(s, d, c) for c in contigs] for d in range(2)] for s in range(self.n_samples)]
        names = [self.names()[c] for c in contigs]
        return TagData.from_segments([int(self.contig_sizes[c]) for c in contigs], segs, names)

    # ---- serialisation ------------------------------------------------------------------
    def write_upbi(self, path: str) -> None:
        """Binary container read by oracle/ref_replay.cpp."""
        with open(path, "wb") as f:
            f.write(b"UPBI0001")
            f.write(struct.pack("<II", self.n_contigs, self.n_samples))
            f.write(self.contig_sizes.astype("<u4").tobytes())
            f.write(self.off.astype("<u8").tobytes())
            f.write(struct.pack("<Q", int(self.pos.shape[0])))
            f.write(self.pos.astype("<u4").tobytes())
            f.write(self.cnt.astype("<u4").tobytes())

    def write_wig(self, path: str, sample: int, name: str, tags_header: int | None = None) -> None:
        """Canonical directional tag-frequency wiggle, as bin/convert_align writes it
        (misc/format.cpp:1042-1060, :1164-1219)."""
        names = self.names()
        with open(path, "w") as f:
            f.write("# tags=%d\n" % (self.tags(sample) if tags_header is None else tags_header))
            for d, sign in ((0, "+"), (1, "-")):
                if all(self.seg(sample, d, c)[0].shape[0] == 0 for c in range(self.n_contigs)):
                    continue
                if d == 0:
                    f.write('track name="%s +" description="%s" priority=3 visibility=full type=wiggle_0 '
                            'alwaysZero=on color=0,0,255\n' % (name, name))
                else:
                    f.write('track name="%s -" description=" " priority=3 visibility=full type=wiggle_0 '
                            'alwaysZero=on color=255,0,0 altColor=255,0,0\n' % name)
                for c in range(self.n_contigs):
                    p, k = self.seg(sample, d, c)
                    if p.shape[0] == 0:
                        continue
                    f.write("variableStep chrom=%s\n" % names[c])
                    neg = "-" if d == 1 else ""
                    f.write("".join("%d %s%d\n" % (a, neg, b) for a, b in zip(p.tolist(), k.tolist())))

    def write_contig_table(self, path: str) -> None:
        with open(path, "w") as f:
            for nm, sz in zip(self.names(), self.contig_sizes.tolist()):
                f.write("%s\t%d\n" % (nm, sz))


def _unique_counts(p: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    if p.shape[0] == 0:
        return np.zeros(0, np.uint32), np.zeros(0, np.uint32)
    u, k = np.unique(p, return_counts=True)
    return u.astype(np.uint32), k.astype(np.uint32)


def make_synthetic(contig_sizes: Sequence[int], n_samples: int, tags_per_sample: int, *, directional: bool = True,
                   seed: int = 1234, bandwidth: int = 50, peak_fraction: float = 0.7, peak_spacing: int = 10000,
                   contig_names: Sequence[str] | None = None, keep_clear_of_edges: bool = True) -> TagData:
    """70 % of each sample's tags fall in peaks shared by all samples (centres
    uniform, one per `peak_spacing` bp, Pareto(1.2)+1 weights), 30 % are uniform background.
    Directional (3SEQ / RNA-seq like): independent peak sets per strand, tags = round(centre + N(0,40)).
    Non-directional (ChIP like): fragment midpoints centre + N(0,30), fragment length N(180,20), '+' tag
    at mid - len/2 and '-' tag at mid + len/2 (true half-shift 90)."""
    sizes = np.asarray(contig_sizes, dtype=np.int64)
    C = len(sizes)
    G = int(sizes.sum())
    starts = np.concatenate([[0], np.cumsum(sizes)])
    prng = np.random.default_rng(seed)  # peaks are shared by all samples
    lo_margin = bandwidth + 1 if keep_clear_of_edges else 1

    def peak_set(k):
        npk = max(1, G // peak_spacing)
        g = prng.integers(0, G, size=npk)
        w = prng.pareto(1.2, size=npk) + 1.0
        return g, w / w.sum()

    peak_sets = [peak_set(0), peak_set(1)] if directional else [peak_set(0)]

    segs = []
    for s in range(n_samples):
        rng = np.random.default_rng(seed + 1 + s)
        n_peak = int(round(tags_per_sample * peak_fraction))
        n_bg = tags_per_sample - n_peak
        glob = [[], []]  # per strand: global coordinates (0-based across concatenated contigs)
        if directional:
            for d in range(2):
                m = n_peak // 2 if d == 0 else n_peak - n_peak // 2
                g, w = peak_sets[d]
                which = rng.choice(g.shape[0], size=m, p=w)
                glob[d].append(np.rint(g[which] + rng.normal(0, 40, size=m)).astype(np.int64))
        else:
            g, w = peak_sets[0]
            m = n_peak // 2
            which = rng.choice(g.shape[0], size=m, p=w)
            mid = g[which] + rng.normal(0, 30, size=m)
            ln = rng.normal(180, 20, size=m)
            glob[0].append(np.rint(mid - ln / 2).astype(np.int64))
            glob[1].append(np.rint(mid + ln / 2).astype(np.int64))
            n_bg = tags_per_sample - 2 * m
        for d in range(2):
            m = n_bg // 2 if d == 0 else n_bg - n_bg // 2
            glob[d].append(rng.integers(0, G, size=m))
        per = [[None] * C for _ in range(2)]
        for d in range(2):
            x = np.concatenate(glob[d])
            x = x[(x >= 0) & (x < G)]
            c = np.searchsorted(starts, x, side="right") - 1
            local = x - starts[c] + 1  # 1-based
            ok = (local >= lo_margin) & (local <= sizes[c] - (bandwidth if keep_clear_of_edges else 0))
            c, local = c[ok], local[ok]
            for ci in range(C):
                per[d][ci] = _unique_counts(local[c == ci])
        segs.append(per)
    return TagData.from_segments(sizes, segs, contig_names)
