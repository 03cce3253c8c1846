"""Test-side helpers: ctypes binding of the CPU oracle (oracle/unipeak_oracle.c), runners for the
compiled reference (oracle/_ref, when present), and the regions-table text format.

Nothing here is imported by the product package.
"""
from __future__ import annotations

import ctypes as C
import math
import os
import struct
import subprocess
import sys
from typing import Dict, List, Optional, Sequence

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
ORACLE_DIR = os.path.join(ROOT, "oracle")
REF_DIR = os.path.join(ORACLE_DIR, "_ref")
ORACLE_SO = os.path.join(ORACLE_DIR, "_build", "libupb_oracle.so")

from unipeak_b200.synth import TagData  # noqa: E402


def have_ref() -> bool:
    return all(os.access(os.path.join(REF_DIR, b), os.X_OK) for b in ("regions", "strand_shift", "tags_in_regions", "ref_replay"))


# ------------------------------------------------------------------------------------------------
# oracle binding
# ------------------------------------------------------------------------------------------------
class UpoInput(C.Structure):
    _fields_ = [("n_contigs", C.c_uint32), ("contig_sizes", C.POINTER(C.c_uint32)), ("n_samples", C.c_uint32),
                ("off", C.POINTER(C.c_uint64)), ("pos", C.POINTER(C.c_uint32)), ("cnt", C.POINTER(C.c_uint32))]


class UpoParams(C.Structure):
    _fields_ = [("bandwidth", C.c_uint32), ("region_threshold", C.c_double), ("kurtosis_threshold", C.c_double),
                ("corr_threshold", C.c_double), ("hit_threshold", C.c_double), ("directional", C.c_int),
                ("control", C.POINTER(C.c_uint8)), ("n_coeffs", C.c_uint32), ("coeffs", C.POINTER(C.c_double)),
                ("background", C.c_double), ("want_profile", C.c_int)]


class UpoRegion(C.Structure):
    _fields_ = [("forward_label", C.c_uint8), ("buffer", C.c_uint8), ("contig", C.c_uint32), ("left", C.c_uint32),
                ("len", C.c_uint32), ("peak_pos", C.c_uint32), ("peak_score", C.c_double), ("kurtosis", C.c_double),
                ("corr", C.c_double)]


class UpoProfilePoint(C.Structure):
    _fields_ = [("buffer", C.c_uint8), ("contig", C.c_uint32), ("pos", C.c_uint32), ("score", C.c_double)]


class UpoRegionsResult(C.Structure):
    _fields_ = [("n_rows", C.c_size_t), ("rows", C.POINTER(UpoRegion)), ("counts", C.POINTER(C.c_uint32)),
                ("n_dropped", C.c_size_t), ("dropped", C.POINTER(UpoRegion)), ("dropped_counts", C.POINTER(C.c_uint32)),
                ("n_passed", C.c_uint64), ("n_rejected", C.c_uint64), ("tags_in_regions", C.POINTER(C.c_uint64)),
                ("n_profile", C.c_size_t), ("profile", C.POINTER(UpoProfilePoint)), ("status", C.c_int)]


class UpoShiftParams(C.Structure):
    _fields_ = [("bandwidth", C.c_uint32), ("region_threshold", C.c_double), ("kurtosis_threshold", C.c_double),
                ("corr_threshold", C.c_double), ("hit_threshold", C.c_uint32), ("min_shift", C.c_uint32),
                ("max_shift", C.c_uint32), ("n_regions_to_test", C.c_uint32), ("background", C.c_double)]


class UpoShiftResult(C.Structure):
    _fields_ = [("n_regions_found", C.c_uint64), ("n_regions_tested", C.c_uint32), ("n_tags_in_regions", C.c_uint64),
                ("shift_freq", C.POINTER(C.c_uint64)), ("best_shift", C.c_uint32), ("n_tested_rows", C.c_size_t),
                ("tested_contig", C.POINTER(C.c_uint32)), ("tested_left", C.POINTER(C.c_uint32)),
                ("tested_len", C.POINTER(C.c_uint32)), ("tested_best_shift", C.POINTER(C.c_uint32)),
                ("tested_best_corr", C.POINTER(C.c_double)), ("status", C.c_int)]


class UpoRegionKey(C.Structure):
    _fields_ = [("contig", C.c_uint32), ("start", C.c_uint32), ("end", C.c_uint32)]


_oracle = None


def build_oracle() -> None:
    subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "port"])


def oracle():
    global _oracle
    if _oracle is None:
        if not os.path.exists(ORACLE_SO) or os.path.getmtime(ORACLE_SO) < os.path.getmtime(os.path.join(ORACLE_DIR, "unipeak_oracle.c")):
            build_oracle()
        lib = C.CDLL(ORACLE_SO)
        lib.upo_kernel.argtypes = [C.c_uint32, C.c_double, C.POINTER(C.c_double)]
        lib.upo_kernel.restype = None
        lib.upo_regions.argtypes = [C.POINTER(UpoInput), C.POINTER(UpoParams), C.POINTER(UpoRegionsResult)]
        lib.upo_regions_free.argtypes = [C.POINTER(UpoRegionsResult)]
        lib.upo_strand_shift.argtypes = [C.POINTER(UpoInput), C.POINTER(UpoShiftParams), C.POINTER(UpoShiftResult)]
        lib.upo_shift_free.argtypes = [C.POINTER(UpoShiftResult)]
        lib.upo_tags_in_regions.argtypes = [C.POINTER(UpoInput), C.c_int, C.c_uint32, C.c_size_t, C.POINTER(UpoRegionKey), C.POINTER(C.c_uint32)]
        lib.upo_sort_by_sum_desc.argtypes = [C.c_size_t, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        lib.upo_sort_by_sum_desc.restype = None
        lib.upo_strerror.restype = C.c_char_p
        _oracle = lib
    return _oracle


def _as_input(data: TagData):
    sizes = np.ascontiguousarray(data.contig_sizes, dtype=np.uint32)
    off = np.ascontiguousarray(data.off, dtype=np.uint64)
    pos = np.ascontiguousarray(data.pos, dtype=np.uint32)
    cnt = np.ascontiguousarray(data.cnt, dtype=np.uint32)
    inp = UpoInput(data.n_contigs, sizes.ctypes.data_as(C.POINTER(C.c_uint32)), data.n_samples,
                   off.ctypes.data_as(C.POINTER(C.c_uint64)), pos.ctypes.data_as(C.POINTER(C.c_uint32)),
                   cnt.ctypes.data_as(C.POINTER(C.c_uint32)))
    return inp, (sizes, off, pos, cnt)


def region_settings(data: TagData, *, directional=True, bandwidth=50, region_threshold=25.0, kurtosis_threshold=50.0,
                    corr_threshold=0.3, hit_threshold=10.0, control: Sequence[int] = (), coeffs: Optional[Sequence[float]] = None,
                    mappable: int = 0, proportional=False) -> Dict:
    """Derive what src/regions.cpp:104-232 derives from the flags: scaled hit threshold, corr threshold,
    background, normalised coefficients.  `control` holds 0-based sample indices."""
    n = data.n_samples
    ctrl = np.zeros(n, dtype=np.uint8)
    for i in control:
        ctrl[i] = 1
    n_control = int(ctrl.sum())
    tags = [data.tags(s) for s in range(n)]
    n_noncontrol = sum(t for t, c in zip(tags, ctrl) if not c)
    mappable = mappable or data.genome_size()
    background = float(n_noncontrol) / float(mappable)
    if directional:
        background /= 2
    cf: List[float] = []
    if proportional:
        cf = [float(n_noncontrol) / (float(t) * float(n - n_control)) for t, c in zip(tags, ctrl) if not c]
    elif coeffs:
        cf = [float(x) for x in coeffs]
        scaled = 0.0
        nc_tags = [t for t, c in zip(tags, ctrl) if not c]
        for x, t in zip(cf, nc_tags):
            scaled += x * float(t)
        cf = [x * (n_noncontrol / scaled) for x in cf]   # regions.cpp:228: Counter / double
    return dict(bandwidth=bandwidth, region_threshold=float(region_threshold), kurtosis_threshold=float(kurtosis_threshold),
                corr_threshold=(-1.0 if directional else float(corr_threshold)),
                hit_threshold=float(hit_threshold) * (n - n_control), directional=bool(directional), control=ctrl,
                coeffs=cf, background=background, tags=tags, n_noncontrol=n_noncontrol, mappable=mappable)


def _rows_to_dicts(rows, counts, n_rows, n) -> List[Dict]:
    out = []
    for k in range(n_rows):
        r = rows[k]
        out.append(dict(forward_label=int(r.forward_label), buffer=int(r.buffer), contig=int(r.contig), left=int(r.left),
                        len=int(r.len), peak=int(r.peak_pos), peak_score=float(r.peak_score), kurtosis=float(r.kurtosis),
                        corr=float(r.corr), counts=[int(counts[k * n + e]) for e in range(n)]))
    return out


def oracle_regions(data: TagData, st: Dict, want_profile=False) -> Dict:
    lib = oracle()
    inp, keep = _as_input(data)
    ctrl = np.ascontiguousarray(st["control"], dtype=np.uint8)
    cf = np.ascontiguousarray(st["coeffs"], dtype=np.float64)
    p = UpoParams(st["bandwidth"], st["region_threshold"], st["kurtosis_threshold"], st["corr_threshold"], st["hit_threshold"],
                  1 if st["directional"] else 0, ctrl.ctypes.data_as(C.POINTER(C.c_uint8)), cf.shape[0],
                  cf.ctypes.data_as(C.POINTER(C.c_double)), st["background"], 1 if want_profile else 0)
    res = UpoRegionsResult()
    status = lib.upo_regions(C.byref(inp), C.byref(p), C.byref(res))
    n = data.n_samples
    out = dict(status=status, rows=_rows_to_dicts(res.rows, res.counts, res.n_rows, n),
               dropped=_rows_to_dicts(res.dropped, res.dropped_counts, res.n_dropped, n),
               n_passed=int(res.n_passed), n_rejected=int(res.n_rejected),
               tags_in_regions=[int(res.tags_in_regions[e]) for e in range(n)] if res.tags_in_regions else [])
    if want_profile:
        m = res.n_profile
        out["profile"] = [(int(res.profile[i].buffer), int(res.profile[i].contig), int(res.profile[i].pos), float(res.profile[i].score)) for i in range(m)]
    lib.upo_regions_free(C.byref(res))
    del keep
    return out


def oracle_strand_shift(data: TagData, *, bandwidth=50, region_threshold=25.0, kurtosis_threshold=50.0, corr_threshold=0.3,
                        hit_threshold=10, min_shift=25, max_shift=150, n_regions_to_test=1000, mappable=0) -> Dict:
    lib = oracle()
    inp, keep = _as_input(data)
    total = sum(data.tags(s) for s in range(data.n_samples))
    background = float(total) / float(mappable or data.genome_size())
    p = UpoShiftParams(bandwidth, region_threshold, kurtosis_threshold, corr_threshold, hit_threshold, min_shift, max_shift,
                       n_regions_to_test, background)
    res = UpoShiftResult()
    status = lib.upo_strand_shift(C.byref(inp), C.byref(p), C.byref(res))
    out = dict(status=status, n_regions_found=int(res.n_regions_found), n_regions_tested=int(res.n_regions_tested),
               n_tags_in_regions=int(res.n_tags_in_regions), best_shift=int(res.best_shift), background=background, total_tags=total,
               shift_freq=[int(res.shift_freq[i]) for i in range(max_shift + 1)] if res.shift_freq else [],
               tested=[dict(contig=int(res.tested_contig[i]), left=int(res.tested_left[i]), len=int(res.tested_len[i]),
                            best_shift=int(res.tested_best_shift[i]), best_corr=float(res.tested_best_corr[i]))
                       for i in range(res.n_tested_rows)])
    lib.upo_shift_free(C.byref(res))
    del keep
    return out


def oracle_tags_in_regions(data: TagData, keys: Sequence[Sequence[int]], *, directional=True, extension=0):
    lib = oracle()
    inp, keep = _as_input(data)
    arr = (UpoRegionKey * max(1, len(keys)))()
    for i, (c, s, e) in enumerate(keys):
        arr[i] = UpoRegionKey(c, s, e)
    out = np.zeros(max(1, len(keys) * data.n_samples), dtype=np.uint32)
    status = lib.upo_tags_in_regions(C.byref(inp), 1 if directional else 0, extension, len(keys), arr,
                                     out.ctypes.data_as(C.POINTER(C.c_uint32)))
    del keep
    return status, out[:len(keys) * data.n_samples].reshape(len(keys), data.n_samples)


# ------------------------------------------------------------------------------------------------
# regions-table text (misc/format.cpp:1141-1162) — the drop-in output contract
# ------------------------------------------------------------------------------------------------
def _fmt_g(x: float) -> str:
    if math.isnan(x):
        return "-nan" if math.copysign(1.0, x) < 0 else "nan"
    return "%g" % x


def _fmt_2f(x: float) -> str:
    if math.isnan(x):
        return "-nan" if math.copysign(1.0, x) < 0 else "nan"
    return "%.2f" % x


def format_region_row(row: Dict, names: Sequence[str], peaks=False, corrs=False) -> str:
    right = row["left"] + row["len"] - 1
    coord = "%d-%d" % (row["left"], right) if row["forward_label"] else "%d-%d" % (right, row["left"])
    s = "%s:%s" % (names[row["contig"]], coord)
    if peaks:
        s += "\t%d" % row["peak"]
    if corrs:
        s += "\t" + _fmt_g(row["corr"])
    s += "\t" + _fmt_2f(row["kurtosis"])
    for c in row["counts"]:
        s += "\t%d" % c
    return s


def read_table_rows(path: str) -> List[str]:
    rows = []
    with open(path) as f:
        for line in f:
            line = line.rstrip("\n")
            if not line or line[0] == "#" or line[0] == "\t":
                continue
            rows.append(line)
    return rows


# ------------------------------------------------------------------------------------------------
# compiled reference runners (only where oracle/_ref exists)
# ------------------------------------------------------------------------------------------------
def write_inputs(data: TagData, workdir: str, prefix="s") -> Dict:
    os.makedirs(workdir, exist_ok=True)
    ctab = os.path.join(workdir, "contigs.txt")
    data.write_contig_table(ctab)
    wigs = []
    for s in range(data.n_samples):
        p = os.path.join(workdir, "%s%d.wig" % (prefix, s))
        data.write_wig(p, s, "%s%d" % (prefix, s))
        wigs.append(p)
    return dict(contigs=ctab, wigs=wigs)


def run_ref_cli(binary: str, args: Sequence[str], check=True):
    proc = subprocess.run([os.path.join(REF_DIR, binary)] + list(args), stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    if check and proc.returncode != 0:
        raise RuntimeError("%s failed (%d): %s" % (binary, proc.returncode, proc.stderr[-2000:]))
    return proc


def _hex(x: float) -> str:
    return float(x).hex()


def run_ref_replay_regions(data: TagData, st: Dict, workdir: str, reps=1) -> Dict:
    os.makedirs(workdir, exist_ok=True)
    inp = os.path.join(workdir, "in.upbi")
    outp = os.path.join(workdir, "out.bin")
    data.write_upbi(inp)
    args = ["regions", inp, outp, "bw=%d" % st["bandwidth"], "r=" + _hex(st["region_threshold"]), "k=" + _hex(st["kurtosis_threshold"]),
            "u=" + _hex(st["corr_threshold"]), "t=" + _hex(st["hit_threshold"]), "background=" + _hex(st["background"]),
            "directional=%d" % (1 if st["directional"] else 0), "reps=%d" % reps]
    ctrl = [str(i) for i, c in enumerate(st["control"]) if c]
    if ctrl:
        args.append("control=" + ",".join(ctrl))
    if len(st["coeffs"]):
        args.append("coeffs=" + ",".join(_hex(x) for x in st["coeffs"]))
    proc = run_ref_cli("ref_replay", args)
    seconds = float(proc.stdout.strip().split("=")[1])
    n = data.n_samples
    with open(outp, "rb") as f:
        buf = f.read()
    o = 0
    n_rows, n_dropped, n_passed, n_rejected = struct.unpack_from("<QQQQ", buf, o); o += 32
    tags = list(struct.unpack_from("<%dQ" % n, buf, o)); o += 8 * n

    def rd(k):
        nonlocal o
        rows = []
        for _ in range(k):
            fwd, contig, left, ln, peak, pscore, kurt, corr = struct.unpack_from("<BIIIIddd", buf, o); o += 1 + 16 + 24
            counts = list(struct.unpack_from("<%dI" % n, buf, o)); o += 4 * n
            rows.append(dict(forward_label=fwd, contig=contig, left=left, len=ln, peak=peak, peak_score=pscore, kurtosis=kurt, corr=corr, counts=counts))
        return rows
    rows = rd(n_rows)
    dropped = rd(n_dropped)
    return dict(rows=rows, dropped=dropped, n_passed=n_passed, n_rejected=n_rejected, tags_in_regions=tags, seconds=seconds)


def run_ref_replay_shift(data: TagData, workdir: str, *, bandwidth=50, region_threshold=25.0, kurtosis_threshold=50.0,
                         corr_threshold=0.3, hit_threshold=10, min_shift=25, max_shift=150, n_regions_to_test=1000, mappable=0) -> Dict:
    os.makedirs(workdir, exist_ok=True)
    inp = os.path.join(workdir, "in.upbi")
    outp = os.path.join(workdir, "shift.bin")
    data.write_upbi(inp)
    total = sum(data.tags(s) for s in range(data.n_samples))
    background = float(total) / float(mappable or data.genome_size())
    args = ["shift", inp, outp, "bw=%d" % bandwidth, "r=" + _hex(region_threshold), "k=" + _hex(kurtosis_threshold), "u=" + _hex(corr_threshold),
            "t=%d" % hit_threshold, "background=" + _hex(background), "min_shift=%d" % min_shift, "max_shift=%d" % max_shift, "g=%d" % n_regions_to_test]
    proc = run_ref_cli("ref_replay", args)
    with open(outp, "rb") as f:
        buf = f.read()
    o = 0
    n_found, n_tested, n_tags, best = struct.unpack_from("<QIQI", buf, o); o += 24
    (m,) = struct.unpack_from("<I", buf, o); o += 4
    freq = list(struct.unpack_from("<%dQ" % m, buf, o)); o += 8 * m
    tested = []
    for _ in range(n_tested):
        c, l, ln, b, corr = struct.unpack_from("<IIIId", buf, o); o += 24
        tested.append(dict(contig=c, left=l, len=ln, best_shift=b, best_corr=corr))
    return dict(n_regions_found=n_found, n_regions_tested=n_tested, n_tags_in_regions=n_tags, best_shift=best, shift_freq=freq,
                tested=tested, seconds=float(proc.stdout.strip().split("=")[1]))


# ------------------------------------------------------------------------------------------------
# golden fixtures (tests/golden/*.npz, written by tests/golden/make_golden.py from the reference itself)
# ------------------------------------------------------------------------------------------------
GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def golden_names(kind: str) -> List[str]:
    out = []
    for f in sorted(os.listdir(GOLDEN_DIR)):
        if f.endswith(".npz") and f.startswith(kind + "_"):
            out.append(f[:-4])
    return out


def _f(a) -> float:
    return float(np.asarray(a, dtype=np.uint64).view(np.float64).reshape(-1)[0])


def load_golden(name: str):
    """Returns (TagData, settings dict, expected dict) with doubles restored from their bit patterns."""
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    data = TagData(z["contig_sizes"].astype(np.uint32), int(z["n_samples"]), z["off"].astype(np.uint64), z["pos"].astype(np.uint32), z["cnt"].astype(np.uint32))
    kind = str(z["kind"])
    if kind == "tir":
        keys = [tuple(int(x) for x in row) for row in z["keys"]]
        return data, dict(directional=bool(int(z["directional"])), extension=int(z["extension"]), keys=keys), z["counts"].astype(np.uint32)
    if kind == "regions":
        st = dict(bandwidth=int(z["bandwidth"]), region_threshold=_f(z["region_threshold"]), kurtosis_threshold=_f(z["kurtosis_threshold"]),
                  corr_threshold=_f(z["corr_threshold"]), hit_threshold=_f(z["hit_threshold"]), directional=bool(int(z["directional"])),
                  control=z["control"].astype(np.uint8), coeffs=[float(x) for x in z["coeffs"].view(np.float64)], background=_f(z["background"]))

        def rows(prefix):
            k = z[prefix + "left"].shape[0]
            return [dict(forward_label=int(z[prefix + "forward_label"][i]), contig=int(z[prefix + "contig"][i]), left=int(z[prefix + "left"][i]),
                         len=int(z[prefix + "len"][i]), peak=int(z[prefix + "peak"][i]), peak_score=float(z[prefix + "peak_score"].view(np.float64)[i]),
                         kurtosis=float(z[prefix + "kurtosis"].view(np.float64)[i]), corr=float(z[prefix + "corr"].view(np.float64)[i]),
                         counts=[int(x) for x in z[prefix + "counts"][i]]) for i in range(k)]
        want = dict(rows=rows("row_"), dropped=rows("drop_"), n_passed=int(z["n_passed"]), n_rejected=int(z["n_rejected"]),
                    tags_in_regions=[int(x) for x in z["tags_in_regions"]])
        return data, st, want
    p = {}
    for k in z.files:
        if k.startswith("p_"):
            v = z[k]
            p[k[2:]] = _f(v) if v.dtype == np.uint64 else int(v)
    want = dict(n_regions_found=int(z["n_regions_found"]), n_regions_tested=int(z["n_regions_tested"]), n_tags_in_regions=int(z["n_tags_in_regions"]),
                best_shift=int(z["best_shift"]), shift_freq=[int(x) for x in z["shift_freq"]], background=_f(z["background"]),
                tested=[dict(contig=int(z["tested_contig"][i]), left=int(z["tested_left"][i]), len=int(z["tested_len"][i]),
                             best_shift=int(z["tested_best_shift"][i]), best_corr=float(z["tested_best_corr"].view(np.float64)[i]))
                        for i in range(z["tested_contig"].shape[0])])
    return data, p, want


def same_fp(a: float, b: float) -> bool:
    return a == b or (math.isnan(a) and math.isnan(b))


def assert_rows_equal(got, want, what, fp_keys=("peak_score", "kurtosis", "corr")):
    """Bit-exact comparison of region rows (tolerance 0 on the FP64 fields too).  The reference computes
    strandCorr only when asked; rows whose expected corr is NaN are compared on the other fields."""
    assert len(got) == len(want), "%s: %d rows, expected %d" % (what, len(got), len(want))
    for i, (g, w) in enumerate(zip(got, want)):
        for k in ("forward_label", "contig", "left", "len", "peak", "counts"):
            assert g[k] == w[k], "%s row %d field %s: got %r, expected %r (%r)" % (what, i, k, g[k], w[k], w)
        for k in fp_keys:
            assert same_fp(g[k], w[k]), "%s row %d field %s: got %r, expected %r" % (what, i, k, g[k], w[k])


# ---- fixtures for the reference's left-edge quirk in its rarer shapes ------------------------------------
def chain_leak_data(variant: int, seed: int = 5):
    """Contigs whose EVERY tag lies within `bandwidth` (50) of the contig start: the reference's flush cannot drain the
    window (misc/peakcall.cpp:176-184,224-231), so stale cells - and sometimes an open region - travel into the next
    contig the same ProfileBuffer sees.  Returns (TagData, control list).

    variant 0: normal contig, all-early contig, contig whose first tag is far from the start, normal contig
    variant 1: all-early contig followed by a contig that only the CONTROL sample touches, then a normal one
    variant 2: two all-early contigs in a row, then one whose own first tags are early as well (but not all of them)
    variant 3: the all-early contig is the last one (nothing inherits the window)"""
    rng = np.random.default_rng(seed + 100 * variant)
    n = 3 if variant == 1 else 2

    def normal(size, first=60):
        centres = rng.integers(first + 200, size - 300, size=max(2, size // 4000))
        pos = np.unique(np.clip(np.concatenate([c + rng.integers(-60, 60, size=40) for c in centres]), first, size - 60)).astype(np.uint32)
        return pos, rng.integers(1, 7, size=pos.shape[0]).astype(np.uint32)

    def early(lo=2, hi=48, k=14, heavy=True):
        pos = np.sort(rng.choice(np.arange(lo, hi + 1), size=k, replace=False)).astype(np.uint32)
        return pos, rng.integers(3 if heavy else 1, 12 if heavy else 3, size=k).astype(np.uint32)

    def none():
        return np.zeros(0, np.uint32), np.zeros(0, np.uint32)

    shapes = {
        0: ["normal", "early", "far", "normal"],
        1: ["normal", "early", "control_only", "normal"],
        2: ["early", "early", "early_then_normal", "normal"],
        3: ["normal", "normal", "early"],
    }[variant]
    sizes = [30000, 20000, 25000, 30000][:len(shapes)]
    segs = []
    for s in range(n):
        per = [[None] * len(shapes), [None] * len(shapes)]
        is_control = variant == 1 and s == 2
        for d in range(2):
            for c, shape in enumerate(shapes):
                size = sizes[c]
                if shape == "control_only":
                    per[d][c] = normal(size, first=70 + 40 * d) if is_control else none()
                elif is_control:
                    per[d][c] = normal(size) if shape == "normal" else none()
                elif shape == "normal":
                    per[d][c] = normal(size)
                elif shape == "early":
                    per[d][c] = early(k=10 + 3 * s + d)
                elif shape == "far":
                    p, k = normal(size, first=4000)
                    per[d][c] = (p, k)
                elif shape == "early_then_normal":
                    p0, k0 = early(lo=5, hi=40, k=8)
                    p1, k1 = normal(size, first=200)
                    per[d][c] = (np.concatenate([p0, p1]), np.concatenate([k0, k1]))
        segs.append(per)
    data = TagData.from_segments(sizes, segs, ["chr%s" % "ABCD"[c] for c in range(len(shapes))])
    return data, ([2] if variant == 1 else [])


def long_start_region_data(seed: int = 9):
    """Non-directional sample with a chrM-like contig: dense from position 1 on, reverse tags displaced by ~80, so the region that
    starts at the contig start is long enough for the shift sweep of bin/strand_shift (src/strand_shift.cpp:206)."""
    rng = np.random.default_rng(seed)
    sizes = [16000, 600000]
    per = [[None, None], [None, None]]
    base = np.arange(1, 2400, dtype=np.uint32)
    base = base[rng.random(base.shape[0]) < 0.7]
    per[0][0] = (base, rng.integers(1, 5, size=base.shape[0]).astype(np.uint32))
    rev = np.unique(np.clip(base.astype(np.int64) + 80 + rng.integers(-3, 4, size=base.shape[0]), 1, sizes[0])).astype(np.uint32)
    per[1][0] = (rev, rng.integers(1, 5, size=rev.shape[0]).astype(np.uint32))
    # ordinary peaks on the second contig
    centres = rng.integers(2000, sizes[1] - 2000, size=60)
    for d in range(2):
        pos = np.unique(np.concatenate([c + 90 * d + rng.integers(-150, 150, size=260) for c in centres])).astype(np.uint32)
        per[d][1] = (pos, rng.integers(1, 4, size=pos.shape[0]).astype(np.uint32))
    return TagData.from_segments(sizes, [per], ["chrM", "chr1"])
