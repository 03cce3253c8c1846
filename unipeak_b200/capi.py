"""ctypes binding of the C-ABI in include/unipeak_b200.h (the CUDA library built in-tree as
unipeak_b200/libunipeak_b200.so).  There is no fallback: if the library is missing, or no CUDA device is
present, every call fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, List, Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("UPB_LIB") or os.path.join(_HERE, "libunipeak_b200.so")   # UPB_LIB: experiment builds (tools/)

UPB_OK = 0
ERRORS = {
    -1: "UPB_ERR_CUDA", -2: "UPB_ERR_BAD_ARG", -3: "UPB_ERR_NO_DEVICE", -4: "UPB_ERR_NOT_LOADED", -5: "UPB_ERR_UNSORTED",
    -6: "UPB_ERR_LEFT_EDGE", -7: "UPB_ERR_REFERENCE_ABORTS", -8: "UPB_ERR_NO_REGIONS", -9: "UPB_ERR_REVERSE_NONDIR",
    -10: "UPB_ERR_NOMEM", -11: "UPB_ERR_UNSUPPORTED",
}


class UpbError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__("%s (%d): %s" % (ERRORS.get(code, "UPB_ERR"), code, msg))
        self.code = code


class RegionParams(C.Structure):
    _fields_ = [("bandwidth", C.c_uint32), ("region_threshold", C.c_double), ("kurtosis_threshold", C.c_double),
                ("corr_threshold", C.c_double), ("hit_threshold", C.c_double), ("directional", C.c_int),
                ("control", C.POINTER(C.c_uint8)), ("n_coeffs", C.c_uint32), ("coeffs", C.POINTER(C.c_double)),
                ("background", C.c_double), ("reference_order", C.c_int)]


class RegionTable(C.Structure):
    _fields_ = [("n_rows", C.c_uint64), ("n_samples", C.c_uint32), ("buffer", C.POINTER(C.c_uint8)),
                ("forward_label", C.POINTER(C.c_uint8)), ("dropped", C.POINTER(C.c_uint8)), ("contig", C.POINTER(C.c_uint32)),
                ("left", C.POINTER(C.c_uint32)), ("len", C.POINTER(C.c_uint32)), ("peak", C.POINTER(C.c_uint32)),
                ("hit_sum", C.POINTER(C.c_uint32)), ("cand_index", C.POINTER(C.c_uint64)), ("peak_score", C.POINTER(C.c_double)), ("kurtosis", C.POINTER(C.c_double)),
                ("corr", C.POINTER(C.c_double)), ("counts", C.POINTER(C.c_uint32)), ("n_passed", C.c_uint64),
                ("n_rejected", C.c_uint64), ("tags_in_regions", C.POINTER(C.c_uint64)), ("unit_candidates", C.POINTER(C.c_uint64)),
                ("unit_last_tag", C.POINTER(C.c_uint32)), ("unit_rows", C.POINTER(C.c_uint64)), ("_slab", C.c_void_p), ("_slab_bytes", C.c_uint64), ("_ctx", C.c_void_p)]


class Profile(C.Structure):
    _fields_ = [("n", C.c_uint64), ("pos", C.POINTER(C.c_uint32)), ("score", C.POINTER(C.c_double))]


class ShiftParams(C.Structure):
    _fields_ = [("bandwidth", C.c_uint32), ("region_threshold", C.c_double), ("kurtosis_threshold", C.c_double),
                ("corr_threshold", C.c_double), ("hit_threshold", C.c_uint32), ("min_shift", C.c_uint32),
                ("max_shift", C.c_uint32), ("n_regions_to_test", C.c_uint32), ("background", C.c_double)]


class ShiftReport(C.Structure):
    _fields_ = [("n_regions_found", C.c_uint64), ("n_regions_tested", C.c_uint32), ("n_tags_in_regions", C.c_uint64),
                ("best_shift", C.c_uint32), ("n_freq", C.c_uint32), ("shift_freq", C.POINTER(C.c_uint64)),
                ("tested_contig", C.POINTER(C.c_uint32)), ("tested_left", C.POINTER(C.c_uint32)), ("tested_len", C.POINTER(C.c_uint32)),
                ("tested_best_shift", C.POINTER(C.c_uint32)), ("tested_best_corr", C.POINTER(C.c_double))]


class ShiftRegionList(C.Structure):
    _fields_ = [("n", C.c_uint64), ("contig", C.POINTER(C.c_uint32)), ("left", C.POINTER(C.c_uint32)), ("len", C.POINTER(C.c_uint32)),
                ("sum", C.POINTER(C.c_uint32)), ("replayed", C.POINTER(C.c_uint8))]


class RegionKey(C.Structure):
    _fields_ = [("contig", C.c_uint32), ("start", C.c_uint32), ("end", C.c_uint32)]


class Stats(C.Structure):
    _fields_ = [("ms_total", C.c_float), ("ms_kde", C.c_float), ("launches", C.c_uint32), ("kde_tiles", C.c_uint64),
                ("kde_hot_tiles", C.c_uint64), ("n_candidates", C.c_uint64), ("n_entries", C.c_uint64), ("hot_blocks", C.c_uint64),
                ("exact_evals", C.c_uint64), ("spans_with_tags", C.c_uint64), ("n_stitched", C.c_uint64), ("warm_spans", C.c_uint64), ("ms_call", C.c_float),
                ("ms_host_prepare", C.c_float), ("ms_host_enqueue", C.c_float), ("ms_host_wait", C.c_float), ("ms_host_post", C.c_float), ("ms_host_assemble", C.c_float)]


EXPORTS = [
    "upb_create", "upb_destroy", "upb_last_error", "upb_device_count", "upb_host_alloc", "upb_host_free", "upb_set_contigs",
    "upb_load_tags", "upb_call_regions", "upb_free_region_table", "upb_reference_order", "upb_profile_contig", "upb_free_profile",
    "upb_strand_shift", "upb_free_shift_report", "upb_count_tags_in_regions", "upb_kernel_taps", "upb_last_stats",
    "upb_shared_table_bytes", "upb_shared_table_init", "upb_call_regions_shared", "upb_set_profiling", "upb_last_stage_times", "upb_stream", "upb_host_register", "upb_host_unregister", "upb_load_tags_async", "upb_load_wait", "upb_reindex", "upb_aggregate_tags", "upb_free_tag_counts",
    "upb_load_tags_packed2", "upb_pack_tags2", "upb_load_tags_vbyte", "upb_pack_tags_vbyte", "upb_last_load_ms",
    "upb_shift_regions", "upb_free_shift_regions", "upb_shift_examine_order", "upb_shift_sweep", "upb_shift_mode",
    "upb_multi_create", "upb_multi_destroy", "upb_multi_last_error", "upb_multi_gpus", "upb_multi_context", "upb_multi_set_contigs", "upb_multi_owner",
    "upb_multi_load_tags", "upb_multi_load_wait", "upb_multi_call_regions", "upb_multi_profile_contig", "upb_multi_strand_shift",
    "upb_multi_load_tags_by_sample", "upb_multi_count_tags_in_regions",
]


class TagSource(C.Structure):
    _fields_ = [("n_samples", C.c_uint32), ("off", C.c_void_p), ("pos", C.c_void_p), ("cnt", C.c_void_p), ("words", C.c_void_p),
                ("n_exceptions", C.c_uint64), ("exceptions", C.c_void_p), ("keep_contig", C.c_void_p), ("directional_hint", C.c_int),
                ("vbytes", C.c_void_p), ("vbyte_off", C.c_void_p), ("vtile_off", C.c_void_p), ("n_vtiles", C.c_uint64)]


class SharedTable(C.Structure):
    _fields_ = [("host", C.c_void_p), ("device", C.c_void_p), ("bytes", C.c_uint64), ("rank", C.c_uint32), ("world", C.c_uint32),
                ("n_contigs", C.c_uint32), ("n_samples", C.c_uint32), ("row_capacity", C.c_uint64)]


_lib = None


class TagCounts(C.Structure):
    _fields_ = [("n", C.c_uint64), ("tag_count", C.c_uint64), ("off", C.POINTER(C.c_uint64)), ("pos", C.POINTER(C.c_uint32)), ("cnt", C.POINTER(C.c_uint32))]


PACKED_EXC_DTYPE = np.dtype([("index", np.uint64), ("pos", np.uint32), ("cnt", np.uint32)])


def pack_tags2(n_samples: int, n_contigs: int, off, pos, cnt, out=None):
    """upb_pack_tags2: (words u16 = 12-bit position delta | 4-bit count << 12, exceptions) of the plain arrays."""
    L = lib()
    off = np.ascontiguousarray(off, dtype=np.uint64)
    pos = np.ascontiguousarray(pos, dtype=np.uint32)
    cnt = np.ascontiguousarray(cnt, dtype=np.uint32)
    E = int(pos.shape[0])
    n_exc = C.c_uint64(0)
    L.upb_pack_tags2(n_samples, n_contigs, off.ctypes.data, pos.ctypes.data, cnt.ctypes.data, None, None, 0, C.byref(n_exc))
    words = out if out is not None else np.empty(max(E, 1), np.uint16)
    exc = np.zeros(max(int(n_exc.value), 1), dtype=PACKED_EXC_DTYPE)
    rc = L.upb_pack_tags2(n_samples, n_contigs, off.ctypes.data, pos.ctypes.data, cnt.ctypes.data, words.ctypes.data,
                          exc.ctypes.data, int(n_exc.value), C.byref(n_exc))
    if rc:
        raise UpbError(rc, "upb_pack_tags2 failed")
    return words, exc[:int(n_exc.value)]


def pack_tags_vbyte(n_samples: int, n_contigs: int, off, pos, cnt, out=None):
    """upb_pack_tags_vbyte: the variable-length byte form of the plain arrays -> (bytes u8, byte_off u64 [layout of off], tile_off u64
    [tiles + 1], exceptions).  `out`: a u8 array to encode into (e.g. pinned); with it the encoder runs once (the tile count follows from `off`,
    the exception list is sized by guess and the call repeated if the guess was short), without it a first call reports the sizes."""
    L = lib()
    off = np.ascontiguousarray(off, dtype=np.uint64)
    pos = np.ascontiguousarray(pos, dtype=np.uint32)
    cnt = np.ascontiguousarray(cnt, dtype=np.uint32)
    nb, nt, nx = C.c_uint64(0), C.c_uint64(0), C.c_uint64(0)
    args = (n_samples, n_contigs, off.ctypes.data, pos.ctypes.data, cnt.ctypes.data)
    E = int(pos.shape[0])
    if out is None:
        L.upb_pack_tags_vbyte(*args, None, 0, C.byref(nb), None, None, 0, C.byref(nt), None, 0, C.byref(nx))
        n_bytes, n_tiles, exc_cap = int(nb.value), int(nt.value), int(nx.value)
        buf = np.empty(max(n_bytes, 1), np.uint8)
    else:
        lens = np.diff(off.reshape(n_samples * 2, n_contigs + 1).astype(np.int64), axis=1)
        n_tiles = int(((lens + 2047) // 2048).sum())
        exc_cap = max(4096, E // 64)
        buf = out
    byte_off = np.zeros(max(off.shape[0], 1), np.uint64)
    for _ in range(2):
        tile_off = np.zeros(n_tiles + 1, np.uint64)
        exc = np.zeros(max(exc_cap, 1), dtype=PACKED_EXC_DTYPE)
        rc = L.upb_pack_tags_vbyte(*args, buf.ctypes.data, int(buf.shape[0]), C.byref(nb), byte_off.ctypes.data, tile_off.ctypes.data, n_tiles + 1, C.byref(nt),
                                   exc.ctypes.data, exc_cap, C.byref(nx))
        if rc == 0:
            return buf[:int(nb.value)], byte_off, tile_off, exc[:int(nx.value)]
        if int(nb.value) > buf.shape[0]:
            raise ValueError("pack_tags_vbyte: `out` holds %d bytes, %d needed" % (buf.shape[0], int(nb.value)))
        n_tiles, exc_cap = int(nt.value), int(nx.value)                    # the sizes the encoder reported: once more with those
    raise UpbError(rc, "upb_pack_tags_vbyte failed")


def host_register(ptr: int, nbytes: int) -> int:
    """Page-lock + map caller-owned host memory (upb_host_register); returns the device-side address of it."""
    dp = C.c_void_p()
    rc = lib().upb_host_register(C.c_void_p(ptr), nbytes, C.byref(dp))
    if rc:
        raise UpbError(rc, lib().upb_last_error(None).decode())
    return int(dp.value)


def host_unregister(ptr: int) -> None:
    lib().upb_host_unregister(C.c_void_p(ptr))


def lib():
    """Load the CUDA library; raise if it has not been built (python __graft_entry__.py / unipeak_b200.build)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("%s is missing: build it with `python -m unipeak_b200.build` (nvcc, sm_100a). "
                              "There is no CPU fallback." % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        L.upb_create.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
        L.upb_destroy.argtypes = [C.c_void_p]
        L.upb_destroy.restype = None
        L.upb_last_error.argtypes = [C.c_void_p]
        L.upb_last_error.restype = C.c_char_p
        L.upb_host_alloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t]
        L.upb_host_free.argtypes = [C.c_void_p]
        L.upb_host_free.restype = None
        L.upb_set_contigs.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p]
        L.upb_load_tags.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]
        L.upb_call_regions.argtypes = [C.c_void_p, C.POINTER(RegionParams), C.POINTER(RegionTable)]
        L.upb_free_region_table.argtypes = [C.POINTER(RegionTable)]
        L.upb_free_region_table.restype = None
        L.upb_profile_contig.argtypes = [C.c_void_p, C.POINTER(RegionParams), C.c_int, C.c_uint32, C.POINTER(Profile)]
        L.upb_free_profile.argtypes = [C.POINTER(Profile)]
        L.upb_free_profile.restype = None
        L.upb_strand_shift.argtypes = [C.c_void_p, C.POINTER(ShiftParams), C.POINTER(ShiftReport)]
        L.upb_free_shift_report.argtypes = [C.POINTER(ShiftReport)]
        L.upb_free_shift_report.restype = None
        L.upb_shift_regions.argtypes = [C.c_void_p, C.POINTER(ShiftParams), C.POINTER(ShiftRegionList)]
        L.upb_free_shift_regions.argtypes = [C.POINTER(ShiftRegionList)]
        L.upb_free_shift_regions.restype = None
        L.upb_shift_examine_order.argtypes = [C.c_uint64, C.c_void_p, C.c_uint64, C.c_void_p, C.POINTER(C.c_uint64)]
        L.upb_shift_sweep.argtypes = [C.c_void_p, C.POINTER(ShiftParams), C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.upb_shift_mode.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.POINTER(C.c_uint32)]
        L.upb_count_tags_in_regions.argtypes = [C.c_void_p, C.c_int, C.c_uint32, C.c_uint64, C.POINTER(RegionKey), C.c_void_p]
        L.upb_kernel_taps.argtypes = [C.c_uint32, C.c_double, C.POINTER(C.c_double)]
        L.upb_last_stats.argtypes = [C.c_void_p, C.POINTER(Stats)]
        L.upb_stream.argtypes = [C.c_void_p]
        L.upb_load_tags_async.argtypes = [C.c_void_p, C.POINTER(TagSource)]
        L.upb_load_wait.argtypes = [C.c_void_p]
        L.upb_last_load_ms.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_float)]
        L.upb_reindex.argtypes = [C.c_void_p]
        L.upb_load_tags_packed2.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]
        L.upb_load_tags_vbyte.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_void_p]
        L.upb_pack_tags_vbyte.argtypes = [C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64), C.c_void_p,
                                          C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64), C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64)]
        L.upb_pack_tags2.argtypes = [C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64)]
        L.upb_host_register.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p)]
        L.upb_host_unregister.argtypes = [C.c_void_p]
        L.upb_host_unregister.restype = None
        L.upb_stream.restype = C.c_void_p
        L.upb_shared_table_bytes.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint64]
        L.upb_shared_table_bytes.restype = C.c_uint64
        L.upb_shared_table_init.argtypes = [C.POINTER(SharedTable)]
        L.upb_call_regions_shared.argtypes = [C.c_void_p, C.POINTER(RegionParams), C.POINTER(SharedTable), C.c_uint64, C.POINTER(RegionTable)]
        L.upb_set_profiling.argtypes = [C.c_void_p, C.c_int]
        L.upb_last_stage_times.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(C.c_char_p), C.POINTER(C.c_float), C.POINTER(C.c_uint32)]
        L.upb_reference_order.argtypes = [C.c_uint32, C.c_uint32, C.c_int, C.c_uint32, C.c_void_p, C.c_void_p, C.POINTER(RegionTable), C.c_void_p]
        L.upb_multi_create.argtypes = [C.c_int, C.c_void_p, C.POINTER(C.c_void_p)]
        L.upb_multi_destroy.argtypes = [C.c_void_p]
        L.upb_multi_destroy.restype = None
        L.upb_multi_last_error.argtypes = [C.c_void_p]
        L.upb_multi_last_error.restype = C.c_char_p
        L.upb_multi_gpus.argtypes = [C.c_void_p]
        L.upb_multi_context.argtypes = [C.c_void_p, C.c_int]
        L.upb_multi_context.restype = C.c_void_p
        L.upb_multi_set_contigs.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p]
        L.upb_multi_owner.argtypes = [C.c_void_p, C.c_void_p]
        L.upb_multi_load_tags.argtypes = [C.c_void_p, C.POINTER(TagSource)]
        L.upb_multi_load_wait.argtypes = [C.c_void_p]
        L.upb_multi_call_regions.argtypes = [C.c_void_p, C.POINTER(RegionParams), C.POINTER(RegionTable)]
        L.upb_multi_profile_contig.argtypes = [C.c_void_p, C.POINTER(RegionParams), C.c_int, C.c_uint32, C.POINTER(Profile)]
        L.upb_multi_strand_shift.argtypes = [C.c_void_p, C.POINTER(ShiftParams), C.POINTER(ShiftReport)]
        L.upb_multi_load_tags_by_sample.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]
        L.upb_multi_count_tags_in_regions.argtypes = [C.c_void_p, C.c_int, C.c_uint32, C.c_uint64, C.POINTER(RegionKey), C.c_void_p]
        _lib = L
    return _lib


def shared_table_bytes(n_contigs: int, n_samples: int, world: int, row_capacity: int) -> int:
    return int(lib().upb_shared_table_bytes(n_contigs, n_samples, world, row_capacity))


def shift_examine_order(sums: np.ndarray, want: int):
    """upb_shift_examine_order (host only): (order, sorted_end) — order[:sorted_end] are the indices of the first sorted_end
    elements of the reference's std::sort of the regions by Region::sum() descending (src/strand_shift.cpp:198)."""
    sums = np.ascontiguousarray(sums, dtype=np.uint32)
    order = np.empty(sums.shape[0], dtype=np.uint32)
    end = C.c_uint64(0)
    rc = lib().upb_shift_examine_order(sums.shape[0], sums.ctypes.data, int(want), order.ctypes.data, C.byref(end))
    if rc:
        raise UpbError(rc, "upb_shift_examine_order failed")
    return order, int(end.value)


def shift_mode(shift_freq: np.ndarray, min_shift: int) -> int:
    """upb_shift_mode (host only): src/strand_shift.cpp:239-258."""
    f = np.ascontiguousarray(shift_freq, dtype=np.uint64)
    best = C.c_uint32(0)
    rc = lib().upb_shift_mode(f.ctypes.data, f.shape[0], int(min_shift), C.byref(best))
    if rc:
        raise UpbError(rc, "upb_shift_mode failed")
    return int(best.value)


def kernel_taps(bandwidth: int, total: float = 1.0) -> np.ndarray:
    """Kernel(bandwidth, sum) of the reference (misc/kernel.cpp:16-35)."""
    out = np.zeros(2 * bandwidth + 1, dtype=np.float64)
    rc = lib().upb_kernel_taps(bandwidth, total, out.ctypes.data_as(C.POINTER(C.c_double)))
    if rc:
        raise UpbError(rc, "upb_kernel_taps")
    return out


def _np(ptr, n, dtype):
    if n == 0:
        return np.zeros(0, dtype=dtype)
    return np.ctypeslib.as_array(ptr, shape=(n,)).astype(dtype, copy=True)


class RegionView:
    """Lazy, zero-copy access to a upb_region_table (see Context.call_regions_view)."""

    def __init__(self, table, n_contigs):
        self._t = table
        self._n_contigs = n_contigs
        self._cols = None
        self.n_rows = int(table.n_rows)
        self.n_passed = int(table.n_passed)
        self.n_rejected = int(table.n_rejected)

    def columns(self) -> Dict:
        if self._cols is None:
            self._cols = Context._views(self._t, self._n_contigs)
        return self._cols

    def __getitem__(self, key):
        if key in ("n_rows", "n_passed", "n_rejected"):
            return getattr(self, key)
        return self.columns()[key]


class Context:
    """One device context (one per host thread / rank)."""

    def __init__(self, device: int = 0):
        self._lib = lib()
        self._h = C.c_void_p()
        rc = self._lib.upb_create(device, C.byref(self._h))
        if rc:
            raise UpbError(rc, self._lib.upb_last_error(None).decode())
        self.n_contigs = 0
        self.n_samples = 0

    def close(self):
        if self._h:
            self.release_view()
            self._lib.upb_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int):
        if rc:
            raise UpbError(rc, self._lib.upb_last_error(self._h).decode())

    def set_contigs(self, sizes: Sequence[int]):
        arr = np.ascontiguousarray(sizes, dtype=np.uint32)
        self._check(self._lib.upb_set_contigs(self._h, arr.shape[0], arr.ctypes.data))
        self.n_contigs = int(arr.shape[0])

    def load_tags(self, n_samples: int, off: np.ndarray, pos: np.ndarray, cnt: np.ndarray):
        off = np.ascontiguousarray(off, dtype=np.uint64)
        pos = np.ascontiguousarray(pos, dtype=np.uint32)
        cnt = np.ascontiguousarray(cnt, dtype=np.uint32)
        self._check(self._lib.upb_load_tags(self._h, n_samples, off.ctypes.data, pos.ctypes.data, cnt.ctypes.data))
        self.n_samples = n_samples

    def load_tags_ptr(self, n_samples: int, off_ptr: int, pos_ptr: int, cnt_ptr: int):
        """Same, from raw host addresses (e.g. pinned buffers owned by the caller)."""
        self._check(self._lib.upb_load_tags(self._h, n_samples, C.c_void_p(off_ptr), C.c_void_p(pos_ptr), C.c_void_p(cnt_ptr)))
        self.n_samples = n_samples

    def load_tags_packed2_ptr(self, n_samples: int, off_ptr: int, words_ptr: int, n_exc: int, exc_ptr: int):
        """upb_load_tags_packed2 from raw host addresses (see pack_tags2)."""
        self._check(self._lib.upb_load_tags_packed2(self._h, n_samples, C.c_void_p(off_ptr), C.c_void_p(words_ptr), n_exc, C.c_void_p(exc_ptr)))
        self.n_samples = n_samples

    def load_packed2(self, data) -> None:
        """Load a TagData through the 2-byte host form."""
        self.set_contigs(data.contig_sizes)
        off = np.ascontiguousarray(data.off, dtype=np.uint64)
        words, exc = pack_tags2(data.n_samples, self.n_contigs, off, data.pos, data.cnt)
        self.load_tags_packed2_ptr(data.n_samples, off.ctypes.data, words.ctypes.data, exc.shape[0], exc.ctypes.data)

    def load_vbyte(self, data, keep_contigs: Optional[Sequence[int]] = None) -> None:
        """Load a TagData through the variable-length byte form (all contigs, or — what a rank of a multi-GPU job does — only the given ones)."""
        self.set_contigs(data.contig_sizes)
        off = np.ascontiguousarray(data.off, dtype=np.uint64)
        vb, boff, toff, exc = pack_tags_vbyte(data.n_samples, self.n_contigs, off, data.pos, data.cnt)
        keep = None
        if keep_contigs is not None:
            keep = np.zeros(self.n_contigs, dtype=np.uint8)
            keep[list(keep_contigs)] = 1
        self._hold = (off, vb, boff, toff, exc)
        self.load_async_ptr(data.n_samples, off.ctypes.data, vbytes_ptr=vb.ctypes.data, vbyte_off_ptr=boff.ctypes.data, vtile_off_ptr=toff.ctypes.data,
                            n_vtiles=toff.shape[0] - 1, n_exc=exc.shape[0], exc_ptr=exc.ctypes.data, keep_contig=keep)
        self.load_wait()
        self._hold = None

    def load_async_ptr(self, n_samples: int, off_ptr: int, *, pos_ptr: int = 0, cnt_ptr: int = 0, words_ptr: int = 0, n_exc: int = 0, exc_ptr: int = 0,
                       keep_contig: Optional[np.ndarray] = None, directional_hint: bool = True,
                       vbytes_ptr: int = 0, vbyte_off_ptr: int = 0, vtile_off_ptr: int = 0, n_vtiles: int = 0) -> None:
        """upb_load_tags_async from raw host addresses: returns once the transfers are queued.  The buffers (and `keep_contig`) must stay
        alive until load_wait() or the next call that uses the data returns."""
        self._keep = None if keep_contig is None else np.ascontiguousarray(keep_contig, dtype=np.uint8)
        src = TagSource(n_samples, off_ptr, pos_ptr or None, cnt_ptr or None, words_ptr or None, n_exc, exc_ptr or None,
                        None if self._keep is None else self._keep.ctypes.data, 1 if directional_hint else 0,
                        vbytes_ptr or None, vbyte_off_ptr or None, vtile_off_ptr or None, n_vtiles)
        self._check(self._lib.upb_load_tags_async(self._h, C.byref(src)))
        self.n_samples = n_samples

    def load_wait(self) -> None:
        self._check(self._lib.upb_load_wait(self._h))

    def last_load_ms(self):
        """(ms until the last byte of the last load was on the device, ms until its index was complete), from the first queued copy."""
        a, b = C.c_float(-1), C.c_float(-1)
        rc = self._lib.upb_last_load_ms(self._h, C.byref(a), C.byref(b))
        return (float(a.value), float(b.value)) if rc == 0 else (-1.0, -1.0)

    def reindex(self) -> None:
        """upb_reindex: rebuild the tile-major copy and the zone maps from the resident sample-major arrays."""
        self._check(self._lib.upb_reindex(self._h))

    def load_subset(self, data, keep_contigs: Sequence[int], packed: bool = True) -> None:
        """Load only the segments of the given contigs out of arrays that describe all of them (what a rank of a multi-GPU job does)."""
        self.set_contigs(data.contig_sizes)
        keep = np.zeros(self.n_contigs, dtype=np.uint8)
        keep[list(keep_contigs)] = 1
        off = np.ascontiguousarray(data.off, dtype=np.uint64)
        if packed:
            words, exc = pack_tags2(data.n_samples, self.n_contigs, off, data.pos, data.cnt)
            self._hold = (off, words, exc)
            self.load_async_ptr(data.n_samples, off.ctypes.data, words_ptr=words.ctypes.data, n_exc=exc.shape[0], exc_ptr=exc.ctypes.data, keep_contig=keep)
        else:
            pos = np.ascontiguousarray(data.pos, dtype=np.uint32); cnt = np.ascontiguousarray(data.cnt, dtype=np.uint32)
            self._hold = (off, pos, cnt)
            self.load_async_ptr(data.n_samples, off.ctypes.data, pos_ptr=pos.ctypes.data, cnt_ptr=cnt.ctypes.data, keep_contig=keep)
        self.load_wait()
        self._hold = None

    def aggregate_tags(self, key, pos, cnt=None, combine_strands=False):
        """upb_aggregate_tags: (off[2C+1], pos, cnt, tag_count) — distinct 5' positions with their tag counts per (strand, contig)."""
        key = np.ascontiguousarray(key, dtype=np.uint32)
        pos = np.ascontiguousarray(pos, dtype=np.uint32)
        cnt_p = None
        if cnt is not None:
            cnt = np.ascontiguousarray(cnt, dtype=np.uint32)
            cnt_p = cnt.ctypes.data
        t = TagCounts()
        self._lib.upb_aggregate_tags.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(TagCounts)]
        self._lib.upb_free_tag_counts.argtypes = [C.POINTER(TagCounts)]
        self._check(self._lib.upb_aggregate_tags(self._h, int(key.shape[0]), key.ctypes.data, pos.ctypes.data, cnt_p, 1 if combine_strands else 0, C.byref(t)))
        try:
            n = int(t.n)
            return (_np(t.off, 2 * self.n_contigs + 1, np.uint64), _np(t.pos, n, np.uint32), _np(t.cnt, n, np.uint32), int(t.tag_count))
        finally:
            self._lib.upb_free_tag_counts(C.byref(t))

    def load(self, data) -> None:
        """Load a unipeak_b200.synth.TagData."""
        self.set_contigs(data.contig_sizes)
        self.load_tags(data.n_samples, data.off, data.pos, data.cnt)

    def _region_params(self, bandwidth, region_threshold, kurtosis_threshold, corr_threshold, hit_threshold, directional, control, coeffs,
                       background, reference_order):
        ctrl = np.ascontiguousarray(control if control is not None else np.zeros(self.n_samples, np.uint8), dtype=np.uint8)
        cf = np.ascontiguousarray(coeffs if coeffs is not None else [], dtype=np.float64)
        p = RegionParams(bandwidth, region_threshold, kurtosis_threshold, corr_threshold, hit_threshold, 1 if directional else 0,
                         ctrl.ctypes.data_as(C.POINTER(C.c_uint8)), cf.shape[0], cf.ctypes.data_as(C.POINTER(C.c_double)),
                         background, 1 if reference_order else 0)
        return p, (ctrl, cf)

    @staticmethod
    def _views(t, n_contigs):
        """numpy views (no copies) of the table's columns; valid until the table is freed."""
        R, n = int(t.n_rows), int(t.n_samples)

        def v(ptr, count, dtype):
            if count == 0 or not ptr:
                return np.zeros(0, dtype=dtype)
            nbytes = count * np.dtype(dtype).itemsize                # from_address: ~2 us per view (np.ctypeslib.as_array costs ~60 us each)
            return np.frombuffer((C.c_uint8 * nbytes).from_address(C.addressof(ptr.contents)), dtype=dtype)
        return dict(buffer=v(t.buffer, R, np.uint8), forward_label=v(t.forward_label, R, np.uint8), dropped=v(t.dropped, R, np.uint8),
                    contig=v(t.contig, R, np.uint32), left=v(t.left, R, np.uint32), len=v(t.len, R, np.uint32), peak=v(t.peak, R, np.uint32),
                    hit_sum=v(t.hit_sum, R, np.uint32), cand_index=v(t.cand_index, R, np.uint64), peak_score=v(t.peak_score, R, np.float64),
                    kurtosis=v(t.kurtosis, R, np.float64), corr=v(t.corr, R, np.float64), counts=v(t.counts, R * n, np.uint32).reshape(R, n),
                    tags_in_regions=v(t.tags_in_regions, n, np.uint64), unit_candidates=v(t.unit_candidates, 2 * n_contigs, np.uint64),
                    unit_last_tag=v(t.unit_last_tag, 2 * n_contigs, np.uint32), unit_rows=v(t.unit_rows, 2 * n_contigs, np.uint64))

    @staticmethod
    def _table_dict(t, n_contigs, want_rows=True) -> Dict:
        R = int(t.n_rows)
        cols = {k: np.array(v, copy=True) for k, v in Context._views(t, n_contigs).items()}
        rows, dropped = [], []
        for i in range(R if want_rows else 0):
            row = dict(forward_label=int(cols["forward_label"][i]), buffer=int(cols["buffer"][i]), contig=int(cols["contig"][i]),
                       left=int(cols["left"][i]), len=int(cols["len"][i]), peak=int(cols["peak"][i]),
                       peak_score=float(cols["peak_score"][i]), kurtosis=float(cols["kurtosis"][i]), corr=float(cols["corr"][i]),
                       counts=[int(x) for x in cols["counts"][i]])
            (dropped if cols["dropped"][i] else rows).append(row)
        return dict(rows=rows, dropped=dropped, n_passed=int(t.n_passed), n_rejected=int(t.n_rejected),
                    tags_in_regions=[int(x) for x in cols["tags_in_regions"]], columns=cols,
                    unit_candidates=cols["unit_candidates"], unit_last_tag=cols["unit_last_tag"])

    @staticmethod
    def _shift_dict(rc, r) -> Dict:
        return dict(status=rc, n_regions_found=int(r.n_regions_found), n_regions_tested=int(r.n_regions_tested),
                    n_tags_in_regions=int(r.n_tags_in_regions), best_shift=int(r.best_shift),
                    shift_freq=[int(r.shift_freq[i]) for i in range(int(r.n_freq))] if r.shift_freq else [],
                    tested=[dict(contig=int(r.tested_contig[i]), left=int(r.tested_left[i]), len=int(r.tested_len[i]),
                                 best_shift=int(r.tested_best_shift[i]), best_corr=float(r.tested_best_corr[i]))
                            for i in range(int(r.n_regions_tested))])

    def call_regions_view(self, *, bandwidth=50, region_threshold=25.0, kurtosis_threshold=50.0, corr_threshold=-1.0, hit_threshold=10.0,
                          directional=True, control: Optional[Sequence[int]] = None, coeffs: Optional[Sequence[float]] = None,
                          background: float, reference_order=True, **_ignored) -> "RegionView":
        """upb_call_regions without any Python-side copy: the returned object exposes the table's columns as numpy views
        of the library's pinned result slab (created on first access); they stay valid until the next
        call_regions_view / release_view / close on this context."""
        self.release_view()
        p, keep = self._region_params(bandwidth, region_threshold, kurtosis_threshold, corr_threshold, hit_threshold, directional, control,
                                      coeffs, background, reference_order)
        t = RegionTable()
        self._check(self._lib.upb_call_regions(self._h, C.byref(p), C.byref(t)))
        self._view_table = t
        return RegionView(t, self.n_contigs)

    def release_view(self):
        t = getattr(self, "_view_table", None)
        if t is not None:
            self._lib.upb_free_region_table(C.byref(t))
            self._view_table = None

    def call_regions(self, *, bandwidth=50, region_threshold=25.0, kurtosis_threshold=50.0, corr_threshold=-1.0,
                     hit_threshold=10.0, directional=True, control: Optional[Sequence[int]] = None,
                     coeffs: Optional[Sequence[float]] = None, background: float, reference_order=True, rows=True, **_ignored) -> Dict:
        p, keep = self._region_params(bandwidth, region_threshold, kurtosis_threshold, corr_threshold, hit_threshold, directional, control,
                                      coeffs, background, reference_order)
        t = RegionTable()
        self._check(self._lib.upb_call_regions(self._h, C.byref(p), C.byref(t)))
        try:
            return self._table_dict(t, self.n_contigs, rows)
        finally:
            self._lib.upb_free_region_table(C.byref(t))

    def profile_contig(self, buffer: int, contig: int, *, bandwidth=50, region_threshold=25.0, directional=True,
                       control: Optional[Sequence[int]] = None, background: float, **_ignored):
        ctrl = np.ascontiguousarray(control if control is not None else np.zeros(self.n_samples, np.uint8), dtype=np.uint8)
        p = RegionParams(bandwidth, region_threshold, 0.0, -1.0, 0.0, 1 if directional else 0, ctrl.ctypes.data_as(C.POINTER(C.c_uint8)),
                         0, None, background, 0)
        pr = Profile()
        self._check(self._lib.upb_profile_contig(self._h, C.byref(p), buffer, contig, C.byref(pr)))
        try:
            return _np(pr.pos, int(pr.n), np.uint32), _np(pr.score, int(pr.n), np.float64)
        finally:
            self._lib.upb_free_profile(C.byref(pr))

    def strand_shift(self, *, bandwidth=50, region_threshold=25.0, kurtosis_threshold=50.0, corr_threshold=0.3, hit_threshold=10,
                     min_shift=25, max_shift=150, n_regions_to_test=1000, background: float) -> Dict:
        p = ShiftParams(bandwidth, region_threshold, kurtosis_threshold, corr_threshold, hit_threshold, min_shift, max_shift,
                        n_regions_to_test, background)
        r = ShiftReport()
        rc = self._lib.upb_strand_shift(self._h, C.byref(p), C.byref(r))
        try:
            out = self._shift_dict(rc, r)
        finally:
            self._lib.upb_free_shift_report(C.byref(r))
        if rc and rc != -8:
            self._check(rc)
        return out

    @staticmethod
    def _shift_params(bandwidth=50, region_threshold=25.0, kurtosis_threshold=50.0, corr_threshold=0.3, hit_threshold=10,
                      min_shift=25, max_shift=150, n_regions_to_test=1000, *, background: float) -> "ShiftParams":
        return ShiftParams(bandwidth, region_threshold, kurtosis_threshold, corr_threshold, hit_threshold, min_shift, max_shift,
                           n_regions_to_test, background)

    def shift_regions(self, **params) -> Dict[str, np.ndarray]:
        """upb_shift_regions: the regions of the contigs loaded on this context, in the reference's order (contig, left)."""
        p = self._shift_params(**params)
        l = ShiftRegionList()
        self._check(self._lib.upb_shift_regions(self._h, C.byref(p), C.byref(l)))
        try:
            n = int(l.n)
            return dict(contig=_np(l.contig, n, np.uint32), left=_np(l.left, n, np.uint32), len=_np(l.len, n, np.uint32),
                        sum=_np(l.sum, n, np.uint32), replayed=_np(l.replayed, n, np.uint8))
        finally:
            self._lib.upb_free_shift_regions(C.byref(l))

    def shift_sweep(self, contig: np.ndarray, left: np.ndarray, length: np.ndarray, **params):
        """upb_shift_sweep: (best shift, best correlation) of every given region (src/strand_shift.cpp:206-219)."""
        p = self._shift_params(**params)
        contig = np.ascontiguousarray(contig, dtype=np.uint32); left = np.ascontiguousarray(left, dtype=np.uint32)
        length = np.ascontiguousarray(length, dtype=np.uint32)
        n = int(contig.shape[0])
        bs = np.zeros(n, dtype=np.uint32); bc = np.zeros(n, dtype=np.float64)
        if n:
            self._check(self._lib.upb_shift_sweep(self._h, C.byref(p), n, contig.ctypes.data, left.ctypes.data, length.ctypes.data,
                                                  bs.ctypes.data, bc.ctypes.data))
        return bs, bc

    def count_tags_in_regions(self, keys: Sequence[Sequence[int]], *, directional=True, extension=0) -> np.ndarray:
        arr = (RegionKey * max(1, len(keys)))()
        for i, (c, s, e) in enumerate(keys):
            arr[i] = RegionKey(c, s, e)
        out = np.zeros(max(1, len(keys) * self.n_samples), dtype=np.uint32)
        self._check(self._lib.upb_count_tags_in_regions(self._h, 1 if directional else 0, extension, len(keys), arr, out.ctypes.data))
        return out[:len(keys) * self.n_samples].reshape(len(keys), self.n_samples)

    def count_tags_in_regions_arr(self, keys: np.ndarray, *, directional=True, extension=0) -> np.ndarray:
        """Same, keys as a contiguous (n, 3) uint32 array of (contig, start, end) — no per-row Python work."""
        keys = np.ascontiguousarray(keys, dtype=np.uint32).reshape(-1, 3)
        out = np.zeros(max(1, keys.shape[0] * self.n_samples), dtype=np.uint32)
        self._check(self._lib.upb_count_tags_in_regions(self._h, 1 if directional else 0, extension, keys.shape[0], C.cast(keys.ctypes.data, C.POINTER(RegionKey)), out.ctypes.data))
        return out[:keys.shape[0] * self.n_samples].reshape(keys.shape[0], self.n_samples)

    def call_regions_shared(self, table: "SharedTable", step: int, *, bandwidth=50, region_threshold=25.0, kurtosis_threshold=50.0, corr_threshold=-1.0,
                            hit_threshold=10.0, directional=True, control: Optional[Sequence[int]] = None, coeffs: Optional[Sequence[float]] = None,
                            background: float, reference_order=True, **_ignored) -> "RegionView":
        """upb_call_regions_shared: this rank's part of one table-wide call; on rank 0 the view is the whole table (columns inside the
        shared segment), on the other ranks only n_rows (their own rows) is meaningful."""
        self.release_view()
        p, keep = self._region_params(bandwidth, region_threshold, kurtosis_threshold, corr_threshold, hit_threshold, directional, control,
                                      coeffs, background, reference_order)
        t = RegionTable()
        self._check(self._lib.upb_call_regions_shared(self._h, C.byref(p), C.byref(table), int(step), C.byref(t)))
        return RegionView(t, self.n_contigs)

    def set_profiling(self, on: bool) -> None:
        self._check(self._lib.upb_set_profiling(self._h, 1 if on else 0))

    def stage_times(self) -> List:
        """[(stage name, device ms)] of the last region-calling call (set_profiling(True) first)."""
        n = C.c_uint32(0)
        names = (C.c_char_p * 32)()
        ms = (C.c_float * 32)()
        self._check(self._lib.upb_last_stage_times(self._h, 32, names, ms, C.byref(n)))
        return [(names[k].decode(), float(ms[k])) for k in range(min(int(n.value), 32))]

    def stream_ptr(self) -> int:
        """cudaStream_t of this context (for CUDA-event timing by the caller)."""
        return int(self._lib.upb_stream(self._h) or 0)

    def stats(self) -> Dict:
        s = Stats()
        self._check(self._lib.upb_last_stats(self._h, C.byref(s)))
        return dict(ms_total=float(s.ms_total), ms_kde=float(s.ms_kde), launches=int(s.launches), kde_tiles=int(s.kde_tiles),
                    kde_hot_tiles=int(s.kde_hot_tiles), n_candidates=int(s.n_candidates), n_entries=int(s.n_entries),
                    hot_blocks=int(s.hot_blocks), exact_evals=int(s.exact_evals), spans_with_tags=int(s.spans_with_tags),
                    n_stitched=int(s.n_stitched), warm_spans=int(s.warm_spans), ms_call=float(s.ms_call),
                    host_ms=[float(s.ms_host_prepare), float(s.ms_host_enqueue), float(s.ms_host_wait), float(s.ms_host_post), float(s.ms_host_assemble)])


class MultiContext:
    """upb_multi: the GPUs of one box behind one handle (one context and one worker thread per GPU inside the library).  Same calls and
    same results as Context; contigs are spread over the GPUs for regions / strand_shift, samples for tags_in_regions."""

    def __init__(self, n_gpus: int, devices: Optional[Sequence[int]] = None):
        self._lib = lib()
        self._h = C.c_void_p()
        dev = None if devices is None else np.ascontiguousarray(devices, dtype=np.int32)
        rc = self._lib.upb_multi_create(n_gpus, None if dev is None else dev.ctypes.data, C.byref(self._h))
        if rc:
            raise UpbError(rc, "upb_multi_create(%d) failed" % n_gpus)
        self.n_gpus, self.n_contigs, self.n_samples = n_gpus, 0, 0
        self._hold = None

    def close(self):
        if self._h:
            self._lib.upb_multi_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int):
        if rc:
            raise UpbError(rc, self._lib.upb_multi_last_error(self._h).decode())

    def set_contigs(self, sizes: Sequence[int]):
        arr = np.ascontiguousarray(sizes, dtype=np.uint32)
        self._check(self._lib.upb_multi_set_contigs(self._h, arr.shape[0], arr.ctypes.data))
        self.n_contigs = int(arr.shape[0])

    def owner(self) -> np.ndarray:
        out = np.zeros(self.n_contigs, dtype=np.uint32)
        self._check(self._lib.upb_multi_owner(self._h, out.ctypes.data))
        return out

    def load(self, data, packed: bool = False) -> None:
        self.set_contigs(data.contig_sizes)
        off = np.ascontiguousarray(data.off, dtype=np.uint64)
        if packed:
            words, exc = pack_tags2(data.n_samples, self.n_contigs, off, data.pos, data.cnt)
            self._hold = (off, words, exc)
            src = TagSource(data.n_samples, off.ctypes.data, None, None, words.ctypes.data, exc.shape[0], exc.ctypes.data, None, 1)
        else:
            pos = np.ascontiguousarray(data.pos, dtype=np.uint32); cnt = np.ascontiguousarray(data.cnt, dtype=np.uint32)
            self._hold = (off, pos, cnt)
            src = TagSource(data.n_samples, off.ctypes.data, pos.ctypes.data, cnt.ctypes.data, None, 0, None, None, 1)
        self._check(self._lib.upb_multi_load_tags(self._h, C.byref(src)))
        self._check(self._lib.upb_multi_load_wait(self._h))
        self._hold = None
        self.n_samples = data.n_samples

    def load_by_sample(self, data) -> None:
        self.set_contigs(data.contig_sizes)
        off = np.ascontiguousarray(data.off, dtype=np.uint64)
        pos = np.ascontiguousarray(data.pos, dtype=np.uint32); cnt = np.ascontiguousarray(data.cnt, dtype=np.uint32)
        self._check(self._lib.upb_multi_load_tags_by_sample(self._h, data.n_samples, off.ctypes.data, pos.ctypes.data, cnt.ctypes.data))
        self.n_samples = data.n_samples

    def call_regions(self, *, bandwidth=50, region_threshold=25.0, kurtosis_threshold=50.0, corr_threshold=-1.0, hit_threshold=10.0,
                     directional=True, control: Optional[Sequence[int]] = None, coeffs: Optional[Sequence[float]] = None,
                     background: float, reference_order=True, rows=True, **_ignored) -> Dict:
        ctrl = np.ascontiguousarray(control if control is not None else np.zeros(self.n_samples, np.uint8), dtype=np.uint8)
        cf = np.ascontiguousarray(coeffs if coeffs is not None else [], dtype=np.float64)
        p = RegionParams(bandwidth, region_threshold, kurtosis_threshold, corr_threshold, hit_threshold, 1 if directional else 0,
                         ctrl.ctypes.data_as(C.POINTER(C.c_uint8)), cf.shape[0], cf.ctypes.data_as(C.POINTER(C.c_double)),
                         background, 1 if reference_order else 0)
        t = RegionTable()
        self._check(self._lib.upb_multi_call_regions(self._h, C.byref(p), C.byref(t)))
        return Context._table_dict(t, self.n_contigs, rows)

    def profile_contig(self, buffer: int, contig: int, *, bandwidth=50, region_threshold=25.0, directional=True,
                       control: Optional[Sequence[int]] = None, background: float, **_ignored):
        ctrl = np.ascontiguousarray(control if control is not None else np.zeros(self.n_samples, np.uint8), dtype=np.uint8)
        p = RegionParams(bandwidth, region_threshold, 0.0, -1.0, 0.0, 1 if directional else 0, ctrl.ctypes.data_as(C.POINTER(C.c_uint8)),
                         0, None, background, 0)
        pr = Profile()
        self._check(self._lib.upb_multi_profile_contig(self._h, C.byref(p), buffer, contig, C.byref(pr)))
        try:
            return _np(pr.pos, int(pr.n), np.uint32), _np(pr.score, int(pr.n), np.float64)
        finally:
            self._lib.upb_free_profile(C.byref(pr))

    def strand_shift(self, *, bandwidth=50, region_threshold=25.0, kurtosis_threshold=50.0, corr_threshold=0.3, hit_threshold=10,
                     min_shift=25, max_shift=150, n_regions_to_test=1000, background: float) -> Dict:
        p = ShiftParams(bandwidth, region_threshold, kurtosis_threshold, corr_threshold, hit_threshold, min_shift, max_shift,
                        n_regions_to_test, background)
        r = ShiftReport()
        rc = self._lib.upb_multi_strand_shift(self._h, C.byref(p), C.byref(r))
        try:
            out = Context._shift_dict(rc, r)
        finally:
            self._lib.upb_free_shift_report(C.byref(r))
        if rc and rc != -8:
            self._check(rc)
        return out

    def count_tags_in_regions(self, keys: Sequence[Sequence[int]], *, directional=True, extension=0) -> np.ndarray:
        keys = np.ascontiguousarray(keys, dtype=np.uint32).reshape(-1, 3)
        out = np.zeros(max(1, keys.shape[0] * self.n_samples), dtype=np.uint32)
        self._check(self._lib.upb_multi_count_tags_in_regions(self._h, 1 if directional else 0, extension, keys.shape[0],
                                                              C.cast(keys.ctypes.data, C.POINTER(RegionKey)), out.ctypes.data))
        return out[:keys.shape[0] * self.n_samples].reshape(keys.shape[0], self.n_samples)
