"""unipeak_b200 — B200-native implementation of UniPeak's region-calling hot path
(the work behind bin/regions, bin/strand_shift and bin/tags_in_regions).

The product is the CUDA library `libunipeak_b200.so` behind the C-ABI in include/unipeak_b200.h and the
C++ command-line front ends in unipeak_b200/host.  This Python package only binds that C-ABI (ctypes) for
the tests and the benchmark harness; it contains no compute path of its own.
"""
from .capi import Context, UpbError, kernel_taps, lib, pack_tags2, pack_tags_vbyte, shared_table_bytes, LIB_PATH  # noqa: F401
