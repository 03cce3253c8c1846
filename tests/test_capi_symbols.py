"""The C-ABI library loads and exports every symbol include/unipeak_b200.h declares; host-only entry points
(no device needed) behave; device entry points fail loudly without a GPU (there is no CPU path)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import unipeak_b200
from unipeak_b200 import capi
from refutil import oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "unipeak_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(upb_[a-z_0-9]+)\s*\(", text)))


def test_header_symbols_are_exported():
    lib = capi.lib()
    names = declared_symbols()
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), "libunipeak_b200.so does not export %s" % n
    assert sorted(capi.EXPORTS) == names, "capi.EXPORTS and the header disagree"


def test_library_is_in_tree():
    assert os.path.dirname(capi.LIB_PATH) == os.path.join(ROOT, "unipeak_b200")
    assert os.path.exists(capi.LIB_PATH)


@pytest.mark.parametrize("bw,total", [(50, 1.0), (5, 1.0), (200, 1 / 0.0775), (10, 123.456), (1, 1.0)])
def test_kernel_taps_match_oracle(bw, total):
    # Kernel::Kernel (misc/kernel.cpp:16-35): host arithmetic on both sides, must be bit-identical
    got = unipeak_b200.kernel_taps(bw, total)
    want = np.zeros(2 * bw + 1)
    oracle().upo_kernel(bw, total, want.ctypes.data_as(C.POINTER(C.c_double)))
    assert np.array_equal(got.view(np.uint64), want.view(np.uint64))
    assert got[0] == 0.0 and got[-1] == 0.0


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    assert capi.lib().upb_device_count() == 0
    with pytest.raises(unipeak_b200.UpbError) as ei:
        unipeak_b200.Context(0)
    assert ei.value.code == -3
