#!/usr/bin/env python
"""Where the e2e step goes: load alone, call alone, load + call, with the library's own timeline (UPB_TRACE=1) on stderr."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import unipeak_b200
from unipeak_b200.synth import HG19_SIZES
from unipeak_b200.synth_det import make_synthetic_det

n, tags = 16, int(os.environ.get("TAGS", 30_000_000))
data = make_synthetic_det(HG19_SIZES, n, tags, directional=True, seed=20261018, device="cuda:0")
G = int(sum(HG19_SIZES))
st = dict(bandwidth=50, region_threshold=25.0, kurtosis_threshold=50.0, corr_threshold=-1.0, hit_threshold=10.0 * n, directional=True,
          background=sum(data.tags(s) for s in range(n)) / G / 2)
E = data.pos.shape[0]
h_off = torch.from_numpy(data.off.view(np.int64)).pin_memory()
h_w = torch.empty(E, dtype=torch.int16).pin_memory()
_, exc = unipeak_b200.pack_tags2(n, 25, data.off, data.pos, data.cnt, out=h_w.numpy().view(np.uint16))
h_exc = torch.from_numpy(exc.view(np.uint8).copy()).pin_memory()
FORM = os.environ.get("FORM", "vbyte")
h_vb = torch.empty(3 * E, dtype=torch.uint8).pin_memory()
vb, vb_off, vb_tiles, exc_v = unipeak_b200.pack_tags_vbyte(n, 25, data.off, data.pos, data.cnt, out=h_vb.numpy())
h_excv = torch.from_numpy(exc_v.view(np.uint8).copy()).pin_memory()
ctx = unipeak_b200.Context(0)
ctx.set_contigs(HG19_SIZES)
def load():
    if FORM == "vbyte":
        ctx.load_async_ptr(n, h_off.data_ptr(), vbytes_ptr=h_vb.data_ptr(), vbyte_off_ptr=vb_off.ctypes.data, vtile_off_ptr=vb_tiles.ctypes.data,
                           n_vtiles=vb_tiles.shape[0] - 1, n_exc=exc_v.shape[0], exc_ptr=h_excv.data_ptr(), directional_hint=True)
    else:
        ctx.load_async_ptr(n, h_off.data_ptr(), words_ptr=h_w.data_ptr(), n_exc=exc.shape[0], exc_ptr=h_exc.data_ptr(), directional_hint=True)
print("host form:", FORM, "bytes", vb.shape[0] if FORM == "vbyte" else 2 * E)
load(); ctx.load_wait(); ctx.call_regions_view(**st); ctx.call_regions_view(**st)
def t(fn, reps=3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps * 1e3
print("load + wait        %.2f ms" % t(lambda: (load(), ctx.load_wait())))
print("call (resident)    %.2f ms" % t(lambda: ctx.call_regions_view(**st)))
print("load + call        %.2f ms" % t(lambda: (load(), ctx.call_regions_view(**st))))
t0 = time.perf_counter(); load(); t1 = time.perf_counter(); ctx.load_wait()
print("load_async returns after %.2f ms" % ((t1 - t0) * 1e3))
print("reindex            %.2f ms" % t(lambda: ctx.reindex()))
os.environ["UPB_TRACE"] = "1"; print("load + call (5x)  %.2f ms" % t(lambda: (load(), ctx.call_regions_view(**st)), 5))
load(); ctx.call_regions_view(**st)
