import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..")); sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
import numpy as np
import unipeak_b200
from unipeak_b200.capi import MultiContext
from unipeak_b200.synth import make_synthetic
from refutil import region_settings
d = make_synthetic([400000, 250000, 150000, 90000, 60000], 3, 40000, directional=True, seed=41, peak_spacing=5000)
ctx = unipeak_b200.Context(0)
st = region_settings(d, control=[1], region_threshold=10.0)
ctx.load(d)
want = ctx.call_regions(**st)["columns"]["counts"]
print("single", want[:3].tolist(), want.sum(axis=0))
for packed in (False, True):
    for prime in (False, True):
        m = MultiContext(2, [0, 0])
        if prime:
            m.load(d, packed=packed)
            m.call_regions(**region_settings(d))
        m.load(d, packed=packed)
        for k in range(2):
            got = m.call_regions(**st)["columns"]["counts"]
            print("packed", packed, "primed", prime, "call", k, got[:3].tolist(), got.sum(axis=0), "same" if np.array_equal(got, want) else "DIFF")
        m.close()
# one context, keep_contig subset
c2 = unipeak_b200.Context(0)
c2.load_subset(d, [0, 2, 4], packed=False)
g = c2.call_regions(**dict(st, reference_order=False))["columns"]
print("subset plain", g["counts"][:3].tolist())
c2.load_subset(d, [0, 2, 4], packed=True)
g = c2.call_regions(**dict(st, reference_order=False))["columns"]
print("subset packed", g["counts"][:3].tolist())
