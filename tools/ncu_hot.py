#!/usr/bin/env python
"""Per-CUDA-line hot spots of one kernel from an .ncu-rep (needs -lineinfo and --import-source on).
usage: python tools/ncu_hot.py report.ncu-rep [top]"""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur = None; data = []
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1].split('/')[-1]; continue
    if r[0] == "Line No":
        hdr = r; iN = hdr.index('# Samples'); iI = hdr.index('Instructions Executed'); iL = hdr.index('stall_long_sb'); iS = hdr.index('stall_short_sb'); continue
    if r[0].isdigit():
        try: data.append((int(r[iI] or 0), int(r[iN] or 0), int(r[iL] or 0), int(r[iS] or 0), cur, int(r[0]), r[1]))
        except Exception: pass
tot = sum(d[0] for d in data); tots = sum(d[1] for d in data)
print("total warp instructions %d, samples %d" % (tot, tots))
print("--- by instructions")
for d in sorted(data, reverse=True)[:top]:
    print("%5.1f%% inst %5.1f%% smp (long %5d short %5d) %s:%d | %s" % (100*d[0]/tot, 100*d[1]/tots, d[2], d[3], d[4], d[5], d[6].strip()[:105]))
print("--- by samples")
for d in sorted(data, key=lambda x: -x[1])[:top//2]:
    print("%5.1f%% inst %5.1f%% smp (long %5d short %5d) %s:%d | %s" % (100*d[0]/tot, 100*d[1]/tots, d[2], d[3], d[4], d[5], d[6].strip()[:105]))
