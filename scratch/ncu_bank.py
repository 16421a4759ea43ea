#!/usr/bin/env python
"""Shared-memory wavefronts (actual vs ideal) per CUDA source line from an .ncu-rep."""
import collections, csv, io, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
wf = collections.Counter(); ideal = collections.Counter(); text = {}
fname = None; h = None
for r in rows:
    if not r: continue
    if r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if r[0] == "Line No": h = r; iW = h.index("L1 Wavefronts Shared"); iI = h.index("L1 Wavefronts Shared Ideal"); continue
    if h is None or len(r) <= iI: continue
    if r[0].isdigit():
        cur = (fname, int(r[0])); text[cur] = r[1]
    if r[2] and r[iW].isdigit():
        wf[cur] += int(r[iW]); ideal[cur] += int(r[iI]) if r[iI].isdigit() else 0
tot = sum(wf.values()) or 1
print("total wavefronts", tot, "ideal", sum(ideal.values()))
for k in sorted(wf):
    if wf[k] * 100 >= tot:
        print(f"{k[0]}:{k[1]:4d} wf {wf[k]:10d} ideal {ideal[k]:10d} x{wf[k]/max(ideal[k],1):.2f} | {text.get(k,'')[:100]}")
