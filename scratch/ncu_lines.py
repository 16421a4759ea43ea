#!/usr/bin/env python
"""Dynamic warp-instructions per CUDA source line from an .ncu-rep (cuda,sass view)."""
import collections, csv, io, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cnt = collections.Counter(); samples = collections.Counter(); text = {}
fname = None; h = None
for r in rows:
    if not r: continue
    if r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if r[0] == "Line No": h = r; iE = h.index("Instructions Executed"); iSm = h.index("# Samples"); continue
    if h is None or len(r) <= iE: continue
    if r[0].isdigit():
        cur = (fname, int(r[0])); text[cur] = r[1]
    if r[2] and r[iE].isdigit():  # a sass row
        cnt[cur] += int(r[iE]); samples[cur] += int(r[iSm]) if r[iSm].isdigit() else 0
tot = sum(cnt.values()); ts = sum(samples.values()) or 1
print("total", tot)
for k in sorted(cnt):
    if cnt[k] * 200 >= tot or samples[k] * 100 >= ts:
        print(f"{k[0]}:{k[1]:4d} {100*cnt[k]/tot:5.1f}% inst {100*samples[k]/ts:5.1f}% samples | {text.get(k,'')[:110]}")
