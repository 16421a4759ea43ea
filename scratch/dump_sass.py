#!/usr/bin/env python
"""SASS of the hot kernels (cuobjdump -sass of the built object), one file per kernel, trimmed to the instruction text
(no encodings), gzip'd, plus an index with the mnemonic counts that show what each kernel is made of:
UBLKCP = cp.async.bulk (TMA engine) issue, SYNCS = mbarrier ops, IDP = dp2a, REDUX = warp reduction, ATOMS = shared atomics."""
import collections, gzip, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
obj = os.path.join(ROOT, "ako_b200", "csrc", "build", "ako_device.o")
out_dir = os.path.join(ROOT, "profiles", "sass")
WANT = {
    "_Z12k_lift_stripILi0ELi2EEv11StripParams": "k_lift_strip_dd137_gate",
    "_Z12k_lift_stripILi0ELi0EEv11StripParams": "k_lift_strip_dd137_plain",
    "_Z12k_lift_stripILi1ELi0EEv11StripParams": "k_lift_strip_cdf53_plain",
    "_Z14k_unlift_stripILi0EEv13UnstripParams": "k_unlift_strip_dd137",
    "_Z14k_unlift_stripILi1EEv13UnstripParams": "k_unlift_strip_cdf53",
    "_Z13k_lift_strip4ILi0ELi2EEv12Strip4Params": "k_lift_strip4_dd137_gate",
    "_Z12k_kg_lengthsPKsmmPKjPjjS3_S2_PKhj": "k_kg_lengths",
    "_Z11k_kg_startsPKsmmPjPhj": "k_kg_starts",
    "_Z9k_kg_packPKsmmPKjPKmS2_jPhmmS2_": "k_kg_pack",
    "_Z11k_kd_decodePKhPKmS2_jPK10KdSubStateS2_P6KfLookPmPjP7KdImagePsmmP5KtRunS9_j": "k_kd_decode",
    "_Z9k_kd_syncPKhPKmS2_jiS2_PmP10KdSubStateP7KdImage": "k_kd_sync",
    "_Z20k_format_fwd_rgba8x8PKhPsjjmiimm8FmtTiles": "k_format_fwd_rgba8x8",
    "_Z20k_format_inv_rgba8x8PKsPhjjmimm8FmtTiles": "k_format_inv_rgba8x8",
}
text = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
cur, funcs = None, collections.OrderedDict()
for line in text.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        funcs[cur] = []
        continue
    if cur is not None:
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);\s*/\*", line)
        if m:
            funcs[cur].append(f"/*{m.group(1)}*/ {m.group(2).strip()} ;")
index = ["# SASS of the hot kernels (sm_100a), from `cuobjdump -sass ako_b200/csrc/build/ako_device.o`", "",
         "Regenerate with `python scratch/dump_sass.py` after `make -C ako_b200/csrc`. One gzip'd listing per kernel;",
         "the table counts the mnemonics that identify the mechanisms DESIGN.md names.", "",
         "| kernel | instructions | UBLKCP (TMA bulk copy) | SYNCS (mbarrier) | IDP (dp2a) | IMAD | PRMT | REDUX | ATOMS | BAR | LDS | STS | LDG | STG |",
         "|---|---|---|---|---|---|---|---|---|---|---|---|---|---|"]
for mangled, name in WANT.items():
    ins = funcs.get(mangled)
    if not ins:
        print("missing", mangled, file=sys.stderr)
        continue
    with gzip.open(os.path.join(out_dir, name + ".sass.gz"), "wt") as f:
        f.write("\n".join(ins) + "\n")
    ops = collections.Counter()
    for i in ins:
        t = i.split("*/", 1)[1].split()
        op = t[1] if t[0].startswith("@") else t[0]
        ops[op.split(".")[0]] += 1
    index.append(f"| `{name}` | {len(ins)} | " + " | ".join(str(ops.get(k, 0)) for k in
                 ("UBLKCP", "SYNCS", "IDP", "IMAD", "PRMT", "REDUX", "ATOMS", "BAR", "LDS", "STS", "LDG", "STG")) + " |")
open(os.path.join(out_dir, "README.md"), "w").write("\n".join(index) + "\n")
print("\n".join(index[5:]))
