#!/usr/bin/env python
"""Generates tests/golden/golden.json from the UNMODIFIED reference (oracle/_ref/libako_ref.so,
compiled from /root/reference/library by oracle/Makefile). Run in the build container only:

    python tests/golden/make_golden.py

Each case stores the settings, the seeded input recipe, the reference's .ako blob (base64, small
cases only) and SHA-256 of blob and of the reference-decoded pixels.
"""
import base64
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_lib as ol  # noqa: E402
from cases import noise_image, smooth_image  # noqa: E402

CASES = []
for i, (w, h, ch) in enumerate([(64, 48, 4), (61, 35, 4), (40, 40, 3), (24, 33, 1), (19, 16, 2), (96, 80, 4)]):
    for wavelet in (0, 1, 2):
        for wrap in ((0, 1, 2, 3) if i < 2 else (0,)):
            for (q, g) in ((0, 0), (16, 0), (9, 9)):
                CASES.append(dict(kind="smooth", w=w, h=h, ch=ch, seed=10 + i, wavelet=wavelet, wrap=wrap, q=q, g=g,
                                  tiles=0, color=0, discard=0, chroma_loss=1))
for tiles in (8, 32):
    for wavelet in (0, 1, 2):
        CASES.append(dict(kind="synth", w=100, h=75, ch=4, seed=3, wavelet=wavelet, wrap=0, q=12, g=2, tiles=tiles,
                          color=0, discard=0, chroma_loss=1))
for color in (0, 1, 2):
    for discard in (0, 1):
        CASES.append(dict(kind="synth", w=160, h=130, ch=4, seed=4, wavelet=0, wrap=0, q=8, g=0, tiles=0,
                          color=color, discard=discard, chroma_loss=2))
CASES.append(dict(kind="noise", w=48, h=48, ch=4, seed=5, wavelet=1, wrap=0, q=0, g=0, tiles=0, color=0, discard=0,
                  chroma_loss=1))


def make_input(orc, c):
    if c["kind"] == "synth":
        return ol.synth(orc, c["w"], c["h"], c["seed"])
    if c["kind"] == "noise":
        return noise_image(c["w"], c["h"], c["ch"], c["seed"])
    return smooth_image(c["w"], c["h"], c["ch"], c["seed"])


def main():
    orc, ref = ol.load_oracle(), ol.load_ref()
    assert ref is not None, "needs /root/reference to build oracle/_ref"
    out = []
    for c in CASES:
        img = make_input(orc, c)
        blob, st = ol.ref_encode(ref, img, wavelet=c["wavelet"], wrap=c["wrap"], q=c["q"], g=c["g"],
                                 tiles=c["tiles"], color=c["color"], discard=c["discard"],
                                 chroma_loss=c["chroma_loss"])
        assert st == 0, (c, st)
        dec, st = ol.ref_decode(ref, blob)
        e = dict(c)
        e["input_sha256"] = hashlib.sha256(img.tobytes()).hexdigest()
        e["blob_sha256"] = hashlib.sha256(blob).hexdigest()
        e["blob_len"] = len(blob)
        e["decoded_sha256"] = hashlib.sha256(dec.tobytes()).hexdigest()
        if len(blob) <= 2500:
            e["blob_b64"] = base64.b64encode(blob).decode()
        out.append(e)
    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump(out, f, indent=0, sort_keys=True)
    print(len(out), "cases,", os.path.getsize(os.path.join(HERE, "golden.json")), "bytes")


if __name__ == "__main__":
    main()
