"""Randomised soak: GPU (akoEncodeExt / akoDecodeExt and the batch entry points) against the oracle over random shapes
and settings, including corrupted blobs. python scratch/soak.py SECONDS [SEED]"""
import os, sys, time
import numpy as np
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import ako_b200, oracle_lib as ol
from cases import *
budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
orc = ol.load_oracle()
rs = np.random.RandomState(seed)
S = ako_b200.default_settings
t0 = time.time(); n = 0; bad = 0; enc_fail = 0; corrupt_ok = 0
while time.time() - t0 < budget:
    w = int(rs.choice([9, 40, 63, 64, 65, 100, 129, 200, 333, 520, 777, 1031, 1921, 2050])) + int(rs.randint(0, 9))
    h = int(rs.choice([8, 17, 40, 64, 131, 264, 601])) + int(rs.randint(0, 5))
    ch = int(rs.choice([1, 3, 4, 4, 4, 2]))
    if w * h * ch > 2_500_000:
        h = max(8, 2_500_000 // (w * ch))
    kind = rs.randint(0, 4)
    if kind == 0: img = noise_image(w, h, ch, n)
    elif kind == 1: img = smooth_image(w, h, ch, n)
    elif kind == 2:
        img = np.full((h, w, ch), int(rs.randint(0, 256)), np.uint8); img[rs.randint(0, h), rs.randint(0, w)] ^= 0x55
    else:
        img = np.ascontiguousarray(ol.synth(orc, w, h, n)[:, :, :min(ch, 4)]); ch = img.shape[2]
    lossless = rs.rand() < 0.3
    kw = dict(wavelet=int(rs.choice([0, 1, 2])), wrap=int(rs.choice([0, 0, 0, 1, 2, 3])), color=int(rs.choice([0, 1, 2])),
              q=0 if lossless else int(rs.choice([1, 3, 16, 60, 400])), g=0 if lossless else int(rs.choice([0, 0, 4, 16, 90])),
              chroma_loss=int(rs.randint(0, 4)), discard=int(rs.randint(0, 2)))
    tiles = int(rs.choice([0, 0, 0, 8, 64, 128, 256]))
    if tiles and not (0 < w % tiles < 3 or 0 < h % tiles < 3):
        kw["tiles"] = tiles
    want, wst = ol.orc_encode(orc, img, **kw)
    alias = {"q": "quantization", "g": "gate", "tiles": "tiles_dimension", "discard": "discard_non_visible"}
    s = S(**{alias.get(k, k): v for k, v in kw.items()})
    if rs.rand() < 0.3:
        k = int(rs.randint(2, 6))
        blobs, st, done = ako_b200.encode_batch([img] * k, s)
        got, gst = (blobs[0], st) if done == k else (None, st)
        if done == k and any(b != blobs[0] for b in blobs): bad += 1; print("BATCH MISMATCH", w, h, ch, kw)
    else:
        got, gst = ako_b200.encode(img, s)
    n += 1
    if want is None:
        enc_fail += 1
        if got is not None or gst != wst: bad += 1; print("STATUS", w, h, ch, kw, gst, wst)
        continue
    if got != want:
        bad += 1; print("ENCODE MISMATCH", w, h, ch, kw); continue
    want_px, _ = ol.orc_decode(orc, want)
    px, st, _ = ako_b200.decode(want)
    if st != 0 or not np.array_equal(px, want_px):
        bad += 1; print("DECODE MISMATCH", w, h, ch, kw, st); continue
    if rs.rand() < 0.5:
        b = bytearray(want)
        for _ in range(int(rs.randint(1, 4))):
            p = int(rs.randint(16, len(b))); b[p] ^= 1 << int(rs.randint(0, 8))
        if rs.rand() < 0.3: b = b[:len(b) - int(rs.randint(1, 9))]
        b = bytes(b)
        cpx, cst = ol.orc_decode(orc, b)
        gpx, gst, _ = ako_b200.decode(b)
        if gst != cst or (cpx is not None and not np.array_equal(gpx, cpx)):
            bad += 1; print("CORRUPT MISMATCH", w, h, ch, kw, gst, cst)
        corrupt_ok += cpx is not None
print(f"soak: {n} cases in {time.time()-t0:.0f} s, {enc_fail} encodes refused alike, {corrupt_ok} corrupted blobs still accepted alike, {bad} MISMATCHES")
sys.exit(1 if bad else 0)
