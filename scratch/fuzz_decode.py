"""Corrupted .ako blobs: akoDecodeExt of the CUDA library against the reference (unmodified, oracle/_ref) and the
oracle on the SAME corrupted bytes. Reports disagreements in status or pixels; never asserts (exploration)."""
import sys, os, json
R = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import numpy as np
import ako_b200, oracle_lib as ol
orc = ol.load_oracle(); ref = ol.load_ref()
rs = np.random.RandomState(123)
stats = {"cases": 0, "status_mismatch_orc": 0, "pixel_mismatch_orc": 0, "status_mismatch_ref": 0, "pixel_mismatch_ref": 0}
examples = []
for (w, h, kw) in [(96, 80, dict(wavelet=0, q=16, g=0)), (200, 131, dict(wavelet=1, q=0, g=0)), (64, 64, dict(wavelet=2, q=8, g=4, tiles=32))]:
    img = ol.synth(orc, w, h, w)
    blob, _ = ol.orc_encode(orc, img, **kw)
    for trial in range(60):
        b = bytearray(blob)
        mode = trial % 4
        if mode == 0:      # flip a few body bytes
            for _ in range(rs.randint(1, 4)):
                b[rs.randint(16, len(b))] ^= 1 << rs.randint(0, 8)
        elif mode == 1:    # truncate
            b = b[:rs.randint(16, len(b))]
        elif mode == 2:    # corrupt a block size field (first block head at 16)
            b[16 + rs.randint(0, 4)] ^= 1 << rs.randint(0, 8)
        else:              # append garbage / zero a span
            a = rs.randint(20, len(b) - 4); b[a:a + rs.randint(1, 16)] = bytes(rs.randint(1, 16))
        b = bytes(b)
        got, gst, _ = ako_b200.decode(b)
        wo, wst = ol.orc_decode(orc, b)
        stats["cases"] += 1
        if gst != wst:
            stats["status_mismatch_orc"] += 1
            if len(examples) < 6: examples.append(("orc", w, h, trial, mode, gst, wst))
        elif got is not None and not np.array_equal(got, wo):
            stats["pixel_mismatch_orc"] += 1
            if len(examples) < 6: examples.append(("orc-px", w, h, trial, mode))
        if ref is not None:
            wr, rst = ol.ref_decode(ref, b)
            if gst != rst:
                stats["status_mismatch_ref"] += 1
                if len(examples) < 12: examples.append(("ref", w, h, trial, mode, gst, rst))
            elif got is not None and not np.array_equal(got, wr):
                stats["pixel_mismatch_ref"] += 1
                if len(examples) < 12: examples.append(("ref-px", w, h, trial, mode))
print(json.dumps(stats)); print(examples)
