"""Which of bench.py's SHAPES passes the oracle gate (blob bytes, decoded pixels), via the batch device API."""
import os, sys
import numpy as np, torch
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import ako_b200, oracle_lib as ol
import bench
from ako_b200.synth import synth_rgba8_torch
orc = ol.load_oracle()
ctx = ako_b200.Context(0)
for name, (w, h, ch, wavelet, q, g, tiles, B) in bench.SHAPES.items():
    distinct = min(B, 4)
    base = synth_rgba8_torch(w, h, [40 + i for i in range(distinct)], device="cuda")[..., :ch].contiguous()
    base = base.repeat((B + distinct - 1) // distinct, 1, 1, 1)[:B].contiguous()
    s = ako_b200.default_settings(wavelet=wavelet, quantization=q, gate=g, tiles_dimension=tiles)
    dc = bench.DeviceCodec(torch, ako_b200, ctx, 0, w, h, ch, s, B, base)
    sizes = dc.step(0); ctx.sync()
    wants = {}
    for i in range(B):
        img = base[i].cpu().numpy()
        if i >= distinct:
            got = dc.blobs[i, :sizes[i]].cpu().numpy().tobytes()
            ok = got == wants[i % distinct][0] and np.array_equal(dc.out[i].cpu().numpy(), wants[i % distinct][1])
            if not ok:
                print(name, i, "MISMATCH vs image", i % distinct, "blob", got == wants[i % distinct][0], flush=True)
            continue
        want_blob, st = ol.orc_encode(orc, img, wavelet=wavelet, q=q, g=g, tiles=tiles)
        got = dc.blobs[i, :sizes[i]].cpu().numpy().tobytes()
        want_px, _ = ol.orc_decode(orc, want_blob)
        px_ok = np.array_equal(dc.out[i].cpu().numpy(), want_px)
        wants[i] = (want_blob, want_px)
        # the single-image host API on the same input
        blob2, st2 = ako_b200.encode(img, s)
        print(name, i, "blob", got == want_blob, len(got), len(want_blob), "pixels", px_ok, "host-api blob", blob2 == want_blob, flush=True)
        if got != want_blob:
            a = np.frombuffer(got, np.uint8); b = np.frombuffer(want_blob, np.uint8)
            m = min(len(a), len(b)); d = np.nonzero(a[:m] != b[:m])[0]
            print("   first diff at", d[:5], "of", m)
