"""Small end-to-end + stage run for compute-sanitizer (memcheck): strip kernels at aligned and edge shapes, the
small-pyramid kernels, Kagari encode/decode incl. long runs, tiles, batch API."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import ako_b200, oracle_lib as ol
orc = ol.load_oracle()
S = ako_b200.default_settings
for (w, h) in [(528, 70), (1040, 264), (408, 616), (264, 40), (130, 67)]:
    img = ol.synth(orc, w, h, w)
    for wavelet, q, g in ((0, 16, 16), (1, 0, 0), (2, 7, 0)):
        blob, st = ako_b200.encode(img, S(wavelet=wavelet, quantization=q, gate=g))
        want, _ = ol.orc_encode(orc, img, wavelet=wavelet, q=q, g=g)
        assert st == 0 and blob == want, (w, h, wavelet)
        px, st, _ = ako_b200.decode(blob)
        assert st == 0 and np.array_equal(px, ol.orc_decode(orc, want)[0])
img = ol.synth(orc, 200, 150, 9)
for tiles in (8, 64):
    blob, st = ako_b200.encode(img, S(wavelet=0, quantization=5, tiles_dimension=tiles))
    assert st == 0 and blob == ol.orc_encode(orc, img, wavelet=0, q=5, tiles=tiles)[0]
    px, st, _ = ako_b200.decode(blob)
    assert st == 0
ctx = ako_b200.Context()
v = np.concatenate([np.zeros(70000, np.int16), np.arange(-50, 50).astype(np.int16), np.full(9000, 3, np.int16)])
enc = ctx.kagari_encode(v, len(v) * 4 + 64)
used, back = ctx.kagari_decode(enc, len(v))
assert used == len(enc) and np.array_equal(back, v)
ctx.close()
print("sanitize_small ok")
