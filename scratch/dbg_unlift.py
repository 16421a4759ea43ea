import os, sys, ctypes as C
import numpy as np
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import ako_b200, oracle_lib as ol
from cases import *
P, u8p, i16p = ol._p, ol.u8p, ol.i16p
orc = ol.load_oracle()
ctx = ako_b200.Context(0)
w, h, ch = [int(x) for x in sys.argv[1:4]]
wavelet = int(sys.argv[4]) if len(sys.argv) > 4 else 0
rs = np.random.RandomState(1)
planes = rs.randint(-300, 600, size=(ch, h, w)).astype(np.int16)
n = orc.orc_tile_data_size(w, h) * ch // 2
want = np.zeros(n, np.int16)
os_ = ol.make_settings(ol.OrcSettings, wavelet=wavelet, wrap=0, q=0, g=0)
orc.orc_lift(C.byref(os_), ch, w, h, P(planes.copy(), i16p), P(want, i16p))
s = ako_b200.default_settings(wavelet=wavelet, wrap=0, quantization=0, gate=0)
back = ctx.unlift(want, s, ch, w, h)
bad = back != planes
print("mismatches", int(bad.sum()), "of", bad.size)
for c in range(ch):
    b = bad[c]
    if b.any():
        ys, xs = np.nonzero(b)
        print("ch", c, "rows", ys.min(), ys.max(), "cols", xs.min(), xs.max(), "count", len(ys))
        print(" bad cols histogram (by 8):", np.bincount(xs // 8)[:40])
        print(" bad rows head:", np.unique(ys)[:20])
