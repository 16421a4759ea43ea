import os, sys, ctypes as C, subprocess
import numpy as np
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import ako_b200, oracle_lib as ol
orc = ol.load_oracle()
ctx = ako_b200.Context()
w, h, ch = int(sys.argv[1]), int(sys.argv[2]), 1
wl = int(sys.argv[3])
rs = np.random.RandomState(1)
planes = rs.randint(-300, 600, size=(ch, h, w)).astype(np.int16)
n = orc.orc_tile_data_size(w, h) * ch // 2
want = np.zeros(n, np.int16); tmp = planes.copy()
os_ = ol.make_settings(ol.OrcSettings, wavelet=wl, q=0, g=0)
orc.orc_lift(C.byref(os_), ch, w, h, ol._p(tmp, ol.i16p), ol._p(want, ol.i16p))
got = ctx.lift(planes, ako_b200.default_settings(wavelet=wl, q=0, g=0))
bad = np.nonzero(want != got)[0]
print("mismatches", len(bad), "of", n)
# locate: finest level block is at the end: [q][C][B][D] each tw*th
tw, th = (w + 1) // 2, (h + 1) // 2
band = tw * th
base = n - 3 * band
for name, off in (("C", base), ("B", base + band), ("D", base + 2 * band)):
    b = bad[(bad >= off) & (bad < off + band)] - off
    if len(b):
        rows, cols = b // tw, b % tw
        print(name, len(b), "rows", rows.min(), rows.max(), "cols", cols.min(), cols.max(), "first", rows[0], cols[0],
              "want", want[off + b[0]], "got", got[off + b[0]])
print("below finest:", len(bad[bad < base]))
