"""Randomised soak of the strip + frame path: lift / unlift of random planes with MIRROR / REPEAT / ZERO against the
oracle, sizes drawn around the thin-last-tile cases. python scratch/soak_wrap.py SECONDS [SEED]"""
import os, sys, time, ctypes as C
import numpy as np
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import ako_b200, oracle_lib as ol
budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
orc = ol.load_oracle(); ctx = ako_b200.Context(0)
rs = np.random.RandomState(seed)
i16p = C.POINTER(C.c_int16)
P = lambda a, t: a.ctypes.data_as(t)
t0 = time.time(); n = 0; bad = 0
while time.time() - t0 < budget:
    # coefficient sizes near multiples of the 64 x 32 tile (+0..7) and anything else
    tw = int(rs.choice([64 * int(rs.randint(5, 20)) + int(rs.randint(0, 8)), int(rs.randint(300, 1300))]))
    th = int(rs.choice([32 * int(rs.randint(5, 24)) + int(rs.randint(0, 8)), int(rs.randint(150, 800))]))
    w = 2 * tw - int(rs.randint(0, 2)); h = 2 * th - int(rs.randint(0, 2))
    ch = int(rs.choice([1, 1, 2, 3]))
    wavelet = int(rs.choice([0, 1])); wrap = int(rs.choice([1, 2, 3]))
    q, g = [(0, 0), (7, 9), (16, 0), (3, 40)][int(rs.randint(0, 4))]
    planes = rs.randint(-300, 600, size=(ch, h, w)).astype(np.int16)
    nvals = orc.orc_tile_data_size(w, h) * ch // 2
    want = np.zeros(nvals, np.int16); tmp = planes.copy()
    os_ = ol.make_settings(ol.OrcSettings, wavelet=wavelet, wrap=wrap, q=q, g=g)
    orc.orc_lift(C.byref(os_), ch, w, h, P(tmp, i16p), P(want, i16p))
    s = ako_b200.default_settings(wavelet=wavelet, wrap=wrap, quantization=q, gate=g)
    got = ctx.lift(planes, s)
    n += 1
    if not np.array_equal(want, got):
        bad += 1; print("LIFT MISMATCH", w, h, ch, wavelet, wrap, q, g, int(np.argmax(want != got))); continue
    back_want = np.zeros((ch, h, w), np.int16); st = want.copy()
    orc.orc_unlift(C.byref(os_), ch, w, h, P(st, i16p), P(back_want, i16p))
    back = ctx.unlift(want, s, ch, w, h)
    if not np.array_equal(back_want, back):
        bad += 1; print("UNLIFT MISMATCH", w, h, ch, wavelet, wrap, q, g)
print(f"soak_wrap: {n} cases in {time.time()-t0:.0f} s, {bad} MISMATCHES")
sys.exit(1 if bad else 0)
