"""Pins the oracle (oracle/ako_oracle.c) to the unmodified reference and to SURVEY Appendix B.

CPU only. If these fail the oracle is wrong and no GPU parity claim means anything.
"""
import ctypes as C
import hashlib
import itertools

import numpy as np
import pytest

import oracle_lib as ol
from cases import *

u8p, i16p = ol.u8p, ol.i16p
P = ol._p


def sha(b):
    return hashlib.sha256(b).hexdigest()


def test_synth_generator_matches_survey_hash(orc):
    img = ol.synth(orc, 1024, 1280, 1)
    assert sha(img.tobytes()) == "7e0d163cd4d552f8d11de3434628916df0009b37ae114b724b99892cca2c071f"


def test_numpy_synth_generator_matches_oracle(orc):
    """ako_b200/synth.py (what bench.py generates its inputs with) == the oracle's generator == SURVEY Appendix C."""
    from ako_b200.synth import synth_rgba8
    assert sha(synth_rgba8(1024, 1280, 1).tobytes()) == "7e0d163cd4d552f8d11de3434628916df0009b37ae114b724b99892cca2c071f"
    for (w, h, seed) in [(64, 48, 1), (333, 257, 7), (130, 700, 1003), (1, 1, 0), (129, 3, 4000000000)]:
        assert np.array_equal(synth_rgba8(w, h, seed), ol.synth(orc, w, h, seed)), (w, h, seed)


@pytest.mark.parametrize("kat", KATS[:2] + KATS[3:4], ids=lambda k: f"{k[0]}x{k[1]}-w{k[2]}-q{k[3]}-g{k[4]}")
def test_known_answers_oracle(orc, kat):
    w, h, wavelet, q, g, seed, size, blob_sha, dec_sha = kat
    img = ol.synth(orc, w, h, seed)
    blob, st = ol.orc_encode(orc, img, wavelet=wavelet, q=q, g=g)
    assert st == 0 and len(blob) == size and sha(blob) == blob_sha
    if dec_sha:
        out, st = ol.orc_decode(orc, blob)
        assert st == 0 and sha(out.tobytes()) == dec_sha


@pytest.mark.parametrize("kat", KATS, ids=lambda k: f"{k[0]}x{k[1]}-w{k[2]}-q{k[3]}-g{k[4]}")
def test_known_answers_reference(orc, ref, kat):
    """The compiled reference itself reproduces the survey's hashes (sanity of oracle/_ref)."""
    w, h, wavelet, q, g, seed, size, blob_sha, dec_sha = kat
    img = ol.synth(orc, w, h, seed)
    blob, st = ol.ref_encode(ref, img, wavelet=wavelet, q=q, g=g)
    assert st == 0 and len(blob) == size and sha(blob) == blob_sha
    if dec_sha:
        out, st = ol.ref_decode(ref, blob)
        assert sha(out.tobytes()) == dec_sha


def test_geometry_and_schedule(orc, ref):
    for w, h in [(1024, 1280), (1632, 2464), (1920, 1080), (8192, 8192), (3, 3), (5, 1000), (1021, 1277),
                 (16384, 16384), (9, 9), (17, 33)]:
        assert orc.orc_tile_data_size(w, h) == ref.akoTileDataSize(w, h)
        cw, ch = w, h
        while cw > 2 and ch > 2:
            for f, mul in itertools.product([1, 5, 16, 100, 1000, 0, -3], [1, 2, 3]):
                assert orc.orc_quantization(f, mul, w, h, cw, ch) == ref.akoQuantization(f, mul, w, h, cw, ch)
                assert orc.orc_gate(f, mul, w, h, cw, ch) == ref.akoGate(f, mul, w, h, cw, ch)
            cw, ch = (cw + 1) // 2, (ch + 1) // 2
    # Appendix B schedule rows
    assert orc.orc_quantization(16, 1, 1024, 1280, 1024, 1280) == 12
    assert orc.orc_quantization(16, 2, 1024, 1280, 1024, 1280) == 25
    assert orc.orc_quantization(16, 1, 8192, 8192, 8192, 8192) == 88
    assert orc.orc_gate(16, 2, 1632, 2464, 1632, 2464) == 43


@pytest.mark.parametrize("wrap", [0, 1, 2, 3])
@pytest.mark.parametrize("wavelet", [W_DD137, W_CDF53, W_HAAR])
def test_lift_1d_vs_reference_rows(orc, ref, wavelet, wrap):
    """orc_lift_1d == ako<W>LiftH on one row; orc_unlift_1d == ako<W>UnliftH; every wrap, odd/even lengths.
    Lengths follow the reference's own tests (tests/cdf53-test.c:238-271, dd137-test.c)."""
    rs = np.random.RandomState(7 + wavelet * 4 + wrap)
    lengths = [22, 16, 13, 17, 512, 150, 300, 31, 64] + ([10, 9, 8, 7, 6, 5, 4, 3] if wavelet != W_DD137 else [15])
    for n in lengths:
        for amp in (64, 2000, 32767):
            x = rs.randint(-amp, amp + 1, size=n).astype(np.int16)
            t = (n + 1) // 2
            fake = 2 * t - n
            lp, hp = np.zeros(t, np.int16), np.zeros(t, np.int16)
            orc.orc_lift_1d(wavelet, wrap, n, P(x, i16p), 1, P(lp, i16p), P(hp, i16p), 1)
            out = np.zeros(2 * t, np.int16)
            if wavelet == W_HAAR:
                ref.akoHaarLiftH(1, t, fake, n, x.ctypes.data, out.ctypes.data)
            elif wavelet == W_CDF53:
                ref.akoCdf53LiftH(wrap, 1, t, fake, n, x.ctypes.data, out.ctypes.data)
            else:
                ref.akoDd137LiftH(wrap, 1, t, fake, n, x.ctypes.data, out.ctypes.data)
            assert np.array_equal(out[:t], lp) and np.array_equal(out[t:], hp), (n, amp)

            # inverse on arbitrary (not necessarily lifted) coefficients
            lp2 = rs.randint(-amp, amp + 1, size=t).astype(np.int16)
            hp2 = rs.randint(-amp, amp + 1, size=t).astype(np.int16)
            mine = np.full(n, -1, np.int16)
            orc.orc_unlift_1d(wavelet, wrap, n, P(lp2, i16p), P(hp2, i16p), 1, P(mine, i16p), 1)
            theirs = np.full(2 * t, -1, np.int16)
            if wavelet == W_HAAR:
                ref.akoHaarUnliftH(t, 1, 2 * t, fake, lp2.ctypes.data, hp2.ctypes.data, theirs.ctypes.data)
            elif wavelet == W_CDF53:
                ref.akoCdf53UnliftH(wrap, t, 1, 2 * t, fake, lp2.ctypes.data, hp2.ctypes.data, theirs.ctypes.data)
            else:
                ref.akoDd137UnliftH(wrap, t, 1, 2 * t, fake, lp2.ctypes.data, hp2.ctypes.data, theirs.ctypes.data)
            assert np.array_equal(theirs[:n], mine), (n, amp)

            # and the reference's own property: unlift(lift(x)) == x
            back = np.zeros(n, np.int16)
            orc.orc_unlift_1d(wavelet, wrap, n, P(lp, i16p), P(hp, i16p), 1, P(back, i16p), 1)
            assert np.array_equal(back, x)


def _ref_lift_stream(ref, img_planes_dense, s, ch, w, h):
    """Run the reference akoLift on dense planes (re-laid out with the reference's plane spacing)."""
    spacing = ref.akoPlanesSpacing(w, h)
    size = ref.akoTileDataSize(w, h) * ch
    stride = w * h + spacing
    a = np.zeros(stride * ch + 64, np.int16)
    for c in range(ch):
        a[stride * c: stride * c + w * h] = img_planes_dense[c].ravel()
    b = np.zeros(size // 2 + stride * ch + 64, np.int16)
    ref.akoLift(0, C.byref(s), ch, w, h, spacing, a.ctypes.data, b.ctypes.data)
    return b[: size // 2].copy()


@pytest.mark.parametrize("wrap", [0, 1, 2, 3])
@pytest.mark.parametrize("wavelet", [W_DD137, W_CDF53, W_HAAR])
def test_lift_unlift_stream_vs_reference(orc, ref, wavelet, wrap):
    rs = np.random.RandomState(100 + wavelet * 4 + wrap)
    for (w, h, ch) in [(64, 64, 4), (65, 63, 3), (33, 47, 1), (16, 17, 2), (3, 3, 4), (5, 9, 3), (100, 7, 4),
                       (129, 70, 4)]:
        for q, g in [(0, 0), (16, 0), (7, 9)]:
            planes = rs.randint(-300, 600, size=(ch, h, w)).astype(np.int16)
            rs_set = ol.make_settings(ol.AkoSettings, wavelet=wavelet, wrap=wrap, q=q, g=g)
            os_set = ol.make_settings(ol.OrcSettings, wavelet=wavelet, wrap=wrap, q=q, g=g)
            want = _ref_lift_stream(ref, planes, rs_set, ch, w, h)
            got = np.zeros_like(want)
            tmp = planes.copy()
            orc.orc_lift(C.byref(os_set), ch, w, h, P(tmp, i16p), P(got, i16p))
            assert np.array_equal(want, got), (w, h, ch, q, g)

            # inverse: reference akoUnlift vs orc_unlift on the same stream
            spacing = ref.akoPlanesSpacing(w, h)
            stride = w * h + spacing
            a = np.zeros(len(want) + stride * ch + 64, np.int16)
            a[: len(want)] = want
            b = np.zeros(stride * ch + 64, np.int16)
            ref.akoUnlift(C.byref(rs_set), ch, 0, w, h, spacing, a.ctypes.data, b.ctypes.data)
            ref_planes = np.stack([b[stride * c: stride * c + w * h].reshape(h, w) for c in range(ch)])
            mine = np.zeros((ch, h, w), np.int16)
            st = want.copy()
            orc.orc_unlift(C.byref(os_set), ch, w, h, P(st, i16p), P(mine, i16p))
            assert np.array_equal(ref_planes, mine), (w, h, ch, q, g)
            if q == 0 and g == 0:
                assert np.array_equal(mine, planes)


def _runs_vector(rs, n, zero_heavy):
    out = []
    while len(out) < n:
        v = 0 if (zero_heavy and rs.rand() < 0.6) else int(rs.randint(-40, 41))
        if rs.rand() < 0.02:
            v = int(rs.randint(-32767, 32768))
        run = int(rs.choice([1, 1, 1, 2, 3, 4, 7, 50, 400]))
        out += [v] * run
    return np.array(out[:n], np.int16)


def test_kagari_vs_reference(orc, ref):
    rs = np.random.RandomState(5)
    vectors = [np.array([5], np.int16), np.array([0, 0], np.int16), np.array([0, 0, 0], np.int16),
               np.array([3, 3, 3, 3], np.int16), np.array([1, 2, 3, 4, 5, 6, 7, 8], np.int16),
               np.array([32767, -32767, 32767, -32767], np.int16)]
    vectors += [_runs_vector(rs, n, z) for n in (10, 100, 5000, 70000) for z in (False, True)]
    # overflow cases of the run counter (kagari.c:265-271): zero runs around 65 535 elements
    for n in (65534, 65535, 65536, 65537, 65538, 2 * 65534 + 10, 3 * 65535 + 7):
        vectors.append(np.zeros(n, np.int16))
        vectors.append(np.concatenate([np.array([9], np.int16), np.full(n, -2, np.int16), np.array([9, 9], np.int16)]))
    for v in vectors:
        n = len(v)
        cap = n * 4 + 64
        want = np.zeros(cap, np.uint8)
        wn = ref.akoKagariEncode(n * 2, cap, v.ctypes.data, want.ctypes.data)
        got = np.zeros(cap, np.uint8)
        gn = orc.orc_kagari_encode(n, P(v, i16p), cap, P(got, u8p))
        assert wn == gn and wn > 0 and np.array_equal(want[:wn], got[:gn]), n
        assert (orc.orc_kagari_bits(n, P(v, i16p)) + 7) // 8 == wn
        # decode both ways
        back = np.zeros(n + 8, np.int16)
        used = orc.orc_kagari_decode(n, wn, P(got, u8p), P(back, i16p))
        assert used == wn and np.array_equal(back[:n], v)
        back2 = np.zeros(n + 70000, np.int16)
        used2 = ref.akoKagariDecode(n, wn, (n + 70000) * 2, got.ctypes.data, back2.ctypes.data)
        assert used2 == wn and np.array_equal(back2[:n], v)


def test_kagari_minus_32768_encoder_bits(orc, ref):
    """-32768 zig-zags to 0xFFFF and +1 wraps the uint16 parameter of akoEliasEncodeStep to 0 (kagari.c:214-217 with
    :38-45): the reference emits ONE zero bit for it, which no decoder can read back. SURVEY 7.3: replicate the
    encoder's bits anyway (a 16-bit image cannot produce the value, a raw stage call can)."""
    rs = np.random.RandomState(32768)
    vectors = [np.array([-32768], np.int16), np.array([-32768, -32768], np.int16), np.full(9, -32768, np.int16),
               np.array([5, -32768, 5, 5, 5, -32768, -32768, -32768, -32768, 7], np.int16),
               np.concatenate([np.array([1], np.int16), np.full(70000, -32768, np.int16), np.array([2], np.int16)])]
    for n in (100, 5000):
        v = rs.randint(-40, 41, size=n).astype(np.int16)
        v[rs.rand(n) < 0.1] = -32768
        vectors.append(v)
    for v in vectors:
        n = len(v)
        cap = n * 4 + 64
        want = np.zeros(cap, np.uint8)
        wn = ref.akoKagariEncode(n * 2, cap, v.ctypes.data, want.ctypes.data)
        got = np.zeros(cap, np.uint8)
        gn = orc.orc_kagari_encode(n, P(v, i16p), cap, P(got, u8p))
        assert wn == gn and wn > 0 and np.array_equal(want[:wn], got[:gn]), n
        assert (orc.orc_kagari_bits(n, P(v, i16p)) + 7) // 8 == wn


def test_kagari_capacity_rule(orc, ref):
    """Success iff ceil(bits/8) < capacity (derived from kagari.c:65-68, :93-107)."""
    rs = np.random.RandomState(11)
    for n in (1, 2, 7, 64, 300):
        v = rs.randint(-3000, 3000, size=n).astype(np.int16)
        need = (orc.orc_kagari_bits(n, P(v, i16p)) + 7) // 8
        for cap in range(max(1, need - 3), need + 4):
            want = np.zeros(cap + 16, np.uint8)
            got = np.zeros(cap + 16, np.uint8)
            wn = ref.akoKagariEncode(n * 2, cap, v.ctypes.data, want.ctypes.data)
            gn = orc.orc_kagari_encode(n, P(v, i16p), cap, P(got, u8p))
            assert wn == gn, (n, cap, need)
            assert (wn != 0) == (need < cap)


@pytest.mark.parametrize("color", [C_YCOCG, C_SUBG, C_NONE, C_YCOCG_Q])
def test_format_vs_reference(orc, ref, color):
    rs = np.random.RandomState(color)
    for (w, h, ch) in [(16, 9, 4), (7, 5, 3), (9, 4, 1), (8, 8, 2), (5, 5, 6)]:
        for discard in (0, 1):
            img = noise_image(w, h, ch, 3 + w)
            img[rs.rand(h, w) < 0.3, ch - 1] = 0
            mine = np.zeros((ch, h, w), np.int16)
            orc.orc_format_forward(discard, color, ch, w, h, w, P(img, u8p), P(mine, i16p))
            theirs = np.zeros((ch, h, w), np.int16)
            ref.akoFormatToPlanarI16Yuv(discard, color, ch, w, h, w, 0, img.ctypes.data, theirs.ctypes.data)
            assert np.array_equal(mine, theirs)
            # inverse on arbitrary planes (saturation paths)
            pl = rs.randint(-600, 900, size=(ch, h, w)).astype(np.int16)
            a, b = pl.copy(), pl.copy()
            o1, o2 = np.zeros((h, w, ch), np.uint8), np.zeros((h, w, ch), np.uint8)
            orc.orc_format_inverse(color, ch, w, h, w, P(a, i16p), P(o1, u8p))
            ref.akoFormatToInterleavedU8Rgb(color, ch, w, h, 0, w, b.ctypes.data, o2.ctypes.data)
            assert np.array_equal(o1, o2)


def _e2e(orc, ref, img, **kw):
    rb, rst = ol.ref_encode(ref, img, **kw)
    ob, ost = ol.orc_encode(orc, img, **kw)
    assert rst == ost, (kw, rst, ost)
    assert rb == ob, kw
    if rb is None:
        return None
    rd, _ = ol.ref_decode(ref, rb)
    od, st = ol.orc_decode(orc, rb)
    assert st == 0 and np.array_equal(rd, od), kw
    return rb


@pytest.mark.parametrize("shape", SHAPES, ids=lambda s: "x".join(map(str, s)))
def test_end_to_end_shapes(orc, ref, shape):
    w, h, ch = shape
    for wavelet, wrap in itertools.product([W_DD137, W_CDF53, W_HAAR], [0, 1, 2, 3]):
        for q, g in [(0, 0), (16, 0), (5, 12)]:
            img = smooth_image(w, h, ch, 1 + w * 3 + h)
            blob = _e2e(orc, ref, img, wavelet=wavelet, wrap=wrap, q=q, g=g)
            if q == 0 and g == 0 and blob is not None:
                out, _ = ol.orc_decode(orc, blob)
                assert np.array_equal(out, img)


def test_end_to_end_options(orc, ref):
    img = ol.synth(orc, 200, 150, 9)
    for color in (C_YCOCG, C_SUBG, C_NONE):
        for discard in (0, 1):
            for cl in (0, 1, 3):
                _e2e(orc, ref, img, wavelet=W_CDF53, color=color, discard=discard, chroma_loss=cl, q=10, g=3)
    for comp in (0, 1, 2):  # Kagari, "Manbavaran" (== Kagari with another flag, compression.c:39), none
        _e2e(orc, ref, img, wavelet=W_DD137, compression=comp, q=16)
        _e2e(orc, ref, img, wavelet=W_DD137, compression=comp, q=0)
    # noise image, lossless: blob larger than the u8 input but within the int16 stream
    _e2e(orc, ref, noise_image(96, 96, 4, 2), wavelet=W_CDF53, q=0)


@pytest.mark.parametrize("tiles", [8, 32, 64, 256])
def test_end_to_end_tiles(orc, ref, tiles):
    for (w, h) in [(200, 150), (256, 256), (67, 131)]:
        img = ol.synth(orc, w, h, tiles + w)
        # edge tiles narrower than 3 px are outside the reference's own domain
        if 0 < w % tiles < 3 or 0 < h % tiles < 3:
            continue
        for wavelet in (W_DD137, W_CDF53, W_HAAR):
            _e2e(orc, ref, img, wavelet=wavelet, tiles=tiles, q=0)
            _e2e(orc, ref, img, wavelet=wavelet, tiles=tiles, q=16, g=4)


def test_status_codes(orc, ref):
    img = ol.synth(orc, 40, 40, 1)
    for kw in (dict(tiles=24), dict(tiles=4), dict(wrap=7), dict(wavelet=9), dict(color=5), dict(compression=3)):
        rb, rst = ol.ref_encode(ref, img, **kw)
        ob, ost = ol.orc_encode(orc, img, **kw)
        assert rb is None and ob is None and rst == ost and rst != 0, kw
    big = np.zeros((8, 8, 17), np.uint8)
    assert ol.ref_encode(ref, big)[1] == ol.orc_encode(orc, big)[1] == 2
    good, _ = ol.ref_encode(ref, img, wavelet=W_CDF53)
    for mutate in (lambda b: b"Bko" + b[3:], lambda b: b[:3] + b"\x03" + b[4:],
                   lambda b: b[:14] + b"\x80" + b[15:], lambda b: b[:4] + b"\0\0\0\0" + b[8:],
                   lambda b: b[:-5], lambda b: b[:16] + b"\x10\0\0\0" + b[20:]):
        bad = mutate(good)
        rd, rst = ol.ref_decode(ref, bad)
        od, ost = ol.orc_decode(orc, bad)
        assert rd is None and od is None and rst == ost, (rst, ost)


RATIO_CASES = [  # (w, h, seed, ratio, settings)
    (320, 200, 11, 10, dict(wavelet=0, g=0)),
    (320, 200, 11, 30, dict(wavelet=1, g=8)),
    (257, 131, 12, 6, dict(wavelet=2, g=0)),
    (300, 260, 13, 20, dict(wavelet=0, g=0, tiles=128)),
    (200, 160, 14, 4000, dict(wavelet=1, g=0)),   # unreachable target: the q *= 4 loop runs to its end
    (200, 160, 14, 1, dict(wavelet=0, q=40, g=5)),  # ratio 1 = lossless
    (200, 160, 14, 0, dict(wavelet=0, q=12, g=0)),  # ratio 0 = one plain pass
    (128, 128, 15, 12, dict(wavelet=0, color=1, g=0)),
    (1296, 1040, 16, 15, dict(wavelet=0, wrap=2, g=4)),  # REPEAT at a size where levels are strip + frame on the GPU
]


@pytest.mark.parametrize("case", RATIO_CASES, ids=lambda c: f"{c[0]}x{c[1]}-r{c[3]}")
def test_ratio_search_vs_reference(orc, ref, case):
    """orc_encode_pass (EncodePass, tools/akoenc.cpp:111-213) against the same loop driven over the unmodified
    reference's akoEncodeExt: same blob, same quantisation, same number of passes."""
    w, h, seed, ratio, kw = case
    img = ol.synth(orc, w, h, seed)
    want, want_q, want_passes = ol.ref_encode_pass(ref, img, ratio, **kw)
    got, st, q, passes = ol.orc_encode_pass(orc, img, ratio, **kw)
    assert got == want
    assert (q, passes) == (want_q, want_passes)


def test_corrupted_blobs_oracle_vs_reference(orc, ref):
    """Corrupted blobs: the oracle returns the reference's status and pixels, with ONE documented exception -- a block
    size field that points past the end of the input. The reference does not look at input_size there (decode.c:71 is
    a tautology) and reads beyond the caller's buffer; when its accumulator's read-ahead happens to stop
    exactly at the forged size it even answers OK (kagari.c:119-143, compression.c:69-70). The oracle, like the CUDA
    library, answers AKO_BROKEN_INPUT for every block that overruns the input."""
    import struct
    rs = np.random.RandomState(321)

    def overruns(b, kw):
        if kw.get("tiles"):
            return True  # a forged size in a multi-tile blob shifts every later block head: not analysed here
        return len(b) >= 20 and 20 + struct.unpack("<I", b[16:20])[0] > len(b)

    for (w, h, kw) in [(96, 80, dict(wavelet=0, q=16, g=0)), (200, 131, dict(wavelet=1, q=0, g=0)),
                       (64, 64, dict(wavelet=2, q=8, g=4, tiles=32))]:
        blob, _ = ol.orc_encode(orc, ol.synth(orc, w, h, w), **kw)
        for trial in range(60):
            b = bytearray(blob)
            mode = trial % 3
            if mode == 0:
                for _ in range(rs.randint(1, 4)):
                    b[rs.randint(20, len(b))] ^= 1 << rs.randint(0, 8)
            elif mode == 1:
                b[16 + rs.randint(0, 4)] ^= 1 << rs.randint(0, 8)
            else:
                a = rs.randint(20, len(b) - 4)
                n = rs.randint(1, 16)
                b[a:a + n] = bytes(rs.randint(0, 256, size=n).astype(np.uint8))
            b = bytes(b)
            if overruns(b, kw):
                got, st = ol.orc_decode(orc, b)
                if not kw.get("tiles"):
                    assert st == 15 and got is None
                continue
            want, wst = ol.ref_decode(ref, b)
            got, st = ol.orc_decode(orc, b)
            assert st == wst, (w, h, trial, st, wst)
            if want is not None:
                assert np.array_equal(got, want), (w, h, trial)
