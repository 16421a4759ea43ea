"""GPU parity: the CUDA path (through the C-ABI of libako_b200.so) against the oracle, the committed golden
vectors and the survey's known answers. Bit-exact everywhere (integer / byte work)."""
import base64
import ctypes as C
import hashlib
import itertools
import json
import os

import numpy as np
import pytest

import ako_b200
import oracle_lib as ol
from cases import *
from golden.make_golden import make_input

pytestmark = pytest.mark.gpu
P, u8p, i16p = ol._p, ol.u8p, ol.i16p


def sha(b):
    return hashlib.sha256(b).hexdigest()


@pytest.fixture(scope="module")
def ctx():
    c = ako_b200.Context()
    yield c
    c.close()


def S(**kw):
    return ako_b200.default_settings(**kw)


def OS(**kw):
    return ol.make_settings(ol.OrcSettings, **kw)


# ------------------------------------------------------------------ stages

@pytest.mark.parametrize("color", [C_YCOCG, C_SUBG, C_NONE, C_YCOCG_Q])
def test_format_stage(orc, ctx, color):
    rs = np.random.RandomState(color)
    for (w, h, ch) in [(16, 9, 4), (64, 33, 4), (7, 5, 3), (9, 4, 1), (8, 8, 2), (5, 5, 6), (24, 3, 16), (64, 33, 3), (24, 9, 3),
                       (200, 17, 3)]:
        for discard in (0, 1):
            img = noise_image(w, h, ch, 3 + w)
            img[rs.rand(h, w) < 0.3, ch - 1] = 0
            want = np.zeros((ch, h, w), np.int16)
            orc.orc_format_forward(discard, color, ch, w, h, w, P(img, u8p), P(want, i16p))
            got = ctx.format_forward(img, S(color=color, discard=discard))
            assert np.array_equal(want, got), (w, h, ch, discard)
            pl = rs.randint(-600, 900, size=(ch, h, w)).astype(np.int16)
            a = pl.copy()
            want_px = np.zeros((h, w, ch), np.uint8)
            orc.orc_format_inverse(color, ch, w, h, w, P(a, i16p), P(want_px, u8p))
            assert np.array_equal(want_px, ctx.format_inverse(pl, color)), (w, h, ch)


LIFT_SHAPES = [(64, 64, 4), (65, 63, 3), (33, 47, 1), (16, 17, 2), (3, 3, 4), (5, 9, 3), (100, 7, 4), (129, 70, 4),
               (200, 131, 2), (257, 66, 1), (128, 128, 1), (15, 300, 1),
               # strip-kernel geometry: several / partial 128-column strips, several row splits, odd heights
               (528, 70, 1), (272, 601, 2), (1040, 48, 1), (64, 17, 3), (2064, 1100, 1), (264, 40, 1), (408, 616, 2),
               (248, 33, 1), (968, 72, 1)]


@pytest.mark.parametrize("wavelet", [W_DD137, W_CDF53, W_HAAR])
def test_strip_kernels_any_width(orc, ctx, wavelet):
    """The strip kernels at every residue of the width: 34 consecutive widths (coefficient columns 65 .. 82: every
    position of the right edge inside a 16-column chunk of the forward H pass and an 8-column chunk of the inverse one,
    every row misalignment of the subbands), two strips, several row splits, both directions against the oracle."""
    rs = np.random.RandomState(500 + wavelet)
    for w in list(range(130, 164)) + [263, 519, 1031]:
        h = 40 + (w % 5)
        for ch, (q, g) in [(1, (0, 0)), (2, (7, 9))]:
            planes = rs.randint(-300, 600, size=(ch, h, w)).astype(np.int16)
            n = orc.orc_tile_data_size(w, h) * ch // 2
            want = np.zeros(n, np.int16)
            os_ = OS(wavelet=wavelet, wrap=0, q=q, g=g)
            orc.orc_lift(C.byref(os_), ch, w, h, P(planes.copy(), i16p), P(want, i16p))
            s = S(wavelet=wavelet, wrap=0, q=q, g=g)
            assert np.array_equal(want, ctx.lift(planes, s)), (w, h, ch, q)
            back_want = np.zeros((ch, h, w), np.int16)
            orc.orc_unlift(C.byref(os_), ch, w, h, P(want.copy(), i16p), P(back_want, i16p))
            assert np.array_equal(back_want, ctx.unlift(want, s, ch, w, h)), (w, h, ch, q)


@pytest.mark.parametrize("wrap", [0, 1, 2, 3])
@pytest.mark.parametrize("wavelet", [W_DD137, W_CDF53, W_HAAR])
def test_lift_unlift_stage(orc, ctx, wavelet, wrap):
    rs = np.random.RandomState(100 + wavelet * 4 + wrap)
    for (w, h, ch) in LIFT_SHAPES:
        for q, g in [(0, 0), (16, 0), (7, 9)]:
            amp = 30000 if (q == 0 and w == 64) else 600
            planes = rs.randint(-amp // 2, amp, size=(ch, h, w)).astype(np.int16)
            n = orc.orc_tile_data_size(w, h) * ch // 2
            want = np.zeros(n, np.int16)
            tmp = planes.copy()
            os_ = OS(wavelet=wavelet, wrap=wrap, q=q, g=g)
            orc.orc_lift(C.byref(os_), ch, w, h, P(tmp, i16p), P(want, i16p))
            s = S(wavelet=wavelet, wrap=wrap, q=q, g=g)
            got = ctx.lift(planes, s)
            assert np.array_equal(want, got), (w, h, ch, q, g, int(np.argmax(want != got)))

            back_want = np.zeros((ch, h, w), np.int16)
            st = want.copy()
            orc.orc_unlift(C.byref(os_), ch, w, h, P(st, i16p), P(back_want, i16p))
            back = ctx.unlift(want, s, ch, w, h)
            assert np.array_equal(back_want, back), (w, h, ch, q, g)
            if q == 0 and g == 0:
                assert np.array_equal(back, planes)


@pytest.mark.parametrize("wrap", [1, 2, 3])
@pytest.mark.parametrize("wavelet", [W_DD137, W_CDF53])
def test_wrap_modes_strip_plus_frame(orc, ctx, wavelet, wrap):
    """MIRROR / REPEAT / ZERO at sizes where a level is done by the CLAMP strip kernel and then has the tiles along its
    edge computed again with the real wrap mode (ako_device.cu frame_worth): last tile column / row thinner than the
    reach of the edge taps (two columns / rows of tiles in the frame), odd sizes, several channels."""
    rs = np.random.RandomState(700 + wavelet * 4 + wrap)
    for (w, h, ch) in [(1500, 1200, 2), (1290, 1034, 1), (1026, 2100, 1), (1281, 1027, 1), (2064, 1100, 3)]:
        for q, g in [(0, 0), (7, 9)]:
            planes = rs.randint(-300, 600, size=(ch, h, w)).astype(np.int16)
            n = orc.orc_tile_data_size(w, h) * ch // 2
            want = np.zeros(n, np.int16)
            tmp = planes.copy()
            os_ = OS(wavelet=wavelet, wrap=wrap, q=q, g=g)
            orc.orc_lift(C.byref(os_), ch, w, h, P(tmp, i16p), P(want, i16p))
            s = S(wavelet=wavelet, wrap=wrap, q=q, g=g)
            got = ctx.lift(planes, s)
            assert np.array_equal(want, got), (w, h, ch, q, g, int(np.argmax(want != got)))
            back_want = np.zeros((ch, h, w), np.int16)
            st = want.copy()
            orc.orc_unlift(C.byref(os_), ch, w, h, P(st, i16p), P(back_want, i16p))
            back = ctx.unlift(want, s, ch, w, h)
            assert np.array_equal(back_want, back), (w, h, ch, q, g)
            if q == 0 and g == 0:
                assert np.array_equal(back, planes)
    # and through the whole codec
    img = ol.synth(orc, 1290, 1034, 77)
    _e2e(orc, img, wavelet=wavelet, wrap=wrap, q=0)
    _e2e(orc, img, wavelet=wavelet, wrap=wrap, q=12, g=5)
    # level 0 fused with the RGBA8 read (width a multiple of 4): the frame converts the pixels it needs itself
    img = ol.synth(orc, 1296, 1040, 79)
    img[:40, :, 3] = 0
    img[:, -30:, 3] = 0
    img[500:520, 600:640, 3] = 0
    for color in (C_YCOCG, C_SUBG, C_NONE):
        _e2e(orc, img, wavelet=wavelet, wrap=wrap, color=color, discard=1, q=8, g=2)
    _e2e(orc, img, wavelet=wavelet, wrap=wrap, q=0)
    if wavelet == W_DD137:
        # tiles large enough to have a frame of their own, as batch members (four tile shapes). Encoder only: the
        # reference's reader refuses tiles above 512 (AKO_INVALID_FLAGS), and so do the oracle and this library.
        _e2e(orc, ol.synth(orc, 3100, 2900, 78), wavelet=wavelet, wrap=wrap, q=12, g=5, tiles=2048)


def _runs_vector(rs, n, zero_heavy):
    out = []
    while len(out) < n:
        v = 0 if (zero_heavy and rs.rand() < 0.6) else int(rs.randint(-40, 41))
        if rs.rand() < 0.02:
            v = int(rs.randint(-32767, 32768))
        out += [v] * int(rs.choice([1, 1, 1, 2, 3, 4, 7, 50, 400, 5000]))
    return np.array(out[:n], np.int16)


def test_kagari_stage(orc, ctx):
    rs = np.random.RandomState(5)
    vectors = [np.array([5], np.int16), np.array([0, 0], np.int16), np.array([0, 0, 0], np.int16),
               np.array([3, 3, 3, 3], np.int16), np.arange(1, 9).astype(np.int16),
               np.array([32767, -32767, 32767, -32767], np.int16),
               rs.randint(-32767, 32768, size=5000).astype(np.int16)]
    vectors += [_runs_vector(rs, n, z) for n in (10, 100, 2047, 2048, 2049, 5000, 70000, 300001) for z in (False, True)]
    for n in (65534, 65535, 65536, 65537, 65538, 2 * 65534 + 10, 3 * 65535 + 7):
        vectors.append(np.zeros(n, np.int16))
        vectors.append(np.concatenate([np.array([9], np.int16), np.full(n, -2, np.int16), np.array([9, 9], np.int16)]))
    for v in vectors:
        n = len(v)
        cap = n * 4 + 64
        want = np.zeros(cap, np.uint8)
        wn = orc.orc_kagari_encode(n, P(v, i16p), cap, P(want, u8p))
        got = ctx.kagari_encode(v, cap)
        assert got is not None and len(got) == wn and np.array_equal(want[:wn], got), n
        used, back = ctx.kagari_decode(got, n)
        assert used == wn and np.array_equal(back, v), n


def test_kagari_minus_32768_encoder_bits(orc, ctx):
    """The reference's encoder emits a single zero bit for -32768 (kagari.c:214-217 with :38-45; undecodable, SURVEY
    7.3 "replicate encoder bits anyway"): the CUDA encoder must produce the same bytes; nothing is decoded back."""
    rs = np.random.RandomState(32768)
    vectors = [np.array([-32768], np.int16), np.array([-32768, -32768], np.int16), np.full(9, -32768, np.int16),
               np.array([5, -32768, 5, 5, 5, -32768, -32768, -32768, -32768, 7], np.int16),
               np.concatenate([np.array([1], np.int16), np.full(70000, -32768, np.int16), np.array([2], np.int16)])]
    for n in (100, 5000, 9000):
        v = rs.randint(-40, 41, size=n).astype(np.int16)
        v[rs.rand(n) < 0.1] = -32768
        vectors.append(v)
    for v in vectors:
        n = len(v)
        cap = n * 4 + 64
        want = np.zeros(cap, np.uint8)
        wn = orc.orc_kagari_encode(n, P(v, i16p), cap, P(want, u8p))
        got = ctx.kagari_encode(v, cap)
        assert got is not None and len(got) == wn and np.array_equal(want[:wn], got), n


def test_strip_kernels_full_amplitude(orc, ctx):
    """+-32767 planes through several strips and row splits (2064 x 1100), all three wavelets: the hi-domain
    arithmetic of the strip kernels (|n| < 2^26, lift_strip.cuh hi_step) at the largest tap sums, lossless and
    quantised, forward and inverse."""
    rs = np.random.RandomState(77)
    w, h = 2064, 1100
    for wavelet in (W_DD137, W_CDF53, W_HAAR):
        planes = rs.randint(-32767, 32768, size=(1, h, w)).astype(np.int16)
        planes[0, ::97, ::53] = 32767
        planes[0, 5::89, 7::61] = -32767
        for q, g in ((0, 0), (9, 13)):
            n = orc.orc_tile_data_size(w, h) // 2
            want = np.zeros(n, np.int16)
            tmp = planes.copy()
            os_ = OS(wavelet=wavelet, q=q, g=g)
            orc.orc_lift(C.byref(os_), 1, w, h, P(tmp, i16p), P(want, i16p))
            s = S(wavelet=wavelet, q=q, g=g)
            got = ctx.lift(planes, s)
            assert np.array_equal(want, got), (wavelet, q, int(np.argmax(want != got)))
            back_want = np.zeros((1, h, w), np.int16)
            st = want.copy()
            orc.orc_unlift(C.byref(os_), 1, w, h, P(st, i16p), P(back_want, i16p))
            back = ctx.unlift(want, s, 1, w, h)
            assert np.array_equal(back_want, back), (wavelet, q)
            if q == 0:
                assert np.array_equal(back, planes)


def test_kagari_capacity_rule(orc, ctx):
    rs = np.random.RandomState(11)
    for n in (7, 64, 300, 3000):
        v = rs.randint(-3000, 3000, size=n).astype(np.int16)
        need = (orc.orc_kagari_bits(n, P(v, i16p)) + 7) // 8
        for cap in (need - 5, need - 1, need, need + 1, need + 4, need + 9):
            got = ctx.kagari_encode(v, cap)
            # the C-ABI packs whole words, so besides "bytes < cap" (the reference's rule) it needs bytes <= cap&~3
            if need < cap and need <= (cap & ~3):
                assert got is not None and len(got) == need
            elif need >= cap:
                assert got is None


def test_kagari_decode_rejects_broken(ctx):
    v = np.arange(-100, 100).astype(np.int16)
    good = ctx.kagari_encode(v)
    used, _ = ctx.kagari_decode(good[:-3], len(v))
    assert used == 0
    used, _ = ctx.kagari_decode(np.zeros(40, np.uint8), 16)
    assert used == 0


def _orc_kagari_decode(orc, data, n):
    data = np.ascontiguousarray(data, dtype=np.uint8)
    out = np.zeros(n, np.int16)
    used = orc.orc_kagari_decode(n, len(data), P(data, u8p), P(out, i16p))
    return used, out


def test_kagari_decode_answers_like_the_reference_on_any_bytes(orc, ctx):
    """What the parallel decoder does not accept goes to a kernel that IS the reference's decoder (64-bit accumulator,
    refill policy, kagari.c:119-163): blocks with trailing bytes (the reference reports the bytes its read-ahead has
    fetched, so some are accepted), codewords of 16..27 leading zeros (accepted, value truncated to 16 bits), random
    bytes, truncated blocks. Consumed bytes and -- whenever the block is accepted -- the values are the oracle's."""
    rs = np.random.RandomState(77)
    cases = []
    for trial in range(24):
        n = int(rs.randint(20, 6000))
        v = (rs.randint(-40, 41, size=n) * (rs.rand(n) < 0.4)).astype(np.int16)
        good = ctx.kagari_encode(v, cap=4 * n + 64)
        for tail in (0, 1, 2, 3, 5, 6, 7, 9):
            extra = rs.randint(0, 256, size=tail).astype(np.uint8) if trial % 2 else np.zeros(tail, np.uint8)
            cases.append((np.concatenate([good, extra]), n))
        cases.append((good[:len(good) - 1 - trial % 4], n))
        cases.append((good, n + 1 + trial))      # asks for more values than the block holds
        cases.append((good, max(1, n - 1 - trial)))  # ... and for fewer
    for trial in range(60):
        # long zero spans between set bits: codewords with many leading zeros
        size = int(rs.randint(8, 400))
        raw = (rs.randint(0, 256, size=size) * (rs.rand(size) < 0.25)).astype(np.uint8)
        cases.append((raw, int(rs.randint(1, 200))))
    # 17 zeros, a one, 17 more bits: a codeword of 35 bits whose value is truncated to 16 bits
    bits = "0" * 17 + "1" + "01100110011001101" + "1" * 11
    raw = np.frombuffer(int(bits, 2).to_bytes(len(bits) // 8 + 1, "big")[-(len(bits) + 7) // 8:], dtype=np.uint8)
    cases.append((np.concatenate([raw, np.full(8, 0xFF, np.uint8)]), 4))
    accepted = 0
    for data, n in cases:
        want_used, want = _orc_kagari_decode(orc, data, n)
        used, got = ctx.kagari_decode(data, n)
        assert used == want_used, (len(data), n, used, want_used)
        if want_used:
            accepted += 1
            assert np.array_equal(got, want), (len(data), n)
    assert accepted >= 40


# ------------------------------------------------------------------ whole codec through akoEncodeExt / akoDecodeExt

def _e2e(orc, img, **kw):
    want, wst = ol.orc_encode(orc, img, **kw)
    got, gst = ako_b200.encode(img, S(**kw))
    assert wst == gst, (kw, wst, gst)
    assert want == got, kw
    if want is None:
        return None
    want_px, dst = ol.orc_decode(orc, want)
    got_px, st, s = ako_b200.decode(want)
    if want_px is None:
        # (tiles of 1024 and up encode, but their tile code sets a flag bit the reference's reader refuses, head.c:124)
        assert st == dst and got_px is None, (kw, st, dst)
        return want
    assert st == 0 and np.array_equal(want_px, got_px), kw
    return want


@pytest.mark.parametrize("shape", SHAPES, ids=lambda s: "x".join(map(str, s)))
def test_end_to_end_shapes(orc, shape):
    w, h, ch = shape
    for wavelet, wrap in itertools.product([W_DD137, W_CDF53, W_HAAR], [0, 1, 2, 3]):
        for q, g in [(0, 0), (16, 0), (5, 12)]:
            img = smooth_image(w, h, ch, 1 + w * 3 + h)
            blob = _e2e(orc, img, wavelet=wavelet, wrap=wrap, q=q, g=g)
            if q == 0 and g == 0 and blob is not None:
                out, st, _ = ako_b200.decode(blob)
                assert np.array_equal(out, img)


@pytest.mark.parametrize("shape", [(1921, 1081, 4), (1000, 1000, 3), (777, 555, 2), (2050, 300, 4), (1028, 516, 4)],
                         ids=lambda s: "x".join(map(str, s)))
def test_end_to_end_odd_sizes(orc, shape):
    """Sizes off the multiples of 8 at the scale where the strip kernels carry every level (several strips and row
    splits per level, odd widths and heights down the pyramid), all wavelets, lossless and quantised."""
    w, h, ch = shape
    img = ol.synth(orc, w, h, 3 + w)[..., :ch].copy()
    for wavelet in (W_DD137, W_CDF53, W_HAAR):
        for q, g in [(0, 0), (16, 16)]:
            _e2e(orc, img, wavelet=wavelet, q=q, g=g)


def test_end_to_end_options(orc):
    img = ol.synth(orc, 200, 150, 9)
    for color in (C_YCOCG, C_SUBG, C_NONE):
        for discard in (0, 1):
            for cl in (0, 1, 3):
                _e2e(orc, img, wavelet=W_CDF53, color=color, discard=discard, chroma_loss=cl, q=10, g=3)
    for comp in (0, 1, 2):
        _e2e(orc, img, wavelet=W_DD137, compression=comp, q=16)
        _e2e(orc, img, wavelet=W_DD137, compression=comp, q=0)
    _e2e(orc, noise_image(96, 96, 4, 2), wavelet=W_CDF53, q=0)
    _e2e(orc, img, wavelet=W_NONE, compression=2, q=0)


@pytest.mark.parametrize("wavelet", [W_DD137, W_CDF53, W_HAAR])
def test_end_to_end_rgba_level0_fused(orc, wavelet):
    """4-channel CLAMP images whose level 0 runs in the fused colour + lifting kernel (lift_strip4.cuh): noise RGBA
    with transparent pixels, one / several / partial 64-column strips, several row splits, odd heights, every colour
    model with and without alpha discard, every quantiser mode -- .ako bytes against the oracle's."""
    rs = np.random.RandomState(7 + wavelet)
    for (w, h) in [(64, 16), (72, 17), (136, 72), (200, 151), (1032, 530), (2064, 1100)]:
        img = noise_image(w, h, 4, w + h)
        img[rs.rand(h, w) < 0.3, 3] = 0
        big = w * h > 100000
        for color in ((C_YCOCG,) if big else (C_YCOCG, C_SUBG, C_NONE)):
            for discard in ((1,) if big else (0, 1)):
                for q, g in [(0, 0), (16, 0), (5, 12)]:
                    _e2e(orc, img, wavelet=wavelet, color=color, discard=discard, q=q, g=g)
    # smooth content as well (the quantised planes of noise are all escapes)
    img = ol.synth(orc, 1632, 616, 3)
    _e2e(orc, img, wavelet=wavelet, q=16, g=16)


@pytest.mark.parametrize("tiles", [8, 32, 64, 256])
def test_end_to_end_tiles(orc, tiles):
    for (w, h) in [(200, 150), (256, 256), (67, 131)]:
        if 0 < w % tiles < 3 or 0 < h % tiles < 3:
            continue
        img = ol.synth(orc, w, h, tiles + w)
        for wavelet in (W_DD137, W_CDF53, W_HAAR):
            _e2e(orc, img, wavelet=wavelet, tiles=tiles, q=0)
            _e2e(orc, img, wavelet=wavelet, tiles=tiles, q=16, g=4)


def test_end_to_end_rgb_tiles(orc):
    """3-channel images whose tiles start on multiples of 8 pixels take the 8-pixels-per-thread RGB8 format kernels with
    the tile addressing (full tiles, right / bottom / corner tiles of another shape)."""
    for (w, h, tiles) in [(256, 192, 64), (200, 136, 64), (520, 264, 128)]:
        img = ol.synth(orc, w, h, 5 + w)[..., :3].copy()
        for wavelet in (W_DD137, W_CDF53):
            _e2e(orc, img, wavelet=wavelet, tiles=tiles, q=0)
            _e2e(orc, img, wavelet=wavelet, tiles=tiles, q=12, g=4, color=C_SUBG)


def test_more_tiles_than_a_grid_dimension(orc):
    """2048 x 2056 with tiles_dimension = 8 is 65 792 tiles (a grid dimension holds 65 535); all four tile shape
    groups with every wavelet at a smaller size, through events (tile by tile) and without (one pass per shape)."""
    from ako_b200 import lib as akolib
    img = ol.synth(orc, 2048, 2056, 3)
    _e2e(orc, img, wavelet=W_CDF53, tiles=8, q=0)
    img = ol.synth(orc, 333, 222, 4)[..., :3].copy()
    seen = []
    cb = ako_b200.load().akoDefaultCallbacks()
    fn = akolib.EVENTS_FN(lambda t, n, e, d: seen.append((t, n, e)))
    cb.events = C.cast(fn, C.c_void_p)
    for wavelet in (W_DD137, W_CDF53, W_HAAR):
        for tiles in (16, 64):
            want = _e2e(orc, img, wavelet=wavelet, tiles=tiles, q=16, g=8)
            del seen[:]
            got, st = ako_b200.encode(img, S(wavelet=wavelet, tiles=tiles, q=16, g=8), cb)
            assert st == 0 and got == want
            n_tiles = -(-333 // tiles) * -(-222 // tiles)
            assert len(seen) == 6 * n_tiles and [t for t, _, _ in seen] == sorted(t for t, _, _ in seen)
            px, st, _ = ako_b200.decode(want, cb)
            want_px, _ = ol.orc_decode(orc, want)
            assert st == 0 and np.array_equal(px, want_px)


with open(os.path.join(os.path.dirname(__file__), "golden", "golden.json")) as f:
    GOLDEN = json.load(f)


def test_golden_vectors():
    """Committed outputs of the unmodified reference (tests/golden/make_golden.py)."""
    orc = ol.load_oracle()
    for c in GOLDEN:
        img = make_input(orc, c)
        blob, st = ako_b200.encode(img, S(wavelet=c["wavelet"], wrap=c["wrap"], q=c["q"], g=c["g"], tiles=c["tiles"],
                                          color=c["color"], discard=c["discard"], chroma_loss=c["chroma_loss"]))
        assert st == 0 and len(blob) == c["blob_len"] and sha(blob) == c["blob_sha256"], c
        if "blob_b64" in c:
            assert blob == base64.b64decode(c["blob_b64"])
        dec, st, _ = ako_b200.decode(blob)
        assert st == 0 and sha(dec.tobytes()) == c["decoded_sha256"], c


@pytest.mark.parametrize("kat", KATS, ids=lambda k: f"{k[0]}x{k[1]}-w{k[2]}-q{k[3]}-g{k[4]}")
def test_known_answers(orc, kat):
    """SHA-256 of the reference's blobs and decoded pixels at the BASELINE shapes."""
    w, h, wavelet, q, g, seed, size, blob_sha, dec_sha = kat
    img = ol.synth(orc, w, h, seed)
    blob, st = ako_b200.encode(img, S(wavelet=wavelet, q=q, g=g))
    assert st == 0 and len(blob) == size and sha(blob) == blob_sha
    out, st, s = ako_b200.decode(blob)
    assert st == 0
    if dec_sha:
        assert sha(out.tobytes()) == dec_sha
    if q == 0:
        assert np.array_equal(out, img)
    assert (s.wavelet, s.wrap, s.compression) == (wavelet, 0, 0)


def test_reference_decodes_our_blobs_and_we_decode_its(orc, ref):
    """Cross-decoding with the real reference library (oracle/_ref)."""
    img = ol.synth(orc, 333, 222, 12)
    for wavelet in (W_DD137, W_CDF53, W_HAAR):
        ours, st = ako_b200.encode(img, S(wavelet=wavelet, q=8, g=2))
        theirs, _ = ol.ref_encode(ref, img, wavelet=wavelet, q=8, g=2)
        assert ours == theirs
        a, _ = ol.ref_decode(ref, ours)
        b, st, _ = ako_b200.decode(theirs)
        assert np.array_equal(a, b)


def test_status_codes(orc):
    img = ol.synth(orc, 40, 40, 1)
    good, _ = ol.orc_encode(orc, img, wavelet=W_CDF53)
    for mutate in (lambda b: b[:-5], lambda b: b[:16] + b"\x10\0\0\0" + b[20:], lambda b: b[:20],
                   lambda b: b[:16] + b"\0\0\0\0" + b[20:]):
        bad = mutate(good)
        od, ost = ol.orc_decode(orc, bad)
        gd, gst, _ = ako_b200.decode(bad)
        assert od is None and gd is None and gst == ost == 15
    # incompressible edge tiles: the block would not be smaller than its stream -> AKO_ERROR (encode.c:159-164)
    noise = noise_image(9, 9, 4, 1)
    want, wst = ol.orc_encode(orc, noise, wavelet=W_CDF53, q=0, tiles=8)
    got, gst = ako_b200.encode(noise, S(wavelet=W_CDF53, q=0, tiles=8))
    assert (want is None) == (got is None) and wst == gst


def test_events_order():
    """ako.h:75-84: per tile FORMAT, WAVELET, COMPRESSION on encode; the reverse stage order on decode."""
    from ako_b200 import lib as akolib
    L = ako_b200.load()
    seen = []
    fn = akolib.EVENTS_FN(lambda t, n, e, d: seen.append((t, n, e)))
    cb = L.akoDefaultCallbacks()
    cb.events = C.cast(fn, C.c_void_p)
    img = noise_image(40, 24, 4, 3)
    blob, st = ako_b200.encode(img, S(wavelet=W_CDF53, tiles=16, q=4), cb)
    assert st == 0
    tiles = 3 * 2
    assert seen == [(t, tiles, e) for t in range(tiles) for e in (1, 2, 3, 4, 5, 6)]
    seen.clear()
    out, st, _ = ako_b200.decode(blob, cb)
    assert st == 0
    assert seen == [(t, tiles, e) for t in range(tiles) for e in (5, 6, 3, 4, 1, 2)]


# ------------------------------------------------------------------ device-resident and batched API

def test_device_resident_and_batch(orc, ctx):
    w, h, ch, n = 320, 200, 4, 5
    imgs = np.stack([ol.synth(orc, w, h, 1000 + i) for i in range(n)])
    s = S(wavelet=W_DD137, q=16, g=16)
    want = [ol.orc_encode(orc, imgs[i], wavelet=W_DD137, q=16, g=16)[0] for i in range(n)]

    bound = ctx.encode_bound(s, ch, w, h)
    stride = (bound + 255) // 256 * 256
    d_in = ctx.to_device(imgs)
    d_out = ctx.alloc(stride * n)
    # single image, device resident
    size, st = ctx.encode_device(s, ch, w, h, d_in, d_out, stride)
    assert st == 0 and size == len(want[0])
    assert ctx.to_host(d_out, (size,), np.uint8).tobytes() == want[0]
    # batch
    done, st, sizes = ctx.encode_batch_device(s, ch, w, h, n, d_in, w * h * ch, d_out, stride)
    assert done == n and st == 0 and sizes == [len(b) for b in want]
    blobs = ctx.to_host(d_out, (n, stride), np.uint8)
    for i in range(n):
        assert blobs[i, :sizes[i]].tobytes() == want[i]
    # batch decode of those blobs
    d_px = ctx.alloc(w * h * ch * n)
    done, st = ctx.decode_batch_device(n, d_out, stride, sizes, d_px, w * h * ch)
    assert done == n and st == 0
    px = ctx.to_host(d_px, (n, h, w, ch), np.uint8)
    for i in range(n):
        assert np.array_equal(px[i], ol.orc_decode(orc, want[i])[0])
    # single decode, device resident
    st, dims, s2 = ctx.decode_device(sizes[2], d_out + 2 * stride, d_px, w * h * ch)
    assert st == 0 and dims == (ch, w, h)
    assert np.array_equal(ctx.to_host(d_px, (h, w, ch), np.uint8), px[2])
    # a batch of one untiled image: its block walk rides on the head's read-back (ako_host.c first_block)
    done, st = ctx.decode_batch_device(1, d_out + 3 * stride, stride, sizes[3:4], d_px, w * h * ch)
    assert done == 1 and st == 0
    assert np.array_equal(ctx.to_host(d_px, (h, w, ch), np.uint8), px[3])
    # ... and answers like the block walk for a blob cut short or with a forged / zero block size
    for cut in (17, 20, sizes[3] - 1):
        done, st = ctx.decode_batch_device(1, d_out + 3 * stride, stride, [cut], d_px, w * h * ch)
        assert done == 0 and st == ol.orc_decode(orc, want[3][:cut])[1] == 15, (cut, st)
        st, dims, _ = ctx.decode_device(cut, d_out + 3 * stride, d_px, w * h * ch)
        assert st == 15, (cut, st)
    bad = np.frombuffer(want[3], np.uint8).copy()
    bad[16:20] = 0
    d_bad = ctx.to_device(bad)
    done, st = ctx.decode_batch_device(1, d_bad, stride, [len(bad)], d_px, w * h * ch)
    assert done == 0 and st == ol.orc_decode(orc, bad.tobytes())[1] == 15
    st, dims, _ = ctx.decode_device(len(bad), d_bad, d_px, w * h * ch)
    assert st == 15
    ctx.free(d_bad)
    assert ctx.launch_count() > 0
    for p in (d_in, d_out, d_px):
        ctx.free(p)


# ------------------------------------------------------------------ BASELINE full sizes: size-independent properties

@pytest.mark.parametrize("kat", KATS_BIG[:3], ids=lambda k: f"{k[0]}x{k[1]}-w{k[2]}")
def test_big_known_answers(kat):
    orc = ol.load_oracle()
    w, h, wavelet, q, g, seed, size, blob_sha, _ = kat
    img = ol.synth(orc, w, h, seed)
    blob, st = ako_b200.encode(img, S(wavelet=wavelet, q=q, g=g))
    assert st == 0 and len(blob) == size and sha(blob) == blob_sha


@pytest.mark.parametrize("wavelet", [W_DD137, W_CDF53, W_HAAR])
def test_big_lossless_roundtrip_property(wavelet):
    """8192x8192 RGBA8, q=0: decode(encode(x)) == x, through the public API."""
    orc = ol.load_oracle()
    img = ol.synth(orc, 8192, 8192, 4)
    blob, st = ako_b200.encode(img, S(wavelet=wavelet, q=0))
    assert st == 0
    out, st, _ = ako_b200.decode(blob)
    assert st == 0 and np.array_equal(out, img)


def test_c5_lossless_16384_known_answer_and_roundtrip():
    """BASELINE configs[4]: 16384x16384 RGBA8, CDF 5/3, q=0 g=0, one tile, one 3.66 Gbit entropy stream.
    Known answer from the unmodified reference + exact round trip."""
    orc = ol.load_oracle()
    w, h, wavelet, q, g, seed, size, blob_sha, _ = KATS_BIG[3]
    img = ol.synth(orc, w, h, seed)
    blob, st = ako_b200.encode(img, S(wavelet=wavelet, q=q, g=g))
    assert st == 0 and len(blob) == size
    assert sha(blob) == blob_sha
    out, st, _ = ako_b200.decode(blob)
    assert st == 0 and np.array_equal(out, img)


RATIO_CASES = [  # (w, h, seed, ratio, settings) -- the CPU twin of this list pins the oracle to the reference
    (320, 200, 11, 10, dict(wavelet=0, g=0)),
    (320, 200, 11, 30, dict(wavelet=1, g=8)),
    (257, 131, 12, 6, dict(wavelet=2, g=0)),
    (300, 260, 13, 20, dict(wavelet=0, g=0, tiles=128)),
    (200, 160, 14, 4000, dict(wavelet=1, g=0)),
    (200, 160, 14, 1, dict(wavelet=0, q=40, g=5)),
    (200, 160, 14, 0, dict(wavelet=0, q=12, g=0)),
    (128, 128, 15, 12, dict(wavelet=0, color=1, g=0)),
    (1296, 1040, 16, 15, dict(wavelet=0, wrap=2, g=4)),  # REPEAT at a size where levels are strip + frame on the GPU
    (1632, 2464, 2, 25, dict(wavelet=0, g=16)),  # configs[1] shape
]


@pytest.mark.parametrize("case", RATIO_CASES, ids=lambda c: f"{c[0]}x{c[1]}-r{c[3]}")
def test_ratio_search(orc, case):
    """akoB200EncodeRatio against the oracle's EncodePass (tools/akoenc.cpp:111-213): same blob, same quantisation,
    same pass count -- while running the wavelet once per colour model instead of once per pass."""
    w, h, seed, ratio, kw = case
    img = ol.synth(orc, w, h, seed)
    want, want_st, want_q, want_passes = ol.orc_encode_pass(orc, img, ratio, **kw)
    alias = {"tiles": "tiles_dimension"}
    s = S(**{alias.get(k, k): v for k, v in kw.items()})
    got, st, q, passes = ako_b200.encode_ratio(img, ratio, s)
    assert got == want
    assert (q, passes) == (want_q, want_passes)
    if got is not None:
        px, st, _ = ako_b200.decode(got)
        want_px, _ = ol.orc_decode(orc, want)
        assert st == 0 and np.array_equal(px, want_px)


@pytest.mark.parametrize("n", [1, 2, 5, 19])
def test_host_batch_api(orc, n):
    """akoB200EncodeBatch / akoB200DecodeBatch (host pointers, chunks pipelined over several streams): per image the
    bytes akoEncodeExt / akoDecodeExt return, i.e. the oracle's."""
    w, h = 200, 152
    imgs = [ol.synth(orc, w, h, 40 + i) for i in range(n)]
    for kw in (dict(wavelet=0, q=16, g=16), dict(wavelet=1, q=0, g=0, tiles=64)):
        alias = {"tiles": "tiles_dimension"}
        s = S(**{alias.get(k, k): v for k, v in kw.items()})
        blobs, st, done = ako_b200.encode_batch(imgs, s)
        assert (st, done) == (0, n)
        want = [ol.orc_encode(orc, im, **kw)[0] for im in imgs]
        assert blobs == want
        px, st, done = ako_b200.decode_batch(blobs)
        assert (st, done) == (0, n)
        for i in range(n):
            assert np.array_equal(px[i], ol.orc_decode(orc, want[i])[0])


@pytest.mark.parametrize("bands", [2, 3, 7])
def test_tiled_image_over_devices(orc, bands, monkeypatch):
    """One tiled image cut into bands of tile rows, one per entry of AKO_CUDA_DEVICES (akoEncodeExt / akoDecodeExt;
    encode.c:115-205, decode.c:113-230: tiles are independent blocks): the same .ako bytes and pixels as from one
    device. On a one-GPU box the bands share device 0 (one pooled context each); with more GPUs they spread."""
    import torch
    ndev = max(1, torch.cuda.device_count())
    monkeypatch.setenv("AKO_CUDA_DEVICES", ",".join(str(i % ndev) for i in range(bands)))
    monkeypatch.setenv("AKO_B200_BANDS_MIN_PIXELS", "1")
    made = 0
    for (w, h, ch, tiles) in [(600, 500, 4, 64), (333, 222, 3, 32), (256, 1024, 4, 128), (200, 130, 1, 64)]:
        img = ol.synth(orc, w, h, 11 + w)[..., :ch].copy()
        for kw in (dict(wavelet=W_DD137, q=16, g=16), dict(wavelet=W_CDF53, q=0, g=0), dict(wavelet=W_HAAR, q=4, g=0, compression=2)):
            made += _e2e(orc, img, tiles=tiles, **kw) is not None  # (an incompressible tile fails both encoders alike)
    assert made >= 9
    # a blob whose last band is cut short: the same refusal as from one device
    img = ol.synth(orc, 600, 500, 3)
    blob, _ = ol.orc_encode(orc, img, wavelet=W_CDF53, q=8, tiles=64)
    px, st, _ = ako_b200.decode(blob[:len(blob) - 40])
    assert px is None and st == 15


def test_host_batch_api_failures(orc):
    imgs = [ol.synth(orc, 96, 80, 60 + i) for i in range(12)]
    blobs, st, done = ako_b200.encode_batch(imgs, S(wavelet=1, q=8))
    assert (st, done) == (0, 12)
    bad = list(blobs)
    bad[9] = bad[9][:len(bad[9]) // 2]                      # truncated: AKO_BROKEN_INPUT
    px, st, done = ako_b200.decode_batch(bad)
    assert done == 9 and st == 15 and px[9] is None
    assert all(np.array_equal(px[i], ol.orc_decode(orc, blobs[i])[0]) for i in range(9))
    other, _ = ol.orc_encode(orc, ol.synth(orc, 64, 64, 1), wavelet=1, q=8)
    bad = list(blobs)
    bad[3] = other                                          # another shape: refused, the rest of the batch is unaffected
    px, st, done = ako_b200.decode_batch(bad)
    assert done == 3 and st == 9 and px[3] is None and px[11] is not None
    # encode: invalid settings are refused before any work, like akoEncodeExt
    blobs, st, done = ako_b200.encode_batch(imgs, S(tiles_dimension=100))
    assert done == 0 and st == 4 and all(b is None for b in blobs)
    # an incompressible image fails alone (AKO_ERROR), its neighbours encode
    noise = [noise_image(96, 80, 4, 5) if i == 4 else imgs[i] for i in range(12)]
    blobs, st, done = ako_b200.encode_batch(noise, S(wavelet=2, q=0, g=0))
    ref_blob, ref_st = ol.orc_encode(orc, noise[4], wavelet=2, q=0, g=0)
    if ref_blob is None:
        assert done == 4 and st == ref_st and blobs[4] is None and blobs[0] is not None
    else:
        assert done == 12


def test_one_process_spreads_a_batch_over_the_listed_devices(orc):
    """$AKO_CUDA_DEVICES = "0,1": akoB200EncodeBatch / akoB200DecodeBatch deal their chunks over both GPUs, and
    concurrent akoEncodeExt callers are dealt over them in turn; every result is what one GPU gives. Needs two GPUs."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    old = os.environ.get("AKO_CUDA_DEVICES")
    os.environ["AKO_CUDA_DEVICES"] = "0,1"
    try:
        imgs = [ol.synth(orc, 200, 136, 90 + i) for i in range(37)]
        want = [ol.orc_encode(orc, im, wavelet=1, q=16, g=4)[0] for im in imgs]
        s = S(wavelet=1, quantization=16, gate=4)
        blobs, st, done = ako_b200.encode_batch(imgs, s)
        assert (st, done) == (0, 37) and blobs == want
        px, st, done = ako_b200.decode_batch(blobs)
        assert (st, done) == (0, 37)
        assert all(np.array_equal(px[i], ol.orc_decode(orc, want[i])[0]) for i in range(37))
        for i in range(6):  # single calls alternate between the devices
            assert ako_b200.encode(imgs[i], s)[0] == want[i]
    finally:
        if old is None:
            os.environ.pop("AKO_CUDA_DEVICES", None)
        else:
            os.environ["AKO_CUDA_DEVICES"] = old


def test_pooled_contexts_follow_the_device_not_the_thread(orc):
    """CUDA's current device is per thread; pooled contexts are not. With $AKO_CUDA_DEVICE naming a device other than
    0, calls from fresh threads (whose current device is 0) and the library's own worker threads must still run on the
    context's device. Needs two GPUs (skipped on the one-GPU box)."""
    import threading
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    old = os.environ.get("AKO_CUDA_DEVICE")
    os.environ["AKO_CUDA_DEVICE"] = "1"
    try:
        imgs = [ol.synth(orc, 160, 120, 70 + i) for i in range(20)]
        want = [ol.orc_encode(orc, im, wavelet=0, q=16, g=0)[0] for im in imgs]
        s = S(wavelet=0, quantization=16, gate=0)
        assert ako_b200.encode(imgs[0], s)[0] == want[0]          # creates a pooled context on device 1, main thread
        results = {}

        def fresh_thread(i):
            results[i] = ako_b200.encode(imgs[i], s)[0]           # pooled context, thread whose current device is 0
        for i in range(1, 4):
            t = threading.Thread(target=fresh_thread, args=(i,))
            t.start()
            t.join()
        assert all(results[i] == want[i] for i in range(1, 4))
        blobs, st, done = ako_b200.encode_batch(imgs, s)          # the library's own worker threads
        assert (st, done) == (0, 20) and blobs == want
        px, st, done = ako_b200.decode_batch(blobs)
        assert (st, done) == (0, 20)
        assert all(np.array_equal(px[i], ol.orc_decode(orc, want[i])[0]) for i in range(20))
    finally:
        if old is None:
            os.environ.pop("AKO_CUDA_DEVICE", None)
        else:
            os.environ["AKO_CUDA_DEVICE"] = old


@pytest.mark.parametrize("seed", range(8))
def test_randomised_settings_sweep(orc, seed):
    """Seeded random walk over the whole settings space (shape, channels, wavelet, wrap, colour, q, gate, chroma loss,
    discard, tiles, image statistics) at sizes the oracle does in milliseconds: .ako bytes and decoded pixels against
    the oracle, case by case. Sizes straddle the kernels' switch points (strip / tile / small-pyramid kernels, widths
    whose halves are 0 and 4 mod 8, full and partial 2048-value Kagari blocks)."""
    rs = np.random.RandomState(9000 + seed)
    widths = [8, 9, 24, 40, 63, 64, 72, 130, 200, 264, 408, 520, 528, 1048]
    for case in range(14):
        w = int(rs.choice(widths)) + int(rs.randint(0, 3)) * int(rs.randint(0, 2))
        h = int(rs.choice(widths[:11])) + int(rs.randint(0, 3))
        ch = int(rs.choice([1, 2, 3, 4, 4, 4, 5, 8]))
        if w * h * ch > 1_200_000:
            h = max(8, 1_200_000 // (w * ch))
        kind = rs.randint(0, 4)
        if kind == 0:
            img = noise_image(w, h, ch, seed * 100 + case)
        elif kind == 1:
            img = smooth_image(w, h, ch, seed * 100 + case)
        elif kind == 2:
            img = np.full((h, w, ch), int(rs.randint(0, 256)), np.uint8)      # one long run per plane
            img[rs.randint(0, h), rs.randint(0, w)] ^= 0x55
        else:
            img = np.ascontiguousarray(ol.synth(orc, w, h, seed * 100 + case)[:, :, :min(ch, 4)])
            ch = img.shape[2]
        lossless = rs.rand() < 0.3
        kw = dict(wavelet=int(rs.choice([W_DD137, W_CDF53, W_HAAR])), wrap=int(rs.randint(0, 4)),
                  color=int(rs.choice([C_YCOCG, C_SUBG, C_NONE])), q=0 if lossless else int(rs.choice([1, 3, 16, 60, 400])),
                  g=0 if lossless else int(rs.choice([0, 0, 4, 16, 90])), chroma_loss=int(rs.randint(0, 4)),
                  discard=int(rs.randint(0, 2)))
        tiles = int(rs.choice([0, 0, 0, 8, 16, 64, 128]))
        if tiles and not (0 < w % tiles < 3 or 0 < h % tiles < 3):
            kw["tiles"] = tiles
        blob = _e2e(orc, img, **kw)
        if blob is not None and lossless and kw["discard"] == 0:
            out, st, _ = ako_b200.decode(blob)
            assert st == 0 and np.array_equal(out, img), (w, h, ch, kw)


def _corrupt(rs, blob, mode):
    b = bytearray(blob)
    if mode == 0:      # flip a few body bits
        for _ in range(rs.randint(1, 4)):
            b[rs.randint(16, len(b))] ^= 1 << rs.randint(0, 8)
    elif mode == 1:    # truncate
        b = b[:rs.randint(16, len(b))]
    elif mode == 2:    # corrupt the first block's size field
        b[16 + rs.randint(0, 4)] ^= 1 << rs.randint(0, 8)
    else:              # overwrite a span
        a = rs.randint(20, len(b) - 4)
        n = rs.randint(1, 16)
        b[a:a + n] = bytes(rs.randint(0, 256, size=n).astype(np.uint8))
    return bytes(b)


def test_corrupted_blobs_decode_like_the_oracle(orc):
    """Malformed input (decode.c:145-156, compression.c:58-73, kagari.c:301-366): on the SAME corrupted bytes the
    CUDA decoder returns the oracle's status, and the oracle's pixels whenever that status is OK. (The oracle's own
    agreement with the reference on corrupted input is pinned on the CPU: test_corrupted_blobs_oracle_vs_reference.)"""
    rs = np.random.RandomState(123)
    for (w, h, kw) in [(96, 80, dict(wavelet=0, q=16, g=0)), (200, 131, dict(wavelet=1, q=0, g=0)),
                       (64, 64, dict(wavelet=2, q=8, g=4, tiles=32)), (333, 40, dict(wavelet=0, q=3, g=0, wrap=2))]:
        blob, _ = ol.orc_encode(orc, ol.synth(orc, w, h, w), **kw)
        for trial in range(48):
            bad = _corrupt(rs, blob, trial % 4)
            got, gst, _ = ako_b200.decode(bad)
            want, wst = ol.orc_decode(orc, bad)
            assert gst == wst, (w, h, trial, gst, wst)
            if want is not None:
                assert np.array_equal(got, want), (w, h, trial)


def test_encode_accepts_a_device_pointer(orc):
    """akoEncodeExt copies its input with cudaMemcpyDefault: under unified addressing `in` may just as well be device
    memory (the blob still comes back through callbacks->malloc). akoDecodeExt parses the container on the host, so its
    input stays host memory; akoB200DecodeDevice is the entry point for device-resident blobs."""
    import torch
    w, h = 200, 152
    img = ol.synth(orc, w, h, 3)
    dev = torch.from_numpy(img).cuda()
    L = ako_b200.load()
    s = S(wavelet=0, quantization=16, gate=0)
    out, st = C.c_void_p(), C.c_int(0)
    n = L.akoEncodeExt(None, C.byref(s), 4, w, h, dev.data_ptr(), C.byref(out), C.byref(st))
    assert n and st.value == 0
    blob = C.string_at(out.value, n)
    L.akoDefaultFree(out)
    want, _ = ol.orc_encode(orc, img, wavelet=0, q=16, g=0)
    assert blob == want
