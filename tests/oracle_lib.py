"""ctypes bindings for the CHECKERS (test infrastructure only).

  * ``orc``  -- oracle/libako_oracle.so, our CPU restatement (oracle/ako_oracle.c)
  * ``ref``  -- oracle/_ref/libako_ref.so, the unmodified reference compiled from
               /root/reference/library (built by oracle/Makefile; may be absent)

Nothing under ako_b200/ imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")

c_size_t, c_int, c_void_p = C.c_size_t, C.c_int, C.c_void_p
u8p = C.POINTER(C.c_uint8)
i16p = C.POINTER(C.c_int16)


class OrcSettings(C.Structure):
    _fields_ = [("wavelet", c_int), ("color", c_int), ("wrap", c_int), ("compression", c_int),
                ("tiles_dimension", C.c_uint64), ("quantization", c_int), ("gate", c_int),
                ("chroma_loss", c_int), ("discard_non_visible", c_int)]


class AkoSettings(C.Structure):  # library/ako.h:86-99
    _fields_ = [("wavelet", c_int), ("color", c_int), ("wrap", c_int), ("compression", c_int),
                ("tiles_dimension", c_size_t), ("quantization", c_int), ("gate", c_int),
                ("chroma_loss", c_int), ("discard_non_visible", c_int)]


def _p(a, t):
    return a.ctypes.data_as(t)


def build_oracle():
    """(Re)build the oracle .so files if sources are newer / missing."""
    subprocess.run(["make", "-C", ORACLE_DIR, "-s"], check=True, stdout=subprocess.DEVNULL,
                   stderr=subprocess.DEVNULL)


def load_oracle():
    path = os.path.join(ORACLE_DIR, "libako_oracle.so")
    if not os.path.exists(path):
        build_oracle()
    L = C.CDLL(path)
    L.orc_synth_rgba8.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, u8p]
    L.orc_half.restype = c_size_t
    L.orc_half.argtypes = [c_size_t]
    L.orc_tile_data_size.restype = c_size_t
    L.orc_tile_data_size.argtypes = [c_size_t, c_size_t]
    L.orc_levels.restype = c_size_t
    L.orc_levels.argtypes = [c_size_t, c_size_t]
    L.orc_tile_dimension.restype = c_size_t
    L.orc_tile_dimension.argtypes = [c_size_t] * 3
    L.orc_tiles_no.restype = c_size_t
    L.orc_tiles_no.argtypes = [c_size_t] * 3
    for f in (L.orc_quantization, L.orc_gate):
        f.restype = C.c_int16
        f.argtypes = [c_int, c_int, c_size_t, c_size_t, c_size_t, c_size_t]
    L.orc_format_forward.argtypes = [c_int, c_int, c_size_t, c_size_t, c_size_t, c_size_t, u8p, i16p]
    L.orc_format_inverse.argtypes = [c_int, c_size_t, c_size_t, c_size_t, c_size_t, i16p, u8p]
    L.orc_lift_1d.argtypes = [c_int, c_int, c_size_t, i16p, c_size_t, i16p, i16p, c_size_t]
    L.orc_unlift_1d.argtypes = [c_int, c_int, c_size_t, i16p, i16p, c_size_t, i16p, c_size_t]
    L.orc_lift.argtypes = [C.POINTER(OrcSettings), c_size_t, c_size_t, c_size_t, i16p, i16p]
    L.orc_unlift.argtypes = [C.POINTER(OrcSettings), c_size_t, c_size_t, c_size_t, i16p, i16p]
    L.orc_kagari_encode.restype = c_size_t
    L.orc_kagari_encode.argtypes = [c_size_t, i16p, c_size_t, u8p]
    L.orc_kagari_decode.restype = c_size_t
    L.orc_kagari_decode.argtypes = [c_size_t, c_size_t, u8p, i16p]
    L.orc_kagari_bits.restype = C.c_uint64
    L.orc_kagari_bits.argtypes = [c_size_t, i16p]
    L.orc_head_write.argtypes = [c_size_t, c_size_t, c_size_t, C.POINTER(OrcSettings), u8p]
    L.orc_head_read.argtypes = [u8p, C.POINTER(c_size_t), C.POINTER(c_size_t), C.POINTER(c_size_t),
                                C.POINTER(OrcSettings)]
    L.orc_encode_bound.restype = c_size_t
    L.orc_encode_bound.argtypes = [c_size_t] * 3
    L.orc_encode.restype = c_size_t
    L.orc_encode.argtypes = [C.POINTER(OrcSettings), c_size_t, c_size_t, c_size_t, u8p, u8p, C.POINTER(c_int)]
    L.orc_decode.argtypes = [c_size_t, u8p, u8p, C.POINTER(OrcSettings)]
    L.orc_encode_pass.restype = c_size_t
    L.orc_encode_pass.argtypes = [c_int, C.POINTER(OrcSettings), c_size_t, c_size_t, c_size_t, u8p, u8p,
                                  C.POINTER(c_int), C.POINTER(c_size_t), C.POINTER(c_int)]
    return L


def ref_path():
    return os.path.join(ORACLE_DIR, "_ref", "libako_ref.so")


def load_ref():
    """The compiled reference, or None when it was never built (no /root/reference)."""
    path = ref_path()
    if not os.path.exists(path):
        try:
            build_oracle()
        except Exception:
            return None
    if not os.path.exists(path):
        return None
    L = C.CDLL(path)
    L.akoDefaultSettings.restype = AkoSettings
    L.akoEncodeExt.restype = c_size_t
    L.akoEncodeExt.argtypes = [c_void_p, C.POINTER(AkoSettings), c_size_t, c_size_t, c_size_t, c_void_p,
                               C.POINTER(c_void_p), C.POINTER(c_int)]
    L.akoDecodeExt.restype = c_void_p
    L.akoDecodeExt.argtypes = [c_void_p, c_size_t, c_void_p, C.POINTER(AkoSettings), C.POINTER(c_size_t),
                               C.POINTER(c_size_t), C.POINTER(c_size_t), C.POINTER(c_int)]
    L.akoDefaultFree.argtypes = [c_void_p]
    L.akoTileDataSize.restype = c_size_t
    L.akoTileDataSize.argtypes = [c_size_t, c_size_t]
    L.akoPlanesSpacing.restype = c_size_t
    L.akoPlanesSpacing.argtypes = [c_size_t, c_size_t]
    for f in (L.akoQuantization, L.akoGate):
        f.restype = C.c_int16
        f.argtypes = [c_int, c_int, c_size_t, c_size_t, c_size_t, c_size_t]
    L.akoKagariEncode.restype = c_size_t
    L.akoKagariEncode.argtypes = [c_size_t, c_size_t, c_void_p, c_void_p]
    L.akoKagariDecode.restype = c_size_t
    L.akoKagariDecode.argtypes = [c_size_t, c_size_t, c_size_t, c_void_p, c_void_p]
    L.akoFormatToPlanarI16Yuv.argtypes = [c_int, c_int, c_size_t, c_size_t, c_size_t, c_size_t, c_size_t,
                                          c_void_p, c_void_p]
    L.akoFormatToInterleavedU8Rgb.argtypes = [c_int, c_size_t, c_size_t, c_size_t, c_size_t, c_size_t,
                                              c_void_p, c_void_p]
    L.akoLift.argtypes = [c_size_t, C.POINTER(AkoSettings), c_size_t, c_size_t, c_size_t, c_size_t, c_void_p,
                          c_void_p]
    L.akoUnlift.argtypes = [C.POINTER(AkoSettings), c_size_t, c_size_t, c_size_t, c_size_t, c_size_t,
                            c_void_p, c_void_p]
    for name in ("akoCdf53LiftH", "akoDd137LiftH"):
        getattr(L, name).argtypes = [c_int, c_size_t, c_size_t, c_size_t, c_size_t, c_void_p, c_void_p]
    for name in ("akoCdf53UnliftH", "akoDd137UnliftH"):
        getattr(L, name).argtypes = [c_int, c_size_t, c_size_t, c_size_t, c_size_t, c_void_p, c_void_p, c_void_p]
    L.akoHaarLiftH.argtypes = [c_size_t, c_size_t, c_size_t, c_size_t, c_void_p, c_void_p]
    L.akoHaarUnliftH.argtypes = [c_size_t, c_size_t, c_size_t, c_size_t, c_void_p, c_void_p, c_void_p]
    return L


# ---------------------------------------------------------------- helpers


def make_settings(cls, wavelet=0, color=0, wrap=0, compression=0, tiles=0, q=16, g=0, chroma_loss=1, discard=0):
    s = cls()
    s.wavelet, s.color, s.wrap, s.compression = wavelet, color, wrap, compression
    s.tiles_dimension = tiles
    s.quantization, s.gate, s.chroma_loss, s.discard_non_visible = q, g, chroma_loss, discard
    return s


def synth(orc, w, h, seed):
    img = np.empty((h, w, 4), dtype=np.uint8)
    orc.orc_synth_rgba8(w, h, seed, _p(img, u8p))
    return img


def orc_encode(orc, img, **kw):
    h, w, ch = img.shape
    s = make_settings(OrcSettings, **kw)
    out = np.zeros(orc.orc_encode_bound(ch, w, h), dtype=np.uint8)
    st = c_int(0)
    n = orc.orc_encode(C.byref(s), ch, w, h, _p(np.ascontiguousarray(img), u8p), _p(out, u8p), C.byref(st))
    return (out[:n].tobytes() if n else None), st.value


def orc_encode_pass(orc, img, ratio, **kw):
    """EncodePass of the encoder tool: (blob | None, status, q of the blob, passes)."""
    h, w, ch = img.shape
    s = make_settings(OrcSettings, **kw)
    out = np.zeros(orc.orc_encode_bound(ch, w, h), dtype=np.uint8)
    st, q, passes = c_int(0), c_int(0), c_size_t(0)
    n = orc.orc_encode_pass(ratio, C.byref(s), ch, w, h, _p(np.ascontiguousarray(img), u8p), _p(out, u8p),
                            C.byref(q), C.byref(passes), C.byref(st))
    return (out[:n].tobytes() if n else None), st.value, q.value, passes.value


def ref_encode_pass(ref, img, ratio, **kw):
    """EncodePass (tools/akoenc.cpp:111-213) transcribed over the UNMODIFIED reference's akoEncodeExt: pins
    orc_encode_pass. Returns (blob | None, q of the blob, passes)."""
    passes = [0]

    def enc(q, g=None):
        k = dict(kw)
        k["q"] = q
        if g is not None:
            k["g"] = g
        passes[0] += 1
        blob, _ = ref_encode(ref, img, **k)
        return blob

    size = lambda b: len(b) if b else 0
    if ratio == 0 or kw.get("wavelet", 0) == 3 or kw.get("compression", 0) == 2:
        return enc(kw.get("q", 16)), kw.get("q", 16), 1
    if ratio == 1:
        return enc(0, 0), 0, 1
    h, w, ch = img.shape
    target = (w * h * ch) // ratio
    margin = (target * 4) // 100
    last = enc(0)
    last_q = 0
    ceil_size = size(last)
    q, floor_size, floor_q, ceil_q = 1, ceil_size, 0, 0
    while True:
        q *= 4
        ceil_size, ceil_q = floor_size, floor_q
        last, last_q = enc(q), q
        floor_size, floor_q = size(last), q
        if not (floor_size > target and q <= (1 << 28)):
            break
    last_size = floor_size
    while abs(floor_size - ceil_size) > margin and abs(floor_q - ceil_q) > 1:
        q = (ceil_q + floor_q) // 2
        last, last_q = enc(q), q
        last_size = size(last)
        if last_size > target:
            ceil_size, ceil_q = last_size, q
        else:
            floor_size, floor_q = last_size, q
    if abs(floor_size - target) < abs(ceil_size - target):
        chosen_size, chosen_q = floor_size, floor_q
    else:
        chosen_size, chosen_q = ceil_size, ceil_q
    if last_size == chosen_size:
        return last, last_q, passes[0]
    return enc(chosen_q), chosen_q, passes[0]


def orc_decode(orc, blob):
    b = np.frombuffer(blob, dtype=np.uint8).copy()
    ch, w, h = c_size_t(), c_size_t(), c_size_t()
    s = OrcSettings()
    st = orc.orc_head_read(_p(b, u8p), C.byref(ch), C.byref(w), C.byref(h), C.byref(s))
    if st != 0:
        return None, st
    out = np.zeros((h.value, w.value, ch.value), dtype=np.uint8)
    st = orc.orc_decode(len(b), _p(b, u8p), _p(out, u8p), C.byref(s))
    return (out if st == 0 else None), st


def ref_encode(ref, img, **kw):
    h, w, ch = img.shape
    s = make_settings(AkoSettings, **kw)
    out = c_void_p()
    st = c_int(0)
    img = np.ascontiguousarray(img)
    n = ref.akoEncodeExt(None, C.byref(s), ch, w, h, img.ctypes.data, C.byref(out), C.byref(st))
    if n == 0:
        return None, st.value
    blob = C.string_at(out.value, n)
    ref.akoDefaultFree(out)
    return blob, st.value


def ref_decode(ref, blob):
    ch, w, h = c_size_t(), c_size_t(), c_size_t()
    st = c_int(0)
    s = AkoSettings()
    buf = C.create_string_buffer(blob, len(blob))
    p = ref.akoDecodeExt(None, len(blob), buf, C.byref(s), C.byref(ch), C.byref(w), C.byref(h), C.byref(st))
    if not p:
        return None, st.value
    n = ch.value * w.value * h.value
    out = np.frombuffer(C.string_at(p, n), dtype=np.uint8).reshape(h.value, w.value, ch.value).copy()
    ref.akoDefaultFree(p)
    return out, st.value
