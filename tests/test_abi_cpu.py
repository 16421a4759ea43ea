"""CPU-only checks of the drop-in boundary: the library loads and exports every symbol that include/*.h
declares, with the reference's struct layouts. No compute calls (there is no CPU path to call)."""
import ctypes as C
import os
import re

import ako_b200
from ako_b200 import lib as akolib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

REFERENCE_EXPORTS = ["akoEncodeExt", "akoDecodeExt", "akoDefaultSettings", "akoDefaultCallbacks", "akoDefaultFree",
                     "akoStatusString", "akoVersionMajor", "akoVersionMinor", "akoVersionPatch", "akoFormatVersion"]


def _declared(header):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ako[A-Z]\w*)\s*\(", text)))


def test_library_builds_and_loads():
    ako_b200.build()
    assert os.path.exists(ako_b200.lib_path())
    ako_b200.load()


def test_every_declared_symbol_is_exported():
    L = ako_b200.load()
    names = _declared("ako.h") + _declared("ako_b200.h")
    assert set(REFERENCE_EXPORTS) <= set(names)
    assert len(names) > 30
    for n in names:
        assert hasattr(L, n), n


def test_struct_layouts_match_reference_abi():
    # library/ako.h:86-109 on LP64: 4 enums + size_t + 4 ints = 40 B; five pointers = 40 B
    assert C.sizeof(akolib.AkoSettings) == 40
    assert C.sizeof(akolib.AkoCallbacks) == 40
    assert akolib.AkoSettings.tiles_dimension.offset == 16
    assert akolib.AkoSettings.quantization.offset == 24


def test_defaults_strings_versions():
    L = ako_b200.load()
    s = L.akoDefaultSettings()  # misc.c:30-47
    assert (s.wavelet, s.color, s.wrap, s.compression, s.tiles_dimension) == (0, 0, 0, 0, 0)
    assert (s.quantization, s.gate, s.chroma_loss, s.discard_non_visible) == (16, 0, 1, 0)
    cb = L.akoDefaultCallbacks()
    assert cb.malloc and cb.realloc and cb.free and not cb.events
    assert (L.akoVersionMajor(), L.akoVersionMinor(), L.akoVersionPatch(), L.akoFormatVersion()) == (0, 2, 0, 2)
    assert ako_b200.status_string(0) == "Everything Ok!"
    assert ako_b200.status_string(15) == "Broken input/premature end"
    assert ako_b200.status_string(99) == "Unknown status code"


def test_status_strings_match_reference(ref):
    ref.akoStatusString.restype = C.c_char_p
    ref.akoStatusString.argtypes = [C.c_int]
    for i in range(-1, 18):
        assert ako_b200.status_string(i) == ref.akoStatusString(i).decode()


def test_argument_errors_need_no_device():
    """Checks that come before any GPU work (encode.c:53-82 order) answer with the reference's statuses."""
    import numpy as np
    img = np.zeros((8, 8, 4), np.uint8)
    L = ako_b200.load()
    st = C.c_int(0)
    out = C.c_void_p()
    s = ako_b200.default_settings()
    assert L.akoEncodeExt(None, C.byref(s), 4, 8, 8, None, C.byref(out), C.byref(st)) == 0 and st.value == 9
    bad = akolib.AkoCallbacks()
    assert L.akoEncodeExt(C.byref(bad), C.byref(s), 4, 8, 8, img.ctypes.data, C.byref(out), C.byref(st)) == 0
    assert st.value == 10
    for kw, want in ((dict(tiles=24), 4), (dict(tiles=4), 4), (dict(wrap=7), 5), (dict(wavelet=9), 6),
                     (dict(color=5), 7), (dict(compression=3), 8)):
        s2 = ako_b200.default_settings(**kw)
        assert L.akoEncodeExt(None, C.byref(s2), 4, 8, 8, img.ctypes.data, C.byref(out), C.byref(st)) == 0
        assert st.value == want, kw
    assert L.akoEncodeExt(None, C.byref(s), 17, 8, 8, img.ctypes.data, C.byref(out), C.byref(st)) == 0 and st.value == 2
    assert L.akoEncodeExt(None, C.byref(s), 4, 0, 8, img.ctypes.data, C.byref(out), C.byref(st)) == 0 and st.value == 3
    # decode: header checks
    for blob, want in ((b"Bko\x02" + b"\0" * 28, 11), (b"Ako\x03" + b"\0" * 28, 12),
                       (b"Ako\x02" + b"\x08\0\0\0" * 2 + b"\0\x80\0\0" + b"\0" * 16, 14),
                       (b"Ako\x02" + b"\0" * 28, 3), (b"Ako", 15)):
        img2, st2, _ = ako_b200.decode(blob)
        assert img2 is None and st2 == want, (blob[:16], st2)


def test_torch_synth_matches_numpy_synth():
    """bench.py makes its large batches (configs[3], configs[4]) with the torch restatement of SURVEY Appendix C."""
    import numpy as np
    from ako_b200.synth import synth_rgba8, synth_rgba8_torch
    for w, h, seeds in ((300, 200, [1, 2, 1000, 77777]), (520, 391, [5])):
        got = synth_rgba8_torch(w, h, seeds, device="cpu").numpy()
        for k, seed in enumerate(seeds):
            assert np.array_equal(got[k], synth_rgba8(w, h, seed)), (w, h, seed)
    band = synth_rgba8_torch(300, 50, [9], device="cpu", y0=100).numpy()[0]
    assert np.array_equal(band, synth_rgba8(300, 200, 9)[100:150])


def test_crafted_header_sizes_are_refused_before_any_device_work():
    """A header whose dimensions wrap size_t products, or exceed what the kernels index, must fail in the header
    checks (no context, no allocation): the reference overflows in the same place (decode.c:87-110)."""
    import struct
    for w, h, flags in ((1 << 31, 1 << 31, 3), (0xFFFFFFFF, 0xFFFFFFFF, 15), (1 << 20, 1 << 20, 3)):
        blob = b"Ako\x02" + struct.pack("<III", w, h, flags) + b"\0" * 64
        img, st, _ = ako_b200.decode(blob)
        assert img is None and st == 13, (w, h, st)
