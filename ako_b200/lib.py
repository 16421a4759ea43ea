"""ctypes binding of libako_b200.so (include/ako.h + include/ako_b200.h).

Mirrors the reference's public interface (library/ako.h:130-145): ``encode`` == akoEncodeExt,
``decode`` == akoDecodeExt, same argument meaning, same status codes. ``Context`` exposes the additive
device-resident / batched / single-stage entry points. No oracle, no CPU path: if the library or a CUDA
device is missing the call fails loudly.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
c_size_t, c_int, c_void_p = C.c_size_t, C.c_int, C.c_void_p

STATUS = ["AKO_OK", "AKO_ERROR", "AKO_INVALID_CHANNELS_NO", "AKO_INVALID_DIMENSIONS", "AKO_INVALID_TILES_DIMENSIONS",
          "AKO_INVALID_WRAP_MODE", "AKO_INVALID_WAVELET_TRANSFORMATION", "AKO_INVALID_COLOR_TRANSFORMATION",
          "AKO_INVALID_COMPRESSION_METHOD", "AKO_INVALID_INPUT", "AKO_INVALID_CALLBACKS", "AKO_INVALID_MAGIC",
          "AKO_UNSUPPORTED_VERSION", "AKO_NO_ENOUGH_MEMORY", "AKO_INVALID_FLAGS", "AKO_BROKEN_INPUT"]


class AkoError(RuntimeError):
    def __init__(self, status, what=""):
        self.status = status
        name = STATUS[status] if 0 <= status < len(STATUS) else str(status)
        super().__init__(f"{what}: {name}" if what else name)


class AkoSettings(C.Structure):  # include/ako.h, struct akoSettings (40 bytes on LP64)
    _fields_ = [("wavelet", c_int), ("color", c_int), ("wrap", c_int), ("compression", c_int),
                ("tiles_dimension", c_size_t), ("quantization", c_int), ("gate", c_int),
                ("chroma_loss", c_int), ("discard_non_visible", c_int)]


class AkoCallbacks(C.Structure):  # include/ako.h, struct akoCallbacks (40 bytes on LP64)
    _fields_ = [("malloc", c_void_p), ("realloc", c_void_p), ("free", c_void_p), ("events", c_void_p),
                ("events_data", c_void_p)]


EVENTS_FN = C.CFUNCTYPE(None, c_size_t, c_size_t, c_int, c_void_p)


def lib_path():
    # AKO_B200_LIB: another build of the same library (A/B experiments of kernel variants)
    return os.environ.get("AKO_B200_LIB") or os.path.join(HERE, "libako_b200.so")


def build(verbose=False):
    """Compile libako_b200.so in-tree with nvcc for sm_100a (works without a GPU)."""
    out = subprocess.run(["make", "-C", os.path.join(HERE, "csrc")], capture_output=True, text=True)
    if out.returncode != 0:
        raise RuntimeError("building libako_b200.so failed:\n" + out.stdout + out.stderr)
    if verbose:
        print(out.stdout)
    return lib_path()


_LIB = None


def load():
    """The shared library; raises if it has not been built (there is no fallback implementation)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(lib_path()):
        raise RuntimeError(f"{lib_path()} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(ako_b200 has no CPU fallback)")
    L = C.CDLL(lib_path())
    sp = C.POINTER(AkoSettings)
    stp = C.POINTER(c_int)
    L.akoEncodeExt.restype = c_size_t
    L.akoEncodeExt.argtypes = [C.POINTER(AkoCallbacks), sp, c_size_t, c_size_t, c_size_t, c_void_p,
                               C.POINTER(c_void_p), stp]
    L.akoDecodeExt.restype = c_void_p
    L.akoDecodeExt.argtypes = [C.POINTER(AkoCallbacks), c_size_t, c_void_p, sp, C.POINTER(c_size_t),
                               C.POINTER(c_size_t), C.POINTER(c_size_t), stp]
    L.akoDefaultSettings.restype = AkoSettings
    L.akoDefaultCallbacks.restype = AkoCallbacks
    L.akoDefaultFree.argtypes = [c_void_p]
    L.akoStatusString.restype = C.c_char_p
    L.akoStatusString.argtypes = [c_int]
    for f in ("akoVersionMajor", "akoVersionMinor", "akoVersionPatch", "akoFormatVersion"):
        getattr(L, f).restype = c_int
    # extension
    L.akoB200ContextCreate.restype = c_void_p
    L.akoB200ContextCreate.argtypes = [c_int, stp]
    L.akoB200ContextDestroy.argtypes = [c_void_p]
    L.akoB200ContextStream.restype = c_void_p
    L.akoB200ContextStream.argtypes = [c_void_p]
    L.akoB200Synchronize.argtypes = [c_void_p]
    L.akoB200DeviceAlloc.restype = c_void_p
    L.akoB200DeviceAlloc.argtypes = [c_void_p, c_size_t]
    L.akoB200DeviceFree.argtypes = [c_void_p, c_void_p]
    L.akoB200PinnedAlloc.restype = c_void_p
    L.akoB200PinnedAlloc.argtypes = [c_size_t]
    L.akoB200PinnedFree.argtypes = [c_void_p]
    L.akoB200CopyToDevice.argtypes = [c_void_p, c_void_p, c_void_p, c_size_t]
    L.akoB200CopyToHost.argtypes = [c_void_p, c_void_p, c_void_p, c_size_t]
    L.akoB200PinnedCallbacks.restype = AkoCallbacks
    L.akoB200ProfileEnable.argtypes = [c_void_p, c_int]
    L.akoB200ProfileReset.argtypes = [c_void_p]
    L.akoB200ProfileGet.restype = c_size_t
    L.akoB200ProfileGet.argtypes = [c_void_p, c_size_t, C.POINTER(C.c_char_p), C.POINTER(C.c_uint64),
                                    C.POINTER(C.c_double)]
    L.akoB200ProfileGetBytes.restype = c_size_t
    L.akoB200ProfileGetBytes.argtypes = [c_void_p, c_size_t, C.POINTER(C.c_uint64)]
    L.akoB200LaunchCount.restype = C.c_uint64
    L.akoB200LaunchCount.argtypes = [c_void_p]
    L.akoB200EncodeBound.restype = c_size_t
    L.akoB200EncodeBound.argtypes = [sp, c_size_t, c_size_t, c_size_t]
    L.akoB200EncodeDevice.restype = c_size_t
    L.akoB200EncodeDevice.argtypes = [c_void_p, sp, c_size_t, c_size_t, c_size_t, c_void_p, c_void_p, c_size_t, stp]
    L.akoB200DecodeDevice.argtypes = [c_void_p, c_size_t, c_void_p, c_void_p, c_void_p, c_size_t, sp,
                                      C.POINTER(c_size_t), C.POINTER(c_size_t), C.POINTER(c_size_t)]
    L.akoB200EncodeBatchDevice.restype = c_size_t
    L.akoB200EncodeBatchDevice.argtypes = [c_void_p, sp, c_size_t, c_size_t, c_size_t, c_size_t, c_void_p, c_size_t,
                                           c_void_p, c_size_t, C.POINTER(c_size_t), stp]
    L.akoB200DecodeBatchDevice.restype = c_size_t
    L.akoB200DecodeBatchDevice.argtypes = [c_void_p, c_size_t, c_void_p, c_size_t, C.POINTER(c_size_t), c_void_p,
                                           c_size_t, stp]
    vpp = C.POINTER(c_void_p)
    L.akoB200EncodeBatch.restype = c_size_t
    L.akoB200EncodeBatch.argtypes = [C.POINTER(AkoCallbacks), sp, c_size_t, c_size_t, c_size_t, c_size_t, vpp, vpp,
                                     C.POINTER(c_size_t), stp]
    L.akoB200DecodeBatch.restype = c_size_t
    L.akoB200DecodeBatch.argtypes = [C.POINTER(AkoCallbacks), c_size_t, vpp, C.POINTER(c_size_t), vpp, sp,
                                     C.POINTER(c_size_t), C.POINTER(c_size_t), C.POINTER(c_size_t), stp]
    L.akoB200EncodeRatio.restype = c_size_t
    L.akoB200EncodeRatio.argtypes = [C.POINTER(AkoCallbacks), sp, c_int, c_size_t, c_size_t, c_size_t, c_void_p,
                                     C.POINTER(c_void_p), stp, C.POINTER(c_size_t), stp]
    L.akoB200StreamSize.restype = c_size_t
    L.akoB200StreamSize.argtypes = [c_size_t, c_size_t, c_size_t]
    L.akoB200FormatForward.argtypes = [c_void_p, sp, c_size_t, c_size_t, c_size_t, c_size_t, c_void_p, c_void_p]
    L.akoB200FormatInverse.argtypes = [c_void_p, c_int, c_size_t, c_size_t, c_size_t, c_size_t, c_void_p, c_void_p]
    L.akoB200Lift.argtypes = [c_void_p, sp, c_size_t, c_size_t, c_size_t, c_void_p, c_void_p]
    L.akoB200Unlift.argtypes = [c_void_p, sp, c_size_t, c_size_t, c_size_t, c_void_p, c_void_p]
    L.akoB200KagariEncode.restype = c_size_t
    L.akoB200KagariEncode.argtypes = [c_void_p, c_size_t, c_void_p, c_void_p, c_size_t, stp]
    L.akoB200KagariDecode.restype = c_size_t
    L.akoB200KagariDecode.argtypes = [c_void_p, c_size_t, c_size_t, c_void_p, c_void_p, stp]
    _LIB = L
    return L


def default_settings(**kw):
    """akoDefaultSettings() (misc.c:30-47) with keyword overrides, e.g. wavelet=1, quantization=0."""
    s = load().akoDefaultSettings()
    alias = {"q": "quantization", "g": "gate", "tiles": "tiles_dimension", "discard": "discard_non_visible"}
    for k, v in kw.items():
        setattr(s, alias.get(k, k), v)
    return s


def status_string(status):
    return load().akoStatusString(status).decode()


def encode(image, settings=None, callbacks=None):
    """akoEncodeExt on a host array of shape (h, w, channels) uint8. Returns (blob bytes | None, status).
    ``callbacks`` is an AkoCallbacks structure (or None for the defaults)."""
    L = load()
    image = np.ascontiguousarray(image, dtype=np.uint8)
    h, w, ch = image.shape
    out = c_void_p()
    st = c_int(0)
    n = L.akoEncodeExt(C.byref(callbacks) if callbacks is not None else None,
                       C.byref(settings) if settings is not None else None, ch, w, h,
                       image.ctypes.data, C.byref(out), C.byref(st))
    if n == 0:
        return None, st.value
    blob = C.string_at(out.value, n)
    free = C.CFUNCTYPE(None, c_void_p)(callbacks.free) if callbacks is not None else L.akoDefaultFree
    free(out)
    return blob, st.value


def encode_ratio(image, ratio, settings=None):
    """akoB200EncodeRatio == the encoder tool's EncodePass (tools/akoenc.cpp:111-213) on a host array.
    Returns (blob bytes | None, status, quantization of the blob, passes the tool would have made)."""
    L = load()
    image = np.ascontiguousarray(image, dtype=np.uint8)
    h, w, ch = image.shape
    out = c_void_p()
    st, q, passes = c_int(0), c_int(0), c_size_t(0)
    n = L.akoB200EncodeRatio(None, C.byref(settings) if settings is not None else None, ratio, ch, w, h,
                             image.ctypes.data, C.byref(out), C.byref(q), C.byref(passes), C.byref(st))
    if n == 0:
        return None, st.value, q.value, passes.value
    blob = C.string_at(out.value, n)
    L.akoDefaultFree(out)
    return blob, st.value, q.value, passes.value


def encode_batch(images, settings=None):
    """akoB200EncodeBatch on a list of same-shape (h, w, channels) uint8 arrays.
    Returns (list of blob bytes | None, status, leading images that succeeded)."""
    L = load()
    images = [np.ascontiguousarray(im, dtype=np.uint8) for im in images]
    n = len(images)
    h, w, ch = images[0].shape if n else (0, 0, 0)
    ins = (c_void_p * max(n, 1))(*[im.ctypes.data for im in images])
    outs = (c_void_p * max(n, 1))()
    sizes = (c_size_t * max(n, 1))()
    st = c_int(0)
    done = L.akoB200EncodeBatch(None, C.byref(settings) if settings is not None else None, ch, w, h, n, ins, outs, sizes,
                                C.byref(st))
    blobs = []
    for i in range(n):
        if outs[i]:
            blobs.append(C.string_at(outs[i], sizes[i]))
            L.akoDefaultFree(outs[i])
        else:
            blobs.append(None)
    return blobs, st.value, done


def decode_batch(blobs):
    """akoB200DecodeBatch on a list of blobs of one shape. Returns (list of images | None, status, leading successes)."""
    L = load()
    n = len(blobs)
    bufs = [np.frombuffer(b, dtype=np.uint8) for b in blobs]
    ins = (c_void_p * max(n, 1))(*[b.ctypes.data for b in bufs])
    sizes = (c_size_t * max(n, 1))(*[len(b) for b in blobs])
    outs = (c_void_p * max(n, 1))()
    s = AkoSettings()
    ch, w, h = c_size_t(), c_size_t(), c_size_t()
    st = c_int(0)
    done = L.akoB200DecodeBatch(None, n, ins, sizes, outs, C.byref(s), C.byref(ch), C.byref(w), C.byref(h), C.byref(st))
    images = []
    for i in range(n):
        if outs[i]:
            nb = ch.value * w.value * h.value
            images.append(np.frombuffer(C.string_at(outs[i], nb), dtype=np.uint8).reshape(h.value, w.value, ch.value).copy())
            L.akoDefaultFree(outs[i])
        else:
            images.append(None)
    return images, st.value, done


def decode(blob, callbacks=None):
    """akoDecodeExt on host bytes. Returns (image (h, w, channels) uint8 | None, status, settings)."""
    L = load()
    buf = np.frombuffer(blob, dtype=np.uint8)
    s = AkoSettings()
    ch, w, h = c_size_t(), c_size_t(), c_size_t()
    st = c_int(0)
    p = L.akoDecodeExt(C.byref(callbacks) if callbacks is not None else None, len(blob), buf.ctypes.data, C.byref(s),
                       C.byref(ch), C.byref(w), C.byref(h), C.byref(st))
    if not p:
        return None, st.value, None
    n = ch.value * w.value * h.value
    img = np.frombuffer(C.string_at(p, n), dtype=np.uint8).reshape(h.value, w.value, ch.value).copy()
    free = C.CFUNCTYPE(None, c_void_p)(callbacks.free) if callbacks is not None else L.akoDefaultFree
    free(p)
    return img, st.value, s


class Context:
    """akoB200Context: one CUDA stream + workspace on one device. Device pointers are plain ints."""

    def __init__(self, device=-1):
        self.L = load()
        st = c_int(0)
        self.h = self.L.akoB200ContextCreate(device, C.byref(st))
        if not self.h:
            raise AkoError(st.value, "akoB200ContextCreate (no usable CUDA device? ako_b200 has no CPU fallback)")

    def close(self):
        if self.h:
            self.L.akoB200ContextDestroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- memory
    def alloc(self, nbytes):
        p = self.L.akoB200DeviceAlloc(self.h, nbytes)
        if not p:
            raise AkoError(13, "akoB200DeviceAlloc")
        return p

    def free(self, p):
        self.L.akoB200DeviceFree(self.h, p)

    def to_device(self, arr):
        arr = np.ascontiguousarray(arr)
        p = self.alloc(max(arr.nbytes, 16) + 64)
        self._check(self.L.akoB200CopyToDevice(self.h, p, arr.ctypes.data, arr.nbytes), "CopyToDevice")
        self.sync()
        return p

    def to_host(self, p, shape, dtype):
        out = np.empty(shape, dtype=dtype)
        self._check(self.L.akoB200CopyToHost(self.h, out.ctypes.data, p, out.nbytes), "CopyToHost")
        self.sync()
        return out

    def sync(self):
        self._check(self.L.akoB200Synchronize(self.h), "Synchronize")

    @property
    def stream(self):
        return self.L.akoB200ContextStream(self.h)

    @staticmethod
    def _check(st, what):
        if st != 0:
            raise AkoError(st, what)

    # -- profiling
    def profile(self, enable=True):
        self.L.akoB200ProfileEnable(self.h, 1 if enable else 0)

    def profile_reset(self):
        self.L.akoB200ProfileReset(self.h)

    def profile_get(self):
        names = (C.c_char_p * 64)()
        launches = (C.c_uint64 * 64)()
        ms = (C.c_double * 64)()
        n = self.L.akoB200ProfileGet(self.h, 64, names, launches, ms)
        return {names[i].decode(): (int(launches[i]), float(ms[i])) for i in range(min(n, 64))}

    def profile_get_bytes(self):
        """kernel name -> algorithmic bytes its launches moved."""
        names = (C.c_char_p * 64)()
        launches = (C.c_uint64 * 64)()
        ms = (C.c_double * 64)()
        nbytes = (C.c_uint64 * 64)()
        n = self.L.akoB200ProfileGet(self.h, 64, names, launches, ms)
        self.L.akoB200ProfileGetBytes(self.h, 64, nbytes)
        return {names[i].decode(): int(nbytes[i]) for i in range(min(n, 64))}

    def launch_count(self):
        return int(self.L.akoB200LaunchCount(self.h))

    # -- whole codec, device resident
    def encode_bound(self, settings, ch, w, h):
        return self.L.akoB200EncodeBound(C.byref(settings), ch, w, h)

    def encode_device(self, settings, ch, w, h, d_in, d_out, cap):
        st = c_int(0)
        n = self.L.akoB200EncodeDevice(self.h, C.byref(settings), ch, w, h, d_in, d_out, cap, C.byref(st))
        return n, st.value

    def decode_device(self, size, d_in, d_out, cap, head16=None):
        s = AkoSettings()
        ch, w, h = c_size_t(), c_size_t(), c_size_t()
        st = self.L.akoB200DecodeDevice(self.h, size, d_in, head16, d_out, cap, C.byref(s), C.byref(ch), C.byref(w),
                                        C.byref(h))
        return st, (ch.value, w.value, h.value), s

    def encode_batch_device(self, settings, ch, w, h, n, d_in, in_stride, d_out, out_stride):
        sizes = (c_size_t * n)()
        st = c_int(0)
        done = self.L.akoB200EncodeBatchDevice(self.h, C.byref(settings), ch, w, h, n, d_in, in_stride, d_out,
                                               out_stride, sizes, C.byref(st))
        return done, st.value, list(sizes)

    def decode_batch_device(self, n, d_in, in_stride, sizes, d_out, out_stride):
        arr = (c_size_t * n)(*sizes)
        st = c_int(0)
        done = self.L.akoB200DecodeBatchDevice(self.h, n, d_in, in_stride, arr, d_out, out_stride, C.byref(st))
        return done, st.value

    # -- single stages (host numpy in/out convenience, used by the parity tests)
    def format_forward(self, img, settings, in_stride_px=None):
        h, w, ch = img.shape
        d_in = self.to_device(img)
        d_pl = self.alloc(ch * w * h * 2 + 64)
        self._check(self.L.akoB200FormatForward(self.h, C.byref(settings), ch, w, h, in_stride_px or w, d_in, d_pl),
                    "FormatForward")
        out = self.to_host(d_pl, (ch, h, w), np.int16)
        self.free(d_in)
        self.free(d_pl)
        return out

    def format_inverse(self, planes, color):
        ch, h, w = planes.shape
        d_pl = self.to_device(planes.astype(np.int16))
        d_out = self.alloc(ch * w * h + 64)
        self._check(self.L.akoB200FormatInverse(self.h, color, ch, w, h, w, d_pl, d_out), "FormatInverse")
        out = self.to_host(d_out, (h, w, ch), np.uint8)
        self.free(d_pl)
        self.free(d_out)
        return out

    def stream_size(self, ch, w, h):
        return self.L.akoB200StreamSize(ch, w, h)

    def lift(self, planes, settings):
        ch, h, w = planes.shape
        n = self.stream_size(ch, w, h) // 2
        d_pl = self.to_device(planes.astype(np.int16))
        d_st = self.alloc(n * 2 + 64)
        self._check(self.L.akoB200Lift(self.h, C.byref(settings), ch, w, h, d_pl, d_st), "Lift")
        out = self.to_host(d_st, (n,), np.int16)
        self.free(d_pl)
        self.free(d_st)
        return out

    def unlift(self, stream, settings, ch, w, h):
        d_st = self.to_device(stream.astype(np.int16))
        d_pl = self.alloc(ch * w * h * 2 + 64)
        self._check(self.L.akoB200Unlift(self.h, C.byref(settings), ch, w, h, d_st, d_pl), "Unlift")
        out = self.to_host(d_pl, (ch, h, w), np.int16)
        self.free(d_st)
        self.free(d_pl)
        return out

    def kagari_encode(self, values, cap=None):
        values = np.ascontiguousarray(values, dtype=np.int16)
        n = len(values)
        cap = cap if cap is not None else n * 4 + 64
        d_in = self.to_device(values)
        d_out = self.alloc(cap + 64)
        st = c_int(0)
        size = self.L.akoB200KagariEncode(self.h, n, d_in, d_out, cap, C.byref(st))
        self._check(st.value, "KagariEncode")
        out = self.to_host(d_out, (size,), np.uint8) if size else None
        self.free(d_in)
        self.free(d_out)
        return out

    def kagari_decode(self, data, n_values):
        data = np.ascontiguousarray(data, dtype=np.uint8)
        d_in = self.to_device(data)
        d_out = self.alloc(n_values * 2 + 64)
        st = c_int(0)
        used = self.L.akoB200KagariDecode(self.h, n_values, len(data), d_in, d_out, C.byref(st))
        self._check(st.value, "KagariDecode")
        out = self.to_host(d_out, (n_values,), np.int16)
        self.free(d_in)
        self.free(d_out)
        return used, out
