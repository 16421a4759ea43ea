"""ako_b200 -- B200-native (sm_100a) implementation of the Ako codec's encode/decode hot path.

The product is the C-ABI shared library ``ako_b200/libako_b200.so`` (sources in ``ako_b200/csrc``, public
headers in ``include/``). This package is a thin ctypes mirror of that ABI for tests and benchmarks; it adds
no functionality and has no CPU fallback: every call runs the CUDA kernels or raises.
"""
from .lib import (AkoError, AkoSettings, Context, STATUS, build, decode, decode_batch, default_settings, encode,
                  encode_batch, encode_ratio, lib_path, load, status_string)

__all__ = ["AkoError", "AkoSettings", "Context", "STATUS", "build", "decode", "decode_batch", "default_settings",
           "encode", "encode_batch", "encode_ratio", "lib_path", "load", "status_string"]
