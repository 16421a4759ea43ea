"""the seeded, integer-only synthetic RGBA8 image every measurement of this repo uses
(periodic ramps + 64x64 block offsets + +-4 hash noise + structured alpha). numpy restatement for bench.py, so that
generating the inputs does not go through the oracle; tests/test_abi_cpu.py checks it against the oracle's copy and
the survey's SHA-256."""
import numpy as np


def _mix(v):
    v = v.astype(np.uint32)
    v ^= v >> np.uint32(16)
    v = (v * np.uint32(0x7FEB352D)).astype(np.uint32)
    v ^= v >> np.uint32(15)
    v = (v * np.uint32(0x846CA68B)).astype(np.uint32)
    v ^= v >> np.uint32(16)
    return v


def _tri(t, period):
    p = (t.astype(np.uint32) % np.uint32(period)).astype(np.int64)
    h = period // 2
    return np.where(p < h, p, period - p)


def synth_rgba8(w, h, seed):
    """(h, w, 4) uint8, identical to orc_synth_rgba8(w, h, seed)."""
    with np.errstate(over="ignore"):
        seed = np.uint32(seed)
        X = np.arange(w, dtype=np.uint32)[None, :]
        Y = np.arange(h, dtype=np.uint32)[:, None]
        n = _mix((X * np.uint32(0x9E3779B1)) ^ _mix(Y + seed * np.uint32(0x85EBCA6B)))
        n0 = (n & 7).astype(np.int64) - 4
        n1 = ((n >> 8) & 7).astype(np.int64) - 4
        n2 = ((n >> 16) & 7).astype(np.int64) - 4
        blk = (_mix(((X >> 6) * np.uint32(73856093)) ^ ((Y >> 6) * np.uint32(19349663)) ^ seed) & 63).astype(np.int64)
        r = _tri(X + np.uint32(3) * seed, 509) * 255 // 254
        g = _tri(Y + np.uint32(5) * seed, 383) * 255 // 191
        b = _tri(X + Y, 251) * 255 // 125
        out = np.empty((h, w, 4), np.uint8)
        out[..., 0] = np.clip(r // 2 + blk + 32 + n0, 0, 255)
        out[..., 1] = np.clip(g // 2 + blk + 32 + n1, 0, 255)
        out[..., 2] = np.clip(b // 2 + (63 - blk) + 32 + n2, 0, 255)
        alpha_on = (((X >> 7) + (Y >> 7) + seed) % np.uint32(5)) == 0
        out[..., 3] = np.where(alpha_on, np.clip(_tri(np.broadcast_to(X, (h, w)), 128) * 4, 0, 255), 255)
    return out
