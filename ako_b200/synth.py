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


def synth_rgba8_torch(w, h, seeds, device="cuda", y0=0):
    """The same generator as synth_rgba8 on a torch device, for the large batches of bench.py (configs[3]: 4096
    images, configs[4]: 16384x16384): (len(seeds), h, w, 4) uint8, rows y0 .. y0+h of the image (large images are
    made band by band). uint32 arithmetic is carried in int64 and masked; tests/test_abi_cpu.py checks it against
    synth_rgba8."""
    import torch

    M = 0xFFFFFFFF

    def mix(v):
        v = v ^ (v >> 16)
        v = (v * 0x7FEB352D) & M
        v = v ^ (v >> 15)
        v = (v * 0x846CA68B) & M  # the int64 product wraps; its low 32 bits are exact
        return v ^ (v >> 16)

    def tri(t, period):
        p = t % period
        return torch.where(p < period // 2, p, period - p)

    i64 = dict(dtype=torch.int64, device=device)
    S = torch.as_tensor(list(seeds), **i64).view(-1, 1, 1)
    X = torch.arange(w, **i64).view(1, 1, w)
    Y = torch.arange(y0, y0 + h, **i64).view(1, h, 1)
    n = mix(((X * 0x9E3779B1) & M) ^ mix((Y + ((S * 0x85EBCA6B) & M)) & M))
    n0, n1, n2 = (n & 7) - 4, ((n >> 8) & 7) - 4, ((n >> 16) & 7) - 4
    blk = mix(((((X >> 6) * 73856093) & M) ^ (((Y >> 6) * 19349663) & M)) ^ S) & 63
    r = tri((X + 3 * S) & M, 509) * 255 // 254
    g = tri((Y + 5 * S) & M, 383) * 255 // 191
    b = tri((X + Y) & M, 251) * 255 // 125
    out = torch.empty((S.shape[0], h, w, 4), dtype=torch.uint8, device=device)
    out[..., 0] = (r // 2 + blk + 32 + n0).clamp_(0, 255)
    out[..., 1] = (g // 2 + blk + 32 + n1).clamp_(0, 255)
    out[..., 2] = (b // 2 + (63 - blk) + 32 + n2).clamp_(0, 255)
    alpha_on = ((((X >> 7) + (Y >> 7) + S) & M) % 5) == 0
    out[..., 3] = torch.where(alpha_on, (tri(X, 128) * 4).clamp(0, 255).expand(S.shape[0], h, w), 255)
    return out
