"""Where the host-pointer API's time goes: per-call latency and thread scaling (configs[1] shape)."""
import sys, os, time, json, ctypes as C
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import torch, numpy as np
from concurrent.futures import ThreadPoolExecutor
import ako_b200, oracle_lib as ol
sys.setswitchinterval(0.0005)
L = ako_b200.load(); orc = ol.load_oracle()
w, h = 1632, 2464
N = 16
pool = torch.empty((N, h, w, 4), dtype=torch.uint8).pin_memory()
for i in range(4): pool[i].copy_(torch.from_numpy(ol.synth(orc, w, h, 2 + i)))
for i in range(4, N): pool[i].copy_(pool[i % 4])
s = ako_b200.default_settings(wavelet=0, quantization=16, gate=16)
cb = L.akoB200PinnedCallbacks()
free_fn = C.CFUNCTYPE(None, C.c_void_p)(cb.free)
def enc(i):
    out = C.c_void_p(); st = C.c_int(0)
    n = L.akoEncodeExt(C.byref(cb), C.byref(s), 4, w, h, pool[i % N].data_ptr(), C.byref(out), C.byref(st))
    assert n
    return out, n
def dec(out, n):
    st = C.c_int(0); a, b, c = C.c_size_t(), C.c_size_t(), C.c_size_t()
    p = L.akoDecodeExt(C.byref(cb), n, out, None, C.byref(a), C.byref(b), C.byref(c), C.byref(st))
    assert p
    return p
def both(i):
    out, n = enc(i); p = dec(out, n); free_fn(out); free_fn(p)
for i in range(4): both(i)
res = {}
def best(fn, items, T, reps=4):
    ex = ThreadPoolExecutor(T)
    out = None
    bt = 1e9
    for r in range(reps + 2):
        t = time.perf_counter(); out = list(ex.map(fn, items)); dt = time.perf_counter() - t
        if r >= 2: bt = min(bt, dt)
        if r < reps + 1 and fn is enc:
            for (o, n) in out: free_fn(o)
    ex.shutdown()
    return bt, out
for T in (1, 2, 3, 4, 6, 8, 12):
    dt, _ = best(both, range(64), T)
    res[f"both_T{T}"] = round(64 * w * h / dt / 1e6, 1)
for T in (1, 2, 4, 8):
    dt, bl = best(enc, range(64), T)
    res[f"enc_T{T}"] = round(64 * w * h / dt / 1e6, 1)
    def d(b):
        p = dec(*b); free_fn(p)
    dt, _ = best(d, bl, T)
    res[f"dec_T{T}"] = round(64 * w * h / dt / 1e6, 1)
    for (o, n) in bl: free_fn(o)
print(json.dumps(res))
