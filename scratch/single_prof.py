"""One C2 image alone: per-kernel times (CUDA events around every launch) next to the end-to-end device time."""
import os, sys, json, time
import numpy as np, torch
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import ako_b200
from ako_b200.synth import synth_rgba8_torch
w, h = 1632, 2464
ctx = ako_b200.Context(0)
img = synth_rgba8_torch(w, h, [5], device="cuda")[0].contiguous()
s = ako_b200.default_settings(wavelet=0, quantization=16, gate=16)
bound = ctx.encode_bound(s, 4, w, h)
blob = torch.empty(bound, dtype=torch.uint8, device="cuda")
out = torch.empty_like(img)
ts = torch.cuda.ExternalStream(ctx.stream)
def once():
    done, st, sz = ctx.encode_batch_device(s, 4, w, h, 1, img.data_ptr(), img.numel(), blob.data_ptr(), bound)
    assert done == 1, st
    done2, st2 = ctx.decode_batch_device(1, blob.data_ptr(), bound, sz, out.data_ptr(), img.numel())
    assert done2 == 1
    return sz
for _ in range(3): once()
ctx.sync()
ts_ms = []
for _ in range(10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    e0.record(ts); once(); e1.record(ts); ctx.sync()
    ts_ms.append((e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3))
print("device ms / wall ms (median):", np.median([a for a, _ in ts_ms]), np.median([b for _, b in ts_ms]))
n0 = ctx.launch_count()
ctx.profile_reset(); ctx.profile(True)
once(); ctx.sync()
prof = ctx.profile_get(); ctx.profile(False)
print("launches per encode+decode:", sum(v[0] for v in prof.values()), "kernel time sum ms:", round(sum(v[1] for v in prof.values()), 4))
for k, v in sorted(prof.items(), key=lambda kv: -kv[1][1]):
    print(f"  {k:28s} x{v[0]:2d} {v[1]*1e3:8.1f} us")
