"""Kernel split of one encode+decode step of a named shape of bench.SHAPES: python scratch/shape_prof2.py NAME"""
import os, sys, json
import numpy as np, torch
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import ako_b200, bench
from ako_b200.synth import synth_rgba8_torch
name = sys.argv[1]
w, h, ch, wavelet, q, g, tiles, B = bench.SHAPES[name][:8]
wrap = bench.SHAPES[name][8] if len(bench.SHAPES[name]) > 8 else 0
ctx = ako_b200.Context(0)
imgs = synth_rgba8_torch(w, h, list(range(B)), device="cuda")[..., :ch].contiguous()
s = ako_b200.default_settings(wavelet=wavelet, quantization=q, gate=g, tiles_dimension=tiles, wrap=wrap)
bound = ctx.encode_bound(s, ch, w, h)
stride = (bound + 255) & ~255
blobs = torch.empty(B * stride, dtype=torch.uint8, device="cuda")
out = torch.empty_like(imgs)
torch.cuda.synchronize()
def once():
    done, st, sz = ctx.encode_batch_device(s, ch, w, h, B, imgs.data_ptr(), w * h * ch, blobs.data_ptr(), stride)
    assert done == B, st
    done2, st2 = ctx.decode_batch_device(B, blobs.data_ptr(), stride, sz, out.data_ptr(), w * h * ch)
    assert done2 == B
for _ in range(2): once()
_d, _st, _sz = ctx.encode_batch_device(s, ch, w, h, B, imgs.data_ptr(), w * h * ch, blobs.data_ptr(), stride)
print("compressed bytes per pixel", round(float(np.sum(_sz)) / (w * h * B), 4), "largest blob", int(np.max(_sz)))
ctx.sync()
ctx.profile_reset(); ctx.profile(True); once(); ctx.sync()
prof = ctx.profile_get(); ctx.profile(False)
tot = sum(v[1] for v in prof.values())
print(name, "launches", sum(v[0] for v in prof.values()), "kernel ms", round(tot, 3), "ns/px", round(tot * 1e6 / (w * h * B), 4))
for k, v in sorted(prof.items(), key=lambda kv: -kv[1][1])[:16]:
    print(f"  {k:28s} x{v[0]:3d} {v[1]:8.3f} ms")
