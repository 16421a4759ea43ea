"""Lossless single-image encode+decode driver for profiling: python scratch/c5_drv.py SIZE"""
import os, sys
import torch
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import ako_b200
from ako_b200.synth import synth_rgba8_torch
size = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
ctx = ako_b200.Context(0)
img = torch.empty((size, size, 4), dtype=torch.uint8, device="cuda")
for y0 in range(0, size, 1024):
    img[y0:y0 + 1024] = synth_rgba8_torch(size, min(1024, size - y0), [5], device="cuda", y0=y0)[0]
s = ako_b200.default_settings(wavelet=1, quantization=0, gate=0)
bound = ctx.encode_bound(s, 4, size, size)
blob = torch.empty(bound, dtype=torch.uint8, device="cuda")
out = torch.empty_like(img)
torch.cuda.synchronize()
for it in range(2):
    n, st = ctx.encode_device(s, 4, size, size, img.data_ptr(), blob.data_ptr(), bound)
    assert st == 0 and n
    st, _, _ = ctx.decode_device(n, blob.data_ptr(), out.data_ptr(), size * size * 4)
    assert st == 0
ctx.sync()
assert torch.equal(out, img)
print("ok", n)
