"""Forward + inverse lifting of one 8192^2 x 4 image per wrap mode (strip + frame vs the general kernels)."""
import os, sys, numpy as np, torch
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R)
import ako_b200
from ako_b200.synth import synth_rgba8_torch
w = h = 4096
ctx = ako_b200.Context(0)
img = synth_rgba8_torch(w, h, [1, 2, 3, 4], device="cuda").contiguous()
out = torch.empty_like(img)
ts = torch.cuda.ExternalStream(ctx.stream)
for wavelet in (0, 1):
    for wrap in (0, 1, 2, 3):
        s = ako_b200.default_settings(wavelet=wavelet, wrap=wrap, quantization=16, gate=16)
        bound = ctx.encode_bound(s, 4, w, h); stride = -(-bound // 256) * 256
        blobs = torch.empty((4, stride), dtype=torch.uint8, device="cuda")
        def once():
            d, st, sz = ctx.encode_batch_device(s, 4, w, h, 4, img.data_ptr(), w * h * 4, blobs.data_ptr(), stride)
            assert d == 4, st
            d2, st2 = ctx.decode_batch_device(4, blobs.data_ptr(), stride, sz, out.data_ptr(), w * h * 4)
            assert d2 == 4
        for _ in range(2): once()
        ctx.sync(); best = 1e9
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(ts); once(); e1.record(ts); ctx.sync()
            best = min(best, e0.elapsed_time(e1))
        print(f"wavelet {wavelet} wrap {wrap} frame={'off' if os.environ.get('AKO_B200_NO_FRAME') else 'on'}: {best:.2f} ms enc+dec of 4 x 4096^2 RGBA", flush=True)
