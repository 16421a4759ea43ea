import sys, os, time
R = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
sys.setswitchinterval(0.0005)
import torch, numpy as np
import bench, ako_b200, oracle_lib as ol
orc = ol.load_oracle()
w, h, B = 1632, 2464, 32
ctx = ako_b200.Context(0)
s = ako_b200.default_settings(wavelet=0, quantization=16, gate=16)
img = torch.from_numpy(ol.synth(orc, w, h, 2)).cuda()
pool = img.unsqueeze(0).repeat(B, 1, 1, 1).contiguous()
bound = ctx.encode_bound(s, 4, w, h); stride = -(-bound // 256) * 256
blobs = torch.empty((B, stride), dtype=torch.uint8, device="cuda"); out = torch.empty_like(pool)
def step():
    done, st, sizes = ctx.encode_batch_device(s, 4, w, h, B, pool.data_ptr(), w * h * 4, blobs.data_ptr(), stride)
    ctx.decode_batch_device(B, blobs.data_ptr(), stride, sizes, out.data_ptr(), w * h * 4)
for _ in range(3): step()
torch.cuda.synchronize()
sm = bench.ClockSampler(0)
t0 = time.perf_counter()
sm.start()
for _ in range(20): step()
torch.cuda.synchronize()
t1 = time.perf_counter()
r = sm.stop()
print(r, "region ms", (t1 - t0) * 1e3, [round((x - t0) * 1e3, 1) for x in sm.stamps][:30])
