"""configs[4]: 16384x16384 RGBA8, CDF 5/3, lossless, one tile: device-resident encode / decode times and the
kernel split. Correctness of this case is tests/test_gpu_parity.py::test_c5_lossless_16384_known_answer_and_roundtrip."""
import sys, os, time, json
import numpy as np, torch
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import ako_b200, oracle_lib as ol
orc = ol.load_oracle()
w = h = 16384
img = ol.synth(orc, w, h, 5)
ctx = ako_b200.Context(0)
s = ako_b200.default_settings(wavelet=1, quantization=0, gate=0)
d_in = torch.from_numpy(img).cuda()
bound = ctx.encode_bound(s, 4, w, h)
d_out = torch.empty(bound, dtype=torch.uint8, device="cuda")
d_px = torch.empty_like(d_in)
ts = torch.cuda.ExternalStream(ctx.stream)
res = {}
for it in range(3):
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record(ts)
    size, st = ctx.encode_device(s, 4, w, h, d_in.data_ptr(), d_out.data_ptr(), bound)
    e1.record(ts)
    assert st == 0
    st, dims, _ = ctx.decode_device(size, d_out.data_ptr(), d_px.data_ptr(), w * h * 4)
    e2.record(ts)
    ctx.sync()
    assert st == 0
    res = {"blob_bytes": size, "encode_ms": e0.elapsed_time(e1), "decode_ms": e1.elapsed_time(e2)}
assert torch.equal(d_px, d_in)
res["encode_MPix_s"] = w * h / res["encode_ms"] / 1e3
res["decode_MPix_s"] = w * h / res["decode_ms"] / 1e3
ctx.profile_reset(); ctx.profile(True)
size, st = ctx.encode_device(s, 4, w, h, d_in.data_ptr(), d_out.data_ptr(), bound)
ctx.decode_device(size, d_out.data_ptr(), d_px.data_ptr(), w * h * 4)
ctx.sync()
prof = ctx.profile_get()
res["kernels_ms"] = {k: round(v[1], 3) for k, v in sorted(prof.items(), key=lambda kv: -kv[1][1])[:12]}
print(json.dumps(res))
