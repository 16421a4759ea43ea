"""Full-pyramid DWT time per image size: t(8192) - t(4096) is what level 0 of the 8192 pyramid costs, and so on."""
import sys, os, json, ctypes as C
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch, numpy as np
import ako_b200
ctx = ako_b200.Context(0); L = ako_b200.load()
ts = torch.cuda.ExternalStream(ctx.stream)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
res = {}
for wavelet, name in ((1, "cdf53"), (0, "dd137"), (2, "haar")):
    s = ako_b200.default_settings(wavelet=wavelet, quantization=0, gate=0)
    for size in (8192, 4096, 2048, 1024, 512, 256):
        w = h = size
        n = ctx.stream_size(4, w, h) // 2
        base = torch.randint(-255, 256, (4, h, w), dtype=torch.int16, device="cuda")
        planes = torch.empty_like(base); st_t = torch.empty(n + 64, dtype=torch.int16, device="cuda")
        for direction in ("fwd", "inv"):
            times = []
            for it in range(6):
                if direction == "fwd": planes.copy_(base)
                if size >= 4096: flush.zero_()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(ts)
                if direction == "fwd": L.akoB200Lift(ctx.h, C.byref(s), 4, w, h, planes.data_ptr(), st_t.data_ptr())
                else: L.akoB200Unlift(ctx.h, C.byref(s), 4, w, h, st_t.data_ptr(), planes.data_ptr())
                e1.record(ts); ctx.sync()
                if it >= 2: times.append(e0.elapsed_time(e1))
            res[f"{name}_{direction}_{size}"] = round(float(np.median(times)) * 1e3, 1)
        if size == 8192:
            for direction in ("fwd", "inv"):
                planes.copy_(base); ctx.profile_reset(); ctx.profile(True)
                if direction == "fwd": L.akoB200Lift(ctx.h, C.byref(s), 4, w, h, planes.data_ptr(), st_t.data_ptr())
                else: L.akoB200Unlift(ctx.h, C.byref(s), 4, w, h, st_t.data_ptr(), planes.data_ptr())
                ctx.sync(); ctx.profile(False)
                res[f"{name}_{direction}_kernels"] = {k: (v[0], round(v[1] * 1e3, 1)) for k, v in ctx.profile_get().items()}
        del base, planes, st_t
print(json.dumps(res))
