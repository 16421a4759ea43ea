"""Kernel split (CUDA events per launch) of encode+decode for one shape: python scratch/shape_prof.py w h ch wavelet q g tiles B"""
import os, sys, json
import numpy as np, torch
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import ako_b200, bench
from ako_b200.synth import synth_rgba8_torch
w, h, ch, wavelet, q, g, tiles, B = (int(x) for x in sys.argv[1:9])
ctx = ako_b200.Context(0)
stream = torch.cuda.ExternalStream(ctx.stream)
distinct = min(B, 4)
base = synth_rgba8_torch(w, h, [40 + i for i in range(distinct)], device="cuda")[..., :ch].contiguous()
pool = base.repeat((B + distinct - 1) // distinct, 1, 1, 1)[:B].contiguous()
s = ako_b200.default_settings(wavelet=wavelet, quantization=q, gate=g, tiles_dimension=tiles)
dc = bench.DeviceCodec(torch, ako_b200, ctx, 0, w, h, ch, s, B, pool)
for i in range(3):
    dc.step(i)
ms = bench._events_ms(torch, stream, dc.step, 5)
ctx.profile_reset(); ctx.profile(True)
dc.step(0); ctx.sync()
prof = ctx.profile_get(); ctx.profile(False)
print(f"{w}x{h}x{ch} wl{wavelet} q{q} g{g} tiles{tiles} B{B}: {ms:.3f} ms/step = {w*h*B/ms/1e3:.0f} MPix/s, {ms*1e6/(w*h*B):.4f} ns/px")
tot = sum(v[1] for v in prof.values())
for k, v in sorted(prof.items(), key=lambda kv: -kv[1][1])[:16]:
    print(f"   {k:26s} x{v[0]:<4d} {v[1]:.4f} ms  {100*v[1]/tot:.1f}%")
