"""Stale-workspace hunt: dirty a context's workspaces with a large noisy encode, then encode a tiled image and
compare block by block with the oracle."""
import os, sys, struct
import numpy as np, torch
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import ako_b200, oracle_lib as ol
from ako_b200.synth import synth_rgba8_torch
orc = ol.load_oracle()
ctx = ako_b200.Context(0)
which = sys.argv[1] if len(sys.argv) > 1 else "noise"
W = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
if which == "noise":
    big = torch.randint(0, 256, (4096, 4096, 4), dtype=torch.uint8, device="cuda")
    s0 = ako_b200.default_settings(wavelet=1, quantization=0, gate=0)
    bound = ctx.encode_bound(s0, 4, 4096, 4096)
    out = torch.empty(bound, dtype=torch.uint8, device="cuda")
    n, st = ctx.encode_device(s0, 4, 4096, 4096, big.data_ptr(), out.data_ptr(), bound)
    print("dirty encode", n, st)
    px = torch.empty_like(big)
    print("dirty decode", ctx.decode_device(n, out.data_ptr(), px.data_ptr(), big.numel())[0], bool(torch.equal(px, big)))
if which == "c2":
    import bench
    w2, h2 = 1632, 2464
    pool = synth_rgba8_torch(w2, h2, list(range(2, 10)), device="cuda").repeat(8, 1, 1, 1).contiguous()
    s2 = ako_b200.default_settings(wavelet=0, quantization=16, gate=16)
    dc = bench.DeviceCodec(torch, ako_b200, ctx, 0, w2, h2, 4, s2, 64, pool)
    for i in range(3):
        dc.step(i)
    ctx.sync()
    del dc, pool
    torch.cuda.empty_cache()
    print("dirtied with the c2 batch")
w = h = W
img_t = synth_rgba8_torch(w, h, [40], device="cuda")[0].contiguous()
img = img_t.cpu().numpy()
for tiles in (256, 64):
    s = ako_b200.default_settings(wavelet=0, quantization=16, gate=0, tiles_dimension=tiles)
    bound = ctx.encode_bound(s, 4, w, h)
    out = torch.empty(bound, dtype=torch.uint8, device="cuda")
    n, st = ctx.encode_device(s, 4, w, h, img_t.data_ptr(), out.data_ptr(), bound)
    got = out[:n].cpu().numpy().tobytes()
    want, _ = ol.orc_encode(orc, img, wavelet=0, q=16, g=0, tiles=tiles)
    print("tiles", tiles, "equal", got == want, len(got), len(want), flush=True)
    n2, st = ctx.encode_device(s, 4, w, h, img_t.data_ptr(), out.data_ptr(), bound)
    got2 = out[:n2].cpu().numpy().tobytes()
    print("   second encode equal to oracle", got2 == want, "equal to first", got2 == got, flush=True)
    if got != want:
        # walk the blocks of both
        pa = pb = 16
        t = 0
        while pa < len(got) and pb < len(want):
            sa, sb = struct.unpack_from("<I", got, pa)[0], struct.unpack_from("<I", want, pb)[0]
            if sa != sb or got[pa + 4:pa + 4 + sa] != want[pb + 4:pb + 4 + sb]:
                print("  tile", t, "differs: sizes", sa, sb)
                if t > 40:
                    break
            pa += 4 + sa; pb += 4 + sb; t += 1
