"""Measures pinned H2D / D2H / concurrent bandwidth on the box (planning number for the e2e leg)."""
import torch, time, json
n = 128 << 20
h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_a = torch.empty(n, dtype=torch.uint8, device="cuda")
d_b = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
res = {}
def run(label, h2d, d2h, reps=10):
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1):
                d_a.copy_(h_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2):
                h_out.copy_(d_b, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t
    res[label] = round(n * reps / dt / 1e9, 2)
for _ in range(2):
    run("h2d_GBps", True, False)
    run("d2h_GBps", False, True)
    run("both_each_GBps", True, True)
# small chunks: 16 MB
m = 16 << 20
torch.cuda.synchronize(); t = time.perf_counter()
for i in range(64):
    with torch.cuda.stream(s1):
        d_a[:m].copy_(h_in[:m], non_blocking=True)
torch.cuda.synchronize(); res["h2d_16MB_chunks_GBps"] = round(m * 64 / (time.perf_counter() - t) / 1e9, 2)
print(json.dumps(res))
