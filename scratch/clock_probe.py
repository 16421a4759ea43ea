import sys, os, time, threading
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import pynvml
pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)
import torch
x = torch.zeros(1, device="cuda")
stamps = []
stop = False
def poll():
    while not stop:
        t0 = time.perf_counter()
        pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
        t1 = time.perf_counter()
        pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
        t2 = time.perf_counter()
        stamps.append((round((t0 - T0) * 1e3, 2), round((t1 - t0) * 1e3, 2), round((t2 - t1) * 1e3, 2)))
        time.sleep(0.002)
T0 = time.perf_counter()
th = threading.Thread(target=poll, daemon=True); th.start()
time.sleep(0.1)
print("idle:", stamps[:6]); n0 = len(stamps)
a = torch.randn(8192, 8192, device="cuda")
t = time.perf_counter()
for _ in range(200): b = a @ a
torch.cuda.synchronize()
print("busy ms", (time.perf_counter() - t) * 1e3, "samples", len(stamps) - n0, stamps[n0:n0 + 6])
stop = True
