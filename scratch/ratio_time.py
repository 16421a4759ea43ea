"""Ratio search (akoB200EncodeRatio) against the tool's loop over the reference's akoEncodeExt, configs[1] shape."""
import sys, os, time, json
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import ako_b200, oracle_lib as ol
orc = ol.load_oracle(); ref = ol.load_ref()
img = ol.synth(orc, 1632, 2464, 2)
s = ako_b200.default_settings(wavelet=0, gate=16)
for ratio in (10, 25, 60):
    ako_b200.encode_ratio(img, ratio, s)
    t = time.perf_counter(); blob, st, q, passes = ako_b200.encode_ratio(img, ratio, s); gpu = time.perf_counter() - t
    t = time.perf_counter(); want, wq, wp = ol.ref_encode_pass(ref, img, ratio, wavelet=0, g=16); cpu = time.perf_counter() - t
    # the same passes as plain akoEncodeExt calls on the GPU library
    t = time.perf_counter()
    for _ in range(passes):
        ako_b200.encode(img, s)
    naive = time.perf_counter() - t
    print(json.dumps({"ratio": ratio, "q": q, "passes": passes, "same_blob": blob == want, "size": len(blob),
                      "gpu_search_ms": round(gpu * 1e3, 2), "gpu_naive_passes_ms": round(naive * 1e3, 2),
                      "reference_cpu_ms": round(cpu * 1e3, 1)}))
