#!/bin/bash
# quick GPU check: kagari + end-to-end parity tests, then the headline bench with C5 beside it
TAG=${1:-q}
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "${2:-kagari or end_to_end or golden or known}" 2>&1 | tail -3
python bench.py --steps 10 --warmup 3 --skip-cpu --secondaries ${3:-c5} > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err || tail -5 gpurun_out/${TAG}_bench.err
python - <<PY
import json
p=json.loads(open("gpurun_out/${TAG}_bench.json").read().strip().splitlines()[-1])
print(p["value"], p["ms_per_step"], "enc", p["secondary"]["encode_ms_per_step"], "dec", p["secondary"]["decode_ms_per_step"])
print({k:v["ms_per_step"] for k,v in list(p["kernels"].items())[:14]})
for k,v in p["secondary"].items():
    if isinstance(v, dict): print(k, {a:b for a,b in v.items() if a not in ("workload","how")})
PY
