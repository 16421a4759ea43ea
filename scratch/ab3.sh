#!/bin/bash
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "kagari or golden or known or end_to_end_shapes or sweep or ratio" 2>&1 | tail -2
for W in c2 c4 c1; do python bench.py --workload $W --steps 10 --warmup 3 --skip-cpu --secondaries none 2>/dev/null | python -c "
import json,sys
p=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=p['kernels']; print('$W', p['value'], p['ms_per_step'], 'lengths', k['kagari_lengths']['ms_per_step'])"; done
python scratch/c5_time.py 2>/dev/null | python -c "
import json,sys
p=json.loads(sys.stdin.read()); print('c5', p['encode_ms'], p['decode_ms'], p['kernels_ms'].get('kagari_lengths'))"
