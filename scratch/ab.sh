#!/bin/bash
# A/B of library variants on the C2 bench: scratch/ab.sh lib1.so lib2.so ...   ("" = the in-tree build)
for L in "" "$@"; do AKO_B200_LIB=$L python bench.py --steps 10 --warmup 3 --skip-cpu --secondaries ${SEC:-none} 2>/dev/null | python -c "
import json,sys
p=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=p['kernels']; print('[$L]', p['value'], p['ms_per_step'], {a:k[a]['ms_per_step'] for a in list(k)[:6]})
d=(p.get('secondary') or {}).get('dwt')
if d:
    for w,r in d['results'].items(): print('   ', w, {a:(v['ms'], v['level0']['ms']) for a,v in r.items()})"; done
