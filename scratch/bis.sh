for sec in shapes c1,shapes c4,shapes c5,shapes; do
  python bench.py --steps 3 --warmup 3 --skip-cpu --secondaries $sec > gpurun_out/bis.json 2>gpurun_out/bis.err
  python -c "
import json
p=json.loads(open('gpurun_out/bis.json').read().strip().splitlines()[-1])
print('$sec', p['secondary']['shapes']['rgba_8192x8192_tiles256'])"
done
