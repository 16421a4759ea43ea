#!/bin/bash
# Standard evidence run on one B200 (through gpurun): tests, benches of every workload, the reference arm.
TAG=${1:-r1}
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/${TAG}_tests.log
python bench.py --steps 20 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2>> gpurun_out/${TAG}_bench.err
python bench.py --workload dwt --steps 5 --warmup 2 > gpurun_out/${TAG}_dwt.json 2>> gpurun_out/${TAG}_bench.err
python bench.py --workload c1 --steps 10 --warmup 3 --skip-cpu > gpurun_out/${TAG}_c1.json 2>> gpurun_out/${TAG}_bench.err
python bench.py --workload c4 --steps 10 --warmup 3 --skip-cpu > gpurun_out/${TAG}_c4.json 2>> gpurun_out/${TAG}_bench.err
python scratch/c5_time.py > gpurun_out/${TAG}_c5.json 2>> gpurun_out/${TAG}_bench.err
cat gpurun_out/${TAG}_tests.log; head -c 700 gpurun_out/${TAG}_bench.json; echo; head -c 400 gpurun_out/${TAG}_bench_ref.json; echo; tail -2 gpurun_out/${TAG}_bench.err
