#!/bin/bash
# Standard evidence run on one B200 (through gpurun): tests, benches, launch list, one full capture.
TAG=${1:-r1}
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/${TAG}_tests.log
python bench.py --steps 20 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2>> gpurun_out/${TAG}_bench.err
python bench.py --workload dwt --steps 5 --warmup 2 > gpurun_out/${TAG}_dwt.json 2>> gpurun_out/${TAG}_bench.err
CMD="python bench.py --steps 2 --warmup 3 --skip-cpu"
$CMD > gpurun_out/${TAG}_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_lift_strip -s 0 -c 1 -f -o gpurun_out/${TAG}_c2_lift_strip_l0 $CMD > /dev/null 2>&1
cat gpurun_out/${TAG}_tests.log; cat gpurun_out/${TAG}_bench.json | head -c 600; echo; cat gpurun_out/${TAG}_bench_ref.json | head -c 400; echo; tail -2 gpurun_out/${TAG}_bench.err
