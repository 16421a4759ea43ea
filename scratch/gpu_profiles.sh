#!/bin/bash
# ncu captures of the dominant kernels (one B200). TAG names the round/capture.
TAG=${1:-r1e}
CMD="python bench.py --steps 1 --warmup 3 --skip-cpu"
$CMD > gpurun_out/${TAG}_plain_c2.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/${TAG}_c2_launches.csv $CMD > /dev/null 2>&1
for K in k_lift_strip k_unlift_strip k_kg_lengths k_kt_expand k_kg_starts; do
  ncu --set full --clock-control none --import-source on -k regex:$K -s 0 -c 1 -f -o gpurun_out/${TAG}_c2_$K $CMD > /dev/null 2>&1
done
for WL in cdf53 dd137; do
  CMD2="python bench.py --workload dwt --dwt-wavelets $WL --steps 1 --warmup 0"
  $CMD2 > gpurun_out/${TAG}_plain_dwt_$WL.log 2>&1 || exit 1
  ncu --set full --clock-control none --import-source on -k regex:k_lift_strip -s 0 -c 1 -f -o gpurun_out/${TAG}_dwt_${WL}_k_lift_strip $CMD2 > /dev/null 2>&1
  ncu --set full --clock-control none --import-source on -k regex:k_unlift_strip -s 5 -c 1 -f -o gpurun_out/${TAG}_dwt_${WL}_k_unlift_strip $CMD2 > /dev/null 2>&1
done
ls gpurun_out/${TAG}_*
