#!/bin/bash
# ncu captures of the dominant kernels (one B200). TAG names the round/capture. Each capture runs only after the same
# command has exited 0 without ncu. The reports are summarised ON THE BOX (scratch/ncu_summary.py, ncu_lines.py,
# ncu_bank.py -> gpurun_out/<name>.{summary,lines,bank}.txt); only the dominant kernel's .ncu-rep is kept, because
# gpurun brings back at most 64 MiB.
TAG=${1:-r1h}
O=gpurun_out
summarise() { # $1 = report path without extension
  python scratch/ncu_summary.py $1.ncu-rep > $1.summary.txt 2>/dev/null
  python scratch/ncu_lines.py $1.ncu-rep > $1.lines.txt 2>/dev/null
  python scratch/ncu_bank.py $1.ncu-rep > $1.bank.txt 2>/dev/null
}
CMD="python bench.py --steps 1 --warmup 3 --skip-cpu"
$CMD > $O/${TAG}_plain_c2.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_c2_launches.csv $CMD > /dev/null 2>&1
for K in k_lift_strip k_unlift_strip k_kg_lengths k_kt_expand k_kg_starts; do
  ncu --set full --clock-control none --import-source on -k regex:$K -s 0 -c 1 -f -o $O/${TAG}_c2_$K $CMD > /dev/null 2>&1
  summarise $O/${TAG}_c2_$K
  [ $K = k_lift_strip ] || rm -f $O/${TAG}_c2_$K.ncu-rep
done
for WL in cdf53 dd137; do
  CMD2="python bench.py --workload dwt --dwt-wavelets $WL --steps 1 --warmup 0"
  $CMD2 > $O/${TAG}_plain_dwt_$WL.log 2>&1 || exit 1
  ncu --set full --clock-control none --import-source on -k regex:k_lift_strip -s 0 -c 1 -f -o $O/${TAG}_dwt_${WL}_k_lift_strip $CMD2 > /dev/null 2>&1
  ncu --set full --clock-control none --import-source on -k regex:k_unlift_strip -s 5 -c 1 -f -o $O/${TAG}_dwt_${WL}_k_unlift_strip $CMD2 > /dev/null 2>&1
  for K in k_lift_strip k_unlift_strip; do
    summarise $O/${TAG}_dwt_${WL}_$K
    rm -f $O/${TAG}_dwt_${WL}_$K.ncu-rep
  done
done
ls -la $O/${TAG}_* | awk '{print $5, $9}'
