#!/bin/bash
# Round-2 evidence: ncu launch list + --set full captures of the dominant kernels of the C2 bench, summarised on the box.
TAG=${1:-r2}
O=gpurun_out
summarise() { # $1 = report path without extension
  python scratch/ncu_summary.py $1.ncu-rep > $1.summary.txt 2>/dev/null
  python scratch/ncu_lines.py $1.ncu-rep > $1.lines.txt 2>/dev/null
  python scratch/ncu_bank.py $1.ncu-rep > $1.bank.txt 2>/dev/null
}
CMD="python bench.py --steps 1 --warmup 3 --skip-cpu --secondaries none"
$CMD > $O/${TAG}_plain_c2.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 400 --csv --log-file $O/${TAG}_c2_launches.csv $CMD > /dev/null 2>&1
for K in k_lift_strip4 k_unlift_strip k_kg_lengths k_kd_decode k_kd_sync k_kt_fill; do
  ncu --set full --clock-control none --import-source on -k regex:$K -s 0 -c 1 -f -o $O/${TAG}_c2_$K $CMD > /dev/null 2>&1
  summarise $O/${TAG}_c2_$K
  rm -f $O/${TAG}_c2_$K.ncu-rep
done
# configs[4]: the decoder on a lossless 8192^2 image
CMD5="python scratch/c5_drv.py 8192"
$CMD5 > $O/${TAG}_plain_c5.log 2>&1 || exit 1
for K in k_kd_decode k_kd_sync k_kg_lengths; do
  ncu --set full --clock-control none --import-source on -k regex:$K -s 0 -c 1 -f -o $O/${TAG}_c5_$K $CMD5 > /dev/null 2>&1
  summarise $O/${TAG}_c5_$K
  rm -f $O/${TAG}_c5_$K.ncu-rep
done
ls -la $O/${TAG}_* | awk '{print $5, $9}'
