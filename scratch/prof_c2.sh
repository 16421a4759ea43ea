#!/bin/bash
# ncu --set full captures of chosen kernels of the C2 bench: scratch/prof_c2.sh TAG "k_a k_b ..." [bench args]
TAG=$1; O=gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --skip-cpu --secondaries none ${3:-}"
$CMD > $O/${TAG}_plain.log 2>&1 || { tail -5 $O/${TAG}_plain.log; exit 1; }
for K in $2; do
  ncu --set full --clock-control none --import-source on -k regex:$K -s ${SKIP:-0} -c 1 -f -o $O/${TAG}_$K $CMD > /dev/null 2>&1
  python scratch/ncu_summary.py $O/${TAG}_$K.ncu-rep > $O/${TAG}_$K.summary.txt 2>/dev/null
  python scratch/ncu_lines.py $O/${TAG}_$K.ncu-rep > $O/${TAG}_$K.lines.txt 2>/dev/null
  python scratch/ncu_bank.py $O/${TAG}_$K.ncu-rep > $O/${TAG}_$K.bank.txt 2>/dev/null
  rm -f $O/${TAG}_$K.ncu-rep
  echo "=== $K"; grep -E "==|duration|dram__bytes|pipe_alu.avg|pipe_fma.avg|pipe_lsu.avg|warps_active|issue_active|registers|opcode|stall|total warp" $O/${TAG}_$K.summary.txt
  awk '{print $5+0, $0}' $O/${TAG}_$K.lines.txt | sort -n -r | head -${TOP:-25} | cut -d' ' -f2- | cut -c1-160
done
