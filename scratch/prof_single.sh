#!/bin/bash
# ncu --set full of chosen kernels for ONE C2 image (scratch/single_prof.py): scratch/prof_single.sh TAG "k_a k_b"
TAG=$1; O=gpurun_out
CMD="python scratch/single_prof.py"
for K in $2; do
  ncu --set full --clock-control none --import-source on -k regex:$K -s ${SKIP:-3} -c 1 -f -o $O/${TAG}_$K $CMD > /dev/null 2>&1
  python scratch/ncu_summary.py $O/${TAG}_$K.ncu-rep > $O/${TAG}_$K.summary.txt 2>/dev/null
  python scratch/ncu_lines.py $O/${TAG}_$K.ncu-rep > $O/${TAG}_$K.lines.txt 2>/dev/null
  rm -f $O/${TAG}_$K.ncu-rep
  echo "=== $K"; grep -E "==|duration|dram__bytes|pipe_alu.avg|warps_active|issue_active|registers|grid_size|opcode|stall|total warp" $O/${TAG}_$K.summary.txt
  awk '{print $7+0, $0}' $O/${TAG}_$K.lines.txt | sort -n -r | head -${TOP:-22} | cut -d' ' -f2- | cut -c1-170
done
