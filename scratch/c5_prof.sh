#!/bin/bash
# ncu --set full captures of the decoder kernels on a lossless image (tag $1, size $2)
TAG=$1; SIZE=${2:-8192}; O=gpurun_out
CMD="python scratch/c5_drv.py $SIZE"
$CMD > $O/${TAG}_plain.log 2>&1 || { tail -5 $O/${TAG}_plain.log; exit 1; }
for K in ${3:-k_kd_sync k_kd_extract k_kt_spans k_kt_expand}; do
  ncu --set full --clock-control none --import-source on -k regex:$K -s 0 -c 1 -f -o $O/${TAG}_$K $CMD > /dev/null 2>&1
  python scratch/ncu_summary.py $O/${TAG}_$K.ncu-rep > $O/${TAG}_$K.summary.txt 2>/dev/null
  python scratch/ncu_lines.py $O/${TAG}_$K.ncu-rep > $O/${TAG}_$K.lines.txt 2>/dev/null
  rm -f $O/${TAG}_$K.ncu-rep
  echo "=== $K"; head -30 $O/${TAG}_$K.summary.txt | grep -E "==|duration|dram__bytes|pipe_alu.avg|pipe_lsu.avg|warps_active|issue_active|registers|opcode|stall|total warp"; awk '$3+0 >= 2.0 || $5+0 >= 3.0' $O/${TAG}_$K.lines.txt | head -40
done
