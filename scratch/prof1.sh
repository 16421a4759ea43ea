#!/bin/bash
# one ncu --set full capture of kernel regex $2 from the C2 bench (tag $1), summarised on the box
TAG=$1; K=$2; SKIP=${3:-0}
O=gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --skip-cpu --no-secondary ${4:-}"
$CMD > $O/${TAG}_plain.log 2>&1 || { tail -5 $O/${TAG}_plain.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:$K -s $SKIP -c 1 -f -o $O/${TAG}_$K $CMD > /dev/null 2>&1
python scratch/ncu_summary.py $O/${TAG}_$K.ncu-rep > $O/${TAG}_$K.summary.txt 2>/dev/null
python scratch/ncu_lines.py $O/${TAG}_$K.ncu-rep > $O/${TAG}_$K.lines.txt 2>/dev/null
python scratch/ncu_bank.py $O/${TAG}_$K.ncu-rep > $O/${TAG}_$K.bank.txt 2>/dev/null
rm -f $O/${TAG}_$K.ncu-rep
cat $O/${TAG}_$K.summary.txt; cat $O/${TAG}_$K.lines.txt
