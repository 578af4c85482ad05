#!/bin/bash
# one `ncu --set full` capture of the kernels matching a regex inside one eager training step
#   bash tools/gpu_ncu_one.sh TAG REGEX [COUNT]
TAG=$1; REGEX=$2; CNT=${3:-2}
O=gpurun_out/$TAG
mkdir -p $O
WCMC_BRANCH_STREAMS=0 timeout 600 ncu --set full --import-source on --clock-control none -k "regex:$REGEX" -c $CNT -f -o $O/cap \
  python bench.py --no-graph --steps 1 --warmup 1 --no-cpu-baseline --no-720p > $O/ncu.log 2>&1
echo "ncu exit $?"; tail -3 $O/ncu.log
python tools/ncu_summary.py $O/cap.ncu-rep > $O/summary.txt 2>&1; cat $O/summary.txt | head -60
ls -la $O
