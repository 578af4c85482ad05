#!/bin/bash
# ncu evidence for profiles/: launch list of one step (with DRAM bytes) + full captures of the main kernels.
# Reports are summarised ON the box (tools/ncu_summary.py) and only small .ncu-rep files are kept: gpurun
# copies back at most 64 MiB.
TAG=${1:-prof}
O=gpurun_out/$TAG
mkdir -p $O
BENCH="python bench.py --no-graph --steps 1 --warmup 3 --no-cpu-baseline --no-720p"
cap() {  # name, kernel regex, extra ncu args..., -- command
  local name=$1 regex=$2; shift 2
  local extra=()
  while [ "$1" != "--" ]; do extra+=("$1"); shift; done; shift
  timeout 600 ncu --set full --clock-control none -k "regex:$regex" "${extra[@]}" -f -o $O/$name "$@" > $O/$name.log 2>&1
  python tools/ncu_summary.py $O/$name.ncu-rep > $O/$name.txt 2>&1
  local sz=$(stat -c %s $O/$name.ncu-rep 2>/dev/null || echo 0)
  if [ "$sz" -gt 6000000 ]; then rm -f $O/$name.ncu-rep; fi
  tail -2 $O/$name.log
}
WCMC_BRANCH_STREAMS=0 timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 3200 --csv --log-file $O/launches.csv $BENCH > $O/ncu_launch.log 2>&1
echo "launch list exit $?"
gzip -f $O/launches.csv
cap conv_layers conv_igemm -c 24 -- python tools/conv_bench.py fwd 1
cap wgrad_layers conv_wgrad_kernel -c 8 -- python tools/conv_bench.py wgrad 1
cap kernel_apply kernel_apply -c 12 -- python tools/ka_bench.py 1
cap misc "pathnet|fmse|adam_clip|wgrad_reduce_batch|slab_reduce" -s 14 -c 14 -- env WCMC_BRANCH_STREAMS=0 $BENCH
cap allpairs allpairs_kernel -c 2 -- python tools/loss_sweep.py 1
du -sh $O; ls -la $O
