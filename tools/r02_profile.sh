#!/bin/bash
# Round 2 ncu evidence for profiles/: launch list of the bench command (2 steps, with DRAM bytes), ncu --set full of the
# dominant kernels, sanitizer on the kernels added in round 2, repeated 720p timings.
O=gpurun_out/r02p
mkdir -p $O
BENCH="python bench.py --no-graph --steps 2 --warmup 3 --no-cpu-baseline --no-720p"
WCMC_BRANCH_STREAMS=0 timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 3000 --csv --log-file $O/launches.csv $BENCH > $O/ncu_launch.log 2>&1
echo "launch list exit $?"
gzip -f $O/launches.csv
python tools/launch_shares.py $O/launches.csv.gz > $O/launch_shares.txt; head -45 $O/launch_shares.txt
cap() {  # name, kernel regex, extra ncu args..., -- command
  local name=$1 regex=$2; shift 2
  local extra=()
  while [ "$1" != "--" ]; do extra+=("$1"); shift; done; shift
  timeout 600 ncu --set full --import-source on --clock-control none -k "regex:$regex" "${extra[@]}" -f -o $O/$name "$@" > $O/$name.log 2>&1
  python tools/ncu_summary.py $O/$name.ncu-rep > $O/$name.txt 2>&1
  local sz=$(stat -c %s $O/$name.ncu-rep 2>/dev/null || echo 0)
  if [ "$sz" -gt 9000000 ]; then rm -f $O/$name.ncu-rep; fi
  tail -1 $O/$name.log
}
cap wgrad_group "conv_wgrad_group|wgrad_reduce" -s 2 -c 4 -- python tools/wgrad_group_bench.py 1 both group
cap conv_step "conv_igemm" -s 40 -c 40 -- env WCMC_BRANCH_STREAMS=0 $BENCH
cap frame720 "conv_igemm_kernel<.*1>|conv_igemm_kernel<1, 5, 1>|recombine" -c 24 -- python tools/frame_bench.py 1
cap mlp "pathnet" -s 8 -c 8 -- env WCMC_BRANCH_STREAMS=0 $BENCH
SAN="compute-sanitizer --error-exitcode 7 --print-limit 20"
timeout 900 $SAN --tool memcheck python -m pytest tests/test_gpu_parity.py -x -q -k "(grouped_wgrad_vs_fp64 and float16) or (fused_last_conv and 2-44-36) or step_glue or device_permutation or batched_pack" > $O/sanitizer_memcheck_r02.log 2>&1; echo "memcheck exit $?"; tail -3 $O/sanitizer_memcheck_r02.log
timeout 600 $SAN --tool racecheck python -m pytest tests/test_gpu_parity.py -x -q -k "(grouped_wgrad_vs_fp64 and float16-1) or (fused_last_conv and 2097152-2-44-36) or step_glue" > $O/sanitizer_racecheck_r02.log 2>&1; echo "racecheck exit $?"; tail -3 $O/sanitizer_racecheck_r02.log
for i in 1 2 3; do python tools/frame_bench.py 10; done > $O/frame_bench.txt 2>&1; cat $O/frame_bench.txt
du -sh $O
