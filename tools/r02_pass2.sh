#!/bin/bash
# Round 2, GPU pass 2: grouped weight-gradient kernel -- timings against the round-1 launches, ncu --set full.
O=gpurun_out/r02b
mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_parity.py -q -k "grouped_wgrad" > $O/pytest_wgrad.log 2>&1; echo "grouped wgrad pytest exit $?"; tail -3 $O/pytest_wgrad.log
timeout 300 python tools/wgrad_group_bench.py 7 both all > $O/wgrad_group_bench.txt 2>&1; cat $O/wgrad_group_bench.txt
timeout 600 ncu --set full --import-source on --clock-control none -k "regex:conv_wgrad_group|wgrad_reduce" -s 2 -c 4 -f -o $O/wgrad_group \
  python tools/wgrad_group_bench.py 1 both group > $O/ncu.log 2>&1; echo "ncu exit $?"; tail -2 $O/ncu.log
python tools/ncu_summary.py $O/wgrad_group.ncu-rep > $O/ncu_wgrad_group.txt 2>&1; head -120 $O/ncu_wgrad_group.txt
