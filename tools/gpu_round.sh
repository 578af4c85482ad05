#!/bin/bash
# One GPU-box visit: parity tests, bench line, ncu launch list and full captures of the top kernels.
# Usage (here): gpurun --timeout 1500 -- 'bash tools/gpu_round.sh <tag>'
TAG=${1:-r01}
O=gpurun_out/$TAG
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > $O/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log
tail -3 $O/pytest.log
timeout 600 python bench.py > $O/bench.json 2> $O/bench.err; echo "bench exit $?"
cat $O/bench.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_ref.json 2>> $O/bench.err
BENCH="python bench.py --no-graph --steps 1 --warmup 3 --no-cpu-baseline"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $O/launches.csv $BENCH > $O/ncu_launch.log 2>&1
for K in conv_igemm conv_wgrad kernel_apply; do
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$K -s 60 -c 3 -f -o $O/prof_$K $BENCH > $O/ncu_$K.log 2>&1
done
ls -la $O
