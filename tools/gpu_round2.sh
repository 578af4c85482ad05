#!/bin/bash
# Second visit: new loss / full-frame tests, conv knob experiments, bench with 720p, kernel-apply capture.
TAG=${1:-r01b}
O=gpurun_out/$TAG
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_loss.py tests/test_gpu_parity.py -k "loss or fmse or grs or full_frame or golden" -x -q > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log
tail -15 $O/pytest.log
timeout 300 python tools/conv_bench.py knobs 5 > $O/conv_knobs.txt 2>&1; cat $O/conv_knobs.txt
timeout 300 python tools/conv_bench.py all 5 > $O/conv_all.txt 2>&1; cat $O/conv_all.txt
timeout 600 python bench.py --no-cpu-baseline > $O/bench.json 2> $O/bench.err; echo "bench exit $?"; tail -3 $O/bench.err
cat $O/bench.json
BENCH="python bench.py --no-graph --steps 1 --warmup 3 --no-cpu-baseline --no-720p"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:kernel_apply -s 4 -c 4 -f -o $O/prof_kernel_apply $BENCH > $O/ncu_ka.log 2>&1
ls -la $O
