#!/bin/bash
# Round-2 closing pass on one B200: GPU tests, smoke(), the default bench line, the ncu launch list of the bench
# command, memcheck of the exchange kernel (one rank).
O=gpurun_out/r02z
mkdir -p $O
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -4
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py > $O/bench.json 2> $O/bench.err; python - <<PY
import json
d = json.loads(open("$O/bench.json").read().strip().splitlines()[-1])
print("bench: %.1f patches/s %.3f ms e2e %.1f roofline %s frac %.3f 720p %.2f ms clocks %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["kernel"][:30], d["roofline"]["frac"], d["denoise_720p"]["ms_per_frame"], d["clocks"]))
PY
BENCH="python bench.py --no-graph --steps 2 --warmup 3 --no-cpu-baseline --no-720p"
WCMC_BRANCH_STREAMS=0 timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 3000 --csv --log-file $O/launches.csv $BENCH > $O/ncu_launch.log 2>&1
echo "launch list exit $?"
gzip -f $O/launches.csv
python tools/launch_shares.py $O/launches.csv.gz > $O/launch_shares.txt; head -30 $O/launch_shares.txt
mkdir -p /tmp/xw
RANK=0 WORLD_SIZE=1 LOCAL_RANK=0 MASTER_ADDR=127.0.0.1 MASTER_PORT=29977 timeout 300 compute-sanitizer --error-exitcode 7 --print-limit 10 --tool memcheck python tests/exchange_worker.py /tmp/xw > $O/sanitizer_memcheck_exchange.log 2>&1; echo "memcheck exchange exit $?"; tail -4 $O/sanitizer_memcheck_exchange.log; cat /tmp/xw/rank0.json 2>/dev/null | head -c 600
