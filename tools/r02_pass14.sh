#!/bin/bash
OUT=gpurun_out/r02q
mkdir -p $OUT
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
for V in a b c; do
  timeout 600 python bench.py --steps 100 --warmup 5 --kernel-pass-steps 3 --no-720p > $OUT/bench3_$V.json 2> $OUT/bench3_$V.err
  python - <<PY
import json
try:
    d = json.loads(open("$OUT/bench3_$V.json").read().strip().splitlines()[-1])
    print("$V  %8.1f patches/s  %.3f ms  e2e %.1f (%.3f ms)  clocks %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["clocks"]["sm_mhz"]))
except Exception as e:
    print("$V no line", e); print(open("$OUT/bench3_$V.err").read()[-600:])
PY
done
timeout 900 python bench.py > $OUT/bench_default.json 2> $OUT/bench_default.err; tail -c 1500 $OUT/bench_default.json
