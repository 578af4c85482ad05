#!/bin/bash
OUT=gpurun_out/r02q
mkdir -p $OUT
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
for V in 1 0 1 0; do
  WCMC_CONV_SHARE=$V timeout 600 python bench.py --steps 80 --warmup 5 --kernel-pass-steps 3 --no-720p > $OUT/bench_share$V.json 2> $OUT/bench_share$V.err
  python - <<PY
import json
try:
    d = json.loads(open("$OUT/bench_share$V.json").read().strip().splitlines()[-1])
    print("share=$V  %8.1f patches/s  %.3f ms  e2e %.1f  clocks %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["clocks"]["sm_mhz"]))
except Exception as e:
    print("share=$V no line", e); print(open("$OUT/bench_share$V.err").read()[-600:])
PY
done
timeout 300 python tools/step_timeline.py 3 > $OUT/timeline.txt 2>&1; tail -90 $OUT/timeline.txt
cp gpurun_out/timeline_last_replay.txt $OUT/ 2>/dev/null
