#!/bin/bash
OUT=gpurun_out/r02q
mkdir -p $OUT
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "graphed" 2>&1 | tail -40
timeout 600 python tools/flake_hunt.py 300 100 2>&1 | grep -v Warning | tail -30
for V in 1 1; do
  WCMC_CONV_SHARE=$V timeout 600 python bench.py --steps 80 --warmup 5 --kernel-pass-steps 3 --no-720p > $OUT/bench2_share$V.json 2> $OUT/bench2_share$V.err
  python - <<PY
import json
try:
    d = json.loads(open("$OUT/bench2_share$V.json").read().strip().splitlines()[-1])
    print("share=$V  %8.1f patches/s  %.3f ms  e2e %.1f (%.3f ms)  clocks %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["clocks"]["sm_mhz"]))
except Exception as e:
    print("share=$V no line", e); print(open("$OUT/bench2_share$V.err").read()[-600:])
PY
done
