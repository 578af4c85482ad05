#!/bin/bash
# conv_share A/B: every convolution launch plans for 1/n of the SMs (two streams run side by side).
OUT=gpurun_out/r02share
mkdir -p $OUT
for V in ${VARIANTS:-"conv_share=1" "conv_share=2"}; do
  WCMC_TUNE=$V timeout 600 python bench.py --steps 80 --warmup 5 --kernel-pass-steps 3 --no-720p > "$OUT/bench_$V.json" 2> "$OUT/bench_$V.err"
  python - <<PY
import json
try:
    d = json.loads(open("$OUT/bench_$V.json").read().strip().splitlines()[-1])
    print("$V  %8.1f patches/s  %.3f ms  e2e %.1f  clocks %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["clocks"]["sm_mhz"]))
except Exception as e:
    print("$V no line", e); print(open("$OUT/bench_$V.err").read()[-800:])
PY
done
