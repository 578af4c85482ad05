#!/bin/bash
# Programmatic dependent launch A/B: GPU tests with PDL on, then the bench with and without it on the same box.
OUT=gpurun_out/r02pdl
mkdir -p $OUT
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -8
for V in 1 0 1 0; do
  WCMC_TUNE=pdl=$V timeout 600 python bench.py --steps 100 --warmup 5 --kernel-pass-steps 3 --no-720p > $OUT/bench_pdl$V.json 2> $OUT/bench_pdl$V.err
  python - <<PY
import json
try:
    d = json.loads(open("$OUT/bench_pdl$V.json").read().strip().splitlines()[-1])
    print("pdl=$V  %8.1f patches/s  %.3f ms  e2e %.1f  clocks %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["clocks"]))
except Exception as e:
    print("pdl=$V no line", e); print(open("$OUT/bench_pdl$V.err").read()[-1500:])
PY
done
