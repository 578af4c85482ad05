#!/bin/bash
# usage: gpu_quick.sh <tag> "<pytest -k expr or empty for all gpu tests>" [bench]
O=gpurun_out/${1:-q}; mkdir -p $O
if [ -n "$2" ]; then K="-k"; E="$2"; else K=""; E=""; fi
timeout 900 python -m pytest tests -m gpu -x -q $K "$E" > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log
grep -v Warning $O/pytest.log | tail -25
if [ "$3" = "bench" ]; then
  timeout 600 python bench.py --no-cpu-baseline > $O/bench.json 2> $O/bench.err; echo "bench exit $?"; grep -v Warn $O/bench.err | tail -5
  python - <<PY
import json
d=json.load(open("$O/bench.json"))
print({k:d[k] for k in ("value","ms_per_step","gpu_launches")}, d["e2e"]["value"])
for k,v in sorted(d["kernels"].items(), key=lambda kv:-kv[1]["ms_per_step"]): print("  %-20s"%k, v)
print(d.get("denoise_720p"))
print(d["roofline"])
PY
fi
