#!/bin/bash
# pair + row-stage conv as the default: regression, pair_check, knobs, bench
TAG=${1:-r01e}
O=gpurun_out/$TAG
mkdir -p $O
timeout 300 python tools/pair_check.py 5 > $O/pair_check.txt 2>&1; PC=$?; echo "pair_check exit $PC"; tail -24 $O/pair_check.txt | cut -c1-200
timeout 700 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log
tail -4 $O/pytest.log
timeout 200 python tools/conv_bench.py knobs 5 > $O/conv_knobs.txt 2>&1; cat $O/conv_knobs.txt | cut -c1-120
timeout 400 python bench.py --no-cpu-baseline > $O/bench.json 2> $O/bench.err; echo "bench exit $?"; tail -2 $O/bench.err
WCMC_TUNE=conv_row_stages=0 timeout 400 python bench.py --no-cpu-baseline --no-720p > $O/bench_tapstages.json 2> $O/bench_tapstages.err; echo "bench exit $?"
python - <<PY
import json, glob
for f in sorted(glob.glob("$O/bench*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, {k: d.get(k) for k in ("value", "ms_per_step", "gpu_launches")}, d.get("e2e", {}).get("value"), d.get("roofline", {}).get("achieved"), d.get("denoise_720p"))
    except Exception as e:
        print(f, "unreadable", e)
PY
