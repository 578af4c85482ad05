#!/bin/bash
# generic regression + bench pass:  bash tools/gpu_round6.sh TAG [extra bench args]
TAG=${1:-r01f}
O=gpurun_out/$TAG
mkdir -p $O
timeout 700 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log
tail -3 $O/pytest.log
timeout 300 python tools/conv_bench.py all 5 > $O/conv_layers.txt 2>&1; grep -v "kernel only\|tensor-core" $O/conv_layers.txt | cut -c1-120
timeout 400 python bench.py --no-cpu-baseline > $O/bench.json 2> $O/bench.err; echo "bench exit $?"; tail -2 $O/bench.err
python - <<PY
import json, glob
for f in sorted(glob.glob("$O/bench*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, {k: d.get(k) for k in ("value", "ms_per_step", "gpu_launches")}, d.get("e2e", {}).get("value"), d.get("roofline", {}).get("achieved"), (d.get("denoise_720p") or {}).get("ms_per_frame"))
        for k, v in sorted(d.get("kernels", {}).items(), key=lambda kv: -kv[1].get("ms_per_step", 0))[:14]:
            print("   %-22s %3d calls %7.3f ms  %s" % (k, v["calls_per_step"], v["ms_per_step"], {a: b for a, b in v.items() if a in ("tflops", "gbs")}))
    except Exception as e:
        print(f, "unreadable", e)
PY
