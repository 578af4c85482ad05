#!/bin/bash
TAG=${1:-r01final}
O=gpurun_out/$TAG
mkdir -p $O
timeout 400 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log
tail -3 $O/pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke exit $?"; tail -3 $O/smoke.log
timeout 300 python bench.py > $O/bench.json 2> $O/bench.err; echo "bench exit $?"; tail -2 $O/bench.err
python - <<PY
import json
d = json.loads(open("$O/bench.json").read().strip().splitlines()[-1])
print({k: d.get(k) for k in ("value", "ms_per_step", "gpu_launches")}, d.get("e2e", {}).get("value"), d.get("roofline", {}).get("achieved"), (d.get("denoise_720p") or {}).get("ms_per_frame"), d.get("cpu_baseline"))
for k, v in sorted(d.get("kernels", {}).items(), key=lambda kv: -kv[1].get("ms_per_step", 0))[:14]:
    print("   %-22s %3d calls %7.3f ms  %s" % (k, v["calls_per_step"], v["ms_per_step"], {a: b for a, b in v.items() if a in ("tflops", "gbs")}))
PY
