#!/bin/bash
# fused MLP backward: focused tests first (own process: a pipeline bug traps the context), then the regression + bench
TAG=${1:-r01g}
O=gpurun_out/$TAG
mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -s -k "fused_mlp_backward" > $O/pytest_mlp.log 2>&1; echo "mlp pytest exit $?"
grep -i "PathNet grads\|worst\|passed\|failed\|error\|wcmc:" $O/pytest_mlp.log | head -20
timeout 700 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log
tail -5 $O/pytest.log
timeout 400 python bench.py --no-cpu-baseline > $O/bench.json 2> $O/bench.err; echo "bench exit $?"; tail -2 $O/bench.err
WCMC_FUSED_MLP_BWD=0 timeout 400 python bench.py --no-cpu-baseline --no-720p > $O/bench_generic_mlp_bwd.json 2> $O/bench2.err; echo "bench exit $?"
python - <<PY
import json, glob
for f in sorted(glob.glob("$O/bench*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, {k: d.get(k) for k in ("value", "ms_per_step", "gpu_launches")}, d.get("e2e", {}).get("value"), d.get("roofline", {}).get("achieved"), (d.get("denoise_720p") or {}).get("ms_per_frame"))
        for k, v in sorted(d.get("kernels", {}).items(), key=lambda kv: -kv[1].get("ms_per_step", 0))[:16]:
            print("   %-22s %3d calls %7.3f ms  %s" % (k, v["calls_per_step"], v["ms_per_step"], {a: b for a, b in v.items() if a in ("tflops", "gbs")}))
    except Exception as e:
        print(f, "unreadable", e)
PY
