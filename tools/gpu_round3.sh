#!/bin/bash
TAG=${1:-r01c}
O=gpurun_out/$TAG
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log
tail -6 $O/pytest.log
timeout 600 python bench.py --no-cpu-baseline > $O/bench.json 2> $O/bench.err; echo "bench exit $?"; tail -3 $O/bench.err
python - <<PY
import json
d=json.load(open("$O/bench.json"))
print({k:d[k] for k in ("value","ms_per_step","e2e","gpu_launches")})
print(d["kernels"]["kernel_apply_fwd"], d["kernels"]["kernel_apply_bwd"])
print(d.get("denoise_720p"))
PY
