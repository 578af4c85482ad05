#!/bin/bash
O=gpurun_out/r02h
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_preprocess.py tests/test_gpu_nccl.py -q > $O/pytest.log 2>&1; echo "pytest exit $?"; tail -25 $O/pytest.log
python - <<PY
import json, sys
sys.path.insert(0, ".")
import bench
pk, _ = bench.peaks()
from wcmc_b200 import lib
lib.init(0)
print(json.dumps(bench.bench_preprocess(pk)))
PY
