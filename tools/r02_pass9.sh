#!/bin/bash
O=gpurun_out/r02i
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_preprocess.py -q > $O/pytest.log 2>&1; echo "pytest exit $?"; tail -5 $O/pytest.log
python - <<PY
import json, sys
sys.path.insert(0, ".")
import bench
pk, _ = bench.peaks()
from wcmc_b200 import lib
lib.init(0)
print(json.dumps(bench.bench_preprocess(pk)))
PY
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:preprocess -c 12 python - <<PY 2>&1 | grep -E "preprocess|duration|dram__bytes" | head -40
import sys, torch
sys.path.insert(0, ".")
from wcmc_b200 import lib, preprocess
lib.init(0)
raw = torch.rand(720, 1280, 4, 104, device="cuda")
for _ in range(2):
    preprocess.preprocess_kpcn(raw); preprocess.preprocess_llpm(raw)
torch.cuda.synchronize()
PY
