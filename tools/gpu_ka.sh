#!/bin/bash
O=gpurun_out/${1:-ka}; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_parity.py -k "kernel_apply or kpcn_matches" -x -q 2>&1 | tail -3
timeout 300 python tools/ka_bench.py 7 2>&1 | tee $O/ka_bench.txt
