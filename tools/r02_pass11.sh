#!/bin/bash
O=gpurun_out/r02k
mkdir -p $O
timeout 600 ncu --set full --import-source on --clock-control none -k conv_igemm_kernel -s 9 -c 1 -f -o $O/conv64 python tools/conv_bench.py fwd 1 > $O/ncu.log 2>&1; echo "ncu exit $?"; tail -2 $O/ncu.log
python tools/ncu_summary.py $O/conv64.ncu-rep | head -30
ls -la $O
