#!/bin/bash
# Round-1 re-entry GPU pass: regression (pytest -m gpu + bench), tcgen05 issue-rate probe, CTA-pair conv check + A/B.
TAG=${1:-r01d}
O=gpurun_out/$TAG
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > $O/gpu.txt 2>&1
timeout 700 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log
tail -4 $O/pytest.log
timeout 400 python bench.py > $O/bench.json 2> $O/bench.err; echo "bench exit $?"; tail -2 $O/bench.err
timeout 120 build/mma_probe > $O/mma_probe.jsonl 2> $O/mma_probe.err; echo "probe exit $?"
cut -c1-260 $O/mma_probe.jsonl | python -c "
import sys, json
for l in sys.stdin:
    try:
        d = json.loads(l)
        print('%-40s clk/mma %7.2f (floor %5.1f)  ghz %.3f  TF/s %7.1f' % (d['variant'], d['clk_per_mma'], d['floor_clk'], d['sm_ghz'], d['tflops_in_kernel']))
    except Exception as e:
        print(l.strip()[:200])
"
timeout 300 python tools/pair_check.py 5 > $O/pair_check.txt 2>&1; PC=$?; echo "pair_check exit $PC"; tail -30 $O/pair_check.txt | cut -c1-220
timeout 200 python tools/conv_bench.py knobs 5 > $O/conv_knobs.txt 2>&1; grep -i "interleav\|production" $O/conv_knobs.txt
if [ $PC -eq 0 ]; then
  for T in conv_pair=1 conv_interleave=1 conv_pair=1,conv_interleave=1; do
    WCMC_TUNE=$T timeout 300 python bench.py --no-cpu-baseline > $O/bench_$T.json 2> $O/bench_$T.err; echo "bench $T exit $?"
  done
fi
python - <<PY
import json, glob
for f in sorted(glob.glob("$O/bench*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, {k: d.get(k) for k in ("value", "ms_per_step", "gpu_launches")}, d.get("e2e", {}).get("value"), d.get("roofline", {}).get("achieved"), d.get("denoise_720p"))
    except Exception as e:
        print(f, "unreadable", e)
PY
