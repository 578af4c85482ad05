#!/bin/bash
O=gpurun_out/r02j
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "resident or conv_tilings or conv_fwd_dgrad or conv_pair" > $O/pytest_new.log 2>&1; echo "new tests exit $?"; tail -8 $O/pytest_new.log
timeout 1800 python -m pytest tests -m gpu -q > $O/pytest.log 2>&1; echo "pytest exit $?"; tail -4 $O/pytest.log
timeout 300 python tools/conv_bench.py fwd 5 > $O/conv_bench_fwd.txt 2>&1; grep unet $O/conv_bench_fwd.txt
timeout 300 python tools/conv_bench.py dgrad 5 > $O/conv_bench_dgrad.txt 2>&1; grep unet $O/conv_bench_dgrad.txt
for T in conv_resident=1 conv_resident=0; do
  WCMC_TUNE=$T timeout 300 python bench.py --steps 20 --no-cpu-baseline --no-720p > $O/bench_$T.json 2> $O/bench_$T.err; echo "bench $T exit $?"
done
python - <<PY
import json, glob
for f in sorted(glob.glob("$O/bench*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        k = d["kernels"]
        print(f, round(d["value"], 1), "patches/s", round(d["ms_per_step"], 3), "ms; e2e", round(d["e2e"]["value"], 1),
              {n: k[n]["ms_per_step"] for n in ("conv2d_k5", "conv2d_k3", "conv2d_wgrad_k5", "conv2d_wgrad_k3") if n in k})
    except Exception as e:
        print(f, "unreadable", e)
PY
