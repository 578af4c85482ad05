#!/bin/bash
# First GPU pass of round 2: what round 1 left unvalidated or unmeasured.
O=gpurun_out/${1:-r02a}
mkdir -p $O
# 1. N3 preprocessing kernels (never run on a GPU)
WCMC_UNVALIDATED=1 timeout 300 python -m pytest tests/test_gpu_preprocess.py -x -q > $O/pytest_n3.log 2>&1; echo "N3 pytest exit $?"; tail -3 $O/pytest_n3.log
# 2. regression incl. the pair-vs-single bit-identity test added after the last GPU pass
timeout 600 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest exit $?"; tail -3 $O/pytest.log
# 3. launch-shape knobs: deeper halo ring for the short 3x3 layers, pair threshold
for T in conv_plane_slots=4 conv_plane_slots=3 conv_pair_min_clk=8000 conv_item_clk=3000; do
  WCMC_TUNE=$T timeout 300 python bench.py --no-cpu-baseline --no-720p > $O/bench_$T.json 2> $O/bench_$T.err; echo "bench $T exit $?"
done
timeout 300 python bench.py --no-cpu-baseline > $O/bench.json 2> $O/bench.err
python - <<PY
import json, glob
for f in sorted(glob.glob("$O/bench*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        k = d.get("kernels", {})
        print(f, round(d["value"], 1), round(d["ms_per_step"], 3), {n: k[n]["ms_per_step"] for n in ("conv2d_k5", "conv2d_k3") if n in k})
    except Exception as e:
        print(f, "unreadable", e)
PY
