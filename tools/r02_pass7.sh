#!/bin/bash
# Round 2, GPU pass 7: glue round 2 (absmax scale, cropped loss, bias pool, Adam non-finite guard), N3 kernels rewritten,
# full regression, bench, conv knob experiments.
O=gpurun_out/r02g
mkdir -p $O
rm -f gpurun_out/parity_records.jsonl
timeout 1800 python -m pytest tests -m gpu -q > $O/pytest.log 2>&1; echo "pytest exit $?"; tail -8 $O/pytest.log
cp gpurun_out/parity_records.jsonl $O/ 2>/dev/null
timeout 600 python bench.py --steps 20 --no-cpu-baseline > $O/bench.json 2> $O/bench.err; echo "bench exit $?"; tail -c 300 $O/bench.err
for T in conv_plane_slots=3 conv_plane_slots=4; do
  WCMC_TUNE=$T timeout 300 python bench.py --steps 20 --no-cpu-baseline --no-720p > $O/bench_$T.json 2> $O/bench_$T.err; echo "bench $T exit $?"
done
python - <<PY
import json, glob
for f in sorted(glob.glob("$O/bench*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        k = d["kernels"]
        print(f, round(d["value"], 1), "patches/s", round(d["ms_per_step"], 3), "ms; e2e", round(d["e2e"]["value"], 1), "launches", d["gpu_launches"],
              {n: k[n]["ms_per_step"] for n in ("conv2d_k5", "conv2d_k3", "conv2d_wgrad_k5", "conv2d_wgrad_k3") if n in k})
        if "denoise_720p" in d: print("  720p:", {a: b for a, b in d["denoise_720p"].items() if "ms" in a}); print("  n3:", d.get("preprocess_n3"))
    except Exception as e:
        print(f, "unreadable", e)
PY
