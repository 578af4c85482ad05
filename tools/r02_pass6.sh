#!/bin/bash
# Round 2, GPU pass 6: fused last conv + softmax + kernel-apply (N1): parity, 720p timing fused vs not, regression.
O=gpurun_out/r02f
mkdir -p $O
rm -f gpurun_out/parity_records.jsonl
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "fused_last_conv or 720p or full_frame or conv_pair or conv_fwd_dgrad" > $O/pytest_new.log 2>&1; echo "new tests exit $?"; tail -15 $O/pytest_new.log
timeout 1800 python -m pytest tests -m gpu -q > $O/pytest.log 2>&1; echo "pytest exit $?"; tail -4 $O/pytest.log
cp gpurun_out/parity_records.jsonl $O/ 2>/dev/null
timeout 600 python bench.py --steps 20 --no-cpu-baseline > $O/bench.json 2> $O/bench.err; echo "bench exit $?"
WCMC_FUSE_KA=0 timeout 600 python bench.py --steps 20 --no-cpu-baseline > $O/bench_unfused.json 2> $O/bench_unfused.err; echo "bench unfused exit $?"
python - <<PY
import json
for f in ("$O/bench.json", "$O/bench_unfused.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["value"], 1), "patches/s", round(d["ms_per_step"], 3), "ms; e2e", round(d["e2e"]["value"], 1))
        print("  roofline:", d["roofline"]["kernel"][:40], d["roofline"]["achieved"], d["roofline"]["frac"])
        if "denoise_720p" in d: print("  720p:", {k: v for k, v in d["denoise_720p"].items() if k not in ("e2e_note", "workload")})
    except Exception as e:
        print(f, "unreadable", e)
PY
