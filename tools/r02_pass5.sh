#!/bin/bash
# Round 2, GPU pass 5: full regression + bench after the wgrad chains / DDP changes; stream experiment.
O=gpurun_out/r02e
mkdir -p $O
rm -f gpurun_out/parity_records.jsonl
timeout 1800 python -m pytest tests -m gpu -q > $O/pytest.log 2>&1; echo "pytest exit $?"; tail -4 $O/pytest.log
cp gpurun_out/parity_records.jsonl $O/ 2>/dev/null
timeout 600 python bench.py --steps 20 > $O/bench.json 2> $O/bench.err; echo "bench exit $?"
WCMC_BRANCH_STREAMS=0 timeout 600 python bench.py --steps 20 --no-cpu-baseline --no-720p > $O/bench_onestream.json 2> $O/bench_onestream.err; echo "bench one stream exit $?"
timeout 600 python bench.py --steps 300 --no-cpu-baseline --no-720p --kernel-pass-steps 10 > $O/bench_300steps.json 2> $O/bench_300.err; echo "bench 300 steps exit $?"
python - <<PY
import json
for f in ("$O/bench.json", "$O/bench_onestream.json", "$O/bench_300steps.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["value"], 1), "patches/s", round(d["ms_per_step"], 3), "ms; e2e", round(d["e2e"]["value"], 1), "launches", d["gpu_launches"], d["clocks"])
        print("  roofline:", d["roofline"]["kernel"][:40], d["roofline"]["achieved"], d["roofline"]["frac"], "| step:", d.get("roofline_step"))
        for m in d["roofline_more"]:
            print("   ", m["kernel"][:60], m["achieved"], m["unit"], m["frac"], m["ms_per_step"])
        if "denoise_720p" in d: print("  720p:", d["denoise_720p"]); print("  n3:", d.get("preprocess_n3")); print("  cpu:", d.get("cpu_baseline"))
    except Exception as e:
        print(f, "unreadable", e)
PY
