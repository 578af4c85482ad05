#!/bin/bash
# Round 2, GPU pass 3: glue kernels (K9 / K12, Feistel permutations, row-staged weight packing) + full regression + bench.
O=gpurun_out/r02c
mkdir -p $O
rm -f gpurun_out/parity_records.jsonl
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "step_glue or device_permutation or batched_pack or grouped_wgrad" > $O/pytest_new.log 2>&1; echo "new tests exit $?"; tail -15 $O/pytest_new.log
timeout 1800 python -m pytest tests -m gpu -q > $O/pytest.log 2>&1; echo "pytest exit $?"; tail -12 $O/pytest.log
cp gpurun_out/parity_records.jsonl $O/ 2>/dev/null
timeout 600 python bench.py --steps 20 --no-cpu-baseline > $O/bench.json 2> $O/bench.err; echo "bench exit $?"; tail -c 400 $O/bench.err
python - <<PY
import json
for f in ("$O/bench.json",):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["value"], 1), "patches/s", round(d["ms_per_step"], 3), "ms; e2e", round(d["e2e"]["value"], 1), "launches", d["gpu_launches"])
        print("  roofline:", d["roofline"]["kernel"][:40], d["roofline"]["achieved"], d["roofline"]["frac"], "| step:", d.get("roofline_step"))
        for m in d["roofline_more"]:
            print("   ", m["kernel"][:60], m["achieved"], m["unit"], m["frac"], m["ms_per_step"])
        for k, v in sorted(d["kernels"].items(), key=lambda kv: -kv[1]["ms_per_step"]):
            print("      %-22s %3d calls %7.4f ms %5.1f%%" % (k, v["calls_per_step"], v["ms_per_step"], 100 * v["share_of_kernel_time"]))
        if "denoise_720p" in d: print("  720p:", d["denoise_720p"])
    except Exception as e:
        print(f, "unreadable", e)
PY
