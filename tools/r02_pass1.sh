#!/bin/bash
# Round 2, GPU pass 1: regression of batch A (device guard, optimiser state, staged permutations, unmasked N3), parity
# records, the ReLU-flip ablation at the north-star size, compute-sanitizer on the barrier-heavy kernels, baseline bench.
O=gpurun_out/r02a
mkdir -p $O
rm -f gpurun_out/parity_records.jsonl
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_parity.py -q -k "grouped_wgrad" > $O/pytest_wgrad.log 2>&1; echo "grouped wgrad pytest exit $?"; tail -15 $O/pytest_wgrad.log
timeout 1800 python -m pytest tests -m gpu -q > $O/pytest.log 2>&1; echo "pytest exit $?"; tail -25 $O/pytest.log
cp gpurun_out/parity_records.jsonl $O/ 2>/dev/null
timeout 300 python tests/ablation_relu_flip.py > $O/relu_flip_ablation.txt 2> $O/relu_flip_ablation.err; echo "ablation exit $?"; grep -v Warn $O/relu_flip_ablation.txt | tail -8
# compute-sanitizer: memcheck on the unit tests of the kernels whose correctness rests on mbarrier / cluster protocols
SAN="compute-sanitizer --error-exitcode 7 --print-limit 20"
timeout 900 $SAN --tool memcheck python -m pytest tests/test_gpu_parity.py -x -q -k "test_conv_pair_launch_is_bit_identical or (test_conv_fwd_dgrad_wgrad_vs_fp64 and 5-100-100) or (test_fused_mlp_backward_vs_generic_path and 16) or test_conv_tilings_agree" > $O/sanitizer_memcheck.log 2>&1; echo "memcheck exit $?"; tail -4 $O/sanitizer_memcheck.log
timeout 900 $SAN --tool racecheck python -m pytest tests/test_gpu_parity.py -x -q -k "(test_conv_pair_launch_is_bit_identical and not 441) or (test_fused_mlp_backward_vs_generic_path and 16)" > $O/sanitizer_racecheck.log 2>&1; echo "racecheck exit $?"; tail -4 $O/sanitizer_racecheck.log
timeout 600 $SAN --tool synccheck python -m pytest tests/test_gpu_parity.py -x -q -k "(test_conv_fwd_dgrad_wgrad_vs_fp64 and 5-100-100)" > $O/sanitizer_synccheck.log 2>&1; echo "synccheck exit $?"; tail -4 $O/sanitizer_synccheck.log
timeout 600 python bench.py --steps 20 > $O/bench.json 2> $O/bench.err; echo "bench exit $?"; tail -c 600 $O/bench.err
timeout 300 python bench.py --steps 10 --perm-rng cpu --no-cpu-baseline --no-720p > $O/bench_permcpu.json 2> $O/bench_permcpu.err; echo "bench cpu-perm exit $?"
python - <<PY
import json
for f in ("$O/bench.json", "$O/bench_permcpu.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["value"], 1), "patches/s", round(d["ms_per_step"], 3), "ms; e2e", round(d["e2e"]["value"], 1))
        print("  roofline:", d["roofline"]["kernel"][:40], d["roofline"]["achieved"], d["roofline"]["frac"], "| step:", d.get("roofline_step"))
        for m in d["roofline_more"]:
            print("   ", m["kernel"][:60], m["achieved"], m["unit"], m["frac"], m["ms_per_step"])
        if "denoise_720p" in d: print("  720p:", d["denoise_720p"])
    except Exception as e:
        print(f, "unreadable", e)
PY
