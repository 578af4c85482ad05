#!/bin/bash
# 8-GPU box: exchange kernel vs NCCL (message timings), the step with transport auto and with NCCL, and N = 1 on the same box.
OUT=gpurun_out/r02x8
mkdir -p $OUT/w8
PORT=29831
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $((PORT++)) \
    tests/exchange_worker.py $OUT/w8 --bench > $OUT/exchange_n8.log 2>&1; echo "exchange worker exit $?"
python - <<PY
import json
r = json.load(open("$OUT/w8/rank0.json"))
for k, t in r["transports"].items():
    print(k, {a: b for a, b in t.items() if a != "bench"})
    for name, v in t.get("bench", {}).items():
        print("   ", name, v)
print("nccl", r.get("nccl"))
PY
one() {  # name, N, env...
  name=$1; N=$2; shift 2
  if [ $N -gt 1 ]; then
    env "$@" timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((PORT++)) \
        bench.py --gpus $N --steps 50 --warmup 5 --kernel-pass-steps 3 --no-720p > $OUT/bench_${name}.json 2> $OUT/bench_${name}.err
  else
    env "$@" timeout 400 python bench.py --gpus 1 --steps 50 --warmup 5 --kernel-pass-steps 3 --no-720p > $OUT/bench_${name}.json 2> $OUT/bench_${name}.err
  fi
  python - <<PY
import json
try:
    d = json.loads(open("$OUT/bench_${name}.json").read().strip().splitlines()[-1])
    print("%-12s %8.1f patches/s  %.3f ms  e2e %.1f (%.3f ms)  %s" % ("$name", d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["config"].get("grad_exchange_channels")))
except Exception as e:
    print("$name", "no line", e); print(open("$OUT/bench_${name}.err").read()[-800:])
PY
}
one n8_auto 8 WCMC_EXCHANGE=auto
one n8_nccl 8 WCMC_EXCHANGE=nccl
one n1 1 WCMC_EXCHANGE=auto
one n4_auto 4 WCMC_EXCHANGE=auto
