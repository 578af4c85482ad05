#!/bin/bash
# Round 2, gradient exchange: NVSwitch kernel vs NCCL on N GPUs (correctness + message-size timing), then the step.
# usage (on the GPU box): bash tools/r02_exchange.sh N
N=${1:-2}
OUT=gpurun_out/r02x
mkdir -p $OUT/w$N
PORT=29531
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((PORT++)) "$@"; }
run tests/exchange_worker.py $OUT/w$N --bench > $OUT/exchange_n$N.log 2>&1; echo "exchange worker exit $?"
tail -5 $OUT/exchange_n$N.log
python - <<PY
import json
r = json.load(open("$OUT/w$N/rank0.json"))
for k, t in r["transports"].items():
    print(k, {a: b for a, b in t.items() if a != "bench"})
    for name, v in t.get("bench", {}).items():
        print("   ", name, v)
print("nccl", r.get("nccl"))
PY
for X in nccl peer multimem; do
  WCMC_EXCHANGE=$X run bench.py --gpus $N --steps 60 --warmup 5 --kernel-pass-steps 3 --no-720p > $OUT/bench_n${N}_$X.json 2> $OUT/bench_n${N}_$X.err
  echo "bench $X exit $?"
  python - <<PY
import json
try:
    d = json.loads(open("$OUT/bench_n${N}_$X.json").read().strip().splitlines()[-1])
    print("$X", d["value"], d["ms_per_step"], d["e2e"]["value"], d["config"]["grad_exchange"])
except Exception as e:
    print("$X", "no line", e)
PY
done
