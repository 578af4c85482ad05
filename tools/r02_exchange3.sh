#!/bin/bash
# Gradient exchange at N GPUs: kernel vs NCCL message timings, then the step with transport auto / peer / multimem / nccl.
N=${1:-2}
OUT=gpurun_out/r02x
mkdir -p $OUT/w$N
PORT=29731
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((PORT++)) \
    tests/exchange_worker.py $OUT/w$N --bench > $OUT/exchange3_n$N.log 2>&1; echo "exchange worker exit $?"
python - <<PY
import json
r = json.load(open("$OUT/w$N/rank0.json"))
for k, t in r["transports"].items():
    print(k, {a: b for a, b in t.items() if a != "bench"})
    for name, v in t.get("bench", {}).items():
        print("   ", name, v)
print("nccl", r.get("nccl"))
PY
one() {  # name, env...
  name=$1; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((PORT++)) \
      bench.py --gpus $N --steps 50 --warmup 5 --kernel-pass-steps 3 --no-720p > $OUT/b3_n${N}_$name.json 2> $OUT/b3_n${N}_$name.err
  python - <<PY
import json
try:
    d = json.loads(open("$OUT/b3_n${N}_$name.json").read().strip().splitlines()[-1])
    print("%-20s %8.1f patches/s  %.3f ms  e2e %.1f  %s" % ("$name", d["value"], d["ms_per_step"], d["e2e"]["value"], d["config"].get("grad_exchange_channels")))
except Exception as e:
    print("$name", "no line", e)
PY
}
for X in "$@"; do :; done
one auto WCMC_EXCHANGE=auto
one peer WCMC_EXCHANGE=peer
one multimem WCMC_EXCHANGE=multimem
if [ "${2:-}" != "short" ]; then one nccl WCMC_EXCHANGE=nccl; fi
