#!/bin/bash
# Round 2, gradient exchange inside the step at N GPUs: transports, grid size, overlap on / off, and a dry run
# (everything but the exchange kernel) that shows what the glue around it costs.
N=${1:-2}
OUT=gpurun_out/r02x
mkdir -p $OUT
PORT=29631
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((PORT++)) "$@"; }
one() {  # name, env...
  name=$1; shift
  env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((PORT++)) \
      bench.py --gpus $N --steps 60 --warmup 5 --kernel-pass-steps 3 --no-720p > $OUT/b2_n${N}_$name.json 2> $OUT/b2_n${N}_$name.err
  python - <<PY
import json
try:
    d = json.loads(open("$OUT/b2_n${N}_$name.json").read().strip().splitlines()[-1])
    print("%-28s %8.1f patches/s  %.3f ms  e2e %.1f" % ("$name", d["value"], d["ms_per_step"], d["e2e"]["value"]))
except Exception as e:
    print("$name", "no line", e)
PY
}
one multimem WCMC_EXCHANGE=multimem
one peer_b128 WCMC_EXCHANGE=peer WCMC_TUNE=exchange_blocks=128
one multimem_b16 WCMC_EXCHANGE=multimem WCMC_TUNE=exchange_blocks=16
one dry WCMC_EXCHANGE=multimem WCMC_EXCHANGE_DRY=1
one multimem_nooverlap WCMC_EXCHANGE=multimem WCMC_DDP_OVERLAP=0
one nccl WCMC_EXCHANGE=nccl
