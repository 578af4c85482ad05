#!/bin/bash
# 4 GPUs: what the exchange costs in the step -- dry (no exchange kernel, ranks free-running), auto, multimem, no overlap.
OUT=gpurun_out/r02x4
mkdir -p $OUT
PORT=29931
one() {  # name, env...
  name=$1; shift
  env "$@" timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port $((PORT++)) \
      bench.py --gpus 4 --steps 50 --warmup 5 --kernel-pass-steps 3 --no-720p > $OUT/bench_${name}.json 2> $OUT/bench_${name}.err
  python - <<PY
import json
try:
    d = json.loads(open("$OUT/bench_${name}.json").read().strip().splitlines()[-1])
    print("%-14s %8.1f patches/s  %.3f ms  e2e %.1f (%.3f ms) clk %s %s" % ("$name", d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["clocks"]["sm_mhz"], d["config"].get("grad_exchange_channels")))
except Exception as e:
    print("$name", "no line", e); print(open("$OUT/bench_${name}.err").read()[-800:])
PY
}
one dry WCMC_EXCHANGE=peer WCMC_EXCHANGE_DRY=1
one auto WCMC_EXCHANGE=auto
one multimem WCMC_EXCHANGE=multimem
one peer_nooverlap WCMC_EXCHANGE=peer WCMC_DDP_OVERLAP=0
one peer_b32 WCMC_EXCHANGE=peer WCMC_TUNE=exchange_blocks=32
