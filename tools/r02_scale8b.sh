#!/bin/bash
# 8 GPUs: the step with / without binding each rank to its GPU's NUMA node (end-to-end: 8 x 197 MB of H2D per step)
OUT=gpurun_out/r02x8b
mkdir -p $OUT
PORT=30031
one() {  # name, env...
  name=$1; shift
  env "$@" timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $((PORT++)) \
      bench.py --gpus 8 --steps 50 --warmup 5 --kernel-pass-steps 3 --no-720p > $OUT/bench_${name}.json 2> $OUT/bench_${name}.err
  python - <<PY
import json
try:
    d = json.loads(open("$OUT/bench_${name}.json").read().strip().splitlines()[-1])
    print("%-10s %8.1f patches/s  %.3f ms  e2e %.1f (%.3f ms) clk %s numa %s %s" % ("$name", d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["clocks"]["sm_mhz"], d["config"].get("host_cores_bound_to_gpu_numa_node"), d["config"].get("grad_exchange_channels")))
except Exception as e:
    print("$name", "no line", e); print(open("$OUT/bench_${name}.err").read()[-800:])
PY
}
one numa WCMC_NUMA_BIND=1
one nonuma WCMC_NUMA_BIND=0
numactl -H 2>/dev/null | head -5; nvidia-smi topo -m 2>/dev/null | head -14
