#!/bin/bash
# Round 2: the scaling run as the driver launches it (weak scaling, 8 patches per GPU), N = 1, 2, 4, 8 on one box.
O=gpurun_out/r02s
mkdir -p $O
nvidia-smi --query-gpu=index,name --format=csv,noheader > $O/smi.txt
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu-baseline --no-720p > $O/bench_n1.json 2> $O/bench_n1.err; echo "N=1 exit $?"
for N in 2 4 8; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N \
    bench.py --gpus $N --steps 20 --warmup 3 > $O/bench_n$N.json 2> $O/bench_n$N.err; echo "N=$N exit $?"
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29520 \
    bench.py --gpus 8 --steps 20 --warmup 3 --global-batch 64 > $O/bench_n8_global64.json 2> $O/bench_n8_global64.err; echo "N=8 global 64 exit $?"
python - <<PY
import json
base = None
for n in (1, 2, 4, 8):
    try:
        d = json.loads(open("$O/bench_n%d.json" % n).read().strip().splitlines()[-1])
        if n == 1: base = (d["value"], d["e2e"]["value"])
        print("N=%d  %.1f patches/s  %.3f ms/step  e2e %.1f  eff %.3f  e2e eff %.3f  clocks %s" % (n, d["value"], d["ms_per_step"], d["e2e"]["value"], d["value"] / (n * base[0]), d["e2e"]["value"] / (n * base[1]), d["clocks"]))
    except Exception as e:
        print(n, "unreadable", e)
try:
    d = json.loads(open("$O/bench_n8_global64.json").read().strip().splitlines()[-1])
    print("N=8 global 64:", round(d["value"], 1), round(d["ms_per_step"], 3), d["scaling"], d["config"]["per_gpu_batch"])
except Exception as e:
    print("global64 unreadable", e)
PY
grep -v "Warning\|WeightNorm\|^$\|OMP_NUM\|\*\*\*" $O/bench_n8.err | tail -5
