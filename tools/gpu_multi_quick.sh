#!/bin/bash
# N-GPU bench exactly as the driver launches it (own arm only, short)
N=${1:-2}; O=gpurun_out/multi${N}_final; mkdir -p $O
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 10 --warmup 3 > $O/bench.json 2> $O/bench.err; echo "exit $?"
tail -3 $O/bench.err | cut -c1-300; cut -c1-700 $O/bench.json
