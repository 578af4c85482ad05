#!/bin/bash
# N-GPU bench exactly as the driver launches it
N=${1:-2}; O=gpurun_out/multi$N; mkdir -p $O
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 10 --warmup 3 > $O/bench.json 2> $O/bench.err; echo "exit $?"
tail -3 $O/bench.err; cat $O/bench.json | cut -c1-900
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > $O/bench_ref.json 2>> $O/bench.err; echo "ref exit $?"; cut -c1-400 $O/bench_ref.json
