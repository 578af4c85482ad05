#!/bin/bash
O=gpurun_out/r02d
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_nccl.py -q -x > $O/pytest_nccl.log 2>&1; echo "nccl pytest exit $?"; tail -5 $O/pytest_nccl.log
cp gpurun_out/nccl_parity.json $O/ 2>/dev/null; cat $O/nccl_parity.json
for OV in 1 0 1; do
  WCMC_DDP_OVERLAP=$OV timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 20 --warmup 3 > $O/bench2_ov$OV.json 2> $O/bench2_ov$OV.err; echo "bench N=2 overlap=$OV exit $?"
  python - <<PY
import json
d = json.loads(open("$O/bench2_ov$OV.json").read().strip().splitlines()[-1])
print(round(d["value"], 1), "patches/s", round(d["ms_per_step"], 3), "ms; e2e", round(d["e2e"]["value"], 1), d["config"].get("grad_exchange"))
PY
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus 2 --steps 10 --warmup 3 --global-batch 64 > $O/bench2_global64.json 2> $O/bench2_global64.err; echo "bench N=2 global 64 exit $?"
python - <<PY
import json
d = json.loads(open("$O/bench2_global64.json").read().strip().splitlines()[-1])
print(round(d["value"], 1), "patches/s", round(d["ms_per_step"], 3), "ms; e2e", round(d["e2e"]["value"], 1), d["scaling"], d["config"]["global_batch"], d["config"]["per_gpu_batch"])
PY
grep -v "Warning\|WeightNorm\|^$\|OMP_NUM\|\*\*\*" $O/bench2_ov1.err | tail -5
