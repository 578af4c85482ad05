#!/bin/bash
# Round 2, GPU pass 4 (2 GPUs): on-hardware multi-rank parity (NCCL all-reduce inside the step graph), the overlapped
# gradient exchange against the plain one, teardown without os._exit.
O=gpurun_out/r02d
mkdir -p $O
nvidia-smi --query-gpu=index,name --format=csv > $O/smi.txt
timeout 900 python -m pytest tests/test_gpu_nccl.py tests/test_gpu_preprocess.py -q -x > $O/pytest_nccl.log 2>&1; echo "nccl + n3 pytest exit $?"; tail -15 $O/pytest_nccl.log
cp gpurun_out/nccl_parity.json $O/ 2>/dev/null
for OV in 1 0; do
  WCMC_DDP_OVERLAP=$OV timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 20 --warmup 3 > $O/bench2_ov$OV.json 2> $O/bench2_ov$OV.err; echo "bench N=2 overlap=$OV exit $?"
done
timeout 600 python bench.py --steps 20 --no-cpu-baseline > $O/bench1.json 2> $O/bench1.err; echo "bench N=1 exit $?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus 2 --steps 10 --warmup 3 --global-batch 64 > $O/bench2_global64.json 2> $O/bench2_global64.err; echo "bench N=2 global 64 exit $?"
python - <<PY
import json, glob
for f in sorted(glob.glob("$O/bench*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["value"], 1), "patches/s", round(d["ms_per_step"], 3), "ms; e2e", round(d["e2e"]["value"], 1), d["config"].get("grad_exchange"), d["scaling"], d["config"]["global_batch"])
        if "denoise_720p" in d: print("  720p:", {k: v for k, v in d["denoise_720p"].items() if "ms" in k}, "n3:", d.get("preprocess_n3"))
    except Exception as e:
        print(f, "unreadable", e)
PY
tail -3 $O/bench2_ov1.err
