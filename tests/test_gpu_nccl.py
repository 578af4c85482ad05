"""GPU, >= 2 devices, NCCL: on-hardware parity of the data-parallel step.  The worker
(tests/nccl_parity_worker.py, one rank per GPU under torchrun) drives `GraphedTrainStep` with the gradient
all-reduce captured inside the CUDA graph -- the path bench.py's N > 1 runs time -- and records:
  * gradient after the step == clip(mean over ranks of each rank's own shard gradient), the shard gradients taken
    with the same kernels run eagerly without the exchange (two ranks: one fp32 add, one exact halving -- equal up to
    the order of the fp32 atomics that sum the bias gradients);
  * the same quantity against the ORACLE run per shard with the same per-rank seeds, gradients averaged;
  * replicas bit-identical after two steps."""
import json
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (run under `gpurun --gpus 2`)")
def test_graphed_step_with_nccl_allreduce_world2(tmp_path):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "nccl_parity_worker.py"),
           str(tmp_path)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, (r.stdout[-3000:], r.stderr[-3000:])
    recs = [json.load(open(tmp_path / ("rank%d.json" % i))) for i in range(2)]
    out = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out):
        with open(os.path.join(out, "nccl_parity.json"), "w") as f:
            json.dump(recs, f, indent=1)
    for rec in recs:
        assert rec["replicas_bit_identical_after_2_steps"] and rec["weights_moved"], rec
        # two ranks: (g0 + g1) / 2 is one rounding + an exact halving, identical in NCCL and in torch; what is left
        # is the summation order of the bias-gradient column sums (fp32 atomics in the data-gradient epilogue)
        assert rec["step_vs_mean_of_own_shards_rel"] < 1e-5, rec
        # vs the fp32 oracle per shard, gradients averaged: small-case bound of tests/test_gpu_parity.py
        # (a 16-bit forward flips ~eps of the ReLU masks; 1e-2 holds at the north-star size)
        assert rec["step_vs_oracle_mean_rel"] < 1e-1, rec
        for k, v in rec.items():
            if k.startswith("loss_rel_"):
                assert v < (5e-3 if "manif" in k else 1e-3), (k, v)
