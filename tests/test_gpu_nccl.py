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


def _run_exchange_worker(tmp_path, world, extra=()):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr",
           "127.0.0.1", "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "exchange_worker.py"),
           str(tmp_path), *extra]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, (r.stdout[-3000:], r.stderr[-3000:])
    recs = [json.load(open(tmp_path / ("rank%d.json" % i))) for i in range(world)]
    out = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out):
        with open(os.path.join(out, "peer_exchange_world%d.json" % world), "w") as f:
            json.dump(recs, f, indent=1)
    return recs


def _check_exchange(recs, world):
    ran = 0
    for rec in recs:
        for name, t in rec["transports"].items():
            if "unavailable" in t:
                continue
            ran += 1
            assert t["replicas_identical"], (name, t)
            assert t["max_abs_err"] < 1e-5 and t["graph_two_channels_max_abs_err"] < 1e-5, (name, t)
            if world <= 2:     # one fp32 add (commutative) and an exact scale: nothing left to differ in
                assert t["bitwise_equal_to_nccl"], (name, t)
    return ran


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (run under `gpurun --gpus 2`)")
def test_peer_exchange_kernel_matches_nccl_world2(tmp_path):
    """`wcmc_grad_exchange` (multicast and peer-load variants) == NCCL all-reduce on ragged sizes, both channels in
    flight, eager and graph-replayed.  The symmetric allocation must exist on an NVSwitch box: no skip here."""
    recs = _run_exchange_worker(tmp_path, 2)
    assert _check_exchange(recs, 2) >= 2, recs     # at least one transport on both ranks
    assert all("unavailable" not in r["transports"]["peer"] for r in recs), recs


def test_peer_exchange_kernel_single_rank(tmp_path):
    """The same kernel with a one-rank group (what a 1-GPU box can exercise: flags, epochs, slices, graph replay)."""
    recs = _run_exchange_worker(tmp_path, 1)
    if _check_exchange(recs, 1) == 0:
        pytest.skip("symmetric memory unavailable with one rank: %s" % recs[0]["transports"])


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (run under `gpurun --gpus 2`)")
def test_dataparallel_two_devices_and_non_current_device():
    """The reference's own multi-GPU mode (single process, `nn.DataParallel`, replicas in threads,
    /root/reference/train_kpcn.py:266-269) and its `--single_gpu --device_id 1` mode (models on a device that is not
    torch's current one, :259-263): per-device kernel attributes, thread-local launch state and the device guard of the
    ctypes layer.  Outputs and gradients must equal a plain single-device run."""
    import torch.nn as nn
    from wcmc_b200 import dropin, lib
    from wcmc_b200.synth import make_batch
    dropin.install()
    lib.init(0)
    from sbmc import KPCN
    from support.networks import PathNet
    torch.manual_seed(0)
    base = KPCN(34).cuda(0)
    data = {k: v.cuda(0) for k, v in make_batch(batch=4, size=48, seed=3, paths=False, llpm_channel=False).items()}
    want = base(data)
    (want["diffuse"].mean() + want["specular"].mean()).backward()
    g_want = torch.cat([p.grad.flatten() for p in base.parameters()]).clone()
    # (a) the same model living on cuda:1 while torch's current device stays cuda:0
    assert torch.cuda.current_device() == 0
    other = KPCN(34)
    other.load_state_dict(base.state_dict())
    other.cuda(1)
    out1 = other({k: v.cuda(1) for k, v in data.items()})
    (out1["diffuse"].mean() + out1["specular"].mean()).backward()
    g1 = torch.cat([p.grad.flatten() for p in other.parameters()])
    assert out1["radiance"].device.index == 1
    torch.testing.assert_close(out1["radiance"].cpu(), want["radiance"].cpu(), rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(g1.cpu(), g_want.cpu(), rtol=1e-4, atol=1e-7)
    # (b) nn.DataParallel over both devices: the batch is scattered 2 + 2, replicas run in two threads
    for p in base.parameters():
        p.grad = None
    dp = nn.DataParallel(base, device_ids=[0, 1], output_device=0)
    out = dp(data)
    assert tuple(out["radiance"].shape) == tuple(want["radiance"].shape)
    torch.testing.assert_close(out["radiance"], want["radiance"], rtol=1e-5, atol=1e-6)
    # gradients: mean over the gathered batch == mean of the two half-batch means
    (out["diffuse"].mean() + out["specular"].mean()).backward()
    g_dp = torch.cat([p.grad.flatten() for p in base.parameters()])
    torch.testing.assert_close(g_dp, g_want, rtol=2e-3, atol=1e-6)
    # the path-embedding network (two streams, weight-norm override per thread) under DataParallel as well
    torch.manual_seed(1)
    pn = PathNet(ic=36, outc=3).cuda(0)
    pb = {k: v.cuda(0) for k, v in make_batch(batch=4, spp=2, size=32, seed=5).items()}
    with torch.no_grad():
        p_want = pn(pb)
        p_dp = nn.DataParallel(pn, device_ids=[0, 1], output_device=0)(pb)
    torch.testing.assert_close(p_dp, p_want, rtol=1e-5, atol=1e-6)
