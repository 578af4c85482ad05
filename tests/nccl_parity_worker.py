"""Worker of tests/test_gpu_nccl.py (run under torchrun, one rank per GPU): the data-parallel KPCN+WCMC step as
bench.py / SCALE time it -- GraphedTrainStep with the NCCL gradient all-reduce inside the CUDA graph -- checked
against (a) its own per-shard gradients averaged over the ranks, bit for bit, (b) the oracle run per shard with
the same per-rank seeds and gradients averaged (/root/reference/support/interfaces.py:237-238 ->
:261 -> :271), and (c) replica equality after two steps.  Writes one JSON record per rank."""
import json
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def rel(a, b):
    a, b = a.double(), b.double()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def flat(models, grad=False):
    return torch.cat([(p.grad if grad else p.data).reshape(-1) for m in models.values() for p in m.parameters()])


def main():
    out_dir = sys.argv[1]
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    from wcmc_b200 import ddp, dropin, lib
    from wcmc_b200.synth import make_batch
    dropin.install()
    lib.init(local)
    from sbmc import KPCN
    from support.interfaces import KPCNInterface
    from support.losses import FeatureMSE, RelativeMSE
    from support.networks import PathNet
    from tests._oracle_loader import load_oracle
    from wcmc_b200.engine import GraphedTrainStep
    oracle = load_oracle()

    def make(KP, PN, seed):
        torch.manual_seed(seed)
        return {"dncnn": KP(39), "backbone_diffuse": PN(ic=36, outc=3), "backbone_specular": PN(ic=36, outc=3)}

    def interface(models, graph_sync):
        optims = {"optim_" + k: torch.optim.Adam(m.parameters(), lr=1e-4) for k, m in models.items()}
        lf = {"l_diffuse": torch.nn.L1Loss(), "l_specular": torch.nn.L1Loss(), "l_recon": torch.nn.L1Loss(),
              "l_test": RelativeMSE(), "l_manif": FeatureMSE(non_local=True, rng="cpu")}
        itf = KPCNInterface(models, optims, lf, types.SimpleNamespace(model_name="t"), use_llpm_buf=True,
                            manif_learn=True, w_manif=0.1, train_branches=True)
        if graph_sync:
            itf.grad_sync = ddp.GradAllReduce()
        itf.to_train_mode()
        return itf, optims

    models = make(KPCN, PathNet, 100 + rank)              # replicas start DIFFERENT on purpose
    for m in models.values():
        m.cuda()
    ddp.broadcast_parameters(models)
    w0 = flat(models).clone()
    gathered = [torch.empty_like(w0) for _ in range(world)]
    dist.all_gather(gathered, w0)
    assert all(torch.equal(g, gathered[0]) for g in gathered), "broadcast_parameters did not equalise the replicas"

    ref_models = make(oracle.KPCN, oracle.PathNet, 0)
    shadow = make(KPCN, PathNet, 0)                       # our kernels, eager, no exchange: this rank's own gradients
    for k in models:
        ref_models[k].load_state_dict(models[k].state_dict())
        shadow[k].load_state_dict(models[k].state_dict())
        ref_models[k].cuda()
        shadow[k].cuda()
    batch = {k: v.cuda() for k, v in make_batch(batch=2, spp=2, size=64, seed=1234 + rank).items()}   # this rank's shard

    itf, _ = interface(models, True)
    graphed = GraphedTrainStep(itf, batch)
    assert graphed.sync_in_graph and graphed.fused is not None, "the NCCL all-reduce must be captured in the graph"
    assert torch.equal(flat(models), w0), "warm-up / capture must not update the weights"
    rec = {"rank": rank, "world": world}

    # ---- step 1 --------------------------------------------------------------------------------------------------
    seed = 7000 + rank
    sitf, _ = interface(shadow, False)
    sitf.preprocess(batch)
    torch.manual_seed(seed)
    sitf.train_batch(batch, grad_hook_mode=True)          # gradients only
    g_local = flat(shadow, grad=True).clone()
    g_all = [torch.empty_like(g_local) for _ in range(world)]
    dist.all_gather(g_all, g_local)
    g_mean = g_all[0].clone()
    for g in g_all[1:]:
        g_mean += g
    g_mean /= world
    torch.manual_seed(seed)
    graphed(batch)                                        # graph replay: fwd + 2 x bwd + NCCL all-reduce + clip + Adam
    g_step = flat(models, grad=True).clone()              # the fused kernel leaves the CLIPPED mean gradient
    rec["step_vs_mean_of_own_shards_max_abs"] = float((g_step - g_mean.clamp(-1, 1)).abs().max())
    rec["step_vs_mean_of_own_shards_rel"] = rel(g_step, g_mean.clamp(-1, 1))

    # the oracle per shard: forward + the two backward passes (no clip / Adam), same permutations through the seed
    torch.manual_seed(seed)
    oracle_loss, _, _ = oracle.ref.kpcn_losses(ref_models, batch, use_llpm_buf=True, manif_learn=True, w_manif=0.1)
    go = flat(ref_models, grad=True).clone()
    go_all = [torch.empty_like(go) for _ in range(world)]
    dist.all_gather(go_all, go)
    go_mean = torch.stack(go_all).mean(0).clamp(-1, 1)
    rec["step_vs_oracle_mean_rel"] = rel(g_step, go_mean)
    rec["local_vs_oracle_local_rel"] = rel(g_local, go)
    for k, v in oracle_loss.items():
        rec["loss_rel_" + k] = rel(itf.m_losses["m_" + k], v)

    # ---- step 2 + replica equality -----------------------------------------------------------------------------------
    torch.manual_seed(seed + 1)
    graphed(batch)
    w2 = flat(models)
    gathered = [torch.empty_like(w2) for _ in range(world)]
    dist.all_gather(gathered, w2)
    rec["replicas_bit_identical_after_2_steps"] = bool(all(torch.equal(g, gathered[0]) for g in gathered))
    rec["weights_moved"] = bool(not torch.equal(w2, w0))
    torch.cuda.synchronize()
    with open(os.path.join(out_dir, "rank%d.json" % rank), "w") as f:
        json.dump(rec, f)
    dist.barrier()
    torch.cuda.synchronize()
    sys.stdout.flush()
    os._exit(0)   # see bench.py::_leave


if __name__ == "__main__":
    main()
