"""CPU, world_size 2, gloo: the data-parallel plumbing of the train step (wcmc_b200/ddp.py and the
`grad_sync` hook of the drop-in KPCNInterface).  The CUDA kernels are not involved: stand-in torch
modules with the KPCN / PathNet call contract take their place, so what is checked is the host
logic -- replicas start equal, every rank applies the mean gradient of all shards between backward
and clip/Adam (support/interfaces.py:237-238 -> :261 -> :271), and replicas stay bit-identical."""
import os
import socket
import types

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


class _TinyKPCN(torch.nn.Module):
    """dict -> {'radiance','diffuse','specular'} like sbmc.KPCN, 128->(128-4) valid convs."""

    def __init__(self, n_in):
        super().__init__()
        self.diffuse = torch.nn.Conv2d(n_in, 3, 5)
        self.specular = torch.nn.Conv2d(n_in, 3, 5)

    def forward(self, data):
        d = self.diffuse(data["kpcn_diffuse_in"])
        s = self.specular(data["kpcn_specular_in"])
        from support.utils import crop_like
        albedo = crop_like(data["kpcn_albedo"], d)
        return {"radiance": albedo * d + torch.exp(s) - 1.0, "diffuse": d, "specular": s}


class _TinyPathNet(torch.nn.Module):
    def __init__(self, ic, outc):
        super().__init__()
        self.conv = torch.nn.Conv2d(ic, outc, 1)

    def forward(self, samples):
        p = samples["paths"]
        b, s, c, h, w = p.shape
        return torch.relu(self.conv(p.reshape(b * s, c, h, w))).reshape(b, s, -1, h, w)


class _TinyManifLoss(torch.nn.Module):
    """Stand-in for FeatureMSE (a CUDA kernel in the product): same call contract, any differentiable value."""

    def forward(self, p_buffer, ref):
        return (p_buffer.mean(1) - ref).pow(2).mean()


def _make_itf(seed):
    from wcmc_b200 import dropin
    dropin.install()
    from support.interfaces import KPCNInterface
    from support.losses import RelativeMSE
    torch.manual_seed(seed)
    models = {"dncnn": _TinyKPCN(34 + 1 + 3 + 1), "backbone_diffuse": _TinyPathNet(36, 3),
              "backbone_specular": _TinyPathNet(36, 3)}
    optims = {"optim_" + k: torch.optim.Adam(m.parameters(), lr=1e-3) for k, m in models.items()}
    lf = {"l_diffuse": torch.nn.L1Loss(), "l_specular": torch.nn.L1Loss(), "l_recon": torch.nn.L1Loss(),
          "l_test": RelativeMSE(), "l_manif": _TinyManifLoss()}
    itf = KPCNInterface(models, optims, lf, types.SimpleNamespace(model_name="t"), use_llpm_buf=True,
                        manif_learn=True, w_manif=0.1)
    itf.to_train_mode()
    return itf, models


def _flat(models, grad=False):
    return torch.cat([(p.grad if grad else p.data).reshape(-1) for m in models.values() for p in m.parameters()])


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    from wcmc_b200 import ddp
    from wcmc_b200.synth import make_batch
    itf, models = _make_itf(seed=100 + rank)           # replicas start DIFFERENT on purpose
    ddp.broadcast_parameters(models)
    w0 = _flat(models).clone()
    gathered = [torch.empty_like(w0) for _ in range(world)]
    dist.all_gather(gathered, w0)
    assert all(torch.equal(g, gathered[0]) for g in gathered), "broadcast_parameters did not equalise replicas"

    batch = make_batch(batch=2, spp=2, size=24, seed=1234 + rank)   # each rank its own shard
    # local gradients of this shard (no update), with the pairing permutations pinned by the seed
    itf.preprocess(batch)
    torch.manual_seed(7 + rank)
    itf.train_batch(batch, grad_hook_mode=True)
    g_local = _flat(models, grad=True).clone()
    g_all = [torch.empty_like(g_local) for _ in range(world)]
    dist.all_gather(g_all, g_local)
    g_mean = torch.stack(g_all).mean(0)

    # the real step with the hook: gradient seen by clip/Adam must be the mean over ranks
    seen = {}
    sync = ddp.GradAllReduce()

    def hook(ms):
        sync(ms)
        seen["g"] = _flat(ms, grad=True).clone()
    itf.grad_sync = hook
    itf.preprocess(batch)
    torch.manual_seed(7 + rank)
    itf.train_batch(batch)
    torch.testing.assert_close(seen["g"], g_mean, rtol=1e-6, atol=1e-9)
    assert sync.bytes_last == g_mean.numel() * 4
    w1 = _flat(models)
    assert not torch.equal(w1, w0), "no parameter update happened"
    gathered = [torch.empty_like(w1) for _ in range(world)]
    dist.all_gather(gathered, w1)
    assert all(torch.equal(g, gathered[0]) for g in gathered), "replicas diverged after one step"
    # ---- the overlapping exchange: backward cut at the p-buffers, dncnn's all-reduce started asynchronously while the
    # path-embedding networks still back-propagate (KPCNInterface._run_backward + GradAllReduce.early) ----
    itf.grad_sync = None
    itf.preprocess(batch)
    torch.manual_seed(9 + rank)
    itf.train_batch(batch, grad_hook_mode=True)
    g_local2 = _flat(models, grad=True).clone()
    g_all2 = [torch.empty_like(g_local2) for _ in range(world)]
    dist.all_gather(g_all2, g_local2)
    calls = {"early": 0}

    class Rec(ddp.GradAllReduce):
        def early(self, ms):
            calls["early"] += 1
            assert list(ms) == ["dncnn"]
            super().early(ms)

        def __call__(self, ms):
            super().__call__(ms)
            seen["g2"] = _flat(ms, grad=True).clone()
    itf.grad_sync = Rec()
    itf.preprocess(batch)
    torch.manual_seed(9 + rank)
    itf.train_batch(batch)
    assert calls["early"] == 1, "the overlapped path was not taken"
    torch.testing.assert_close(seen["g2"], torch.stack(g_all2).mean(0), rtol=1e-6, atol=1e-9)
    assert itf.grad_sync.bytes_last == g_local2.numel() * 4
    w2 = _flat(models)
    gathered = [torch.empty_like(w2) for _ in range(world)]
    dist.all_gather(gathered, w2)
    assert all(torch.equal(g, gathered[0]) for g in gathered), "replicas diverged after the overlapped step"
    # ---- the same step through the SYMMETRIC-BUFFER host logic of the NVSwitch transport (16-byte slots, two channels,
    # the flag riding in the late exchange, p.grad adopting views of the buffer) with a stand-in for the kernel: a
    # CPU buffer reduced by gloo, same interface as ddp.PeerExchange ----
    class FakeExchange:
        transport, multicast_ptr = "peer", 0

        def __init__(self, n):
            self.capacity = (n + 3) // 4 * 4
            self.buf = torch.full((self.capacity,), float("nan"))     # padding words are never read
            self.calls = []

        def all_reduce_(self, n, scale=1.0, channel=0, offset=0):
            assert n % 4 == 0 and offset == 0 and n <= self.capacity
            self.calls.append((channel, n))
            region = torch.nan_to_num(self.buf[:n])                   # the pad words may hold anything
            dist.all_reduce(region)
            self.buf[:n] = region * scale
            return self.buf[:n]

    class PeerLogic(ddp.GradAllReduce):
        def _use_peer(self, grads):
            return True

        def _exchange(self, channel, n):
            px = self._px.get(channel)
            if px is None or px.capacity < n:
                px = self._px[channel] = FakeExchange(n)
            return px

    itf.grad_sync = None
    itf.preprocess(batch)
    torch.manual_seed(11 + rank)
    itf.train_batch(batch, grad_hook_mode=True)
    g_local3 = _flat(models, grad=True).clone()
    g_all3 = [torch.empty_like(g_local3) for _ in range(world)]
    dist.all_gather(g_all3, g_local3)
    peer = PeerLogic()
    box = {}

    def hook3(ms):
        box["ok"] = peer(ms, ok=torch.tensor(True))
        box["g"] = _flat(ms, grad=True).clone()
    itf.grad_sync = type("Hook", (), {"early": staticmethod(peer.early), "drain": staticmethod(peer.drain), "world": world,
                                      "__call__": lambda self, ms: hook3(ms)})()
    itf.preprocess(batch)
    torch.manual_seed(11 + rank)
    itf.train_batch(batch)
    torch.testing.assert_close(box["g"], torch.stack(g_all3).mean(0), rtol=1e-6, atol=1e-9)
    assert bool(box["ok"]), "all ranks finite -> combined flag true"
    assert [c for c, _ in peer._px[0].calls] == [0] and [c for c, _ in peer._px[1].calls] == [1]
    # every gradient is now a 16-byte-aligned view of its channel's buffer (no copy back)
    for name, m in models.items():
        px = peer._px[0 if name == "dncnn" else 1]
        lo, hi = px.buf.data_ptr(), px.buf.data_ptr() + px.buf.numel() * 4
        for p_ in m.parameters():
            if p_.grad is not None:
                assert lo <= p_.grad.data_ptr() < hi and (p_.grad.data_ptr() - lo) % 16 == 0
    # a rank with a non-finite loss makes every rank skip: the flag element is summed with the gradients
    bad = peer({k: m for k, m in models.items()}, ok=torch.tensor(rank != 1))
    assert not bool(bad)
    w3 = _flat(models)
    gathered = [torch.empty_like(w3) for _ in range(world)]
    dist.all_gather(gathered, w3)
    assert all(torch.equal(g, gathered[0]) for g in gathered), "replicas diverged after the symmetric-buffer step"
    # expected update: Adam on clip(mean gradient)
    torch.save({"w0": w0, "w1": w1, "g": seen["g"]}, os.path.join(out_dir, "rank%d.pt" % rank))
    dist.barrier()
    dist.destroy_process_group()


def test_grad_allreduce_world2_gloo(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0 = torch.load(tmp_path / "rank0.pt")
    r1 = torch.load(tmp_path / "rank1.pt")
    assert torch.equal(r0["w1"], r1["w1"]) and torch.equal(r0["g"], r1["g"])
    # first Adam step moves every parameter by lr * sign(g) (bias-corrected m/sqrt(v) = g/|g|)
    g = r0["g"].clamp(-1, 1)
    nz = g.abs() > 1e-4
    step = (r0["w1"] - r0["w0"])[nz]
    torch.testing.assert_close(step, -1e-3 * torch.sign(g[nz]), rtol=2e-3, atol=1e-7)


def test_grad_allreduce_single_process_is_noop():
    from wcmc_b200 import ddp
    m = {"a": torch.nn.Linear(3, 2)}
    m["a"].weight.grad = torch.ones(2, 3)
    ddp.GradAllReduce()(m)
    assert torch.equal(m["a"].weight.grad, torch.ones(2, 3))
