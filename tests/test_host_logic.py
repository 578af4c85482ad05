"""CPU: host-side logic that needs no GPU -- stream fork/join is a no-op without CUDA, the fused optimiser's
eligibility rules, the synthetic batch contract, crop_like, and the all-pairs oracle against brute force."""
import importlib.util
import os

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_streams_are_noop_without_cuda():
    from wcmc_b200 import streams
    if torch.cuda.is_available():
        return
    with streams.fork("diffuse"):
        x = torch.ones(3) * 2
    streams.join()
    assert float(x.sum()) == 6.0 and not streams._open


def test_fused_adam_eligibility():
    from wcmc_b200 import optim as wopt
    p = [torch.nn.Parameter(torch.zeros(4))]
    assert not wopt.supported(torch.optim.Adam(p, lr=1e-3))                    # CPU parameters
    assert not wopt.supported(torch.optim.SGD(p, lr=1e-3))
    assert not wopt.supported(torch.optim.AdamW(p, lr=1e-3))
    if torch.cuda.is_available():
        q = [torch.nn.Parameter(torch.zeros(4, device="cuda"))]
        assert wopt.supported(torch.optim.Adam(q, lr=1e-3))
        assert not wopt.supported(torch.optim.Adam(q, lr=1e-3, amsgrad=True))
        assert not wopt.supported(torch.optim.Adam(q, lr=1e-3, weight_decay=0.1))


def test_synthetic_batch_contract():
    """Tensor contract of DenoiseDataset.__getitem__ (/root/reference/support/datasets.py:1078-1126)."""
    from wcmc_b200.synth import make_batch
    b = make_batch(batch=2, spp=3, size=20, seed=1)
    assert tuple(b["kpcn_diffuse_in"].shape) == (2, 35, 20, 20) and tuple(b["paths"].shape) == (2, 3, 36, 20, 20)
    assert torch.equal(b["kpcn_diffuse_in"][:, :3], b["kpcn_diffuse_buffer"])
    assert float(b["kpcn_albedo"].min()) >= 0.00316
    want = b["kpcn_albedo"] * b["target_diffuse"] + torch.exp(b["target_specular"]) - 1.0
    torch.testing.assert_close(b["target_total"], want)
    v = make_batch(batch=1, size=16, paths=False)
    assert tuple(v["kpcn_specular_in"].shape) == (1, 34, 16, 16) and "paths" not in v
    again = make_batch(batch=2, spp=3, size=20, seed=1)
    assert all(torch.equal(b[k], again[k]) for k in b)


def test_crop_like_matches_reference_rule():
    """support/utils.py:24-42: crop = max(delta // 2, 0), crop2 = delta - crop (128 -> 92 = [18:110])."""
    from wcmc_b200 import dropin
    dropin.install()
    from support.utils import crop_like
    src = torch.arange(128 * 128.0).reshape(1, 1, 128, 128)
    out = crop_like(src, torch.empty(1, 1, 92, 92))
    assert torch.equal(out, src[..., 18:110, 18:110])
    odd = crop_like(torch.arange(49.0).reshape(1, 1, 7, 7), torch.empty(1, 1, 4, 4))
    assert torch.equal(odd, torch.arange(49.0).reshape(1, 1, 7, 7)[..., 1:5, 1:5])


def test_allpairs_oracle_vs_brute_force():
    spec = importlib.util.spec_from_file_location("oracle_allpairs_ref", os.path.join(ROOT, "oracle", "allpairs_ref.py"))
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    g = torch.Generator().manual_seed(0)
    p, r = torch.rand(97, 5, generator=g).double(), (torch.rand(97, 3, generator=g) * 3).double()
    t = ref.tonemap_gamma(r)
    e = 0.5 * ((p[:, None] - p[None]) ** 2).sum(-1) - 0.5 * ((t[:, None] - t[None]) ** 2).sum(-1)
    off = ~torch.eye(97, dtype=torch.bool)
    torch.testing.assert_close(ref.allpairs_loss(p, r, chunk=10), (0.5 * e[off] ** 2).sum() / 97 ** 2)
    keep = off & (0.5 * ((t[:, None] - t[None]) ** 2).sum(-1) < 0.05)
    torch.testing.assert_close(ref.allpairs_loss(p, r, tau=0.05, chunk=33), (0.5 * e[keep] ** 2).sum() / 97 ** 2)
    x = 2.0 * torch.cat([e[off], -e[off], torch.zeros(1, dtype=torch.float64)])
    want = (torch.logsumexp(x, 0) - torch.log(torch.tensor(1.0 + 2 * int(off.sum())))) / 2 ** 0.5
    torch.testing.assert_close(ref.allpairs_loss(p, r, mode="lse", chunk=16), want)
