"""GPU parity of the all-pairs loss kernel (K11, wcmc_fmse_allpairs_fwd; an extension, see
wcmc_b200/allpairs.py) against oracle/allpairs_ref.py (chunked fp64 PyTorch), plus size-independent
properties at BASELINE.json configs[4] sizes."""
import importlib.util
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def ref():
    spec = importlib.util.spec_from_file_location("oracle_allpairs_ref", os.path.join(ROOT, "oracle", "allpairs_ref.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.fixture(scope="module")
def ap():
    from wcmc_b200 import allpairs, lib
    lib.init()
    return allpairs


def rows(n, d, seed):
    g = torch.Generator().manual_seed(seed)
    p = torch.rand(n, d, generator=g) * 1.5
    r = torch.rand(n, 3, generator=g) * 4 - 0.3      # some negatives: clamp path of the tone map
    return p, r


@pytest.mark.parametrize("n,d", [(2, 3), (130, 3), (300, 16), (1000, 32), (4153, 16), (2048, 37)])
@pytest.mark.parametrize("mode,tau", [("mse", None), ("lse", None), ("mse", 0.05), ("lse", 0.05)])
def test_allpairs_vs_oracle(ap, ref, n, d, mode, tau):
    p, r = rows(n, d, n + d)
    want = ref.allpairs_loss(p, r, mode=mode, alpha=2.0, tau=tau)
    got, kept = ap.allpairs_loss(p.cuda(), r.cuda(), mode=mode, alpha=2.0, tau=tau)
    assert abs(float(got) - float(want)) <= 2e-5 * abs(float(want)) + 1e-7, (float(got), float(want))
    if tau is None:
        assert float(kept) == n * (n - 1)


def test_allpairs_gradient_closed_form(ap, ref):
    p, r = rows(257, 12, 5)
    pc = p.clone().double().requires_grad_(True)
    ref.allpairs_loss(pc, r).backward()
    pg = p.cuda().requires_grad_(True)
    loss, _ = ap.allpairs_loss(pg, r.cuda())
    (3.0 * loss).backward()
    err = (pg.grad.cpu().double() - 3.0 * pc.grad).norm() / pc.grad.norm() / 3.0
    assert float(err) < 1e-5


def test_allpairs_nonfinite_raises(ap):
    p, r = rows(200, 8, 1)
    p[17, 3] = float("nan")
    with pytest.raises(RuntimeError, match="Infinite loss at train time"):
        ap.allpairs_loss(p.cuda(), r.cuda())


def test_allpairs_properties_config5_size(ap, ref):
    """N = 65536, D = 16 (2.1e9 unordered pairs): (i) invariant under a row permutation; (ii) zero when
    the embedding reproduces the tone-mapped label; (iii) a threshold that keeps everything equals the
    unmasked loss and keeps N(N-1) pairs; (iv) the permutation-paired reference loss is an unbiased
    estimate of the all-pairs mse: averaged over 20 random pairings it agrees within a few per cent."""
    n, d = 65536, 16
    g = torch.Generator(device="cuda").manual_seed(0)
    p = torch.rand(n, d, device="cuda", generator=g)
    r = torch.rand(n, 3, device="cuda", generator=g) * 3
    base, kept = ap.allpairs_loss(p, r)
    assert float(kept) == float(n) * (n - 1)
    perm = torch.randperm(n, device="cuda", generator=g)
    again, _ = ap.allpairs_loss(p[perm].contiguous(), r[perm].contiguous())
    assert abs(float(again) - float(base)) < 1e-5 * float(base)
    t = (r.clamp(min=0) / (1 + r.clamp(min=0))) ** 0.454545
    zero, _ = ap.allpairs_loss(t.contiguous(), r)
    assert float(zero) < 1e-9
    big, kept2 = ap.allpairs_loss(p, r, tau=1e6)
    assert abs(float(big) - float(base)) < 1e-5 * float(base) and float(kept2) == float(kept)
    est = 0.0
    for _ in range(20):
        j = torch.randperm(n, device="cuda", generator=g)
        e = 0.5 * (p - p[j]).pow(2).sum(1) - 0.5 * (t - t[j]).pow(2).sum(1)
        est += float(0.5 * e.pow(2).mean()) / 20
    assert abs(est - float(base)) < 0.03 * float(base)
