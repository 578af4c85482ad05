"""GPU parity of the path-disentangling loss kernels (wcmc_fmse_perm_fwd / _bwd) through the
drop-in `support.losses` API: against vectors produced by the REFERENCE's own support/losses.py
(tests/golden/make_golden.py), against the oracle on cropped (strided) views, and -- at
BASELINE.json's full size -- through size-independent properties."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


@pytest.fixture(scope="module")
def losses():
    from wcmc_b200 import dropin, lib
    dropin.install()
    lib.init()
    import support.losses as L
    return L


@pytest.mark.parametrize("tag", ["a", "b"])
@pytest.mark.parametrize("nl", [True, False])
def test_fmse_matches_reference_golden(losses, golden, tag, nl):
    g = golden["fmse_%s_nl%d" % (tag, nl)]
    p = g["p"].cuda().requires_grad_(True)
    torch.manual_seed(g["seed"])          # CPU default generator: same pairing as the reference run
    loss = losses.FeatureMSE(non_local=nl)(p, g["ref"].cuda())
    loss.backward()
    assert rel(loss, g["loss"]) < 1e-5
    assert rel(p.grad, g["grad"]) < 1e-5
    torch.testing.assert_close(p.grad.cpu(), g["grad"], rtol=1e-3, atol=1e-7 * float(g["grad"].abs().max()) + 1e-12)


@pytest.mark.parametrize("tag", ["a", "b"])
def test_grs_matches_reference_golden(losses, golden, tag):
    g = golden["grs_%s" % tag]
    p = g["p"].cuda().requires_grad_(True)
    torch.manual_seed(g["seed"])
    loss = losses.GlobalRelativeSimilarityLoss()(p, g["ref"].cuda())
    loss.backward()
    assert rel(loss, g["loss"]) < 1e-5
    assert rel(p.grad, g["grad"]) < 1e-4


@pytest.mark.parametrize("shape,crop", [((2, 3, 4, 40, 40), 18), ((1, 2, 12, 33, 47), 5), ((3, 8, 3, 24, 24), 0)])
def test_fmse_strided_crop_vs_oracle(losses, oracle, shape, crop):
    """The kernels read the crop_like() view of the p-buffer / target without a copy."""
    from support.utils import crop_like
    b, s, c, h, w = shape
    g = torch.Generator().manual_seed(3)
    p_full = torch.rand(shape, generator=g)
    ref_full = torch.rand(b, 3, h, w, generator=g) * 3 - 0.5       # some negatives: clamp path
    like = torch.empty(b, 3, h - 2 * crop, w - 2 * crop)
    n = s * (h - 2 * crop) * (w - 2 * crop)
    idx_p, idx_b = torch.randperm(n, generator=g), torch.randperm(b * n, generator=g)

    def run(fn, dev):
        pf = p_full.to(dev).requires_grad_(True)
        loss = fn(crop_like(pf, like), crop_like(ref_full.to(dev), like), idx_p.to(dev), idx_b.to(dev))
        loss.backward()
        return loss.detach(), pf.grad

    fmse = losses.FeatureMSE(non_local=True)
    lo, go = run(lambda p, r, ip, ib: fmse(p, r, ip, ib), "cuda")
    lr, gr = run(lambda p, r, ip, ib: oracle.ref.feature_mse(p, r, True, ip, ib), "cpu")
    assert rel(lo, lr) < 1e-5 and rel(go, gr) < 1e-5
    if crop:
        assert float(go[..., :crop, :].abs().max()) == 0.0      # nothing leaks outside the crop


def test_fmse_nonfinite_raises(losses):
    p = torch.rand(1, 2, 3, 8, 8, device="cuda")
    ref = torch.rand(1, 3, 8, 8, device="cuda")
    f = losses.FeatureMSE(non_local=True)
    f(p, ref)
    p[0, 1, 2, 3, 4] = float("inf")
    with pytest.raises(RuntimeError, match="Infinite loss at train time"):
        f(p, ref)
    p[0, 1, 2, 3, 4] = 0.5
    ref[0, 1, 2, 2] = float("nan")
    with pytest.raises(RuntimeError, match="Infinite loss at train time"):
        f(p, ref)


def test_fmse_full_size_properties(losses):
    """B=8, S=8, C=3, 92x92 (541,696 rows): (i) identity pairing gives exactly zero loss and zero
    gradient; (ii) when P reproduces the tone-mapped label (C=3) every displacement vanishes for ANY
    pairing; (iii) the loss is invariant under pi -> pi^-1 (pairs are unordered) and the gradient
    sums to zero over all rows (each pair contributes +d and -d)."""
    b, s, c, h, w = 8, 8, 3, 92, 92
    n = s * h * w
    g = torch.Generator(device="cuda").manual_seed(0)
    p = torch.rand(b, s, c, h, w, device="cuda", generator=g).requires_grad_(True)
    ref = torch.rand(b, 3, h, w, device="cuda", generator=g) * 4
    f = losses.FeatureMSE(non_local=True)
    ident_p, ident_b = torch.arange(n, device="cuda"), torch.arange(b * n, device="cuda")
    loss = f(p, ref, ident_p, ident_b)
    loss.backward()
    assert float(loss) == 0.0 and float(p.grad.abs().max()) == 0.0
    t = (ref.clamp(min=0) / (1 + ref.clamp(min=0))) ** 0.454545
    pt = t.unsqueeze(1).expand(b, s, 3, h, w).contiguous()
    idx_p, idx_b = torch.randperm(n, device="cuda", generator=g), torch.randperm(b * n, device="cuda", generator=g)
    assert float(f(pt, ref, idx_p, idx_b)) < 1e-10
    p.grad = None
    l1 = f(p, ref, idx_p, idx_b)
    l1.backward()
    inv_p, inv_b = torch.empty_like(idx_p), torch.empty_like(idx_b)
    inv_p[idx_p] = torch.arange(n, device="cuda")
    inv_b[idx_b] = torch.arange(b * n, device="cuda")
    l2 = f(p.detach(), ref, inv_p, inv_b)
    assert rel(l2, l1) < 1e-5
    rows = p.grad.permute(0, 1, 3, 4, 2).reshape(-1, c).double()
    assert float(rows.sum(0).abs().max()) < 1e-6 * float(rows.abs().sum(0).max())


def test_fmse_crop_inside_the_function_matches_crop_like(losses):
    """FeatureMSE.forward_cropped(p, ref): the centred crop taken inside the autograd Function (strided kernel reads,
    gradient written into the interior of a zero-filled full-size tensor) == forward(crop_like(p, ref), ref)."""
    g = torch.Generator(device="cuda").manual_seed(4)
    p1 = torch.rand(2, 3, 4, 40, 44, device="cuda", generator=g).requires_grad_(True)
    p2 = p1.detach().clone().requires_grad_(True)
    ref = torch.rand(2, 3, 20, 28, device="cuda", generator=g) * 2
    lf = losses.FeatureMSE(non_local=True)
    n1, n2 = 3 * 20 * 28, 2 * 3 * 20 * 28
    ip, ib = torch.randperm(n1, generator=torch.Generator().manual_seed(1)).cuda(), torch.randperm(n2, generator=torch.Generator().manual_seed(2)).cuda()
    a = lf.forward_cropped(p1, ref, ip, ib)
    from support.utils import crop_like
    b = lf(crop_like(p2, ref), ref, ip, ib)
    assert torch.equal(a, b)
    (a * 3.0).backward()
    (b * 3.0).backward()
    assert torch.equal(p1.grad, p2.grad) and float(p1.grad[..., :10, :].abs().sum()) == 0.0
    # same size: falls back to the plain call
    q = torch.rand(2, 3, 4, 20, 28, device="cuda", generator=g).requires_grad_(True)
    assert torch.equal(lf.forward_cropped(q, ref, ip, ib), lf(q, ref, ip, ib))


def test_absmax_scale_kernel(losses):
    from wcmc_b200 import lib, ops
    g = torch.Generator(device="cuda").manual_seed(6)
    for n in (1, 5, 1000, 203136 * 3 + 1, 12582912):
        x = torch.randn(n, device="cuda", generator=g) * 3e-4
        out = lib.absmax_scale(x, 256.0)
        amax = float(x.abs().max())
        assert abs(float(out[0]) - 256.0 / amax) <= 1e-6 * 256.0 / amax and abs(float(out[1]) - amax / 256.0) <= 1e-6 * amax
    s, inv = ops.grad_scale(torch.zeros(7, device="cuda"))
    assert torch.isfinite(s).all() and float(s * inv) == pytest.approx(1.0, rel=1e-5)
