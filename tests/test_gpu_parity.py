"""GPU parity tests: the CUDA path (through the C ABI / drop-in modules) against the oracle on the
same seeded inputs and weights.  Tolerances are north_star's: relative L2 <= 1e-3 for denoised
images and losses, <= 1e-2 for gradients on the bf16 tensor-core path.
The oracle is plain PyTorch; for full-size cases it is evaluated on the GPU in fp32 (TF32 off)."""
import types

import pytest
import torch
import torch.nn.functional as F

from tests._record import record
from wcmc_b200.synth import make_batch

pytestmark = pytest.mark.gpu

TOL_IMG = 1e-3
TOL_GRAD = 1e-2
# The raw manifold term on the SMALL golden cases: 2 patches x 2 spp x 40^2 crop to 4 x 4 pixels = 64 rows per loss
# call, so the ~7e-4 fp16 storage error of the p-buffer does not average out over the mean (measured 1.1e-3 /
# 2.0e-3, profiles/parity_r02.txt).  At the north-star size (541,696 rows) the same term sits at 2.4e-4 and is held to
# TOL_IMG = 1e-3 like every other loss (test_full_size_wcmc_step_vs_oracle).
TOL_MANIF = 5e-3


def rel(a, b):
    a, b = a.double(), b.double()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


@pytest.fixture(scope="module")
def backend():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    from wcmc_b200 import dropin, lib
    dropin.install()
    lib.init()
    import sbmc
    import support.interfaces as itf
    import support.losses as losses
    import support.networks as networks
    return types.SimpleNamespace(lib=lib, sbmc=sbmc, KPCN=sbmc.KPCN, PathNet=networks.PathNet, losses=losses,
                                 itf=itf)


def to_cuda(batch):
    return {k: v.cuda() for k, v in batch.items()}


# ---------------------------------------------------------------------------------------------
# kernels through the C ABI
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("act_dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("k,cin,cout,same", [(5, 100, 100, 0), (5, 39, 100, 0), (5, 100, 441, 0), (3, 64, 64, 1),
                                             (3, 384, 128, 1), (1, 36, 64, 0), (1, 128, 3, 0)])
def test_conv_fwd_dgrad_wgrad_vs_fp64(backend, k, cin, cout, same, act_dtype):
    """Activations / weights in `act_dtype`, gradients in bf16 (the tensor cores mix the formats)."""
    lib = backend.lib
    g = torch.Generator(device="cuda").manual_seed(k * 1000 + cin)
    n, h, w = 2, 44, 36
    pad = k // 2 if same else 0
    x = torch.randn(n, cin, h, w, device="cuda", generator=g).to(act_dtype).float()
    wt = (torch.randn(cout, cin, k, k, device="cuda", generator=g) / (cin * k * k) ** 0.5).to(act_dtype).float()
    b = torch.randn(cout, device="cuda", generator=g)
    xd = x.double().requires_grad_(True)
    wd_ = wt.double().requires_grad_(True)
    ref = F.conv2d(xd, wd_, b.double(), padding=pad)
    dy = torch.randn(ref.shape, device="cuda", generator=g).to(act_dtype).float()
    gx, gw = torch.autograd.grad(ref, (xd, wd_), dy.double())
    wf, wdg, bp = lib.pack_weights(wt, b, want_bias=True, dtype=act_dtype)
    xn = lib.nchw_to_nhwc(x, dtype=act_dtype)
    y = lib.conv2d(xn, wf, bp, k, pad, act=0, out_dtype=torch.float32)
    assert rel(y[..., :cout].permute(0, 3, 1, 2), ref) < 2e-5
    dyn = lib.nchw_to_nhwc(dy.to(act_dtype).float(), dtype=act_dtype)
    dx = lib.conv2d(dyn, wdg, None, k, k - 1 - pad, act=0, out_dtype=torch.float32)
    assert rel(dx[..., :cin].permute(0, 3, 1, 2), gx) < 2e-5
    dw = lib.conv2d_wgrad(xn, dyn, cout, cin, k, pad, lib.pad16(cin), lib.pad16(cout))
    assert rel(dw, gw) < 2e-5
    db = lib.bias_grad(dyn, cout)
    assert rel(db, dy.double().sum((0, 2, 3))) < 1e-5


@pytest.mark.parametrize("act_dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("pack", [1, 0])
def test_grouped_wgrad_vs_fp64(backend, act_dtype, pack):
    """wcmc_conv2d_wgrad_group: the weight gradients of a whole backward pass in ONE launch (K-split teams dealt out
    over the SMs, whole-kernel-row tap groups, 100-channel rows packed at a TMEM column stride of 100) against fp64
    autograd -- every layer shape of the hot path in one group (KPCN first / mid / last 5x5 valid layers, U-Net 3x3
    same-padded layers incl. channel slices of a wider tensor, a 1x1 layer), ragged map sizes, `accumulate` and
    `scale`; the un-packed column layout (knob) must give the same numbers."""
    lib = backend.lib
    rt = lib.load()
    assert rt.wcmc_tuning_set(b"wgrad_group_pack", pack) == 0
    try:
        g = torch.Generator(device="cuda").manual_seed(11)
        cases = [  # n, h, w, cin, cout, k, pad
            (2, 37, 29, 39, 100, 5, 0), (2, 33, 25, 100, 100, 5, 0), (1, 29, 21, 100, 441, 5, 0),
            (2, 24, 40, 64, 64, 3, 1), (2, 12, 20, 128, 256, 3, 1), (2, 12, 20, 256, 256, 3, 1),
            (2, 24, 40, 384, 128, 3, 1), (2, 24, 40, 192, 64, 3, 1), (3, 9, 7, 36, 64, 1, 0), (1, 8, 8, 100, 100, 5, 2)]
        scale = torch.full((1,), 0.25, device="cuda")
        want, got, keep = [], [], []
        for i, (n, h, w, cin, cout, k, pad) in enumerate(cases):
            x = torch.randn(n, cin, h, w, device="cuda", generator=g).to(act_dtype).float()
            ho, wo = h + 2 * pad - k + 1, w + 2 * pad - k + 1
            dy = torch.randn(n, cout, ho, wo, device="cuda", generator=g).to(act_dtype).float()
            gw = torch.nn.grad.conv2d_weight(x.double(), (cout, cin, k, k), dy.double(), padding=pad)
            # operands live in channel slices of wider NHWC tensors for some layers (U-Net concatenation buffers)
            xoff, doff = (8, 16) if i % 3 == 1 else (0, 0)
            cin_p, cout_p = lib.pad16(cin), lib.pad16(cout)
            xw = torch.randn(n, h, w, cin_p + xoff + 8, device="cuda", generator=g).to(act_dtype)
            dw_ = torch.randn(n, ho, wo, cout_p + doff + 8, device="cuda", generator=g).to(act_dtype)
            lib.nchw_to_nhwc(x, dst=xw, dst_coff=xoff, c_fill=cin_p, dtype=act_dtype)
            lib.nchw_to_nhwc(dy, dst=dw_, dst_coff=doff, c_fill=cout_p, dtype=act_dtype)
            acc = i % 4 == 2
            out = torch.randn(cout, cin, k, k, device="cuda", generator=g) if acc else None
            base = out.double().clone() if acc else 0.0
            sc = scale if i % 2 == 0 else None
            r = lib.conv2d_wgrad(xw, dw_, cout, cin, k, pad, cin_p, cout_p, x_coff=xoff, dy_coff=doff, out=out,
                                 accumulate=acc, scale=sc, defer=True)
            want.append(base + gw * (0.25 if sc is not None else 1.0))
            got.append(r)
            keep.append((xw, dw_))
        lib.wgrad_flush()
        torch.cuda.synchronize()
        for c, a, b in zip(cases, got, want):
            assert rel(a, b) < 2e-5, (c, rel(a, b))
    finally:
        rt.wcmc_tuning_set(b"wgrad_group_pack", 1)


def test_grouped_wgrad_matches_per_layer_launches_at_full_size(backend):
    """The grouped launch against the round-1 per-layer kernel on the nine layers of a KPCN branch at the north-star
    size (B = 8, 128^2): same gradients (fp32 summation order differs), and the plan uses every SM."""
    lib = backend.lib
    rt = lib.load()
    from wcmc_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(3)
    dt = ops.ACT_DTYPE
    shapes = [(39, 100, 128)] + [(100, 100, 128 - 4 * i) for i in range(1, 8)] + [(100, 441, 96)]
    layers = []
    for cin, cout, h in shapes:
        x = (torch.randn(8, h, h, lib.pad16(cin), device="cuda", generator=g) * 0.5).to(dt)
        x[..., cin:] = 0
        dy = (torch.randn(8, h - 4, h - 4, lib.pad16(cout), device="cuda", generator=g) * 0.1).to(dt)
        dy[..., cout:] = 0
        layers.append((x, dy, cin, cout))
    outs = {}
    for mode in (1, 0):
        assert rt.wcmc_tuning_set(b"wgrad_group", mode) == 0
        try:
            res = [lib.conv2d_wgrad(x, dy, cout, cin, 5, 0, lib.pad16(cin), lib.pad16(cout), defer=True)
                   for x, dy, cin, cout in layers]
            lib.wgrad_flush()
            torch.cuda.synchronize()
            outs[mode] = res
        finally:
            rt.wcmc_tuning_set(b"wgrad_group", 1)
    for a, b in zip(outs[1], outs[0]):
        assert rel(a, b) < 1e-4      # two fp32 summation orders over 8 x 124^2 pixels (measured 3e-5)
    plan, launches = lib.wgrad_group_plan([(8, h, h, cin, cout, 5, 0) for cin, cout, h in shapes])
    # three chains (first layer / the seven identical 100->100 layers / the 441-channel layer) share the 148 SMs; a
    # layer's partial sums come from the 2-6 teams whose piece of the chain's tile line touches it
    assert launches == 1 and all(1 <= p[0] <= 8 for p in plan)
    assert len({p[1] for p in plan[1:8]}) == 1 and 140 <= plan[0][1] + plan[1][1] + plan[8][1] <= 148


def test_batched_pack_matches_single_layer_pack(backend):
    """wcmc_pack_weights_batch (row-staged through shared memory) against the element-wise single-layer kernel:
    forward and data-gradient operand layouts and the padded bias, bit for bit, on every layer shape of the path."""
    lib = backend.lib
    g = torch.Generator(device="cuda").manual_seed(5)
    shapes = [(100, 39, 5), (100, 100, 5), (441, 100, 5), (64, 36, 1), (64, 64, 3), (128, 384, 3), (256, 256, 3), (3, 128, 1),
              (64, 192, 3)]
    for dtype in (torch.float16, torch.bfloat16):
        specs = []
        for cout, cin, k in shapes:
            w = torch.randn(cout, cin, k, k, device="cuda", generator=g)
            b = torch.randn(cout, device="cuda", generator=g)
            specs.append((w, b, lib.pad16(cout), lib.pad16(cin)))
        got = lib.pack_weights_batch(specs, dtype=dtype, dgrad=True)
        for (w, b, cout_p, cin_p), (f, d, bp) in zip(specs, got):
            f1, d1, b1 = lib.pack_weights(w, b, cout_p, cin_p, want_bias=True, dtype=dtype)
            assert torch.equal(f, f1) and torch.equal(d, d1) and torch.equal(bp, b1), tuple(w.shape)
        fwd_only = lib.pack_weights_batch(specs[:2], dtype=dtype, dgrad=False)
        assert fwd_only[0][1] is None and torch.equal(fwd_only[1][0], got[1][0])


def test_device_permutation_is_a_fresh_bijection(backend):
    """wcmc_random_permutation (keyed Feistel network + cycle walking): a bijection of [0, n) for awkward n, a new one
    on every launch (the launch advances its own device counter, so CUDA-graph replays differ too), no visible
    structure (fixed points ~ 1, displacement spread over the whole range)."""
    lib = backend.lib
    st = torch.tensor([12345, 0], dtype=torch.int64, device="cuda")
    for n in (1, 2, 7, 1000, 67712, 541696):
        a = lib.random_permutation(n, st, salt=n)
        b = lib.random_permutation(n, st, salt=n)
        ar = torch.arange(n, device="cuda")
        assert torch.equal(torch.sort(a).values, ar) and torch.equal(torch.sort(b).values, ar), n
        if n >= 1000:
            assert not torch.equal(a, b)
            assert int((a == ar).sum()) < 12 and int((a == b).sum()) < 12
            disp = (a - ar).abs().float().mean() / n
            assert 0.28 < float(disp) < 0.39        # E|i - pi(i)| / n = 1/3 for a uniform permutation
    assert int(st[0]) == 12345 + 12 and int(st[1]) == 0
    graph = torch.cuda.CUDAGraph()
    out = torch.empty(5000, dtype=torch.int64, device="cuda")
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        lib.random_permutation(5000, st, out=out)
    torch.cuda.current_stream().wait_stream(s)
    with torch.cuda.graph(graph):
        lib.random_permutation(5000, st, out=out)
    graph.replay()
    first = out.clone()
    graph.replay()
    assert not torch.equal(first, out) and torch.equal(torch.sort(out).values, torch.arange(5000, device="cuda"))


def test_step_glue_kernels_vs_torch(backend):
    """K9 / K12 against the reference's torch expressions (support/interfaces.py:165-180; nn.L1Loss;
    support/losses.py:255-264; radiance recombination of sbmc.KPCN.forward), forward and backward."""
    from wcmc_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(21)
    b, s, c, h, w, cin = 3, 4, 6, 20, 28, 35
    kin = torch.randn(b, cin, h, w, device="cuda", generator=g)
    for c0, cr in ((0, c), (0, c // 2), (2, 3)):
        p = torch.rand(b, s, c, h, w, device="cuda", generator=g).requires_grad_(True)
        q = p.detach().clone().requires_grad_(True)
        out = ops.PBufferConcatFn.apply(kin, p, c0, cr)
        sl = q[:, :, c0:c0 + cr]
        want = torch.cat([kin, sl.mean(1), sl.var(1).mean(1, keepdim=True).detach() / s], 1)
        assert out.shape == want.shape and rel(out, want) < 1e-6
        wgt = torch.randn_like(out)
        (out * wgt).sum().backward()
        (want * wgt).sum().backward()
        assert rel(p.grad, q.grad) < 1e-6 and float(p.grad[:, :, :c0].abs().sum()) == 0.0
    # recombination + image losses on centred crops of full-size tensors
    hh, ww = 56, 64
    alb = torch.rand(b, 3, hh, ww, device="cuda", generator=g) + 0.01
    tgt = [torch.rand(b, 3, hh, ww, device="cuda", generator=g) * 2 for _ in range(3)]
    rd = torch.rand(b, 3, h, w, device="cuda", generator=g).requires_grad_(True)
    rs = torch.rand(b, 3, h, w, device="cuda", generator=g).requires_grad_(True)
    rd2, rs2 = rd.detach().clone().requires_grad_(True), rs.detach().clone().requires_grad_(True)
    crop = backend.sbmc.crop_like
    rad = ops.RecombineFn.apply(alb, rd, rs)
    rad2 = crop(alb, rd2) * rd2 + torch.exp(rs2) - 1.0
    assert rel(rad, rad2) < 1e-6
    wg = torch.randn_like(rad)
    (rad * wg).sum().backward()
    (rad2 * wg).sum().backward()
    assert rel(rd.grad, rd2.grad) < 1e-6 and rel(rs.grad, rs2.grad) < 1e-6
    rd.grad = rs.grad = rd2.grad = rs2.grad = None
    l_d, l_s, l_t, rmse = ops.ImageLossesFn.apply(rd, rs, rad.detach(), tgt[0], tgt[1], tgt[2], 1e-2)
    l1 = torch.nn.L1Loss()
    want = [l1(rd2, crop(tgt[0], rd2)), l1(rs2, crop(tgt[1], rs2)), l1(rad2.detach(), crop(tgt[2], rad2)),
            backend.losses.RelativeMSE()(rad2.detach(), crop(tgt[2], rad2).contiguous())]
    for a, b_ in zip((l_d, l_s, l_t, rmse), want):
        assert rel(a, b_) < 1e-5
    relm = 0.5 * torch.mean((rad2.detach() - crop(tgt[2], rad2)) ** 2 / (crop(tgt[2], rad2) ** 2 + 1e-2))
    assert rel(rmse, relm) < 1e-5
    (l_d * 0.7 + l_s * 1.3).backward()
    (want[0] * 0.7 + want[1] * 1.3).backward()
    assert rel(rd.grad, rd2.grad) < 1e-6 and rel(rs.grad, rs2.grad) < 1e-6
    # the sums are deterministic (fixed-order partials): a second launch gives the same bits
    again = ops.ImageLossesFn.apply(rd.detach(), rs.detach(), rad.detach(), tgt[0], tgt[1], tgt[2], 1e-2)
    assert all(torch.equal(a, b_) for a, b_ in zip((l_d, l_s, l_t, rmse), again))


@pytest.mark.parametrize("n,h,w", [(2, 44, 36), (1, 37, 61), (8, 96, 96)])
@pytest.mark.parametrize("flags", [0, 1 << 21, (1 << 20) | (1 << 4), (1 << 20) | (2 << 4), (1 << 21) | (2 << 4)])
def test_fused_last_conv_kernel_apply_vs_separate_launches(backend, oracle, n, h, w, flags):
    """wcmc_conv2d_kernel_apply (last 5x5 layer -> running softmax over the four n tiles -> 21x21 gather,
    logits never in HBM) against the two separate launches (fp32 logits + wcmc_kernel_apply_fwd) and against the
    oracle's softmax + kernel_weighting on the same logits; single-CTA and CTA-pair launches, one and two M tiles per
    region (the two warp halves then split columns / tiles differently), ragged maps."""
    lib = backend.lib
    from wcmc_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(n * 100 + h)
    dt = ops.ACT_DTYPE
    x = lib.nchw_to_nhwc(torch.randn(n, 100, h, w, device="cuda", generator=g).relu(), dtype=dt)
    wt = torch.randn(441, 100, 5, 5, device="cuda", generator=g) * 0.02
    b = torch.randn(441, device="cuda", generator=g) * 0.5
    wf, _, bp = lib.pack_weights(wt, b, want_bias=True, dtype=dt)
    data = torch.rand(n, 3, h - 4, w - 4, device="cuda", generator=g) * 3
    logits = lib.conv2d(x, wf, bp, 5, 0, act=0, out_dtype=torch.float32)
    want, _ = lib.kernel_apply_fwd(logits, data, 21, want_stats=False)
    got = lib.conv2d_kernel_apply(x, wf, bp, data, 5, 0, 21, flags=flags)
    assert got.shape == want.shape and rel(got, want) < 2e-5, rel(got, want)
    if h <= 61:
        k = torch.softmax(logits[..., :441].permute(0, 3, 1, 2).double(), 1).view(n, 21, 21, h - 4, w - 4)
        ref, _ = oracle.modules.kernel_weighting(data.double(), k)
        assert rel(got, ref) < 2e-5


@pytest.mark.parametrize("n,h,w,cin,cout,k,pad", [(8, 128, 128, 64, 64, 3, 1), (2, 37, 53, 64, 64, 3, 1),
                                                  (8, 64, 64, 64, 128, 3, 1), (8, 64, 64, 128, 64, 3, 1), (3, 40, 40, 39, 100, 5, 0)])
def test_conv_resident_weights_is_bit_identical(backend, n, h, w, cin, cout, k, pad):
    """Layers whose whole weight tensor fits the shared-memory ring load it once per CTA (`b_resident`) instead of
    re-streaming it for every work item: same MMAs in the same order, so the output must not change by a bit
    (knob conv_resident = 0 restores the streaming ring)."""
    lib = backend.lib
    rt = lib.load()
    from wcmc_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(cin + cout)
    dt = ops.ACT_DTYPE
    x = lib.nchw_to_nhwc(torch.randn(n, cin, h, w, device="cuda", generator=g), dtype=dt)
    wt = torch.randn(cout, cin, k, k, device="cuda", generator=g) * 0.05
    b = torch.randn(cout, device="cuda", generator=g)
    wf, wd, bp = lib.pack_weights(wt, b, want_bias=True, dtype=dt)
    outs = {}
    for res in (1, 0):
        assert rt.wcmc_tuning_set(b"conv_resident", res) == 0
        try:
            y = lib.conv2d(x, wf, bp, k, pad, act=1, flags=1 << 21)          # single-CTA launch (residency needs it)
            dx = lib.conv2d(y, wd, None, k, k - 1 - pad, act=0, mask=x, flags=1 << 21)
            outs[res] = (y, dx)
        finally:
            rt.wcmc_tuning_set(b"conv_resident", 1)
    torch.cuda.synchronize()
    assert torch.equal(outs[1][0], outs[0][0]) and torch.equal(outs[1][1], outs[0][1])
    ref = torch.relu(F.conv2d(lib.nhwc_to_nchw(x, cin).double(), wt.to(dt).double(), b.double(), padding=pad))
    assert rel(lib.nhwc_to_nchw(outs[1][0], cout), ref) < 2e-3


@pytest.mark.parametrize("mt,nt", [(1, 0), (2, 0), (2, 64), (1, 48)])
def test_conv_tilings_agree(backend, mt, nt):
    lib = backend.lib
    g = torch.Generator(device="cuda").manual_seed(7)
    x = torch.randn(3, 48, 37, 29, device="cuda", generator=g)
    wt = torch.randn(96, 48, 3, 3, device="cuda", generator=g) * 0.05
    wf, _ = lib.pack_weights(wt)
    xn = lib.nchw_to_nhwc(x)
    base = lib.conv2d(xn, wf, None, 3, 1, act=1, out_dtype=torch.float32)
    alt = lib.conv2d(xn, wf, None, 3, 1, act=1, out_dtype=torch.float32, flags=(mt << 4) | (nt << 8))
    assert torch.equal(base, alt)


@pytest.mark.parametrize("n,h,w,cin,cout,k,pad,f32", [(1, 44, 44, 100, 100, 5, 0, False),    # 9 regions: a dead second region
                                                         (1, 16, 16, 64, 64, 3, 1, False),      # a single region
                                                         (4, 72, 72, 100, 441, 5, 0, True),     # 4 n tiles, fp32 logits
                                                         (8, 64, 64, 192, 64, 3, 1, False)])
def test_conv_pair_launch_is_bit_identical(backend, n, h, w, cin, cout, k, pad, f32):
    """CTA-pair launches (cluster of 2, tcgen05 cta_group::2, flags bit 20) against single-CTA launches (bit 21), with
    a kernel row of taps per weight stage and with one tap per stage (bit 22): the same MMAs per output element in the
    same K order, so forward, data gradient (+ fused ReLU mask) are bit-identical and the fused bias gradient agrees
    up to the order of its atomics (the product path picks the launch shape per layer; tools/pair_check.py times it)."""
    lib = backend.lib
    PAIR, SINGLE, TAPS = 1 << 20, 1 << 21, 1 << 22
    g = torch.Generator(device="cuda").manual_seed(100 * k + cin)
    dt = _act_dtype()
    ho, wo = h + 2 * pad - k + 1, w + 2 * pad - k + 1
    x = lib.nchw_to_nhwc(torch.randn(n, cin, h, w, device="cuda", generator=g), dtype=dt)
    wt = torch.randn(cout, cin, k, k, device="cuda", generator=g) * 0.03
    wf, wd, bp = lib.pack_weights(wt, torch.randn(cout, device="cuda", generator=g), want_bias=True, dtype=dt)
    dy = lib.nchw_to_nhwc(torch.randn(n, cout, ho, wo, device="cuda", generator=g), dtype=dt)
    od = torch.float32 if f32 else None
    base = lib.conv2d(x, wf, bp, k, pad, act=0 if f32 else 1, out_dtype=od, flags=SINGLE | TAPS)
    cs0 = torch.zeros(lib.pad16(cin), device="cuda")
    dbase = lib.conv2d(dy, wd, None, k, k - 1 - pad, act=0, mask=x, flags=SINGLE | TAPS, colsum=cs0)
    for flags in (SINGLE, PAIR, PAIR | TAPS):
        assert torch.equal(lib.conv2d(x, wf, bp, k, pad, act=0 if f32 else 1, out_dtype=od, flags=flags), base), flags
        cs = torch.zeros(lib.pad16(cin), device="cuda")
        assert torch.equal(lib.conv2d(dy, wd, None, k, k - 1 - pad, act=0, mask=x, flags=flags, colsum=cs), dbase), flags
        assert rel(cs, cs0) < 1e-4, flags


@pytest.mark.parametrize("k,c,shape", [(21, 3, (2, 37, 45)), (21, 3, (8, 92, 92)), (5, 1, (1, 9, 70)), (3, 4, (2, 8, 32))])
def test_kernel_apply_vs_oracle(backend, oracle, k, c, shape):
    lib = backend.lib
    n, h, w = shape
    g = torch.Generator(device="cuda").manual_seed(11)
    taps = k * k
    cs = (taps + 7) // 8 * 8
    logits = torch.randn(n, h, w, cs, device="cuda", generator=g) * 3
    data = torch.rand(n, c, h, w, device="cuda", generator=g) * 4
    gout = torch.randn(n, c, h, w, device="cuda", generator=g)
    z = logits[..., :taps].permute(0, 3, 1, 2).contiguous().double().requires_grad_(True)
    ref, _ = oracle.modules.KernelApply()(data.double(), z)
    (gz,) = torch.autograd.grad(ref, z, gout.double())
    out, stats = lib.kernel_apply_fwd(logits, data, k)
    assert rel(out, ref) < 1e-5
    dl = lib.kernel_apply_bwd(logits, data, out, stats, gout, k, dtype=torch.float32)
    assert rel(dl[..., :taps], gz.permute(0, 2, 3, 1)) < 1e-5
    assert float(dl[..., taps:].abs().max()) == 0.0 if cs > taps else True
    dlb = lib.kernel_apply_bwd(logits, data, out, stats, gout, k, dtype=torch.bfloat16)
    assert rel(dlb[..., :taps].float(), gz.permute(0, 2, 3, 1)) < 4e-3
    # property: a softmax-weighted gather of a constant image is that constant inside, and never
    # exceeds the data range anywhere (zero padding only darkens)
    ones = torch.full_like(data, 2.5)
    o2, _ = lib.kernel_apply_fwd(logits, ones, k)
    r = k // 2
    if h > 2 * r and w > 2 * r:
        assert torch.allclose(o2[..., r:h - r, r:w - r], ones[..., r:h - r, r:w - r], rtol=1e-5)
    assert float(o2.max()) <= 2.5 * (1 + 1e-5) and float(o2.min()) >= 0.0


def test_glue_kernels_vs_torch(backend):
    lib = backend.lib
    g = torch.Generator(device="cuda").manual_seed(5)
    x = torch.randn(2, 16, 12, 20, device="cuda", generator=g).bfloat16().float()
    xn = lib.nchw_to_nhwc(x)
    # max pool fwd / bwd
    xr = x.clone().requires_grad_(True)
    ref = F.max_pool2d(xr, 2, 2)
    got = lib.nhwc_to_nchw(lib.maxpool2_fwd(xn, 16), 16)
    assert torch.equal(got, ref)
    dy = torch.randn(ref.shape, device="cuda", generator=g).bfloat16().float()
    (gx,) = torch.autograd.grad(ref, xr, dy)
    add = torch.randn(x.shape, device="cuda", generator=g).bfloat16().float()
    gotb = lib.nhwc_to_nchw(lib.maxpool2_bwd(xn, lib.nchw_to_nhwc(dy), 16, add=lib.nchw_to_nhwc(add)), 16)
    assert rel(gotb, gx + add) < 4e-3
    # bilinear 2x fwd / bwd
    ref = F.interpolate(xr, scale_factor=2, mode="bilinear", align_corners=False)
    got = lib.nhwc_to_nchw(lib.upsample2_fwd(xn, 16), 16)
    assert rel(got, ref) < 4e-3
    dy = torch.randn(ref.shape, device="cuda", generator=g).bfloat16().float()
    (gx,) = torch.autograd.grad(ref, xr, dy)
    gotb = lib.nhwc_to_nchw(lib.upsample2_bwd(lib.nchw_to_nhwc(dy), 16), 16)
    assert rel(gotb, gx) < 4e-3
    # spp reduce / broadcast
    b, s = 1, 2
    red = lib.nhwc_to_nchw(lib.spp_reduce(xn, b, s, 16, 0.5), 16)
    assert rel(red, x.view(b, s, 16, 12, 20).mean(1)) < 4e-3
    bc = lib.nhwc_to_nchw(lib.spp_broadcast(lib.nchw_to_nhwc(red), b, s, 16, 2.0, add=xn), 16)
    assert rel(bc, x + 2.0 * red.bfloat16().float().repeat(s, 1, 1, 1)) < 4e-3


# ---------------------------------------------------------------------------------------------
# modules against the oracle
#
# Outputs and losses are compared with the fp32 oracle at north_star's 1e-3.  Gradients are compared
#   (a) at 1e-2 with the PRECISION-MATCHED oracle (same fp32 oracle code, but every conv rounds its
#       input / weight to the backend's 16-bit storage type, tests/_oracle_loader.precision_matched):
#       this is the parity statement for the kernels -- same function, same linearisation point;
#   (b) with the plain fp32 oracle at a looser bound.  A 16-bit forward perturbs pre-activations by
#       eps ~ 7e-4 (fp16), which flips the ReLU mask of a fraction ~eps of the units; every flipped
#       unit is an O(1) error in dz, so ANY reduced-precision forward sits at rel-L2 ~ sqrt(eps) ~ 3e-2
#       against fp32 gradients, independent of how exact the backward kernels are (measured in situ:
#       the dgrad kernel reproduces the exact dgrad of its own inputs to 1e-5).
# ---------------------------------------------------------------------------------------------
from tests._oracle_loader import precision_matched, restore_precision  # noqa: E402

TOL_GRAD_FP32_ORACLE = 1e-1


def _act_dtype():
    from wcmc_b200 import ops
    return ops.ACT_DTYPE


def _grads(model):
    return {k: p.grad.detach().clone() for k, p in model.named_parameters() if p.grad is not None}


def _global_rel(ours, ref):
    assert sorted(ours) == sorted(ref)
    flat_o = torch.cat([ours[k].flatten() for k in sorted(ours)])
    flat_r = torch.cat([ref[k].flatten() for k in sorted(ref)])
    return rel(flat_o, flat_r)


def _compare_grads(ours, ref, tol=TOL_GRAD):
    e = _global_rel(ours, ref)
    assert e < tol, "global grad rel-L2 %.3e" % e
    return max(rel(ours[k], ref[k]) for k in ours)


@pytest.mark.parametrize("size,batch,n_in", [(48, 2, 34), (128, 8, 39)])
def test_kpcn_matches_oracle(backend, oracle, size, batch, n_in):
    torch.manual_seed(0)
    ref = oracle.KPCN(n_in).cuda()
    ours = backend.KPCN(n_in).cuda()
    ours.load_state_dict(ref.state_dict())
    data = to_cuda(make_batch(batch=batch, size=size, seed=3, paths=False))
    if n_in > 34:
        g = torch.Generator().manual_seed(9)
        extra = torch.rand(batch, n_in - 34, size, size, generator=g).cuda()
        data["kpcn_diffuse_in"] = torch.cat([data["kpcn_diffuse_in"], extra], 1)
        data["kpcn_specular_in"] = torch.cat([data["kpcn_specular_in"], extra * 0.5], 1)
    xo = data["kpcn_diffuse_in"].clone().requires_grad_(True)
    out_o = ours(dict(data, kpcn_diffuse_in=xo))
    tgt = torch.rand_like(out_o["radiance"])
    F.l1_loss(out_o["diffuse"], tgt).backward(retain_graph=True)
    F.l1_loss(out_o["specular"], tgt).backward()
    g_ours, gx_ours = _grads(ours), xo.grad.clone()

    def run_ref():
        ref.zero_grad()
        xr = data["kpcn_diffuse_in"].clone().requires_grad_(True)
        out_r = ref(dict(data, kpcn_diffuse_in=xr))
        # the L1 sign pattern is taken from OUR outputs so that a 1e-5 output difference cannot flip
        # the (discontinuous) loss gradient at a pixel
        sd = torch.sign(out_o["diffuse"].detach() - tgt) / tgt.numel()
        ss = torch.sign(out_o["specular"].detach() - tgt) / tgt.numel()
        ((out_r["diffuse"] * sd).sum() + (out_r["specular"] * ss).sum()).backward()
        return out_r, _grads(ref), xr.grad.clone()

    out_r, g_ref, gx_ref = run_ref()
    for key in ("radiance", "diffuse", "specular"):
        assert out_o[key].shape == out_r[key].shape
        assert rel(out_o[key], out_r[key]) < TOL_IMG, key
    e32 = _global_rel(g_ours, g_ref)
    h = precision_matched(ref, _act_dtype())
    try:
        out_q, g_q, gx_q = run_ref()
    finally:
        restore_precision(h)
    assert rel(out_o["diffuse"], out_q["diffuse"]) < 1e-4
    # even the precision-matched oracle rounds differently at fp16 ties (fp32 accumulation order), so a
    # ~1e-4 fraction of ReLU masks still differs: sqrt(1e-4) = 1e-2 is the floor of this comparison; the
    # input gradient sits at the end of the 9-layer chain and compounds it
    worst = _compare_grads(g_ours, g_q, 2 * TOL_GRAD)
    assert rel(gx_ours, gx_q) < TOL_GRAD_FP32_ORACLE
    print("KPCN grads: vs precision-matched oracle %.2e (worst tensor %.2e), vs fp32 oracle %.2e"
          % (_global_rel(g_ours, g_q), worst, e32))
    record("test_kpcn_matches_oracle[%d-%d-%d]" % (size, batch, n_in),
           radiance=rel(out_o["radiance"], out_r["radiance"]), diffuse=rel(out_o["diffuse"], out_r["diffuse"]),
           specular=rel(out_o["specular"], out_r["specular"]), bound_img=TOL_IMG,
           grads_vs_precision_matched=_global_rel(g_ours, g_q), worst_tensor_vs_precision_matched=worst,
           bound_pm=2 * TOL_GRAD, grads_vs_fp32=e32, input_grad_vs_pm=rel(gx_ours, gx_q),
           input_grad_vs_fp32=rel(gx_ours, gx_ref), bound_fp32_small_case=TOL_GRAD_FP32_ORACLE)
    assert e32 < TOL_GRAD_FP32_ORACLE
    assert rel(gx_ours, gx_ref) < 2 * TOL_GRAD_FP32_ORACLE


@pytest.mark.parametrize("size,batch,spp,outc", [(16, 1, 2, 3), (32, 2, 3, 4), (128, 2, 8, 3)])
def test_pathnet_matches_oracle(backend, oracle, size, batch, spp, outc):
    torch.manual_seed(0)
    ref = oracle.PathNet(36, outc=outc).cuda()
    ours = backend.PathNet(36, outc=outc).cuda()
    ours.load_state_dict(ref.state_dict())
    data = to_cuda(make_batch(batch=batch, spp=spp, size=size, seed=4))
    po = ours(data)
    w = torch.randn_like(po)
    (po * w).mean().backward()

    def run_ref():
        ref.zero_grad()
        pr = ref(data)
        (pr * w).mean().backward()
        return pr, _grads(ref)

    pr, g_ref = run_ref()
    assert po.shape == pr.shape == (batch, spp, outc, size, size)
    assert rel(po, pr) < 2e-3   # fp16 storage through 20 layers (measured ~7e-4); losses are checked below
    e32 = _global_rel(_grads(ours), g_ref)
    h = precision_matched(ref, _act_dtype())
    try:
        pq, g_q = run_ref()
    finally:
        restore_precision(h)
    assert rel(po, pq) < 2e-3
    # 20 layers + pooling compound the ReLU-mask flips (see the header comment); the north-star bar is
    # checked on the full step below
    _compare_grads(_grads(ours), g_q, TOL_GRAD_FP32_ORACLE)
    print("PathNet grads: vs precision-matched oracle %.2e, vs fp32 oracle %.2e" % (_global_rel(_grads(ours), g_q), e32))
    record("test_pathnet_matches_oracle[%d-%d-%d-%d]" % (size, batch, spp, outc), p_buffer=rel(po, pr), bound_out=2e-3,
           grads_vs_precision_matched=_global_rel(_grads(ours), g_q), grads_vs_fp32=e32,
           bound_small_case=TOL_GRAD_FP32_ORACLE)
    assert e32 < TOL_GRAD_FP32_ORACLE


def _fingerprint_check(tag, models, g, grad_bound, names=None):
    """Per-tensor relative L2 of every gradient / updated parameter against the REFERENCE's own step through the
    sign-projection fingerprints of the golden fixture (tests/_proj.py); every value is recorded."""
    from tests import _proj
    for name, m in models.items():
        if name not in g.get("grad_fp", {}) or any(p.grad is None for p in m.parameters()):
            continue
        pn = [n for n, _ in m.named_parameters()]
        e_g = _proj.est_rel(_proj.fingerprints([p.grad for p in m.parameters()]), g["grad_fp"][name])
        e_all = _proj.est_rel_global(_proj.fingerprints([p.grad for p in m.parameters()]), g["grad_fp"][name])
        e_p = _proj.param_rms_shift_in_lr(m.parameters(), g["param_fp"][name], 1e-4)
        worst = int(e_g.argmax())
        record(tag, model=name, grad_rel_all_tensors=e_all, grad_rel_worst_tensor=float(e_g[worst]),
               worst_tensor=pn[worst], grad_rel_median_tensor=float(e_g.median()),
               param_rms_shift_in_lr_worst=float(e_p.max()), bound=grad_bound)
        assert e_all < grad_bound, (name, e_all)
        # Adam's first step moves every weight by lr * sign(g): an RMS shift of 0.7 lr = 12 % of the signs of the
        # worst tensor (biases whose gradient is below the 16-bit noise floor)
        assert float(e_p.max()) < 0.7, (name, float(e_p.max()))


def _build(KPCN, PathNet, n_in, llpm, outc):
    torch.manual_seed(0)
    models = {"dncnn": KPCN(n_in)}
    if llpm:
        models["backbone_diffuse"] = PathNet(ic=36, outc=outc)
        models["backbone_specular"] = PathNet(ic=36, outc=outc)
    return models


@pytest.mark.parametrize("tag", ["wcmc", "wcmc_m10r01", "vanilla"])
def test_train_step_matches_reference_golden(backend, oracle, golden, tag):
    """KPCNInterface.train_batch + validate_batch on the GPU backend against the vectors the
    REFERENCE's own interfaces.py / losses.py produced (tests/golden/make_golden.py)."""
    g = golden["itf_" + tag]
    cfg = g["cfg"]
    llpm = cfg["use_llpm_buf"]
    ref_models = _build(oracle.KPCN, oracle.PathNet, g["n_in"], llpm, cfg["outc"])
    models = _build(backend.KPCN, backend.PathNet, g["n_in"], llpm, cfg["outc"])
    for k in models:
        models[k].load_state_dict(ref_models[k].state_dict())
        models[k].cuda()
    optims = {"optim_" + k: torch.optim.Adam(m.parameters(), lr=1e-4) for k, m in models.items()}
    loss_funcs = {"l_diffuse": torch.nn.L1Loss(), "l_specular": torch.nn.L1Loss(), "l_recon": torch.nn.L1Loss(),
                  "l_test": backend.losses.RelativeMSE()}
    if cfg["manif_learn"]:
        loss_funcs["l_manif"] = backend.losses.FeatureMSE(non_local=True)
    itf = backend.itf.KPCNInterface(models, optims, loss_funcs, types.SimpleNamespace(model_name="t"),
                                    use_llpm_buf=llpm, manif_learn=cfg["manif_learn"], w_manif=0.1,
                                    train_branches=True, disentanglement_option=cfg["opt"])
    batch = to_cuda(make_batch(batch=2, spp=2, size=40, seed=g["data_seed"], paths=llpm))
    itf.to_train_mode()
    itf.preprocess(batch)
    torch.manual_seed(g["perm_seed"])
    itf.train_batch(batch)
    record("test_train_step_matches_reference_golden[%s]" % tag, bound_loss=TOL_IMG, bound_manif=TOL_MANIF,
           **{k: rel(itf.m_losses[k].cpu(), v) for k, v in g["losses"].items()})
    for k, v in g["losses"].items():
        # the raw manifold term amplifies the p-buffer's fp16 storage error (~7e-4) about 4x
        assert rel(itf.m_losses[k].cpu(), v) < (TOL_MANIF if "manif" in k else TOL_IMG), k
    for name, m in models.items():
        gs = torch.stack([p.grad.detach().double().abs().sum() for p in m.parameters()]).cpu()
        record("test_train_step_matches_reference_golden[%s]" % tag, model=name,
               grad_abs_sums_rel=rel(gs, g["grad_abs_sums"][name]), bound=3e-2)
        assert rel(gs, g["grad_abs_sums"][name]) < 3e-2, name
        # Adam's first step moves every weight by exactly lr * sign(grad): a gradient whose sign differs
        # (|g| below the precision floor) shifts that weight by 2 * lr.  Allow 3 % sign differences.
        for p_, want in zip(m.parameters(), g["param_sums"][name]):
            tol = 1e-4 * (6 * p_.numel() ** 0.5 + 0.03 * p_.numel()) + 1e-3 * abs(float(want))
            assert abs(float(p_.detach().double().sum()) - float(want)) < tol
    # per-tensor relative L2 against the reference's own gradients (small case: the 16-bit forward flips ~eps of
    # the ReLU masks, see TOL_GRAD_FP32_ORACLE above; 1e-2 is asserted at the north-star size)
    _fingerprint_check("test_train_step_matches_reference_golden[%s]" % tag, models, g, TOL_GRAD_FP32_ORACLE)
    itf.to_eval_mode()
    with torch.no_grad():
        rad, _ = itf.validate_batch(batch)
    assert rel(rad.cpu(), g["val_radiance"]) < TOL_IMG
    assert rel(itf.m_losses["m_val"].cpu(), g["m_val"]) < 5e-3


def test_full_size_wcmc_step_vs_oracle(backend, oracle):
    """BASELINE configs[1]: B=8, S=8, 128^2, C=3 (n_in=39), FeatureMSE non-local, w_manif 0.1."""
    ref_models = _build(oracle.KPCN, oracle.PathNet, 39, True, 3)
    models = _build(backend.KPCN, backend.PathNet, 39, True, 3)
    for k in models:
        models[k].load_state_dict(ref_models[k].state_dict())
        models[k].cuda()
        ref_models[k].cuda()
    optims = {"optim_" + k: torch.optim.Adam(m.parameters(), lr=1e-4) for k, m in models.items()}
    ref_optims = {"optim_" + k: torch.optim.Adam(m.parameters(), lr=1e-4) for k, m in ref_models.items()}
    loss_funcs = {"l_diffuse": torch.nn.L1Loss(), "l_specular": torch.nn.L1Loss(), "l_recon": torch.nn.L1Loss(),
                  "l_test": backend.losses.RelativeMSE(), "l_manif": backend.losses.FeatureMSE(non_local=True)}
    itf = backend.itf.KPCNInterface(models, optims, loss_funcs, types.SimpleNamespace(model_name="t"),
                                    use_llpm_buf=True, manif_learn=True, w_manif=0.1, train_branches=True)
    batch = to_cuda(make_batch(batch=8, spp=8, size=128, seed=1234))
    itf.to_train_mode()
    itf.preprocess(batch)
    torch.manual_seed(77)
    itf.train_batch(batch)
    torch.manual_seed(77)
    loss, _, _ = oracle.ref.kpcn_train_step(ref_models, ref_optims, batch, use_llpm_buf=True, manif_learn=True,
                                            w_manif=0.1)
    record("test_full_size_wcmc_step_vs_oracle", bound_loss=TOL_IMG, bound_manif=TOL_IMG,
           **{k: rel(itf.m_losses["m_" + k], v) for k, v in loss.items()})
    for k, v in loss.items():
        assert rel(itf.m_losses["m_" + k], v) < TOL_IMG, k        # north_star: every loss, manifold terms included
    for name in models:
        go_, gr_ = _grads(models[name]), _grads(ref_models[name])
        per = {k: rel(go_[k], gr_[k]) for k in go_}
        record("test_full_size_wcmc_step_vs_oracle", model=name, grads_vs_fp32_oracle=_global_rel(go_, gr_),
               worst_tensor=max(per.values()), worst_tensor_name=max(per, key=per.get), bound=TOL_GRAD)
        # the north-star configuration meets 1e-2 against the plain fp32 oracle
        worst = _compare_grads(_grads(models[name]), _grads(ref_models[name]), tol=TOL_GRAD)
        print(name, "grad rel-L2 vs fp32 oracle", _global_rel(_grads(models[name]), _grads(ref_models[name])),
              "worst tensor", worst)
        po = torch.cat([p.detach().flatten() for p in models[name].parameters()])
        pr = torch.cat([p.detach().flatten() for p in ref_models[name].parameters()])
        assert rel(po, pr) < 5e-3, name
    itf.to_eval_mode()
    with torch.no_grad():
        rad, _ = itf.validate_batch(batch)
        rad_ref, _, relmse_ref = oracle.ref.kpcn_validate(ref_models, batch, use_llpm_buf=True)
    record("test_full_size_wcmc_step_vs_oracle", val_radiance=rel(rad, rad_ref),
           val_relmse=rel(itf.m_losses["m_val"], relmse_ref), bound=TOL_IMG)
    assert rel(rad, rad_ref) < TOL_IMG
    assert rel(itf.m_losses["m_val"], relmse_ref) < TOL_IMG
    # size-independent properties: radiance recombination and finite gradients everywhere
    for m in models.values():
        for p in m.parameters():
            assert torch.isfinite(p).all()


def test_graphed_step_matches_eager(backend, oracle):
    """wcmc_b200.engine.GraphedTrainStep replays the same kernels: identical to the eager step when no
    random permutation is involved; finite and close in loss with the device-side permutations."""
    from wcmc_b200.engine import GraphedTrainStep

    def make(llpm):
        torch.manual_seed(0)
        models = _build(backend.KPCN, backend.PathNet, 39 if llpm else 34, llpm, 3)
        for m in models.values():
            m.cuda()
        optims = {"optim_" + k: torch.optim.Adam(m.parameters(), lr=1e-4) for k, m in models.items()}
        lf = {"l_diffuse": torch.nn.L1Loss(), "l_specular": torch.nn.L1Loss(), "l_recon": torch.nn.L1Loss(),
              "l_test": backend.losses.RelativeMSE()}
        if llpm:
            lf["l_manif"] = backend.losses.FeatureMSE(non_local=True, rng="device")
        itf = backend.itf.KPCNInterface(models, optims, lf, types.SimpleNamespace(model_name="t"), use_llpm_buf=llpm,
                                        manif_learn=llpm, w_manif=0.1, train_branches=True)
        itf.to_train_mode()
        return itf, models

    for llpm in (False, True):
        batch = to_cuda(make_batch(batch=2, spp=2, size=48, seed=31, paths=llpm))
        itf_e, models_e = make(llpm)
        itf_g, models_g = make(llpm)
        step = GraphedTrainStep(itf_g, batch)
        for _ in range(2):
            itf_e.preprocess(batch)
            itf_e.train_batch(batch)
            step(batch)
        assert itf_g.iters == itf_e.iters == 2
        for k in itf_e.m_losses:
            tol = 0.5 if (llpm and "manif" in k) else (5e-2 if llpm else 1e-6)  # different random pairings on a tiny problem
            assert rel(itf_g.m_losses[k], itf_e.m_losses[k]) < tol, k
        if not llpm:
            for name in models_e:
                # bias gradients are reduced with float atomics (order-dependent in the last bits) and
                # Adam's first steps move a weight by lr * sign(g): allow a few sign differences
                pe = torch.cat([p.detach().flatten() for p in models_e[name].parameters()])
                pg = torch.cat([p.detach().flatten() for p in models_g[name].parameters()])
                assert rel(pg, pe) < 5e-3


def test_graphed_step_survives_an_eager_backward_between_replays(backend):
    """An eager backward pass between two replays re-points every `p.grad`; if the optimiser's descriptors are then
    re-read (`load_state_dict`, an lr change: FusedClipAdam.refresh) they must still name the gradient tensors the
    captured kernels write.  Found with bench.py's long-run reset; twin runs (with / without the interleaved eager
    pass, which takes no update of its own) must follow the same trajectory."""
    from wcmc_b200.engine import GraphedTrainStep

    def make():
        torch.manual_seed(0)
        models = _build(backend.KPCN, backend.PathNet, 34, False, 3)
        for m in models.values():
            m.cuda()
        optims = {"optim_" + k: torch.optim.Adam(m.parameters(), lr=1e-4) for k, m in models.items()}
        lf = {"l_diffuse": torch.nn.L1Loss(), "l_specular": torch.nn.L1Loss(), "l_recon": torch.nn.L1Loss(),
              "l_test": backend.losses.RelativeMSE()}
        itf = backend.itf.KPCNInterface(models, optims, lf, types.SimpleNamespace(model_name="t"), use_llpm_buf=False,
                                        manif_learn=False, train_branches=True)
        itf.to_train_mode()
        return itf, models, optims

    batch = to_cuda(make_batch(batch=2, spp=2, size=48, seed=32, paths=False))
    other = to_cuda(make_batch(batch=2, spp=2, size=48, seed=33, paths=False))
    runs = []
    for interleave in (False, True):
        itf, models, optims = make()
        step = GraphedTrainStep(itf, batch)
        step(batch)
        if interleave:
            # what an eager train_batch between two replays does, minus the update itself: zero_grad(set_to_none),
            # a backward pass on another batch, and the fused optimiser re-reading its view of the model (this
            # rewrites, in place, the pinned descriptors the captured graph copies in on every replay)
            for p in step._params:
                p.grad = None
            out = models["dncnn"](other)
            (out["radiance"].mean()).backward()
            assert step._params[0].grad is not step._grads[0]
            assert itf._fused() is step.fused
            step.fused.refresh()
            assert step.fused._key is not step._adam_key
        step(batch)
        # the descriptors name the graph's own gradient tensors again, and p.grad is pinned back
        want = [g.data_ptr() for g in step._grads if g is not None]
        assert [e[1] for e in step.fused._key] == want
        assert all(p.grad is g for p, g in zip(step._params, step._grads))
        step(batch)
        runs.append(torch.cat([p.detach().flatten() for m in models.values() for p in m.parameters()]).clone())
        step.release()
    # same trajectory up to the order of the fp32 atomics of the bias gradients (cf. test_graphed_step_matches_eager)
    assert rel(runs[1], runs[0]) < 2e-3


def test_full_frame_denoise_vs_oracle_and_tiling(backend, oracle):
    """wcmc_b200.inference.denoise_frame (configs[3] at a small frame): equals the oracle KPCN on the
    replicate-padded frame, and -- the property that justifies not tiling -- an interior 92x92 block
    equals what the reference's 128x128 tile protocol computes for that tile."""
    from wcmc_b200 import inference
    torch.manual_seed(0)
    ref = oracle.KPCN(34)
    net = backend.KPCN(34)
    net.load_state_dict(ref.state_dict())
    net.cuda().eval()
    ref.cuda().eval()
    h, w = 150, 212
    batch = to_cuda({k: v for k, v in make_batch(batch=1, size=0, height=h, width=w, seed=9, paths=False,
                                                 llpm_channel=False).items() if k.startswith("kpcn")})
    out = inference.denoise_frame(net, batch)
    assert tuple(out["radiance"].shape) == (1, 3, h, w)
    with torch.no_grad():
        want = ref(inference.pad_frame(batch))
    for k in ("radiance", "diffuse", "specular"):
        assert rel(out[k], want[k]) < TOL_IMG, k
    # tile protocol: a 128x128 crop whose 92x92 centre is at least 10 px (the 21x21 gather radius) away from
    # the crop's valid border sees zero padding where the full frame has data, so compare the part of
    # the centre the tile computes from in-tile data only: the inner 72x72.
    y0, x0 = 11, 40
    tile = {k: v[..., y0:y0 + 128, x0:x0 + 128].contiguous() for k, v in batch.items()}
    with torch.no_grad():
        t = net(tile)["radiance"]                        # (1,3,92,92) = frame rows y0+18 .. y0+110
    full = out["radiance"][..., y0 + 18:y0 + 110, x0 + 18:x0 + 110]
    assert rel(t[..., 10:-10, 10:-10], full[..., 10:-10, 10:-10]) < TOL_IMG


def test_720p_denoise_fused_vs_oracle(backend, oracle):
    """BASELINE configs[3] at its full size: 1280x720, whole frame in one pass, last conv -> softmax -> kernel-apply
    fused (logits never in HBM) against the fp32 oracle KPCN on the replicate-padded frame (north_star: <= 1e-3), and
    against the un-fused launches of the same backend."""
    from wcmc_b200 import inference, ops
    torch.manual_seed(0)
    ref = oracle.KPCN(34)
    net = backend.KPCN(34)
    net.load_state_dict(ref.state_dict())
    net.cuda().eval()
    ref.cuda().eval()
    batch = to_cuda({k: v for k, v in make_batch(batch=1, size=0, height=720, width=1280, seed=7, paths=False,
                                                 llpm_channel=False).items() if k.startswith("kpcn")})
    assert ops.FUSE_KERNEL_APPLY
    out = inference.denoise_frame(net, batch)
    ops.FUSE_KERNEL_APPLY = False
    try:
        unfused = inference.denoise_frame(net, batch)
    finally:
        ops.FUSE_KERNEL_APPLY = True
    with torch.no_grad():
        want = ref(inference.pad_frame(batch))
    errs = {k: rel(out[k], want[k]) for k in ("radiance", "diffuse", "specular")}
    record("test_720p_denoise_fused_vs_oracle", bound=TOL_IMG, fused_vs_unfused=rel(out["radiance"], unfused["radiance"]),
           **errs)
    assert tuple(out["radiance"].shape) == (1, 3, 720, 1280)
    assert all(e < TOL_IMG for e in errs.values()), errs
    assert rel(out["radiance"], unfused["radiance"]) < 1e-4


def test_fused_clip_adam_matches_torch(backend):
    """wcmc_adam_clip_step on torch.optim.Adam's own state == clip_grad_value_ + Adam.step() of torch,
    over several steps, tensors of awkward sizes, two optimisers with different learning rates; the
    optimiser state stays loadable by torch (state_dict round trip); ok_flag = 0 leaves everything alone."""
    from wcmc_b200 import optim as wopt
    g = torch.Generator(device="cuda").manual_seed(0)
    shapes = [(100, 39, 5, 5), (100,), (441, 100, 5, 5), (7,), (64, 36, 1, 1), (4097,), (1,)]

    def make():
        ps = [torch.nn.Parameter(torch.randn(s, device="cuda", generator=torch.Generator(device="cuda").manual_seed(i)))
              for i, s in enumerate(shapes)]
        return ps, [torch.optim.Adam(ps[:4], lr=1e-4), torch.optim.Adam(ps[4:], lr=1e-6)]

    pa, oa = make()
    pb, ob = make()
    fused = wopt.FusedClipAdam(ob)
    assert all(wopt.supported(o) for o in ob)
    for step in range(5):
        grads = [torch.randn(s, device="cuda", generator=g) * (3.0 if step % 2 else 0.3) for s in shapes]
        for p, q, gr in zip(pa, pb, grads):
            p.grad = gr.clone()
            q.grad = gr.clone()
        torch.nn.utils.clip_grad_value_(pa, 1.0)
        for o in oa:
            o.step()
        fused.step(clip=1.0)
        for p, q in zip(pa, pb):
            assert torch.equal(p.grad, q.grad)                       # clipped gradient written back
            torch.testing.assert_close(q.detach(), p.detach(), rtol=1e-6, atol=1e-9)
    for o1, o2 in zip(oa, ob):
        s1, s2 = o1.state_dict(), o2.state_dict()
        for k in s1["state"]:
            assert float(s1["state"][k]["step"]) == float(s2["state"][k]["step"]) == 5.0
            # same operations, but fused multiply-adds round once where torch's separate kernels round twice
            torch.testing.assert_close(s2["state"][k]["exp_avg"], s1["state"][k]["exp_avg"], rtol=1e-5, atol=2e-7)
            torch.testing.assert_close(s2["state"][k]["exp_avg_sq"], s1["state"][k]["exp_avg_sq"], rtol=1e-5, atol=2e-7)
    before = [q.detach().clone() for q in pb]
    for q in pb:
        q.grad = torch.ones_like(q)
    fused.step(clip=1.0, ok_flag=torch.zeros(1, dtype=torch.int32, device="cuda"), count=False)
    for q, b in zip(pb, before):
        assert torch.equal(q.detach(), b)
    assert int(fused.t_dev) == 5
    # a NaN / inf gradient element (fp16 overflow in a 16-bit backward pass) is skipped and counted, never applied
    for q in pb:
        q.grad = torch.full_like(q, 0.5)
    pb[0].grad.view(-1)[3] = float("nan")
    pb[2].grad.view(-1)[10] = float("inf")
    before = [q.detach().clone() for q in pb]
    fused.step(clip=1.0)
    assert fused.nonfinite_count() == 2
    assert float(pb[0].detach().view(-1)[3]) == float(before[0].view(-1)[3])
    assert float(pb[2].detach().view(-1)[10]) == float(before[2].view(-1)[10])
    assert all(torch.isfinite(q).all() for q in pb) and not torch.equal(pb[1].detach(), before[1])
    st = ob[0].state[pb[0]]
    assert torch.isfinite(st["exp_avg"]).all() and torch.isfinite(st["exp_avg_sq"]).all()


def test_fused_adam_checkpoint_is_interchangeable_with_torch(backend, tmp_path):
    """The reference checkpoints by pickling the optimiser OBJECTS (`'optims': itf.optims`, train_kpcn.py:114,143):
    `Optimizer.__getstate__` runs no state-dict hook, so `state['step']` must be current at all times.  After a
    save / load round trip torch's own Adam must continue exactly where the fused kernel stopped (bias corrections
    included), a load_state_dict + lr change must be honoured by the next fused step, and optimisers at different
    step counts make the interface fall back to torch's step instead of raising."""
    import io
    import pickle
    from wcmc_b200 import optim as wopt
    shapes = [(33, 7, 3, 3), (33,), (130,), (5, 5)]

    def make(lr=1e-3):
        ps = [torch.nn.Parameter(torch.randn(s, device="cuda", generator=torch.Generator(device="cuda").manual_seed(i)))
              for i, s in enumerate(shapes)]
        return ps, [torch.optim.Adam(ps[:2], lr=lr), torch.optim.Adam(ps[2:], lr=lr)]

    def set_grads(ps, seed):
        g = torch.Generator(device="cuda").manual_seed(seed)
        for q in ps:
            q.grad = torch.randn(q.shape, device="cuda", generator=g)

    pa, oa = make()
    pb, ob = make()
    fused = wopt.FusedClipAdam(ob)
    for step in range(4):
        set_grads(pa, step)
        set_grads(pb, step)
        torch.nn.utils.clip_grad_value_(pa, 1.0)
        for o in oa:
            o.step()
        fused.step(clip=1.0)
        for o in ob:   # what pickling sees, with no hook involved
            assert all(float(st["step"]) == step + 1 for st in o.state.values())
    buf = io.BytesIO()
    torch.save({"optims": ob, "params": pb}, buf)          # train_kpcn.py:109-116 pickles the objects
    buf.seek(0)
    ck = torch.load(buf, weights_only=False)
    pc, oc = ck["params"], ck["optims"]
    assert all(float(st["step"]) == 4.0 for o in oc for st in o.state.values())
    # torch's Adam continues from the unpickled fused state exactly like from its own state
    set_grads(pa, 99)
    set_grads(pc, 99)
    torch.nn.utils.clip_grad_value_(pa, 1.0)
    torch.nn.utils.clip_grad_value_(pc, 1.0)
    for o in oa + oc:
        o.step()
    for p_, q_ in zip(pa, pc):
        torch.testing.assert_close(q_.detach(), p_.detach(), rtol=1e-5, atol=1e-7)
    # load_state_dict into the LIVE fused optimisers (resume, train_kpcn.py:283-296) + an lr change: both honoured
    for o, src in zip(ob, oa):
        o.load_state_dict(pickle.loads(pickle.dumps(src.state_dict())))
        o.param_groups[0]["lr"] = 5e-4
    for o in oa:
        o.param_groups[0]["lr"] = 5e-4
    for p_, q_ in zip(pa, pb):
        q_.data.copy_(p_.data)
    set_grads(pa, 7)
    set_grads(pb, 7)
    torch.nn.utils.clip_grad_value_(pa, 1.0)
    for o in oa:
        o.step()
    fused.step(clip=1.0)
    assert fused.t == 6 and int(fused.t_dev) == 6
    for p_, q_ in zip(pa, pb):
        torch.testing.assert_close(q_.detach(), p_.detach(), rtol=1e-5, atol=1e-7)
    # different step counts (only one optimiser resumed): consistent() is False -> the interface keeps torch's step
    pd, od = make()
    od[0].load_state_dict(pickle.loads(pickle.dumps(oa[0].state_dict())))
    assert not wopt.FusedClipAdam(od).consistent()
    itf = backend.itf.KPCNInterface.__new__(backend.itf.KPCNInterface)
    itf.fused_optim, itf._fused_adam = True, None
    itf.models = {"a": None, "b": None}
    itf.optims = {"optim_a": od[0], "optim_b": od[1]}
    assert itf._fused() is None



# ---------------------------------------------------------------------------------------------
# the ablation interfaces on the same kernels, against vectors produced by the REFERENCE's
# own KPCNRefInterface / KPCNPreInterface (tests/golden/make_golden_n4.py)
# ---------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def golden_n4():
    import os
    from tests.conftest import ROOT
    return torch.load(os.path.join(ROOT, "tests", "golden", "ref_golden_n4.pt"), weights_only=False)


def _loss_funcs(backend, manif):
    lf = {"l_diffuse": torch.nn.L1Loss(), "l_specular": torch.nn.L1Loss(), "l_recon": torch.nn.L1Loss(),
          "l_test": backend.losses.RelativeMSE()}
    if manif:
        lf["l_manif"] = backend.losses.FeatureMSE(non_local=True)
    return lf


def _check_step_vs_golden(models, itf, g):
    assert set(itf.m_losses) == set(g["losses"])
    for k, v in g["losses"].items():
        assert rel(itf.m_losses[k].cpu(), v) < (TOL_MANIF if "manif" in k else TOL_IMG), k
    for name, m in models.items():
        assert m.training == g["training"][name], name
        has_grads = float(g["grad_abs_sums"][name].min()) >= 0
        if has_grads:
            gs = torch.stack([p.grad.detach().double().abs().sum() for p in m.parameters()]).cpu()
            # models the stage does not train (eval mode) still receive gradients in the reference -- through the
            # regression branch only, two orders of magnitude smaller and never applied: looser bound
            assert rel(gs, g["grad_abs_sums"][name]) < (3e-2 if m.training else 6e-2), name
        for p_, want in zip(m.parameters(), g["param_sums"][name]):
            tol = 1e-4 * (6 * p_.numel() ** 0.5 + 0.03 * p_.numel()) + 1e-3 * abs(float(want))
            assert abs(float(p_.detach().double().sum()) - float(want)) < tol, name
    _fingerprint_check("n4_interfaces[%s]" % "+".join(sorted(itf.m_losses)), models, g, TOL_GRAD_FP32_ORACLE)


def test_ref_interface_matches_reference_golden(backend, oracle, golden_n4):
    g = golden_n4["ref"]
    torch.manual_seed(0)
    ref_models = {"dncnn": oracle.KPCN(g["n_in"])}
    models = {"dncnn": backend.KPCN(g["n_in"])}
    models["dncnn"].load_state_dict(ref_models["dncnn"].state_dict())
    models["dncnn"].cuda()
    optims = {"optim_dncnn": torch.optim.Adam(models["dncnn"].parameters(), lr=1e-4)}
    itf = backend.itf.KPCNRefInterface(models, optims, _loss_funcs(backend, False), types.SimpleNamespace(model_name="t"))
    batch = to_cuda(make_batch(batch=2, spp=2, size=40, seed=g["data_seed"], paths=False))
    itf.to_train_mode()
    itf.preprocess(batch)
    itf.train_batch(batch)
    _check_step_vs_golden(models, itf, g)
    itf.to_eval_mode()
    with torch.no_grad():
        rad, pb = itf.validate_batch(batch)
    assert pb is None
    assert rel(rad.cpu(), g["val_radiance"]) < TOL_IMG
    assert rel(itf.m_losses["m_val"].cpu(), g["m_val"]) < 5e-3


@pytest.mark.parametrize("tag", ["pre_manifold", "pre_regress"])
def test_pre_interface_matches_reference_golden(backend, oracle, golden_n4, tag):
    g = golden_n4[tag]
    ref_models = _build(oracle.KPCN, oracle.PathNet, 39, True, 3)
    models = _build(backend.KPCN, backend.PathNet, 39, True, 3)
    for k in models:
        models[k].load_state_dict(ref_models[k].state_dict())
        models[k].cuda()
    optims = {"optim_" + k: torch.optim.Adam(m.parameters(), lr=1e-4) for k, m in models.items()}
    before = {k: [p.detach().clone() for p in m.parameters()] for k, m in models.items()}
    itf = backend.itf.KPCNPreInterface(models, optims, _loss_funcs(backend, True), types.SimpleNamespace(model_name="t"),
                                       manif_learn=g["manif_learn"], w_manif=0.1)
    batch = to_cuda(make_batch(batch=2, spp=2, size=40, seed=g["data_seed"], paths=True))
    itf.to_train_mode()
    itf.preprocess(batch)
    torch.manual_seed(g["perm_seed"])
    itf.train_batch(batch)
    _check_step_vs_golden(models, itf, g)
    # the frozen side is bit-for-bit untouched
    frozen = ["dncnn"] if g["manif_learn"] else ["backbone_diffuse", "backbone_specular"]
    for k in frozen:
        assert all(torch.equal(a, b) for a, b in zip(before[k], models[k].parameters())), k


def test_tile_protocol_vs_one_pass(backend, oracle):
    """`wcmc_b200.inference.inference` (the reference's tile protocol, test_models.py:49-101,
    through KPCNInterface.validate_batch) against the same frame in ONE pass: identical inside the 28-pixel
    margin `test_models.denoise` crops (:217-219), and the tiled result equals the oracle KPCN tile by tile."""
    from wcmc_b200 import inference as inf
    torch.manual_seed(0)
    ref = oracle.KPCN(34)
    net = backend.KPCN(34)
    net.load_state_dict(ref.state_dict())
    net.cuda()
    ref.cuda().eval()
    itf = backend.itf.KPCNInterface({"dncnn": net}, {"optim_dncnn": torch.optim.Adam(net.parameters(), lr=1e-4)},
                                    _loss_funcs(backend, False), types.SimpleNamespace(model_name="t"))
    h, w = 192, 256
    b = make_batch(batch=1, size=0, height=h, width=w, seed=12, paths=False, llpm_channel=False)
    frame = {k: v[0] for k, v in b.items()}
    ds = inf.FrameTiles(frame)
    loader = torch.utils.data.DataLoader(ds, batch_size=4, shuffle=False)
    tiled, pb = inf.inference(itf, loader)
    assert pb is None and tuple(tiled.shape) == (3, h, w)
    m_val_tiled = float(itf.m_losses["m_val"])
    one, _ = inf.inference_one_pass(itf, frame)
    c = 28
    assert rel(one[:, c:-c, c:-c], tiled[:, c:-c, c:-c]) < 1e-5
    # oracle on one tile, stitched region of that tile
    patch, i0, j0, i1, j1, i, j = ds[4]
    with torch.no_grad():
        want = ref({k: v.unsqueeze(0).cuda() for k, v in patch.items()})["radiance"]
    want = F.pad(want, (18, 18, 18, 18), "replicate")[0]
    assert rel(tiled[:, i0:i1, j0:j1], want[:, i0 - i:i1 - i, j0 - j:j1 - j]) < TOL_IMG
    assert m_val_tiled > 0


@pytest.mark.parametrize("size,batch,spp,outc", [(20, 2, 3, 3), (32, 1, 2, 12), (64, 3, 4, 3)])
def test_fused_mlp_backward_vs_generic_path(backend, oracle, size, batch, spp, outc):
    """K8 / K9 (one kernel per MLP backward) against the generic 1x1-conv dgrad / wgrad launches of the same
    backend and against the fp32 oracle: same gradients for every PathNet parameter.  20x20 = 400 pixels is not a
    multiple of the 128-row tile (partial tiles), batch 3 x 32 tiles x 2 groups exercises the persistent loop."""
    from wcmc_b200 import ops
    torch.manual_seed(0)
    ref = oracle.PathNet(36, outc=outc).cuda()
    ours = backend.PathNet(36, outc=outc).cuda()
    ours.load_state_dict(ref.state_dict())
    data = to_cuda(make_batch(batch=batch, spp=spp, size=size, seed=6))
    w = None
    res = {}
    saved = ops.FUSED_MLP_BWD
    try:
        for mode in (True, False):
            ops.FUSED_MLP_BWD = mode
            ours.zero_grad()
            po = ours(data)
            if w is None:
                w = torch.randn_like(po)
            (po * w).mean().backward()
            res[mode] = _grads(ours)
    finally:
        ops.FUSED_MLP_BWD = saved
    ref.zero_grad()
    (ref(data) * w).mean().backward()
    g_ref = _grads(ref)
    e_fused, e_generic = _global_rel(res[True], g_ref), _global_rel(res[False], g_ref)
    print("PathNet grads vs fp32 oracle: fused MLP backward %.2e, generic %.2e" % (e_fused, e_generic))
    per = {k: rel(res[True][k], res[False][k]) for k in res[True]}
    worst = sorted(per.items(), key=lambda kv: -kv[1])[:4]
    print("fused vs generic, worst tensors:", ["%s %.2e" % kv for kv in worst])
    assert e_fused < max(2e-2, 1.5 * e_generic)
    # every tensor separately (a wrong bias / small weight block would hide in the global norm)
    for k, e in per.items():
        assert e < 5e-2, "%s: fused vs generic rel-L2 %.3e" % (k, e)


def test_batched_weight_norm_vs_torch(backend):
    """ops.BatchedWeightNormFn (one launch for all layers, each direction) against torch._weight_norm + autograd."""
    from wcmc_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(3)
    shapes = [(64, 36, 1, 1), (64, 64, 3, 3), (512, 512, 3, 3), (3, 128, 1, 1), (128, 192, 3, 3)]
    vs = [torch.randn(s, device="cuda", generator=g).requires_grad_(True) for s in shapes]
    gs = [(torch.rand(s[0], 1, 1, 1, device="cuda", generator=g) + 0.5).requires_grad_(True) for s in shapes]
    flat = [t for pair in zip(vs, gs) for t in pair]
    ws = ops.BatchedWeightNormFn.apply(*flat)
    ref = [torch._weight_norm(v, gg, 0) for v, gg in zip(vs, gs)]
    ups = [torch.randn(s, device="cuda", generator=g) for s in shapes]
    for w, r in zip(ws, ref):
        assert rel(w, r) < 1e-6
    got = torch.autograd.grad([(w * u).sum() for w, u in zip(ws, ups)], flat)
    want = torch.autograd.grad([(r * u).sum() for r, u in zip(ref, ups)], flat)
    for a, b in zip(got, want):
        assert a.shape == b.shape and rel(a, b) < 1e-5
