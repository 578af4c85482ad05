"""CTA-pair (cta_group::2) conv launches against the single-CTA launches of the same kernel.
   python tools/pair_check.py [reps]
For every layer shape: forward and data-gradient with flags bit 20 (pair) and bit 21 (single), each with row
stages (a weight stage = one kernel row of taps) and with bit 22 (one tap per stage); the outputs must be
bit-identical (same MMAs per output element, same K order), then they are timed (CUDA events, L2 flushed
between reps).  Run in its own process: a pipeline bug traps the context (bounded mbarrier waits).
Exit code 1 on any mismatch."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wcmc_b200 import lib  # noqa: E402

lib.init()
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
dt = torch.float16
PAIR, SINGLE = 1 << 20, 1 << 21
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
g = torch.Generator(device="cuda").manual_seed(0)


def timeit(fn):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


# (name, N, H, W, cin, cout, k, pad, fp32 out)
cfgs = [("odd regions 3x3 n=1 100->100 @44", 1, 44, 44, 100, 100, 5, 0, False),
        ("tiny 1 region 64->64 3x3 @16", 1, 16, 16, 64, 64, 3, 1, False),
        ("kpcn first 39->100 @128", 8, 128, 128, 39, 100, 5, 0, False),
        ("kpcn mid 100->100 @120", 8, 120, 120, 100, 100, 5, 0, False),
        ("kpcn mid 100->100 @100", 8, 100, 100, 100, 100, 5, 0, False),
        ("kpcn last 100->441 @96 (fp32 logits)", 8, 96, 96, 100, 441, 5, 0, True),
        ("unet 64->64 3x3 @128", 8, 128, 128, 64, 64, 3, 1, False),
        ("unet 192->64 3x3 @128", 8, 128, 128, 192, 64, 3, 1, False),
        ("unet 128->128 3x3 @64", 8, 64, 64, 128, 128, 3, 1, False),
        ("unet 384->128 3x3 @64", 8, 64, 64, 384, 128, 3, 1, False),
        ("unet 256->256 3x3 @32", 8, 32, 32, 256, 256, 3, 1, False)]
bad = 0
TAPS = 1 << 22
for il in (0, 1):
    if il:
        PAIR, SINGLE = PAIR | TAPS, SINGLE | TAPS
    for name, n, h, w, cin, cout, k, pad, f32 in cfgs:
        ho, wo = h + 2 * pad - k + 1, w + 2 * pad - k + 1
        x = lib.nchw_to_nhwc(torch.randn(n, cin, h, w, device="cuda", generator=g), dtype=dt)
        wt = torch.randn(cout, cin, k, k, device="cuda", generator=g) * 0.03
        bias = torch.randn(cout, device="cuda", generator=g)
        wf, wd, bp = lib.pack_weights(wt, bias, want_bias=True, dtype=dt)
        dy = lib.nchw_to_nhwc(torch.randn(n, cout, ho, wo, device="cuda", generator=g), dtype=dt)
        fl = 2.0 * n * ho * wo * k * k * cin * cout
        od = torch.float32 if f32 else None

        def fwd(flags):
            return lib.conv2d(x, wf, bp, k, pad, act=0 if f32 else 1, out_dtype=od, flags=flags)

        def dgrad(flags, colsum=None):
            return lib.conv2d(dy, wd, None, k, k - 1 - pad, act=0, mask=x, flags=flags, colsum=colsum)

        a, b, base = fwd(SINGLE), fwd(PAIR), fwd((1 << 21) | TAPS)   # base: single CTA, one tap per stage
        torch.cuda.synchronize()
        ok_f = torch.equal(a, b) and torch.equal(a, base)
        cs_a = torch.zeros(lib.pad16(cin), device="cuda")
        cs_b = torch.zeros(lib.pad16(cin), device="cuda")
        da, db = dgrad(SINGLE, cs_a), dgrad(PAIR, cs_b)
        torch.cuda.synchronize()
        ok_d = torch.equal(da, db)
        ok_c = torch.allclose(cs_a, cs_b, rtol=1e-4, atol=1e-2 * float(cs_a.abs().max()) + 1e-6)   # atomics: order differs
        if il == 0:
            t_fs, t_fp = timeit(lambda: fwd(SINGLE)), timeit(lambda: fwd(PAIR))
            t_ds, t_dp = timeit(lambda: dgrad(SINGLE)), timeit(lambda: dgrad(PAIR))
            print("%-38s fwd %s single %7.1f us (%6.1f TF/s)  pair %7.1f us (%6.1f TF/s) | dgrad %s/%s single %7.1f us  pair %7.1f us"
                  % (name, "ok " if ok_f else "BAD", t_fs * 1e3, fl / t_fs / 1e9, t_fp * 1e3, fl / t_fp / 1e9,
                     "ok" if ok_d else "BAD", "ok" if ok_c else "BAD", t_ds * 1e3, t_dp * 1e3), flush=True)
        else:
            t_fs, t_fp = timeit(lambda: fwd(SINGLE)), timeit(lambda: fwd(PAIR))
            print("%-38s [tap stages] fwd %s dgrad %s/%s  single %7.1f us  pair %7.1f us (%6.1f TF/s)"
                  % (name, "ok" if ok_f else "BAD", "ok" if ok_d else "BAD", "ok" if ok_c else "BAD", t_fs * 1e3,
                     t_fp * 1e3, fl / t_fp / 1e9), flush=True)
        bad += (not ok_f) + (not ok_d) + (not ok_c)
print("pair_check: %s" % ("all identical" if bad == 0 else "%d MISMATCHES" % bad))
sys.exit(1 if bad else 0)
