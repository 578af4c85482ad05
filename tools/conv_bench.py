"""Single-layer conv micro-benchmark (CUDA events, L2 flushed between reps) for kernel tuning.
   python tools/conv_bench.py [fwd|dgrad|wgrad|all|knobs] [reps]
`knobs` times the forward kernel of the KPCN mid layer with parts of the pipeline disabled (results
are wrong on purpose) to see which stage bounds it."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wcmc_b200 import lib
lib.init()
what = sys.argv[1] if len(sys.argv) > 1 else "all"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
dt = torch.float16
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def timeit(fn, flops, name):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    t = sorted(ts)[len(ts) // 2]
    print("%-44s %8.1f us  %7.1f TFLOP/s (algorithmic)" % (name, t * 1e3, flops / t / 1e9), flush=True)
g = torch.Generator(device="cuda").manual_seed(0)
# (name, N, H, W, cin, cout, k, pad)
cfgs = [("kpcn mid 100->100 @120", 8, 120, 120, 100, 100, 5, 0),
        ("kpcn mid 100->100 @100", 8, 100, 100, 100, 100, 5, 0),
        ("kpcn first 39->100 @128", 8, 128, 128, 39, 100, 5, 0),
        ("kpcn last 100->441 @96", 8, 96, 96, 100, 441, 5, 0),
        ("unet 64->64 3x3 @128", 8, 128, 128, 64, 64, 3, 1),
        ("unet 192->64 3x3 @128", 8, 128, 128, 192, 64, 3, 1),
        ("unet 128->128 3x3 @64", 8, 64, 64, 128, 128, 3, 1),
        ("unet 384->128 3x3 @64", 8, 64, 64, 384, 128, 3, 1),
        ("unet 256->256 3x3 @32", 8, 32, 32, 256, 256, 3, 1),
        ("mlp 1x1 36->64 @64x128x128", 64, 128, 128, 36, 64, 1, 0),
        ("mlp 1x1 64->64 @64x128x128", 64, 128, 128, 64, 64, 1, 0),
        ("mlp 1x1 128->128 @64x128x128", 64, 128, 128, 128, 128, 1, 0)]
if what == "knobs":
    cfgs = [cfgs[0], cfgs[3], cfgs[4]]
for name, n, h, w, cin, cout, k, pad in cfgs:
    ho, wo = h + 2 * pad - k + 1, w + 2 * pad - k + 1
    x = lib.nchw_to_nhwc(torch.randn(n, cin, h, w, device="cuda", generator=g), dtype=dt)
    wt = torch.randn(cout, cin, k, k, device="cuda", generator=g) * 0.03
    wf, wd, bp = lib.pack_weights(wt, torch.zeros(cout, device="cuda"), want_bias=True, dtype=dt)
    dy = lib.nchw_to_nhwc(torch.randn(n, cout, ho, wo, device="cuda", generator=g), dtype=dt)
    fl = 2.0 * n * ho * wo * k * k * cin * cout
    if what == "knobs":
        for label, flags in (("production", 0), ("weights once", 1 << 16), ("halos once", 1 << 17),
                             ("weights+halos once", 3 << 16), ("no epilogue", 1 << 18), ("MMA only", 7 << 16),
                             ("mt=1", 1 << 4), ("mt=1 MMA only", (7 << 16) | (1 << 4)),
                             ("single CTA", 1 << 21), ("single CTA, tap stages", (1 << 21) | (1 << 22)),
                             ("pair, tap stages", 1 << 22), ("pair MMA only", (7 << 16))):
            timeit(lambda: lib.conv2d(x, wf, bp, k, pad, act=1, flags=flags), fl, "fwd %-22s %s" % (name, label))
        continue
    if what in ("fwd", "all"):
        timeit(lambda: lib.conv2d(x, wf, bp, k, pad, act=1), fl, "fwd   " + name)
    if what in ("dgrad", "all"):
        timeit(lambda: lib.conv2d(dy, wd, None, k, k - 1 - pad, act=0, mask=x), fl, "dgrad " + name)
    if what in ("wgrad", "all"):
        timeit(lambda: lib.conv2d_wgrad(x, dy, cout, cin, k, pad, lib.pad16(cin), lib.pad16(cout)), fl, "wgrad " + name)
        def partial_only():
            lib.conv2d_wgrad(x, dy, cout, cin, k, pad, lib.pad16(cin), lib.pad16(cout), defer=True)
            lib._pending().clear()
        timeit(partial_only, fl, "wgrad (tensor-core kernel only) " + name[:12])
        lib.load().wcmc_tuning_set(b"wgrad_uniform", 0)
        timeit(partial_only, fl, "wgrad (kernel only, splits ~ taps) " + name[:9])
        lib.load().wcmc_tuning_set(b"wgrad_uniform", 1)
