"""Grouped weight-gradient micro-benchmark (CUDA events, L2 flushed between reps): the nine 5x5 layers of one KPCN
branch and the fifteen 3x3 layers of one PathNet U-Net at the north-star size (B = 8, 128^2), as ONE grouped launch
each (wcmc_conv2d_wgrad_group) and as round-1 per-layer launches (knob wgrad_group=0).
    python tools/wgrad_group_bench.py [reps] [which=kpcn|unet|both] [mode=all|group]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wcmc_b200 import lib  # noqa: E402

lib.init()
rt = lib.load()
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
which = sys.argv[2] if len(sys.argv) > 2 else "both"
mode = sys.argv[3] if len(sys.argv) > 3 else "all"
dt = torch.float16
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
g = torch.Generator(device="cuda").manual_seed(0)

KPCN = [(39, 100, 128, 5, 0)] + [(100, 100, 128 - 4 * i, 5, 0) for i in range(1, 8)] + [(100, 441, 96, 5, 0)]
UNET = ([(64, 64, 128, 3, 1)] * 3 + [(64, 128, 64, 3, 1), (128, 128, 64, 3, 1), (128, 128, 64, 3, 1), (128, 256, 32, 3, 1),
        (256, 256, 32, 3, 1), (256, 256, 32, 3, 1), (384, 128, 64, 3, 1), (128, 128, 64, 3, 1), (128, 128, 64, 3, 1),
        (192, 64, 128, 3, 1), (64, 64, 128, 3, 1), (64, 64, 128, 3, 1)])


def make(shapes):
    layers, flops = [], 0.0
    for cin, cout, h, k, pad in shapes:
        ho = h + 2 * pad - k + 1
        x = (torch.randn(8, h, h, lib.pad16(cin), device="cuda", generator=g) * 0.5).to(dt)
        x[..., cin:] = 0
        dy = (torch.randn(8, ho, ho, lib.pad16(cout), device="cuda", generator=g) * 0.1).to(dt)
        dy[..., cout:] = 0
        layers.append((x, dy, cin, cout, k, pad))
        flops += 2.0 * 8 * ho * ho * k * k * cin * cout
    return layers, flops


def run(layers):
    for x, dy, cin, cout, k, pad in layers:
        lib.conv2d_wgrad(x, dy, cout, cin, k, pad, lib.pad16(cin), lib.pad16(cout), defer=True)
    lib.wgrad_flush()


def timeit(fn, flops, name):
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
    t = sorted(ts)[len(ts) // 2]
    print("%-66s %8.1f us  %7.1f TFLOP/s (algorithmic)" % (name, t * 1e3, flops / t / 1e9), flush=True)


for tag, shapes in (("kpcn", KPCN), ("unet", UNET)):
    if which not in (tag, "both"):
        continue
    layers, fl = make(shapes)
    plan, launches = lib.wgrad_group_plan([(8, h, h, cin, cout, k, pad) for cin, cout, h, k, pad in shapes])
    print("# %s: %d layers, %.1f GFLOP, plan (teams, ctas, taps/group, column stride): %s" % (tag, len(shapes), fl / 1e9, plan))
    timeit(lambda: run(layers), fl, "%s grouped launch + reduction" % tag)
    if mode == "all":
        rt.wcmc_tuning_set(b"wgrad_group_pack", 0)
        timeit(lambda: run(layers), fl, "%s grouped, rows not packed (column stride = padded cin)" % tag)
        rt.wcmc_tuning_set(b"wgrad_group_pack", 1)
        rt.wcmc_tuning_set(b"wgrad_group", 0)
        timeit(lambda: run(layers), fl, "%s round-1 per-layer launches + one reduction" % tag)
        rt.wcmc_tuning_set(b"wgrad_group", 1)
        for i in (1, len(shapes) - 1):
            one = [layers[i]]
            cin, cout, h, k, pad = shapes[i]
            ho = h + 2 * pad - k + 1
            timeit(lambda: run(one), 2.0 * 8 * ho * ho * k * k * cin * cout, "%s layer %d alone (%d->%d @%d) grouped kernel" % (tag, i, cin, cout, h))
