"""Kernel-apply micro-benchmark (CUDA events, L2 flushed between reps): GB/s of algorithmic bytes.
   python tools/ka_bench.py [reps]"""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wcmc_b200 import lib
l = lib.init()
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 7
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def timeit(fn):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]
g = torch.Generator(device="cuda").manual_seed(0)
for name, n, h, w in (("train 8x92x92", 8, 92, 92), ("frame 1x720x1280", 1, 720, 1280)):
    logits = torch.randn(n, h, w, 448, device="cuda", generator=g)
    data = torch.rand(n, 3, h, w, device="cuda", generator=g)
    gout = torch.randn(n, 3, h, w, device="cuda", generator=g)
    for tw in (8, 16, 32):
        assert l.wcmc_tuning_set(b"ka_tile_w", tw) == 0
        out, stats = lib.kernel_apply_fwd(logits, data, 21)
        t = timeit(lambda: lib.kernel_apply_fwd(logits, data, 21))
        by = n * h * w * (441 * 4 + 12 + 12 + 8)
        print("fwd %-18s tile %2d  %8.1f us  %7.1f GB/s" % (name, tw, t * 1e3, by / t / 1e6), flush=True)
        t = timeit(lambda: lib.kernel_apply_bwd(logits, data, out, stats, gout, 21, dl_cs=448, dtype=torch.float16))
        by = n * h * w * (441 * 4 + 441 * 2 + 12 * 3 + 8)
        print("bwd %-18s tile %2d  %8.1f us  %7.1f GB/s" % (name, tw, t * 1e3, by / t / 1e6), flush=True)
l.wcmc_tuning_set(b"ka_tile_w", 16)
