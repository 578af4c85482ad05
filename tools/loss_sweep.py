"""BASELINE.json configs[4]: path-disentangling loss standalone sweep, N in {16k, 64k, 256k} rows x
D in {16, 32}: permutation-paired form (reference semantics, HBM/L2-bound gather) and all-pairs form
(tensor-core Gram tiles reduced in the epilogue).  CUDA events, L2 flushed between reps.
   python tools/loss_sweep.py [reps]"""
import json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wcmc_b200 import allpairs, lib
lib.init()
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
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
rows = []
for n in (16384, 65536, 262144):
    for d in (16, 32):
        # perm mode wants (B,S,C,H,W): one "patch" of n rows
        hh = 128; s = n // (hh * hh) if n >= hh * hh else 1
        h = hh if n >= hh * hh else int(n ** 0.5)
        p5 = torch.rand(1, s, d, h, n // (s * h), device="cuda", generator=g)
        ref4 = torch.rand(1, 3, h, n // (s * h), device="cuda", generator=g) * 3
        ip = torch.randperm(n, device="cuda", generator=g); ib = torch.randperm(n, device="cuda", generator=g)
        t_perm = timeit(lambda: lib.fmse_perm_fwd(p5, ref4, ip, ib))
        byts = n * ((d + 3) * 4.0 * 3 + 2 * 16.0)
        pr, rr = allpairs.rows_from_pbuffer(p5, ref4)
        res = {"N": n, "D": d, "perm_us": round(t_perm * 1e3, 1), "perm_GBs": round(byts / t_perm / 1e6, 1)}
        for name, mode, tau in (("mse", 0, 0.0), ("lse", 1, 0.0), ("mse_masked", 0, 0.05)):
            t = timeit(lambda: lib.fmse_allpairs_fwd(pr, rr, mode, 2.0, tau))
            res[name + "_ms"] = round(t, 3)
            res[name + "_Gpairs_s"] = round(n * (n - 1) / 2 / t / 1e6, 1)
            res[name + "_TFLOPs_algorithmic"] = round(2.0 * n * n / 2 * (d + 3) / t / 1e9, 2)
        print(json.dumps(res), flush=True)
