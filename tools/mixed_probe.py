"""Probe: which operand-format combinations does tcgen05.mma kind::f16 accept? (one subprocess each)"""
import subprocess, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

def run(xd, wd):
    import torch
    import torch.nn.functional as F
    from wcmc_b200 import lib
    dt = {"f16": torch.float16, "bf16": torch.bfloat16}
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn(2, 64, 40, 36, device="cuda", generator=g)
    wt = torch.randn(64, 64, 3, 3, device="cuda", generator=g) * 0.05
    xn = lib.nchw_to_nhwc(x, dtype=dt[xd])
    wf, _ = lib.pack_weights(wt, dtype=dt[wd])
    y = lib.conv2d(xn, wf, None, 3, 1, act=0, out_dtype=torch.float32)
    torch.cuda.synchronize()
    ref = F.conv2d(x.to(dt[xd]).double(), wt.to(dt[wd]).double(), padding=1)
    e = ((y[..., :64].permute(0, 3, 1, 2).double() - ref).norm() / ref.norm()).item()
    print("RESULT", xd, wd, "rel err", e)

if __name__ == "__main__":
    if len(sys.argv) == 3:
        run(sys.argv[1], sys.argv[2])
    else:
        for xd, wd in [("bf16", "bf16"), ("f16", "f16"), ("bf16", "f16"), ("f16", "bf16")]:
            p = subprocess.run([sys.executable, __file__, xd, wd], capture_output=True, text=True, timeout=120)
            out = [l for l in (p.stdout + p.stderr).splitlines() if "RESULT" in l or "rror" in l or "wcmc:" in l]
            print(xd, wd, "rc", p.returncode, " | ".join(out[-3:])[:400])
