"""GPU bring-up probe: runs every kernel case in its own subprocess (a trapped kernel kills the
CUDA context) and writes one JSON line per case to gpurun_out/probe.jsonl.

    python tools/gpu_probe.py            # all cases
    python tools/gpu_probe.py --case conv:5:112:112:1:2:0   # single case (internal)
"""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def rel(a, b):
    import torch
    return ((a.double() - b.double()).norm() / (b.double().norm() + 1e-30)).item()


def case_conv(k, cin, cout, pad_same, mt, boff, n=2, h=40, w=36, fp32=0, nt=0):
    import torch
    import torch.nn.functional as F
    from wcmc_b200 import lib
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    g = torch.Generator(device="cuda").manual_seed(1)
    pad = k // 2 if pad_same else 0
    x = torch.randn(n, cin, h, w, device="cuda", generator=g)
    wt = torch.randn(cout, cin, k, k, device="cuda", generator=g) / (cin * k * k) ** 0.5
    b = torch.randn(cout, device="cuda", generator=g)
    xb = x.bfloat16().float()
    wb = wt.bfloat16().float()
    ref = F.relu(F.conv2d(xb, wb, b, padding=pad))
    cin_p, cout_p = lib.pad16(cin), lib.pad16(cout)
    xn = lib.nchw_to_nhwc(x)
    wf, wd = lib.pack_weights(wt)
    bp = torch.zeros(cout_p, device="cuda")
    bp[:cout] = b
    flags = boff | (mt << 4) | (nt << 8)
    y = lib.conv2d(xn, wf, bp, k, pad, act=1, out_fp32=bool(fp32), flags=flags)
    torch.cuda.synchronize()
    if fp32:
        got = y[..., :cout].permute(0, 3, 1, 2)
    else:
        got = lib.nhwc_to_nchw(y, cout)
    err = rel(got, ref)
    pad_ok = bool((y[..., cout:].float().abs().max() == 0).item()) if cout_p > cout else True
    # layout round trip sanity
    rt = rel(lib.nhwc_to_nchw(xn, cin), xb)
    out = dict(err=err, pad_zero=pad_ok, roundtrip=rt)
    if err > 2e-2:
        d = (got - ref).abs()
        idx = d.flatten().argmax().item()
        out["worst"] = [int(v) for v in torch.unravel_index(torch.tensor(idx), d.shape)]
        out["per_channel_err"] = [round(rel(got[:, c], ref[:, c]), 4) for c in range(min(cout, 8))]
        # error by output row / col within the first image (find tile-pattern bugs)
        out["row_err"] = [round(rel(got[0, :, r], ref[0, :, r]), 3) for r in range(min(got.shape[2], 36))]
        out["col_err"] = [round(rel(got[0, :, :, c], ref[0, :, :, c]), 3) for c in range(min(got.shape[3], 36))]
    return out


def case_dgrad(k, cin, cout, pad_same, n=2, h=40, w=36):
    import torch
    import torch.nn.functional as F
    from wcmc_b200 import lib
    torch.backends.cudnn.allow_tf32 = False
    g = torch.Generator(device="cuda").manual_seed(2)
    pad = k // 2 if pad_same else 0
    ho, wo = h + 2 * pad - k + 1, w + 2 * pad - k + 1
    wt = torch.randn(cout, cin, k, k, device="cuda", generator=g) / (cout * k * k) ** 0.5
    dy = torch.randn(n, cout, ho, wo, device="cuda", generator=g)
    a_prev = torch.randn(n, cin, h, w, device="cuda", generator=g)   # activation whose relu mask applies
    x = torch.zeros(n, cin, h, w, device="cuda", requires_grad=True)
    y = F.conv2d(x, wt.bfloat16().float(), None, padding=pad)
    (gx,) = torch.autograd.grad(y, x, dy.bfloat16().float())
    ref = gx * (a_prev.bfloat16().float() > 0)
    wf, wd = lib.pack_weights(wt)
    dyn = lib.nchw_to_nhwc(dy)
    mask = lib.nchw_to_nhwc(a_prev)
    dx = lib.conv2d(dyn, wd, None, k, k - 1 - pad, act=0, mask=mask, slope=0.0)
    torch.cuda.synchronize()
    got = lib.nhwc_to_nchw(dx, cin)
    return dict(err=rel(got, ref))


def case_wgrad(k, cin, cout, pad_same, n=2, h=40, w=36):
    import torch
    import torch.nn.functional as F
    from wcmc_b200 import lib
    torch.backends.cudnn.allow_tf32 = False
    g = torch.Generator(device="cuda").manual_seed(3)
    pad = k // 2 if pad_same else 0
    ho, wo = h + 2 * pad - k + 1, w + 2 * pad - k + 1
    x = torch.randn(n, cin, h, w, device="cuda", generator=g)
    dy = torch.randn(n, cout, ho, wo, device="cuda", generator=g)
    wt = torch.zeros(cout, cin, k, k, device="cuda", requires_grad=True)
    y = F.conv2d(x.bfloat16().float(), wt, None, padding=pad)
    (ref,) = torch.autograd.grad(y, wt, dy.bfloat16().float())
    xn = lib.nchw_to_nhwc(x)
    dyn = lib.nchw_to_nhwc(dy)
    got = lib.conv2d_wgrad(xn, dyn, cout, cin, k, pad, lib.pad16(cin), lib.pad16(cout))
    db = lib.bias_grad(dyn, cout)
    torch.cuda.synchronize()
    out = dict(err=rel(got, ref), bias_err=rel(db, dy.bfloat16().float().sum((0, 2, 3))))
    if out["err"] > 2e-2:
        out["per_tap_err"] = [round(rel(got[:, :, t // k, t % k], ref[:, :, t // k, t % k]), 3) for t in range(k * k)]
        out["per_co_err"] = [round(rel(got[c], ref[c]), 3) for c in range(0, cout, max(1, cout // 16))]
        out["per_ci_err"] = [round(rel(got[:, c], ref[:, c]), 3) for c in range(0, cin, max(1, cin // 16))]
    return out


def case_ka(k, c, n=2, h=37, w=45):
    import torch
    import torch.nn.functional as F
    from wcmc_b200 import lib
    g = torch.Generator(device="cuda").manual_seed(4)
    taps = k * k
    cs = (taps + 7) // 8 * 8
    logits = torch.randn(n, h, w, cs, device="cuda", generator=g) * 2
    data = torch.rand(n, c, h, w, device="cuda", generator=g) * 3
    gout = torch.randn(n, c, h, w, device="cuda", generator=g)
    z = logits[..., :taps].permute(0, 3, 1, 2).contiguous().double().requires_grad_(True)
    p = F.softmax(z, dim=1)
    unf = F.unfold(data.double(), (k, k), padding=k // 2).view(n, c, taps, h, w)
    ref = (unf * p.unsqueeze(1)).sum(2)
    (gz,) = torch.autograd.grad(ref, z, gout.double())
    out, stats = lib.kernel_apply_fwd(logits, data, k)
    dl32 = lib.kernel_apply_bwd(logits, data, out, stats, gout, k, bf16=False)
    dl16 = lib.kernel_apply_bwd(logits, data, out, stats, gout, k, bf16=True)
    torch.cuda.synchronize()
    gzn = gz.permute(0, 2, 3, 1)
    res = dict(fwd_err=rel(out, ref), bwd_err_f32=rel(dl32[..., :taps], gzn), bwd_err_bf16=rel(dl16[..., :taps].float(), gzn))
    if cs > taps:
        res["pad_zero"] = bool((dl32[..., taps:].abs().max() == 0).item())
    return res


def run_case(spec):
    parts = spec.split(":")
    kind, args = parts[0], [int(v) for v in parts[1:]]
    fn = dict(conv=case_conv, dgrad=case_dgrad, wgrad=case_wgrad, ka=case_ka)[kind]
    return fn(*args)


CASES = [
    # conv: k, cin, cout, pad_same, mt, base_offset_mode
    "conv:1:64:64:0:2:0",      # canonical descriptors only (pitch multiple of 1024)
    "conv:1:64:64:0:1:0",
    "conv:3:64:64:1:2:0",      # shifted windows, halo pitch 18*128
    "conv:3:64:64:1:2:1",
    "conv:3:64:64:1:1:0",
    "conv:3:64:64:1:1:1",
    "conv:5:112:112:0:2:0",
    "conv:5:112:112:0:2:1",
    "conv:5:100:100:0:1:0",
    "conv:5:34:100:0:2:0",
    "conv:5:100:441:0:2:0:2:40:36:1",   # fp32 logits layer
    "conv:3:192:64:1:2:0",
    "conv:3:384:128:1:1:0",
    "conv:3:128:256:1:2:0",
    "conv:1:36:64:0:2:0",
    "conv:1:128:16:0:2:0",
    "dgrad:5:100:100:0",
    "dgrad:5:39:100:0",
    "dgrad:3:64:128:1",
    "dgrad:1:64:64:0",
    "wgrad:1:64:64:0",
    "wgrad:3:64:64:1",
    "wgrad:5:100:100:0",
    "wgrad:5:39:100:0",
    "wgrad:5:100:441:0",
    "wgrad:3:384:128:1",
    "wgrad:1:36:64:0",
    "ka:5:3",
    "ka:21:3",
    "ka:21:1",
    "ka:3:4",
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--case")
    ap.add_argument("--only", default="")
    a = ap.parse_args()
    if a.case:
        try:
            res = run_case(a.case)
            print("PROBE_RESULT " + json.dumps(dict(case=a.case, ok=True, **res)))
        except Exception as e:  # noqa: BLE001
            print("PROBE_RESULT " + json.dumps(dict(case=a.case, ok=False, error=repr(e)[:600])))
        return
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    out = open(os.path.join(ROOT, "gpurun_out", "probe.jsonl"), "a")
    for spec in CASES:
        if a.only and not spec.startswith(a.only):
            continue
        t0 = time.time()
        try:
            pr = subprocess.run([sys.executable, os.path.abspath(__file__), "--case", spec], capture_output=True,
                                text=True, timeout=120)
            line = [ln for ln in pr.stdout.splitlines() if ln.startswith("PROBE_RESULT ")]
            if line:
                rec = json.loads(line[-1][len("PROBE_RESULT "):])
            else:
                rec = dict(case=spec, ok=False, error="no result; rc=%d" % pr.returncode,
                           stdout=pr.stdout[-600:], stderr=pr.stderr[-1200:])
        except subprocess.TimeoutExpired:
            rec = dict(case=spec, ok=False, error="timeout")
        rec["sec"] = round(time.time() - t0, 1)
        print(json.dumps(rec))
        out.write(json.dumps(rec) + "\n")
        out.flush()


if __name__ == "__main__":
    main()
