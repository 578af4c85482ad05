"""Kernel timeline of the CUDA-graph-replayed KPCN+WCMC step (what bench.py's `value` times), from torch.profiler
(CUPTI activity records: true start / duration / stream of every kernel inside a replay).  Prints per step: span,
per-stream busy time, time with 0 / 1 / >= 2 kernels in flight, the largest idle gaps with their neighbours, and the
per-kernel totals; writes the raw kernel list to gpurun_out/timeline_kernels.json.

    python tools/step_timeline.py [replays]
"""
import json
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

from wcmc_b200 import dropin, lib  # noqa: E402
from wcmc_b200.synth import make_batch  # noqa: E402

dropin.install()
lib.init()
from sbmc import KPCN  # noqa: E402
from support.interfaces import KPCNInterface  # noqa: E402
from support.losses import FeatureMSE, RelativeMSE  # noqa: E402
from support.networks import PathNet  # noqa: E402

from wcmc_b200.engine import GraphedTrainStep  # noqa: E402


def main():
    replays = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    torch.manual_seed(0)
    models = {"dncnn": KPCN(39).cuda(), "backbone_diffuse": PathNet(ic=36, outc=3).cuda(),
              "backbone_specular": PathNet(ic=36, outc=3).cuda()}
    optims = {"optim_" + k: torch.optim.Adam(m.parameters(), lr=1e-4) for k, m in models.items()}
    lf = {"l_diffuse": torch.nn.L1Loss(), "l_specular": torch.nn.L1Loss(), "l_recon": torch.nn.L1Loss(),
          "l_test": RelativeMSE(), "l_manif": FeatureMSE(non_local=True, rng="device")}
    itf = KPCNInterface(models, optims, lf, types.SimpleNamespace(model_name="timeline"), use_llpm_buf=True,
                        manif_learn=True, w_manif=0.1, train_branches=True, disentanglement_option="m11r11")
    dev = {k: v.cuda() for k, v in make_batch(batch=8, spp=8, size=128, seed=1234).items()}
    itf.to_train_mode()
    step = GraphedTrainStep(itf, dev)
    for _ in range(5):
        step(dev)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(replays):
            step(dev)
        torch.cuda.synchronize()
    out_dir = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out_dir, exist_ok=True)
    trace = os.path.join(out_dir, "timeline_trace.json")
    prof.export_chrome_trace(trace)
    ks = []
    for e in json.load(open(trace))["traceEvents"]:
        if e.get("ph") == "X" and e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset"):
            ks.append(dict(name=e["name"], start=float(e["ts"]), end=float(e["ts"]) + float(e["dur"]),
                           stream=e.get("args", {}).get("stream", e.get("tid")), kind=e["cat"]))
    os.remove(trace)
    ks.sort(key=lambda k: k["start"])
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "timeline_kernels.json"), "w") as f:
        json.dump(ks, f)
    # split into replays at the adam kernel
    cuts = [i for i, k in enumerate(ks) if "adam_clip_kernel" in k["name"]]
    print("kernels", len(ks), "replays found", len(cuts))
    if len(cuts) < 2:
        return
    seg = ks[cuts[-2] + 1:cuts[-1] + 1]
    t0, t1 = seg[0]["start"], max(k["end"] for k in seg)
    print("last replay: %d kernels, span %.1f us" % (len(seg), t1 - t0))
    streams = sorted({k["stream"] for k in seg})
    for s_ in streams:
        mine = [k for k in seg if k["stream"] == s_]
        print("  stream %s: %d kernels, busy %.1f us, first %.1f last %.1f" % (
            s_, len(mine), sum(k["end"] - k["start"] for k in mine), mine[0]["start"] - t0, max(k["end"] for k in mine) - t0))
    # concurrency profile by sweeping
    ev = []
    for k in seg:
        ev.append((k["start"], 1))
        ev.append((k["end"], -1))
    ev.sort()
    depth, last, hist = 0, t0, {}
    for t, d in ev:
        hist[min(depth, 2)] = hist.get(min(depth, 2), 0.0) + (t - last)
        depth += d
        last = t
    print("  time with 0 / 1 / >=2 kernels in flight: %.1f / %.1f / %.1f us" % (hist.get(0, 0), hist.get(1, 0), hist.get(2, 0)))
    # idle gaps (nothing in flight)
    gaps, depth, last_end, prev = [], 0, None, None
    cur_end = t0
    order = sorted(seg, key=lambda k: k["start"])
    for k in order:
        if k["start"] > cur_end:
            gaps.append((k["start"] - cur_end, prev["name"][:50] if prev else "-", k["name"][:50], cur_end - t0))
        if k["end"] > cur_end:
            cur_end, prev = k["end"], k
    gaps.sort(reverse=True)
    print("  idle gaps: %d, total %.1f us; largest:" % (len(gaps), sum(g[0] for g in gaps)))
    for g in gaps[:25]:
        print("    %6.1f us at %7.1f  after %-50s before %s" % (g[0], g[3], g[1], g[2]))
    tot = {}
    for k in seg:
        n = k["name"].split("(")[0][:70]
        a = tot.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += k["end"] - k["start"]
    print("  per kernel (count, total us):")
    for n, (c, t) in sorted(tot.items(), key=lambda x: -x[1][1])[:40]:
        print("    %4d %9.1f  %s" % (c, t, n))
    # the sequence with times, for reading the critical path
    with open(os.path.join(ROOT, "gpurun_out", "timeline_last_replay.txt"), "w") as f:
        for k in order:
            f.write("%9.1f %8.1f  s%-4s %s\n" % (k["start"] - t0, k["end"] - k["start"], k["stream"], k["name"][:90]))


if __name__ == "__main__":
    main()
