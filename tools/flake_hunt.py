"""Hunts the intermittent "Infinite loss at train time." of long bench runs: the graphed north-star step driven (a) from
resident inputs and (b) through the DevicePrefetcher from pinned host memory, weights / moments / step count put back
every RESET steps as bench.py does, check="immediate" so that the failing step is known; on a failure prints the step,
the loss terms and which finite flag tripped, then goes on with a restored state.

    python tools/flake_hunt.py [steps per phase] [reset]
"""
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from wcmc_b200 import dropin, lib  # noqa: E402
from wcmc_b200.synth import make_batch  # noqa: E402

dropin.install()
lib.init()
from sbmc import KPCN  # noqa: E402
from support.interfaces import KPCNInterface  # noqa: E402
from support.losses import FeatureMSE, RelativeMSE  # noqa: E402
from support.networks import PathNet  # noqa: E402

from wcmc_b200.engine import DevicePrefetcher, GraphedTrainStep  # noqa: E402


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 300
    reset = int(sys.argv[2]) if len(sys.argv) > 2 else 100
    torch.manual_seed(0)
    models = {"dncnn": KPCN(39).cuda(), "backbone_diffuse": PathNet(ic=36, outc=3).cuda(),
              "backbone_specular": PathNet(ic=36, outc=3).cuda()}
    optims = {"optim_" + k: torch.optim.Adam(m.parameters(), lr=1e-4) for k, m in models.items()}
    lf = {"l_diffuse": torch.nn.L1Loss(), "l_specular": torch.nn.L1Loss(), "l_recon": torch.nn.L1Loss(),
          "l_test": RelativeMSE(), "l_manif": FeatureMSE(non_local=True, rng="device")}
    itf = KPCNInterface(models, optims, lf, types.SimpleNamespace(model_name="hunt"), use_llpm_buf=True,
                        manif_learn=True, w_manif=0.1, train_branches=True, disentanglement_option="m11r11")
    host = {k: v.pin_memory() for k, v in make_batch(batch=8, spp=8, size=128, seed=1234).items()}
    dev = {k: v.cuda() for k, v in host.items()}
    itf.to_train_mode()
    g = GraphedTrainStep(itf, dev, check="immediate")
    for _ in range(5):
        g(dev)
    live = [p for m in models.values() for p in m.parameters()]
    for o in optims.values():
        for st in o.state.values():
            live += [st["exp_avg"], st["exp_avg_sq"]]
    init = [t.detach().clone() for t in live]
    t0 = g.fused.t

    def restore():
        with torch.no_grad():
            torch._foreach_copy_(live, init)
        g.fused.set_step(t0)
        itf.m_losses.clear()

    def report(phase, i, since):
        torch.cuda.synchronize()
        terms = {k: float(v) for k, v in g.loss.items()}
        wmax = max(float(p.abs().max()) for p in live[:len(list(models["dncnn"].parameters()))])
        print("FAIL %s step %d (%d since restore): losses %s flag %s max|w dncnn| %.3g nonfinite grads %d" % (
            phase, i, since, terms, bool(g.flags), wmax, g.fused.nonfinite_count()), flush=True)

    for phase in ("resident", "prefetch", "resident", "prefetch"):
        restore()
        fails, since = 0, 0
        pf = DevicePrefetcher(iter(lambda: host, None)) if phase == "prefetch" else None
        for i in range(steps):
            if since == reset:
                restore()
                since = 0
            since += 1
            try:
                if pf is None:
                    g(g.static)
                else:
                    b = next(pf)
                    g(b)
                    pf.release()
            except RuntimeError as e:
                if "Infinite loss" not in str(e):
                    raise
                fails += 1
                report(phase, i, since)
                restore()
                since = 0
        torch.cuda.synchronize()
        print("phase %-9s %d steps, reset every %d: %d failures, running loss %.4f" % (
            phase, steps, reset, fails, float(itf.m_losses.get("m_l_total", torch.zeros(()))) / max(1, since)), flush=True)


if __name__ == "__main__":
    main()
