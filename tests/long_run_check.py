"""Long-run stability check (test infrastructure; uses the oracle as the fp32 yardstick): the bench's synthetic batch is
trained on repeatedly for N steps by (a) the B200 backend (CUDA-graph step) and (b) the fp32 oracle on the same GPU;
prints the loss trajectory of both and the largest activation magnitudes the backend's 16-bit tensors would hold.
    python tests/long_run_check.py [steps] [ours|oracle|both]"""
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 700
which = sys.argv[2] if len(sys.argv) > 2 else "both"
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
from wcmc_b200 import dropin, lib  # noqa: E402
from wcmc_b200.synth import make_batch  # noqa: E402
dropin.install()
lib.init(0)
from sbmc import KPCN  # noqa: E402
from support.interfaces import KPCNInterface  # noqa: E402
from support.losses import FeatureMSE, RelativeMSE  # noqa: E402
from support.networks import PathNet  # noqa: E402
from tests._oracle_loader import load_oracle  # noqa: E402
oracle = load_oracle()


def make(KP, PN):
    torch.manual_seed(0)
    return {"dncnn": KP(39).cuda(), "backbone_diffuse": PN(ic=36, outc=3).cuda(), "backbone_specular": PN(ic=36, outc=3).cuda()}


batch = {k: v.cuda() for k, v in make_batch(batch=8, spp=8, size=128, seed=1234).items()}
if which in ("ours", "both"):
    models = make(KPCN, PathNet)
    optims = {"optim_" + k: torch.optim.Adam(m.parameters(), lr=1e-4) for k, m in models.items()}
    lf = {"l_diffuse": torch.nn.L1Loss(), "l_specular": torch.nn.L1Loss(), "l_recon": torch.nn.L1Loss(),
          "l_test": RelativeMSE(), "l_manif": FeatureMSE(non_local=True, rng="device")}
    itf = KPCNInterface(models, optims, lf, types.SimpleNamespace(model_name="long"), use_llpm_buf=True, manif_learn=True,
                        w_manif=0.1, train_branches=True)
    itf.to_train_mode()
    from wcmc_b200.engine import GraphedTrainStep
    step = GraphedTrainStep(itf, batch)
    for i in range(steps):
        try:
            loss = step(batch)
        except RuntimeError as e:
            print("ours: step %d raised %s; last losses %s" % (i, e, {k: float(v) for k, v in step.loss.items()}))
            wmax = {n: max(float(p.detach().abs().max()) for p in m.parameters()) for n, m in models.items()}
            print("ours: max |weight| per model", wmax)
            break
        if i % 50 == 0 or i == steps - 1:
            print("ours   step %4d  %s" % (i, {k: round(float(v), 5) for k, v in loss.items()}), flush=True)
if which in ("oracle", "both"):
    ref = make(oracle.KPCN, oracle.PathNet)
    ropt = {"optim_" + k: torch.optim.Adam(m.parameters(), lr=1e-4) for k, m in ref.items()}
    for i in range(steps):
        loss, _, _ = oracle.ref.kpcn_train_step(ref, ropt, batch, use_llpm_buf=True, manif_learn=True, w_manif=0.1)
        if i % 50 == 0 or i == steps - 1:
            print("oracle step %4d  %s" % (i, {k: round(float(v), 5) for k, v in loss.items()}), flush=True)
