"""Golden vectors for the SURVEY.md §8(f) N4 interfaces, made by running the REFERENCE's own
`KPCNRefInterface` / `KPCNPreInterface` (/root/reference/support/interfaces.py:526-750) on CPU.

Run in the authoring container only (needs /root/reference):   python tests/golden/make_golden_n4.py
Same conventions as make_golden.py: reference sources imported unmodified; `sbmc` comes from oracle/sbmc
(parity unpinned for its arithmetic, the step wiring is what these vectors pin).
Output: tests/golden/ref_golden_n4.pt (small; committed)
"""
import os
import sys
import types

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_golden import ROOT, REF, _install_stubs  # noqa: E402


def _record(itf, models, batch):
    rec = dict(losses={k: v.clone() for k, v in itf.m_losses.items()})
    rec["param_sums"] = {k: torch.stack([p.detach().double().sum() for p in m.parameters()]) for k, m in models.items()}
    rec["grad_abs_sums"] = {k: torch.stack([(p.grad.detach().double().abs().sum() if p.grad is not None
                                             else torch.tensor(-1.0, dtype=torch.float64)) for p in m.parameters()])
                            for k, m in models.items()}
    rec["training"] = {k: bool(m.training) for k, m in models.items()}
    from tests import _proj   # per-tensor fingerprints, see make_golden.py
    rec["grad_fp"] = {k: _proj.fingerprints([p.grad for p in m.parameters()]) for k, m in models.items()
                      if all(p.grad is not None for p in m.parameters())}
    rec["param_fp"] = {k: _proj.fingerprints(list(m.parameters())) for k, m in models.items()}
    return rec


def main():
    _install_stubs()
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    sys.path.insert(0, REF)
    sys.path.insert(0, ROOT)
    from support.losses import FeatureMSE, RelativeMSE
    from support.networks import PathNet
    from support.interfaces import KPCNRefInterface, KPCNPreInterface
    from sbmc import KPCN
    from wcmc_b200.synth import make_batch

    G = {}
    args = types.SimpleNamespace(model_name="golden")

    def loss_funcs(manif):
        lf = {"l_diffuse": torch.nn.L1Loss(), "l_specular": torch.nn.L1Loss(), "l_recon": torch.nn.L1Loss(),
              "l_test": RelativeMSE()}
        if manif:
            lf["l_manif"] = FeatureMSE(non_local=True)
        return lf

    # ---- KPCNRefInterface: inputs = [34 channels | reference image] ---------------------------------
    torch.manual_seed(0)
    models = {"dncnn": KPCN(37)}
    optims = {"optim_dncnn": torch.optim.Adam(models["dncnn"].parameters(), lr=1e-4)}
    itf = KPCNRefInterface(models, optims, loss_funcs(False), args)
    batch = make_batch(batch=2, spp=2, size=40, seed=31, paths=False)
    itf.to_train_mode()
    itf.preprocess(batch)
    itf.train_batch(batch)
    rec = _record(itf, models, batch)
    itf.to_eval_mode()
    with torch.no_grad():
        rad, pb = itf.validate_batch(batch)
    rec.update(val_radiance=rad.clone(), m_val=itf.m_losses["m_val"].clone(), data_seed=31, n_in=37, p_buffers_none=pb is None)
    G["ref"] = rec

    # ---- KPCNPreInterface, both stages ---------------------------------------------------------------
    for tag, manif in (("pre_manifold", True), ("pre_regress", False)):
        torch.manual_seed(0)
        models = {"dncnn": KPCN(35 + 3 + 1), "backbone_diffuse": PathNet(ic=36, outc=3),
                  "backbone_specular": PathNet(ic=36, outc=3)}
        optims = {"optim_" + k: torch.optim.Adam(m.parameters(), lr=1e-4) for k, m in models.items()}
        itf = KPCNPreInterface(models, optims, loss_funcs(True), args, manif_learn=manif, w_manif=0.1)
        batch = make_batch(batch=2, spp=2, size=40, seed=32, paths=True)
        itf.to_train_mode()
        itf.preprocess(batch)   # iters = 1 -> the manifold stage calls plt.imsave (stubbed)
        torch.manual_seed(56)
        itf.train_batch(batch)
        rec = _record(itf, models, batch)
        rec.update(data_seed=32, perm_seed=56, manif_learn=manif)
        G[tag] = rec

    out_fn = os.path.join(ROOT, "tests", "golden", "ref_golden_n4.pt")
    torch.save(G, out_fn)
    print("wrote", out_fn, os.path.getsize(out_fn), "bytes")
    for k, v in G.items():
        print(k, {a: float(b) for a, b in v["losses"].items()}, v["training"])


if __name__ == "__main__":
    main()
