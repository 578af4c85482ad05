"""Generate golden vectors by running the REFERENCE's own Python sources.

Run in the authoring container only (needs /root/reference, which does not exist on the GPU
box):   python tests/golden/make_golden.py

What is reference code here and what is not:
  * support/losses.py, support/networks.py, support/interfaces.py, support/utils.py are
    imported unmodified from /root/reference (with no-op stand-ins for the missing
    third-party imports kornia / matplotlib, which the exercised code never calls into
    except ``plt.imsave`` -> no-op).
  * ``sbmc`` (KPCN, ConvChain, Autoencoder, KernelApply) is NOT in the reference tree; the
    oracle restatement under oracle/sbmc is put on sys.path for the reference's
    ``from sbmc import modules as ops``.  Vectors that depend on it pin the *wiring*
    (PathNet.forward, KPCNInterface step) but not sbmc's arithmetic (parity unpinned).
Outputs: tests/golden/ref_golden.pt  (small; committed)
"""
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = "/root/reference"


def _install_stubs():
    kornia = types.ModuleType("kornia")

    def rgb_to_hls(x):
        raise NotImplementedError("kornia stand-in")
    kornia.rgb_to_hls = rgb_to_hls
    sys.modules["kornia"] = kornia
    mpl = types.ModuleType("matplotlib")
    plt = types.ModuleType("matplotlib.pyplot")
    plt.imsave = lambda *a, **k: None
    mpl.pyplot = plt
    sys.modules["matplotlib"] = mpl
    sys.modules["matplotlib.pyplot"] = plt


def main():
    _install_stubs()
    sys.path.insert(0, os.path.join(ROOT, "oracle"))   # oracle `sbmc`
    sys.path.insert(0, REF)                            # reference `support`
    sys.path.insert(0, ROOT)
    from support.losses import FeatureMSE, GlobalRelativeSimilarityLoss, RelativeMSE
    from support.networks import PathNet
    from support.interfaces import KPCNInterface
    from sbmc import KPCN
    from wcmc_b200.synth import make_batch
    from tests import _proj

    G = {}

    # ---- losses (fully reference arithmetic) -------------------------------------------
    g = torch.Generator().manual_seed(7)
    for tag, (b, s, c, h, w) in {"a": (2, 3, 4, 6, 5), "b": (1, 2, 3, 8, 8)}.items():
        p = torch.rand(b, s, c, h, w, generator=g)
        ref = torch.randn(b, 3, h, w, generator=g) * 1.5 + 0.5
        for non_local in (True, False):
            p_ = p.clone().requires_grad_(True)
            torch.manual_seed(100)
            loss = FeatureMSE(non_local=non_local)(p_, ref)
            loss.backward()
            G["fmse_%s_nl%d" % (tag, non_local)] = dict(p=p, ref=ref, loss=loss.detach(),
                                                         grad=p_.grad.clone(), seed=100)
        p_ = p.clone().requires_grad_(True)
        torch.manual_seed(101)
        loss = GlobalRelativeSimilarityLoss()(p_, ref)
        loss.backward()
        G["grs_%s" % tag] = dict(p=p, ref=ref, loss=loss.detach(), grad=p_.grad.clone(), seed=101)
    im = torch.rand(2, 3, 7, 9, generator=g) * 3
    rf = torch.rand(2, 3, 7, 9, generator=g) * 3
    G["relmse"] = dict(im=im, ref=rf, loss=RelativeMSE()(im, rf))

    # ---- PathNet wiring -----------------------------------------------------------------
    torch.manual_seed(0)
    net = PathNet(ic=36, outc=3)
    batch = make_batch(batch=1, spp=2, size=16, seed=11)
    with torch.no_grad():
        out = net(batch)
    G["pathnet"] = dict(seed=0, data_seed=11, out=out, str=str(net),
                        keys=sorted(net.state_dict().keys()))

    # ---- KPCNInterface.train_batch / validate_batch (reference step orchestration) -------
    for tag, cfg in {
        "wcmc": dict(use_llpm_buf=True, manif_learn=True, opt="m11r11", outc=3),
        "wcmc_m10r01": dict(use_llpm_buf=True, manif_learn=True, opt="m10r01", outc=4),
        "vanilla": dict(use_llpm_buf=False, manif_learn=False, opt="m11r11", outc=3),
    }.items():
        torch.manual_seed(0)
        llpm = cfg["use_llpm_buf"]
        c_reg = cfg["outc"] // 2 if cfg["opt"] in ("m10r01", "m11r01") else cfg["outc"]
        n_in = 35 + c_reg + 1 if llpm else 34
        models = {"dncnn": KPCN(n_in)}
        if llpm:
            models["backbone_diffuse"] = PathNet(ic=36, outc=cfg["outc"])
            models["backbone_specular"] = PathNet(ic=36, outc=cfg["outc"])
        optims = {"optim_" + k: torch.optim.Adam(m.parameters(), lr=1e-4) for k, m in models.items()}
        loss_funcs = {"l_diffuse": torch.nn.L1Loss(), "l_specular": torch.nn.L1Loss(),
                      "l_recon": torch.nn.L1Loss(), "l_test": RelativeMSE()}
        if cfg["manif_learn"]:
            loss_funcs["l_manif"] = FeatureMSE(non_local=True)
        args = types.SimpleNamespace(model_name="golden")
        itf = KPCNInterface(models, optims, loss_funcs, args, use_llpm_buf=llpm,
                            manif_learn=cfg["manif_learn"], w_manif=0.1, train_branches=True,
                            disentanglement_option=cfg["opt"])
        batch = make_batch(batch=2, spp=2, size=40, seed=21, paths=llpm)
        itf.to_train_mode()
        itf.preprocess(batch)
        torch.manual_seed(55)
        itf.train_batch(batch)
        rec = dict(cfg=cfg, n_in=n_in, data_seed=21, perm_seed=55,
                   losses={k: v.clone() for k, v in itf.m_losses.items()})
        rec["param_sums"] = {k: torch.stack([p.detach().double().sum() for p in m.parameters()])
                             for k, m in models.items()}
        rec["grad_abs_sums"] = {k: torch.stack([p.grad.detach().double().abs().sum()
                                                for p in m.parameters()])
                                for k, m in models.items()}
        # per-tensor fingerprints (tests/_proj.py): K sign projections + L2 norm of every gradient (left CLIPPED by
        # the step, interfaces.py:261) and of every parameter after the Adam step
        rec["grad_fp"] = {k: _proj.fingerprints([p.grad for p in m.parameters()]) for k, m in models.items()}
        rec["param_fp"] = {k: _proj.fingerprints(list(m.parameters())) for k, m in models.items()}
        rec["param_names"] = {k: [n for n, _ in m.named_parameters()] for k, m in models.items()}
        itf.to_eval_mode()
        with torch.no_grad():
            rad, pb = itf.validate_batch(batch)
        rec["val_radiance"] = rad.clone()
        rec["m_val"] = itf.m_losses["m_val"].clone()
        G["itf_" + tag] = rec

    out_fn = os.path.join(ROOT, "tests", "golden", "ref_golden.pt")
    torch.save(G, out_fn)
    print("wrote", out_fn, os.path.getsize(out_fn), "bytes")


if __name__ == "__main__":
    main()
