"""CPU: the drop-in step orchestration (wcmc_b200/dropin/support/interfaces.py) against the REFERENCE's own classes
(/root/reference/support/interfaces.py: KPCNInterface :80-333, KPCNRefInterface :526-585, KPCNPreInterface :588-750),
both driven with the same stand-in torch modules (the CUDA kernels are not involved): same losses, same gradients, same
parameters after the Adam step, same train / eval modes, same validate_batch outputs.  Needs the reference checkout
(present where the CPU tests run; skipped on the GPU box)."""
import importlib.util
import os
import sys
import types

import pytest
import torch

REF = "/root/reference/support/interfaces.py"
pytestmark = pytest.mark.skipif(not os.path.exists(REF), reason="reference checkout not present")

from tests.test_ddp_gloo import _TinyKPCN, _TinyPathNet  # noqa: E402


class _TinyManifLoss(torch.nn.Module):
    """Stand-in for FeatureMSE (a CUDA kernel in the product): same call contract, any channel count."""

    def forward(self, p_buffer, ref):
        return (p_buffer.mean(1).mean(1, keepdim=True) - ref.mean(1, keepdim=True)).pow(2).mean() + 0.1 * p_buffer.pow(2).mean()


@pytest.fixture(scope="module")
def both():
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    from make_golden import _install_stubs
    _install_stubs()                      # matplotlib.pyplot.imsave -> no-op (interfaces.py:130-137 writes PNGs)
    from wcmc_b200 import dropin
    dropin.install()
    import support.interfaces as ours
    spec = importlib.util.spec_from_file_location("wcmc_reference_interfaces", REF)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    from support.losses import RelativeMSE
    return types.SimpleNamespace(ours=ours, ref=ref, RelativeMSE=RelativeMSE)


def _models(n_in, llpm, outc=3):
    torch.manual_seed(0)
    m = {"dncnn": _TinyKPCN(n_in)}
    if llpm:
        m["backbone_diffuse"] = _TinyPathNet(36, outc)
        m["backbone_specular"] = _TinyPathNet(36, outc)
    return m


def _loss_funcs(both, manif):
    lf = {"l_diffuse": torch.nn.L1Loss(), "l_specular": torch.nn.L1Loss(), "l_recon": torch.nn.L1Loss(),
          "l_test": both.RelativeMSE()}
    if manif:
        lf["l_manif"] = _TinyManifLoss()
    return lf


def _run(cls, models, lf, batch, **kw):
    optims = {"optim_" + k: torch.optim.Adam(m.parameters(), lr=1e-3) for k, m in models.items()}
    itf = cls(models, optims, lf, types.SimpleNamespace(model_name="t"), **kw)
    itf.to_train_mode()
    modes = {k: m.training for k, m in models.items()}
    itf.preprocess(batch)
    itf.train_batch(batch)
    losses = {k: v.clone() for k, v in itf.m_losses.items()}
    grads = {k: [None if p.grad is None else p.grad.clone() for p in m.parameters()] for k, m in models.items()}
    params = {k: [p.detach().clone() for p in m.parameters()] for k, m in models.items()}
    itf.to_eval_mode()
    with torch.no_grad():
        rad, pb = itf.validate_batch(batch)
    return dict(losses=losses, grads=grads, params=params, modes=modes, rad=rad, pb=pb, m_val=itf.m_losses["m_val"].clone(),
                summary=itf.get_epoch_summary("eval", 1.0))


def _same(a, b):
    assert list(a["losses"]) == list(b["losses"])          # same keys in the same order
    for k in a["losses"]:
        torch.testing.assert_close(a["losses"][k], b["losses"][k], rtol=1e-6, atol=1e-8, msg=k)
    assert a["modes"] == b["modes"]
    for name in a["params"]:
        for x, y in zip(a["params"][name], b["params"][name]):
            torch.testing.assert_close(x, y, rtol=1e-6, atol=1e-8)
        for x, y in zip(a["grads"][name], b["grads"][name]):
            assert (x is None) == (y is None), name
            if x is not None:
                torch.testing.assert_close(x, y, rtol=1e-5, atol=1e-9)
    torch.testing.assert_close(a["rad"], b["rad"], rtol=1e-6, atol=1e-8)
    torch.testing.assert_close(a["m_val"], b["m_val"], rtol=1e-6, atol=1e-8)
    assert abs(a["summary"] - b["summary"]) < 1e-7
    if a["pb"] is None:
        assert b["pb"] is None
    else:
        assert sorted(a["pb"]) == sorted(b["pb"])
        for k in a["pb"]:
            torch.testing.assert_close(a["pb"][k], b["pb"][k], rtol=1e-6, atol=1e-8)


@pytest.mark.parametrize("cfg", [dict(use_llpm_buf=True, manif_learn=True, disentanglement_option="m11r11"),
                                 dict(use_llpm_buf=True, manif_learn=True, disentanglement_option="m10r01"),
                                 dict(use_llpm_buf=True, manif_learn=True, disentanglement_option="m11r01"),
                                 dict(use_llpm_buf=True, manif_learn=True, disentanglement_option="m10r11"),
                                 dict(use_llpm_buf=True, manif_learn=False),
                                 dict(use_llpm_buf=False, manif_learn=False),
                                 dict(use_llpm_buf=False, manif_learn=False, train_branches=False)],
                         ids=lambda c: "-".join("%s=%s" % kv for kv in c.items()))
def test_kpcn_interface_matches_reference_class(both, cfg):
    from wcmc_b200.synth import make_batch
    llpm = cfg["use_llpm_buf"]
    outc = 4
    half = cfg.get("disentanglement_option", "m11r11") in ("m10r01", "m11r01")
    n_in = 34 + (1 + (outc // 2 if half else outc) + 1 if llpm else 0)
    batch = make_batch(batch=2, spp=2, size=24, seed=3, paths=llpm)
    kw = dict(cfg, w_manif=0.1)
    ref = _run(both.ref.KPCNInterface, _models(n_in, llpm, outc), _loss_funcs(both, cfg["manif_learn"]), batch, **kw)
    ours = _run(both.ours.KPCNInterface, _models(n_in, llpm, outc), _loss_funcs(both, cfg["manif_learn"]), batch, **kw)
    _same(ours, ref)


def test_ref_interface_matches_reference_class(both):
    from wcmc_b200.synth import make_batch
    batch = make_batch(batch=2, spp=2, size=24, seed=4, paths=False)
    ref = _run(both.ref.KPCNRefInterface, _models(37, False), _loss_funcs(both, False), batch)
    ours = _run(both.ours.KPCNRefInterface, _models(37, False), _loss_funcs(both, False), batch)
    _same(ours, ref)


@pytest.mark.parametrize("manif_learn", [True, False])
def test_pre_interface_matches_reference_class(both, manif_learn):
    from wcmc_b200.synth import make_batch
    batch = make_batch(batch=2, spp=2, size=24, seed=5, paths=True)
    ref = _run(both.ref.KPCNPreInterface, _models(39, True), _loss_funcs(both, True), batch, manif_learn=manif_learn)
    ours = _run(both.ours.KPCNPreInterface, _models(39, True), _loss_funcs(both, True), batch, manif_learn=manif_learn)
    _same(ours, ref)
    frozen = ["dncnn"] if manif_learn else ["backbone_diffuse", "backbone_specular"]
    fresh = _models(39, True)
    for k in frozen:    # the stage's frozen side keeps its initial weights in both implementations
        for p0, p1 in zip(fresh[k].parameters(), ours["params"][k]):
            assert torch.equal(p0.detach(), p1)


def test_error_behaviour_matches_reference(both):
    """Constructor assertions and the non-finite-loss exception (interfaces.py:84-95, :255-257)."""
    lf = _loss_funcs(both, False)
    for mod in (both.ref, both.ours):
        with pytest.raises(AssertionError):
            mod.KPCNInterface({"backbone_diffuse": _TinyPathNet(36, 3)}, {}, lf, None)          # no 'dncnn'
        with pytest.raises(AssertionError):
            mod.KPCNInterface({"dncnn": _TinyKPCN(34)}, {}, lf, None, manif_learn=True)         # no backbones
        with pytest.raises(AssertionError):
            mod.KPCNInterface({"dncnn": _TinyKPCN(34)}, {}, {"l_recon": lf["l_recon"], "l_test": lf["l_test"]}, None)
        with pytest.raises(AssertionError):
            mod.KPCNRefInterface({"dncnn": _TinyKPCN(37)}, {}, lf, None, use_llpm_buf=True)
    from wcmc_b200.synth import make_batch
    batch = make_batch(batch=1, spp=2, size=24, seed=6, paths=False)
    batch["target_diffuse"] = batch["target_diffuse"].clone()
    batch["target_diffuse"][0, 0, 12, 12] = float("nan")
    for mod in (both.ref, both.ours):
        models = _models(34, False)
        optims = {"optim_dncnn": torch.optim.Adam(models["dncnn"].parameters(), lr=1e-3)}
        itf = mod.KPCNInterface(models, optims, lf, types.SimpleNamespace(model_name="t"))
        itf.to_train_mode()
        itf.preprocess(batch)
        with pytest.raises(RuntimeError, match="Non-finite loss at train time"):
            itf.train_batch(batch)


def test_train_kpcn_imports_unchanged_against_the_dropin():
    """`import train_kpcn` (the reference's training script, unmodified) resolves every hot-path import to the drop-in
    packages (train_kpcn.py:19-33: PathNet, the three losses, KPCNInterface / KPCNRefInterface / KPCNPreInterface,
    sbmc.KPCN, ttools crop_like) while datasets / utils fall through to the reference checkout.  Own process: the
    script's imports must not leak into the other tests."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = r'''
import sys, types
for name in ("visdom", "kornia"):
    sys.modules[name] = types.ModuleType(name)
mpl = types.ModuleType("matplotlib"); plt = types.ModuleType("matplotlib.pyplot"); plt.imsave = lambda *a, **k: None
mpl.pyplot = plt; sys.modules["matplotlib"] = mpl; sys.modules["matplotlib.pyplot"] = plt
sys.path.insert(0, "/root/reference")
sys.path.insert(0, %r)
from wcmc_b200 import dropin
dropin.install()
import train_kpcn as T
import support.interfaces as I, support.networks as N, support.losses as L, sbmc
drop = dropin.__file__.rsplit("/", 1)[0]
for obj in (T.KPCNInterface, T.KPCNRefInterface, T.KPCNPreInterface, T.PathNet, T.FeatureMSE, T.GlobalRelativeSimilarityLoss,
            T.RelativeMSE, T.KPCN, T.crop_like):
    mod = sys.modules[obj.__module__]
    assert mod.__file__.startswith(drop), (obj, mod.__file__)
assert T.MSDenoiseDataset.__module__ == "support.datasets" and sys.modules["support.datasets"].__file__.startswith("/root/reference")
assert T.BasicArgumentParser.__module__ == "support.utils"
assert issubclass(T.KPCNRefInterface, T.KPCNInterface) and issubclass(T.KPCNPreInterface, T.KPCNInterface)
assert callable(T.train_epoch_kpcn) and callable(T.validate_kpcn) and callable(T.train)
print("OK")
''' % root
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300, cwd=root)
    assert r.returncode == 0 and r.stdout.strip().endswith("OK"), (r.stdout[-500:], r.stderr[-2000:])


def test_dropin_utils_match_reference_utils():
    """support/utils.py of the drop-in shadows the reference's: same CLI flags on BasicArgumentParser (utils.py:70-100),
    same tone mappers (:44-67), same crop rule (:24-42)."""
    import numpy as np
    spec = importlib.util.spec_from_file_location("wcmc_reference_utils", "/root/reference/support/utils.py")
    ru = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ru)
    from wcmc_b200 import dropin
    dropin.install()
    import support.utils as du

    def flags(parser):
        return {a.dest: (tuple(a.option_strings), a.default, getattr(a, "type", None), type(a).__name__)
                for a in parser._actions}
    assert flags(du.BasicArgumentParser()) == flags(ru.BasicArgumentParser())
    g = np.random.default_rng(0)
    c = g.random((5, 7, 3)).astype(np.float32) * 4
    np.testing.assert_allclose(du.ToneMap(c), ru.ToneMap(c), rtol=1e-6)
    np.testing.assert_allclose(du.ToneMap(c, 2.5), ru.ToneMap(c, 2.5), rtol=1e-6)
    np.testing.assert_allclose(du.LinearToSrgb(c), ru.LinearToSrgb(c))
    b = (g.random((2, 3, 5, 7)).astype(np.float32) - 0.2) * 4
    np.testing.assert_allclose(du.ToneMapBatch(b), ru.ToneMapBatch(b), rtol=1e-6)
    assert not np.shares_memory(du.ToneMap(c), c)       # the reference copies, too
    for shape, tgt in (((1, 1, 128, 128), (1, 1, 92, 92)), ((2, 3, 7, 9), (2, 3, 4, 4)), ((1, 2, 5, 5), (1, 2, 5, 5))):
        src = torch.arange(float(np.prod(shape))).reshape(shape)
        assert torch.equal(du.crop_like(src, torch.empty(tgt)), ru.crop_like(src, torch.empty(tgt)))


def test_image_metrics_match_reference_losses():
    """RelativeMSE / SMAPE / TonemappedMSE / TonemappedRelativeMSE of the drop-in against the reference's own
    classes (support/losses.py:245-320), values and gradients (pure torch on any device)."""
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    from make_golden import _install_stubs
    _install_stubs()
    spec = importlib.util.spec_from_file_location("wcmc_reference_losses", "/root/reference/support/losses.py")
    rl = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(rl)
    from wcmc_b200 import dropin
    dropin.install()
    import support.losses as dl
    g = torch.Generator().manual_seed(2)
    im = (torch.rand(2, 3, 9, 11, generator=g) * 4 - 0.5)
    ref = torch.rand(2, 3, 9, 11, generator=g) * 4
    for name in ("RelativeMSE", "SMAPE", "TonemappedMSE", "TonemappedRelativeMSE"):
        a = im.clone().requires_grad_(True)
        b = im.clone().requires_grad_(True)
        la, lb = getattr(dl, name)()(a, ref), getattr(rl, name)()(b, ref)
        torch.testing.assert_close(la, lb, rtol=1e-6, atol=1e-8, msg=name)
        la.backward()
        lb.backward()
        torch.testing.assert_close(a.grad, b.grad, rtol=1e-5, atol=1e-9, msg=name)


@pytest.mark.parametrize("flags", ["--train_branches",
                                   "--use_llpm_buf --manif_learn --manif_loss FMSE --train_branches --pnet_out_size 3",
                                   "--use_llpm_buf --manif_learn --manif_loss GRS --train_branches --pnet_out_size 4 "
                                   "--disentangle m10r01",
                                   "--kpcn_ref --train_branches",
                                   "--kpcn_pre --use_llpm_buf --manif_learn --manif_loss FMSE --train_branches"])
def test_train_kpcn_init_model_against_the_dropin(flags, tmp_path):
    """The reference's own `train_kpcn.init_model` (train_kpcn.py:194-338, unmodified) builds its models, optimisers,
    losses and interface from the drop-in packages: constructor signatures, `str(model)`, parameters(), state_dict
    round trip through the checkpoint layout of train_kpcn.train (:108-121).  `.cuda()` is made a no-op (no GPU here;
    nothing is run forward)."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = r'''
import sys, types, os
for name in ("visdom", "kornia"):
    sys.modules[name] = types.ModuleType(name)
mpl = types.ModuleType("matplotlib"); plt = types.ModuleType("matplotlib.pyplot"); plt.imsave = lambda *a, **k: None
mpl.pyplot = plt; sys.modules["matplotlib"] = mpl; sys.modules["matplotlib.pyplot"] = plt
sys.path.insert(0, "/root/reference")
sys.path.insert(0, %r)
from wcmc_b200 import dropin
dropin.install()
import torch
torch.nn.Module.cuda = lambda self, *a, **k: self
torch.cuda.device_count = lambda: 1
import train_kpcn as T
p = T.BasicArgumentParser()
# the flags train_kpcn.py adds under __main__ (train_kpcn.py:386-427)
p.add_argument('--desc', type=str, default="t"); p.add_argument('--lr_dncnn', type=float, default=1e-4)
p.add_argument('--lr_pnet', type=float, nargs='+', default=[0.0001]); p.add_argument('--lr_ckpt', action='store_true')
p.add_argument('--best_err', type=float); p.add_argument('--pnet_out_size', type=int, nargs='+', default=[3])
p.add_argument('--manif_loss', type=str); p.add_argument('--train_branches', action='store_true')
p.add_argument('--use_llpm_buf', action='store_true'); p.add_argument('--manif_learn', action='store_true')
p.add_argument('--w_manif', type=float, nargs='+', default=[0.1]); p.add_argument('--disentangle', type=str, default='m11r11')
p.add_argument('--single_gpu', action='store_true'); p.add_argument('--device_id', type=int, default=0)
p.add_argument('--kpcn_ref', action='store_true'); p.add_argument('--kpcn_pre', action='store_true')
p.add_argument('--not_save', action='store_true'); p.add_argument('--local', action='store_true')
args = p.parse_args(%r.split() + ["--single_gpu", "--save", %r, "--model_name", "m"])
llpm = args.use_llpm_buf
ds = types.SimpleNamespace(dncnn_in_size=34 + (3 + 2 if llpm else 0), pnet_in_size=36 if llpm else 0, pnet_out_size=3)
interfaces, params = T.init_model({"train": ds}, args)
assert len(interfaces) == 1
itf = interfaces[0]
want = "KPCNRefInterface" if args.kpcn_ref else ("KPCNPreInterface" if args.kpcn_pre else "KPCNInterface")
assert type(itf).__name__ == want and type(itf).__module__ == "support.interfaces"
assert set(itf.models) == ({"dncnn", "backbone_diffuse", "backbone_specular"} if llpm else {"dncnn"})
assert all("optim_" + k in itf.optims for k in itf.models)
n_first = next(itf.models["dncnn"].parameters()).shape[1]
half = args.disentangle in ("m10r01", "m11r01")
exp = 34 + 3 if args.kpcn_ref else (34 + 1 + (args.pnet_out_size[0] // 2 if half else args.pnet_out_size[0]) + 1 if llpm else 34)
assert n_first == exp, (n_first, exp)
itf.to_train_mode()
# the checkpoint train_kpcn.train() writes, and the way init_model reads it back
state = {"model": str(itf.models["dncnn"]), "optims": itf.optims, "best_err": itf.best_err}
for name in itf.models:
    state["state_dict_" + name] = itf.models[name].state_dict()
fn = os.path.join(args.save, "latest_m.pth")
torch.save(state, fn)
ck = torch.load(fn, weights_only=False)
for name in itf.models:
    itf.models[name].load_state_dict(ck["state_dict_" + name])
    ck["optims"]["optim_" + name].state_dict()
if llpm:
    assert str(itf.models["backbone_diffuse"]) == "PathNet i36in64o%%d" %% args.pnet_out_size[0]
print("OK")
''' % (root, flags, str(tmp_path))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300, cwd=root)
    assert r.returncode == 0 and r.stdout.strip().endswith("OK"), (r.stdout[-800:], r.stderr[-2500:])


def test_train_kpcn_train_loop_drives_the_dropin_interface(tmp_path):
    """The reference's own `train_kpcn.train` (train_kpcn.py:87-160: epoch loop, checkpoint dicts, validation, best-error
    bookkeeping) driving the drop-in KPCNInterface for one epoch.  Stand-in torch models replace the CUDA-backed ones
    (no GPU here) and Tensor.cuda is a no-op; what is checked is the interface contract the loop relies on."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = r'''
import sys, types, os
for name in ("visdom", "kornia"):
    sys.modules[name] = types.ModuleType(name)
mpl = types.ModuleType("matplotlib"); plt = types.ModuleType("matplotlib.pyplot"); plt.imsave = lambda *a, **k: None
mpl.pyplot = plt; sys.modules["matplotlib"] = mpl; sys.modules["matplotlib.pyplot"] = plt
sys.path.insert(0, "/root/reference")
sys.path.insert(0, %r)
from wcmc_b200 import dropin
dropin.install()
import torch
torch.Tensor.cuda = lambda self, *a, **k: self
import train_kpcn as T
from tests.test_ddp_gloo import _TinyKPCN, _TinyPathNet
from tests.test_interfaces_vs_reference import _TinyManifLoss
from wcmc_b200.synth import make_batch
torch.manual_seed(0)
models = {"dncnn": _TinyKPCN(39), "backbone_diffuse": _TinyPathNet(36, 3), "backbone_specular": _TinyPathNet(36, 3)}
optims = {"optim_" + k: torch.optim.Adam(m.parameters(), lr=1e-3) for k, m in models.items()}
lf = {"l_diffuse": torch.nn.L1Loss(), "l_specular": torch.nn.L1Loss(), "l_recon": torch.nn.L1Loss(), "l_test": T.RelativeMSE(),
      "l_manif": _TinyManifLoss()}
args = types.SimpleNamespace(desc="t", start_epoch=0, num_epoch=2, model_name="m", visual=False, val_epoch=1, not_save=False,
                             save=%r)
itf = T.KPCNInterface(models, optims, lf, args, use_llpm_buf=True, manif_learn=True, w_manif=0.1, train_branches=True)
loaders = {"train": [make_batch(batch=2, spp=2, size=24, seed=s) for s in (1, 2)],
           "val": [make_batch(batch=2, spp=2, size=24, seed=3)]}
w0 = [p.detach().clone() for p in models["dncnn"].parameters()]
T.train([itf], loaders, {"data_device": 0}, args)
assert itf.iters == 4 and itf.best_err < 1e10
assert any(not torch.equal(a, b) for a, b in zip(w0, models["dncnn"].parameters()))
for fn in ("latest_m.pth", "m.pth"):
    ck = torch.load(os.path.join(args.save, fn), weights_only=False)
    assert ck["start_epoch"] in (1, 2) and ck["model"] == str(models["dncnn"]) and "state_dict_backbone_specular" in ck
    assert abs(ck["best_err"] - itf.best_err) < 1e-12 or fn == "latest_m.pth"
print("OK")
''' % (root, str(tmp_path))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300, cwd=root)
    assert r.returncode == 0 and r.stdout.strip().endswith("OK"), (r.stdout[-1500:], r.stderr[-2500:])


class _TinyBackbone(torch.nn.Module):
    """Stand-in for the single path-embedding network of train_sbmc.py: dict -> (B,S,C,H,W)."""

    def __init__(self, ic, outc):
        super().__init__()
        self.conv = torch.nn.Conv2d(ic, outc, 1)

    def forward(self, samples):
        p = samples["paths"]
        b, s, c, h, w = p.shape
        return torch.relu(self.conv(p.reshape(b * s, c, h, w))).reshape(b, s, -1, h, w)


class _TinyMultisteps(torch.nn.Module):
    """Stand-in for sbmc.Multisteps: per-sample features + radiance -> a (B,3,h,w) image, valid 3x3 conv (so that
    crop_like has something to crop)."""

    def __init__(self, n_feat):
        super().__init__()
        self.conv = torch.nn.Conv2d(n_feat + 3, 3, 3)

    def forward(self, batch):
        x = torch.cat([batch["features"], batch["radiance"]], 2).mean(1)
        return self.conv(x)


def itf_clip(which):
    return 1000.0 if which == "SBMCInterface" else 250.0


def _sbmc_batch(seed, llpm):
    g = torch.Generator().manual_seed(seed)
    b, s, h, w, f = 2, 3, 20, 20, 7
    batch = {"target_image": torch.rand(b, 3, h, w, generator=g), "radiance": torch.rand(b, s, 3, h, w, generator=g),
             "features": torch.rand(b, s, f, h, w, generator=g)}
    if llpm:
        batch["paths"] = torch.rand(b, s, 36, h, w, generator=g)
    return batch, f


@pytest.mark.parametrize("cfg", [dict(use_llpm_buf=True, manif_learn=True, disentangle="m11r11"),
                                 dict(use_llpm_buf=True, manif_learn=True, disentangle="m10r01"),
                                 dict(use_llpm_buf=True, manif_learn=True, disentangle="m11r01"),
                                 dict(use_llpm_buf=True, manif_learn=True, disentangle="m10r11"),
                                 dict(use_llpm_buf=True, manif_learn=False),
                                 dict(use_llpm_buf=False, manif_learn=False)],
                         ids=lambda c: "-".join("%s=%s" % kv for kv in c.items()))
@pytest.mark.parametrize("which", ["SBMCInterface", "LBMCInterface"])
def test_sbmc_interface_matches_reference_class(both, cfg, which):
    """SBMCInterface / LBMCInterface (/root/reference/support/interfaces.py:336-523, :753-839) with stand-in backbone /
    Multisteps modules: same losses, gradients (after clip_grad_norm_; the inputs are scaled so that the LBMC clip of
    250 actually bites), parameters after the step, validate outputs and epoch summary."""
    llpm = cfg["use_llpm_buf"]
    outc = 4
    half = cfg.get("disentangle", "m11r11") in ("m10r01", "m11r01")
    batch, f = _sbmc_batch(5, llpm)
    batch["target_image"] = batch["target_image"] * 3000.0      # large residuals -> gradient norms above both clips
    n_feat = f + ((outc // 2 if half else outc) + 1 if llpm else 0)

    def run(cls):
        torch.manual_seed(0)
        models = {"dncnn": _TinyMultisteps(n_feat)}
        if llpm:
            models["backbone"] = _TinyBackbone(36, outc)
        lf = {"l_recon": torch.nn.MSELoss(), "l_test": both.RelativeMSE()}     # MSE: its gradient grows with the residual
        if cfg["manif_learn"]:
            lf["l_manif"] = _TinyManifLoss()
        optims = {"optim_" + k: torch.optim.SGD(m.parameters(), lr=1e-6) for k, m in models.items()}   # SGD: the update
        itf = cls(models, optims, lf, types.SimpleNamespace(model_name="t"), w_manif=0.1, **cfg)       # shows the clip
        itf.to_train_mode()
        itf.preprocess(batch)
        itf.train_batch(batch)
        gn = torch.sqrt(sum(p.grad.pow(2).sum() for p in models["dncnn"].parameters()))
        assert abs(float(gn) - itf_clip(which)) < 1e-2 * itf_clip(which), "the gradient-norm clip did not bite"
        res = dict(losses={k: v.clone() for k, v in itf.m_losses.items()}, iters=itf.iters,
                   grads=[None if p.grad is None else p.grad.clone() for m in models.values() for p in m.parameters()],
                   params=[p.detach().clone() for m in models.values() for p in m.parameters()])
        itf.to_eval_mode()
        with torch.no_grad():
            res["out"], res["pb"] = itf.validate_batch(batch)
        res["m_val"] = itf.m_losses["m_val"].clone()
        res["summary"] = itf.get_epoch_summary("eval", 1.0)
        res["str"] = str(itf)
        return res

    ours, ref = run(getattr(both.ours, which)), run(getattr(both.ref, which))
    assert list(ours["losses"]) == list(ref["losses"]) and ours["iters"] == ref["iters"] and ours["str"] == ref["str"]
    for k in ref["losses"]:
        torch.testing.assert_close(ours["losses"][k], ref["losses"][k], rtol=1e-6, atol=1e-8, msg=k)
    for x, y in zip(ours["grads"], ref["grads"]):
        assert (x is None) == (y is None)
        if x is not None:
            torch.testing.assert_close(x, y, rtol=1e-5, atol=1e-9)
    for x, y in zip(ours["params"], ref["params"]):
        torch.testing.assert_close(x, y, rtol=1e-6, atol=1e-8)
    torch.testing.assert_close(ours["out"], ref["out"], rtol=1e-6, atol=1e-8)
    torch.testing.assert_close(ours["m_val"], ref["m_val"], rtol=1e-6, atol=1e-8)
    assert abs(ours["summary"] - ref["summary"]) < 1e-7
    assert (ours["pb"] is None) == (ref["pb"] is None)
    if ref["pb"] is not None:
        torch.testing.assert_close(ours["pb"], ref["pb"], rtol=1e-6, atol=1e-8)


def test_sbmc_interface_names_the_non_finite_term(both):
    batch, f = _sbmc_batch(6, False)
    models = {"dncnn": _TinyMultisteps(f)}
    optims = {"optim_dncnn": torch.optim.Adam(models["dncnn"].parameters(), lr=1e-3)}
    itf = both.ours.SBMCInterface(models, optims, {"l_recon": torch.nn.L1Loss(), "l_test": both.RelativeMSE()},
                                  types.SimpleNamespace(model_name="t"))
    itf.to_train_mode()
    bad = dict(batch, target_image=batch["target_image"] * float("nan"))
    itf.preprocess(bad)
    with pytest.raises(RuntimeError, match="l_total: Non-finite loss at train time."):
        itf.train_batch(bad)
