"""CPU: the C-ABI library loads and exports every symbol include/wcmc.h declares; the product
path refuses to run without a CUDA device (no CPU fallback)."""
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "wcmc.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(wcmc_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from wcmc_b200 import lib
    handle = lib.load()
    names = _declared()
    assert len(names) >= 15
    for n in names:
        assert hasattr(handle, n), "libwcmc.so does not export %s" % n
        assert n in lib.SIGNATURES, "%s has no ctypes signature in wcmc_b200/lib.py" % n
    assert sorted(lib.SIGNATURES) == names
    assert b"sm_100a" in handle.wcmc_version()


def test_no_cpu_fallback():
    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    from wcmc_b200 import dropin, lib
    with pytest.raises(lib.WcmcError):
        lib.init()
    dropin.install()
    from sbmc import KPCN
    from wcmc_b200.synth import make_batch
    net = KPCN(34)
    with pytest.raises((lib.WcmcError, AssertionError, RuntimeError)):
        net(make_batch(batch=1, size=40, paths=False))


def test_product_path_never_imports_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may touch oracle/."""
    bad = []
    for base, _, files in os.walk(os.path.join(ROOT, "wcmc_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(base, f)).read()
                if re.search(r"^\s*(from|import)\s+oracle|_oracle_loader|[\"']oracle[\"']", txt, flags=re.M):
                    bad.append(os.path.join(base, f))
    assert not bad, bad


def test_dropin_api_surface():
    from wcmc_b200 import dropin
    dropin.install()
    import inspect
    from sbmc import KPCN, modules
    from support.interfaces import KPCNInterface
    from support.losses import FeatureMSE, GlobalRelativeSimilarityLoss, RelativeMSE  # noqa: F401
    from support.networks import PathNet
    from support.utils import BasicArgumentParser, crop_like  # noqa: F401
    from ttools.modules.image_operators import crop_like as c2  # noqa: F401
    assert list(inspect.signature(KPCNInterface.__init__).parameters)[1:] == [
        "models", "optims", "loss_funcs", "args", "visual", "use_llpm_buf", "manif_learn", "w_manif",
        "train_branches", "disentanglement_option"]
    assert list(inspect.signature(modules.ConvChain.__init__).parameters)[1:] == [
        "ninputs", "noutputs", "ksize", "width", "depth", "stride", "pad", "normalize", "normalization_type",
        "output_type", "activation", "weight_norm"]
    assert list(inspect.signature(PathNet.__init__).parameters)[1:] == ["ic", "intermc", "outc"]
    assert str(PathNet(36, outc=3)) == "PathNet i36in64o3"
    for m in ("to_train_mode", "preprocess", "train_batch", "validate_batch", "to_eval_mode", "get_epoch_summary"):
        assert callable(getattr(KPCNInterface, m))
    assert KPCN(34).n_in == 34


def test_state_dict_keys_match_oracle(oracle):
    from wcmc_b200 import dropin
    dropin.install()
    from sbmc import KPCN
    from support.networks import PathNet
    assert sorted(KPCN(39).state_dict()) == sorted(oracle.KPCN(39).state_dict())
    assert sorted(PathNet(36, outc=4).state_dict()) == sorted(oracle.PathNet(36, outc=4).state_dict())
    ours, ref = PathNet(36, outc=4), oracle.PathNet(36, outc=4)
    ours.load_state_dict(ref.state_dict())
    for (k1, v1), (k2, v2) in zip(sorted(ours.state_dict().items()), sorted(ref.state_dict().items())):
        assert k1 == k2 and torch.equal(v1, v2)


def test_losses_refuse_cpu_tensors(golden):
    """The path-disentangling loss is a CUDA kernel pair; on a CPU tensor it raises (no fallback).
    RelativeMSE is a one-line torch expression and is pinned here against the reference's vector."""
    from wcmc_b200 import dropin, lib
    dropin.install()
    from support.losses import FeatureMSE, GlobalRelativeSimilarityLoss, RelativeMSE
    g = golden["fmse_a_nl1"]
    with pytest.raises(lib.WcmcError):
        FeatureMSE(non_local=True)(g["p"], g["ref"])
    with pytest.raises(lib.WcmcError):
        GlobalRelativeSimilarityLoss()(g["p"], g["ref"])
    g = golden["relmse"]
    torch.testing.assert_close(RelativeMSE()(g["im"], g["ref"]), g["loss"], rtol=1e-5, atol=1e-7)
