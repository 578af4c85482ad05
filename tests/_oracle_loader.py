"""Loads the CPU oracle under private module names (``oracle_sbmc``, ``oracle_wcmc_ref``) so it
can coexist with the product's drop-in ``sbmc`` package in one process."""
import importlib.util
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_cache = {}


def load_oracle():
    if "o" in _cache:
        return _cache["o"]
    odir = os.path.join(ROOT, "oracle")
    saved = {k: sys.modules.pop(k) for k in list(sys.modules) if k == "sbmc" or k.startswith("sbmc.")}
    saved_path = list(sys.path)
    try:
        sys.path.insert(0, odir)
        import sbmc as osbmc  # oracle/sbmc
        spec = importlib.util.spec_from_file_location("oracle_wcmc_ref", os.path.join(odir, "wcmc_ref.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        sys.path[:] = saved_path
        for k in [k for k in sys.modules if k == "sbmc" or k.startswith("sbmc.")]:
            sys.modules["oracle_" + k] = sys.modules.pop(k)
        sys.modules.update(saved)
    o = types.SimpleNamespace(sbmc=osbmc, ref=mod, KPCN=osbmc.KPCN, modules=osbmc.modules,
                              PathNet=mod.PathNet)
    _cache["o"] = o
    return o


def precision_matched(module, dtype):
    """Makes an ORACLE module evaluate with the backend's storage precision: every Conv2d rounds its
    input activation and its (effective) weight to `dtype` before an fp32 convolution -- exactly the
    points where the CUDA path stores 16-bit tensors (biases, accumulation, logits, losses stay
    fp32).  Gradients pass straight through the rounding.  Used to separate "the kernels compute the
    wrong thing" from "a 16-bit forward flips a few ReLU masks" (precision policy)."""
    import torch
    import torch.nn.functional as F

    def rnd(t):
        return t + (t.to(dtype).to(t.dtype) - t).detach()

    handles = []
    for m in module.modules():
        if isinstance(m, torch.nn.Conv2d):
            def fwd(x, m=m):
                return F.conv2d(rnd(x), rnd(m.weight), m.bias, m.stride, m.padding)
            m._orig_forward = m.forward
            m.forward = fwd
            handles.append(m)
    return handles


def restore_precision(handles):
    for m in handles:
        m.forward = m._orig_forward
        del m._orig_forward
