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
