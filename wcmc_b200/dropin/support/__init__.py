"""B200 drop-in for the reference's `support` package (hot-path modules only).

`networks`, `losses`, `interfaces` and `utils` are re-implemented here on top of libwcmc.so.
Everything else the reference scripts import from `support` (datasets, metrics, img_utils: CPU
numpy I/O, out of scope) falls through to the reference checkout when one is reachable: set
WCMC_REFERENCE=/path/to/WCMC or keep the checkout on sys.path.
"""
import os
import sys

_here = os.path.dirname(os.path.abspath(__file__))


def _reference_support_dirs():
    cands = []
    env = os.environ.get("WCMC_REFERENCE")
    if env:
        cands.append(os.path.join(env, "support"))
    for p in sys.path:
        d = os.path.join(p or ".", "support")
        if os.path.isdir(d) and os.path.abspath(d) != _here and os.path.exists(os.path.join(d, "datasets.py")):
            cands.append(os.path.abspath(d))
    return [c for c in cands if os.path.isdir(c)]


for _d in _reference_support_dirs():
    if _d not in __path__:
        __path__.append(_d)
