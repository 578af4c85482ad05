"""Golden vectors for SURVEY.md §8(f) N3 (preprocessing of raw OptaGen sample buffers), made by running the
REFERENCE's own `DenoiseDataset._preprocess_kpcn` / `_preprocess_llpm` / `_gradients`
(/root/reference/support/datasets.py:286-361, :487-582) on a seeded synthetic raw buffer.

Run in the authoring container only (needs /root/reference):   python tests/golden/make_golden_n3.py
The dataset object is built over an empty temporary directory tree (its constructor only lists files).
Output: tests/golden/ref_golden_n3.npz (small; committed)
"""
import os
import sys
import tempfile

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_golden import ROOT, REF, _install_stubs  # noqa: E402


def raw_buffer(h, w, s, seed):
    """Seeded (H,W,S,104) fp32 buffer in the value ranges OptaGen writes: non-negative radiances with a few
    negative / non-finite outliers (the reference clamps them), unit normals, depths up to ~50, albedo in [0,1],
    path weights / throughputs spanning decades, integer bounce-type tags, roughness in [0,1]."""
    g = np.random.default_rng(seed)
    x = g.random((h, w, s, 104), dtype=np.float32)
    x[..., 2:8] = g.gamma(0.7, 1.5, (h, w, s, 6)).astype(np.float32)
    x[..., 2:8] -= 0.05                                   # some negatives: np.maximum(., 0) on the path
    x[..., 66:69] = g.random((h, w, s, 3), dtype=np.float32)                     # albedo at the first diffuse bounce
    n = g.normal(size=(h, w, s, 3)).astype(np.float32)
    x[..., 69:72] = n / np.linalg.norm(n, axis=-1, keepdims=True)
    x[..., 72:73] = (g.random((h, w, s, 1), dtype=np.float32) * 50.0)
    x[..., 73:74] = np.exp(g.normal(-3, 3, (h, w, s, 1))).astype(np.float32)    # path weight
    x[..., 74:80] = np.exp(g.normal(0, 2, (h, w, s, 6))).astype(np.float32)     # radiance w/o weight, light intensity
    x[..., 80:98] = np.exp(g.normal(-1, 2, (h, w, s, 18))).astype(np.float32)   # throughputs
    x[..., 60:66] = g.integers(0, 20, (h, w, s, 6)).astype(np.float32)          # bounce types
    if seed == 3:   # outliers: the reference clamps them to 1e38, whose variance overflows to inf / NaN downstream
        x[1, 2, 0, 3] = np.inf
        x[3, 1, 1, 6] = np.nan
        x[2, 2, 1, 72] = np.float32(3.0e38)
    return x


def main():
    _install_stubs()
    sys.path.insert(0, REF)
    sys.path.insert(0, ROOT)
    from support import datasets as D
    d = tempfile.mkdtemp()
    os.makedirs(os.path.join(d, "train", "gt"))
    ds = D.DenoiseDataset(d, 4, "kpcn", "train", 8, "random", True, False, True, 3)
    assert ds.MAX_DEPTH == 5
    out = {}
    for tag, (h, w, s, seed) in {"a": (12, 10, 4, 3), "b": (9, 16, 2, 4)}.items():
        raw = raw_buffer(h, w, s, seed)
        san = np.where(np.isfinite(raw), raw, 1.0e+38)          # datasets.py:621-624, verbatim
        san = np.where(san < 1.0e+38, san, 1.0e+38)
        kp = ds._preprocess_kpcn(san.astype(np.float32))
        ll = ds._preprocess_llpm(san.astype(np.float32))
        out["raw_" + tag] = raw
        out["kpcn_" + tag] = kp.astype(np.float32)
        out["llpm_" + tag] = ll.astype(np.float32)
        out["grad_" + tag] = ds._gradients(kp[..., :3]).astype(np.float32)
        print(tag, raw.shape, "->", kp.shape, ll.shape, "finite:", np.isfinite(kp).all(), np.isfinite(ll).all())
    fn = os.path.join(ROOT, "tests", "golden", "ref_golden_n3.npz")
    np.savez_compressed(fn, **out)
    print("wrote", fn, os.path.getsize(fn), "bytes")


if __name__ == "__main__":
    main()
