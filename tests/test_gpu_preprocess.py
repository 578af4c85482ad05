"""GPU parity of the N3 preprocessing kernels (wcmc_b200/preprocess.py) against the reference-generated vectors.

The kernels were written after round 1's GPU budget was spent: their per-element arithmetic is checked on the CPU
(tests/test_host_logic.py::test_preprocess_kernel_arithmetic_on_host) but they had never been launched when the round
closed, and nothing on the product path calls them.  Until a GPU run has been seen green the tests are non-strict
xfail -- they RUN (this file is the last GPU file of the suite), a pass is reported as XPASS, a failure cannot turn the
tier red.  WCMC_UNVALIDATED=1 makes them ordinary tests (tools/round2_first.sh); drop the marker once green."""
import os

import numpy as np
import pytest
import torch

pytestmark = [pytest.mark.gpu]
if os.environ.get("WCMC_UNVALIDATED") != "1":
    pytestmark.append(pytest.mark.xfail(reason="first GPU run of the N3 preprocessing kernels (never launched in round 1)",
                                        strict=False))

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("tag", ["a", "b"])
def test_preprocess_kernels_vs_reference_golden(tag):
    from wcmc_b200 import preprocess
    g = np.load(os.path.join(ROOT, "tests", "golden", "ref_golden_n3.npz"))
    raw = torch.from_numpy(g["raw_" + tag]).cuda()
    kp = preprocess.preprocess_kpcn(raw).cpu().numpy()
    ll = preprocess.preprocess_llpm(raw).cpu().numpy()
    want_k, want_l = g["kpcn_" + tag], g["llpm_" + tag]
    ok = np.isfinite(want_k)     # tag "a": the reference itself yields NaN where a 1e38 outlier overflows the variance
    np.testing.assert_allclose(kp[ok], want_k[ok], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(ll, want_l, rtol=1e-5, atol=1e-6)
    if tag == "b":
        assert ok.all()
    t = preprocess.kpcn_batch_tensors(torch.from_numpy(want_k).cuda(), torch.from_numpy(want_l).cuda())
    h, w, s = raw.shape[:3]
    assert tuple(t["kpcn_diffuse_in"].shape) == (35, h, w) and tuple(t["paths"].shape) == (s, 36, h, w)
