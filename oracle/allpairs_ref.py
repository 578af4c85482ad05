"""CPU oracle of the ALL-PAIRS form of the path-disentangling loss (test infrastructure only).

NOT reference behaviour: /root/reference/support/losses.py pairs every row with ONE random partner
(`FeatureMSE`, losses.py:33-61, :82-113).  The all-pairs quantity is what that
permutation loss estimates (one sample per row) and is what BASELINE.json's north_star (4) / configs[4]
describe; it is an extension of this repository, restated here in plain chunked PyTorch (fp64) so the
tensor-core kernel (wcmc_fmse_allpairs_fwd) has something independent to be checked against.

Rows: P (N, D) embeddings, R (N, 3) reference radiance (tone-mapped inside exactly like
losses.py:63-65).  For every ordered pair i != j
    d_p = 1/2 |P_i - P_j|^2 ,  d_t = 1/2 |t_i - t_j|^2 ,  e = d_p - d_t
    mse : L = (1 / N^2) * sum_{i != j, keep(i,j)} 1/2 e^2            (the i = j terms are zero)
    lse : L = (logsumexp(alpha * [e, -e over kept pairs, 0]) - log(1 + 2 * #kept)) / sqrt(alpha)
          (the all-pairs analogue of GlobalRelativeSimilarityLoss, losses.py:185-211)
    keep(i,j) = d_t < tau   when a weak-label threshold tau is given ("weak-label masking": only pairs
                whose reference colours are close act as positives), else all pairs.
"""
import math

import torch


def tonemap_gamma(img):
    img = torch.clamp(img, min=0)
    return (img / (1 + img)) ** 0.454545


def allpairs_loss(p, ref, mode="mse", alpha=2.0, tau=None, chunk=2048):
    p = p.double()
    t = tonemap_gamma(ref.double())
    n = p.shape[0]
    total = torch.zeros((), dtype=torch.float64)
    kept = 0
    m_run, s_run = 0.0, 1.0  # running logsumexp of the multiset {0}: max 0, sum exp(0 - 0) = 1
    for i0 in range(0, n, chunk):
        pi, ti = p[i0:i0 + chunk], t[i0:i0 + chunk]
        d_p = 0.5 * torch.cdist(pi, p).pow(2)
        d_t = 0.5 * torch.cdist(ti, t).pow(2)
        e = d_p - d_t
        keep = torch.ones_like(e, dtype=torch.bool)
        idx = torch.arange(i0, min(i0 + chunk, n))
        keep[torch.arange(idx.numel()), idx] = False
        if tau is not None:
            keep &= d_t < tau
        ek = e[keep]
        kept += int(ek.numel())
        if mode == "mse":
            total += 0.5 * (ek ** 2).sum()
        else:
            x = alpha * torch.cat([ek, -ek])
            if x.numel():
                m_new = max(m_run, float(x.max()))
                s_run = s_run * math.exp(m_run - m_new) + float(torch.exp(x - m_new).sum())
                m_run = m_new
    if mode == "mse":
        return total / (n * n)
    return torch.tensor((m_run + math.log(s_run) - math.log(1 + 2 * kept)) / math.sqrt(alpha), dtype=torch.float64)
