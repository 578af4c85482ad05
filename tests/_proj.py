"""Compact per-tensor gradient fingerprints for the golden fixtures.

A full gradient set of the three models is 47 MB -- too large for a committed fixture -- and the round-1 proxies
(sum of |g| per tensor) cannot see a wrong direction.  A fingerprint is K = 16 dot products of the flattened tensor
with fixed pseudo-random sign vectors plus its L2 norm: for any difference d = g_ours - g_ref,
E[(r . d)^2] = |d|^2 over random signs r, so   sqrt(mean_k (r_k . g_ours - r_k . g_ref)^2) / |g_ref|   is an unbiased
(Johnson-Lindenstrauss) estimate of the per-tensor relative L2 error, good to ~20 % with K = 16.  The sign vectors
come from an integer hash of (element index, k), identical on every device and torch version."""
import torch

K = 16


def _s64(c):
    return c - (1 << 64) if c >= (1 << 63) else c


_C1, _C2, _C3 = _s64(0x9E3779B97F4A7C15), _s64(0xBF58476D1CE4E5B9), _s64(0x94D049BB133111EB)


def signs(n, k, device):
    i = torch.arange(1, n + 1, dtype=torch.int64, device=device)
    kk = torch.full((1,), k + 1, dtype=torch.int64, device=device) * _C2     # wraps in int64, like the lines below
    x = i * _C1 + kk
    x = (x ^ (x >> 31)) * _C3
    x = x ^ (x >> 29)
    return ((x >> 17) & 1).to(torch.float64) * 2.0 - 1.0


def fingerprint(t):
    """-> float64 tensor (K + 1,): K sign projections and the L2 norm of `t` (CPU)."""
    v = t.detach().reshape(-1).double()
    out = [torch.dot(signs(v.numel(), k, v.device), v) for k in range(K)]
    out.append(v.norm())
    return torch.stack(out).cpu()


def fingerprints(tensors):
    return torch.stack([fingerprint(t) for t in tensors])


def est_rel(fp_ours, fp_ref):
    """Per-tensor estimated relative L2 error, (n,) float64, from two (n, K + 1) fingerprint tables."""
    d = fp_ours[:, :K] - fp_ref[:, :K]
    return d.pow(2).mean(1).sqrt() / (fp_ref[:, K] + 1e-30)


def est_rel_global(fp_ours, fp_ref):
    """The same estimate for the concatenation of all tensors of a model."""
    d = fp_ours[:, :K] - fp_ref[:, :K]
    return float((d.pow(2).mean(1).sum().sqrt() / (fp_ref[:, K].pow(2).sum().sqrt() + 1e-30)))


def est_abs(fp_ours, fp_ref):
    """Per-tensor estimated L2 norm of the difference, (n,) float64."""
    return (fp_ours[:, :K] - fp_ref[:, :K]).pow(2).mean(1).sqrt()


def param_rms_shift_in_lr(params, fp_ref, lr):
    """RMS per-element difference of every updated parameter tensor to the reference's, in units of the learning
    rate.  Adam's first step moves each weight by lr * sign(g), so a weight whose (tiny) gradient has the other sign
    differs by 2 * lr: a value of x means a fraction (x / 2)^2 of the signs differ."""
    params = list(params)
    d = est_abs(fingerprints(params), fp_ref)
    n = torch.tensor([p.numel() for p in params], dtype=torch.float64)
    return d / n.sqrt() / lr
