"""Seeded synthetic KPCN / WCMC batches honouring the reference tensor contract.

Follows the batch dict produced by `DenoiseDataset.__getitem__`
(/root/reference/support/datasets.py:128-159 channel order, :286-299 gradients,
:1078-1126 slices / targets).  Value ranges follow SURVEY.md §8(d).  CPU generator only, so
the same seed gives the same batch in the authoring container and on the GPU box.
"""
import torch
import torch.nn.functional as F


def _grads(buf):
    """datasets.py:286-299: forward differences, zero padded on the left / top."""
    dx = F.pad(buf[..., :, 1:] - buf[..., :, :-1], (1, 0, 0, 0))
    dy = F.pad(buf[..., 1:, :] - buf[..., :-1, :], (0, 0, 1, 0))
    return torch.cat([dx, dy], 1)


def _blur5(x):
    c = x.shape[1]
    k = torch.full((c, 1, 5, 5), 1.0 / 25.0, dtype=x.dtype)
    return F.conv2d(F.pad(x, (2, 2, 2, 2), mode="replicate"), k, groups=c)


def make_batch(batch=8, spp=8, size=128, seed=1234, paths=True, llpm_channel=None,
               height=None, width=None):
    """Returns a dict of fp32 CPU tensors.

    kpcn_*_in are (B,34,H,W), or (B,35,H,W) with the path-weight mean channel when
    ``llpm_channel`` (defaults to ``paths``); paths is (B,S,36,H,W)."""
    if llpm_channel is None:
        llpm_channel = paths
    h = height or size
    w = width or size
    g = torch.Generator().manual_seed(seed)

    def U(*shape, lo=0.0, hi=1.0):
        return torch.rand(*shape, generator=g) * (hi - lo) + lo

    diffuse = U(batch, 3, h, w, hi=2.0)
    specular = torch.log1p(-torch.log(U(batch, 3, h, w, lo=1e-6)))
    normals = F.normalize(U(batch, 3, h, w, lo=-1.0), dim=1)
    depth = U(batch, 1, h, w)
    albedo = U(batch, 3, h, w)

    def group(buf, nvar=1):
        return torch.cat([buf, U(batch, nvar, h, w, hi=0.1), _grads(buf)], 1)

    shared = torch.cat([group(normals), group(depth), group(albedo)], 1)  # 10 + 4 + 10
    d_in = torch.cat([group(diffuse), shared], 1)
    s_in = torch.cat([group(specular), shared], 1)
    if llpm_channel:
        pw = U(batch, 1, h, w, lo=-0.15, hi=0.0)
        d_in = torch.cat([d_in, pw], 1)
        s_in = torch.cat([s_in, pw], 1)
    albedo_eps = albedo + 0.00316
    t_d = _blur5(diffuse)
    t_s = _blur5(specular)
    out = {
        "kpcn_diffuse_in": d_in.contiguous(),
        "kpcn_specular_in": s_in.contiguous(),
        "kpcn_diffuse_buffer": diffuse.contiguous(),
        "kpcn_specular_buffer": specular.contiguous(),
        "kpcn_albedo": albedo_eps.contiguous(),
        "target_diffuse": t_d.contiguous(),
        "target_specular": t_s.contiguous(),
        "target_total": (albedo_eps * t_d + torch.exp(t_s) - 1.0).contiguous(),
    }
    if paths:
        p = (torch.randn(batch, spp, 36, h, w, generator=g) * 0.3).clamp_(-1.0, 1.0)
        bounce = torch.randint(0, 20, (batch, spp, 6, h, w), generator=g).float() / 19.0
        p[:, :, 24:30] = bounce
        p[:, :, 30:36] = U(batch, spp, 6, h, w)
        out["paths"] = p.contiguous()
    return out
