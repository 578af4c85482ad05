"""ORACLE (test infrastructure, never shipped on the product path).

Pure-PyTorch CPU restatement of the in-repo part of WCMC's hot path.  Unlike the ``sbmc``
pieces this part IS pinned: ``tests/golden/make_golden.py`` imports the reference's own
`support/losses.py`, `support/networks.py` and `support/interfaces.py` in the authoring
container and the restatement is checked against those outputs (tests/test_oracle.py).

  PathNet              follows /root/reference/support/networks.py:7-42
  feature_mse          follows /root/reference/support/losses.py:33-65, 82-113
  grs_loss             follows /root/reference/support/losses.py:131-211
  relative_mse         follows /root/reference/support/losses.py:255-264
  pbuffer_concat       follows /root/reference/support/interfaces.py:165-180
  kpcn_train_step      follows /root/reference/support/interfaces.py:122-271
  kpcn_validate        follows /root/reference/support/interfaces.py:278-318
"""
import math
import os
import sys

import torch
import torch.nn as nn

_here = os.path.dirname(os.path.abspath(__file__))
if _here not in sys.path:
    sys.path.insert(0, _here)

from sbmc import modules as ops  # noqa: E402  (oracle/sbmc)
from sbmc.modules import crop_like  # noqa: E402


class PathNet(nn.Module):
    """networks.py:7-42"""

    def __init__(self, ic, intermc=64, outc=3):
        super().__init__()
        self.ic, self.intermc, self.outc = ic, intermc, outc
        self.final_ic = intermc + intermc
        self.embedding = ops.ConvChain(ic, intermc, width=intermc, depth=3, ksize=1, pad=False)
        self.propagation = ops.Autoencoder(intermc, intermc, num_levels=3, increase_factor=2.0,
                                           num_convs=3, width=intermc, ksize=3,
                                           output_type="leaky_relu", pooling="max")
        self.final = ops.ConvChain(self.final_ic, outc, width=self.final_ic, depth=2, ksize=1,
                                   pad=False, output_type="relu")

    def __str__(self):
        return "PathNet i{}in{}o{}".format(self.ic, self.intermc, self.outc)

    def forward(self, samples):
        paths = samples["paths"]
        bs, spp, nf, h, w = paths.shape
        flat = self.embedding(paths.contiguous().view(bs * spp, nf, h, w))
        flat = flat.view(bs, spp, self.intermc, h, w)
        reduced = flat.mean(1)
        propagated = self.propagation(reduced)
        both = torch.cat([flat, propagated.unsqueeze(1).expand(-1, spp, -1, -1, -1)], 2)
        out = self.final(both.reshape(bs * spp, self.final_ic, h, w))
        return out.view(bs, spp, self.outc, h, w)


def tonemap_gamma(img):
    """losses.py:63-65"""
    img = torch.clamp(img, min=0)
    return (img / (1 + img)) ** 0.454545


def _paired(p_rows, r_rows, idx, dim):
    d_r = 0.5 * ((r_rows - r_rows.index_select(dim, idx)) ** 2).sum(-1)
    d_p = 0.5 * ((p_rows - p_rows.index_select(dim, idx)) ** 2).sum(-1)
    return d_p - d_r


def feature_mse(p_buffer, ref, non_local=True, idx_patch=None, idx_batch=None):
    """FeatureMSE.forward (losses.py:82-113).  Without explicit indices the permutations are
    drawn exactly like the reference: CPU default generator, ``randperm(s*h*w)`` then
    ``randperm(b*s*h*w)`` (losses.py:35, :50)."""
    b, s, c, h, w = p_buffer.shape
    ref = tonemap_gamma(ref)
    ref = torch.stack((ref,) * s, dim=1)
    if not torch.isfinite(p_buffer).all() or not torch.isfinite(ref).all():
        raise RuntimeError("Infinite loss at train time.")
    if idx_patch is None:
        idx_patch = torch.randperm(s * h * w)
    r1 = ref.permute(0, 1, 3, 4, 2).reshape(b, s * h * w, 3)
    p1 = p_buffer.permute(0, 1, 3, 4, 2).reshape(b, s * h * w, c)
    e = _paired(p1, r1, idx_patch.to(p1.device), 1)
    loss_p = 0.5 * torch.mean(e ** 2)
    if not non_local:
        return loss_p + loss_p
    if idx_batch is None:
        idx_batch = torch.randperm(b * s * h * w)
    e = _paired(p1.reshape(-1, c), r1.reshape(-1, 3), idx_batch.to(p1.device), 0)
    loss_b = 0.5 * torch.mean(e ** 2)
    return loss_p + loss_b


def grs_loss(p_buffer, ref, alpha=2, idx_patch=None, idx_batch=None):
    """GlobalRelativeSimilarityLoss.forward (losses.py:185-211)."""
    if not torch.isfinite(p_buffer).all() or not torch.isfinite(ref).all():
        raise RuntimeError("Infinite loss at train time.")
    b, s, c, h, w = p_buffer.shape
    ref = tonemap_gamma(ref)
    ref = torch.stack((ref,) * s, dim=1)
    if idx_patch is None:
        idx_patch = torch.randperm(s * h * w)
    if idx_batch is None:
        idx_batch = torch.randperm(b * s * h * w)
    r1 = ref.permute(0, 1, 3, 4, 2).reshape(b, s * h * w, 3)
    p1 = p_buffer.permute(0, 1, 3, 4, 2).reshape(b, s * h * w, c)
    d_p = _paired(p1, r1, idx_patch, 1).reshape(-1)
    d_b = _paired(p1.reshape(-1, c), r1.reshape(-1, 3), idx_batch, 0)
    zero = torch.zeros(1, dtype=p_buffer.dtype)
    ex = alpha * torch.cat([d_p, d_b, -d_p, -d_b, zero], 0)
    return (torch.logsumexp(ex, 0) - math.log(1 + 4 * b * s * h * w)) / math.sqrt(alpha)


def relative_mse(im, ref, eps=1e-2):
    """RelativeMSE.forward (losses.py:255-264)."""
    return 0.5 * torch.mean((im - ref) ** 2 / (ref ** 2 + eps))


def split_disentangle(p_buffers, option):
    """interfaces.py:139-163 -> (p_buffers for regression, out_manif for the loss)."""
    c = p_buffers["diffuse"].shape[2]
    assert c >= 2
    lo = {k: v[:, :, :c // 2] for k, v in p_buffers.items()}
    hi = {k: v[:, :, c // 2:] for k, v in p_buffers.items()}
    if option == "m11r11":
        return p_buffers, p_buffers
    if option == "m10r01":
        return lo, hi
    if option == "m11r01":
        return lo, p_buffers
    if option == "m10r11":
        return p_buffers, hi
    raise ValueError(option)


def pbuffer_concat(kpcn_in, p):
    """interfaces.py:165-180: cat[kpcn_in, mean_S(p), var_S(p).mean(C)/S (detached)]."""
    p_var = p.var(1).mean(1, keepdim=True).detach() / p.shape[1]
    return torch.cat([kpcn_in, p.mean(1), p_var], 1)


def kpcn_losses(models, batch, *, use_llpm_buf, manif_learn, w_manif=0.1, train_branches=True,
                disentangle="m11r11", non_local=True, perms=None, manif_fn=None):
    """Forward + backward of KPCNInterface.train_batch up to (not including) `_logging`
    (interfaces.py:122-251).  Grad accumulates into the models' parameters.
    ``perms`` optionally = dict(diffuse=(idx_patch, idx_batch), specular=(...)).
    Returns (loss_dict, out, p_buffers)."""
    out_manif = None
    p_buffers = None
    if use_llpm_buf:
        models["backbone_diffuse"].zero_grad()
        models["backbone_specular"].zero_grad()
        p_buffers = {"diffuse": models["backbone_diffuse"](batch),
                     "specular": models["backbone_specular"](batch)}
        p_reg, out_manif = split_disentangle(p_buffers, disentangle)
        batch = dict(batch)
        batch["kpcn_diffuse_in"] = pbuffer_concat(batch["kpcn_diffuse_in"], p_reg["diffuse"])
        batch["kpcn_specular_in"] = pbuffer_concat(batch["kpcn_specular_in"], p_reg["specular"])
    models["dncnn"].zero_grad()
    out = models["dncnn"](batch)
    total, diffuse, specular = out["radiance"], out["diffuse"], out["specular"]
    loss = {}
    tgt_total = crop_like(batch["target_total"], total)
    l1 = nn.functional.l1_loss
    if train_branches:
        tgt_d = crop_like(batch["target_diffuse"], diffuse)
        tgt_s = crop_like(batch["target_specular"], specular)
        L_d = l1(diffuse, tgt_d)
        L_s = l1(specular, tgt_s)
        # NOTE: the reference stores `L_diffuse.detach()` (shares storage) and then does the
        # in-place `L_diffuse += L_manif * w_manif` (interfaces.py:221-232), so the logged
        # l_diffuse / l_specular INCLUDE the weighted manifold term.  Restated faithfully.
        if manif_learn:
            fn = manif_fn or (lambda p, r, pp: feature_mse(p, r, non_local, *(pp or (None, None))))
            pd = perms["diffuse"] if perms else None
            ps = perms["specular"] if perms else None
            Lm_d = fn(crop_like(out_manif["diffuse"], diffuse), tgt_d, pd)
            L_d = L_d + Lm_d * w_manif
            Lm_s = fn(crop_like(out_manif["specular"], specular), tgt_s, ps)
            L_s = L_s + Lm_s * w_manif
            loss["l_manif_diffuse"], loss["l_manif_specular"] = Lm_d.detach(), Lm_s.detach()
        loss["l_diffuse"], loss["l_specular"] = L_d.detach(), L_s.detach()
        L_d.backward()
        L_s.backward()
        with torch.no_grad():
            loss["l_total"] = l1(total, tgt_total)
    else:
        L_t = l1(total, tgt_total)
        loss["l_total"] = L_t.detach()
        L_t.backward()
    with torch.no_grad():
        loss["rmse"] = relative_mse(total, tgt_total)
    return loss, out, p_buffers


def clip_and_step(models, optims):
    """interfaces.py:259-261, 269-271: clip_grad_value_(1.0) for every model, then Adam."""
    for name in models:
        nn.utils.clip_grad_value_(models[name].parameters(), clip_value=1.0)
    for name in models:
        optims["optim_" + name].step()


def kpcn_train_step(models, optims, batch, **kw):
    loss, out, p_buffers = kpcn_losses(models, batch, **kw)
    for k, v in loss.items():
        if not torch.isfinite(v).all():
            raise RuntimeError("%s: Non-finite loss at train time." % k)
    clip_and_step(models, optims)
    return loss, out, p_buffers


@torch.no_grad()
def kpcn_validate(models, batch, *, use_llpm_buf, disentangle="m11r11"):
    """interfaces.py:278-318 -> (radiance, p_buffers, relmse)."""
    p_buffers = None
    if use_llpm_buf:
        p_buffers = {"diffuse": models["backbone_diffuse"](batch),
                     "specular": models["backbone_specular"](batch)}
        if disentangle in ("m10r01", "m11r01"):
            c = p_buffers["diffuse"].shape[2]
            p_buffers = {k: v[:, :, :c // 2] for k, v in p_buffers.items()}
        batch = dict(batch)
        batch["kpcn_diffuse_in"] = pbuffer_concat(batch["kpcn_diffuse_in"], p_buffers["diffuse"])
        batch["kpcn_specular_in"] = pbuffer_concat(batch["kpcn_specular_in"], p_buffers["specular"])
    out = models["dncnn"](batch)
    tgt = crop_like(batch["target_total"], out["radiance"])
    return out["radiance"], p_buffers, relative_mse(out["radiance"], tgt)


# ------------------------------------------------------------------------------------------------
# the ablation interfaces (re-wirings of the same step)
# ------------------------------------------------------------------------------------------------
def kpcn_ref_batch(batch):
    """interfaces.py:543-552 / :566-575 (KPCNRefInterface): KPCN inputs = cat[inputs, reference image]."""
    new = dict(batch)
    new["kpcn_diffuse_in"] = torch.cat([batch["kpcn_diffuse_in"], batch["target_diffuse"]], 1)
    new["kpcn_specular_in"] = torch.cat([batch["kpcn_specular_in"], batch["target_specular"]], 1)
    return new


def kpcn_ref_train_step(models, optims, batch, train_branches=True):
    """KPCNRefInterface.train_batch (interfaces.py:540-561): the vanilla step on the widened inputs."""
    return kpcn_train_step(models, optims, kpcn_ref_batch(batch), use_llpm_buf=False, manif_learn=False,
                           train_branches=train_branches)


def kpcn_pre_train_step(models, optims, batch, *, manif_learn, w_manif=0.1, train_branches=True, non_local=True):
    """KPCNPreInterface.train_batch (interfaces.py:616-672) with its own _backward / _logging / _optimization
    (:674-750).  manif_learn=True: only the two PathNets run, the manifold loss is taken on the UNcropped
    p-buffers and targets (:686-699), only the backbones are clipped and stepped.  manif_learn=False: PathNets
    forward (their graph is kept, as in the reference), KPCN trained on [in | mean p | var p]; only `dncnn` is
    clipped and stepped.  Returns the loss dict."""
    l1 = nn.functional.l1_loss
    models["backbone_diffuse"].zero_grad()
    models["backbone_specular"].zero_grad()
    loss = {}
    if manif_learn:
        p = {"diffuse": models["backbone_diffuse"](batch), "specular": models["backbone_specular"](batch)}
        Lm_d = feature_mse(p["diffuse"], batch["target_diffuse"], non_local) * w_manif
        Lm_s = feature_mse(p["specular"], batch["target_specular"], non_local) * w_manif
        loss["l_manif_diffuse"], loss["l_manif_specular"] = Lm_d.detach() / w_manif, Lm_s.detach() / w_manif
        Lm_d.backward()
        Lm_s.backward()
    else:
        models["dncnn"].zero_grad()
        p = {"diffuse": models["backbone_diffuse"](batch), "specular": models["backbone_specular"](batch)}
        nb = dict(batch)
        nb["kpcn_diffuse_in"] = pbuffer_concat(batch["kpcn_diffuse_in"], p["diffuse"])
        nb["kpcn_specular_in"] = pbuffer_concat(batch["kpcn_specular_in"], p["specular"])
        out = models["dncnn"](nb)
        total, diffuse, specular = out["radiance"], out["diffuse"], out["specular"]
        tgt_total = crop_like(batch["target_total"], total)
        if train_branches:
            L_d = l1(diffuse, crop_like(batch["target_diffuse"], diffuse))
            L_s = l1(specular, crop_like(batch["target_specular"], specular))
            loss["l_diffuse"], loss["l_specular"] = L_d.detach(), L_s.detach()
            L_d.backward()
            L_s.backward()
            with torch.no_grad():
                loss["l_total"] = l1(total, tgt_total)
        else:
            L_t = l1(total, tgt_total)
            loss["l_total"] = L_t.detach()
            L_t.backward()
    for k, v in loss.items():
        if not torch.isfinite(v).all():
            raise RuntimeError("%s: Non-finite loss at train time." % k)
    trained = [n for n in models if (("backbone" in n) if manif_learn else ("dncnn" in n))]
    for n in trained:
        nn.utils.clip_grad_value_(models[n].parameters(), clip_value=1.0)
    for n in trained:
        optims["optim_" + n].step()
    return loss
