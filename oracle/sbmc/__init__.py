"""ORACLE (test infrastructure): CPU restatement of the external ``sbmc`` package surface
that WCMC uses (`/root/reference/train_kpcn.py:28-33`: ``from sbmc import KPCN``;
`/root/reference/support/networks.py:5`: ``from sbmc import modules as ops``).

Parity unpinned - see ``oracle/sbmc/modules.py`` header.
"""
import torch
import torch.nn as nn

from . import modules
from .modules import ConvChain, KernelApply, crop_like

__all__ = ["KPCN", "modules"]


class KPCN(nn.Module):
    """two 9-layer 5x5 valid-conv chains predicting 21x21
    softmax kernels that are applied to the (cropped) noisy diffuse / specular buffers."""

    def __init__(self, n_in, ksize=21, depth=9, width=100):
        super().__init__()
        self.n_in, self.ksize, self.depth, self.width = n_in, ksize, depth, width
        self.diffuse = ConvChain(n_in, ksize * ksize, depth=depth, width=width, ksize=5,
                                 activation="relu", weight_norm=False, pad=False,
                                 output_type="linear")
        self.specular = ConvChain(n_in, ksize * ksize, depth=depth, width=width, ksize=5,
                                  activation="relu", weight_norm=False, pad=False,
                                  output_type="linear")
        self.kernel_apply = KernelApply(softmax=True, splat=False)

    def forward(self, data):
        k_d = self.diffuse(data["kpcn_diffuse_in"])
        k_s = self.specular(data["kpcn_specular_in"])
        b_d = crop_like(data["kpcn_diffuse_buffer"], k_d).contiguous()
        b_s = crop_like(data["kpcn_specular_buffer"], k_s).contiguous()
        r_d, _ = self.kernel_apply(b_d, k_d)
        r_s, _ = self.kernel_apply(b_s, k_s)
        albedo = crop_like(data["kpcn_albedo"], r_d)
        radiance = albedo * r_d + torch.exp(r_s) - 1.0
        return dict(radiance=radiance, diffuse=r_d, specular=r_s)
