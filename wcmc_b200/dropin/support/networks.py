"""B200 drop-in for /root/reference/support/networks.py: the path-embedding network."""
import torch.nn as nn

from sbmc import modules as ops
from wcmc_b200 import ops as wops


class PathNet(nn.Module):
    """Path embedding network (networks.py:7-42): per-sample 1x1 MLP -> mean over spp -> 3-level
    3x3 U-Net -> per-sample 1x1 MLP on [sample feature | propagated pixel feature].

    Same constructor, attributes and sub-module wiring as the reference; forward() runs the whole
    network as one fused autograd Function on NHWC bf16 buffers (the reference's repeat + cat of a
    (B*S,128,H,W) tensor, networks.py:39-40, is replaced by channel-slice writes)."""

    def __init__(self, ic, intermc=64, outc=3):
        super().__init__()
        self.ic, self.intermc, self.outc = ic, intermc, outc
        self.final_ic = intermc + intermc
        self.embedding = ops.ConvChain(ic, intermc, width=intermc, depth=3, ksize=1, pad=False)
        self.propagation = ops.Autoencoder(intermc, intermc, num_levels=3, increase_factor=2.0, num_convs=3,
                                           width=intermc, ksize=3, output_type="leaky_relu", pooling="max")
        self.final = ops.ConvChain(self.final_ic, outc, width=self.final_ic, depth=2, ksize=1, pad=False,
                                   output_type="relu")

    def __str__(self):
        return "PathNet i{}in{}o{}".format(self.ic, self.intermc, self.outc)

    def forward(self, samples):
        paths = samples["paths"]
        if paths.shape[-2] % 4 or paths.shape[-1] % 4:
            raise ValueError("PathNet: spatial size %s must be divisible by 4" % (tuple(paths.shape[-2:]),))
        with ops.batched_weight_norm(self):   # all 20 layers' g * v / ||v|| in one launch (and one in backward)
            e_layers, e_params = self.embedding.spec()
            u_spec, u_params = self.propagation.spec()
            f_layers, f_params = self.final.spec()
        spec = wops.PathNetSpec(embedding=e_layers, unet=u_spec, final=f_layers)
        return wops.apply(wops.PathNetFn, paths, spec, *(e_params + u_params + f_params))
