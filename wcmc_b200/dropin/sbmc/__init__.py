"""B200 drop-in for `from sbmc import KPCN` (/root/reference/train_kpcn.py:30, :213, :229)."""
import torch
import torch.nn as nn

from wcmc_b200 import ops

from . import modules
from .modules import ConvChain, KernelApply, crop_like

__all__ = ["KPCN", "modules"]


class KPCN(nn.Module):
    """Kernel-predicting denoiser (Bako et al. 2017) as shipped by sbmc: two 9-layer 5x5 valid-conv
    chains predict 21x21 softmax kernels that filter the noisy diffuse / specular buffers
    (SURVEY.md 3.3, Appendix A.2/A.6).  Each branch runs as ONE fused autograd Function:
    tcgen05 convolutions -> fp32 logits -> fused softmax + kernel-apply."""

    def __init__(self, n_in, ksize=21, depth=9, width=100):
        super().__init__()
        self.n_in, self.ksize, self.depth, self.width = n_in, ksize, depth, width
        self.diffuse = ConvChain(n_in, ksize * ksize, depth=depth, width=width, ksize=5, activation="relu",
                                 weight_norm=False, pad=False, output_type="linear")
        self.specular = ConvChain(n_in, ksize * ksize, depth=depth, width=width, ksize=5, activation="relu",
                                  weight_norm=False, pad=False, output_type="linear")
        self.kernel_apply = KernelApply(softmax=True, splat=False)

    def _branch(self, chain, x, buf):
        shrink = self.depth * (chain.ksize - 1)
        ho, wo = x.shape[-2] - shrink, x.shape[-1] - shrink
        if ho <= 0 or wo <= 0:
            raise ValueError("input %s too small for %d valid %dx%d convolutions" %
                             (tuple(x.shape[-2:]), self.depth, chain.ksize, chain.ksize))
        tgt = torch.empty((0, 0, ho, wo), device="meta")
        layers, params = chain.spec()
        return ops.apply(ops.KPCNBranchFn, x, crop_like(buf, tgt), self.ksize, layers, *params)

    def forward(self, data):
        from wcmc_b200 import streams
        with streams.fork("diffuse"):     # independent branches on two streams (see wcmc_b200/streams.py)
            r_d = self._branch(self.diffuse, data["kpcn_diffuse_in"], data["kpcn_diffuse_buffer"])
        with streams.fork("specular"):
            r_s = self._branch(self.specular, data["kpcn_specular_in"], data["kpcn_specular_buffer"])
        # The caller's stream idles until the join: a caller may hang work here that does not depend on the branches
        # (KPCNInterface: the two path-disentangling losses, which only need the p-buffers and the output SIZE).
        hook = self.__dict__.get("while_branches_run")
        if hook is not None:
            hook(tuple(r_d.shape[-2:]))
        streams.join()
        radiance = ops.RecombineFn.apply(data["kpcn_albedo"], r_d, r_s)     # albedo * r_d + exp(r_s) - 1, centred crop
        return dict(radiance=radiance, diffuse=r_d, specular=r_s)
