"""ORACLE (test infrastructure, never shipped on the product path).

Pure-PyTorch CPU restatement of the un-vendored ``sbmc.modules`` pieces that WCMC's
hot path calls: ``ConvChain``, ``Autoencoder``, ``KernelApply`` and the native
``kernel_weighting`` op.  The reference imports them from an external, un-pinned package
(`/root/reference/support/networks.py:4-5`, `:18-24`; `train_kpcn.py:28-33`), so this
restatement follows SURVEY.md Appendix A.  **Parity unpinned** for these pieces: the
reference repository holds no tests or golden vectors for them.

Choices this oracle fixes (SURVEY §8c "unpinned choices"):
  1. kernel index k = dy*K + dx, offset (dy - K//2, dx - K//2), cross-correlation;
  2. the gather reads zero outside the (already cropped) data tensor;
  3. xavier-uniform init with activation gain, zero bias;
  4. ``weight_norm=True`` default (PathNet chains), off for KPCN chains;
  5. U-Net widths 64/128/256, relu inside, top ``right`` chain output ``leaky_relu(0.01)``;
  6. bilinear up-sampling with ``align_corners=False``.
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

__all__ = ["ConvChain", "Autoencoder", "KernelApply", "kernel_weighting", "crop_like"]


def crop_like(src, tgt):
    """Centred crop of the last two dims (follows reference support/utils.py:24-42)."""
    dh = src.shape[-2] - tgt.shape[-2]
    dw = src.shape[-1] - tgt.shape[-1]
    ch, cw = max(dh // 2, 0), max(dw // 2, 0)
    ch2, cw2 = dh - ch, dw - cw
    if ch > 0 or cw > 0 or ch2 > 0 or cw2 > 0:
        return src[..., ch:src.shape[-2] - ch2, cw:src.shape[-1] - cw2]
    return src


def _activation(name):
    if name == "relu":
        return nn.ReLU(inplace=False)
    if name == "leaky_relu":
        return nn.LeakyReLU(negative_slope=0.01, inplace=False)
    if name == "sigmoid":
        return nn.Sigmoid()
    if name == "tanh":
        return nn.Tanh()
    if name == "elu":
        return nn.ELU()
    if name == "softplus":
        return nn.Softplus()
    if name in ("linear", None):
        return None
    raise ValueError("unknown activation %s" % name)


def _gain(name):
    if name in ("elu", "softplus"):
        name = "relu"
    if name is None:
        name = "linear"
    return nn.init.calculate_gain(name)


class _ConvAct(nn.Module):
    """conv (+activation) block; child ``layer`` holds ``conv`` and ``activation``."""

    def __init__(self, cin, cout, ksize, stride, padding, activation, weight_norm):
        super().__init__()
        conv = nn.Conv2d(cin, cout, ksize, stride=stride, padding=padding, bias=True)
        nn.init.xavier_uniform_(conv.weight, gain=_gain(activation))
        nn.init.zeros_(conv.bias)
        if weight_norm:
            conv = nn.utils.weight_norm(conv)
        self.layer = nn.Sequential()
        self.layer.add_module("conv", conv)
        act = _activation(activation)
        if act is not None:
            self.layer.add_module("activation", act)

    def forward(self, x):
        return self.layer(x)


class ConvChain(nn.Module):
    """(depth-1) x (conv + act) then ``prediction`` conv (+ output act)."""

    def __init__(self, ninputs, noutputs, ksize=3, width=64, depth=3, stride=1, pad=True,
                 normalize=False, normalization_type="batch", output_type="linear",
                 activation="relu", weight_norm=True):
        super().__init__()
        if depth <= 0:
            raise ValueError("negative network depth.")
        if normalize:
            raise NotImplementedError("normalize=True is never used on the WCMC hot path")
        padding = ksize // 2 if pad else 0
        self.ksize, self.padding, self.depth = ksize, padding, depth
        self.activation, self.output_type = activation, output_type
        cin = ninputs
        for i in range(depth - 1):
            self.add_module("layer_%d" % i,
                            _ConvAct(cin, width, ksize, stride, padding, activation, weight_norm))
            cin = width
        pred = nn.Conv2d(cin, noutputs, ksize, stride=stride, padding=padding, bias=True)
        nn.init.xavier_uniform_(pred.weight, gain=_gain(output_type))
        nn.init.zeros_(pred.bias)
        if weight_norm:
            pred = nn.utils.weight_norm(pred)
        self.add_module("prediction", pred)
        act = _activation(output_type)
        if act is not None:
            self.add_module("output_activation", act)

    def forward(self, x):
        for m in self.children():
            x = m(x)
        return x


class _UNetLevel(nn.Module):
    def __init__(self, n_in, n_out, level, num_levels, ksize, width, num_convs, max_width,
                 increase_factor, output_type, activation, pooling):
        super().__init__()
        self.is_last = level == num_levels - 1
        w = min(int(width * (increase_factor ** level)), max_width)
        if self.is_last:
            self.left = ConvChain(n_in, n_out, ksize=ksize, width=w, depth=num_convs,
                                  pad=True, output_type=activation, activation=activation)
            return
        self.left = ConvChain(n_in, w, ksize=ksize, width=w, depth=num_convs, pad=True,
                              output_type=activation, activation=activation)
        if pooling == "max":
            self.downsample = nn.MaxPool2d(2, 2)
        elif pooling == "average":
            self.downsample = nn.AvgPool2d(2, 2)
        else:
            raise ValueError("unknown pooling %s" % pooling)
        w_next = min(int(width * (increase_factor ** (level + 1))), max_width)
        self.next_level = _UNetLevel(w, w_next, level + 1, num_levels, ksize, width, num_convs,
                                     max_width, increase_factor, activation, activation, pooling)
        self.right = ConvChain(w_next + w, n_out, ksize=ksize, width=w, depth=num_convs,
                               pad=True, output_type=output_type, activation=activation)

    def forward(self, x):
        left = self.left(x)
        if self.is_last:
            return left
        ds = self.downsample(left)
        nxt = self.next_level(ds)
        us = F.interpolate(nxt, size=left.shape[-2:], mode="bilinear", align_corners=False)
        return self.right(torch.cat([us, left], 1))


class Autoencoder(nn.Module):
    """SURVEY Appendix A.3 recursive U-Net."""

    def __init__(self, ninputs, noutputs, ksize=3, width=64, num_levels=3, num_convs=2,
                 max_width=512, increase_factor=1.0, normalize=False,
                 normalization_type="batch", output_type="linear", activation="relu",
                 pooling="max"):
        super().__init__()
        if normalize:
            raise NotImplementedError
        self.unet = _UNetLevel(ninputs, noutputs, 0, num_levels, ksize, width, num_convs,
                               max_width, increase_factor, output_type, activation, pooling)

    def forward(self, x):
        return self.unet(x)


def kernel_weighting(data, weights):
    """SURVEY Appendix A.5 (Halide ``kernel_weighting``): gather form, zero exterior.

    data (B,C,H,W), weights (B,kh,kw,H,W) -> out (B,C,H,W), sum_w (B,H,W)
    out[b,c,y,x] = sum_{dy,dx} w[b,dy,dx,y,x] * data0[b,c,y+dy-kh//2,x+dx-kw//2]
    """
    b, c, h, w = data.shape
    kh, kw = weights.shape[1], weights.shape[2]
    unf = F.unfold(data, (kh, kw), padding=(kh // 2, kw // 2))  # (B, C*kh*kw, H*W)
    unf = unf.view(b, c, kh * kw, h, w)
    wts = weights.reshape(b, 1, kh * kw, h, w)
    out = (unf * wts).sum(2)
    return out, weights.sum((1, 2))


class KernelApply(nn.Module):
    """SURVEY Appendix A.4."""

    def __init__(self, softmax=True, splat=False):
        super().__init__()
        if splat:
            raise NotImplementedError("splat=True is SBMC-only (out of scope)")
        self.softmax = softmax
        self.splat = splat

    def forward(self, data, kernels):
        bs, k2, h, w = kernels.shape
        if data.shape[-2:] != kernels.shape[-2:]:
            raise ValueError("data and kernels must share spatial size")
        k = int(math.isqrt(k2))
        if self.softmax:
            kernels = F.softmax(kernels, dim=1)
        kernels = kernels.view(bs, k, k, h, w)
        return kernel_weighting(data, kernels)
