"""B200 drop-in for the `sbmc.modules` surface WCMC uses (`from sbmc import modules as ops`,
/root/reference/support/networks.py:5, :18-24): ConvChain, Autoencoder, KernelApply.

Same constructor keywords, child-module names and state-dict keys as the upstream package
(SURVEY.md Appendix A.1-A.4), so checkpoints and `str(model)` look alike; the arithmetic runs in
libwcmc.so (tcgen05 implicit-GEMM convolutions, fused softmax + kernel-apply).  The nn.Conv2d
children only HOLD the parameters (and torch's weight-norm reparametrisation); they are never
called.  CUDA only - there is no CPU fallback.
"""
import math
import threading

import torch
import torch.nn as nn

from wcmc_b200 import ops

__all__ = ["ConvChain", "Autoencoder", "KernelApply", "crop_like"]

_SUPPORTED_ACT = ("relu", "leaky_relu", "linear", None)


def crop_like(src, tgt):
    """Centred crop of the last two dims of `src` to those of `tgt` (support/utils.py:24-42)."""
    dh = src.shape[-2] - tgt.shape[-2]
    dw = src.shape[-1] - tgt.shape[-1]
    top, left = max(dh // 2, 0), max(dw // 2, 0)
    bottom, right = dh - top, dw - left
    if top > 0 or left > 0 or bottom > 0 or right > 0:
        return src[..., top:src.shape[-2] - bottom, left:src.shape[-1] - right]
    return src


def _check_act(name):
    if name not in _SUPPORTED_ACT:
        raise NotImplementedError("activation %r has no fused epilogue in libwcmc.so (the WCMC hot path "
                                  "only uses relu / leaky_relu / linear)" % (name,))


def _gain(name):
    return nn.init.calculate_gain("linear" if name is None else name)


def _make_conv(cin, cout, ksize, padding, gain_of, weight_norm):
    conv = nn.Conv2d(cin, cout, ksize, stride=1, padding=padding, bias=True)
    nn.init.xavier_uniform_(conv.weight, gain=_gain(gain_of))
    nn.init.zeros_(conv.bias)
    if weight_norm:
        conv = nn.utils.weight_norm(conv)
    return conv


_WN = threading.local()   # .override = {id(conv): effective weight} while batched_weight_norm() is active (per thread:
                          # nn.DataParallel runs one replica per Python thread)


def _effective_weight(conv):
    override = getattr(_WN, "override", None)
    if override is not None and id(conv) in override:
        return override[id(conv)]
    if hasattr(conv, "weight_g"):
        return torch._weight_norm(conv.weight_v, conv.weight_g, 0)
    return conv.weight


class batched_weight_norm:
    """Context manager: the effective weights of every weight-normalised convolution under `module` come from ONE
    launch (ops.BatchedWeightNormFn) instead of one torch._weight_norm kernel per layer and direction; the `spec()`
    calls made inside the context pick them up."""

    def __init__(self, module):
        self.convs = [m for m in module.modules() if isinstance(m, nn.Conv2d) and hasattr(m, "weight_g")]

    def __enter__(self):
        self.prev = getattr(_WN, "override", None)
        if self.convs and self.convs[0].weight_v.is_cuda:
            flat = []
            for c in self.convs:
                flat += [c.weight_v, c.weight_g]
            ws = ops.BatchedWeightNormFn.apply(*flat)
            merged = dict(self.prev or {})
            merged.update({id(c): w for c, w in zip(self.convs, ws)})
            _WN.override = merged
        return self

    def __exit__(self, *exc):
        _WN.override = self.prev
        return False


class _ConvAct(nn.Module):
    """Parameter holder named like upstream's block: `layer.conv` (+ `layer.activation`)."""

    def __init__(self, cin, cout, ksize, padding, activation, weight_norm):
        super().__init__()
        self.layer = nn.Sequential()
        self.layer.add_module("conv", _make_conv(cin, cout, ksize, padding, activation, weight_norm))
        if activation not in ("linear", None):
            self.layer.add_module("activation", nn.ReLU() if activation == "relu" else nn.LeakyReLU(0.01))


class ConvChain(nn.Module):
    def __init__(self, ninputs, noutputs, ksize=3, width=64, depth=3, stride=1, pad=True, normalize=False,
                 normalization_type="batch", output_type="linear", activation="relu", weight_norm=True):
        super().__init__()
        if depth <= 0:
            raise ValueError("negative network depth.")
        if normalize:
            raise NotImplementedError("normalize=True is never used on the WCMC hot path")
        if stride != 1:
            raise NotImplementedError("only stride 1 is implemented")
        if ksize not in (1, 3, 5):
            raise NotImplementedError("libwcmc.so convolutions support ksize 1, 3, 5 (got %d)" % ksize)
        _check_act(activation)
        _check_act(output_type)
        padding = ksize // 2 if pad else 0
        self.ninputs, self.noutputs, self.ksize, self.padding, self.depth = ninputs, noutputs, ksize, padding, depth
        self.activation, self.output_type = activation, output_type
        cin = ninputs
        for i in range(depth - 1):
            self.add_module("layer_%d" % i, _ConvAct(cin, width, ksize, padding, activation, weight_norm))
            cin = width
        self.add_module("prediction", _make_conv(cin, noutputs, ksize, padding, output_type, weight_norm))
        if output_type not in ("linear", None):
            self.add_module("output_activation", nn.ReLU() if output_type == "relu" else nn.LeakyReLU(0.01))

    def _convs(self):
        return [getattr(self, "layer_%d" % i).layer.conv for i in range(self.depth - 1)] + [self.prediction]

    def spec(self):
        """(list of ops.LayerSpec, flat [w0, b0, w1, b1, ...] effective parameters)."""
        layers, params = [], []
        convs = self._convs()
        for i, conv in enumerate(convs):
            act = self.output_type if i == len(convs) - 1 else self.activation
            layers.append(ops.LayerSpec(conv.in_channels, conv.out_channels, self.ksize, self.padding,
                                        ops.ACT[act]))
            params += [_effective_weight(conv), conv.bias]
        return layers, params

    def forward(self, x):
        layers, params = self.spec()
        return ops.apply(ops.ConvChainFn, x, layers, *params)


class _UNetLevel(nn.Module):
    def __init__(self, n_in, n_out, level, num_levels, ksize, width, num_convs, max_width, increase_factor,
                 output_type, activation, pooling):
        super().__init__()
        self.is_last = level == num_levels - 1
        w = min(int(width * (increase_factor ** level)), max_width)
        if self.is_last:
            self.left = ConvChain(n_in, n_out, ksize=ksize, width=w, depth=num_convs, pad=True,
                                  output_type=activation, activation=activation)
            return
        self.left = ConvChain(n_in, w, ksize=ksize, width=w, depth=num_convs, pad=True, output_type=activation,
                              activation=activation)
        if pooling != "max":
            raise NotImplementedError("only max pooling is implemented (PathNet uses pooling='max')")
        self.downsample = nn.MaxPool2d(2, 2)
        w_next = min(int(width * (increase_factor ** (level + 1))), max_width)
        self.next_level = _UNetLevel(w, w_next, level + 1, num_levels, ksize, width, num_convs, max_width,
                                     increase_factor, activation, activation, pooling)
        self.right = ConvChain(w_next + w, n_out, ksize=ksize, width=w, depth=num_convs, pad=True,
                               output_type=output_type, activation=activation)

    def spec(self):
        l_layers, l_params = self.left.spec()
        if self.is_last:
            return ops.UNetSpec(left=l_layers), l_params
        n_spec, n_params = self.next_level.spec()
        r_layers, r_params = self.right.spec()
        return ops.UNetSpec(left=l_layers, right=r_layers, nxt=n_spec), l_params + n_params + r_params


class Autoencoder(nn.Module):
    def __init__(self, ninputs, noutputs, ksize=3, width=64, num_levels=3, num_convs=2, max_width=512,
                 increase_factor=1.0, normalize=False, normalization_type="batch", output_type="linear",
                 activation="relu", pooling="max"):
        super().__init__()
        if normalize:
            raise NotImplementedError("normalize=True is never used on the WCMC hot path")
        self.num_levels = num_levels
        self.unet = _UNetLevel(ninputs, noutputs, 0, num_levels, ksize, width, num_convs, max_width,
                               increase_factor, output_type, activation, pooling)

    def spec(self):
        return self.unet.spec()

    def forward(self, x):
        div = 2 ** (self.num_levels - 1)
        if x.shape[-2] % div or x.shape[-1] % div:
            raise ValueError("Autoencoder: spatial size %s must be divisible by %d (2x pooling / exact 2x "
                             "bilinear up-sampling per level)" % (tuple(x.shape[-2:]), div))
        spec, params = self.spec()
        return ops.apply(ops.AutoencoderFn, x, spec, *params)


class KernelApply(nn.Module):
    """softmax over the k*k axis + per-pixel weighted gather; returns (output, sum_w)."""

    def __init__(self, softmax=True, splat=False):
        super().__init__()
        if splat:
            raise NotImplementedError("splat=True is SBMC-only (out of scope for the KPCN hot path)")
        if not softmax:
            raise NotImplementedError("softmax=False is never used by sbmc.KPCN")
        self.softmax, self.splat = softmax, splat

    def forward(self, data, kernels):
        bs, k2, h, w = kernels.shape
        if tuple(data.shape[-2:]) != (h, w):
            raise ValueError("data and kernels must share spatial size")
        k = int(math.isqrt(k2))
        if k * k != k2 or k % 2 == 0 or k > 21:
            raise ValueError("kernel size must be an odd square <= 21x21 (got %d channels)" % k2)
        out = ops.KernelApplyFn.apply(data, kernels, k)
        # softmax weights sum to one; upstream returns the Halide op's sum_w
        return out, torch.ones((bs, h, w), dtype=out.dtype, device=out.device)
