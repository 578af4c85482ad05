"""Host-side operators of the B200 backend: autograd Functions that drive libwcmc.so.

Public tensors stay in the reference's contract (NCHW fp32, `support/datasets.py:760-793`);
inside a Function everything is NHWC bf16 with channels padded to 16 and "concatenation" is a
channel-slice write (see include/wcmc.h).  Three fused operators cover the hot path:

  ConvChainFn    sbmc.modules.ConvChain                       
  KPCNBranchFn   ConvChain(9 x 5x5) -> softmax -> 21x21 kernel-apply   (sbmc.KPCN branch, A.2/A.4)
  PathNetFn      embedding MLP -> spp mean -> U-Net -> broadcast -> final MLP
                 (/root/reference/support/networks.py:29-42)
  AutoencoderFn  the U-Net alone (sbmc.modules.Autoencoder, A.3)

Precision policy: 16-bit tensor-core operands with fp32 accumulation in TMEM.  Activations, packed
weights AND the gradients flowing between layers are fp16 (11-bit significand: bf16's 8 bits put the
gradients at 1.5-3e-2 relative error through the 9-layer stack, outside north_star's 1e-2; the tensor
cores trap on mixed f16 x bf16 operands, so one format is used throughout).  fp16's narrow exponent
is handled by a per-call loss scale: every backward pass multiplies the incoming gradient by
s = 256 / max|g| (computed on the device, no host sync), keeps all 16-bit gradient tensors scaled by
s, and multiplies the fp32 results (weight / bias / input gradients) by 1/s.  fp32 master weights /
optimiser state / biases, fp32 logits + softmax + kernel-apply, fp32 PathNet output, fp32
weight-gradient accumulation.  WCMC_ACT_DTYPE=bf16 switches every 16-bit tensor to bf16.
"""
import os
import threading
from dataclasses import dataclass
from typing import List, Optional

import torch

from . import lib

ACT = {"linear": 0, None: 0, "relu": 1, "leaky_relu": 2}
ACT_DTYPE = {"f16": torch.float16, "fp16": torch.float16, "bf16": torch.bfloat16}[
    os.environ.get("WCMC_ACT_DTYPE", "f16").lower()]
GRAD_DTYPE = ACT_DTYPE
GRAD_TARGET = 256.0  # max |scaled incoming gradient|
DEBUG_TAP = None     # tests may set this to a list to capture the per-layer dz tensors


_CALL = threading.local()


def apply(fn, *args):
    """fn.apply(*args) with the caller's grad mode visible to fn.forward: inside a Function's forward grad mode is
    always off and `ctx.needs_input_grad` is True for parameters even under torch.no_grad(), so without this a
    validation pass would save (and write to HBM) every activation a backward pass would need."""
    prev = getattr(_CALL, "grad", True)
    _CALL.grad = torch.is_grad_enabled()
    try:
        return fn.apply(*args)
    finally:
        _CALL.grad = prev


def _needs_backward(ctx):
    return getattr(_CALL, "grad", True) and any(ctx.needs_input_grad)


def grad_scale(g):
    """-> (s, 1/s) as 1-element fp32 device tensors, s = GRAD_TARGET / max|g| (one launch, no host sync)."""
    g = g.detach()
    if g.dtype != torch.float32 or not g.is_contiguous():
        g = g.float().contiguous()
    out = lib.absmax_scale(g, GRAD_TARGET)
    return out[0:1], out[1:2]
LEAKY_SLOPE = 0.01


@dataclass
class LayerSpec:
    cin: int
    cout: int
    ksize: int
    pad: int
    act: int  # 0 linear, 1 relu, 2 leaky

    @property
    def cin_p(self):
        return lib.pad16(self.cin)

    @property
    def cout_p(self):
        return lib.pad16(self.cout)


@dataclass
class Slice:
    """Channel slice [coff, coff+c) of an NHWC bf16 (or fp32) tensor."""
    t: torch.Tensor
    coff: int
    c: int


def pack_chain(layers: List[LayerSpec], params, need_dgrad=True):
    """params = [w0, b0, w1, b1, ...] (torch layout fp32) -> [(w_fwd, w_dgrad, bias_p)]."""
    specs = [(params[2 * i], params[2 * i + 1], l.cout_p, l.cin_p) for i, l in enumerate(layers)]
    return lib.pack_weights_batch(specs, dtype=ACT_DTYPE, dgrad=need_dgrad)


def chain_forward(x: Slice, layers, packed, out: Optional[Slice] = None, last_fp32=False):
    """Runs the conv stack; returns the list of activations [x, y0, y1, ...] as Slices."""
    acts = [x]
    cur = x
    for i, l in enumerate(layers):
        wf, _, bp = packed[i]
        last = i == len(layers) - 1
        assert cur.c == l.cin, "chain input has %d channels, layer %d expects %d" % (cur.c, i, l.cin)
        if last and out is not None:
            y = lib.conv2d(cur.t, wf, bp, l.ksize, l.pad, act=l.act, slope=LEAKY_SLOPE, x_coff=cur.coff, out=out.t,
                           out_coff=out.coff, cin=l.cin, cout=l.cout)
            cur = Slice(y, out.coff, l.cout)
        else:
            y = lib.conv2d(cur.t, wf, bp, l.ksize, l.pad, act=l.act, slope=LEAKY_SLOPE, x_coff=cur.coff,
                           out_dtype=torch.float32 if (last and last_fp32) else None, cin=l.cin, cout=l.cout)
            cur = Slice(y, 0, l.cout)
        acts.append(cur)
    return acts


def chain_backward(dz: Slice, acts, layers, packed, need_dx, want_param_grads=True, inv_scale=None, db_pool=None):
    """dz = (loss-scaled) gradient w.r.t. the PRE-activation output of the last layer (16-bit NHWC
    slice).  Returns (dx Slice or None [still scaled], [dw0, db0, dw1, db1, ...] [un-scaled fp32]).
    db_pool: optional [zero-filled fp32 tensor, offset]: the bias gradients of this chain are carved out of it (the
    U-Net's five chains share one memset instead of one fill launch each)."""
    grads = [None] * (2 * len(layers))
    # bias gradients of layers 0..L-2 come for free from the epilogue of the data-gradient launch that
    # produces their dz (column sums); the last layer's is a column-sum kernel accumulating into the same
    # zero-filled buffer (one memset per chain instead of a zero kernel per bias)
    db_all = None
    if want_param_grads:
        offs = [0]
        for l in layers:
            offs.append(offs[-1] + l.cout_p)
        if db_pool is not None:
            db_all = db_pool[0][db_pool[1]:db_pool[1] + offs[-1]]
            db_pool[1] += offs[-1]
        else:
            db_all = torch.zeros(offs[-1], dtype=torch.float32, device=dz.t.device)
    for i in range(len(layers) - 1, -1, -1):
        l = layers[i]
        xin = acts[i]
        if DEBUG_TAP is not None:
            DEBUG_TAP.append((i, dz.t.detach().clone(), dz.coff, l.cout, inv_scale))
        if want_param_grads:
            grads[2 * i] = lib.conv2d_wgrad(xin.t, dz.t, l.cout, l.cin, l.ksize, l.pad, l.cin_p, l.cout_p,
                                            x_coff=xin.coff, dy_coff=dz.coff, scale=inv_scale, defer=True)
            grads[2 * i + 1] = db_all[offs[i]:offs[i] + l.cout]
            if i == len(layers) - 1:
                lib.bias_grad(dz.t, l.cout, dy_coff=dz.coff, out=grads[2 * i + 1], accumulate=True, scale=inv_scale)
        if i == 0 and not need_dx:
            return None, grads
        wd = packed[i][1]
        mask = None
        slope = 0.0
        if i > 0 and layers[i - 1].act != 0:
            mask = xin
            slope = LEAKY_SLOPE if layers[i - 1].act == 2 else 0.0
        d = lib.conv2d(dz.t, wd, None, l.ksize, l.ksize - 1 - l.pad, act=0, x_coff=dz.coff,
                       mask=None if mask is None else mask.t, mask_coff=0 if mask is None else mask.coff,
                       slope=slope, cin=l.cout, cout=l.cin,
                       colsum=db_all[offs[i - 1]:] if (db_all is not None and i > 0) else None,
                       colsum_scale=inv_scale, alg_hw=tuple(dz.t.shape[1:3]))
        dz = Slice(d, 0, l.cin)
    return dz, grads


def _to_nhwc(x, c_fill=None, dtype=None, scale=None):
    x = x.contiguous()
    if x.dtype != torch.float32:
        x = x.float()
    return Slice(lib.nchw_to_nhwc(x, c_fill=c_fill, dtype=dtype or ACT_DTYPE, scale=scale), 0, x.shape[1])


def _grad_nhwc(g, c_fill, scale=None):
    """fp32 NCHW gradient -> 16-bit NHWC, multiplied by the loss scale (channels zero padded)."""
    return _to_nhwc(g, c_fill, GRAD_DTYPE, scale)


# ------------------------------------------------------------------------------------------------
# weight normalisation of every convolution of a network in one launch per direction
# ------------------------------------------------------------------------------------------------
class BatchedWeightNormFn(torch.autograd.Function):
    """(v0, g0, v1, g1, ...) -> (w0, w1, ...) with w = g * v / ||v|| over all dims but 0 (torch._weight_norm(v, g, 0))."""

    @staticmethod
    def forward(ctx, *vg):
        vs = [t.detach().float().contiguous() for t in vg[0::2]]
        gs = [t.detach().float().contiguous() for t in vg[1::2]]
        ws, norms = lib.weight_norm_batch_fwd(vs, gs)
        ctx.saved = (vs, gs, norms)
        return tuple(ws)

    @staticmethod
    def backward(ctx, *dws):
        vs, gs, norms = ctx.saved
        dws = [torch.zeros_like(v) if d is None else d for d, v in zip(dws, vs)]
        dvs, dgs = lib.weight_norm_batch_bwd(dws, vs, gs, norms)
        out = []
        for dv, dg in zip(dvs, dgs):
            out += [dv, dg]
        return tuple(out)


# ------------------------------------------------------------------------------------------------
# ConvChain
# ------------------------------------------------------------------------------------------------
class ConvChainFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, layers, *params):
        need = _needs_backward(ctx)
        packed = pack_chain(layers, params, need_dgrad=need)
        xin = _to_nhwc(x, layers[0].cin_p)
        acts = chain_forward(xin, layers, packed)
        y = acts[-1]
        out = lib.nhwc_to_nchw(y.t, layers[-1].cout, y.coff)
        ctx.layers = layers
        ctx.need_dx = ctx.needs_input_grad[0]
        if need:
            ctx.acts = acts
            ctx.packed = packed
        return out

    @staticmethod
    def backward(ctx, g):
        layers = ctx.layers
        last = layers[-1]
        s, inv_s = grad_scale(g)
        dy = _grad_nhwc(g, last.cout_p, s)
        dz = _act_bwd_full(dy, ctx.acts[-1], last)
        dx, grads = chain_backward(dz, ctx.acts, layers, ctx.packed, ctx.need_dx, inv_scale=inv_s)
        gx = lib.nhwc_to_nchw(dx.t, layers[0].cin, dx.coff, scale=inv_s) if dx is not None else None
        lib.wgrad_flush()
        ctx.acts = ctx.packed = None
        return (gx, None) + tuple(grads)


def _act_bwd_full(dy: Slice, y: Slice, layer: LayerSpec):
    """dy covers the padded channel range of `layer`'s output; applies the output activation's
    derivative (no-op for linear)."""
    if layer.act == 0:
        return dy
    d = lib.act_bwd(dy.t, y.t, layer.cout_p, layer.act, LEAKY_SLOPE, dy_coff=dy.coff, y_coff=y.coff)
    return Slice(d, 0, dy.c)


# ------------------------------------------------------------------------------------------------
# KPCN branch: conv chain -> softmax -> kernel apply
# ------------------------------------------------------------------------------------------------
# inference: last layer -> softmax -> kernel-apply in ONE kernel (wcmc_conv2d_kernel_apply; 0 = separate launches)
FUSE_KERNEL_APPLY = os.environ.get("WCMC_FUSE_KA", "1") != "0"


class KPCNBranchFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, data, ksize, layers, *params):
        need = _needs_backward(ctx)
        packed = pack_chain(layers, params, need_dgrad=need)
        xin = _to_nhwc(x, layers[0].cin_p)
        last = layers[-1]
        if (FUSE_KERNEL_APPLY and not need and ksize == 21 and last.ksize == 5 and last.cout == 441 and last.act == 0
                and data.shape[1] == 3 and len(layers) > 1):
            # nothing is saved for a backward pass, so the 441 logits never have to exist in HBM
            acts = chain_forward(xin, layers[:-1], packed[:-1])
            cur = acts[-1]
            data = data.contiguous().float()
            ho, wo = cur.t.shape[1] + 2 * last.pad - 4, cur.t.shape[2] + 2 * last.pad - 4
            assert tuple(data.shape[-2:]) == (ho, wo), "data and kernels must share spatial size"
            wf, _, bp = packed[-1]
            ctx.layers, ctx.ksize, ctx.need_dx = layers, ksize, False
            return lib.conv2d_kernel_apply(cur.t, wf, bp, data, last.ksize, last.pad, ksize, x_coff=cur.coff,
                                           cin=last.cin, cout=last.cout)
        acts = chain_forward(xin, layers, packed, last_fp32=True)
        logits = acts[-1].t  # (N,Ho,Wo,cout_p) fp32
        data = data.contiguous().float()
        assert tuple(data.shape[-2:]) == tuple(logits.shape[1:3]), "data and kernels must share spatial size"
        out, stats = lib.kernel_apply_fwd(logits, data, ksize, want_stats=need)
        ctx.layers, ctx.ksize, ctx.need_dx = layers, ksize, ctx.needs_input_grad[0]
        if need:
            ctx.acts, ctx.packed, ctx.aux = acts, packed, (data, stats)
            ctx.save_for_backward(out)
        return out

    @staticmethod
    def backward(ctx, g):
        layers = ctx.layers
        data, stats = ctx.aux
        (out,) = ctx.saved_tensors
        logits = ctx.acts[-1].t
        g = g.contiguous().float()
        s, inv_s = grad_scale(g)
        dl = lib.kernel_apply_bwd(logits, data, out, stats, g, ctx.ksize, dl_cs=layers[-1].cout_p,
                                  dtype=GRAD_DTYPE, scale=s)
        dz = Slice(dl, 0, layers[-1].cout)
        dx, grads = chain_backward(dz, ctx.acts, layers, ctx.packed, ctx.need_dx, inv_scale=inv_s)
        gx = lib.nhwc_to_nchw(dx.t, layers[0].cin, dx.coff, scale=inv_s) if dx is not None else None
        lib.wgrad_flush()
        ctx.acts = ctx.packed = ctx.aux = None
        return (gx, None, None, None) + tuple(grads)


class KernelApplyFn(torch.autograd.Function):
    """Stand-alone softmax + kernel-apply on an NCHW logits tensor (module-level API)."""

    @staticmethod
    def forward(ctx, data, kernels, ksize):
        n, k2, h, w = kernels.shape
        cs = (k2 + 7) // 8 * 8
        logits = torch.zeros((n, h, w, cs), dtype=torch.float32, device=kernels.device)
        logits[..., :k2] = kernels.permute(0, 2, 3, 1)
        data = data.contiguous().float()
        need = _needs_backward(ctx)
        out, stats = lib.kernel_apply_fwd(logits, data, ksize, want_stats=need)
        ctx.k2 = k2
        ctx.ksize = ksize
        if need:
            ctx.save_for_backward(logits, data, out, stats)
        return out

    @staticmethod
    def backward(ctx, g):
        logits, data, out, stats = ctx.saved_tensors
        dl = lib.kernel_apply_bwd(logits, data, out, stats, g.contiguous().float(), ctx.ksize, dtype=torch.float32)
        return None, dl[..., :ctx.k2].permute(0, 3, 1, 2), None


# ------------------------------------------------------------------------------------------------
# U-Net (sbmc.modules.Autoencoder) on NHWC bf16 buffers
# ------------------------------------------------------------------------------------------------
@dataclass
class UNetSpec:
    """One level of the recursive U-Net: `left` chain, then (unless last) pool -> next -> up ->
    cat([up, left]) -> `right` chain.  Parameter order: left, next..., right."""
    left: List[LayerSpec]
    right: Optional[List[LayerSpec]] = None
    nxt: Optional["UNetSpec"] = None

    def n_params(self):
        n = 2 * len(self.left)
        if self.nxt is not None:
            n += self.nxt.n_params() + 2 * len(self.right)
        return n


def unet_layers(spec: UNetSpec):
    """Flat list of the LayerSpecs in parameter order (left, next level..., right)."""
    out = list(spec.left)
    if spec.nxt is not None:
        out += unet_layers(spec.nxt) + list(spec.right)
    return out


def unet_forward(spec: UNetSpec, x: Slice, params, need_dgrad=True, packed=None):
    """Returns (output Slice, ctx).  params is the flat [w,b,...] list in spec order; packed (optional) the
    already packed weights of every layer in the same order (one pack launch for the whole network)."""
    nl = 2 * len(spec.left)
    if packed is None:
        packed = pack_chain(unet_layers(spec), params, need_dgrad)
    p_left = packed[:nl // 2]
    if spec.nxt is None:
        acts = chain_forward(x, spec.left, p_left)
        return acts[-1], dict(acts_left=acts, p_left=p_left)
    n, h, w, _ = x.t.shape
    assert h % 2 == 0 and w % 2 == 0, "U-Net levels need even spatial sizes (got %dx%d)" % (h, w)
    c_left = spec.left[-1].cout
    c_up = spec.right[0].cin - c_left
    assert c_up % 8 == 0 and c_left % 8 == 0
    cat = torch.empty((n, h, w, c_up + c_left), dtype=ACT_DTYPE, device=x.t.device)
    acts_left = chain_forward(x, spec.left, p_left, out=Slice(cat, c_up, c_left))
    pooled = lib.maxpool2_fwd(cat, c_left, x_coff=c_up)
    nn_ = spec.nxt.n_params()
    y_next, ctx_next = unet_forward(spec.nxt, Slice(pooled, 0, c_left), params[nl:nl + nn_], need_dgrad,
                                    packed[nl // 2:(nl + nn_) // 2])
    assert y_next.c == c_up
    lib.upsample2_fwd(y_next.t, c_up, x_coff=y_next.coff, out=cat, out_coff=0)
    p_right = packed[(nl + nn_) // 2:]
    acts_right = chain_forward(Slice(cat, 0, c_up + c_left), spec.right, p_right)
    ctx = dict(acts_left=acts_left, p_left=p_left, cat=cat, pooled=pooled, ctx_next=ctx_next, y_next=y_next,
               acts_right=acts_right, p_right=p_right, c_up=c_up, c_left=c_left)
    return acts_right[-1], ctx


def unet_backward(spec: UNetSpec, ctx, dy: Slice, need_dx=True, inv_scale=None, db_pool=None):
    """dy = (loss-scaled) gradient w.r.t. the post-activation output.  Returns (dx Slice, flat grads)."""
    if db_pool is None:   # one zero-filled buffer for the bias gradients of every chain of the U-Net
        total = sum(l.cout_p for l in unet_layers(spec))
        db_pool = [torch.zeros(total, dtype=torch.float32, device=dy.t.device), 0]
    if spec.nxt is None:
        dz = _act_bwd_full(dy, ctx["acts_left"][-1], spec.left[-1])
        return chain_backward(dz, ctx["acts_left"], spec.left, ctx["p_left"], need_dx, inv_scale=inv_scale,
                              db_pool=db_pool)
    c_up, c_left = ctx["c_up"], ctx["c_left"]
    dz = _act_bwd_full(dy, ctx["acts_right"][-1], spec.right[-1])
    d_cat, g_right = chain_backward(dz, ctx["acts_right"], spec.right, ctx["p_right"], True, inv_scale=inv_scale,
                                    db_pool=db_pool)
    d_up = lib.upsample2_bwd(d_cat.t, c_up, dy_coff=d_cat.coff)
    d_pool, g_next = unet_backward(spec.nxt, ctx["ctx_next"], Slice(d_up, 0, c_up), True, inv_scale, db_pool)
    d_left = lib.maxpool2_bwd(ctx["cat"], d_pool.t, c_left, x_coff=c_up, dy_coff=d_pool.coff, add=d_cat.t,
                              add_coff=d_cat.coff + c_up)
    left_out = ctx["acts_left"][-1]
    dzl = _act_bwd_full(Slice(d_left, 0, c_left), left_out, spec.left[-1])
    dx, g_left = chain_backward(dzl, ctx["acts_left"], spec.left, ctx["p_left"], need_dx, inv_scale=inv_scale,
                                db_pool=db_pool)
    return dx, g_left + g_next + g_right


class AutoencoderFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, spec, *params):
        need = _needs_backward(ctx)
        xin = _to_nhwc(x, spec.left[0].cin_p)
        y, uctx = unet_forward(spec, xin, list(params), need)
        ctx.spec, ctx.need_dx, ctx.cin = spec, ctx.needs_input_grad[0], x.shape[1]
        ctx.uctx = uctx if need else None
        return lib.nhwc_to_nchw(y.t, y.c, y.coff)

    @staticmethod
    def backward(ctx, g):
        spec = ctx.spec
        c_out = g.shape[1]
        s, inv_s = grad_scale(g)
        dy = _grad_nhwc(g, lib.pad16(c_out), s)
        dx, grads = unet_backward(spec, ctx.uctx, dy, ctx.need_dx, inv_s)
        gx = lib.nhwc_to_nchw(dx.t, ctx.cin, dx.coff, scale=inv_s) if dx is not None else None
        lib.wgrad_flush()
        ctx.uctx = None
        return (gx, None) + tuple(grads)


# ------------------------------------------------------------------------------------------------
# PathNet (support/networks.py:29-42)
# ------------------------------------------------------------------------------------------------
@dataclass
class PathNetSpec:
    embedding: List[LayerSpec]
    unet: UNetSpec
    final: List[LayerSpec]


def fused_mlp_ok(spec):
    """The fused K6 / K7 kernels cover the shapes PathNet is built with (intermc = 64, networks.py:11-24)."""
    e, f = spec.embedding, spec.final
    return (len(e) == 3 and all(l.ksize == 1 and l.pad == 0 and l.cout == 64 for l in e) and e[0].cin <= 64
            and e[1].cin == 64 and e[2].cin == 64 and len(f) == 2 and all(l.ksize == 1 and l.pad == 0 for l in f)
            and f[0].cin == 128 and f[0].cout == 128 and f[1].cin == 128 and f[1].cout <= 32)


FUSED_MLP = os.environ.get("WCMC_FUSED_MLP", "1") != "0"
# K8 / K9: the backward passes of the two MLPs as one kernel each (0: generic 1x1 conv dgrad / wgrad launches)
FUSED_MLP_BWD = os.environ.get("WCMC_FUSED_MLP_BWD", "1") != "0"


class PathNetFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, paths, spec, *params):
        b, s, nf, h, w = paths.shape
        need = _needs_backward(ctx)
        ne, nu = 2 * len(spec.embedding), spec.unet.n_params()
        # one launch packs the weights of all 20 layers (embedding, U-Net, final)
        p_all = pack_chain(list(spec.embedding) + unet_layers(spec.unet) + list(spec.final), params, need)
        p_emb, p_unet, p_fin = p_all[:ne // 2], p_all[ne // 2:(ne + nu) // 2], p_all[(ne + nu) // 2:]
        c_emb = spec.embedding[-1].cout
        c_prop = spec.final[0].cin - c_emb
        outc = spec.final[-1].cout
        dev = paths.device
        if FUSED_MLP and fused_mlp_ok(spec):
            # K6: fp32 NCHW paths -> 3-layer MLP -> emb (+ spp mean) in one kernel; K7: [emb | prop] -> out
            px = paths.contiguous().float()
            fused_bwd = need and FUSED_MLP_BWD and c_emb == 64 and c_prop == 64
            both = torch.empty((b * s, h, w, c_emb + c_prop if (need and not fused_bwd) else c_emb), dtype=ACT_DTYPE,
                               device=dev)
            reduced = torch.empty((b, h, w, c_emb), dtype=ACT_DTYPE, device=dev)
            x16 = h1 = h2 = hfin = None
            if need:
                x16 = torch.empty((b * s, h, w, 64), dtype=ACT_DTYPE, device=dev)   # zero padded to 64 channels
                h1 = torch.empty((b * s, h, w, c_emb), dtype=ACT_DTYPE, device=dev)
                h2 = torch.empty((b * s, h, w, c_emb), dtype=ACT_DTYPE, device=dev)
                hfin = torch.empty((b * s, h, w, spec.final[0].cout), dtype=ACT_DTYPE, device=dev)
            lib.pathnet_embed_fwd(px, p_emb, [l.act for l in spec.embedding], LEAKY_SLOPE, both, 0, reduced,
                                  x16=x16, h1=h1, h2=h2)
            prop, uctx = unet_forward(spec.unet, Slice(reduced, 0, c_emb), list(params[ne:ne + nu]), need, p_unet)
            assert prop.c == c_prop
            out = lib.pathnet_final_fwd(both, 0, prop.t, prop.coff, p_fin, [l.act for l in spec.final], LEAKY_SLOPE,
                                        outc, b, s, hfin=hfin)
            if fused_bwd:
                # K8 / K9 read emb per sample and prop per pixel: no materialised [emb | prop] tensor at all
                prop_t = prop.t if (prop.coff == 0 and prop.t.is_contiguous()) else \
                    prop.t[..., prop.coff:prop.coff + c_prop].contiguous()
                acts_emb = ("fused", x16, h1, h2, both)
                acts_fin = ("fused", prop_t, hfin)
            elif need:
                # the backward pass (generic wgrad / dgrad kernels) reads [emb | prop] as one 128-channel tensor
                lib.spp_broadcast(prop.t, b, s, c_prop, 1.0, x_coff=prop.coff, out=both, out_coff=c_emb)
                acts_emb = [Slice(x16, 0, nf), Slice(h1, 0, c_emb), Slice(h2, 0, c_emb), Slice(both, 0, c_emb)]
                acts_fin = [Slice(both, 0, c_emb + c_prop), Slice(hfin, 0, spec.final[0].cout), None]
        else:
            x = _to_nhwc(paths.reshape(b * s, nf, h, w), spec.embedding[0].cin_p)
            both = torch.empty((b * s, h, w, c_emb + c_prop), dtype=ACT_DTYPE, device=dev)
            acts_emb = chain_forward(x, spec.embedding, p_emb, out=Slice(both, 0, c_emb))
            reduced = lib.spp_reduce(both, b, s, c_emb, 1.0 / s)
            prop, uctx = unet_forward(spec.unet, Slice(reduced, 0, c_emb), list(params[ne:ne + nu]), need, p_unet)
            assert prop.c == c_prop
            lib.spp_broadcast(prop.t, b, s, c_prop, 1.0, x_coff=prop.coff, out=both, out_coff=c_emb)
            acts_fin = chain_forward(Slice(both, 0, c_emb + c_prop), spec.final, p_fin, last_fp32=True)
            y = acts_fin[-1].t  # (B*S,H,W,outc_p) fp32
            out = y[..., :outc].permute(0, 3, 1, 2).reshape(b, s, outc, h, w).contiguous()
        ctx.spec, ctx.dims = spec, (b, s, h, w, c_emb, c_prop, outc)
        if need:
            ctx.saved = (acts_emb, p_emb, uctx, acts_fin, p_fin)
            ctx.save_for_backward(out)
        return out

    @staticmethod
    def backward(ctx, g):
        spec = ctx.spec
        b, s, h, w, c_emb, c_prop, outc = ctx.dims
        acts_emb, p_emb, uctx, acts_fin, p_fin = ctx.saved
        (out,) = ctx.saved_tensors
        last = spec.final[-1]
        if isinstance(acts_fin, tuple) and acts_fin[0] == "fused":
            _, prop_t, hfin = acts_fin
            _, x16, h1, h2, emb = acts_emb
            g = g.reshape(b, s, outc, h, w).float().contiguous()
            gs, inv_s = grad_scale(g)
            d_emb, d_prop, g_fin = lib.pathnet_final_bwd(g, out, gs, inv_s, emb, prop_t, hfin, p_fin,
                                                         [l.act for l in spec.final], LEAKY_SLOPE, outc, b, s)
            d_red, g_unet = unet_backward(spec.unet, uctx, Slice(d_prop, 0, c_prop), True, inv_s)
            dr = d_red.t if (d_red.coff == 0 and d_red.t.shape[-1] == c_emb) else \
                d_red.t[..., d_red.coff:d_red.coff + c_emb].contiguous()
            g_emb = lib.pathnet_embed_bwd(d_emb, dr, inv_s, emb, h2, h1, x16, p_emb, [l.act for l in spec.embedding],
                                          LEAKY_SLOPE, spec.embedding[0].cin, b, s)
            lib.wgrad_flush()
            ctx.saved = None
            return (None, None) + tuple(g_emb) + tuple(g_unet) + tuple(g_fin)
        g = g.reshape(b * s, outc, h, w).float()
        o = out.reshape(b * s, outc, h, w)
        if last.act == 1:
            g = g * (o > 0)
        elif last.act == 2:
            g = torch.where(o > 0, g, g * LEAKY_SLOPE)
        gs, inv_s = grad_scale(g)
        dz = _grad_nhwc(g, last.cout_p, gs)
        d_both, g_fin = chain_backward(dz, acts_fin, spec.final, p_fin, True, inv_scale=inv_s)
        d_prop = lib.spp_reduce(d_both.t, b, s, c_prop, 1.0, x_coff=d_both.coff + c_emb)
        d_red, g_unet = unet_backward(spec.unet, uctx, Slice(d_prop, 0, c_prop), True, inv_s)
        dz_emb = lib.spp_broadcast(d_red.t, b, s, c_emb, 1.0 / s, x_coff=d_red.coff, add=d_both.t,
                                   add_coff=d_both.coff)
        dze = _act_bwd_full(Slice(dz_emb, 0, c_emb), acts_emb[-1], spec.embedding[-1])
        _, g_emb = chain_backward(dze, acts_emb, spec.embedding, p_emb, False, inv_scale=inv_s)
        lib.wgrad_flush()
        ctx.saved = None
        return (None, None) + tuple(g_emb) + tuple(g_unet) + tuple(g_fin)


# ------------------------------------------------------------------------------------------------
# K9 / K12: step glue (support/interfaces.py:165-180, :206-251; sbmc.KPCN.forward's recombination)
# ------------------------------------------------------------------------------------------------
class PBufferConcatFn(torch.autograd.Function):
    """cat[kpcn_in, mean_S(p[:, :, c0:c0+cr]), var_S(.).mean(C, keepdim) / S] in one launch; the variance channel is
    detached as in the reference, the mean carries the gradient back to the path-embedding network."""

    @staticmethod
    def forward(ctx, kpcn_in, p, c0, cr):
        kin = kpcn_in.detach().contiguous().float()
        pc = p.detach().contiguous().float()
        ctx.meta = (tuple(pc.shape), c0, cr, kin.shape[1])
        return lib.pbuffer_concat_fwd(kin, pc, c0, cr)

    @staticmethod
    def backward(ctx, g):
        shape, c0, cr, cin = ctx.meta
        g = g.contiguous().float()
        dp = lib.pbuffer_concat_bwd(g, shape, c0, cr, cin) if ctx.needs_input_grad[1] else None
        return (g[:, :cin] if ctx.needs_input_grad[0] else None), dp, None, None


class RecombineFn(torch.autograd.Function):
    """radiance = crop_like(albedo, r_d) * r_d + exp(r_s) - 1 (sbmc.KPCN.forward) in one launch."""

    @staticmethod
    def forward(ctx, albedo, r_d, r_s):
        rd, rs = r_d.detach().contiguous().float(), r_s.detach().contiguous().float()
        alb = albedo.detach().float()
        if alb.stride(3) != 1:
            alb = alb.contiguous()
        ctx.save_for_backward(alb, rs)
        return lib.recombine(alb, rd, rs)

    @staticmethod
    def backward(ctx, g):
        alb, rs = ctx.saved_tensors
        h, w = rs.shape[-2:]
        y0, x0 = max((alb.shape[-2] - h) // 2, 0), max((alb.shape[-1] - w) // 2, 0)
        a = alb[..., y0:y0 + h, x0:x0 + w]
        return None, g * a, g * torch.exp(rs)


class ImageLossesFn(torch.autograd.Function):
    """(L1(r_d, t_d), L1(r_s, t_s), L1(radiance, t_t), RelativeMSE(radiance, t_t)) in one reduction launch; targets are
    the full-size tensors (centred crop inside the kernel).  Gradients flow to r_d and r_s only: the step takes the
    recombined image's losses under no_grad (support/interfaces.py:240-249)."""

    @staticmethod
    def forward(ctx, r_d, r_s, rad, t_d, t_s, t_t, eps):
        need = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]
        fix = lambda t: None if t is None else t.detach().contiguous().float()   # noqa: E731
        sums, sgn_d, sgn_s = lib.image_losses(fix(r_d), t_d, fix(r_s), t_s, fix(rad), t_t, eps, need)
        ctx.sgn = (sgn_d, sgn_s)
        return sums[0], sums[1], sums[2], sums[3]

    @staticmethod
    def backward(ctx, g_d, g_s, g_t, g_r):
        sgn_d, sgn_s = ctx.sgn
        ctx.sgn = None
        gd = sgn_d * g_d if (sgn_d is not None and g_d is not None) else None
        gs = sgn_s * g_s if (sgn_s is not None and g_s is not None) else None
        return gd, gs, None, None, None, None, None
