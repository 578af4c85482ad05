"""ctypes binding of libwcmc.so (the C ABI declared in include/wcmc.h).

The product path fails loudly when the CUDA library is missing or the device is not sm_100:
there is no CPU fallback and nothing here ever imports ``oracle/``.
"""
import ctypes
import os
import threading

import torch

from . import streams as _streams

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libwcmc.so")

_lib = None
_lock = threading.Lock()
_inited_devices = set()

c_int, c_float, c_void_p, c_size_t = ctypes.c_int, ctypes.c_float, ctypes.c_void_p, ctypes.c_size_t
c_long = ctypes.c_long

class PackDesc(ctypes.Structure):
    """wcmc_pack_desc of include/wcmc.h"""
    _fields_ = [("w", c_void_p), ("bias", c_void_p), ("dst_fwd", c_void_p), ("dst_dgrad", c_void_p),
                ("dst_bias", c_void_p), ("cout", c_int), ("cin", c_int), ("ksize", c_int), ("cout_p", c_int),
                ("cin_p", c_int)]


class WnDesc(ctypes.Structure):
    """wcmc_wn_desc of include/wcmc.h"""
    _fields_ = [("v", c_void_p), ("g", c_void_p), ("w", c_void_p), ("norm", c_void_p), ("dw", c_void_p),
                ("dv", c_void_p), ("dg", c_void_p), ("rows", c_int), ("cols", c_int)]


class WgradReduceDesc(ctypes.Structure):
    """wcmc_wgrad_reduce_desc of include/wcmc.h"""
    _fields_ = [("ws", c_void_p), ("dw", c_void_p), ("scale", c_void_p), ("nsplit", c_int), ("cout", c_int),
                ("cin", c_int), ("taps", c_int), ("cout_p", c_int), ("cin_p", c_int), ("accumulate", c_int),
                ("nsplit_b", c_int), ("taps_a", c_int)]


class WgradLayer(ctypes.Structure):
    """wcmc_wgrad_layer of include/wcmc.h"""
    _fields_ = [("x", c_void_p), ("dy", c_void_p), ("dw", c_void_p), ("scale", c_void_p), ("N", c_int), ("H", c_int),
                ("W", c_int), ("x_cs", c_int), ("x_coff", c_int), ("cin_p", c_int), ("cin", c_int), ("dy_cs", c_int),
                ("dy_coff", c_int), ("cout_p", c_int), ("cout", c_int), ("ksize", c_int), ("pad", c_int),
                ("accumulate", c_int)]


class AdamTensor(ctypes.Structure):
    """wcmc_adam_tensor of include/wcmc.h"""
    _fields_ = [("p", c_void_p), ("g", c_void_p), ("m", c_void_p), ("v", c_void_p), ("n", ctypes.c_long),
                ("lr", c_float), ("beta1", c_float), ("beta2", c_float), ("eps", c_float)]


# name -> (restype, argtypes); must list every symbol of include/wcmc.h (tests check this).
SIGNATURES = {
    "wcmc_last_error": (ctypes.c_char_p, []),
    "wcmc_version": (ctypes.c_char_p, []),
    "wcmc_init": (c_int, [c_int]),
    "wcmc_tuning_set": (c_int, [ctypes.c_char_p, c_int]),
    "wcmc_nchw_f32_to_nhwc": (c_int, [c_void_p, c_void_p] + [c_int] * 8 + [c_void_p, c_void_p]),
    "wcmc_nhwc_to_nchw_f32": (c_int, [c_void_p, c_int, c_void_p] + [c_int] * 7 + [c_void_p, c_void_p]),
    "wcmc_pack_weights": (c_int, [c_void_p] * 5 + [c_int] * 6 + [c_void_p]),
    "wcmc_pack_weights_batch": (c_int, [ctypes.POINTER(PackDesc), c_int, c_int, c_void_p]),
    "wcmc_conv2d": (c_int, [c_void_p] + [c_int] * 7 + [c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_void_p]
                    + [c_int] * 4 + [c_void_p, c_int, c_int, c_float, c_void_p, c_void_p, c_int, c_void_p]),
    "wcmc_conv2d_kernel_apply": (c_int, [c_void_p] + [c_int] * 7 + [c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_void_p,
                                         c_void_p, c_int, c_int, c_void_p]),
    "wcmc_conv2d_wgrad_workspace": (c_size_t, [c_int] * 7),
    "wcmc_conv2d_wgrad": (c_int, [c_void_p] + [c_int] * 7 + [c_void_p] + [c_int] * 6 + [c_void_p]
                          + [c_int] * 3 + [c_void_p, c_void_p, c_size_t, c_void_p]),
    "wcmc_conv2d_wgrad_partial": (c_int, [c_void_p] + [c_int] * 7 + [c_void_p] + [c_int] * 6 + [c_void_p]
                                  + [c_int] * 3 + [c_void_p, c_void_p, c_size_t, ctypes.POINTER(WgradReduceDesc),
                                                   c_void_p]),
    "wcmc_wgrad_reduce_batch": (c_int, [ctypes.POINTER(WgradReduceDesc), c_int, c_void_p]),
    "wcmc_conv2d_wgrad_group_workspace": (c_size_t, [ctypes.POINTER(WgradLayer), c_int]),
    "wcmc_conv2d_wgrad_group": (c_int, [ctypes.POINTER(WgradLayer), c_int, c_int, c_void_p, c_size_t, c_void_p]),
    "wcmc_conv2d_wgrad_group_plan": (c_int, [ctypes.POINTER(WgradLayer), c_int] + [ctypes.POINTER(c_int)] * 4),
    "wcmc_bias_grad": (c_int, [c_void_p] + [c_int] * 5 + [c_void_p, c_int, c_void_p, c_void_p]),
    "wcmc_kernel_apply_fwd": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p] + [c_int] * 5 + [c_void_p]),
    "wcmc_kernel_apply_bwd": (c_int, [c_void_p, c_int] + [c_void_p] * 5 + [c_int] * 7 + [c_void_p, c_void_p]),
    "wcmc_maxpool2_fwd": (c_int, [c_void_p, c_int, c_int] * 2 + [c_int] * 5 + [c_void_p]),
    "wcmc_maxpool2_bwd": (c_int, [c_void_p, c_int, c_int] * 4 + [c_int] * 5 + [c_void_p]),
    "wcmc_upsample2_fwd": (c_int, [c_void_p, c_int, c_int] * 2 + [c_int] * 5 + [c_void_p]),
    "wcmc_upsample2_bwd": (c_int, [c_void_p, c_int, c_int] * 2 + [c_int] * 5 + [c_void_p]),
    "wcmc_spp_reduce": (c_int, [c_void_p, c_int, c_int] * 2 + [c_int] * 4 + [c_float, c_int, c_void_p]),
    "wcmc_spp_broadcast": (c_int, [c_void_p, c_int, c_int] * 3 + [c_int] * 4 + [c_float, c_int, c_void_p]),
    "wcmc_fmse_perm_workspace": (c_size_t, [c_int] * 4),
    "wcmc_fmse_perm_fwd": (c_int, [c_void_p] + [c_long] * 4 + [c_void_p] + [c_long] * 3 + [c_void_p, c_void_p]
                           + [c_int] * 5 + [c_void_p] * 7 + [c_size_t, c_void_p]),
    "wcmc_fmse_perm_bwd": (c_int, [c_void_p] + [c_long] * 4 + [c_void_p] * 7 + [c_float, c_float] + [c_int] * 5
                           + [c_void_p, c_void_p]),
    "wcmc_pathnet_embed_fwd": (c_int, [c_void_p] + [c_int] * 4 + [c_void_p] * 6 + [c_int] * 5 + [c_float]
                               + [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_int, c_int,
                                  c_void_p]),
    "wcmc_pathnet_final_fwd": (c_int, [c_void_p, c_int, c_int] * 2 + [c_void_p] * 4 + [c_int] * 5 + [c_float]
                               + [c_void_p, c_void_p] + [c_int] * 3 + [c_void_p]),
    "wcmc_pathnet_bwd_workspace": (c_size_t, [c_int]),
    "wcmc_pathnet_final_bwd": (c_int, [c_void_p] * 4 + [c_void_p, c_int] * 2 + [c_void_p] * 3 + [c_int] * 5 + [c_float]
                               + [c_void_p] * 6 + [c_int] * 3 + [c_void_p, c_size_t, c_void_p]),
    "wcmc_pathnet_embed_bwd": (c_int, [c_void_p] * 4 + [c_int] + [c_void_p] * 5 + [c_int] * 5 + [c_float]
                               + [c_void_p] * 6 + [c_int] * 3 + [c_void_p, c_size_t, c_void_p]),
    "wcmc_weight_norm_batch": (c_int, [ctypes.POINTER(WnDesc), c_int, c_int, c_void_p]),
    "wcmc_preprocess_kpcn_workspace": (c_size_t, [c_int, c_int]),
    "wcmc_preprocess_kpcn": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_size_t, c_void_p]),
    "wcmc_preprocess_llpm": (c_int, [c_void_p, ctypes.c_long, c_void_p, c_void_p]),
    "wcmc_fmse_perm_bwd_strided": (c_int, [c_void_p] + [c_long] * 4 + [c_void_p] * 7 + [c_float, c_float] + [c_int] * 5
                                   + [c_void_p] + [c_long] * 4 + [c_void_p]),
    "wcmc_absmax_scale_workspace": (c_size_t, []),
    "wcmc_absmax_scale": (c_int, [c_void_p, c_long, c_float, c_void_p, c_void_p, c_size_t, c_void_p]),
    "wcmc_pbuffer_concat_fwd": (c_int, [c_void_p] * 3 + [c_int] * 7 + [c_void_p]),
    "wcmc_pbuffer_concat_bwd": (c_int, [c_void_p] * 2 + [c_int] * 7 + [c_void_p]),
    "wcmc_recombine": (c_int, [c_void_p, c_long, c_long, c_long, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "wcmc_image_losses_workspace": (c_size_t, []),
    "wcmc_image_losses": (c_int, [c_void_p] * 6 + [ctypes.POINTER(c_long), c_int, c_int, c_int, c_float] + [c_void_p] * 4
                          + [c_size_t, c_void_p]),
    "wcmc_random_permutation": (c_int, [c_void_p, c_long, c_void_p, ctypes.c_uint, c_void_p]),
    "wcmc_grad_exchange_flag_bytes": (c_size_t, []),
    "wcmc_grad_exchange": (c_int, [c_void_p, c_void_p, c_void_p, c_long, c_long, c_long, c_int, c_int, c_int, c_float,
                                   c_void_p]),
    "wcmc_adam_chunk": (c_int, []),
    "wcmc_adam_clip_step": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_float, c_void_p, c_void_p]),
    "wcmc_fmse_allpairs_workspace": (c_size_t, [c_int, c_int]),
    "wcmc_fmse_allpairs_fwd": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_float, c_float, c_void_p, c_void_p,
                                       c_void_p, c_size_t, c_void_p]),
    "wcmc_act_bwd": (c_int, [c_void_p, c_int, c_int] * 3 + [ctypes.c_long, c_int, c_int, c_float, c_int, c_void_p]),
}


class WcmcError(RuntimeError):
    pass


# ---- bookkeeping for bench.py: kernels launched, and (optionally) per-launch device time -------
LAUNCHES = {"count": 0}
_profile = None  # when a list: (name, algorithmic_work, start_event, end_event) per timed call
_KERNELS_PER_CALL = {"conv2d_wgrad": 2, "bias_grad": 2, "fmse_perm_fwd": 2, "adam_clip_step": 2,
                     "fmse_allpairs_fwd": 3, "pathnet_final_bwd": 2, "pathnet_embed_bwd": 2}


def _pending():
    """Per thread AND device: (WgradReduceDesc, keep-alive tensors) of deferred weight-gradient reductions
    (nn.DataParallel drives one replica per thread; a flush must only see its own replica's layers)."""
    d = getattr(_tls, "pending", None)
    if d is None:
        d = _tls.pending = {}
    return d.setdefault(_tls.dev, [])


def profile_start():
    global _profile
    _profile = []


def profile_stop():
    """-> {name: (n_calls, total_ms, total_algorithmic_work)}; synchronises the device."""
    global _profile
    rec, _profile = _profile, None
    torch.cuda.synchronize()
    out = {}
    for name, work, e0, e1 in rec or []:
        n, ms, wk = out.get(name, (0, 0.0, 0.0))
        out[name] = (n + 1, ms + e0.elapsed_time(e1), wk + work)
    return out


def _run(fn, what, work, *args):
    """Calls one C-ABI entry point; counts its kernel launches; optionally brackets it with CUDA
    events on the launching stream (bench.py's live per-kernel timing)."""
    LAUNCHES["count"] += _KERNELS_PER_CALL.get(what, 1)
    dev = _tls.dev
    if torch.cuda.current_device() != dev:
        # the library launches on the CUDA runtime's current device; the tensors (and the stream fetched by
        # _stream()) live on `dev`
        with torch.cuda.device(dev):
            rc = fn(*args)
    elif _profile is None:
        rc = fn(*args)
    else:
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = fn(*args)
        e1.record()
        _profile.append((what, work, e0, e1))
    if rc != 0:
        raise WcmcError("%s failed (%d): %s" % (what, rc, load().wcmc_last_error().decode()))


def load():
    """Loads libwcmc.so (no device needed); raises if it has not been built."""
    global _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise WcmcError("%s not found: run `make` (or __graft_entry__.build()) first; "
                                "there is no CPU fallback" % LIB_PATH)
            lib = ctypes.CDLL(LIB_PATH)
            for name, (res, args) in SIGNATURES.items():
                fn = getattr(lib, name)
                fn.restype = res
                fn.argtypes = args
            _lib = lib
    return _lib


def _check(rc, what):
    if rc != 0:
        raise WcmcError("%s failed (%d): %s" % (what, rc, load().wcmc_last_error().decode()))


_tls = threading.local()   # .dev: device index of the call being issued; .pending: deferred weight-gradient reductions


def _resolve(device):
    if isinstance(device, int):
        return device
    if device is not None:
        idx = torch.device(device).index
        if idx is not None:
            return idx
    return torch.cuda.current_device()


def init(device=None):
    """Loads the library, checks `device` once (sm_100 or WcmcError) and makes it the device of this thread's next
    launches.  `device`: None (torch's current device), int, or torch.device -- the wrappers below pass the device
    of their tensors, so a model living on cuda:1 launches on cuda:1 whatever torch's current device is (the
    reference never calls set_device: /root/reference/train_kpcn.py:259-269 `.cuda(args.device_id)`, nn.DataParallel
    replicas in threads).  Afterwards a dictionary lookup: this sits on the launch path of every kernel."""
    lib = _lib
    if lib is None:
        lib = load()
    if not _inited_devices and not torch.cuda.is_available():
        raise WcmcError("wcmc_b200 needs a CUDA device (sm_100a); none is visible and there is no CPU fallback")
    dev = _resolve(device)
    if dev not in _inited_devices:
        with _lock:
            if dev not in _inited_devices:
                _check(lib.wcmc_init(dev), "wcmc_init")
                if not _inited_devices:
                    # measurement aid: WCMC_TUNE="knob=value,..." forwards to wcmc_tuning_set (tools/, bench A/B runs)
                    for item in filter(None, os.environ.get("WCMC_TUNE", "").split(",")):
                        name, _, val = item.partition("=")
                        _check(lib.wcmc_tuning_set(name.strip().encode(), int(val)), "wcmc_tuning_set")
                _inited_devices.add(dev)
    _tls.dev = dev
    return lib


def _stream():
    # raw cudaStream_t of torch's current stream ON THE DEVICE OF THE CALL (torch.cuda.current_stream() costs ~15 us)
    return torch._C._cuda_getCurrentRawStream(_tls.dev)


def _p(t):
    return 0 if t is None else t.data_ptr()


def pad16(c):
    return (c + 15) // 16 * 16


BF16, F16, F32 = 0, 1, 2
_DT = {torch.bfloat16: BF16, torch.float16: F16, torch.float32: F32}


def _dt(t):
    return _DT[t.dtype]


def _h16(t):
    assert t.dtype in (torch.bfloat16, torch.float16) and t.is_contiguous(), "expected contiguous 16-bit NHWC tensor"
    return t


# ------------------------------------------------------------------------------------------------
# thin typed wrappers (torch tensors in, torch tensors out); all run on the current stream.
# 16-bit activations may be torch.float16 or torch.bfloat16 (WCMC_F16 / WCMC_BF16); gradients are bf16.
# ------------------------------------------------------------------------------------------------
def nchw_to_nhwc(src, dst=None, dst_coff=0, c_fill=None, dtype=torch.bfloat16, scale=None):
    """src (N,C,H,W) fp32 -> dst (N,H,W,Cs) 16-bit, channels [dst_coff, dst_coff+c_fill)."""
    lib = init(src.device)
    assert src.dtype == torch.float32 and src.is_contiguous()
    n, c, h, w = src.shape
    if c_fill is None:
        c_fill = pad16(c)
    if dst is None:
        dst = torch.empty((n, h, w, c_fill), dtype=dtype, device=src.device)
    assert _h16(dst).shape[:3] == (n, h, w)
    _run(lib.wcmc_nchw_f32_to_nhwc, "nchw_f32_to_nhwc", src.numel() * 4.0 + n * h * w * c_fill * 2.0,
         src.data_ptr(), dst.data_ptr(), _dt(dst), n, c, h, w, dst.shape[3], dst_coff, c_fill, _p(scale),
         _stream())
    return dst


def nhwc_to_nchw(src, c, src_coff=0, out=None, accumulate=False, scale=None):
    lib = init(src.device)
    n, h, w, cs = _h16(src).shape
    if out is None:
        out = torch.empty((n, c, h, w), dtype=torch.float32, device=src.device)
        accumulate = False
    assert out.dtype == torch.float32 and out.is_contiguous() and tuple(out.shape) == (n, c, h, w)
    _run(lib.wcmc_nhwc_to_nchw_f32, "nhwc_to_nchw_f32", out.numel() * 6.0,
         src.data_ptr(), _dt(src), out.data_ptr(), n, c, h, w, cs, src_coff, int(accumulate), _p(scale),
         _stream())
    return out


def pack_weights(w, bias=None, cout_p=None, cin_p=None, fwd=True, dgrad=True, want_bias=False,
                 dtype=torch.bfloat16):
    """torch (Cout,Cin,k,k) fp32 -> (fwd [cout_p,k*k,cin_p], dgrad [cin_p,k*k,cout_p]) in `dtype`
    [, zero-padded fp32 bias [cout_p] when ``want_bias``]."""
    lib = init(w.device)
    w = w.detach()
    assert w.dtype == torch.float32
    w = w.contiguous()
    cout, cin, k, k2 = w.shape
    assert k == k2
    cout_p = cout_p or pad16(cout)
    cin_p = cin_p or pad16(cin)
    f = torch.empty((cout_p, k * k, cin_p), dtype=dtype, device=w.device) if fwd else None
    d = torch.empty((cin_p, k * k, cout_p), dtype=dtype, device=w.device) if dgrad else None
    bp = torch.empty((cout_p,), dtype=torch.float32, device=w.device) if want_bias else None
    if bias is not None:
        bias = bias.detach().contiguous()
        assert bias.dtype == torch.float32 and bias.numel() == cout
    _run(lib.wcmc_pack_weights, "pack_weights", w.numel() * 8.0,
         w.data_ptr(), _p(bias), _p(f), _p(d), _p(bp), _DT[dtype], cout, cin, k, cout_p, cin_p, _stream())
    if want_bias:
        return f, d, bp
    return f, d


def pack_weights_batch(specs, dtype=torch.bfloat16, dgrad=True):
    """specs: [(w, bias, cout_p, cin_p)] -> [(fwd, dgrad or None, bias_p)], one kernel launch."""
    lib = init(specs[0][0].device)
    n = len(specs)
    descs = (PackDesc * n)()
    out, keep = [], []
    for i, (w, bias, cout_p, cin_p) in enumerate(specs):
        w = w.detach()
        if w.dtype != torch.float32 or not w.is_contiguous():
            w = w.float().contiguous()
        cout, cin, k, _ = w.shape
        f = torch.empty((cout_p, k * k, cin_p), dtype=dtype, device=w.device)
        d = torch.empty((cin_p, k * k, cout_p), dtype=dtype, device=w.device) if dgrad else None
        bp = torch.empty((cout_p,), dtype=torch.float32, device=w.device)
        if bias is not None:
            bias = bias.detach()
            if bias.dtype != torch.float32 or not bias.is_contiguous():
                bias = bias.float().contiguous()
        keep.append((w, bias))
        descs[i] = PackDesc(w.data_ptr(), _p(bias), f.data_ptr(), _p(d), bp.data_ptr(), cout, cin, k, cout_p, cin_p)
        out.append((f, d, bp))
    _run(lib.wcmc_pack_weights_batch, "pack_weights", sum(s[0].numel() for s in specs) * 8.0, descs, n,
         _DT[dtype], _stream())
    return out


SHARE_STREAMS = os.environ.get("WCMC_CONV_SHARE", "1") != "0"


def conv2d(x, w_packed, bias, ksize, pad, act=0, out=None, out_coff=0, out_dtype=None, x_coff=0, mask=None,
           mask_coff=0, slope=0.0, flags=0, cin=None, cout=None, colsum=None, colsum_scale=None, alg_hw=None):
    """x (N,H,W,Cs) 16-bit NHWC; w_packed (cout_p, k*k, cin_p) 16-bit; returns the NHWC output
    (dtype `out_dtype`, default = x's dtype; torch.float32 for the logits layer).
    alg_hw: spatial extent the ALGORITHMIC work of this launch is counted on (bench accounting only): a data-gradient
    launch is a "full"-padded convolution over the forward layer's output map, and its algorithmic FLOPs are the
    forward layer's 2*N*Ho*Wo*k^2*Cin*Cout, not this launch's larger output extent."""
    lib = init(x.device)
    n, h, w, xcs = _h16(x).shape
    cout_p, taps, cin_p = _h16(w_packed).shape
    assert taps == ksize * ksize
    ho, wo = h + 2 * pad - ksize + 1, w + 2 * pad - ksize + 1
    if out is None:
        out = torch.empty((n, ho, wo, cout_p), dtype=out_dtype or x.dtype, device=x.device)
    assert out.is_contiguous() and tuple(out.shape[:3]) == (n, ho, wo) and out.dtype in _DT
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.numel() >= cout_p
    if mask is not None:
        assert tuple(_h16(mask).shape[:3]) == (n, ho, wo)
    ah, aw = alg_hw or (ho, wo)
    if SHARE_STREAMS and not (flags >> 24) & 15:
        flags |= _streams.share() << 24      # two halves side by side on two streams: plan for half of the SMs each
    _run(lib.wcmc_conv2d, "conv2d_k%d" % ksize, 2.0 * n * ah * aw * ksize * ksize * (cin or cin_p) * (cout or cout_p),
         x.data_ptr(), _dt(x), n, h, w, xcs, x_coff, cin_p, w_packed.data_ptr(), _dt(w_packed), cout_p, _p(bias),
         ksize, pad, out.data_ptr(), _dt(out), out.shape[3], out_coff, act, _p(mask),
         0 if mask is None else mask.shape[3], mask_coff, float(slope), _p(colsum), _p(colsum_scale), flags,
         _stream())
    return out


def conv2d_kernel_apply(x, w_packed, bias, data, ksize, pad, ka_ksize=21, x_coff=0, flags=0, cin=None, cout=None):
    """Last KPCN layer fused with softmax + kernel-apply (inference): x (N,H,W,Cs) 16-bit NHWC, w_packed
    (448, k*k, cin_p), bias fp32 (448), data (N,3,Ho,Wo) fp32 -> out (N,3,Ho,Wo) fp32.  The logits stay on chip."""
    lib = init(x.device)
    n, h, w, xcs = _h16(x).shape
    cout_p, taps, cin_p = _h16(w_packed).shape
    assert taps == ksize * ksize and bias.dtype == torch.float32 and bias.numel() >= cout_p
    ho, wo = h + 2 * pad - ksize + 1, w + 2 * pad - ksize + 1
    assert data.dtype == torch.float32 and data.is_contiguous() and tuple(data.shape) == (n, 3, ho, wo)
    out = torch.empty_like(data)
    _run(lib.wcmc_conv2d_kernel_apply, "conv2d_k%d" % ksize,
         2.0 * n * ho * wo * ksize * ksize * (cin or cin_p) * (cout or cout_p),
         x.data_ptr(), _dt(x), n, h, w, xcs, x_coff, cin_p, w_packed.data_ptr(), _dt(w_packed), cout_p, bias.data_ptr(),
         ksize, pad, data.data_ptr(), out.data_ptr(), ka_ksize, flags, _stream())
    return out


_ws_cache = {}


def _workspace(nbytes, device):
    key = (device.index, torch.cuda.current_stream().cuda_stream)
    ws = _ws_cache.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, device=device)
        _ws_cache[key] = ws
    return ws


def conv2d_wgrad(x, dy, cout, cin, ksize, pad, cin_p, cout_p, x_coff=0, dy_coff=0, out=None, accumulate=False,
                 scale=None, defer=False):
    """dw (cout,cin,k,k) fp32 (+)= sum dy (x) x ; x (N,H,W,Cs) / dy (N,Ho,Wo,Cs') 16-bit NHWC.
    defer=True: only the tensor-core kernel runs now; the split-K reduction of every deferred layer
    happens in ONE launch at the next wgrad_flush() (dw is not valid before that)."""
    lib = init(x.device)
    n, h, w, xcs = _h16(x).shape
    ho, wo = h + 2 * pad - ksize + 1, w + 2 * pad - ksize + 1
    assert tuple(_h16(dy).shape[:3]) == (n, ho, wo), (dy.shape, (n, ho, wo))
    if out is None:
        out = torch.empty((cout, cin, ksize, ksize), dtype=torch.float32, device=x.device)
        accumulate = False
    assert out.is_contiguous() and out.dtype == torch.float32
    work = 2.0 * n * ho * wo * ksize * ksize * cin * cout
    if defer:
        # nothing is launched now: the layer joins the backward pass's group and wgrad_flush() runs ONE grouped
        # launch + one reduction for all of them (x, dy, out and scale are kept alive until then)
        layer = WgradLayer(x.data_ptr(), dy.data_ptr(), out.data_ptr(), _p(scale), n, h, w, xcs, x_coff, cin_p, cin,
                           dy.shape[3], dy_coff, cout_p, cout, ksize, pad, int(accumulate))
        _pending().append((layer, (x, dy, out, scale), work, _dt(x)))
        return out
    need = lib.wcmc_conv2d_wgrad_workspace(n, h, w, cin_p, cout_p, ksize, pad)
    ws = _workspace(need, x.device)
    _run(lib.wcmc_conv2d_wgrad, "conv2d_wgrad", work,
         x.data_ptr(), _dt(x), n, h, w, xcs, x_coff, cin_p, dy.data_ptr(), _dt(dy), dy.shape[3], dy_coff, cout_p,
         ksize, pad, out.data_ptr(), cout, cin, int(accumulate), _p(scale), ws.data_ptr(), ws.numel(), _stream())
    return out


WGRAD_GROUP_MAX = 64


def wgrad_flush(device=None):
    """Runs every deferred weight gradient of this thread on `device` (default: the device of the last call): one
    grouped tensor-core launch per <= 16 layers (wcmc_conv2d_wgrad_group) + one batched reduction."""
    lib = init(device if device is not None else getattr(_tls, "dev", None))
    pend = _pending()
    if not pend:
        return
    try:
        for base in range(0, len(pend), WGRAD_GROUP_MAX):
            part = pend[base:base + WGRAD_GROUP_MAX]
            n = len(part)
            layers = (WgradLayer * n)(*[p[0] for p in part])
            dtype = part[0][3]
            assert all(p[3] == dtype for p in part), "deferred weight gradients must share one 16-bit format"
            need = lib.wcmc_conv2d_wgrad_group_workspace(layers, n)
            if need == 0:
                raise WcmcError("wgrad_group: %s" % lib.wcmc_last_error().decode())
            dev = part[0][1][0].device
            ws = torch.empty(need + 256, dtype=torch.uint8, device=dev)
            off = (-ws.data_ptr()) % 256
            # bench accounting: the call is booked per filter size (the 5x5 KPCN stacks and the 3x3 U-Net differ by
            # a factor of two in achievable rate); a mixed group goes under the size with the most work
            by_k = {}
            for p in part:
                by_k[p[0].ksize] = by_k.get(p[0].ksize, 0.0) + p[2]
            k_main = max(by_k, key=by_k.get)
            LAUNCHES["count"] += 1     # grouped kernel + reduction
            _run(lib.wcmc_conv2d_wgrad_group, "conv2d_wgrad_k%d" % k_main, sum(by_k.values()), layers, n, dtype,
                 ws.data_ptr() + off, need, _stream())
    finally:
        pend.clear()


def wgrad_group_plan(specs):
    """Host-only view of the grouped launch plan.  specs: [(N, H, W, cin, cout, ksize, pad)] ->
    [(teams, ctas, taps_per_group, column_stride)] per layer, number of launches."""
    lib = load()
    n = len(specs)
    layers = (WgradLayer * n)()
    for i, (nb, h, w, cin, cout, k, pad) in enumerate(specs):
        layers[i] = WgradLayer(1, 1, 1, 0, nb, h, w, pad16(cin), 0, pad16(cin), cin, pad16(cout), 0, pad16(cout), cout, k,
                               pad, 0)
    outs = [(c_int * n)() for _ in range(4)]
    rc = lib.wcmc_conv2d_wgrad_group_plan(layers, n, *outs)
    if rc < 0:
        raise WcmcError("wgrad_group_plan failed (%d): %s" % (rc, lib.wcmc_last_error().decode()))
    return [tuple(o[i] for o in outs) for i in range(n)], rc


def bias_grad(dy, cout, dy_coff=0, out=None, accumulate=False, scale=None):
    lib = init(dy.device)
    _h16(dy)
    npix = dy.shape[0] * dy.shape[1] * dy.shape[2]
    if out is None:
        out = torch.empty((cout,), dtype=torch.float32, device=dy.device)
        accumulate = False
    if accumulate:
        LAUNCHES["count"] -= 1   # no zero kernel
    _run(lib.wcmc_bias_grad, "bias_grad", npix * cout * 2.0,
         dy.data_ptr(), _dt(dy), npix, dy.shape[3], dy_coff, cout, out.data_ptr(), int(accumulate), _p(scale),
         _stream())
    return out


def kernel_apply_fwd(logits_nhwc, data, ksize, want_stats=True):
    """logits (N,H,W,Cs>=k*k) fp32 NHWC, data (N,C,H,W) fp32 -> out (N,C,H,W), stats (N,H,W,2)."""
    lib = init(data.device)
    assert logits_nhwc.dtype == torch.float32 and logits_nhwc.is_contiguous()
    assert data.dtype == torch.float32 and data.is_contiguous()
    n, c, h, w = data.shape
    assert tuple(logits_nhwc.shape[:3]) == (n, h, w)
    out = torch.empty_like(data)
    stats = torch.empty((n, h, w, 2), dtype=torch.float32, device=data.device) if want_stats else None
    _run(lib.wcmc_kernel_apply_fwd, "kernel_apply_fwd", n * h * w * (ksize * ksize * 4.0 + c * 8.0),
         logits_nhwc.data_ptr(), logits_nhwc.shape[3], data.data_ptr(), out.data_ptr(), _p(stats), n, c, h, w,
         ksize, _stream())
    return out, stats


def kernel_apply_bwd(logits_nhwc, data, out, stats, grad_out, ksize, dl_cs=None, dtype=torch.bfloat16, scale=None):
    lib = init(data.device)
    n, c, h, w = data.shape
    grad_out = grad_out.contiguous()
    assert grad_out.dtype == torch.float32
    dl_cs = dl_cs or logits_nhwc.shape[3]
    dl = torch.empty((n, h, w, dl_cs), dtype=dtype, device=data.device)
    bf16 = dtype != torch.float32
    _run(lib.wcmc_kernel_apply_bwd, "kernel_apply_bwd",
         n * h * w * (ksize * ksize * (4.0 + (2.0 if bf16 else 4.0)) + c * 12.0 + 8.0),
         logits_nhwc.data_ptr(), logits_nhwc.shape[3], data.data_ptr(), out.data_ptr(), stats.data_ptr(),
         grad_out.data_ptr(), dl.data_ptr(), dl_cs, _dt(dl), n, c, h, w, ksize, _p(scale), _stream())
    return dl


# ---- NHWC glue (PathNet / U-Net); tensors are (.., Cs) 16-bit contiguous, slices by (coff, C) --------
def maxpool2_fwd(x, c, x_coff=0, out=None, out_coff=0):
    lib = init(x.device)
    n, h, w, xcs = _h16(x).shape
    if out is None:
        out = torch.empty((n, h // 2, w // 2, c), dtype=x.dtype, device=x.device)
    assert out.dtype == x.dtype
    _run(lib.wcmc_maxpool2_fwd, "maxpool2_fwd", 0.0, x.data_ptr(), xcs, x_coff, _h16(out).data_ptr(),
         out.shape[3], out_coff, n, h, w, c, _dt(x), _stream())
    return out


def maxpool2_bwd(x, dy, c, x_coff=0, dy_coff=0, add=None, add_coff=0, out=None, out_coff=0):
    """x: forward input; dy / add / out: gradients (all one 16-bit dtype)."""
    lib = init(x.device)
    n, h, w, xcs = _h16(x).shape
    if out is None:
        out = torch.empty((n, h, w, c), dtype=x.dtype, device=x.device)
    assert dy.dtype == x.dtype and out.dtype == x.dtype and (add is None or add.dtype == x.dtype)
    _run(lib.wcmc_maxpool2_bwd, "maxpool2_bwd", 0.0, x.data_ptr(), xcs, x_coff, _h16(dy).data_ptr(), dy.shape[3],
         dy_coff, _p(add), 0 if add is None else add.shape[3], add_coff, _h16(out).data_ptr(), out.shape[3],
         out_coff, n, h, w, c, _dt(x), _stream())
    return out


def upsample2_fwd(x, c, x_coff=0, out=None, out_coff=0):
    lib = init(x.device)
    n, h, w, xcs = _h16(x).shape
    if out is None:
        out = torch.empty((n, 2 * h, 2 * w, c), dtype=x.dtype, device=x.device)
    assert out.dtype == x.dtype
    _run(lib.wcmc_upsample2_fwd, "upsample2_fwd", 0.0, x.data_ptr(), xcs, x_coff, _h16(out).data_ptr(),
         out.shape[3], out_coff, n, h, w, c, _dt(x), _stream())
    return out


def upsample2_bwd(dy, c, dy_coff=0, out=None, out_coff=0):
    lib = init(dy.device)
    n, hh, ww, dcs = _h16(dy).shape
    h, w = hh // 2, ww // 2
    if out is None:
        out = torch.empty((n, h, w, c), dtype=dy.dtype, device=dy.device)
    assert out.dtype == dy.dtype
    _run(lib.wcmc_upsample2_bwd, "upsample2_bwd", 0.0, dy.data_ptr(), dcs, dy_coff, _h16(out).data_ptr(),
         out.shape[3], out_coff, n, h, w, c, _dt(dy), _stream())
    return out


def spp_reduce(x, b, s, c, scale, x_coff=0, out=None, out_coff=0):
    """x (B*S,H,W,Cs) -> out (B,H,W,.) = scale * sum over S (same 16-bit dtype)."""
    lib = init(x.device)
    bs, h, w, xcs = _h16(x).shape
    assert bs == b * s
    if out is None:
        out = torch.empty((b, h, w, c), dtype=x.dtype, device=x.device)
    assert out.dtype == x.dtype
    _run(lib.wcmc_spp_reduce, "spp_reduce", 0.0, x.data_ptr(), xcs, x_coff, _h16(out).data_ptr(), out.shape[3],
         out_coff, b, s, h * w, c, float(scale), _dt(x), _stream())
    return out


def spp_broadcast(x, b, s, c, scale=1.0, x_coff=0, add=None, add_coff=0, out=None, out_coff=0):
    """out (B*S,H,W,.) slice = (add or 0) + scale * x (B,H,W,.) broadcast over S (same 16-bit dtype)."""
    lib = init(x.device)
    bb, h, w, xcs = _h16(x).shape
    assert bb == b
    if out is None:
        out = torch.empty((b * s, h, w, c), dtype=x.dtype, device=x.device)
    assert out.dtype == x.dtype and (add is None or add.dtype == x.dtype)
    _run(lib.wcmc_spp_broadcast, "spp_broadcast", 0.0, x.data_ptr(), xcs, x_coff, _p(add),
         0 if add is None else add.shape[3], add_coff, _h16(out).data_ptr(), out.shape[3], out_coff, b, s, h * w, c,
         float(scale), _dt(x), _stream())
    return out


def act_bwd(dy, y, c, act, slope=0.01, dy_coff=0, y_coff=0, out=None, out_coff=0):
    """dz = dy * act'(y): dy / dz gradients, y the activation output (all one 16-bit dtype)."""
    lib = init(dy.device)
    assert dy.dtype == y.dtype
    npix = dy.shape[0] * dy.shape[1] * dy.shape[2]
    if out is None:
        out = torch.empty(tuple(dy.shape[:3]) + (c,), dtype=dy.dtype, device=dy.device)
    _run(lib.wcmc_act_bwd, "act_bwd", 0.0, _h16(dy).data_ptr(), dy.shape[3], dy_coff, _h16(y).data_ptr(),
         y.shape[3], y_coff, _h16(out).data_ptr(), out.shape[3], out_coff, npix, c, act, float(slope), _dt(dy),
         _stream())
    return out


# ---- K10: path-disentangling loss (permutation-paired) ------------------------------------------
def _pview(p):
    """(B,S,C,H,W) fp32 view with unit x stride -> (ptr, sb, ss, sc, sh)."""
    assert p.dtype == torch.float32 and p.dim() == 5 and p.stride(4) == 1, "p-buffer view must be fp32 with unit W stride"
    return (p.data_ptr(),) + tuple(p.stride()[:4])


def fmse_perm_fwd(p, ref, idx_patch, idx_batch=None):
    """p (B,S,C,H,W) / ref (B,3,H,W): fp32 (possibly cropped, strided) views; idx_*: int64 device
    permutations.  -> dict(loss (2,), e_patch, e_batch, inv_patch, inv_batch, nonfinite (int32 flag))."""
    lib = init(p.device)
    b, s, c, h, w = p.shape
    assert ref.dtype == torch.float32 and tuple(ref.shape) == (b, 3, h, w) and ref.stride(3) == 1
    n = s * h * w
    dev = p.device
    assert idx_patch.dtype == torch.int64 and idx_patch.numel() == n and idx_patch.is_contiguous()
    assert idx_patch.device == dev
    if idx_batch is not None:
        assert idx_batch.dtype == torch.int64 and idx_batch.numel() == b * n and idx_batch.is_contiguous()
        assert idx_batch.device == dev
    r = dict(loss=torch.empty(2, dtype=torch.float32, device=dev),
             e_patch=torch.empty(b * n, dtype=torch.float32, device=dev),
             e_batch=torch.empty(b * n, dtype=torch.float32, device=dev) if idx_batch is not None else None,
             inv_patch=torch.empty(n, dtype=torch.int32, device=dev),
             inv_batch=torch.empty(b * n, dtype=torch.int32, device=dev) if idx_batch is not None else None,
             nonfinite=torch.zeros(1, dtype=torch.int32, device=dev))
    need = lib.wcmc_fmse_perm_workspace(b, s, h, w)
    ws = _workspace(need, dev)
    modes = 2 if idx_batch is not None else 1
    work = b * n * ((c + 3) * 4.0 * (1 + modes) + modes * 16.0)   # row + partner rows, idx + e + inv per mode
    _run(lib.wcmc_fmse_perm_fwd, "fmse_perm_fwd", work, *_pview(p), ref.data_ptr(), ref.stride(0), ref.stride(1),
         ref.stride(2), idx_patch.data_ptr(), _p(idx_batch), b, s, c, h, w, r["e_patch"].data_ptr(),
         _p(r["e_batch"]), r["inv_patch"].data_ptr(), _p(r["inv_batch"]), r["loss"].data_ptr(),
         r["nonfinite"].data_ptr(), ws.data_ptr(), ws.numel(), _stream())
    return r


def fmse_perm_bwd(p, idx_patch, idx_batch, inv_patch, inv_batch, w_patch, w_batch, scale, coef_patch, coef_batch,
                  out=None):
    """-> dp (B,S,C,H,W) contiguous fp32 (see include/wcmc.h); or written into `out`, a (B,S,C,H,W) fp32 view with
    unit x stride (the interior of a zero-filled tensor of the uncropped p-buffer's shape)."""
    lib = init(p.device)
    b, s, c, h, w = p.shape
    dp = torch.empty((b, s, c, h, w), dtype=torch.float32, device=p.device) if out is None else out
    assert dp.dtype == torch.float32 and tuple(dp.shape) == (b, s, c, h, w) and dp.stride(4) == 1
    for t in (w_patch, w_batch):
        assert t is None or (t.dtype == torch.float32 and t.is_contiguous() and t.numel() == b * s * h * w)
    modes = 2 if idx_batch is not None else 1
    work = b * s * h * w * (c * 4.0 * (2 + 2 * modes) + modes * 20.0)
    _run(lib.wcmc_fmse_perm_bwd_strided, "fmse_perm_bwd", work, *_pview(p), idx_patch.data_ptr(), _p(idx_batch),
         inv_patch.data_ptr(), _p(inv_batch), w_patch.data_ptr(), _p(w_batch), _p(scale), float(coef_patch),
         float(coef_batch), b, s, c, h, w, dp.data_ptr(), *dp.stride()[:4], _stream())
    return dp


# ---- K9 / K12: step glue ------------------------------------------------------------------------------------
def pbuffer_concat_fwd(kpcn_in, p, c0, cr):
    """kpcn_in (B,Cin,H,W), p (B,S,C,H,W) fp32 contiguous -> (B, Cin+cr+1, H, W) = cat[kpcn_in, mean_S p[:, :, c0:c0+cr],
    var_S(.).mean(C)/S]   (support/interfaces.py:165-180)."""
    lib = init(p.device)
    b, s, c, h, w = p.shape
    cin = kpcn_in.shape[1]
    assert kpcn_in.dtype == torch.float32 and kpcn_in.is_contiguous() and tuple(kpcn_in.shape) == (b, cin, h, w)
    assert p.dtype == torch.float32 and p.is_contiguous()
    out = torch.empty((b, cin + cr + 1, h, w), dtype=torch.float32, device=p.device)
    _run(lib.wcmc_pbuffer_concat_fwd, "pbuffer_concat_fwd", (kpcn_in.numel() + out.numel() + b * s * cr * h * w) * 4.0,
         kpcn_in.data_ptr(), p.data_ptr(), out.data_ptr(), b, s, c, c0, cr, cin, h * w, _stream())
    return out


def pbuffer_concat_bwd(grad_out, shape, c0, cr, cin):
    """grad_out (B, Cin+cr+1, H, W) -> d p (B,S,C,H,W) (zeros outside the channel range)."""
    lib = init(grad_out.device)
    b, s, c, h, w = shape
    grad_out = grad_out.contiguous()
    assert grad_out.dtype == torch.float32 and tuple(grad_out.shape) == (b, cin + cr + 1, h, w)
    dp = torch.empty(shape, dtype=torch.float32, device=grad_out.device)
    _run(lib.wcmc_pbuffer_concat_bwd, "pbuffer_concat_bwd", (dp.numel() + b * cr * h * w) * 4.0, grad_out.data_ptr(),
         dp.data_ptr(), b, s, c, c0, cr, cin, h * w, _stream())
    return dp


def _crop_view(t, h, w):
    """Centred (h, w) crop of a (B,3,H,W) fp32 tensor as (data_ptr at the crop origin, batch / channel / row strides)."""
    assert t.dtype == torch.float32 and t.dim() == 4 and t.shape[1] == 3 and t.stride(3) == 1
    dh, dw = t.shape[2] - h, t.shape[3] - w
    assert dh >= 0 and dw >= 0
    y0, x0 = max(dh // 2, 0), max(dw // 2, 0)
    return t.data_ptr() + 4 * (y0 * t.stride(2) + x0), t.stride(0), t.stride(1), t.stride(2)


def recombine(albedo, r_d, r_s):
    """radiance = crop(albedo) * r_d + exp(r_s) - 1; albedo (B,3,H,W) any size >= r_d's (centred crop)."""
    lib = init(r_d.device)
    b, _, h, w = r_d.shape
    assert r_d.dtype == torch.float32 and r_d.is_contiguous() and r_s.dtype == torch.float32 and r_s.is_contiguous()
    ptr, sb, sc, sh = _crop_view(albedo, h, w)
    out = torch.empty_like(r_d)
    _run(lib.wcmc_recombine, "recombine", r_d.numel() * 16.0, ptr, sb, sc, sh, r_d.data_ptr(), r_s.data_ptr(),
         out.data_ptr(), b, h, w, _stream())
    return out


_loss_ws = {}
_scale_ws = {}


def absmax_scale(g, target):
    """-> fp32 (2,) device tensor [target / max|g|, max|g| / target] in one launch (no host sync)."""
    lib = init(g.device)
    assert g.dtype == torch.float32 and g.is_contiguous()
    if g.data_ptr() % 16:
        g = g.clone()
    key = (g.device.index, torch.cuda.current_stream().cuda_stream)
    ws = _scale_ws.get(key)
    if ws is None:
        ws = _scale_ws[key] = torch.zeros(lib.wcmc_absmax_scale_workspace(), dtype=torch.uint8, device=g.device)
    out = torch.empty(2, dtype=torch.float32, device=g.device)
    _run(lib.wcmc_absmax_scale, "absmax_scale", g.numel() * 4.0, g.data_ptr(), g.numel(), float(target), out.data_ptr(),
         ws.data_ptr(), ws.numel(), _stream())
    return out



def image_losses(r_d, t_d, r_s, t_s, rad, t_t, eps, want_signs):
    """-> (sums (4,) = [L1 diffuse, L1 specular, L1 total, RelMSE total], sgn_d, sgn_s).  Predictions (B,3,h,w) fp32
    contiguous or None; targets full-size (B,3,H,W) tensors read through a centred crop."""
    ref = next(t for t in (r_d, r_s, rad) if t is not None)
    lib = init(ref.device)
    b, _, h, w = ref.shape
    key = (ref.device.index, torch.cuda.current_stream().cuda_stream)
    ws = _loss_ws.get(key)
    if ws is None:
        ws = _loss_ws[key] = torch.zeros(lib.wcmc_image_losses_workspace(), dtype=torch.uint8, device=ref.device)
    ptrs, strides = [], []
    for pred, tgt in ((r_d, t_d), (r_s, t_s), (rad, t_t)):
        if pred is None:
            ptrs.append(0)
            strides += [0, 0, 0]
            continue
        assert pred.dtype == torch.float32 and pred.is_contiguous() and tuple(pred.shape) == (b, 3, h, w)
        ptr, sb, sc, sh = _crop_view(tgt, h, w)
        ptrs.append(ptr)
        strides += [sb, sc, sh]
    sums = torch.empty(4, dtype=torch.float32, device=ref.device)
    sgn_d = torch.empty_like(r_d) if (want_signs and r_d is not None) else None
    sgn_s = torch.empty_like(r_s) if (want_signs and r_s is not None) else None
    _run(lib.wcmc_image_losses, "image_losses", ref.numel() * 4.0 * 8, _p(r_d), _p(r_s), _p(rad), ptrs[0], ptrs[1], ptrs[2],
         (c_long * 9)(*strides), b, h, w, float(eps), _p(sgn_d), _p(sgn_s), sums.data_ptr(), ws.data_ptr(), ws.numel(),
         _stream())
    return sums, sgn_d, sgn_s


def random_permutation(n, state, salt=0, out=None):
    """-> int64 (n,) pseudo-random permutation of [0, n) (keyed Feistel network, no sort).  state: int64 (2,) device
    tensor [draw counter, 0]; every call advances the counter on the device (CUDA-graph replays draw anew)."""
    lib = init(state.device)
    assert state.dtype == torch.int64 and state.numel() == 2 and state.is_contiguous()
    if out is None:
        out = torch.empty(n, dtype=torch.int64, device=state.device)
    _run(lib.wcmc_random_permutation, "random_permutation", n * 8.0, out.data_ptr(), n, state.data_ptr(), int(salt) & 0xFFFFFFFF,
         _stream())
    return out


def tuning_set(name, value):
    """Measurement aid (tools/, A/B runs): forwards to wcmc_tuning_set."""
    _check(init().wcmc_tuning_set(name.encode(), int(value)), "wcmc_tuning_set")


def grad_exchange_flag_bytes():
    return int(init().wcmc_grad_exchange_flag_bytes())


def grad_exchange(buf, multicast_ptr, peers_dev_ptr, flag_offset_bytes, offset, n, rank, world, channel, scale):
    """In-place sum x scale over the ranks of buf[offset:offset+n] (fp32 symmetric buffer; wcmc_b200/ddp.py owns the
    mappings).  One kernel per rank, synchronised across the ranks by flags inside the buffer."""
    lib = init(buf.device)
    assert buf.dtype == torch.float32 and buf.is_contiguous()
    _run(lib.wcmc_grad_exchange, "grad_exchange", n * 4.0, buf.data_ptr(), int(multicast_ptr) or None, int(peers_dev_ptr),
         int(flag_offset_bytes), int(offset), int(n), int(rank), int(world), int(channel), float(scale), _stream())


# ---- K6/K7: fused PathNet MLPs ----------------------------------------------------------------------
def pathnet_embed_fwd(paths, packed, acts, slope, emb, emb_coff, mean, x16=None, h1=None, h2=None):
    """paths (B,S,Cin,H,W) fp32; packed = [(w_fwd, _, bias_p)] x 3 from pack_weights_batch; acts = 3
    activation codes.  Writes emb[..., emb_coff:emb_coff+64] (B*S,H,W,cs), mean (B,H,W,64) and, when
    given, the 16-bit input copy x16 (B*S,H,W,cin_p) and the hidden activations h1, h2 (B*S,H,W,64)."""
    lib = init(paths.device)
    b, s, cin, h, w = paths.shape
    assert paths.dtype == torch.float32 and paths.is_contiguous()
    (w1, _, b1), (w2, _, b2), (w3, _, b3) = packed
    cin_p = w1.shape[2]
    assert tuple(w1.shape) == (64, 1, cin_p) and tuple(w2.shape) == (64, 1, 64) and tuple(w3.shape) == (64, 1, 64)
    for t in (emb, mean, x16, h1, h2):
        assert t is None or (t.dtype == w1.dtype and t.is_contiguous())
    for t in (x16, h1, h2):
        assert t is None or t.shape[-1] == 64
    hw = h * w
    work = b * s * hw * (cin * 4.0 + 64 * 2.0 * (1 + (h1 is not None) + (h2 is not None)) +
                         (0 if x16 is None else x16.shape[-1] * 2.0)) + b * hw * 128.0
    _run(lib.wcmc_pathnet_embed_fwd, "pathnet_embed_fwd", work, paths.data_ptr(), b, s, cin, hw, w1.data_ptr(),
         b1.data_ptr(), w2.data_ptr(), b2.data_ptr(), w3.data_ptr(), b3.data_ptr(), cin_p, _dt(w1), acts[0], acts[1],
         acts[2], float(slope), _p(x16), _p(h1), _p(h2), emb.data_ptr(),
         emb.shape[-1], emb_coff, _p(mean), 0 if mean is None else mean.shape[-1], 0, _stream())


def pathnet_final_fwd(emb, emb_coff, prop, prop_coff, packed, acts, slope, outc, b, s, hfin=None):
    """emb (B*S,H,W,cs) / prop (B,H,W,cs') 16-bit NHWC -> out (B,S,outc,H,W) fp32 [+ hfin (B*S,H,W,128)]."""
    lib = init(emb.device)
    bs, h, w, ecs = _h16(emb).shape
    assert bs == b * s and tuple(_h16(prop).shape[:3]) == (b, h, w) and prop.dtype == emb.dtype
    (w1, _, b1), (w2, _, b2) = packed
    outc_p = w2.shape[0]
    assert tuple(w1.shape) == (128, 1, 128) and tuple(w2.shape) == (outc_p, 1, 128) and w1.dtype == emb.dtype
    assert hfin is None or (hfin.dtype == emb.dtype and hfin.is_contiguous() and hfin.shape[-1] == 128)
    out = torch.empty((b, s, outc, h, w), dtype=torch.float32, device=emb.device)
    hw = h * w
    work = bs * hw * (128.0 + outc * 4.0 + (0.0 if hfin is None else 256.0)) + b * hw * 128.0
    _run(lib.wcmc_pathnet_final_fwd, "pathnet_final_fwd", work, emb.data_ptr(), ecs, emb_coff, prop.data_ptr(),
         prop.shape[-1], prop_coff, w1.data_ptr(), b1.data_ptr(), w2.data_ptr(), b2.data_ptr(), outc, outc_p,
         _dt(emb), acts[0], acts[1], float(slope), _p(hfin), out.data_ptr(), b, s, hw, _stream())
    return out


# ---- preprocessing of raw sample buffers -------------------------------------------
def preprocess_kpcn(raw):
    """raw (H,W,S,104) fp32 cuda -> (H,W,44) fp32: DenoiseDataset._preprocess_kpcn (datasets.py:487-582)."""
    lib = init(raw.device)
    assert raw.dtype == torch.float32 and raw.is_contiguous() and raw.dim() == 4 and raw.shape[3] == 104
    h, w, s, _ = raw.shape
    out = torch.empty((h, w, 44), dtype=torch.float32, device=raw.device)
    need = lib.wcmc_preprocess_kpcn_workspace(h, w)
    ws = torch.empty(need, dtype=torch.uint8, device=raw.device)
    LAUNCHES["count"] += 1
    _run(lib.wcmc_preprocess_kpcn, "preprocess_kpcn", raw.numel() * 4.0 * 0.3 + out.numel() * 4.0, raw.data_ptr(), h, w, s,
         out.data_ptr(), ws.data_ptr(), need, _stream())
    return out


def preprocess_llpm(raw):
    """raw (H,W,S,104) fp32 cuda -> (H,W,S,37) fp32: DenoiseDataset._preprocess_llpm (datasets.py:301-361)."""
    lib = init(raw.device)
    assert raw.dtype == torch.float32 and raw.is_contiguous() and raw.dim() == 4 and raw.shape[3] == 104
    h, w, s, _ = raw.shape
    out = torch.empty((h, w, s, 37), dtype=torch.float32, device=raw.device)
    _run(lib.wcmc_preprocess_llpm, "preprocess_llpm", h * w * s * (176.0 + 148.0), raw.data_ptr(), h * w * s,
         out.data_ptr(), _stream())
    return out


# ---- batched weight normalisation ------------------------------------------------------------------
def weight_norm_batch_fwd(vs, gs):
    """vs[i] (cout, ...) / gs[i] (cout, 1, ...) fp32 -> (ws, norms): torch._weight_norm(v, g, 0) for every layer, one launch."""
    lib = init(vs[0].device)
    n = len(vs)
    descs = (WnDesc * n)()
    ws, norms = [], []
    for i, (v, g) in enumerate(zip(vs, gs)):
        assert v.dtype == torch.float32 and v.is_contiguous() and g.dtype == torch.float32 and g.is_contiguous()
        rows = v.shape[0]
        assert g.numel() == rows
        w = torch.empty_like(v)
        nrm = torch.empty(rows, dtype=torch.float32, device=v.device)
        descs[i] = WnDesc(v.data_ptr(), g.data_ptr(), w.data_ptr(), nrm.data_ptr(), 0, 0, 0, rows, v.numel() // rows)
        ws.append(w)
        norms.append(nrm)
    _run(lib.wcmc_weight_norm_batch, "weight_norm_fwd", sum(v.numel() for v in vs) * 8.0, descs, n, 0, _stream())
    return ws, norms


def weight_norm_batch_bwd(dws, vs, gs, norms):
    """-> (dvs, dgs) for every layer, one launch."""
    lib = init(vs[0].device)
    n = len(vs)
    descs = (WnDesc * n)()
    dvs, dgs, keep = [], [], []
    for i, (dw, v, g, nrm) in enumerate(zip(dws, vs, gs, norms)):
        if dw.dtype != torch.float32 or not dw.is_contiguous():
            dw = dw.float().contiguous()
        keep.append(dw)
        rows = v.shape[0]
        dv, dg = torch.empty_like(v), torch.empty_like(g)
        descs[i] = WnDesc(v.data_ptr(), g.data_ptr(), 0, nrm.data_ptr(), dw.data_ptr(), dv.data_ptr(), dg.data_ptr(), rows,
                          v.numel() // rows)
        dvs.append(dv)
        dgs.append(dg)
    _run(lib.wcmc_weight_norm_batch, "weight_norm_bwd", sum(v.numel() for v in vs) * 12.0, descs, n, 1, _stream())
    return dvs, dgs


# ---- K8/K9: fused PathNet MLP backward passes -------------------------------------------------------
def pathnet_final_bwd(g, out, gscale, inv_scale, emb, prop, hfin, packed, acts, slope, outc, b, s):
    """g, out (B,S,outc,H,W) fp32; emb (B*S,H,W,cs>=64) [channels 0..63], prop (B,H,W,cs') [0..63], hfin (B*S,H,W,128)
    16-bit; packed = [(_, w_dgrad, _)] x 2.  -> (d_emb (B*S,H,W,64), d_prop (B,H,W,64) 16-bit loss-scaled,
    [dw1, db1, dw2, db2] fp32 in torch layout)."""
    lib = init(emb.device)
    bs, h, w, ecs = _h16(emb).shape
    assert bs == b * s and tuple(_h16(prop).shape[:3]) == (b, h, w) and prop.dtype == emb.dtype
    assert g.dtype == torch.float32 and g.is_contiguous() and out.dtype == torch.float32 and out.is_contiguous()
    assert tuple(g.shape) == tuple(out.shape) == (b, s, outc, h, w)
    assert hfin.dtype == emb.dtype and hfin.is_contiguous() and hfin.shape[-1] == 128
    (_, w1t, _), (_, w2t, _) = packed
    outc_p = w2t.shape[2]
    assert tuple(w1t.shape) == (128, 1, 128) and tuple(w2t.shape) == (128, 1, outc_p) and w1t.dtype == emb.dtype
    dev = emb.device
    d_emb = torch.empty((bs, h, w, 64), dtype=emb.dtype, device=dev)
    d_prop = torch.empty((b, h, w, 64), dtype=emb.dtype, device=dev)
    dw1 = torch.empty((128, 128, 1, 1), dtype=torch.float32, device=dev)
    db1 = torch.empty((128,), dtype=torch.float32, device=dev)
    dw2 = torch.empty((outc, 128, 1, 1), dtype=torch.float32, device=dev)
    db2 = torch.empty((outc,), dtype=torch.float32, device=dev)
    need = lib.wcmc_pathnet_bwd_workspace(0)
    ws = torch.empty(need, dtype=torch.uint8, device=dev)
    hw = h * w
    work = bs * hw * (outc * 8.0 + 256.0 + 128.0 + 128.0) + b * hw * 256.0
    _run(lib.wcmc_pathnet_final_bwd, "pathnet_final_bwd", work, g.data_ptr(), out.data_ptr(), _p(gscale), _p(inv_scale),
         emb.data_ptr(), ecs, prop.data_ptr(), prop.shape[-1], hfin.data_ptr(), w1t.data_ptr(), w2t.data_ptr(), outc,
         outc_p, _dt(emb), acts[0], acts[1], float(slope), d_emb.data_ptr(), d_prop.data_ptr(), dw1.data_ptr(),
         db1.data_ptr(), dw2.data_ptr(), db2.data_ptr(), b, s, hw, ws.data_ptr(), need, _stream())
    return d_emb, d_prop, [dw1, db1, dw2, db2]


def pathnet_embed_bwd(d_emb, d_red, inv_scale, emb, h2, h1, x16, packed, acts, slope, cin, b, s):
    """d_emb (B*S,H,W,64), d_red (B,H,W,>=64 contiguous 64-channel) or None, saved activations (64 channels each);
    packed = [(_, w_dgrad, _)] x 3 (layer order 1, 2, 3).  -> [dw1, db1, dw2, db2, dw3, db3] fp32."""
    lib = init(d_emb.device)
    bs, h, w, _ = _h16(d_emb).shape
    assert bs == b * s and d_emb.shape[-1] == 64
    for t in (h2, h1, x16):
        assert t.dtype == d_emb.dtype and t.is_contiguous() and tuple(t.shape) == (bs, h, w, 64)
    assert emb.dtype == d_emb.dtype and emb.is_contiguous() and tuple(emb.shape[:3]) == (bs, h, w)
    assert d_red is None or (d_red.dtype == d_emb.dtype and d_red.is_contiguous() and tuple(d_red.shape) == (b, h, w, 64))
    (_, _w1t, _), (_, w2t, _), (_, w3t, _) = packed
    assert tuple(w2t.shape) == (64, 1, 64) and tuple(w3t.shape) == (64, 1, 64)
    dev = d_emb.device
    f32 = dict(dtype=torch.float32, device=dev)
    dw1, dw2, dw3 = torch.empty((64, cin, 1, 1), **f32), torch.empty((64, 64, 1, 1), **f32), torch.empty((64, 64, 1, 1), **f32)
    db1, db2, db3 = torch.empty((64,), **f32), torch.empty((64,), **f32), torch.empty((64,), **f32)
    need = lib.wcmc_pathnet_bwd_workspace(1)
    ws = torch.empty(need, dtype=torch.uint8, device=dev)
    hw = h * w
    work = bs * hw * 5 * 128.0 + b * hw * 128.0 * s
    _run(lib.wcmc_pathnet_embed_bwd, "pathnet_embed_bwd", work, d_emb.data_ptr(), _p(d_red), _p(inv_scale),
         emb.data_ptr(), emb.shape[-1], h2.data_ptr(), h1.data_ptr(), x16.data_ptr(), w3t.data_ptr(), w2t.data_ptr(), cin,
         _dt(d_emb), acts[0], acts[1], acts[2], float(slope), dw3.data_ptr(), db3.data_ptr(), dw2.data_ptr(),
         db2.data_ptr(), dw1.data_ptr(), db1.data_ptr(), b, s, hw, ws.data_ptr(), need, _stream())
    return [dw1, db1, dw2, db2, dw3, db3]


# ---- K12: fused clip + Adam ------------------------------------------------------------------------
def adam_clip_step(dev_tensors, dev_blocks, nblocks, dev_step, ok_flag, clip, nbytes=0.0, nonfinite=None):
    """dev_tensors: uint8 device tensor holding AdamTensor structs; dev_blocks: int32 (nblocks, 2); nonfinite: optional
    int64 (1,) device counter of skipped NaN / inf gradient elements."""
    lib = init(dev_step.device)
    _run(lib.wcmc_adam_clip_step, "adam_clip_step", nbytes, dev_tensors.data_ptr(), dev_blocks.data_ptr(), nblocks,
         dev_step.data_ptr(), _p(ok_flag), float(clip), _p(nonfinite), _stream())


# ---- K11: all-pairs loss (extension) --------------------------------------------------------------------
def fmse_allpairs_fwd(p_rows, ref_rows, mode=0, alpha=2.0, tau=0.0):
    """p_rows (N,D), ref_rows (N,3) fp32 contiguous -> (out (2,) = [loss, kept ordered pairs], nonfinite flag)."""
    lib = init(p_rows.device)
    n, d = p_rows.shape
    assert p_rows.dtype == torch.float32 and p_rows.is_contiguous()
    assert ref_rows.dtype == torch.float32 and ref_rows.is_contiguous() and tuple(ref_rows.shape) == (n, 3)
    out = torch.empty(2, dtype=torch.float32, device=p_rows.device)
    flag = torch.zeros(1, dtype=torch.int32, device=p_rows.device)
    need = lib.wcmc_fmse_allpairs_workspace(n, d)
    ws = _workspace(need + 256, p_rows.device)
    off = (-ws.data_ptr()) % 256
    _run(lib.wcmc_fmse_allpairs_fwd, "fmse_allpairs_fwd", 2.0 * n * n * (d + 3) / 2, p_rows.data_ptr(),
         ref_rows.data_ptr(), n, d, int(mode), float(alpha), float(tau or 0.0), out.data_ptr(), flag.data_ptr(),
         ws.data_ptr() + off, need, _stream())
    return out, flag
