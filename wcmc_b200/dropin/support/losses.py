"""B200 drop-in for the hot-path losses of /root/reference/support/losses.py.

FeatureMSE (path disentangling loss, losses.py:9-113), GlobalRelativeSimilarityLoss (:116-211) and
RelativeMSE (:245-264).  The displacement / loss arithmetic runs in two CUDA launches
(wcmc_fmse_perm_fwd / _bwd, wcmc_b200/csrc/fmse.cu) on the cropped p-buffer view as it is -- no
permute / reshape copies.  RNG contract kept from the reference: the pairing permutations come from
the CPU default generator, `randperm(S*H*W)` then `randperm(B*S*H*W)` per call (losses.py:35, :50),
so a seeded run pairs exactly the same samples as the reference.  CUDA tensors only.
"""
import math
import os

import torch

from wcmc_b200 import lib

__all__ = ["GlobalRelativeSimilarityLoss", "FeatureMSE", "RelativeMSE"]


_PERM_STATE = {}    # device index -> int64 (2,) [draw counter, ticket] of wcmc_random_permutation (the launch's
                    # ticket is not re-entrant: every draw of a device goes through ONE stream, the loss's)


def _randperm(n, device, rng):
    """rng='cpu': the reference's contract (CPU default generator, then a host->device copy; 13 ms
    of host time for n = 541,696).  rng='device': a keyed Feistel bijection computed on the GPU
    (wcmc_random_permutation: no sort, one launch, a new draw on every CUDA-graph replay) -- a different random
    stream than the reference's, used by the throughput benchmark.  The counter is seeded from torch's CPU
    generator at first use, so torch.manual_seed() makes runs repeatable."""
    if rng == "device":
        st = _PERM_STATE.get(device.index)
        if st is None:
            seed = int(torch.randint(0, 2 ** 62, (1,)).item())
            st = _PERM_STATE[device.index] = torch.tensor([seed, 0], dtype=torch.int64, device=device)
        return lib.random_permutation(n, st, salt=n)
    return torch.randperm(n).to(device, non_blocking=True)


def _default_rng():
    return os.environ.get("WCMC_PERM_RNG", "cpu")


# When True the finite flag of the last call is left on the device in `FINITE_FLAGS` instead of being
# checked with a host sync: wcmc_b200.engine.GraphedTrainStep (CUDA-graph capture) reads it once per step.
DEFER_FINITE_CHECK = False
FINITE_FLAGS = []


def _check_flag(nonfinite):
    ok = nonfinite[0] == 0
    if DEFER_FINITE_CHECK:
        FINITE_FLAGS.append(ok)
        return
    if not bool(ok):
        raise RuntimeError("Infinite loss at train time.")


def _views(p_buffer, ref):
    """The kernels read strided (cropped) fp32 views directly; anything else is normalised here."""
    if not p_buffer.is_cuda:
        raise lib.WcmcError("the path-disentangling loss runs on the B200 only (got a %s tensor); there is no CPU "
                            "fallback" % p_buffer.device)
    if p_buffer.dtype != torch.float32:
        p_buffer = p_buffer.float()
    if p_buffer.stride(4) != 1:
        p_buffer = p_buffer.contiguous()
    ref = ref.to(p_buffer.device, torch.float32)
    if ref.stride(3) != 1:
        ref = ref.contiguous()
    return p_buffer, ref


class PermStage:
    """Pre-staged pairing permutations: the reference's RNG contract (CPU default generator, `randperm(S*H*W)` then
    `randperm(B*S*H*W)` per loss call, losses.py:35, :50) under a captured CUDA graph.  A graph cannot draw from the
    CPU generator, so every loss call of a step owns a slot of STATIC device index buffers; `begin_step()` (called
    by wcmc_b200.engine.GraphedTrainStep before each replay) draws the step's permutations on the host, in the
    order the calls will consume them, into pinned memory and copies them over.  The graph only reads the buffers."""

    def __init__(self):
        self.slots = []      # one per loss call of a step, in call order
        self.cursor = 0

    @staticmethod
    def _draw(slot):
        for host, dev in slot:
            if host is not None:
                torch.randperm(host.numel(), out=host)          # CPU default generator, like the reference
                dev.copy_(host, non_blocking=True)

    def begin_step(self):
        """Draws this step's permutations for every known slot (host work: ~13 ms per 541,696 elements)."""
        self.cursor = 0
        for slot in self.slots:
            self._draw(slot)

    def take(self, n_patch, n_batch, device):
        if self.cursor == len(self.slots):
            if torch.cuda.is_current_stream_capturing():
                raise RuntimeError("PermStage: a loss call without a staged slot inside a CUDA-graph capture")
            mk = lambda n: (None, None) if not n else (  # noqa: E731
                torch.empty(n, dtype=torch.int64).pin_memory(), torch.empty(n, dtype=torch.int64, device=device))
            slot = (mk(n_patch), mk(n_batch))
            self._draw(slot)                 # first use: drawn right here, i.e. still in call order
            self.slots.append(slot)
        (hp, dp), (hb, db) = self.slots[self.cursor]
        self.cursor += 1
        assert dp.numel() == n_patch and (db is None) == (not n_batch) and (db is None or db.numel() == n_batch), \
            "PermStage: the loss calls of a step changed shape"
        return dp, db


def _perms(p, rng, non_local, idx_patch, idx_batch, stage=None):
    b, s, c, h, w = p.shape
    if stage is not None and idx_patch is None and idx_batch is None:
        return stage.take(s * h * w, b * s * h * w if non_local else 0, p.device)
    if idx_patch is None:
        idx_patch = _randperm(s * h * w, p.device, rng)
    if non_local and idx_batch is None:
        idx_batch = _randperm(b * s * h * w, p.device, rng)
    fix = lambda i: None if i is None else i.to(p.device, torch.int64).contiguous()  # noqa: E731
    return fix(idx_patch), fix(idx_batch) if non_local else None


class _FmseFn(torch.autograd.Function):
    """loss = 1/2 mean e_patch^2 + 1/2 mean e_batch^2 (local only: the patch term twice), one fused
    forward launch (wcmc_fmse_perm_fwd) and one gather-only backward launch (wcmc_fmse_perm_bwd)."""

    @staticmethod
    def forward(ctx, p, ref, idx_patch, idx_batch, crop=None):
        # crop = (y0, x0, h, w): `p` is the UNcropped p-buffer and the centred crop of crop_like() is taken here, as a
        # strided view the kernels read directly; the backward then writes straight into the interior of a zero-filled
        # tensor of p's shape (autograd's own slice-backward would allocate, fill and copy once per cropped dimension)
        ctx.crop = crop
        pv = p if crop is None else p[..., crop[0]:crop[0] + crop[2], crop[1]:crop[1] + crop[3]]
        r = lib.fmse_perm_fwd(pv, ref, idx_patch, idx_batch)
        _check_flag(r["nonfinite"])
        ctx.save_for_backward(p, idx_patch, idx_batch, r["inv_patch"], r["inv_batch"], r["e_patch"], r["e_batch"])
        loss = r["loss"]
        return loss[0] + (loss[1] if idx_batch is not None else loss[0])

    @staticmethod
    def backward(ctx, g):
        p, idx_patch, idx_batch, inv_patch, inv_batch, e_patch, e_batch = ctx.saved_tensors
        rows = e_patch.numel()
        coef_p = (1.0 if idx_batch is not None else 2.0) / rows
        crop = ctx.crop
        pv, out, full = p, None, None
        if crop is not None:
            pv = p[..., crop[0]:crop[0] + crop[2], crop[1]:crop[1] + crop[3]]
            full = torch.zeros(p.shape, dtype=torch.float32, device=p.device)
            out = full[..., crop[0]:crop[0] + crop[2], crop[1]:crop[1] + crop[3]]
        dp = lib.fmse_perm_bwd(pv, idx_patch, idx_batch, inv_patch, inv_batch, e_patch, e_batch,
                               g.reshape(1).float().contiguous(), coef_p, 1.0 / rows, out=out)
        return (dp if full is None else full), None, None, None, None


class _PairDisplacementFn(torch.autograd.Function):
    """-> (e_patch, e_batch) displacement vectors with a gradient to p (used by GRS)."""

    @staticmethod
    def forward(ctx, p, ref, idx_patch, idx_batch):
        r = lib.fmse_perm_fwd(p, ref, idx_patch, idx_batch)
        _check_flag(r["nonfinite"])
        ctx.save_for_backward(p, idx_patch, idx_batch, r["inv_patch"], r["inv_batch"])
        return r["e_patch"], r["e_batch"]

    @staticmethod
    def backward(ctx, g_p, g_b):
        p, idx_patch, idx_batch, inv_patch, inv_batch = ctx.saved_tensors
        dp = lib.fmse_perm_bwd(p, idx_patch, idx_batch, inv_patch, inv_batch, g_p.contiguous(), g_b.contiguous(),
                               None, 1.0, 1.0)
        return dp, None, None, None


class FeatureMSE(torch.nn.Module):
    """Path disentangling loss: for random sample pairs (i, pi(i)) penalise
    (1/2|p_i - p_j|^2 - 1/2|t_i - t_j|^2)^2, once with pairs inside each patch and (non_local)
    once with pairs across the whole batch.  `idx_patch` / `idx_batch` (optional, extension) pin the
    pairing permutations; by default they are drawn exactly as the reference draws them."""

    def __init__(self, color="rgb", non_local=True, rng=None):
        super().__init__()
        if color != "rgb":
            raise NotImplementedError("color='hls' is never selected by the reference scripts")
        self.color = color
        self.non_local = non_local
        self.rng = rng or _default_rng()
        assert self.rng in ("cpu", "device")
        self.stage = None    # PermStage while a GraphedTrainStep drives this loss with rng="cpu"
        print("FeatureMSE locality: %s" % ("Non-local" if non_local else "Local"))

    def forward(self, p_buffer, ref, idx_patch=None, idx_batch=None):
        p, r = _views(p_buffer, ref)
        idx_patch, idx_batch = _perms(p, self.rng, self.non_local, idx_patch, idx_batch, self.stage)
        return _FmseFn.apply(p, r, idx_patch, idx_batch, None)

    def forward_cropped(self, p_buffer, ref, idx_patch=None, idx_batch=None):
        """forward(crop_like(p_buffer, ref), ref) with the centred crop (support/utils.py:24-42) taken inside the
        autograd Function: same value and gradient, no slice-backward copies of the (B,S,C,H,W) gradient."""
        h, w = ref.shape[-2:]
        dh, dw = p_buffer.shape[-2] - h, p_buffer.shape[-1] - w
        if dh < 0 or dw < 0 or (dh == 0 and dw == 0) or not p_buffer.is_cuda or p_buffer.dtype != torch.float32 \
                or p_buffer.stride(4) != 1:
            from support.utils import crop_like
            return self.forward(crop_like(p_buffer, ref), ref, idx_patch, idx_batch)
        crop = (max(dh // 2, 0), max(dw // 2, 0), h, w)
        r = ref.to(p_buffer.device, torch.float32)
        if r.stride(3) != 1:
            r = r.contiguous()
        pv = p_buffer[..., crop[0]:crop[0] + h, crop[1]:crop[1] + w]
        idx_patch, idx_batch = _perms(pv, self.rng, self.non_local, idx_patch, idx_batch, self.stage)
        return _FmseFn.apply(p_buffer, r, idx_patch, idx_batch, crop)


class GlobalRelativeSimilarityLoss(torch.nn.Module):
    """(LSE(alpha * [d_p, d_b, -d_p, -d_b, 0]) - log(1 + 4N)) / sqrt(alpha)   (losses.py:185-211)."""

    def __init__(self, alpha=2, color="rgb", rng=None):
        super().__init__()
        self.color = color
        self.alpha = alpha
        self.rng = rng or _default_rng()
        self.stage = None

    def forward(self, p_buffer, ref, idx_patch=None, idx_batch=None):
        p, r = _views(p_buffer, ref)
        b, s, c, h, w = p.shape
        idx_patch, idx_batch = _perms(p, self.rng, True, idx_patch, idx_batch, self.stage)
        d_p, d_b = _PairDisplacementFn.apply(p, r, idx_patch, idx_batch)
        zero = torch.zeros(1, dtype=p.dtype, device=p.device)
        ex = self.alpha * torch.cat([d_p, d_b, -d_p, -d_b, zero], 0)
        return (torch.logsumexp(ex, 0) - math.log(1 + 4 * b * s * h * w)) / math.sqrt(self.alpha)


class RelativeMSE(torch.nn.Module):
    """0.5 * mean((im - ref)^2 / (ref^2 + eps))   (losses.py:245-264)."""

    def __init__(self, eps=1e-2):
        super().__init__()
        self.eps = eps

    def forward(self, im, ref):
        if (im.is_cuda and ref.is_cuda and not (torch.is_grad_enabled() and (im.requires_grad or ref.requires_grad))
                and im.dim() == 4 and im.shape[1] == 3 and im.shape == ref.shape and im.dtype == torch.float32
                and ref.dtype == torch.float32 and ref.stride(3) == 1):
            # metric use (validate_batch, the `rmse` entry of the step): one reduction launch (K12, wcmc_image_losses)
            sums, _, _ = lib.image_losses(None, None, None, None, im.detach().contiguous(), ref.detach(), self.eps, False)
            return sums[3]
        return 0.5 * torch.mean((im - ref) ** 2 / (ref ** 2 + self.eps))


# ---- the remaining image metrics of support/losses.py (used by train_sbmc.py / train_lbmc.py and as test metrics):
# a handful of elementwise torch ops on (B,3,h,w) images, kept so that `from support.losses import ...` of any
# reference script resolves against this package -------------------------------------------------------------
def _tonemap(im):
    """Reinhard: clamp(im, 0) / (1 + clamp(im, 0))   (losses.py:234-242)."""
    im = torch.clamp(im, min=0)
    return im / (1 + im)


class SMAPE(torch.nn.Module):
    """mean(|im - ref| / (eps + |im| + |ref|)), denominator detached   (losses.py:267-284)."""

    def __init__(self, eps=1e-2):
        super().__init__()
        self.eps = eps

    def forward(self, im, ref):
        return (torch.abs(im - ref) / (self.eps + im.detach().abs() + ref.detach().abs())).mean()


class TonemappedMSE(torch.nn.Module):
    """0.5 * mean((T(im) - T(ref))^2)   (losses.py:287-302)."""

    def __init__(self, eps=1e-2):
        super().__init__()
        self.eps = eps

    def forward(self, im, ref):
        return 0.5 * torch.mean((_tonemap(im) - _tonemap(ref)) ** 2)


class TonemappedRelativeMSE(torch.nn.Module):
    """RelativeMSE on Reinhard-tonemapped images   (losses.py:305-320)."""

    def __init__(self, eps=1e-2):
        super().__init__()
        self.eps = eps

    def forward(self, im, ref):
        im, ref = _tonemap(im), _tonemap(ref)
        return 0.5 * torch.mean((im - ref) ** 2 / (ref ** 2 + self.eps))
