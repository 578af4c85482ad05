"""All-pairs form of the path-disentangling loss (extension; K11, wcmc_fmse_allpairs_fwd).

The reference's `FeatureMSE` (/root/reference/support/losses.py:33-61) pairs every embedded sample
with ONE random partner; this module evaluates the same displacement error over ALL pairs with the
N x N Gram matrix computed tile by tile on the tensor cores and reduced in the GEMM epilogue
(BASELINE.json north_star (4), configs[4]).  It is NOT what `train_kpcn.py` runs -- use
`support.losses.FeatureMSE` for that -- and is kept separate for that reason.

    loss, kept = allpairs_loss(p_rows, ref_rows, mode="mse" | "lse", alpha=2.0, tau=None)

`p_rows` (N, D) embeddings, `ref_rows` (N, 3) reference radiance (tone-mapped inside), `tau`: optional
weak-label threshold on 1/2 |t_i - t_j|^2.  The unmasked mse loss is differentiable w.r.t. `p_rows`
(its gradient has a closed form in the second moments of the rows: O(N D^2), no second N x N pass);
the masked and lse variants are forward-only.
"""
import torch

from . import lib

MODES = {"mse": 0, "lse": 1}


def rows_from_pbuffer(p_buffer, ref):
    """(B,S,C,H,W), (B,3,H,W) -> (N,C), (N,3) rows in the reference's (b,s,y,x) order (losses.py:88-97)."""
    b, s, c, h, w = p_buffer.shape
    p = p_buffer.permute(0, 1, 3, 4, 2).reshape(-1, c).contiguous()
    r = ref.permute(0, 2, 3, 1).unsqueeze(1).expand(b, s, h, w, 3).reshape(-1, 3).contiguous()
    return p, r


class _AllPairsMSE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, p_rows, ref_rows):
        out, flag = lib.fmse_allpairs_fwd(p_rows, ref_rows, 0)
        ctx.save_for_backward(p_rows, ref_rows)
        ctx.mark_non_differentiable(flag)
        return out[0], out[1].detach(), flag

    @staticmethod
    def backward(ctx, g, _gk, _gf):
        # dL/dP_i = (2/N^2) [ r_i P_i - sum_j e_ij P_j ],  e_ij = (hp_i - ht_i) + (hp_j - ht_j) - P_i.P_j + t_i.t_j
        p, ref = ctx.saved_tensors
        n = p.shape[0]
        P = p.double()
        r = ref.double().clamp(min=0)
        T = (r / (1 + r)) ** 0.454545
        hd = 0.5 * (P * P).sum(1) - 0.5 * (T * T).sum(1)                     # (N,)
        sP, sT, shd = P.sum(0), T.sum(0), hd.sum()
        row = n * hd + shd - P @ sP + T @ sT                                  # r_i = sum_j e_ij
        eP = hd[:, None] * sP[None] + (hd[:, None] * P).sum(0)[None] - P @ (P.t() @ P) + T @ (T.t() @ P)
        grad = (2.0 / (float(n) * n)) * (row[:, None] * P - eP)
        return (g * grad).to(p.dtype), None


def allpairs_loss(p_rows, ref_rows, mode="mse", alpha=2.0, tau=None):
    """-> (loss scalar tensor, number of kept ordered pairs as a 0-d tensor).  Raises RuntimeError on
    non-finite input like the reference's loss does (losses.py:99-102)."""
    if not p_rows.is_cuda:
        raise lib.WcmcError("allpairs_loss runs on the B200 only; there is no CPU fallback")
    p_rows = p_rows.float().contiguous()
    ref_rows = ref_rows.to(p_rows.device, torch.float32).contiguous()
    if mode == "mse" and not tau:
        loss, kept, flag = _AllPairsMSE.apply(p_rows, ref_rows)
    else:
        if p_rows.requires_grad:
            raise NotImplementedError("the masked / lse all-pairs variants are forward-only")
        out, flag = lib.fmse_allpairs_fwd(p_rows, ref_rows, MODES[mode], alpha, tau or 0.0)
        loss, kept = out[0], out[1]
    if int(flag) != 0:
        raise RuntimeError("Infinite loss at train time.")
    return loss, kept
