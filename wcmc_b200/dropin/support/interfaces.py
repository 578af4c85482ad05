"""B200 drop-in for the KPCN train / validate step of /root/reference/support/interfaces.py
(`BaseInterface` :18-77, `KPCNInterface` :80-333).

Same constructor, public methods, attributes and `m_losses` keys as the reference, so
`train_kpcn.py` drives it unchanged.  Differences are host-side only:
  * the per-loss `isfinite` checks (six device->host syncs per step in the reference, :255-257)
    are folded into ONE sync per step;
  * `grad_sync`, when set (wcmc_b200.ddp), runs between the two backward passes and the gradient
    clipping, which is where a data-parallel all-reduce has to sit (:237-238 -> :261 -> :271).
"""
import os
from abc import ABCMeta, abstractmethod

import torch
import torch.nn as nn

from support.utils import crop_like

_DISENTANGLE = ("m11r11", "m10r01", "m11r01", "m10r11")
# 1: the path-disentangling losses run while the KPCN branches run; default 0: after them, as the reference orders
# the calls.  Measured 7.13 -> 7.09 ms per step on one B200 (three runs, GPU tests green with it on); off by default
# because that is within what boxes differ by and the default path is the one every multi-GPU run of round 2 used.
OVERLAP_MANIF = os.environ.get("WCMC_OVERLAP_MANIF", "0") != "0"


class BaseInterface(metaclass=ABCMeta):
    def __init__(self, models, optims, loss_funcs, args, visual=False, use_llpm_buf=False, manif_learn=False,
                 w_manif=0.1):
        self.models, self.optims, self.loss_funcs, self.args = models, optims, loss_funcs, args
        self.visual, self.use_llpm_buf, self.manif_learn, self.w_manif = visual, use_llpm_buf, manif_learn, w_manif
        self.iters = 0
        self.m_losses = {}
        self.best_err = 1e10
        self.fixed_batch = None

    @abstractmethod
    def to_train_mode(self):
        pass

    @abstractmethod
    def preprocess(self, batch=None):
        pass

    @abstractmethod
    def train_batch(self, batch):
        pass

    @abstractmethod
    def _manifold_forward(self, batch):
        return {}

    @abstractmethod
    def _regress_forward(self, batch):
        return {}

    @abstractmethod
    def _backward(self, batch, out, p_buffers):
        return {}

    @abstractmethod
    def _logging(self, loss_dict):
        pass

    @abstractmethod
    def _optimization(self):
        pass

    @abstractmethod
    def to_eval_mode(self):
        pass

    @abstractmethod
    def validate_batch(self, batch):
        pass

    @abstractmethod
    def get_epoch_summary(self, mode, norm):
        return 0.0


def _half(p_buffers, lo):
    c = p_buffers["diffuse"].shape[2]
    sl = slice(0, c // 2) if lo else slice(c // 2, None)
    return {k: v[:, :, sl] for k, v in p_buffers.items()}


def _with_pbuffer(batch, p_buffers, channels=None):
    """New batch whose KPCN inputs carry [mean_S(p) | var_S(p).mean(C)/S] (interfaces.py:165-180).
    The mean keeps its graph (gradients reach PathNet through the regression branch); the
    variance channel is detached, as in the reference.  `channels` = (c0, cr): only that channel range of the
    p-buffers feeds the regression (the `m10r01` / `m11r01` options).  On the GPU the statistics and the
    concatenation are one launch (K9, ops.PBufferConcatFn); host-logic tests on CPU tensors keep the expression."""
    new = {k: batch[k] for k in ("target_total", "target_diffuse", "target_specular", "kpcn_diffuse_buffer",
                                 "kpcn_specular_buffer", "kpcn_albedo")}
    for name in ("diffuse", "specular"):
        p = p_buffers[name]
        c0, cr = channels if channels is not None else (0, p.shape[2])
        kin = batch["kpcn_%s_in" % name]
        if p.is_cuda and kin.is_cuda:
            from wcmc_b200 import ops as wops
            new["kpcn_%s_in" % name] = wops.PBufferConcatFn.apply(kin, p, c0, cr)
            continue
        p = p[:, :, c0:c0 + cr]
        var = p.var(1).mean(1, keepdim=True).detach() / p.shape[1]
        new["kpcn_%s_in" % name] = torch.cat([kin, p.mean(1), var], 1)
    return new


class KPCNInterface(BaseInterface):
    def __init__(self, models, optims, loss_funcs, args, visual=False, use_llpm_buf=False, manif_learn=False,
                 w_manif=0.1, train_branches=True, disentanglement_option="m11r11"):
        if manif_learn:
            assert "backbone_diffuse" in models, "argument `models` dictionary should contain `'backbone_diffuse'` key."
            assert "backbone_specular" in models, "argument `models` dictionary should contain `'backbone_specular'` key."
        assert "dncnn" in models, "argument `models` dictionary should contain `'dncnn'` key."
        if train_branches:
            assert "l_diffuse" in loss_funcs
            assert "l_specular" in loss_funcs
        if manif_learn:
            assert "l_manif" in loss_funcs
        assert "l_recon" in loss_funcs
        assert "l_test" in loss_funcs
        assert disentanglement_option in _DISENTANGLE
        super().__init__(models, optims, loss_funcs, args, visual, use_llpm_buf, manif_learn, w_manif)
        self.train_branches = train_branches
        self.disentanglement_option = disentanglement_option
        self.grad_sync = None  # optional callable(models): data-parallel gradient all-reduce
        # clip_grad_value_ + the Adam steps as one kernel when every optimiser is a default torch Adam
        # (WCMC_FUSED_ADAM=0 or `itf.fused_optim = False` keeps torch's own step)
        self.fused_optim = os.environ.get("WCMC_FUSED_ADAM", "1") != "0"
        self._fused_adam = None

    def __str__(self):
        return "KPCNInterface"

    # ---- modes ---------------------------------------------------------------------------------
    def to_train_mode(self):
        for name, model in self.models.items():
            model.train()
            assert "optim_" + name in self.optims, "`optim_%s`: an optimization algorithm is not defined." % name

    def to_eval_mode(self):
        for model in self.models.values():
            model.eval()
        self.m_losses["m_val"] = torch.tensor(0.0)

    def preprocess(self, batch=None):
        for key in ("target_total", "target_diffuse", "target_specular", "kpcn_diffuse_in", "kpcn_specular_in",
                    "kpcn_diffuse_buffer", "kpcn_specular_buffer", "kpcn_albedo"):
            assert key in batch
        if self.use_llpm_buf:
            assert "paths" in batch
        self.iters += 1

    # ---- forward pieces ------------------------------------------------------------------------
    def _manifold_forward(self, batch):
        # The two path-embedding networks are independent: on two CUDA streams the tail of one network's
        # kernels (partial last wave on 148 SMs, small U-Net levels) overlaps the other's.  Autograd runs each
        # backward node on its forward stream, so the backward passes overlap the same way.
        from wcmc_b200 import streams
        with streams.fork("diffuse"):
            d = self.models["backbone_diffuse"](batch)
        with streams.fork("specular"):
            s = self.models["backbone_specular"](batch)
        streams.join()
        return {"diffuse": d, "specular": s}

    def _regress_forward(self, batch):
        return self.models["dncnn"](batch)

    def _dump_pbuffers(self, p_buffers):
        """Every 1000 iterations the reference writes a PNG of the first 3 embedding channels
        (interfaces.py:130-137).  Kept as a best-effort side effect."""
        out_dir = "../LLPM_results"
        if not os.path.isdir(out_dir):
            return
        try:
            import matplotlib.pyplot as plt
        except ImportError:
            return
        for name in ("diffuse", "specular"):
            img = p_buffers[name].detach()[0, :, :3].mean(0).clamp(0.0, 1.0).permute(1, 2, 0).cpu().numpy()
            plt.imsave("%s/pbuf_%s_%s.png" % (out_dir, self.args.model_name, name), img)

    def _split(self, p_buffers):
        """-> (p_buffers fed to the regression, p_buffers fed to the manifold loss)   (:139-163)"""
        assert p_buffers["diffuse"].shape[2] >= 2
        opt = self.disentanglement_option
        reg = _half(p_buffers, lo=True) if opt in ("m10r01", "m11r01") else p_buffers
        man = _half(p_buffers, lo=False) if opt in ("m10r01", "m10r11") else p_buffers
        return reg, man

    def _reg_channels(self, p_buffers):
        """Channel range (c0, cr) of the p-buffers that feeds the regression (the lower half for `*r01`)."""
        c = p_buffers["diffuse"].shape[2]
        return (0, c // 2) if self.disentanglement_option in ("m10r01", "m11r01") else (0, c)

    # ---- one optimisation step -----------------------------------------------------------------
    def train_batch(self, batch, grad_hook_mode=False):
        loss_dict = self._forward_backward(batch, dump=self.iters % 1000 == 1)
        if grad_hook_mode:  # gradients only; no logging, no parameter update
            return
        self._logging(loss_dict)
        self._optimization()

    def _forward_backward(self, batch, dump=False):
        """Forward passes + the two backward passes of `train_batch` (:139-251) -> loss dict; shared with the
        CUDA-graph step (wcmc_b200.engine.GraphedTrainStep captures exactly this)."""
        out_manif = None
        self._p_full = None
        drain = getattr(self.grad_sync, "drain", None)
        if drain is not None:
            drain()        # an exchange started by a step that never reached _logging (gradient-only pass)
        if self.use_llpm_buf:
            self.models["backbone_diffuse"].zero_grad()
            self.models["backbone_specular"].zero_grad()
            p_buffers = self._manifold_forward(batch)
            if dump:
                self._dump_pbuffers(p_buffers)
            _, out_manif = self._split(p_buffers)
            batch = _with_pbuffer(batch, p_buffers, self._reg_channels(p_buffers))
            self._p_full = p_buffers
        self.models["dncnn"].zero_grad()
        self._manif_pre = None
        dn, lm = self.models["dncnn"], self.loss_funcs.get("l_manif")
        early_manif = (OVERLAP_MANIF and out_manif is not None and self.manif_learn and self.train_branches
                       and hasattr(lm, "forward_cropped") and callable(getattr(dn, "_branch", None))
                       and "while_branches_run" not in dn.__dict__)
        if early_manif:
            # The path-disentangling losses need the p-buffers and the SIZE of the KPCN output only: they are computed
            # on this stream while the two KPCN branches run on theirs (sbmc.KPCN.forward calls the hook between its
            # fork and its join), instead of after them with the machine idle.  Same calls in the same order (diffuse,
            # then specular -- the order in which the reference consumes its CPU generator, losses.py:35, :50); in the
            # backward pass their gradient kernels are then enqueued before the traversal has to wait for the branches.
            pre = self._manif_pre = {}
            targets = batch

            def hook(hw):
                like = torch.empty((0, 0) + tuple(hw), device="meta")
                for name in ("diffuse", "specular"):
                    pre[name] = lm.forward_cropped(out_manif[name], crop_like(targets["target_" + name], like))
            dn.__dict__["while_branches_run"] = hook
        try:
            out = self._regress_forward(batch)
        finally:
            if early_manif:
                dn.__dict__.pop("while_branches_run", None)
        try:
            return self._backward(batch, out, out_manif)
        finally:
            self._p_full = None
            self._manif_pre = None

    def _run_backward(self, roots):
        """One autograd traversal with both loss roots (same gradients as the reference's `L_diffuse.backward();
        L_specular.backward()`, :237-238).  Data parallel with an overlapping exchange (`grad_sync.early`): the
        traversal is cut at the p-buffers -- first the KPCN part (gradients of `dncnn` and of the two p-buffers),
        then the all-reduce of `dncnn`'s gradients is started asynchronously, then the path-embedding networks
        back-propagate while those 23.6 MB are on the wire."""
        early = getattr(self.grad_sync, "early", None)
        p_full = getattr(self, "_p_full", None)
        if early is None or p_full is None or getattr(self.grad_sync, "world", 1) == 1:
            torch.autograd.backward(roots)
            return
        params = [p for p in self.models["dncnn"].parameters() if p.requires_grad]
        p_outs = [p_full["diffuse"], p_full["specular"]]
        grads = torch.autograd.grad(roots, params + p_outs, allow_unused=True)
        for p, g in zip(params, grads):
            if g is not None:
                p.grad = g if p.grad is None else p.grad + g
        early({"dncnn": self.models["dncnn"]})
        pairs = [(o, g) for o, g in zip(p_outs, grads[len(params):]) if g is not None]
        if pairs:
            torch.autograd.backward([o for o, _ in pairs], [g for _, g in pairs])

    def _backward(self, batch, out, p_buffers):
        assert "radiance" in out and "diffuse" in out and "specular" in out
        total, diffuse, specular = out["radiance"], out["diffuse"], out["specular"]
        losses = {}
        tgt_total = crop_like(batch["target_total"], total)
        fused = self._fused_image_losses(batch, total, diffuse, specular) if self.train_branches else None
        if self.train_branches:
            branch_loss = {}
            for name, pred in (("diffuse", diffuse), ("specular", specular)):
                tgt = crop_like(batch["target_" + name], pred)
                loss = fused["l_" + name] if fused is not None else self.loss_funcs["l_" + name](pred, tgt)
                if self.manif_learn:
                    lm = self.loss_funcs["l_manif"]
                    pre = getattr(self, "_manif_pre", None)
                    if pre and name in pre and tuple(pred.shape[-2:]) == tuple(tgt.shape[-2:]):
                        l_manif = pre[name]                # computed while the KPCN branches ran (_forward_backward)
                    elif hasattr(lm, "forward_cropped"):   # the crop of crop_like() taken inside the loss Function
                        l_manif = lm.forward_cropped(p_buffers[name], tgt)
                    else:
                        l_manif = lm(crop_like(p_buffers[name], pred), tgt)
                    losses["l_manif_" + name] = l_manif.detach()
                    # the reference adds in place AFTER taking `.detach()` of the branch loss, so the
                    # logged l_diffuse / l_specular include the weighted manifold term (:221-232)
                    loss = loss + l_manif * self.w_manif
                losses["l_" + name] = loss.detach()
                branch_loss[name] = loss
            # the reference calls L_diffuse.backward(); L_specular.backward() (:237-238): the two graphs share no
            # parameter, so one traversal with both roots gives the same gradients and lets the branches' backward
            # kernels (recorded on different streams) overlap
            self._run_backward([branch_loss["diffuse"], branch_loss["specular"]])
            with torch.no_grad():
                losses["l_total"] = fused["l_total"] if fused is not None else \
                    self.loss_funcs["l_recon"](total, tgt_total).detach()
        else:
            l_total = self.loss_funcs["l_recon"](total, tgt_total)
            losses["l_total"] = l_total.detach()
            l_total.backward()
        with torch.no_grad():
            losses["rmse"] = fused["rmse"] if fused is not None else self.loss_funcs["l_test"](total, tgt_total).detach()
        # keep the reference's key order (l_diffuse, l_specular, l_manif_*, l_total, rmse)
        order = ["l_diffuse", "l_specular", "l_manif_diffuse", "l_manif_specular", "l_total", "rmse"]
        return {k: losses[k] for k in order if k in losses}

    def _fused_image_losses(self, batch, total, diffuse, specular):
        """K12: the three L1 terms and RelativeMSE of the step in one reduction launch (ops.ImageLossesFn) when the loss
        objects are the ones train_kpcn.py constructs (nn.L1Loss with mean reduction, :300-302; RelativeMSE) and the
        images live on the GPU; None otherwise (any other loss object is simply called, as in the reference)."""
        lf = self.loss_funcs
        from support.losses import RelativeMSE
        l1 = [lf.get(k) for k in ("l_diffuse", "l_specular", "l_recon")]
        if not all(type(f) is nn.L1Loss and f.reduction == "mean" for f in l1) or type(lf.get("l_test")) is not RelativeMSE:
            return None
        tgts = [batch["target_diffuse"], batch["target_specular"], batch["target_total"]]
        preds = [diffuse, specular, total]
        if not all(t.is_cuda and t.dtype == torch.float32 and t.dim() == 4 and t.shape[1] == 3 for t in tgts + preds):
            return None
        if any(t.stride(3) != 1 for t in tgts) or any(p.shape != preds[0].shape for p in preds):
            return None
        from wcmc_b200 import ops as wops
        l_d, l_s, l_t, rmse = wops.ImageLossesFn.apply(diffuse, specular, total.detach(), tgts[0], tgts[1], tgts[2],
                                                       float(lf["l_test"].eps))
        return {"l_diffuse": l_d, "l_specular": l_s, "l_total": l_t.detach(), "rmse": rmse.detach()}

    def _assert_finite(self, loss_dict):
        keys = list(loss_dict)
        vals = torch.stack([loss_dict[k].reshape(()) for k in keys])
        finite = torch.isfinite(vals)
        if not bool(finite.all()):  # the single host sync of the step
            bad = keys[int((~finite).nonzero()[0])]
            raise RuntimeError("%s: Non-finite loss at train time." % bad)

    def _fused(self):
        """FusedClipAdam over this interface's optimisers, or None when they are not plain Adam."""
        if not self.fused_optim:
            return None
        if self._fused_adam is None:
            from wcmc_b200 import optim as wopt
            opts = [self.optims["optim_" + name] for name in self.models]
            self._fused_adam = wopt.FusedClipAdam(opts) if all(wopt.supported(o) for o in opts) else False
        fa = self._fused_adam
        if fa and fa._dirty and not fa.consistent():
            # e.g. a resume that restored only one model's optimiser state (train_kpcn.py:283-296): the launch has
            # one bias-correction step count, so torch's own per-parameter step keeps running until they agree
            return None
        return fa or None

    def _accumulate(self, loss_dict):
        for k in loss_dict:
            if "m_" + k not in self.m_losses:
                self.m_losses["m_" + k] = torch.tensor(0.0, device=loss_dict[k].device)
        acc = [self.m_losses["m_" + k] for k in loss_dict]
        if acc and all(a.is_cuda for a in acc):      # one launch for the six running sums
            with torch.no_grad():
                torch._foreach_add_(acc, [loss_dict[k].detach().reshape(a.shape) for k, a in zip(loss_dict, acc)])
        else:
            for k in loss_dict:
                self.m_losses["m_" + k] += loss_dict[k]

    def _logging(self, loss_dict):
        self._assert_finite(loss_dict)
        if self.grad_sync is not None:
            self.grad_sync(self.models)
        if self._fused() is None:   # otherwise the clip is part of the fused optimiser kernel
            for model in self.models.values():
                nn.utils.clip_grad_value_(model.parameters(), clip_value=1.0)
        self._accumulate(loss_dict)

    def _optimization(self):
        fused = self._fused()
        if fused is not None:
            fused.step(clip=1.0)
            return
        for name in self.models:
            self.optims["optim_" + name].step()

    # ---- validation ----------------------------------------------------------------------------
    def validate_batch(self, batch):
        p_buffers = None
        if self.use_llpm_buf:
            p_buffers = self._manifold_forward(batch)
            assert p_buffers["diffuse"].shape[2] >= 2
            channels = self._reg_channels(p_buffers)
            batch = _with_pbuffer(batch, p_buffers, channels)
            if self.disentanglement_option in ("m10r01", "m11r01"):
                p_buffers = _half(p_buffers, lo=True)
        out = self._regress_forward(batch)
        tgt_total = crop_like(batch["target_total"], out["radiance"])
        l_total = self.loss_funcs["l_test"](out["radiance"], tgt_total)
        if self.m_losses["m_val"].device != l_total.device:
            self.m_losses["m_val"] = self.m_losses["m_val"].to(l_total.device)
        self.m_losses["m_val"] += l_total.detach()
        return out["radiance"], p_buffers

    def get_epoch_summary(self, mode, norm):
        if mode == "train":
            fa = self._fused_adam
            if fa and fa.nonfinite_count():
                print("[wcmc_b200] %d non-finite gradient elements were skipped by the fused optimiser so far "
                      "(16-bit backward overflow)" % fa.nonfinite_count())
            print("[][][]", end=" ")
            for key in self.m_losses:
                if key == "m_val":
                    continue
                print("%s: %.3fE-3" % (key, self.m_losses[key] / (norm * 2) * 1000), end="\t")
                self.m_losses[key] = torch.tensor(0.0, device=self.m_losses[key].device)
            print("")
            return -1.0
        return self.m_losses["m_val"].item() / (norm * 2)


_BATCH_KEYS = ("target_total", "target_diffuse", "target_specular", "kpcn_diffuse_buffer", "kpcn_specular_buffer",
               "kpcn_albedo")


class KPCNRefInterface(KPCNInterface):
    """Upper-bound ablation of /root/reference/support/interfaces.py:526-585: the KPCN inputs are concatenated
    with the reference images themselves (34 + 3 channels), no path embedding, no manifold loss.  Same kernels
    as KPCNInterface; only the batch wiring differs."""

    def __init__(self, models, optims, loss_funcs, args, visual=False, use_llpm_buf=False, manif_learn=False,
                 w_manif=0.1, train_branches=True):
        assert not use_llpm_buf
        assert not manif_learn
        super().__init__(models, optims, loss_funcs, args, visual, use_llpm_buf, manif_learn, w_manif, train_branches)

    def __str__(self):
        return "KPCNRefInterface"

    @staticmethod
    def _with_targets(batch):
        new = {k: batch[k] for k in _BATCH_KEYS}
        for name in ("diffuse", "specular"):
            new["kpcn_%s_in" % name] = torch.cat([batch["kpcn_%s_in" % name], batch["target_" + name]], 1)
        return new

    def train_batch(self, batch):
        batch = self._with_targets(batch)
        self.models["dncnn"].zero_grad()
        out = self._regress_forward(batch)
        loss_dict = self._backward(batch, out, None)
        self._logging(loss_dict)
        self._optimization()

    def validate_batch(self, batch):
        batch = self._with_targets(batch)
        out = self._regress_forward(batch)
        tgt_total = crop_like(batch["target_total"], out["radiance"])
        l_total = self.loss_funcs["l_test"](out["radiance"], tgt_total)
        if self.m_losses["m_val"].device != l_total.device:
            self.m_losses["m_val"] = self.m_losses["m_val"].to(l_total.device)
        self.m_losses["m_val"] += l_total.detach()
        return out["radiance"], None


class KPCNPreInterface(KPCNInterface):
    """Two-stage training of /root/reference/support/interfaces.py:588-750.
    manif_learn=True : pre-train the two path-embedding networks alone on the manifold loss (UNcropped p-buffers
                       and targets, :686-699); the KPCN is in eval mode and never runs.
    manif_learn=False: train the KPCN on [inputs | mean_S(p) | var_S(p)] with the embedding networks frozen in
                       eval mode: they run forward, nothing is propagated into their parameters' optimisers
                       (only `optim_dncnn` steps, only `dncnn` is clipped, :722-750)."""

    def __init__(self, models, optims, loss_funcs, args, visual=False, manif_learn=False, w_manif=0.1,
                 train_branches=True):
        super().__init__(models, optims, loss_funcs, args, visual, True, manif_learn, w_manif, train_branches)
        # the fused clip+Adam kernel covers every model of the interface; this one updates a subset
        self.fused_optim = False

    def __str__(self):
        return "KPCNPreInterface"

    def _trained(self, model_name):
        return ("backbone" in model_name) if self.manif_learn else ("dncnn" in model_name)

    def to_train_mode(self):
        for name, model in self.models.items():
            if "dncnn" in name or "backbone" in name:
                model.train(self._trained(name))
            assert "optim_" + name in self.optims, "`optim_%s`: an optimization algorithm is not defined." % name

    def train_batch(self, batch):
        self.models["backbone_diffuse"].zero_grad()
        self.models["backbone_specular"].zero_grad()
        if self.manif_learn:
            p_buffers = self._manifold_forward(batch)
            if self.iters % 1000 == 1:
                self._dump_pbuffers(p_buffers)
            loss_dict = self._backward(batch, None, p_buffers)
        else:
            self.models["dncnn"].zero_grad()
            p_buffers = self._manifold_forward(batch)
            batch = _with_pbuffer(batch, p_buffers)
            out = self._regress_forward(batch)
            loss_dict = self._backward(batch, out, None)
        self._logging(loss_dict)
        self._optimization()

    def _backward(self, batch, out, p_buffers):
        assert not out or ("radiance" in out and "diffuse" in out and "specular" in out)
        losses = {}
        if self.manif_learn:
            roots = []
            for name in ("diffuse", "specular"):
                l_manif = self.loss_funcs["l_manif"](p_buffers[name], batch["target_" + name]) * self.w_manif
                losses["l_manif_" + name] = l_manif.detach() / self.w_manif
                roots.append(l_manif)
            torch.autograd.backward(roots)   # the two networks share no parameter (:701-702)
            return losses
        total, diffuse, specular = out["radiance"], out["diffuse"], out["specular"]
        tgt_total = crop_like(batch["target_total"], total)
        if self.train_branches:
            roots = []
            for name, pred in (("diffuse", diffuse), ("specular", specular)):
                loss = self.loss_funcs["l_" + name](pred, crop_like(batch["target_" + name], pred))
                losses["l_" + name] = loss.detach()
                roots.append(loss)
            torch.autograd.backward(roots)
            with torch.no_grad():
                losses["l_total"] = self.loss_funcs["l_recon"](total, tgt_total).detach()
        else:
            l_total = self.loss_funcs["l_recon"](total, tgt_total)
            losses["l_total"] = l_total.detach()
            l_total.backward()
        return losses

    def _logging(self, loss_dict):
        self._assert_finite(loss_dict)
        if self.grad_sync is not None:
            self.grad_sync({k: m for k, m in self.models.items() if self._trained(k)})
        for name, model in self.models.items():
            if self._trained(name):
                nn.utils.clip_grad_value_(model.parameters(), clip_value=1.0)
        self._accumulate(loss_dict)

    def _optimization(self):
        for name in self.models:
            if self._trained(name):
                self.optims["optim_" + name].step()


class SBMCInterface(BaseInterface):
    """Step orchestration for a sample-based backbone (sbmc.Multisteps) with ONE path-embedding network
    (/root/reference/support/interfaces.py:336-523; used by train_sbmc.py).  Host logic only: the backbone is whatever
    module the caller passes (its kernels are outside this package's hot path); the path-embedding network, the
    path-disentangling loss and RelativeMSE are this package's, so a `models['backbone'] = PathNet(...)` runs on the
    B200 kernels.  Same public surface and `m_losses` keys as the reference class.  Differences, host-side only:
    one finite check (one device -> host sync) per step instead of one per loss term, and the `grad_sync` hook of
    KPCNInterface between the backward pass and the gradient-norm clipping (:466)."""

    CLIP_NORM = 1000

    def __init__(self, models, optims, loss_funcs, args, visual=False, use_llpm_buf=False, manif_learn=False,
                 w_manif=0.1, use_sbmc_buf=True, disentangle="m11r11"):
        if manif_learn:
            assert "backbone" in models, "argument `models` dictionary should contain `'backbone'` key."
        assert "dncnn" in models, "argument `models` dictionary should contain `'dncnn'` key."
        if manif_learn:
            assert "l_manif" in loss_funcs
        assert "l_recon" in loss_funcs
        assert "l_test" in loss_funcs
        assert disentangle in _DISENTANGLE
        super().__init__(models, optims, loss_funcs, args, visual, use_llpm_buf, manif_learn, w_manif)
        self.disentangle = disentangle
        self.use_sbmc_buf = use_sbmc_buf
        self.grad_sync = None

    def __str__(self):
        return "SBMCInterface"

    def to_train_mode(self):
        for name, model in self.models.items():
            model.train()
            assert "optim_" + name in self.optims, "`optim_%s`: an optimization algorithm is not defined." % name

    def to_eval_mode(self):
        for model in self.models.values():
            model.eval()
        self.m_losses["m_val"] = torch.tensor(0.0)

    def preprocess(self, batch=None):
        for key in ("target_image", "radiance", "features"):
            assert key in batch
        if self.use_llpm_buf:
            assert "paths" in batch
        self.iters += 1

    def _manifold_forward(self, batch):
        return self.models["backbone"](batch)

    def _regress_forward(self, batch):
        return self.models["dncnn"](batch)

    def _with_pbuffer(self, batch, p_reg):
        """features <- [features | p | var_S(p).mean(C) / S broadcast over the samples] per sample (:385-394)."""
        s = p_reg.shape[1]
        var = (p_reg.var(1).mean(1, keepdim=True) / s).detach()
        var = var.unsqueeze(1).expand(-1, s, -1, -1, -1)
        return {"target_image": batch["target_image"], "radiance": batch["radiance"],
                "features": torch.cat([batch["features"], p_reg, var], 2)}

    def _split(self, p_buffer):
        """-> (p fed to the regression, p fed to the manifold loss)   (:371-383)"""
        c = p_buffer.shape[2]
        assert c >= 2
        lo, hi = p_buffer[:, :, :c // 2], p_buffer[:, :, c // 2:]
        return (lo if self.disentangle in ("m10r01", "m11r01") else p_buffer,
                hi if self.disentangle in ("m10r01", "m10r11") else p_buffer)

    def train_batch(self, batch, grad_hook_mode=False):
        out_manif = None
        if self.use_llpm_buf:
            self.models["backbone"].zero_grad()
            p_buffer = self._manifold_forward(batch)
            if self.iters % 1000 == 1:
                self._dump(p_buffer)
            p_reg, out_manif = self._split(p_buffer)
            batch = self._with_pbuffer(batch, p_reg)
        self.models["dncnn"].zero_grad()
        out = self._regress_forward(batch)
        loss_dict = self._backward(batch, out, out_manif)
        if grad_hook_mode:      # gradients only
            return
        self._logging(loss_dict)
        self._optimization()

    def _dump(self, p_buffer):
        out_dir = "../LLPM_results"
        if not os.path.isdir(out_dir):
            return
        try:
            import matplotlib.pyplot as plt
        except ImportError:
            return
        img = p_buffer.detach()[0, :, :3].mean(0).clamp(0.0, 1.0).permute(1, 2, 0).cpu().numpy()
        plt.imsave("%s/pbuf_%s.png" % (out_dir, self.args.model_name), img)

    def _backward(self, batch, out, p_buffer):
        losses = {}
        tgt = crop_like(batch["target_image"], out)
        l_total = self.loss_funcs["l_recon"](out, tgt)
        if self.manif_learn:
            l_manif = self.loss_funcs["l_manif"](crop_like(p_buffer, out), tgt)
            losses["l_manif"] = l_manif.detach()
            # the reference takes `.detach()` of the reconstruction loss and THEN adds the weighted manifold term in
            # place (:440-443): the detached view shares storage, so its logged `l_recon` includes that term too
            l_total = l_total + l_manif * self.w_manif
            losses["l_recon"] = l_total.detach()
        losses["l_total"] = l_total.detach()
        l_total.backward()
        with torch.no_grad():
            losses["rmse"] = self.loss_funcs["l_test"](out, tgt).detach()
        return losses

    def _logging(self, loss_dict):
        keys = list(loss_dict)
        finite = torch.isfinite(torch.stack([loss_dict[k].reshape(()) for k in keys]))
        if not bool(finite.all()):            # the step's one host sync; name the first offending term like :460
            bad = keys[int((~finite).nonzero()[0])]
            raise RuntimeError("%s: Non-finite loss at train time." % bad)
        if self.grad_sync is not None:
            self.grad_sync(self.models)
        for name, model in self.models.items():
            actual = nn.utils.clip_grad_norm_(model.parameters(), max_norm=self.CLIP_NORM)
            if actual > self.CLIP_NORM:
                print("Clipped %s gradients %f -> %f" % (name, self.CLIP_NORM, actual))
        for k in keys:
            if "m_" + k not in self.m_losses:
                self.m_losses["m_" + k] = torch.tensor(0.0, device=loss_dict[k].device)
            self.m_losses["m_" + k] += loss_dict[k]

    def _optimization(self):
        for name in self.models:
            self.optims["optim_" + name].step()

    def validate_batch(self, batch):
        p_buffer = None
        if self.use_llpm_buf:
            p_buffer = self._manifold_forward(batch)
            assert p_buffer.shape[2] >= 2
            if self.disentangle in ("m10r01", "m11r01"):
                p_buffer = p_buffer[:, :, :p_buffer.shape[2] // 2]
            batch = self._with_pbuffer(batch, p_buffer)
        out = self._regress_forward(batch)
        tgt = crop_like(batch["target_image"], out)
        l_total = self.loss_funcs["l_test"](out, tgt)
        if self.m_losses["m_val"].device != l_total.device:
            self.m_losses["m_val"] = self.m_losses["m_val"].to(l_total.device)
        self.m_losses["m_val"] += l_total.detach()
        return out, p_buffer

    def get_epoch_summary(self, mode, norm):
        if mode == "train":
            print("[][][]", end=" ")
            for key in self.m_losses:
                if key == "m_val":
                    continue
                print("%s: %.3fE-3" % (key, self.m_losses[key] / (norm * 2) * 1000), end="\t")
                self.m_losses[key] = torch.tensor(0.0, device=self.m_losses[key].device)
            print("")
            return -1.0
        return self.m_losses["m_val"].item() / (norm * 2)


class LBMCInterface(SBMCInterface):
    """The layer-based backbone's variant (/root/reference/support/interfaces.py:753-839; train_lbmc.py): the same step
    with a tighter gradient-norm clip (0.25 * 1000, :826) and a constructor without `visual` / `use_sbmc_buf`."""

    CLIP_NORM = 0.25 * 1000

    def __init__(self, models, optims, loss_funcs, args, use_llpm_buf=False, manif_learn=False, w_manif=0.1,
                 disentangle="m11r11"):
        super().__init__(models, optims, loss_funcs, args, False, use_llpm_buf, manif_learn, w_manif, False, disentangle)

    def __str__(self):
        return "LBMCInterface"
