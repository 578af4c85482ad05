"""Fused gradient clipping + Adam over the optimisers the reference constructs.

`train_kpcn.py:277` builds one plain `torch.optim.Adam` per model and the step runs
`clip_grad_value_(…, 1.0)` then `optim.step()` three times (`support/interfaces.py:261, :269-271`).
`FusedClipAdam` performs exactly that update with ONE kernel launch (wcmc_adam_clip_step) directly on the
optimisers' own state tensors (`exp_avg`, `exp_avg_sq`, `step`), so `state_dict()` / checkpoints stay
interchangeable with torch's.  Anything but default Adam (amsgrad, weight decay, maximize, non-fp32,
non-CUDA) is reported as unsupported and the caller keeps using torch's step.
"""
import ctypes

import numpy as np
import torch

from . import lib


def supported(opt):
    if type(opt) is not torch.optim.Adam:
        return False
    for g in opt.param_groups:
        if g.get("amsgrad") or g.get("weight_decay", 0) != 0 or g.get("maximize") or g.get("differentiable"):
            return False
        if isinstance(g["lr"], torch.Tensor):
            return False
        for p in g["params"]:
            if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous():
                return False
    return True


class FusedClipAdam:
    """State contract: torch's own per-parameter `state['step' | 'exp_avg' | 'exp_avg_sq']` tensors ARE the state.
    `step` is kept current after every update (one `_foreach_add_` over the CPU scalars), so anything that reads
    the optimisers -- `state_dict()`, or pickling the optimiser objects as /root/reference/train_kpcn.py:114,143
    does with `'optims': itf.optims` (`Optimizer.__getstate__` runs no hook) -- sees what torch's Adam would have
    written.  `load_state_dict()` / an lr change between steps is picked up by the next step (also under a captured
    CUDA graph: the descriptors are re-read from pinned host memory on every replay)."""

    def __init__(self, optimizers):
        self.optims = list(optimizers)
        assert all(supported(o) for o in self.optims)
        self.device = self.optims[0].param_groups[0]["params"][0].device
        self.t = None          # steps taken (python int); device copy in self.t_dev
        self.t_dev = None
        self._key = None
        self._dev_tensors = self._dev_blocks = self._host_keep = None
        self._nblocks = 0
        self._nbytes = 0.0
        self.nonfinite = torch.zeros(1, dtype=torch.int64, device=self.device)   # skipped NaN / inf gradient elements
        self._steps = []       # the CPU `step` scalars of every parameter that takes part in the update
        self._hyper = None
        self._dirty = True     # state tensors may have been replaced (load_state_dict): re-read step counts
        for o in self.optims:
            o.register_load_state_dict_post_hook(lambda opt, self=self: self._mark_dirty())

    def _mark_dirty(self):
        self._dirty = True

    @staticmethod
    def step_counts(optimizers):
        """Set of step counts over every parameter of `optimizers` (a parameter without state counts as 0)."""
        out = set()
        for o in optimizers:
            for g in o.param_groups:
                for p in g["params"]:
                    st = o.state.get(p, {})
                    out.add(int(float(st["step"])) if "step" in st else 0)
        return out

    def consistent(self):
        """One bias-correction step count for the whole launch: every parameter must be at the same step (not the
        case e.g. after resuming with only one model's optimiser state; the caller then keeps torch's step)."""
        return len(self.step_counts(self.optims)) <= 1

    def prepare(self):
        """Creates the state of EVERY parameter and the device step counter now (must happen before a CUDA
        graph capture: allocations / fills inside the capture would be replayed)."""
        for o in self.optims:
            for g in o.param_groups:
                for p in g["params"]:
                    self._state(o, p)
        self._sync_t()

    def _sync_t(self):
        steps = self.step_counts(self.optims)
        if len(steps) != 1:
            raise RuntimeError("FusedClipAdam: parameters with different step counts (%s); use torch's step" % steps)
        t = steps.pop()
        if self.t_dev is None:
            self.t_dev = torch.full((1,), t, dtype=torch.int32, device=self.device)
        elif t != self.t:
            self.t_dev.fill_(t)
        self.t = t
        self._dirty = False

    @staticmethod
    def _state(o, p):
        st = o.state[p]
        if len(st) == 0:
            st["step"] = torch.tensor(0.0, dtype=torch.float32)
            st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
        return st

    # ---- torch-compatible lazy state ---------------------------------------------------------------
    def _entries(self):
        out = []
        for o in self.optims:
            for g in o.param_groups:
                b1, b2 = g["betas"]
                for p in g["params"]:
                    if p.grad is None:
                        continue
                    st = self._state(o, p)
                    assert p.grad.is_contiguous() and p.grad.dtype == torch.float32
                    out.append((p, p.grad, st, float(g["lr"]), float(b1), float(b2), float(g["eps"])))
        return out

    @staticmethod
    def _key_of(entries):
        return tuple((e[0].data_ptr(), e[1].data_ptr(), e[2]["exp_avg"].data_ptr(), e[2]["exp_avg_sq"].data_ptr(),
                      id(e[2]["step"]), e[3], e[4], e[5], e[6]) for e in entries)

    def _build(self, entries):
        n = len(entries)
        descs = (lib.AdamTensor * n)()
        chunk = lib.load().wcmc_adam_chunk()
        blocks = []
        total = 0
        for i, (p, g, st, lr, b1, b2, eps) in enumerate(entries):
            descs[i] = lib.AdamTensor(p.data_ptr(), g.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(),
                                      p.numel(), lr, b1, b2, eps)
            nb = (p.numel() + chunk - 1) // chunk
            blocks.append(np.stack([np.full(nb, i, dtype=np.int32), np.arange(nb, dtype=np.int32)], 1))
            total += p.numel()
        raw = torch.from_numpy(np.frombuffer(bytes(descs), dtype=np.uint8).copy())
        blk = torch.from_numpy(np.concatenate(blocks, 0))
        keep = self._host_keep
        if keep is not None and keep[0].numel() == raw.numel() and keep[1].shape == blk.shape:
            # same launch shape: rewrite the pinned descriptors IN PLACE (a captured graph copies from exactly these
            # buffers on every replay); nothing may still be reading them
            torch.cuda.synchronize(self.device)
            keep[0].copy_(raw)
            keep[1].copy_(blk)
        else:
            keep = self._host_keep = (raw.pin_memory(), blk.pin_memory())
            self._dev_tensors = torch.empty(raw.numel(), dtype=torch.uint8, device=self.device)
            self._dev_blocks = torch.empty(blk.shape, dtype=torch.int32, device=self.device)
        self._dev_tensors.copy_(keep[0], non_blocking=True)
        self._dev_blocks.copy_(keep[1], non_blocking=True)
        self._nblocks = int(blk.shape[0])
        self._nbytes = total * 4.0 * 8
        self._steps = [e[2]["step"] for e in entries]
        self._hyper = self._hyper_now()

    def _hyper_now(self):
        return tuple((g["lr"], tuple(g["betas"]), g["eps"]) for o in self.optims for g in o.param_groups)

    def refresh_if_changed(self):
        """The per-replay check of the CUDA-graph step: a few comparisons unless `load_state_dict()` ran or a
        hyper-parameter (lr schedule) changed, in which case the descriptors the graph re-reads are rewritten."""
        if self._dirty or self._hyper_now() != self._hyper:
            self.refresh()

    def refresh(self):
        """Re-reads hyper-parameters / state tensors / step counts if they changed since the last launch.  Returns
        the entry list (None when no parameter has a gradient).  Cheap when nothing changed."""
        entries = self._entries()
        if not entries:
            return None
        if self.t is None or self._dirty:
            self._sync_t()
        key = self._key_of(entries)
        if key != self._key:
            self._build(entries)
            self._key = key
        return entries

    # ---- the step ------------------------------------------------------------------------------------
    def step(self, clip=1.0, ok_flag=None, count=True):
        """clip_grad_value_(clip) + Adam on every parameter that has a gradient.  ok_flag: optional
        int32 device tensor; 0 = skip the update (non-finite loss).  count=False: the caller (CUDA graph
        replay) advances the host-side step count itself with note_step()."""
        if self.refresh() is None:
            return
        lib.adam_clip_step(self._dev_tensors, self._dev_blocks, self._nblocks, self.t_dev, ok_flag, clip, self._nbytes,
                           self.nonfinite)
        if count:
            self.note_step()

    def note_step(self):
        """One update has been applied: torch's per-parameter `step` scalars follow."""
        self.t += 1
        if self._steps:
            torch._foreach_add_(self._steps, 1.0)

    def set_step(self, t):
        """Puts the update count of every parameter back to `t` (torch's `step` scalars and the device counter the
        kernel's bias correction reads) -- for callers that restore weights and moments to an earlier point."""
        t = int(t)
        for o in self.optims:
            for st in o.state.values():
                if "step" in st:
                    st["step"].fill_(float(t))
        if self.t_dev is not None:
            self.t_dev.fill_(t)
        self.t = t

    def nonfinite_count(self):
        """Gradient elements skipped so far because they were NaN / inf (host sync)."""
        return int(self.nonfinite.item())

    def sync_state(self):
        """Kept for callers of the round-1 API: the step counts are always current now."""
