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
        self._stale = False
        for o in self.optims:
            o.register_state_dict_pre_hook(lambda opt, self=self: self.sync_state())

    def prepare(self):
        """Creates the state of EVERY parameter and the device step counter now (must happen before a CUDA
        graph capture: allocations / fills inside the capture would be replayed)."""
        for o in self.optims:
            for g in o.param_groups:
                for p in g["params"]:
                    self._state(o, p)
        if self.t is None:
            steps = {int(float(st["step"])) for o in self.optims for st in o.state.values()}
            if len(steps) != 1:
                raise RuntimeError("FusedClipAdam: parameters with different step counts (%s)" % steps)
            self.t = steps.pop()
            self.t_dev = torch.full((1,), self.t, dtype=torch.int32, device=self.device)

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

    def _init_step(self, entries):
        steps = {int(float(e[2]["step"])) for e in entries}
        if len(steps) != 1:
            raise RuntimeError("FusedClipAdam: parameters with different step counts (%s); use torch's step" % steps)
        self.t = steps.pop()
        self.t_dev = torch.full((1,), self.t, dtype=torch.int32, device=self.device)

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
        raw = np.frombuffer(bytes(descs), dtype=np.uint8).copy()
        blk = np.concatenate(blocks, 0)
        host_t = torch.from_numpy(raw).pin_memory()
        host_b = torch.from_numpy(blk).pin_memory()
        if self._dev_tensors is None or self._dev_tensors.numel() != host_t.numel():
            self._dev_tensors = torch.empty(host_t.numel(), dtype=torch.uint8, device=self.device)
        if self._dev_blocks is None or self._dev_blocks.shape != host_b.shape:
            self._dev_blocks = torch.empty(host_b.shape, dtype=torch.int32, device=self.device)
        self._dev_tensors.copy_(host_t, non_blocking=True)
        self._dev_blocks.copy_(host_b, non_blocking=True)
        self._host_keep = (host_t, host_b)    # a captured CUDA graph re-copies from these on every replay
        self._nblocks = int(blk.shape[0])
        self._nbytes = total * 4.0 * 8

    # ---- the step ------------------------------------------------------------------------------------
    def step(self, clip=1.0, ok_flag=None, count=True):
        """clip_grad_value_(clip) + Adam on every parameter that has a gradient.  ok_flag: optional
        int32 device tensor; 0 = skip the update (non-finite loss).  count=False: the caller (CUDA graph
        replay) advances the host-side step count itself with note_step()."""
        entries = self._entries()
        if not entries:
            return
        if self.t is None:
            self._init_step(entries)
        key = tuple((e[0].data_ptr(), e[1].data_ptr(), e[3]) for e in entries)
        if key != self._key:
            self._build(entries)
            self._key = key
        lib.adam_clip_step(self._dev_tensors, self._dev_blocks, self._nblocks, self.t_dev, ok_flag, clip, self._nbytes)
        if count:
            self.note_step()

    def note_step(self):
        self.t += 1
        self._stale = True

    def sync_state(self):
        """Writes the step count back into torch's per-parameter `state['step']` (before state_dict())."""
        if not self._stale:
            return
        for o in self.optims:
            for st in o.state.values():
                if "step" in st:
                    st["step"].fill_(float(self.t))
        self._stale = False
