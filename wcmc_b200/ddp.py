"""Data-parallel gradient exchange for the KPCN+WCMC step: one process per GPU, NCCL all-reduce
of the flat fp32 gradients over NVLink / NVSwitch.

The reference's only multi-GPU mode is single-process nn.DataParallel
(/root/reference/train_kpcn.py:266-269), which re-broadcasts the weights every forward and reduces
gradients onto GPU 0.  Here each rank owns a full replica and its own shard of the batch; the one
exchange per step is `all_reduce(sum) / world` of ~11.7 M fp32 gradients (46.9 MB), placed where
the reference's step order requires it: after both backward passes, before clip_grad_value_ and
Adam (support/interfaces.py:237-238 -> :261 -> :271).  A stock DistributedDataParallel wrapper
does not fit: each of the two backward passes touches only half of `dncnn`'s parameters.
"""
import torch
import torch.distributed as dist


class GradAllReduce:
    """Callable(models) for KPCNInterface.grad_sync.  Works with any initialised process group
    (nccl on GPUs, gloo in the CPU tests).

    Overlap: `early(models)` may be called DURING the backward pass for models whose gradients are already complete
    (the interface calls it for `dncnn` once the KPCN part of the traversal is done, while the two path-embedding
    networks still back-propagate): it starts an asynchronous all-reduce of those gradients on the process group's
    own stream.  `__call__` then only reduces what is left and waits for the early part -- under a CUDA-graph
    capture the whole pattern becomes a fork / join inside the graph.  Every rank issues the same collectives in
    the same order."""

    def __init__(self, group=None, overlap=True):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.bytes_last = 0
        self._early = []           # (work, flat buffer, gradient tensors) of reductions in flight
        if not overlap:
            self.early = None      # the interface then keeps the single traversal + one reduction

    def _flatten(self, grads):
        return torch.cat([g.reshape(-1) for g in grads])

    def early(self, models):
        if self.world == 1:
            return
        grads = [p.grad for m in models.values() for p in m.parameters() if p.grad is not None]
        if not grads:
            return
        flat = self._flatten(grads)
        work = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        self._early.append((work, flat, grads))

    def drain(self):
        """Waits for and forgets reductions started by `early` whose step never reached `__call__` (gradient-only
        passes: `train_batch(grad_hook_mode=True)`, the eager warm-up runs before a CUDA-graph capture)."""
        early, self._early = self._early, []
        for work, _, _ in early:
            work.wait()

    def _finish(self, flat, grads):
        flat.div_(self.world)
        torch._foreach_copy_(grads, [t.view_as(g) for t, g in zip(flat.split([g.numel() for g in grads]), grads)])
        return flat.numel() * flat.element_size()

    def __call__(self, models, ok=None):
        """Reduces what `early` has not taken yet and waits for the early part.  `ok` (optional 0-d bool tensor: this
        rank's losses are finite): its negation rides along as one more element of the last gradient buffer, so
        "every rank takes or skips the update together" costs no collective of its own; returns the combined flag."""
        if self.world == 1:
            return ok
        early, self._early = self._early, []
        done = {id(g) for _, _, gs in early for g in gs}
        grads = [p.grad for m in models.values() for p in m.parameters() if p.grad is not None and id(p.grad) not in done]
        nbytes = 0
        all_ok = None
        if grads or ok is not None:
            parts = [g.reshape(-1) for g in grads]
            if ok is not None:
                parts.append((~ok).to(parts[0].dtype if parts else torch.float32).reshape(1))
            flat = torch.cat(parts)
            dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
            if ok is not None:
                all_ok = flat[-1] == 0
                flat = flat[:-1]
            if grads:
                nbytes += self._finish(flat, grads)
        for work, flat_e, grads_e in early:
            work.wait()                       # the current stream waits for the collective
            nbytes += self._finish(flat_e, grads_e)
        self.bytes_last = nbytes
        return all_ok

    def all_ranks(self, ok):
        """Logical AND of a 0-d bool tensor over the ranks (a non-finite loss on one rank must stop them all)."""
        if self.world == 1:
            return ok
        t = ok.to(torch.int32).reshape(1)
        dist.all_reduce(t, op=dist.ReduceOp.MIN, group=self.group)
        return t[0] > 0


def broadcast_parameters(models, src=0, group=None):
    """Replicas start from rank `src`'s weights."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    for m in models.values():
        for t in list(m.parameters()) + list(m.buffers()):
            dist.broadcast(t.data, src, group=group)
