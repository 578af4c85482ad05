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
    (nccl on GPUs, gloo in the CPU tests)."""

    def __init__(self, group=None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.bytes_last = 0

    def __call__(self, models):
        if self.world == 1:
            return
        grads = [p.grad for m in models.values() for p in m.parameters() if p.grad is not None]
        if not grads:
            return
        flat = torch.cat([g.reshape(-1) for g in grads])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
        flat.div_(self.world)
        self.bytes_last = flat.numel() * flat.element_size()
        torch._foreach_copy_(grads, [t.view_as(g) for t, g in zip(flat.split([g.numel() for g in grads]), grads)])


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
