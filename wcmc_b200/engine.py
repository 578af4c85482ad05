"""CUDA-graph execution of the KPCN+WCMC train step.

One step of `KPCNInterface.train_batch` launches ~400 kernels from Python; at B=8 the GPU work is
~12 ms and the host needs longer than that to enqueue it.  The shapes of a training run are fixed
(batch, spp bucket, patch size), so the forward + two backward passes are captured once into a CUDA
graph and replayed: per step the host copies the batch into the graph's static input buffers,
replays, checks the finite flags (the step's single host sync, as in support/interfaces.py) and
runs gradient all-reduce / clipping / Adam eagerly.

Semantics are those of `KPCNInterface.train_batch` (/root/reference/support/interfaces.py:122-192); the
every-1000-iterations PNG dump of the p-buffers is skipped.  Pairing permutations of the path-disentangling
loss: `FeatureMSE(rng="device")` draws them inside the graph; `rng="cpu"` keeps the reference's RNG contract
(CPU default generator, losses.py:35, :50) through pre-staged static index buffers that the host refills
before every replay (support.losses.PermStage) -- exact parity mode, ~27 ms of host randperm per step.
"""
import os

import torch

from support import losses as _losses

from . import ddp as _ddp


class GraphedTrainStep:
    def __init__(self, itf, example_batch, warmup=3):
        self.itf = itf
        lm = itf.loss_funcs.get("l_manif")
        self.stage = None
        if itf.manif_learn and getattr(lm, "rng", "cpu") != "device":
            if not hasattr(lm, "stage"):
                raise ValueError("GraphedTrainStep: the manifold loss draws CPU permutations and cannot stage them "
                                 "(use support.losses.FeatureMSE / GlobalRelativeSimilarityLoss)")
            self.stage = lm.stage = _losses.PermStage()
        self.static = {k: torch.empty_like(v, device="cuda") for k, v in example_batch.items()}
        for k, v in example_batch.items():
            self.static[k].copy_(v)
        self.graph = None
        self.loss = None
        self.flags = None
        if getattr(itf.grad_sync, "prepare", None) is not None:
            itf.grad_sync.prepare(itf.models)   # symmetric exchange buffers: a collective, not capturable
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):   # eager warm-up on a side stream (allocator / lazy init)
                if self.stage is not None:
                    self.stage.begin_step()
                self._fwd_bwd()
            if getattr(itf.grad_sync, "drain", None) is not None:
                itf.grad_sync.drain()      # the warm-up passes stop before _logging: nothing may stay in flight
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self._capture()

    def _fwd_bwd(self):
        itf = self.itf
        _losses.DEFER_FINITE_CHECK = True
        _losses.FINITE_FLAGS.clear()
        try:
            loss = itf._forward_backward(self.static)
            flags = list(_losses.FINITE_FLAGS)
        finally:
            _losses.DEFER_FINITE_CHECK = False
            _losses.FINITE_FLAGS.clear()
        return loss, flags

    def _capture(self):
        itf = self.itf
        # clip + Adam (one kernel, predicated on the finite flag) join the graph.  With data parallelism the
        # gradient exchange sits between backward and update (support/interfaces.py:237-238 -> :261 ->
        # :271) and is captured too (WCMC_GRAPH_ALLREDUCE=0 keeps all-reduce + optimiser outside the graph).
        self.sync_in_graph = itf.grad_sync is not None and os.environ.get("WCMC_GRAPH_ALLREDUCE", "1") != "0"
        self.fused = itf._fused() if (itf.grad_sync is None or self.sync_in_graph) else None
        if self.fused is not None:
            self.fused.prepare()
        if self.stage is not None:
            self.stage.begin_step()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            loss, flags = self._fwd_bwd()
            self.loss = loss
            ok = torch.isfinite(torch.stack([v.reshape(()) for v in loss.values()])).all()
            if flags:
                ok = ok & torch.stack(flags).all()
            self.flags = ok
            if self.fused is not None:
                if self.sync_in_graph:
                    # every rank takes or skips the update together: the flag rides in the gradient all-reduce
                    if isinstance(itf.grad_sync, _ddp.GradAllReduce):
                        ok = itf.grad_sync(itf.models, ok=ok)
                    else:                                 # a user-supplied callable(models) [+ all_ranks]
                        itf.grad_sync(itf.models)
                        ok = itf.grad_sync.all_ranks(ok)
                    self.flags = ok
                self.ok_i32 = ok.to(torch.int32).reshape(1)
                self.fused.step(clip=1.0, ok_flag=self.ok_i32, count=False)

    def release(self):
        """Drops the captured graph and its static buffers (call before tearing a process group down: the graph
        holds the NCCL kernels of the in-graph gradient all-reduce)."""
        torch.cuda.synchronize()
        if self.graph is not None:
            self.graph.reset()
        self.graph = None
        self.loss = self.flags = None

    def __call__(self, batch):
        itf = self.itf
        itf.preprocess(batch)
        for k, v in batch.items():
            if k in self.static:
                self.static[k].copy_(v, non_blocking=True)
        if self.fused is not None:
            self.fused.refresh_if_changed()   # lr schedule / load_state_dict since the last replay
        if self.stage is not None:
            self.stage.begin_step()           # this step's pairing permutations from the CPU generator
        self.graph.replay()
        if self.fused is not None:
            if not bool(self.flags):     # the single host sync of the step; the update was skipped on the device
                raise RuntimeError("Infinite loss at train time.")
            itf._accumulate(self.loss)
            self.fused.note_step()
            return self.loss
        if not bool(self.flags):
            raise RuntimeError("Infinite loss at train time.")
        itf._logging(self.loss)       # finite check of the losses, grad sync, clip, m_losses
        itf._optimization()
        return self.loss


class DevicePrefetcher:
    """Double-buffered host -> device staging of training batches on a copy stream, so the PCIe
    transfer of batch i+1 overlaps the compute of batch i (the reference's loop does a blocking
    `.cuda()` per tensor before every step, /root/reference/train_kpcn.py:47-50).

        pf = DevicePrefetcher(host_batches)      # iterable of dicts of (pinned) CPU tensors
        for dev_batch in pf: step(dev_batch)
    """

    def __init__(self, batches, device=None):
        self.it = iter(batches)
        self.device = device or torch.device("cuda", torch.cuda.current_device())
        self.copy_stream = torch.cuda.Stream(self.device)
        self.bufs = [None, None]
        self.ready = [torch.cuda.Event(), torch.cuda.Event()]
        self.free = [None, None]
        self.i = 0
        self._issue(0)

    def _issue(self, slot):
        try:
            host = next(self.it)
        except StopIteration:
            self.bufs[slot] = None
            return
        with torch.cuda.stream(self.copy_stream):
            if self.free[slot] is not None:
                self.copy_stream.wait_event(self.free[slot])   # the consumer of this slot has finished
            if self.bufs[slot] is None or any(self.bufs[slot][k].shape != v.shape for k, v in host.items()):
                self.bufs[slot] = {k: torch.empty(v.shape, dtype=v.dtype, device=self.device) for k, v in host.items()}
            for k, v in host.items():
                self.bufs[slot][k].copy_(v, non_blocking=True)
            self.ready[slot].record(self.copy_stream)

    def __iter__(self):
        return self

    def __next__(self):
        slot = self.i & 1
        if self.bufs[slot] is None:
            raise StopIteration
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(self.ready[slot])
        batch = self.bufs[slot]
        self.i += 1
        self._issue(self.i & 1)    # start the next transfer while this batch is being consumed
        return batch

    def release(self, batch_slot_index=None):
        """Call after the step that consumed the batch returned by the previous __next__ has been
        enqueued: marks that slot reusable once the current stream reaches this point."""
        slot = (self.i - 1) & 1
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        self.free[slot] = ev
