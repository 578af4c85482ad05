"""CUDA-graph execution of the KPCN+WCMC train step.

One step of `KPCNInterface.train_batch` launches ~400 kernels from Python; at B=8 the GPU work is
~12 ms and the host needs longer than that to enqueue it.  The shapes of a training run are fixed
(batch, spp bucket, patch size), so the forward + two backward passes are captured once into a CUDA
graph and replayed: per step the host copies the batch into the graph's static input buffers (one launch; nothing
when the batch already lives there) and replays -- gradient exchange, clipping and Adam are part of the graph, and
the finite flag of a step is looked at one replay later, so the GPU never waits for the host (see GraphedTrainStep).

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
    """check="deferred" (default): the finite flag of replay i is copied to pinned host memory asynchronously and looked
    at after replay i+1 has been enqueued, so the GPU never waits for the host between steps (the per-step sync +
    Python + graph launch cost ~0.25 ms of idle GPU per 7.4 ms step).  A non-finite step still skips its own update ON
    THE DEVICE (the flag predicates clip + Adam inside the graph); the RuntimeError of
    /root/reference/support/interfaces.py:255-257 surfaces one call later, or in `finish()` / `release()`.
    check="immediate" keeps the reference's timing of the error (one host sync per step).

    A batch whose tensors ARE `self.static[...]` is used in place (no staging copy)."""

    def __init__(self, itf, example_batch, warmup=3, check="deferred"):
        assert check in ("deferred", "immediate")
        self.itf = itf
        self.check = check
        self._pending = None       # (event, slot, step index) of the replay whose flag has not been looked at yet
        self._flag_host = [torch.ones(1, dtype=torch.bool).pin_memory() for _ in range(2)]
        self._n = 0
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
        # the gradient tensors the captured kernels write and the captured Adam reads: an eager backward pass between
        # replays (validation with gradients, a per-kernel timing pass) re-points p.grad; __call__ puts these back so
        # that anything that re-reads the optimiser's view of the model (FusedClipAdam.refresh) sees the graph's own
        self._params = [p for m in itf.models.values() for p in m.parameters()]
        self._grads = [p.grad for p in self._params]
        self._adam_key = self.fused._key if self.fused is not None else None

    def _look(self, pending):
        ev, slot, n = pending
        ev.synchronize()
        if not bool(self._flag_host[slot]):
            raise RuntimeError("Infinite loss at train time. (graph replay %d; its update was skipped on the device)" % n)

    def finish(self):
        """Looks at the flag of the last replay (deferred checking keeps one outstanding)."""
        pending, self._pending = self._pending, None
        if pending is not None:
            self._look(pending)

    def release(self):
        """Drops the captured graph and its static buffers (call before tearing a process group down: the graph
        holds the kernels of the in-graph gradient exchange)."""
        torch.cuda.synchronize()
        try:
            self.finish()
        finally:
            if self.graph is not None:
                self.graph.reset()
            self.graph = None
            self.loss = self.flags = None
            self._params = self._grads = []

    def __call__(self, batch):
        itf = self.itf
        itf.preprocess(batch)
        dst, src = [], []
        for k, v in batch.items():
            if k in self.static and v is not self.static[k]:
                dst.append(self.static[k])
                src.append(v)
        if src and all(v.is_cuda and v.dtype == d.dtype for v, d in zip(src, dst)):
            torch._foreach_copy_(dst, src)            # one launch for the whole batch
        else:
            for d, v in zip(dst, src):
                d.copy_(v, non_blocking=True)
        for p, g in zip(self._params, self._grads):
            if p.grad is not g:
                p.grad = g
        if self.fused is not None:
            self.fused.refresh_if_changed()   # lr schedule / load_state_dict since the last replay
            if self.fused._key is not self._adam_key:
                # an EAGER optimiser step since the last replay rebuilt the descriptors (which the graph re-reads
                # from pinned memory on every replay) for that pass's gradient tensors: point them back at the graph's
                self.fused.refresh()
                self._adam_key = self.fused._key
        if self.stage is not None:
            self.stage.begin_step()           # this step's pairing permutations from the CPU generator
        self.graph.replay()
        n, self._n = self._n, self._n + 1
        if self.fused is not None:
            if self.check == "immediate":
                if not bool(self.flags):     # the single host sync of the step; the update was skipped on the device
                    raise RuntimeError("Infinite loss at train time.")
            else:
                slot = n & 1
                self._flag_host[slot].copy_(self.flags.reshape(1), non_blocking=True)
                ev = torch.cuda.Event()
                ev.record()
                pending, self._pending = self._pending, (ev, slot, n)
                if pending is not None:
                    self._look(pending)      # the PREVIOUS replay's flag: the GPU is already busy with this one
            itf._accumulate(self.loss)
            self.fused.note_step()
            return self.loss
        if not bool(self.flags):
            raise RuntimeError("Infinite loss at train time.")
        itf._logging(self.loss)       # finite check of the losses, grad sync, clip, m_losses
        itf._optimization()
        return self.loss


def bind_host_to_gpu(index):
    """Restricts this process to the CPU cores NVML reports as local to GPU `index` (torch's device index).  With one
    process per GPU on a two-socket host, pinned staging buffers are then first-touched on the GPU's own NUMA node and
    the per-step host -> device copies do not cross the socket interconnect (8 ranks x 197 MB per step).  Returns the
    core list, or None when NVML / the affinity call is not available (nothing changes then)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        uuid = str(torch.cuda.get_device_properties(index).uuid)
        try:
            h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid).encode())
        except Exception:                          # noqa: BLE001 -- older bindings take str
            h = pynvml.nvmlDeviceGetHandleByUUID("GPU-" + uuid)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = [64 * i + b for i, w in enumerate(mask) for b in range(64) if (int(w) >> b) & 1]
        cpus = sorted(set(cpus) & os.sched_getaffinity(0))
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return cpus
    except Exception:                              # noqa: BLE001 -- best effort, never fatal
        return None


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
