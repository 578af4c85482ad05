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
import contextlib
import os
import warnings

import torch
import torch.distributed as dist


_DRY = os.environ.get("WCMC_EXCHANGE_DRY", "0") == "1"


class PeerExchange:
    """A symmetric fp32 buffer (same size on every rank of `group`, every rank's copy mapped into every other rank's
    address space, one multicast mapping where the switch offers it) plus the in-place all-reduce kernel over it
    (`wcmc_grad_exchange`, csrc/grad_exchange.cu).  torch's symmetric-memory allocator provides the mappings; the
    data path is this package's own kernel.  Construction is a collective."""

    def __init__(self, n_floats, group=None, device=None, multicast=True):
        import torch.distributed._symmetric_memory as symm
        from . import lib
        group = group if group is not None else dist.group.WORLD
        device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.capacity = (int(n_floats) + 3) // 4 * 4
        self.flag_offset = self.capacity * 4
        self.buf = symm.empty(self.capacity + lib.grad_exchange_flag_bytes() // 4, dtype=torch.float32, device=device)
        self.buf.zero_()                      # the flags start at zero on every rank ...
        torch.cuda.synchronize(device)
        self.hdl = symm.rendezvous(self.buf, group)
        dist.barrier(group)                   # ... before any rank can launch
        self.rank, self.world = self.hdl.rank, self.hdl.world_size
        self.multicast_ptr = int(self.hdl.multicast_ptr) if multicast else 0
        self.peers_dev = int(self.hdl.buffer_ptrs_dev)
        self._lib = lib

    @property
    def transport(self):
        return "multimem" if self.multicast_ptr else "peer"

    def all_reduce_(self, n, scale=1.0, channel=0, offset=0):
        """buf[offset:offset+n] <- scale * sum over ranks, on the current stream (n is rounded up to 4 floats)."""
        n4 = (int(n) + 3) // 4 * 4
        assert offset % 4 == 0 and offset + n4 <= self.capacity
        if _DRY:       # measurement aid (WCMC_EXCHANGE_DRY=1): everything but the exchange itself, results are wrong
            return self.buf[offset:offset + n]
        self._lib.grad_exchange(self.buf, self.multicast_ptr, self.peers_dev, self.flag_offset, offset, n4, self.rank,
                                self.world, channel, scale)
        return self.buf[offset:offset + n]


class GradAllReduce:
    """Callable(models) for KPCNInterface.grad_sync.  Works with any initialised process group.

    Transport (`transport=` or WCMC_EXCHANGE; default "auto"): on CUDA the gradients travel through this package's own
    NVSwitch kernel over symmetric peer memory (`PeerExchange`: "multimem" = in-switch reduction + multicast, "peer" =
    peer loads / stores in rank order), falling back to "nccl" (`dist.all_reduce`) when the symmetric allocation is
    not available on the machine; gloo groups (the CPU tests) always take `dist.all_reduce`.

    Overlap: `early(models)` may be called DURING the backward pass for models whose gradients are already complete
    (the interface calls it for `dncnn` once the KPCN part of the traversal is done, while the two path-embedding
    networks still back-propagate): it starts the exchange of those gradients on a side stream.  `__call__` then only
    reduces what is left and waits for the early part -- under a CUDA-graph capture the whole pattern becomes a
    fork / join inside the graph.  Every rank issues the same exchanges in the same order."""

    def __init__(self, group=None, overlap=True, transport=None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.bytes_last = 0
        self._early = []           # (work | None, flat buffer, gradient tensors) of reductions in flight
        self.transport = (transport or os.environ.get("WCMC_EXCHANGE", "auto")).lower()
        assert self.transport in ("auto", "nccl", "peer", "multimem")
        self._px = {}              # channel -> PeerExchange
        self._side = None
        if not overlap:
            self.early = None      # the interface then keeps the single traversal + one reduction

    # ---- transport ------------------------------------------------------------------------------------------------
    def _use_peer(self, grads):
        return self.transport != "nccl" and self.world > 1 and bool(grads) and grads[0].is_cuda

    def _exchange(self, channel, n):
        """The symmetric buffer of `channel`, large enough for n floats, or None (-> dist.all_reduce)."""
        px = self._px.get(channel)
        if px is not None and px.capacity >= n:
            return px
        if px is None and channel in self._px:
            return None                              # tried before, not available
        if torch.cuda.is_current_stream_capturing():
            raise RuntimeError("GradAllReduce: the symmetric exchange buffer must exist before a CUDA-graph capture "
                               "(call prepare(models) first)")
        err = None
        try:
            px = PeerExchange(n, self.group, multicast=self.transport != "peer")
            if self.transport == "multimem" and not px.multicast_ptr:
                raise RuntimeError("no multicast mapping on this machine")
        except Exception as e:                       # noqa: BLE001 -- whatever the allocator raises
            px, err = None, e
        agree = torch.tensor([0 if px is None else 1], dtype=torch.int32, device="cuda")
        dist.all_reduce(agree, op=dist.ReduceOp.MIN, group=self.group)
        if int(agree) == 0:
            if self.transport != "auto":
                raise RuntimeError(f"GradAllReduce(transport={self.transport!r}): symmetric memory unavailable: {err}")
            warnings.warn(f"GradAllReduce: symmetric peer memory unavailable ({err}); using NCCL all_reduce")
            px = None
        if px is not None and self.transport == "auto" and px.multicast_ptr:
            self._calibrate(px, n)
        self._px[channel] = px
        return px

    def _calibrate(self, px, n):
        """transport="auto" on a machine with a multicast mapping: time both variants of the kernel on this message
        size (a collective: every rank runs the same launches) and keep the multicast one unless peer loads are
        clearly (25 %) faster -- the in-switch reduction wins with many ranks, plain peer loads with two."""
        mc, ms = px.multicast_ptr, []
        for cand in (mc, 0):
            px.multicast_ptr = cand
            for _ in range(3):
                px.all_reduce_(n, 1.0)
            torch.cuda.synchronize()
            dist.barrier(self.group)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                px.all_reduce_(n, 1.0)
            e1.record()
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1) / 10)
        t = torch.tensor(ms, dtype=torch.float32, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
        # ties go to the multicast variant: its CTAs (64 x 256 threads x 34 registers) fit next to ANY of the step's
        # persistent kernels, the peer variant's (128 x 256 x <= 58) do not fit next to the 544-thread path-network
        # backward kernels and hold those off their SMs while an exchange is in flight (4 GPUs, in the step:
        # 7.147 ms multimem against 7.204 ms peer although peer is 2 % faster alone)
        px.multicast_ptr = mc if float(t[0]) <= 1.25 * float(t[1]) else 0
        px.calibration_us = {"multimem": round(float(t[0]) * 1e3, 1), "peer": round(float(t[1]) * 1e3, 1)}

    def prepare(self, models):
        """Allocates the exchange buffers for all of `models`' parameters (a collective; needed before a CUDA-graph
        capture, otherwise done on first use)."""
        params = [p for m in models.values() for p in m.parameters()]
        if not self._use_peer(params):
            return
        n = self._slots(params) + 4
        for channel in (0, 1) if self.early is not None else (1,):
            self._exchange(channel, n)

    def peer_transport(self):
        px = [p for p in self._px.values() if p is not None]
        return px[-1].transport if px else "nccl"

    def describe(self):
        """{channel: transport [+ calibration timings]} for logs / bench.py's config."""
        return {("early" if c == 0 else "late"): ("nccl" if p is None else
                                                  dict(transport=p.transport, **getattr(p, "calibration_us", {})))
                for c, p in sorted(self._px.items())}

    def _flatten(self, grads):
        return torch.cat([g.reshape(-1) for g in grads])

    def early(self, models):
        if self.world == 1:
            return
        grads = [p.grad for m in models.values() for p in m.parameters() if p.grad is not None]
        if not grads:
            return
        px = self._exchange(0, self._slots(grads)) if self._use_peer(grads) else None
        if px is None:
            flat = self._flatten(grads)
            work = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
            self._early.append((work, flat, grads))
            return
        with self._aside(grads[0]):
            views, n = self._gather(px, grads)
            px.all_reduce_(n, 1.0 / self.world, channel=0)
        self._early.append((None, views, [p for m in models.values() for p in m.parameters() if p.grad is not None]))

    @contextlib.contextmanager
    def _aside(self, ref):
        """The early exchange's own stream (forked from the current one); nothing to fork for CPU tensors."""
        if not ref.is_cuda:
            yield
            return
        if self._side is None:
            self._side = torch.cuda.Stream()
        self._side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self._side):
            yield

    def _rejoin(self):
        if self._side is not None:
            torch.cuda.current_stream().wait_stream(self._side)

    @staticmethod
    def _gather(px, tensors):
        """The gradients (and the flag) into the symmetric buffer, every tensor on a 16-byte boundary -> (views, n).
        Nothing is copied when they already live there (p.grad still holds last step's views and the backward pass
        accumulated in place: `zero_grad(set_to_none=False)`)."""
        views, at = [], 0
        for t in tensors:
            k = t.numel()
            views.append(px.buf[at:at + k].view_as(t))
            at += (k + 3) // 4 * 4
        if any(v.data_ptr() != t.data_ptr() for v, t in zip(views, tensors)):
            torch._foreach_copy_(views, list(tensors))
        return views, at

    @staticmethod
    def _slots(tensors):
        return sum((t.numel() + 3) // 4 * 4 for t in tensors)

    def drain(self):
        """Waits for and forgets reductions started by `early` whose step never reached `__call__` (gradient-only
        passes: `train_batch(grad_hook_mode=True)`, the eager warm-up runs before a CUDA-graph capture)."""
        early, self._early = self._early, []
        for work, _, _ in early:
            if work is not None:
                work.wait()
            else:
                self._rejoin()

    def _finish(self, flat, grads):
        flat.div_(self.world)
        torch._foreach_copy_(grads, [t.view_as(g) for t, g in zip(flat.split([g.numel() for g in grads]), grads)])
        return flat.numel() * flat.element_size()

    @staticmethod
    def _adopt(views, params):
        """Peer transport: the averaged gradients stay where the exchange left them -- every `p.grad` becomes a view
        of the symmetric buffer (no copy back; clip + Adam read them there, the next backward pass replaces them)."""
        for p, v in zip(params, views):
            p.grad = v
        return sum(v.numel() for v in views) * 4

    def __call__(self, models, ok=None):
        """Reduces what `early` has not taken yet and waits for the early part.  `ok` (optional 0-d bool tensor: this
        rank's losses are finite): its negation rides along as one more element of the last gradient buffer, so
        "every rank takes or skips the update together" costs no collective of its own; returns the combined flag."""
        if self.world == 1:
            return ok
        early, self._early = self._early, []
        done = {id(g) for _, _, gs in early for g in gs}      # parameters (peer transport) or gradients (nccl)
        done |= {id(p) for m in models.values() for p in m.parameters() if p.grad is not None and id(p.grad) in done}
        params = [p for m in models.values() for p in m.parameters() if p.grad is not None and id(p) not in done]
        grads = [p.grad for p in params]
        nbytes = 0
        all_ok = None
        if grads or ok is not None:
            flag = [] if ok is None else [(~ok).to(torch.float32).reshape(1)]
            px = self._exchange(1, self._slots(grads + flag)) if self._use_peer(grads + flag) else None
            if px is None:
                flat = torch.cat([g.reshape(-1) for g in grads] + flag)
                dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
                if ok is not None:
                    all_ok = flat[-1] == 0
                    flat = flat[:-1]
                if grads:
                    nbytes += self._finish(flat, grads)
            else:
                views, n = self._gather(px, grads + flag)
                px.all_reduce_(n, 1.0 / self.world, channel=1)
                if ok is not None:
                    all_ok = views.pop()[0] == 0
                nbytes += self._adopt(views, params)
        for work, flat_e, grads_e in early:
            if work is not None:
                work.wait()                       # the current stream waits for the collective
            else:
                self._rejoin()
            nbytes += self._adopt(flat_e, grads_e) if work is None else self._finish(flat_e, grads_e)
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
