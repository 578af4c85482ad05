"""CUDA-graph execution of the KPCN+WCMC train step.

One step of `KPCNInterface.train_batch` launches ~400 kernels from Python; at B=8 the GPU work is
~12 ms and the host needs longer than that to enqueue it.  The shapes of a training run are fixed
(batch, spp bucket, patch size), so the forward + two backward passes are captured once into a CUDA
graph and replayed: per step the host copies the batch into the graph's static input buffers,
replays, checks the finite flags (the step's single host sync, as in support/interfaces.py) and
runs gradient all-reduce / clipping / Adam eagerly.

Semantics are those of `KPCNInterface.train_batch` (/root/reference/support/interfaces.py:122-192)
with two restrictions: the path-disentangling loss must draw its pairing permutations on the device
(`FeatureMSE(rng="device")`: a CPU `randperm` would be frozen into the graph), and the every-1000-
iterations PNG dump of the p-buffers is skipped.
"""
import torch

from support import losses as _losses


class GraphedTrainStep:
    def __init__(self, itf, example_batch, warmup=3):
        self.itf = itf
        lm = itf.loss_funcs.get("l_manif")
        if itf.manif_learn and getattr(lm, "rng", "cpu") != "device":
            raise ValueError("GraphedTrainStep needs FeatureMSE(rng='device'): a CPU permutation would be "
                             "captured once and replayed forever")
        self.static = {k: torch.empty_like(v, device="cuda") for k, v in example_batch.items()}
        for k, v in example_batch.items():
            self.static[k].copy_(v)
        self.graph = None
        self.loss = None
        self.flags = None
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):   # eager warm-up on a side stream (allocator / lazy init)
                self._fwd_bwd()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self._capture()

    def _fwd_bwd(self):
        itf = self.itf
        _losses.DEFER_FINITE_CHECK = True
        _losses.FINITE_FLAGS.clear()
        try:
            batch = self.static
            out_manif = None
            if itf.use_llpm_buf:
                itf.models["backbone_diffuse"].zero_grad()
                itf.models["backbone_specular"].zero_grad()
                p_buffers = itf._manifold_forward(batch)
                p_reg, out_manif = itf._split(p_buffers)
                from support.interfaces import _with_pbuffer
                batch = _with_pbuffer(batch, p_reg)
            itf.models["dncnn"].zero_grad()
            out = itf._regress_forward(batch)
            loss = itf._backward(batch, out, out_manif)
            flags = list(_losses.FINITE_FLAGS)
        finally:
            _losses.DEFER_FINITE_CHECK = False
            _losses.FINITE_FLAGS.clear()
        return loss, flags

    def _capture(self):
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            loss, flags = self._fwd_bwd()
            self.loss = loss
            self.flags = torch.stack(flags).all() if flags else None

    def __call__(self, batch):
        itf = self.itf
        itf.preprocess(batch)
        for k, v in batch.items():
            if k in self.static:
                self.static[k].copy_(v, non_blocking=True)
        self.graph.replay()
        if self.flags is not None and not bool(self.flags):
            raise RuntimeError("Infinite loss at train time.")
        itf._logging(self.loss)       # finite check of the losses, grad sync, clip, m_losses
        itf._optimization()
        return self.loss
