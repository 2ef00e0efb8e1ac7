"""CUDA-graph capture of one ConvNP training step (forward + NLL + backward [+ optimiser step]).

The step of BASELINE configs[1] is ~80 kernel launches for ~7 ms of GPU work; launched eagerly, the gaps between
kernels and the Python/ctypes launch path cost ~5 % of the step.  All shapes of a training step are static as long
as the batch signature (tasks per batch, context / target counts, internal grid) does not change -- which is what
the reference's grouping by station count guarantees (nzdownscale/downscaler/train.py:448-475) -- so the whole
step is captured once and replayed: inputs are copied into static device buffers, `graph.replay()` re-issues every
kernel (encoder, weight packing, tcgen05 convolutions, decoder, head, wgrad, and -- data-parallel -- the two bucketed NCCL
all-reduces).  Release the graph (``del``) and synchronise before ``destroy_process_group``.

Usage (what ``train_epoch(..., use_graph=True)`` and ``bench.py`` do):

    gs = GraphedTrainStep(model, opt, example_batch)        # example_batch: DeviceBatch
    loss = gs.step(batch)                                    # DeviceBatch / HostBatch with the same signature
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from .engine import DeviceBatch, DeviceContext, HostBatch


def batch_signature(b) -> tuple:
    """Everything that is baked into the captured kernel arguments."""
    ctx = []
    for c in b.contexts:
        xs = c.x if isinstance(c.x, tuple) else (c.x,)
        # the banded encoder's band width and workspace size are derived on the host from the gridded coordinates and
        # baked into the captured launch: the coordinates themselves are part of the signature
        coords = None
        if c.gridded and c.x_host is not None:
            coords = tuple(hash(np.ascontiguousarray(v).tobytes()) for v in c.x_host)
        ctx.append((c.gridded, tuple(tuple(v.shape) for v in xs), tuple(c.y.shape), None if c.mask is None else
                    tuple(c.mask.shape), c.mono, c.x_batched, c.y_batched, coords))
    xt = b.xt if isinstance(b.xt, tuple) else (b.xt,)
    g = b.grid
    return (tuple(ctx), tuple(tuple(v.shape) for v in xt), None if b.yt is None else tuple(b.yt.shape),
            None if b.aux_t is None else tuple(b.aux_t.shape), (g.start1, g.n1, g.start2, g.n2, g.res), b.B)


def _tensors(b):
    for c in b.contexts:
        for t in (c.x if isinstance(c.x, tuple) else (c.x,)) + (c.y, c.mask):
            yield t
    for t in (b.xt if isinstance(b.xt, tuple) else (b.xt,)) + (b.yt, b.aux_t):
        yield t


class GraphedTrainStep:
    def __init__(self, model, opt, example: DeviceBatch, capture_optimizer: Optional[bool] = None, warmup: int = 3,
                 warm: bool = False):
        """``warm=True``: an eager step with this batch signature has already run on this model (workspaces, side
        streams and host caches exist), so no warm-up steps -- and no extra optimiser updates -- are made here."""
        eng = model.engine
        # data-parallel: the two bucketed NCCL all-reduces of Engine.backward are captured with the step (they become
        # graph nodes on NCCL's stream, so the first still overlaps the down-path backward in every replay)
        self.model, self.opt, self.eng = model, opt, eng
        self.signature = batch_signature(example)
        clone = lambda t: None if t is None else t.clone()
        ctxs = []
        for c in example.contexts:
            x = tuple(clone(v) for v in c.x) if isinstance(c.x, tuple) else clone(c.x)
            ctxs.append(DeviceContext(c.gridded, x, clone(c.y), clone(c.mask), c.mono, c.x_batched, c.x_host, c.band_cache,
                                      c.y_batched))
        xt = tuple(clone(v) for v in example.xt) if isinstance(example.xt, tuple) else clone(example.xt)
        self.static = DeviceBatch(ctxs, xt, clone(example.yt), clone(example.aux_t), example.grid, example.B, example.Nt)
        if capture_optimizer is None:
            capture_optimizer = all(g.get("capturable", False) for g in opt.param_groups)
        self.capture_optimizer = capture_optimizer
        # warm-up on a side stream (allocates every workspace, creates the side streams, settles the optimiser state)
        if not warm:
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                for _ in range(max(1, warmup)):   # at least one: workspaces, side streams, host-side caches must exist
                    self._eager()
            torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        self.opt.zero_grad(set_to_none=True)
        # the bf16 weight packing must be IN the graph: an eager forward just before the capture (validation loss_fn)
        # leaves the packed copies current, and a capture that then records no pack kernels would replay with the
        # weights frozen at capture time while the fp32 masters keep moving
        eng.force_pack = True
        l0 = eng.launches
        try:
            # thread-local error mode: train_epoch's worker thread may be staging / uploading the next batch (pinned
            # allocations, event synchronisation) while this thread captures
            with torch.cuda.graph(self.graph, capture_error_mode="thread_local"):
                self.loss = self._fwd_bwd()
                if self.capture_optimizer:
                    self.opt.step()
        finally:
            eng.force_pack = False
        self.launches = eng.launches - l0
        eng.pin_signature()        # the graph replays from the engine's workspaces of this batch shape
        if eng.precision == "bf16" and not eng.packs_recorded:
            raise RuntimeError("graph capture recorded no weight-packing kernels")
        # the captured backward writes into THESE gradient tensors (they live in the graph's private pool); an eager
        # step on another batch signature rebinds p.grad to fresh tensors, so step() binds them back before an eager
        # optimiser step reads them
        self._params = [p for p in model.model.parameters() if p.grad is not None]
        self._grads = [p.grad for p in self._params]
        self._loaded = {}

    def _fwd_bwd(self):
        loss = self.model.loss_fn(self.static, normalise=True)
        loss.backward()
        return loss.detach()

    def _eager(self):
        self.opt.zero_grad(set_to_none=True)
        loss = self._fwd_bwd()
        self.opt.step()
        return loss

    def matches(self, batch) -> bool:
        return batch_signature(batch) == self.signature

    def load(self, batch) -> None:
        """Copy a batch (DeviceBatch, or HostBatch in pinned memory) into the static input buffers."""
        static_ids = set()
        for c in batch.contexts:
            if not c.y_batched:       # device-resident static field (staging.BatchStager): copied only when it changes
                static_ids.update(id(t) for t in (c.x if isinstance(c.x, tuple) else (c.x,)) + (c.y, c.mask)
                                  if t is not None)
        for i, (dst, src) in enumerate(zip(_tensors(self.static), _tensors(batch))):
            if dst is None:
                continue
            if id(src) in static_ids and src.device.type == "cuda":
                if self._loaded.get(i) is src:
                    continue
                self._loaded[i] = src
            dst.copy_(src, non_blocking=True)

    def step(self, batch=None) -> torch.Tensor:
        if batch is not None:
            if isinstance(batch, DeviceBatch) and batch.ready is not None:
                torch.cuda.current_stream().wait_event(batch.ready)
                batch.ready = None
            self.load(batch)
        self.graph.replay()
        self.eng.weights_dirty = True          # the replay moved the masters: a later eager forward must re-pack
        if not self.capture_optimizer:
            for p, g in zip(self._params, self._grads):
                p.grad = g
            self.opt.step()
        return self.loss
