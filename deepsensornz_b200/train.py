"""``train_epoch`` / ``set_gpu_default_device`` -- mirror of ``deepsensor.train.train``.

Reference call sites: nzdownscale/downscaler/train.py:48 (set_gpu_default_device) and
train.py:388-394 (``train_epoch(model, tasks_k, batch_size=len(tasks_k), lr=lr, opt=opt)``).
"""
from __future__ import annotations

from typing import List, Optional

import numpy as np
import torch

from .task import Task, concat_tasks


def set_gpu_default_device() -> None:
    """Make CUDA the default torch device (what ``deepsensor.train.set_gpu_default_device`` does)."""
    if not torch.cuda.is_available():
        raise RuntimeError("No GPU available: the B200 ConvNP path has no CPU fallback")
    torch.set_default_device("cuda")


def train_epoch(model, tasks: List[Task], lr: float = 5e-5, batch_size: Optional[int] = None, opt=None,
                progress_bar: bool = False, tqdm_notebook: bool = False, use_graph: Optional[bool] = None) -> List[float]:
    """One pass over ``tasks``.  ``batch_size=None``: one optimiser step per task; otherwise
    ``len(tasks)//batch_size`` steps on concatenated batches (the remainder is dropped, as upstream).

    ``use_graph`` (default: env ``CONVNP_B200_GRAPH=1``): replay forward + backward (+ the optimiser step when it is
    ``capturable``) as a CUDA graph from the second batch with the same shape signature on (graph.py)."""
    import os
    if use_graph is None:
        use_graph = os.environ.get("CONVNP_B200_GRAPH", "0") == "1"
    if opt is None:
        opt = torch.optim.Adam(model.model.parameters(), lr=lr)

    def launch_step(task) -> torch.Tensor:
        """Queue forward + backward + optimiser step; returns the (device) loss without synchronising."""
        opt.zero_grad()
        mean_loss = model.loss_fn(task, normalise=True)
        mean_loss.backward()
        opt.step()
        return mean_loss.detach()

    order = np.random.permutation(len(tasks))
    tasks = [tasks[i] for i in order]
    n_batches = len(tasks) // batch_size if batch_size is not None else len(tasks)

    # Batch i+1 is built (straight into persistent page-locked buffers, static context sets resident on the device:
    # staging.BatchStager) and copied on a side stream by a worker thread while the GPU runs step i; the reference
    # re-concatenates and re-uploads every task synchronously inside loss_fn.
    can_stage = hasattr(model, "stage_task") and torch.cuda.is_available()
    copy_stream = None
    stager = None
    if can_stage:
        from .staging import BatchStager
        copy_stream = model.__dict__.get("_copy_stream")
        if copy_stream is None:
            copy_stream = model.__dict__["_copy_stream"] = torch.cuda.Stream()
        stager = model.__dict__.get("_stager")
        if stager is None:
            stager = model.__dict__["_stager"] = BatchStager(model.engine)
        dev_index = model.engine.device.index

    def make(bi):
        group = tasks[bi * batch_size:(bi + 1) * batch_size] if batch_size is not None else [tasks[bi]]
        if not can_stage:
            return concat_tasks(group) if len(group) > 1 else group[0]
        if dev_index is not None:
            torch.cuda.set_device(dev_index)          # worker thread: the current device is per thread
        if stager.fast_path_ok(group):
            return stager.upload(stager.build(group), stream=copy_stream)
        task = concat_tasks(group) if len(group) > 1 else group[0]
        return model.engine.upload(model.stage_task(task, pinned=True), stream=copy_stream)

    it = range(n_batches)
    if progress_bar:
        try:
            from tqdm import tqdm
            it = tqdm(it)
        except Exception:
            pass
    dp = getattr(model.engine, "world_size", 1) > 1
    use_graph = bool(use_graph) and can_stage
    graphs = model.__dict__.setdefault("_train_graphs", {}) if use_graph else None
    losses = []
    # The loss of batch i is copied to page-locked memory asynchronously and READ while batch i+1 is already queued:
    # every batch's loss still reaches the host (as upstream's train_epoch returns it), but the GPU never waits for the
    # host to come back from a synchronising read between two steps.
    on_gpu = torch.cuda.is_available() and can_stage
    slots = [torch.empty((), dtype=torch.float64, device="cpu", pin_memory=True) for _ in range(2)] if on_gpu else None
    events = [None, None]

    def read_back(j):
        events[j % 2].synchronize()
        losses.append(float(slots[j % 2]))

    from concurrent.futures import ThreadPoolExecutor
    pool = ThreadPoolExecutor(max_workers=1) if can_stage and n_batches > 1 else None
    try:
        nxt = make(0) if n_batches > 0 else None
        for bi in it:
            cur = nxt
            fut = None
            submit = lambda: pool.submit(make, bi + 1) if (pool is not None and bi + 1 < n_batches) else None
            if use_graph and not isinstance(cur, list):
                from .graph import GraphedTrainStep, batch_signature
                key = (id(opt), batch_signature(cur))
                gs = graphs.get(key)
                if gs is None:            # first batch of this shape: eager (it is also the warm-up of the capture)
                    fut = submit()
                    loss_t = launch_step(cur)
                    graphs[key] = False
                else:
                    if gs is False:       # second batch: capture (no concurrent staging while the graph is recorded)
                        gs = graphs[key] = GraphedTrainStep(model, opt, cur, warm=True)
                    fut = submit()
                    loss_t = gs.step(cur)
            else:
                fut = submit()
                loss_t = launch_step(cur)
            if on_gpu and loss_t.is_cuda:
                slots[bi % 2].copy_(loss_t.to(torch.float64), non_blocking=True)
                events[bi % 2] = torch.cuda.Event()
                events[bi % 2].record()
            if fut is not None:
                nxt = fut.result()
            elif bi + 1 < n_batches:
                nxt = make(bi + 1)
            else:
                nxt = None
            if on_gpu and loss_t.is_cuda:
                if bi > 0:
                    read_back(bi - 1)
            else:
                losses.append(float(loss_t.cpu().numpy()))
        if on_gpu and n_batches > 0 and len(losses) < n_batches:
            read_back(n_batches - 1)
    finally:
        if pool is not None:
            pool.shutdown(wait=True)
    return losses
