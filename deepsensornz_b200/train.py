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
                progress_bar: bool = False, tqdm_notebook: bool = False) -> List[float]:
    """One pass over ``tasks``.  ``batch_size=None``: one optimiser step per task; otherwise
    ``len(tasks)//batch_size`` steps on concatenated batches (the remainder is dropped, as upstream)."""
    if opt is None:
        opt = torch.optim.Adam(model.model.parameters(), lr=lr)

    def launch_step(task) -> torch.Tensor:
        """Queue forward + backward + optimiser step; returns the (device) loss without synchronising."""
        opt.zero_grad()
        items = task if isinstance(task, list) else [task]
        losses = [model.loss_fn(t, normalise=True) for t in items]
        mean_loss = torch.stack(losses).mean()
        mean_loss.backward()
        opt.step()
        return mean_loss.detach()

    order = np.random.permutation(len(tasks))
    tasks = [tasks[i] for i in order]
    n_batches = len(tasks) // batch_size if batch_size is not None else len(tasks)

    # Batch i+1 is concatenated, staged in pinned memory and copied on a side stream while the GPU runs step i
    # (the reference re-uploads every task synchronously inside loss_fn).
    can_stage = hasattr(model, "stage_task") and torch.cuda.is_available()
    copy_stream = torch.cuda.Stream() if can_stage else None

    def make(bi):
        if batch_size is not None:
            task = concat_tasks(tasks[bi * batch_size:(bi + 1) * batch_size])
        else:
            task = tasks[bi]
        if not can_stage:
            return task
        return model.engine.upload(model.stage_task(task, pinned=True), stream=copy_stream)

    it = range(n_batches)
    if progress_bar:
        try:
            from tqdm import tqdm
            it = tqdm(it)
        except Exception:
            pass
    losses = []
    nxt = make(0) if n_batches > 0 else None
    for bi in it:
        cur = nxt
        loss_t = launch_step(cur)
        nxt = make(bi + 1) if bi + 1 < n_batches else None
        losses.append(float(loss_t.cpu().numpy()))
    return losses
