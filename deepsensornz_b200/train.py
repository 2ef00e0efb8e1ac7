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

    def train_step(task) -> float:
        opt.zero_grad()
        items = task if isinstance(task, list) else [task]
        losses = [model.loss_fn(t, normalise=True) for t in items]
        mean_loss = torch.stack(losses).mean()
        mean_loss.backward()
        opt.step()
        return float(mean_loss.detach().cpu().numpy())

    order = np.random.permutation(len(tasks))
    tasks = [tasks[i] for i in order]
    n_batches = len(tasks) // batch_size if batch_size is not None else len(tasks)
    it = range(n_batches)
    if progress_bar:
        try:
            from tqdm import tqdm
            it = tqdm(it)
        except Exception:
            pass
    losses = []
    for bi in it:
        if batch_size is not None:
            task = concat_tasks(tasks[bi * batch_size:(bi + 1) * batch_size])
        else:
            task = tasks[bi]
        losses.append(train_step(task))
    return losses
