"""Oracle-side restatement of the host half of the path -- TEST INFRASTRUCTURE ONLY (see convnp_oracle.py).

PARITY UNPINNED (same reason as convnp_oracle.py: upstream ``deepsensor==0.3.6`` is not importable here).

Independent of ``deepsensornz_b200.task`` / ``ConvNP.modify_task``: raw numpy task dicts (the layout
``TaskLoader_SampleStations.task_generation`` builds, /root/reference/nzdownscale/downscaler/train.py:560-637)
go straight to the torch CPU tensors the oracle consumes, following SURVEY.md Appendix A.1 (``modify_task``:
batch axis, float32, NaN -> ``Masked``) and A.8 (``concat_tasks`` / ``merge_contexts``: off-grid context sets
padded to the batch maximum with x = 0 / masked-out y, gridded sets stacked, equal N_t required --
the reason for the reference's grouping rule, train.py:448-475).

Used by the tests to check the product's ``concat_tasks`` / ``modify_task`` bit for bit and to feed the oracle
when golden vectors are generated, so that rows a1 / a10 of SURVEY section 8 are checked against a second
restatement rather than against the code under test.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch


def _f32(a) -> np.ndarray:
    return np.asarray(a, dtype=np.float32)


def _nan_to_mask(y: np.ndarray) -> Tuple[np.ndarray, Optional[np.ndarray]]:
    """A.1: arrays holding NaNs become (y with NaN -> 0, mask [B,1,...] = 1 where no channel is NaN)."""
    nan = np.isnan(y)
    if not nan.any():
        return y, None
    mask = (~nan.any(axis=1, keepdims=True)).astype(np.float32)
    out = y.copy()
    out[nan] = 0.0
    return out, mask


def task_tensors(tasks: Sequence[dict]):
    """-> (contexts [(x, y, mask|None)], xt, yt|None, aux_t|None) as torch CPU float32 tensors with batch axis.

    Off-grid sets: x [B,2,N], y [B,C,N], mask [B,1,N]; gridded: x (x1 [B,1,N1], x2 [B,1,N2]), y [B,C,N1,N2],
    mask [B,1,N1,N2].  On-grid targets: xt = (x1 [B,1,P], x2 [B,1,Q]).
    """
    tasks = list(tasks)
    B = len(tasks)
    n_sets = len(tasks[0]["X_c"])
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a))
    contexts = []
    for k in range(n_sets):
        xs = [t["X_c"][k] for t in tasks]
        ys = [_f32(t["Y_c"][k]) for t in tasks]
        if isinstance(xs[0], tuple):
            x = tuple(np.stack([_f32(xi[d]).reshape(1, -1) for xi in xs], axis=0) for d in range(2))   # [B,1,Nd]
            y = np.stack(ys, axis=0)
        else:
            n_max = max(int(xi.shape[-1]) for xi in xs)
            x = np.zeros((B, 2, n_max), dtype=np.float32)
            y = np.full((B, ys[0].shape[0], n_max), np.nan, dtype=np.float32)   # padding = masked out
            for b, (xi, yi) in enumerate(zip(xs, ys)):
                n = int(xi.shape[-1])
                x[b, :, :n] = _f32(xi)
                y[b, :, :n] = yi
        y, m = _nan_to_mask(y)
        contexts.append((tuple(T(v) for v in x) if isinstance(x, tuple) else T(x), T(y), None if m is None else T(m)))
    if len(tasks[0]["X_t"]) != 1:
        raise ValueError("single target set")
    x0 = tasks[0]["X_t"][0]
    if isinstance(x0, tuple):
        xt = tuple(T(np.stack([_f32(t["X_t"][0][d]).reshape(1, -1) for t in tasks], axis=0)) for d in range(2))
    else:
        nts = {int(t["X_t"][0].shape[-1]) for t in tasks}
        if len(nts) != 1:
            raise ValueError(f"tasks of one batch need the same number of targets, got {sorted(nts)}")
        xt = T(np.stack([_f32(t["X_t"][0]) for t in tasks], axis=0))
    yt = None
    if tasks[0].get("Y_t"):
        yt = T(np.stack([_f32(t["Y_t"][0]) for t in tasks], axis=0))
    aux = None
    if tasks[0].get("Y_t_aux") is not None:
        aux = T(np.stack([_f32(t["Y_t_aux"]) for t in tasks], axis=0))
    return contexts, xt, yt, aux
