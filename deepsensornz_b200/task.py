"""``Task`` / ``concat_tasks`` -- host-side mirror of ``deepsensor.data.task``.

The reference builds one ``Task`` per date (nzdownscale/downscaler/train.py:308-334, :560-637),
groups them by number of target stations (train.py:448-475) and lets ``train_epoch`` concatenate
each group (train.py:388-390).  This module keeps that data layout bit-exactly:

  X_c : list, per context set either (x1 [1,N1], x2 [1,N2]) (gridded) or [2,N] (off-grid)
  Y_c : list, [C,N1,N2] or [C,N]
  X_t : list of [2,N_t] (or a tuple for on-grid targets), Y_t : list of [C_t,N_t], Y_t_aux [C_a,N_t]

All payloads are numpy; nothing here touches the GPU.
"""
from __future__ import annotations

import copy
from typing import Callable, List, Optional, Tuple

import numpy as np


class Masked:
    """A context observation with a 0/1 validity mask (mirror of ``neuralprocesses.mask.Masked``)."""

    __slots__ = ("y", "mask")

    def __init__(self, y, mask):
        self.y = y
        self.mask = mask

    def __repr__(self):
        return f"Masked(y={getattr(self.y, 'shape', None)}, mask={getattr(self.mask, 'shape', None)})"


def _recurse(f: Callable, v):
    if isinstance(v, list):
        return [_recurse(f, e) for e in v]
    if isinstance(v, tuple):
        return tuple(_recurse(f, e) for e in v)
    if isinstance(v, Masked):
        return Masked(_recurse(f, v.y), _recurse(f, v.mask))
    if isinstance(v, (np.ndarray, np.ma.MaskedArray)):
        return f(v)
    return v


class Task(dict):
    """Dict of numpy payloads plus the list of ops already applied (``task['ops']``)."""

    def __init__(self, task_dict: dict):
        super().__init__(task_dict)
        if "ops" not in self:
            self["ops"] = []

    # -- generic op -----------------------------------------------------------------------
    def op(self, f: Callable, op_flag: Optional[str] = None) -> "Task":
        new = Task({k: v for k, v in self.items()})
        new["ops"] = list(self["ops"])
        for k in ("X_c", "Y_c", "X_t", "Y_t", "Y_t_aux"):
            if k in new and new[k] is not None:
                new[k] = _recurse(f, new[k])
        if op_flag:
            new["ops"].append(op_flag)
        return new

    def add_batch_dim(self) -> "Task":
        return self.op(lambda a: a[np.newaxis, ...], "batch_dim")

    def cast_to_float32(self) -> "Task":
        return self.op(lambda a: a.astype(np.float32, copy=False), "float32")   # no copy when already float32

    def remove_target_nans(self) -> "Task":
        """Drop off-grid target points whose observation is NaN (before a batch dim exists)."""
        new = Task({k: v for k, v in self.items()})
        new["ops"] = list(self["ops"]) + ["target_nans_removed"]
        X_t, Y_t = list(new["X_t"]), list(new["Y_t"])
        for i, (x, y) in enumerate(zip(X_t, Y_t)):
            if isinstance(x, tuple) or y is None:
                continue
            bad = np.any(np.isnan(y), axis=0)
            if bad.any():
                X_t[i], Y_t[i] = x[:, ~bad], y[:, ~bad]
                if i == 0 and new.get("Y_t_aux") is not None:
                    new["Y_t_aux"] = new["Y_t_aux"][:, ~bad]
        new["X_t"], new["Y_t"] = X_t, Y_t
        return new

    def mask_nans_numpy(self) -> "Task":
        def f(a):
            if isinstance(a, np.ma.MaskedArray):
                return a
            if is_batch_broadcast(a):                 # one shared slice: scan it once
                one = a[:1]
                if np.issubdtype(a.dtype, np.floating) and np.isnan(one).any():
                    return np.ma.MaskedArray(a, mask=np.broadcast_to(np.isnan(one), a.shape), fill_value=np.nan)
                return a
            if np.issubdtype(a.dtype, np.floating) and np.isnan(a).any():
                return np.ma.MaskedArray(a, mask=np.isnan(a), fill_value=np.nan)
            return a

        return self.op(f, "numpy_mask")

    def mask_nans_nps(self) -> "Task":
        """Context ``MaskedArray`` -> ``Masked(y with NaN->0, mask [B,1,...])``; targets untouched."""
        new = Task({k: v for k, v in self.items()})
        new["ops"] = list(self["ops"]) + ["nps_mask"]

        def f(a):
            if isinstance(a, np.ma.MaskedArray):
                if is_batch_broadcast(a.data):        # one shared slice: mask it once, hand out zero-stride views
                    one = f(np.ma.MaskedArray(a.data[:1], mask=np.ma.getmaskarray(a)[:1]))
                    B = a.shape[0]
                    return Masked(np.broadcast_to(one.y, (B,) + one.y.shape[1:]),
                                  np.broadcast_to(one.mask, (B,) + one.mask.shape[1:]))
                m = np.ma.getmaskarray(a)
                mask = (~np.any(m, axis=1, keepdims=True)).astype(a.dtype)
                y = np.array(a.data, copy=True)
                y[m] = 0.0
                return Masked(y, mask)
            return a

        new["Y_c"] = [f(a) for a in new["Y_c"]]
        return new

    def summary(self) -> str:
        def shp(v):
            if isinstance(v, (list, tuple)):
                return type(v)(shp(e) for e in v)
            if isinstance(v, Masked):
                return ("Masked", v.y.shape)
            return getattr(v, "shape", v)

        return "\n".join(f"{k}: {shp(v)}" for k, v in self.items())


def _pad_last(a: np.ndarray, n: int, value: float) -> np.ndarray:
    if a.shape[-1] == n:
        return a
    pad = np.full(a.shape[:-1] + (n - a.shape[-1],), value, dtype=a.dtype)
    return np.concatenate([a, pad], axis=-1)


def same_buffer(arrays) -> bool:
    """True when every array is the SAME memory (pointer, shape, strides, dtype): a static field -- topography aux,
    land mask -- that the TaskLoader hands to every date without copying (variables without a time axis)."""
    a0 = arrays[0]
    if not isinstance(a0, np.ndarray) or isinstance(a0, np.ma.MaskedArray):
        return False
    k0 = (a0.__array_interface__["data"][0], a0.shape, a0.strides, a0.dtype)
    for a in arrays[1:]:
        if a is a0:
            continue
        if not isinstance(a, np.ndarray) or isinstance(a, np.ma.MaskedArray):
            return False
        if (a.__array_interface__["data"][0], a.shape, a.strides, a.dtype) != k0:
            return False
    return True


def is_batch_broadcast(a) -> bool:
    """A stacked array whose batch axis has stride 0: ``concat_tasks`` made it from one shared buffer."""
    return isinstance(a, np.ndarray) and a.ndim >= 1 and a.shape[0] > 1 and a.strides[0] == 0


def _stack(arrays):
    """Concatenate along the batch axis; arrays that are all the same buffer become a zero-stride view of it (same
    shape and values as the copy, none of its 16 x 7.8 MB of traffic for a 1400 x 1400 land mask)."""
    if len(arrays) > 1 and same_buffer(arrays):
        a = arrays[0]
        return np.broadcast_to(a[:1], (len(arrays) * a.shape[0],) + a.shape[1:]) if a.shape[0] == 1 else \
            np.concatenate(arrays, axis=0)
    return np.concatenate(arrays, axis=0)


def merge_contexts(contexts: List[Tuple], multiple: int = 1):
    """Merge one context set across tasks (mirror of ``neuralprocesses.mask.merge_contexts``).

    Off-grid sets are padded along N to the batch maximum (rounded up to ``multiple``): x with 0,
    y with NaN (turned into mask 0 by the masking ops afterwards).  Gridded sets are stacked.
    """
    xs, ys = [c[0] for c in contexts], [c[1] for c in contexts]
    if isinstance(xs[0], tuple):
        x = tuple(_stack([xi[d] for xi in xs]) for d in range(len(xs[0])))
        return x, _stack(ys)
    n = max(xi.shape[-1] for xi in xs)
    n = ((n + multiple - 1) // multiple) * multiple
    x = np.concatenate([_pad_last(xi, n, 0.0) for xi in xs], axis=0)
    y = np.concatenate([_pad_last(yi, n, np.nan) for yi in ys], axis=0)
    return x, y


def concat_tasks(tasks: List[Task], multiple: int = 1) -> Task:
    """Concatenate tasks along a leading batch axis (mirror of ``deepsensor.data.task.concat_tasks``).

    Requires an equal number of targets per task -- the reason the reference groups tasks by
    station count first (nzdownscale/downscaler/train.py:448-475).
    """
    tasks = list(tasks)
    if len(tasks) == 1:
        return tasks[0]
    for i, t in enumerate(tasks):
        if "numpy_mask" in t["ops"] or "nps_mask" in t["ops"]:
            raise ValueError("Cannot concatenate tasks that have had NaNs masked; masking is applied "
                             "automatically after concatenation.")
        if "target_nans_removed" not in t["ops"]:
            t = t.remove_target_nans()
        if "batch_dim" not in t["ops"]:
            t = t.add_batch_dim()
        if "float32" not in t["ops"]:
            t = t.cast_to_float32()
        tasks[i] = t
    n_sets = [len(t["Y_t"]) for t in tasks]
    if len(set(n_sets)) != 1:
        raise ValueError(f"All tasks must have the same number of target sets, got {n_sets}")
    for i in range(n_sets[0]):
        n_obs = [t["Y_t"][i].shape[-1] if not isinstance(t["X_t"][i], tuple) else t["Y_t"][i].shape[-2:]
                 for t in tasks]
        if len(set(n_obs)) != 1:
            raise ValueError("All tasks must have the same number of targets to concatenate: "
                             f"got {n_obs}. Group tasks by number of targets first.")
    merged = copy.copy(tasks[0])
    merged = Task({k: v for k, v in merged.items()})
    merged["ops"] = list(tasks[0]["ops"])
    contexts = [list(zip(t["X_c"], t["Y_c"])) for t in tasks]
    merged_ctx = [merge_contexts(list(cs), multiple) for cs in zip(*contexts)]
    merged["X_c"] = [c[0] for c in merged_ctx]
    merged["Y_c"] = [c[1] for c in merged_ctx]
    X_t, Y_t = [], []
    for i in range(n_sets[0]):
        if isinstance(tasks[0]["X_t"][i], tuple):
            X_t.append(tuple(np.concatenate([t["X_t"][i][d] for t in tasks], axis=0) for d in range(2)))
        else:
            X_t.append(np.concatenate([t["X_t"][i] for t in tasks], axis=0))
        Y_t.append(np.concatenate([t["Y_t"][i] for t in tasks], axis=0))
    merged["X_t"], merged["Y_t"] = X_t, Y_t
    if tasks[0].get("Y_t_aux") is not None:
        merged["Y_t_aux"] = np.concatenate([t["Y_t_aux"] for t in tasks], axis=0)
    merged["time"] = [t.get("time") for t in tasks]
    return merged.mask_nans_numpy().mask_nans_nps()


def convert_task_to_nps_args(task: Task):
    """(context_data, xt, yt, model_kwargs) as ``deepsensor.model.nps.convert_task_to_nps_args``."""
    context_data = list(zip(task["X_c"], task["Y_c"]))
    if len(task["X_t"]) != 1:
        raise NotImplementedError("the ConvNP hot path supports a single target set")
    xt = task["X_t"][0]
    yt = task["Y_t"][0] if task.get("Y_t") else None
    kw = {}
    if task.get("Y_t_aux") is not None:
        kw["aux_t"] = task["Y_t_aux"]
    return context_data, xt, yt, kw
