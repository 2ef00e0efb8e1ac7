"""``ConvNP.predict`` -- mirror of ``deepsensor.model.model.DeepSensorModel.predict`` for the calls nzdownscale makes.

  self.model.predict(task, X_t=self.ds_elev, progress_bar=True, transform_params=...)   validate_ERA.py:88-92
  self.model.predict(task, X_t=self.ds_elev, progress_bar=True, transform_params=...)   validate_WRF.py:227-231
  model.predict(test_task, X_t=highres_aux_raw_ds, progress_bar=1)                      validate.py:1106
  pred['{var}_station']['mean'].where(mask)                                             validate_ERA.py:94-96

Per task: contexts -> encoder -> UNet -> on-grid (or off-grid) decoder -> MLP head, all on the GPU; the static
aux-at-target tensor (5 x 1400 x 1400 = 39 MB) is uploaded ONCE per call instead of once per task.
``transform_params`` is accepted for call compatibility with the authors' locally modified DeepSensor
 and must be None-equivalent: it is ignored.

Returns ``Prediction``: ``{target_var_ID: dataset}`` where dataset is an ``xarray.Dataset`` with ``mean`` / ``std``
over (time, x1, x2) when xarray is importable and X_t was an xarray object, else a ``GridResult`` / DataFrame.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np
import pandas as pd
import torch

from .task import Task

try:  # pragma: no cover
    import xarray as xr
except Exception:  # noqa: BLE001
    xr = None


class _Field:
    """Tiny stand-in for an xarray.DataArray: ``.values`` [T,N1,N2] and ``.where(mask)``."""

    def __init__(self, values: np.ndarray):
        self.values = values

    def where(self, mask):
        m = np.asarray(getattr(mask, "values", mask), dtype=bool)
        return _Field(np.where(np.broadcast_to(m, self.values.shape), self.values, np.nan))

    def __array__(self, dtype=None):
        return self.values if dtype is None else self.values.astype(dtype)


class GridResult(dict):
    """``{'mean': _Field, 'std': _Field}`` plus coordinates; used when xarray is unavailable."""

    def __init__(self, mean, std, time, x1, x2):
        super().__init__(mean=_Field(mean), std=_Field(std))
        self.coords = {"time": time, "x1": x1, "x2": x2}


class Prediction(dict):
    pass


def _target_coords(model, X_t, X_t_is_normalised: bool):
    """-> (mode, X_t_norm, raw coords for the output).  mode 'on-grid' gives a tuple (x1[N1], x2[N2])."""
    dp = model.data_processor
    if xr is not None and isinstance(X_t, (xr.Dataset, xr.DataArray)):
        Xn = X_t if (X_t_is_normalised or dp is None) else dp.map_coords(X_t)
        x1, x2 = np.asarray(Xn.coords["x1"].values), np.asarray(Xn.coords["x2"].values)
        return "on-grid", (x1.astype(np.float32), x2.astype(np.float32)), X_t
    if isinstance(X_t, (pd.DataFrame, pd.Series, pd.Index)):
        Xn = X_t if (X_t_is_normalised or dp is None) else dp.map_coords(X_t)
        idx = Xn.index if isinstance(Xn, (pd.DataFrame, pd.Series)) else Xn
        arr = np.stack([idx.get_level_values("x1").values, idx.get_level_values("x2").values]).astype(np.float32)
        return "off-grid", arr, X_t
    if isinstance(X_t, tuple):
        if dp is not None and not X_t_is_normalised:
            raise TypeError("pass xarray/pandas X_t (raw coordinates) or set X_t_is_normalised=True for numpy tuples")
        return "on-grid", tuple(np.asarray(v, dtype=np.float32).reshape(-1) for v in X_t), X_t
    arr = np.asarray(X_t, dtype=np.float32)
    if arr.ndim == 2 and arr.shape[0] == 2:
        return "off-grid", arr, X_t
    raise TypeError(f"unsupported X_t type {type(X_t)}")


def predict(model, tasks, X_t, X_t_mask=None, X_t_is_normalised: bool = False, aux_at_targets_override=None,
            resolution_factor: int = 1, pred_params: Sequence[str] = ("mean", "std"), unnormalise: bool = True,
            progress_bar: int = 0, transform_params=None) -> Prediction:
    if isinstance(tasks, Task):
        tasks = [tasks]
    tasks = list(tasks)
    if resolution_factor != 1:
        raise NotImplementedError("resolution_factor != 1 (the reference never passes it)")
    mode, Xn, X_raw = _target_coords(model, X_t, X_t_is_normalised)
    tl = model.task_loader
    # --- aux at targets, once per call ---
    aux = None
    if aux_at_targets_override is not None:
        aux = np.asarray(aux_at_targets_override, dtype=np.float32)
    elif tl is not None and getattr(tl, "aux_at_targets", None) is not None:
        aux = tl.sample_offgrid_aux(Xn, tl.aux_at_targets)
    elif tasks and tasks[0].get("Y_t_aux") is not None:
        aux = np.asarray(tasks[0]["Y_t_aux"], dtype=np.float32)
        want = (len(Xn[0]), len(Xn[1])) if mode == "on-grid" else (Xn.shape[1],)
        if tuple(aux.shape[1:]) != want:
            raise ValueError("task Y_t_aux does not match X_t; pass aux_at_targets_override")
    eng = model.engine
    aux_dev = None
    if aux is not None:
        t = torch.from_numpy(np.ascontiguousarray(aux[np.newaxis]))
        aux_dev = t.pin_memory().to(eng.device, non_blocking=True) if torch.cuda.is_available() else t
    it = tasks
    if progress_bar:
        try:
            from tqdm import tqdm
            it = tqdm(tasks)
        except Exception:  # noqa: BLE001
            pass
    # Per task: only the per-date tensors travel (static context sets are uploaded once, ctx_cache); mean / std come
    # back through two page-locked buffers with asynchronous copies, so the D2H of task i overlaps the kernels of
    # task i+1 (the reference reads every result back synchronously).
    n = len(tasks)
    mean_out = std_out = None
    times = []
    ctx_cache = {}
    cuda = torch.cuda.is_available()
    pin = [None, None, None]
    events = [None, None, None]
    pending = []            # (slot, task index)
    direct = None

    def drain(slot, idx):
        events[slot].synchronize()
        mean_out[idx] = pin[slot][0].numpy()
        std_out[idx] = pin[slot][1].numpy()

    # the pinned -> result-array memcpy (2 x 7.8 MB per 1400^2 task) runs on a worker thread (numpy releases the GIL),
    # so it overlaps the launches of the following tasks
    copy_stream = torch.cuda.Stream() if cuda else None
    from concurrent.futures import ThreadPoolExecutor
    pool = ThreadPoolExecutor(max_workers=1)
    futures = [None, None, None]

    for idx, task in enumerate(it):
        t2 = Task({k: v for k, v in task.items() if k not in ("Y_t", "Y_t_aux", "X_t")})
        t2["ops"] = list(task["ops"])
        if "batch_dim" in t2["ops"]:
            raise ValueError("predict expects un-batched tasks (one per time), as produced by the TaskLoader")
        t2["X_t"] = [tuple(v[np.newaxis] for v in Xn)] if mode == "on-grid" else [Xn]
        t2["Y_t"] = []
        hb = model.stage_task(t2, pinned=False, ctx_cache=ctx_cache)
        hb.aux_t = aux_dev
        # upload on a copy stream: a pageable H2D copy on the compute stream would block the host until the previous
        # task's kernels have drained
        out = model(eng.upload(hb, stream=copy_stream) if copy_stream is not None else hb)
        mean, std = out["mean"][0, 0], out["std"][0, 0]
        if mean_out is None:
            mean_out = np.empty((n,) + tuple(mean.shape), dtype=np.float32)
            std_out = np.empty_like(mean_out)
            if cuda:
                if direct is None:
                    pin = [(torch.empty(mean.shape, dtype=torch.float32).pin_memory(),
                            torch.empty(mean.shape, dtype=torch.float32).pin_memory()) for _ in range(3)]
        if cuda:
            slot = idx % 3
            if futures[slot] is not None:  # the buffer we are about to reuse must have been drained
                futures[slot].result()
            pin[slot][0].copy_(mean, non_blocking=True)
            pin[slot][1].copy_(std, non_blocking=True)
            events[slot] = torch.cuda.Event()
            events[slot].record()
            futures[slot] = pool.submit(drain, slot, idx)
        else:
            mean_out[idx], std_out[idx] = mean.numpy(), std.numpy()
        times.append(task.get("time"))
    for f in futures:
        if f is not None:
            f.result()
    pool.shutdown()
    mean, std = mean_out, std_out
    var_ID = "target"
    if tl is not None and getattr(tl, "target_var_IDs", None):
        var_ID = tl.target_var_IDs[0][0]
    dp = model.data_processor
    if unnormalise and dp is not None:
        mean = dp.map_array(mean, var_ID, unnorm=True)
        std = dp.map_array(std, var_ID, unnorm=True, add_offset=False)
    pred = Prediction()
    if mode == "on-grid":
        if xr is not None and isinstance(X_raw, (xr.Dataset, xr.DataArray)):
            coords = {"time": times, "x1": X_raw.coords[list(X_raw.dims)[-2]], "x2": X_raw.coords[list(X_raw.dims)[-1]]}
            dims = ("time",) + tuple(list(X_raw.dims)[-2:])
            pred[var_ID] = xr.Dataset({"mean": (dims, mean), "std": (dims, std)},
                                      coords={"time": times, dims[1]: X_raw.coords[dims[1]], dims[2]: X_raw.coords[dims[2]]})
        else:
            pred[var_ID] = GridResult(mean, std, times, Xn[0], Xn[1])
    else:
        idx = pd.MultiIndex.from_product([times, range(mean.shape[1])], names=["time", "point"])
        pred[var_ID] = pd.DataFrame({"mean": mean.reshape(-1), "std": std.reshape(-1)}, index=idx)
    return pred
