"""Minimal numpy-backed gridded variable used where xarray is not installed (this build container).

The reference feeds DeepSensor ``xarray`` objects (nzdownscale/downscaler/train.py:141-166).  ``TaskLoader``
(loader.py) accepts real xarray objects when xarray is importable and this ``GridVar`` otherwise; both expose
the tiny surface the loader needs: named data variables, ``x1`` / ``x2`` (and optional ``time``) coordinates,
time selection and nearest-neighbour sampling.
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence

import numpy as np


class GridVar:
    """data_vars: name -> array [T,N1,N2] (or [N1,N2] when static); coords x1 [N1], x2 [N2], time [T]."""

    def __init__(self, data_vars: Dict[str, np.ndarray], x1: np.ndarray, x2: np.ndarray,
                 time: Optional[Sequence] = None):
        self.x1 = np.asarray(x1)
        self.x2 = np.asarray(x2)
        self.time = None if time is None else np.asarray(time, dtype="datetime64[ns]")
        self.data_vars = {k: np.asarray(v) for k, v in data_vars.items()}
        for k, v in self.data_vars.items():
            want = (len(self.x1), len(self.x2)) if self.time is None else (len(self.time), len(self.x1), len(self.x2))
            if v.shape != want:
                raise ValueError(f"{k}: shape {v.shape} != {want}")

    @property
    def var_IDs(self):
        return tuple(self.data_vars)

    def sel_time(self, date) -> "GridVar":
        if self.time is None:
            return self
        d = np.datetime64(date, "ns")
        idx = np.nonzero(self.time == d)[0]
        if idx.size == 0:
            raise KeyError(f"time {date} not in variable")
        return GridVar({k: v[idx[0]] for k, v in self.data_vars.items()}, self.x1, self.x2, None)

    def stack(self) -> np.ndarray:
        """[C,N1,N2] of a time-less variable."""
        if self.time is not None:
            raise ValueError("select a time first")
        return np.stack([self.data_vars[k] for k in self.data_vars], axis=0)

    @staticmethod
    def _nearest(index: np.ndarray, q: np.ndarray) -> np.ndarray:
        """pandas ``get_indexer(method='nearest')`` semantics on a monotone index (ties -> larger index)."""
        idx = np.asarray(index, dtype=np.float64)
        asc = idx[0] <= idx[-1]
        a = idx if asc else idx[::-1]
        pos = np.searchsorted(a, q, side="left")
        lo = np.clip(pos - 1, 0, len(a) - 1)
        hi = np.clip(pos, 0, len(a) - 1)
        pick = np.where(np.abs(q - a[lo]) < np.abs(a[hi] - q), lo, hi)
        return pick if asc else len(a) - 1 - pick

    def sel_nearest(self, x1, x2, grid: bool) -> np.ndarray:
        """grid=True: outer product [C,len(x1),len(x2)]; grid=False: pointwise [C,N]."""
        arr = self.stack()
        i = self._nearest(self.x1, np.asarray(x1, dtype=np.float64))
        j = self._nearest(self.x2, np.asarray(x2, dtype=np.float64))
        return arr[:, i][:, :, j] if grid else arr[:, i, j]
