"""Deterministic synthetic NZ-shaped tasks.

Shapes follow what the reference feeds DeepSensor: ERA5-shaped base grid 140x140 with sea = NaN
(nzdownscale/downscaler/preprocess.py:534-535 descending latitude), 6-channel low-res aux grid,
1400x1400 land-mask context (validation.ipynb:213), ~200 stations split into context / target by
``TaskLoader_SampleStations.sample_df`` (nzdownscale/downscaler/train.py:529-558) and 5 aux-at-target
channels (validation wrf.ipynb:208).  Coordinates are normalised to [0,1]^2 like ``DataProcessor``
does with the hi-res topography extent (preprocess.py:771-778).  There is no network, so these
stand in for the real NetCDFs; everything is float32 from ``numpy.random.default_rng(seed)``.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional

import numpy as np

from .task import Task


def _smooth_field(rng, n: int, coarse: int = 9) -> np.ndarray:
    """Bilinear up-sampling of a coarse random field to n x n (smooth blobs)."""
    c = rng.standard_normal((coarse, coarse))
    xc = np.linspace(0.0, 1.0, coarse)
    xf = np.linspace(0.0, 1.0, n)
    rows = np.stack([np.interp(xf, xc, c[i]) for i in range(coarse)], axis=0)  # [coarse, n]
    return np.stack([np.interp(xf, xc, rows[:, j]) for j in range(n)], axis=1)  # [n, n]


@dataclass
class StaticFields:
    """Fields shared by every date: land mask, low-res aux, hi-res aux-at-target source."""
    x_lo: np.ndarray      # [140] ascending base-grid coordinate
    land_lo: np.ndarray   # [140,140] bool, indexed [x1 ascending, x2]
    c1: np.ndarray        # [6,140,140]
    x_hi: np.ndarray      # [n_hi]
    c2: np.ndarray        # [1,n_hi,n_hi] in {-1,+1}
    aux_hi: Optional[np.ndarray]  # [5,n_hi,n_hi] aux-at-target source (hi-res topo etc.)


def make_static(seed: int = 7, n_lo: int = 140, n_hi: int = 1400, land_frac: float = 0.15,
                with_aux_hi: bool = False) -> StaticFields:
    rng = np.random.default_rng(seed)
    x_lo = np.linspace(0.00357, 0.99643, n_lo).astype(np.float32)
    x_hi = np.linspace(0.0, 1.0, n_hi).astype(np.float32)
    f_hi = _smooth_field(rng, n_hi)
    thr = np.quantile(f_hi, 1.0 - land_frac)
    land_hi = f_hi > thr
    idx = np.clip(np.searchsorted(x_hi, x_lo), 0, n_hi - 1)
    land_lo = land_hi[np.ix_(idx, idx)]
    c1 = rng.uniform(-1.0, 1.0, (6, n_lo, n_lo)).astype(np.float32)
    c1[4] = x_lo[:, None]
    c1[5] = x_lo[None, :]
    c2 = np.where(land_hi, 1.0, -1.0).astype(np.float32)[None]
    aux_hi = rng.uniform(-1.0, 1.0, (5, n_hi, n_hi)).astype(np.float32) if with_aux_hi else None
    return StaticFields(x_lo, land_lo, c1, x_hi, c2, aux_hi)


def make_task(static: StaticFields, seed: int, n_stations: int = 200, context_frac: float = 0.8,
              c0_channels: int = 3, grid_targets: bool = False, all_context: bool = False) -> Task:
    """One synthetic daily/hourly NZ task in the raw (numpy, no batch dim) DeepSensor layout."""
    rng = np.random.default_rng(seed)
    n_lo = static.x_lo.shape[0]
    # c0: ERA5-like, latitude (x1) DESCENDING, sea = NaN on the physical channels
    x1_desc = static.x_lo[::-1].copy()
    c0 = rng.standard_normal((c0_channels, n_lo, n_lo)).astype(np.float32)
    sea = ~static.land_lo[::-1, :]
    n_phys = max(1, c0_channels - 2)
    c0[:n_phys, sea] = np.nan
    # stations on land cells, jittered inside the cell
    land_idx = np.argwhere(static.land_lo)
    pick = rng.choice(land_idx.shape[0], n_stations, replace=False)
    cell = (static.x_lo[1] - static.x_lo[0])
    jit = rng.uniform(-0.45, 0.45, (n_stations, 2)) * cell
    xs = np.stack([static.x_lo[land_idx[pick, 0]] + jit[:, 0],
                   static.x_lo[land_idx[pick, 1]] + jit[:, 1]], axis=0).astype(np.float32)  # [2,N]
    ys = rng.standard_normal((1, n_stations)).astype(np.float32)
    if all_context or grid_targets:
        ci = np.arange(n_stations)
        ti = np.arange(0)
    else:
        n_c = int(context_frac * n_stations)
        ci = rng.choice(n_stations, n_c, replace=False)
        ti = np.setdiff1d(np.arange(n_stations), ci)
    task = {
        "time": seed,
        "ops": [],
        # static fields hand over the SAME coordinate / value arrays for every date (as TaskLoader does for variables
        # without a time axis): predict then uploads and encodes them once
        "X_c": [(x1_desc[None], static.x_lo[None].copy()),
                (static.x_lo[None], static.x_lo[None]),
                (static.x_hi[None], static.x_hi[None]),
                xs[:, ci]],
        "Y_c": [c0, static.c1, static.c2, ys[:, ci]],
    }
    if grid_targets:
        task["X_t"] = [(static.x_hi[None].copy(), static.x_hi[None].copy())]
        task["Y_t"] = []
        task["Y_t_aux"] = static.aux_hi if static.aux_hi is not None else \
            rng.uniform(-1.0, 1.0, (5, static.x_hi.shape[0], static.x_hi.shape[0])).astype(np.float32)
    else:
        task["X_t"] = [xs[:, ti]]
        task["Y_t"] = [ys[:, ti]]
        task["Y_t_aux"] = rng.uniform(-1.0, 1.0, (5, ti.shape[0])).astype(np.float32)
    return Task(task)


def make_tasks(n: int, seed0: int = 20160101, static: Optional[StaticFields] = None, **kw) -> List[Task]:
    static = static or make_static()
    return [make_task(static, seed0 + i, **kw) for i in range(n)]
