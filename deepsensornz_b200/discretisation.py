"""Host-side internal-grid discretisation (the one piece of the path that fixes every kernel shape).

Mirrors upstream ``neuralprocesses/disc.py`` ``Discretisation`` as driven by
``deepsensor.model.convnp.ConvNP`` (variant 1; reference call sites
nzdownscale/downscaler/train.py:238-241, :370).  Kept in ONE function so the [U] assumption is
isolated; kernels only ever see ``(start1, n1, start2, n2, res)``.

Evaluated in float64 from the float32 extrema of every context input and the target input, so
the grid does not depend on float32 rounding near multiples of ``multiple``.
Grid point i along a dimension is ``float32(start + i * res)`` (the kernels compute exactly that).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Sequence

import numpy as np


@dataclass(frozen=True)
class GridSpec:
    start1: float
    n1: int
    start2: float
    n2: int
    res: float

    def points(self, dim: int) -> np.ndarray:
        s, n = (self.start1, self.n1) if dim == 0 else (self.start2, self.n2)
        return (s + np.arange(n, dtype=np.float64) * self.res).astype(np.float32)


def discretise_1d(lo: float, hi: float, ppu: float, margin: float = 0.1, multiple: int = 8):
    res = 1.0 / float(ppu)
    g_lo = float(lo) - margin - res
    g_hi = float(hi) + margin + res
    n_raw = (g_hi - g_lo) / res + 1.0
    n = math.ceil(n_raw / multiple - 1e-9) * multiple
    start = g_lo - (n - n_raw) * res / 2.0
    start = round(start / res) * res
    return start, int(n), res


def extent_of(xs: Sequence, dim: int):
    """Global (min, max) along ``dim`` over off-grid arrays [..,2,N] and gridded tuples (x1, x2).

    Accepts numpy arrays or torch tensors (CPU); empty sets are skipped.
    """
    lo, hi = math.inf, -math.inf
    for x in xs:
        v = x[dim] if isinstance(x, tuple) else x[..., dim, :]
        if hasattr(v, "numel"):
            if v.numel() == 0:
                continue
            lo, hi = min(lo, float(v.min())), max(hi, float(v.max()))
        else:
            v = np.asarray(v)
            if v.size == 0:
                continue
            lo, hi = min(lo, float(v.min())), max(hi, float(v.max()))
    return lo, hi


def discretise(xs: Sequence, ppu: float, margin: float = 0.1, multiple: int = 8) -> GridSpec:
    lo1, hi1 = extent_of(xs, 0)
    lo2, hi2 = extent_of(xs, 1)
    s1, n1, res = discretise_1d(lo1, hi1, ppu, margin, multiple)
    s2, n2, _ = discretise_1d(lo2, hi2, ppu, margin, multiple)
    return GridSpec(s1, n1, s2, n2, res)
