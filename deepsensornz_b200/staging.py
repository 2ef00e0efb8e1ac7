"""Device-resident task pipeline for ``train_epoch`` (VERDICT r01 "next" #2).

The reference keeps every task as numpy (nzdownscale/downscaler/train.py:315-316) and re-concatenates, re-masks and
re-uploads all of it inside every ``train_epoch`` batch (train.py:388-394 -> upstream ``concat_tasks`` + ``loss_fn``):
per 16-task batch that is 16 copies of the 7.84 MB land mask and ~9 MB of H2D per task.  ``BatchStager`` builds the
same batch -- bit for bit what ``concat_tasks`` + ``ConvNP.stage_task`` give (tests/test_staging_cpu.py) -- with none of
that traffic:

  * **static context sets** (every task hands over the SAME buffer: topography aux, land mask -- variables without a
    time axis) are recognised by buffer identity, converted and uploaded ONCE, and their device copy is reused by
    every later batch and epoch (weak references guard against a recycled address);
  * **per-date sets** (base grid, stations, targets, aux-at-targets) are written straight from the tasks into a ring of
    persistent page-locked buffers -- no intermediate concatenated arrays, no fresh ``pin_memory()`` per batch --
    with NaNs left in place: the encoder kernels derive the validity masks on the GPU (identical to the host
    ``Masked`` path, tests/test_gpu_parity.py::test_raw_task_equals_masked_task), off-grid sets are padded with
    x = 0 / y = NaN exactly like ``merge_contexts``;
  * the batch is uploaded on a copy stream by the worker thread that built it, so host work and H2D of batch i+1
    overlap the kernels of batch i.

Only plain (un-batched, un-masked) tasks with off-grid targets take this path; anything else falls back to
``concat_tasks`` + ``stage_task`` (same result, more host work).
"""
from __future__ import annotations

import time
import weakref
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

from .discretisation import discretise
from .engine import DeviceBatch, DeviceContext, Engine, HostBatch, _mono_rows
from .task import Task, same_buffer


def _bufkey(a: np.ndarray):
    return (a.__array_interface__["data"][0], a.shape, a.strides, str(a.dtype))


def _owner(a: np.ndarray):
    while isinstance(a.base, np.ndarray):
        a = a.base
    return a


class _Slot:
    """One set of persistent pinned buffers (keyed by name + shape) and the event of its last upload."""

    def __init__(self):
        self.bufs: Dict[tuple, torch.Tensor] = {}
        self.event: Optional[torch.cuda.Event] = None

    def get(self, name: str, shape: Tuple[int, ...]) -> torch.Tensor:
        k = (name, tuple(shape))
        t = self.bufs.get(k)
        if t is None:
            pin = torch.cuda.is_available()
            t = self.bufs[k] = torch.empty(tuple(shape), dtype=torch.float32, device="cpu", pin_memory=pin)
        return t


class BatchStager:
    def __init__(self, engine: Engine, slots: int = 3):
        self.engine = engine
        self.slots = [_Slot() for _ in range(slots)]
        self.n = 0
        self._static: Dict[tuple, tuple] = {}      # buffer key -> (weakrefs of the owning arrays, DeviceContext, extent)
        self._seen: Dict[tuple, tuple] = {}
        self._coords: Dict[tuple, tuple] = {}      # coordinate bytes -> (x_host, mono, band_cache): host band hints survive batches
        self.host_ms: List[float] = []             # build time of every batch (bench.py reports the mean)
        self.h2d_bytes = 0                         # bytes uploaded for the last batch

    # ------------------------------------------------------------------------------------------
    @staticmethod
    def fast_path_ok(tasks: List[Task]) -> bool:
        t0 = tasks[0]
        n_sets = len(t0["X_c"])
        for t in tasks:
            if any(op in t["ops"] for op in ("batch_dim", "numpy_mask", "nps_mask")):
                return False
            if len(t["X_t"]) != 1 or isinstance(t["X_t"][0], tuple) or not t.get("Y_t"):
                return False
            if len(t["X_c"]) != n_sets or t.get("Y_t_aux") is None:
                return False
            for y in list(t["Y_c"]) + [t["Y_t"][0], t["Y_t_aux"]]:
                if not isinstance(y, np.ndarray) or isinstance(y, np.ma.MaskedArray):
                    return False
        return True

    # ------------------------------------------------------------------------------------------
    def _static_set(self, x, y) -> Optional[tuple]:
        """Device-resident copy of a context set every task shares; None when the buffers are not (yet) known."""
        key = (_bufkey(x[0]), _bufkey(x[1]), _bufkey(y))
        owners = (_owner(x[0]), _owner(x[1]), _owner(y))
        ent = self._static.get(key)
        if ent is not None and all(r() is o for r, o in zip(ent[0], owners)):
            return ent
        try:
            refs = tuple(weakref.ref(o) for o in owners)
        except TypeError:
            return None
        eng = self.engine
        x1 = np.ascontiguousarray(x[0], dtype=np.float32).reshape(1, -1)
        x2 = np.ascontiguousarray(x[1], dtype=np.float32).reshape(1, -1)
        yy = np.ascontiguousarray(y, dtype=np.float32)[np.newaxis]
        mono = (_mono_rows(x1), _mono_rows(x2))
        dev = eng.device
        up = lambda a: torch.from_numpy(a).to(dev)      # one synchronous upload per static field per model
        dc = DeviceContext(True, (up(x1), up(x2)), up(yy), None, mono, False, (x1, x2), y_batched=False)
        ext = ((float(x1.min()), float(x1.max())), (float(x2.min()), float(x2.max())))
        ent = (refs, dc, ext)
        self._static[key] = ent
        return ent

    def _seen_before(self, x, y) -> bool:
        """One-task batches cannot show that a field is shared: it counts as static from its second sighting on (the
        same buffers, still owned by the same live arrays), as in ``predict``'s context cache."""
        key = (_bufkey(x[0]), _bufkey(x[1]), _bufkey(y))
        owners = (_owner(x[0]), _owner(x[1]), _owner(y))
        if key in self._static:
            return True
        refs = self._seen.get(key)
        if refs is not None and all(r() is o for r, o in zip(refs, owners)):
            return True
        try:
            if len(self._seen) > 256:
                self._seen.clear()
            self._seen[key] = tuple(weakref.ref(o) for o in owners)
        except TypeError:
            pass
        return False

    def _coord_info(self, x1: np.ndarray, x2: np.ndarray):
        key = (x1.tobytes(), x2.tobytes())
        ent = self._coords.get(key)
        if ent is None:
            if len(self._coords) > 64:
                self._coords.clear()
            ent = self._coords[key] = ((x1.copy(), x2.copy()), (_mono_rows(x1), _mono_rows(x2)), {})
        return ent

    # ------------------------------------------------------------------------------------------
    def build(self, tasks: List[Task]) -> HostBatch:
        """list[Task] -> HostBatch in the next ring slot (waits for that slot's previous upload to finish)."""
        t_start = time.perf_counter()
        eng = self.engine
        slot = self.slots[self.n % len(self.slots)]
        self.n += 1
        if slot.event is not None:
            slot.event.synchronize()
            slot.event = None
        B = len(tasks)
        tasks = [t.remove_target_nans() if np.isnan(t["Y_t"][0]).any() else t for t in tasks]
        nts = {int(t["X_t"][0].shape[-1]) for t in tasks}
        if len(nts) != 1:
            raise ValueError("All tasks must have the same number of targets to concatenate: "
                             f"got {sorted(nts)}. Group tasks by number of targets first.")
        Nt = nts.pop()
        ctxs: List[DeviceContext] = []
        extents = []                     # per set: ((lo1, hi1), (lo2, hi2))
        for k in range(len(tasks[0]["X_c"])):
            xs = [t["X_c"][k] for t in tasks]
            ys = [t["Y_c"][k] for t in tasks]
            if isinstance(xs[0], tuple):
                x1s, x2s = [x[0] for x in xs], [x[1] for x in xs]
                if same_buffer(ys) and same_buffer(x1s) and same_buffer(x2s) and (B > 1 or self._seen_before(xs[0], ys[0])):
                    ent = self._static_set(xs[0], ys[0])
                    if ent is not None:
                        ctxs.append(ent[1])
                        extents.append(ent[2])
                        continue
                C, N1, N2 = ys[0].shape
                ybuf = slot.get(f"y{k}", (B, C, N1, N2))
                ynp = ybuf.numpy()
                for b, y in enumerate(ys):
                    ynp[b] = y                                   # cast + copy straight into pinned memory
                shared = same_buffer(x1s) and same_buffer(x2s) or all(
                    np.array_equal(x1s[0], a) and np.array_equal(x2s[0], c) for a, c in zip(x1s[1:], x2s[1:]))
                nb = 1 if shared else B
                x1b, x2b = slot.get(f"x1_{k}", (nb, N1)), slot.get(f"x2_{k}", (nb, N2))
                for b in range(nb):
                    x1b.numpy()[b] = np.asarray(x1s[b]).reshape(-1)
                    x2b.numpy()[b] = np.asarray(x2s[b]).reshape(-1)
                x_host, mono, band = self._coord_info(x1b.numpy(), x2b.numpy())
                ctxs.append(DeviceContext(True, (x1b, x2b), ybuf, None, mono, not shared, x_host, band))
                extents.append(((float(x_host[0].min()), float(x_host[0].max())),
                                (float(x_host[1].min()), float(x_host[1].max()))))
            else:
                C = ys[0].shape[0]
                n = max(int(x.shape[-1]) for x in xs)
                xbuf, ybuf = slot.get(f"x{k}", (B, 2, n)), slot.get(f"y{k}", (B, C, n))
                xnp, ynp = xbuf.numpy(), ybuf.numpy()
                for b, (x, y) in enumerate(zip(xs, ys)):
                    m = int(x.shape[-1])
                    xnp[b, :, :m], ynp[b, :, :m] = x, y
                    if m < n:                                    # merge_contexts padding: x = 0, y masked out
                        xnp[b, :, m:], ynp[b, :, m:] = 0.0, np.nan
                ctxs.append(DeviceContext(False, xbuf, ybuf, None))
                extents.append(None if n == 0 else ((float(xnp[:, 0].min()), float(xnp[:, 0].max())),
                                                     (float(xnp[:, 1].min()), float(xnp[:, 1].max()))))
        Ct, Ca = tasks[0]["Y_t"][0].shape[0], tasks[0]["Y_t_aux"].shape[0]
        xt, yt, aux = slot.get("xt", (B, 2, Nt)), slot.get("yt", (B, Ct, Nt)), slot.get("aux", (B, Ca, Nt))
        xtn, ytn, auxn = xt.numpy(), yt.numpy(), aux.numpy()
        for b, t in enumerate(tasks):
            xtn[b], ytn[b], auxn[b] = t["X_t"][0], t["Y_t"][0], t["Y_t_aux"]
        if Nt:
            extents.append(((float(xtn[:, 0].min()), float(xtn[:, 0].max())), (float(xtn[:, 1].min()), float(xtn[:, 1].max()))))
        # the discretisation only needs the extents: hand it 2-point stand-ins
        cfg = eng.cfg
        grid = discretise([np.array([[e[0][0], e[0][1]], [e[1][0], e[1][1]]], dtype=np.float64)
                           for e in extents if e is not None], cfg.points_per_unit, cfg.margin, cfg.grid_multiple)
        hb = HostBatch(ctxs, xt, yt, aux, grid, B)
        hb._slot = slot
        self.host_ms.append((time.perf_counter() - t_start) * 1e3)
        return hb

    def upload(self, hb: HostBatch, stream: Optional[torch.cuda.Stream] = None) -> DeviceBatch:
        db = self.engine.upload(hb, stream=stream)
        self.h2d_bytes = db.h2d_bytes
        slot = getattr(hb, "_slot", None)
        if slot is not None and torch.cuda.is_available():
            slot.event = torch.cuda.Event()
            slot.event.record(stream if stream is not None else torch.cuda.current_stream())
        return db
