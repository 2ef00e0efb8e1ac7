"""Forward / backward orchestration of the ConvNP hot path over libconvnp_b200.so.

One ``Engine`` per model.  It owns the device workspaces, launches the C-ABI kernels on the current
CUDA stream in the order of upstream ``nps.Model.__call__`` + ``nps.loglik``:

    SetConv encoder (1) -> UNet (2) -> SetConv decoder (3) -> aux MLP + Gaussian head + NLL (4)

and runs the hand-written backward of (4), (3) and (2) (the encoder has no trainable inputs).
Two numeric modes:
  * ``fp32``  : everything fp32 on NCHW tensors (CUDA-core convolutions) -- the 1e-5 parity mode;
  * ``bf16``  : UNet activations/weights in bf16 on the blocked layout, tcgen05 implicit-GEMM
                convolutions with fp32 accumulation; encoder, decoder, head stay fp32.
PyTorch is used for device memory, streams and the tiny [B]-sized loss algebra only.
"""
from __future__ import annotations

import collections
import ctypes as C
import math
import os
import weakref
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple, Union

import numpy as np
import torch

from . import _cabi
from ._cabi import CnpBlk, CnpConvOut, CnpMlpParams
from .discretisation import GridSpec, discretise
from .model import ConvNPConfig, ConvNPModule
from .task import is_batch_broadcast


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


@dataclass
class DeviceContext:
    gridded: bool
    x: Union[torch.Tensor, Tuple[torch.Tensor, torch.Tensor]]  # off-grid [B,2,N] | (x1 [Bx,N1], x2 [Bx,N2])
    y: torch.Tensor                                             # [B,C,N] | [B,C,N1,N2]  (may hold NaN)
    mask: Optional[torch.Tensor]                                # [B,1,N] | [B,1,N1,N2] | None
    mono: Tuple[int, int] = (0, 0)
    x_batched: bool = False
    x_host: Optional[Tuple[np.ndarray, np.ndarray]] = None   # gridded: host copy of the coordinates
    band_cache: dict = field(default_factory=dict)           # (grid, scale) -> band hint, shared with uploads
    y_batched: bool = True    # False: every task of the batch carries the same field (static topography / land mask):
    #                           one slice is staged, uploaded and encoded, the encoder output is broadcast


@dataclass
class DeviceBatch:
    """One (possibly concatenated) task resident in HBM."""
    contexts: List[DeviceContext]
    xt: torch.Tensor            # [B,2,Nt]
    yt: Optional[torch.Tensor]  # [B,1,Nt]
    aux_t: Optional[torch.Tensor]  # [B,Ca,Nt]
    grid: GridSpec
    B: int
    Nt: int
    h2d_bytes: int = 0
    ready: Optional[torch.cuda.Event] = None   # set when the upload ran on a side stream (Engine.upload(stream=...))
    xt_host: Optional[Tuple[np.ndarray, np.ndarray]] = None   # on-grid targets: host copy of the coordinates


@dataclass
class HostBatch:
    """A task staged on the host as float32 (optionally pinned) tensors, ready for ``Engine.upload``."""
    contexts: List[DeviceContext]
    xt: torch.Tensor
    yt: Optional[torch.Tensor]
    aux_t: Optional[torch.Tensor]
    grid: GridSpec
    B: int
    xt_host: Optional[Tuple[np.ndarray, np.ndarray]] = None


def _monotone(v: np.ndarray) -> int:
    d = np.diff(v.astype(np.float64))
    if d.size == 0 or np.all(d > 0):
        return 1
    if np.all(d < 0):
        return -1
    return 0


def _mono_rows(rows: np.ndarray) -> int:
    kinds = {_monotone(r) for r in rows}
    return kinds.pop() if len(kinds) == 1 else 0


class _Blk:
    """A blocked bf16 activation buffer [B][CB][H+4][W+4][8] with a zeroed 2-pixel pad."""

    def __init__(self, B: int, CB: int, H: int, W: int, device):
        self.B, self.CB, self.H, self.W = B, CB, H, W
        self.plane = (H + 4) * (W + 4) * 8
        self.bstride = CB * self.plane
        slack = (16 * (W + 4) + 512) * 8  # tile over-reads past the last plane stay inside the allocation
        self.t = torch.zeros(B * self.bstride + slack, dtype=torch.bfloat16, device=device)

    def view(self, cb_off: int = 0, b_off: int = 0) -> CnpBlk:
        """View from chunk ``cb_off`` (and image ``b_off``) on."""
        return CnpBlk(self.t.data_ptr() + 2 * b_off * self.bstride, self.bstride, cb_off, self.H, self.W)

    def zero_(self):
        self.t.zero_()


class Engine:
    def __init__(self, module: ConvNPModule, precision: str = "fp32"):
        if precision not in ("fp32", "bf16"):
            raise ValueError("precision must be 'fp32' or 'bf16'")
        self.module = module
        self.cfg: ConvNPConfig = module.cfg
        self.precision = precision
        if self.cfg.dim_aux_t <= 0:
            raise NotImplementedError("the hot path expects aux-at-target channels (dim_aux_t > 0), as the "
                                      "reference always configures (train.py:160-166)")
        if precision == "bf16" and any(c != 64 for c in self.cfg.unet_channels):
            # the tcgen05 kernels tile exactly 64 output channels (the reference's default, config.py:2685-2689);
            # other widths (train_downscaling.py:116-117 lets the user pick them) run the fp32 CUDA-core kernels
            import warnings
            warnings.warn(f"unet_channels={self.cfg.unet_channels}: the bf16 tensor-core path is specialised for 64 "
                          "channels per level; this model runs the fp32 kernels (precision='fp32')", stacklevel=3)
            self.precision = precision = "fp32"
        self._ws: Dict[tuple, object] = {}
        self._sig_lru: "collections.OrderedDict" = collections.OrderedDict()    # batch signature -> workspace keys
        self._pinned_sigs: set = set()
        self._cur_sig = None
        self._packed: Dict[str, Tuple[int, torch.Tensor]] = {}
        self._scale_cache: Dict[int, Tuple[int, float]] = {}
        self._enc_tabs: Dict[tuple, tuple] = {}      # band tables of the fused encoder, per (coordinates, grid, scale)
        self._pack_reqs: Dict[str, tuple] = {}     # every packing seen so far -> re-issued up front on a side stream
        self._wp_ver: Dict[str, tuple] = {}        # polyphase weights per level: (weight version, forced-pack epoch)
        self._pack_epoch = 0
        self._pack_stream = None
        self._wgrad_stream = None
        self._pack_event = None
        self._pack_event_bwd = None
        self._side_streams: List[torch.cuda.Stream] = []
        self.allreduce_group = None   # set by dist.enable_data_parallel
        self.world_size = 1
        self.launches = 0
        self._prof = None
        self.force_pack = False       # graph capture: re-pack every bf16 weight inside the captured region
        self.packs_recorded = 0
        self.generation = 0           # forward counter: a backward whose activations were overwritten must not run
        # bf16 copies of the weights are re-packed when a master's _version changed -- and, because fused optimisers
        # and CUDA-graph replays update the masters WITHOUT bumping _version (measured: torch.optim.AdamW(fused=True,
        # capturable=True) leaves it untouched), whenever a backward has run since the last packing
        self.weights_dirty = False

    # ------------------------------------------------------------------------------------------
    # helpers
    # ------------------------------------------------------------------------------------------
    @property
    def device(self):
        return self.module.decoder.unet.initial_linear.weight.device

    def _require_cuda(self):
        if self.device.type != "cuda":
            raise _cabi.CnpError("ConvNP parameters are not on a CUDA device: the hot path has no CPU fallback "
                                 "(call set_gpu_default_device() before building the model, or model.model.cuda())")
        _cabi.lib()

    def _call(self, name, *args, work=None):
        """Launch one C-ABI kernel.  ``work`` = (algorithmic flops, algorithmic bytes) for the roofline;
        with profiling on, every launch is bracketed by CUDA events on the launching stream."""
        self.launches += 1
        if self._prof is None:
            _cabi.call(name, *args)
            return
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _cabi.call(name, *args)
        e1.record()
        self._prof.append((name, e0, e1, work or (0.0, 0.0)))

    def profile_start(self):
        self._prof = []

    def profile_stop(self) -> Dict[str, dict]:
        """Per-kernel totals: launches, ms, algorithmic flops / bytes."""
        torch.cuda.synchronize()
        out: Dict[str, dict] = {}
        for name, e0, e1, (fl, by) in self._prof or []:
            d = out.setdefault(name, dict(launches=0, ms=0.0, flops=0.0, bytes=0.0))
            d["launches"] += 1
            d["ms"] += e0.elapsed_time(e1)
            d["flops"] += fl
            d["bytes"] += by
        self._prof = None
        return out

    # ---- workspace lifetime: the reference steps many batch shapes (one group per station count, plus remainders:
    # train.py:448-475) and at internal_density 500 every shape owns several GB of activations.  Workspaces are tagged
    # with the batch signatures that use them; beyond MAX_SIGNATURES live shapes the least recently used one is released
    # (never one a captured CUDA graph replays from: graphs hold raw pointers).
    MAX_SIGNATURES = int(os.environ.get("CONVNP_B200_MAX_SHAPES", "4"))

    def _touch_signature(self, sig) -> None:
        self._cur_sig = sig
        lru = self._sig_lru
        if sig in lru:
            lru.move_to_end(sig)
            return
        lru[sig] = set()
        free = [s for s in lru if s not in self._pinned_sigs and s != sig]
        while len(lru) - len(self._pinned_sigs & set(lru)) > self.MAX_SIGNATURES and free:
            old = free.pop(0)
            keys = lru.pop(old)
            still = set().union(*lru.values()) if lru else set()
            for k in keys - still:
                self._ws.pop(k, None)

    def pin_signature(self) -> None:
        """The current batch shape is being captured into a CUDA graph: its workspaces must outlive the graph."""
        if self._cur_sig is not None:
            self._pinned_sigs.add(self._cur_sig)

    def _tag(self, k) -> None:
        if self._cur_sig is not None:
            self._sig_lru.setdefault(self._cur_sig, set()).add(k)

    def _buf(self, key, shape, dtype=torch.float32, zero=False):
        k = (key, tuple(shape), dtype)
        self._tag(k)
        t = self._ws.get(k)
        if t is None:
            t = torch.zeros(shape, dtype=dtype, device=self.device) if zero else \
                torch.empty(shape, dtype=dtype, device=self.device)
            self._ws[k] = t
        return t

    def _blk(self, key, B, CB, H, W) -> _Blk:
        k = (key, B, CB, H, W, "blk")
        self._tag(k)
        t = self._ws.get(k)
        if t is None:
            t = _Blk(B, CB, H, W, self.device)
            self._ws[k] = t
        return t

    def release_workspaces(self):
        self._ws.clear()
        self._sig_lru.clear()
        self._pinned_sigs.clear()
        self._packed.clear()
        self._pack_reqs.clear()
        self._enc_tabs.clear()

    def _scale2(self, log_scale: torch.Tensor) -> float:
        """exp(2 log_scale) as an fp32-rounded host float.  The length scales are fixed (not learnable), so the value
        is read back from the device once per parameter version -- never inside a step (CUDA-graph capture)."""
        ent = self._scale_cache.get(id(log_scale))
        if ent is None or ent[0] != log_scale._version:
            ent = (log_scale._version, float(np.float32(math.exp(2.0 * float(log_scale)))))
            self._scale_cache[id(log_scale)] = ent
        return ent[1]

    # ------------------------------------------------------------------------------------------
    # host -> device
    # ------------------------------------------------------------------------------------------
    def stage_host(self, contexts, xt, yt, aux_t, pinned: bool = False, ctx_cache: Optional[dict] = None) -> "HostBatch":
        """contexts: list of (x, y, mask|None) numpy/torch CPU arrays with a leading batch axis.
        NaNs may stay in ``y`` (the kernels derive validity on the fly).  Returns float32 contiguous
        CPU tensors (page-locked when ``pinned``) plus the host-side discretisation.

        ``ctx_cache`` (predict over many tasks): gridded context sets whose arrays are the same host buffers as in an
        earlier task (static topography / land mask) are uploaded once (on their second sighting: per-date fields
        never enter the cache) and their device copy -- with its band hint -- is reused; keyed on (data pointer,
        shape) of x, y and mask."""
        def bufkey(a):
            if a is None:
                return None
            a = np.asarray(a)
            return (a.__array_interface__["data"][0], a.shape, a.strides, str(a.dtype))
        def owner(a):
            a = np.asarray(a)
            while isinstance(a.base, np.ndarray):
                a = a.base
            return a

        def host(a):
            return a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)

        def cpu(a):
            if a is None:
                return None
            t = a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(a))
            t = t.to(torch.float32).contiguous()
            if pinned and t.device.type == "cpu" and not t.is_pinned():
                t = t.pin_memory()
            return t

        xs = [c[0] for c in contexts] + [xt]
        xs = [tuple(host(v) for v in x) if isinstance(x, tuple) else host(x) for x in xs]
        grid = discretise(xs, self.cfg.points_per_unit, self.cfg.margin, self.cfg.grid_multiple)
        B = int((xs[-1][0] if isinstance(xs[-1], tuple) else xs[-1]).shape[0])
        if isinstance(xt, tuple):
            x1t, x2t = (v.reshape(v.shape[0], -1) for v in xs[-1])
            if not (np.all(x1t == x1t[:1]) and np.all(x2t == x2t[:1])):
                raise NotImplementedError("on-grid targets must share coordinates across the batch")
            B = int(np.asarray(contexts[0][1].y if hasattr(contexts[0][1], "y") else contexts[0][1]).shape[0])
            xt_h = (cpu(x1t[0]), cpu(x2t[0]))
            xt_np = (np.array(x1t[0], dtype=np.float32), np.array(x2t[0], dtype=np.float32))
        else:
            xt_h = xt_np = None
        hctx = []
        for (x, y, m), xh in zip(contexts, xs[:-1]):
            key = None
            if ctx_cache is not None and isinstance(x, tuple) and not isinstance(y, torch.Tensor):
                key = (bufkey(x[0]), bufkey(x[1]), bufkey(y), bufkey(m))
                bases = tuple(owner(a) for a in (x[0], x[1], y, m) if a is not None)
                ent = ctx_cache.get(key)
                # an entry only counts while the arrays that own its buffers are still the same live objects: a freed
                # temporary (e.g. the float32 copy of a float64 field) can hand its address to the next task's data
                alive = ent is not None and len(ent[0]) == len(bases) and all(r() is b for r, b in zip(ent[0], bases))
                if alive and ent[1] is not None:
                    hctx.append(ent[1])
                    continue
                if not alive:              # first sighting: a per-date field until the same buffers come back
                    try:
                        ctx_cache[key] = (tuple(weakref.ref(b) for b in bases), None)
                    except TypeError:
                        ctx_cache.pop(key, None)
                    key = None
            if isinstance(x, tuple):
                x1h, x2h = xh
                x1h = x1h.reshape(x1h.shape[0], -1)
                x2h = x2h.reshape(x2h.shape[0], -1)
                shared = all(np.array_equal(x1h[0], x1h[i]) and np.array_equal(x2h[0], x2h[i])
                             for i in range(1, x1h.shape[0]))
                mono = (_mono_rows(x1h), _mono_rows(x2h))
                if shared:
                    x1h, x2h = x1h[:1], x2h[:1]
                # static fields (topography aux, land mask) are identical in every task of a batch: keep one slice
                yh = host(y)
                mh = None if m is None else host(m)
                if is_batch_broadcast(yh) and (mh is None or is_batch_broadcast(mh)):
                    y_shared = shared            # concat_tasks stacked ONE buffer (zero-stride view): nothing to compare
                else:
                    # separate copies of the same field: a strided sample rules out per-date fields before the full compare
                    def same(a):
                        flat = a.reshape(a.shape[0], -1)
                        smp = flat[:, ::max(1, flat.shape[1] // 4096)]
                        return all(np.array_equal(smp[0], smp[i], equal_nan=True) for i in range(1, a.shape[0])) and \
                            all(np.array_equal(a[0], a[i], equal_nan=True) for i in range(1, a.shape[0]))
                    y_shared = shared and yh.shape[0] > 1 and same(yh) and (mh is None or same(mh))
                if y_shared:
                    y, m = yh[:1], (None if mh is None else mh[:1])
                hc = DeviceContext(True, (cpu(x1h), cpu(x2h)), cpu(y), cpu(m), mono, not shared, (x1h, x2h),
                                   y_batched=not y_shared)
                if key is not None:       # keep the DEVICE copy: later tasks skip staging and H2D of this set
                    self._require_cuda()
                    dev = self.device
                    up1 = lambda t: None if t is None else t.to(dev, non_blocking=True)
                    hc = DeviceContext(True, tuple(up1(v) for v in hc.x), up1(hc.y), up1(hc.mask), hc.mono, hc.x_batched,
                                       hc.x_host, hc.band_cache, hc.y_batched)
                    ctx_cache[key] = (ctx_cache[key][0], hc)
                hctx.append(hc)
            else:
                hctx.append(DeviceContext(False, cpu(x), cpu(y), cpu(m)))
        return HostBatch(hctx, xt_h if xt_h is not None else cpu(xt), cpu(yt), cpu(aux_t), grid, B, xt_np)

    def upload(self, hb: "HostBatch", stream: Optional[torch.cuda.Stream] = None) -> DeviceBatch:
        """Asynchronous H2D of a staged batch on the current stream, or on ``stream`` (a copy stream: the batch
        then carries a ``ready`` event that ``forward`` waits on, so the copy of step i+1 overlaps step i)."""
        if stream is not None:
            with torch.cuda.stream(stream):
                db = self.upload(hb)
                db.ready = torch.cuda.Event()
                db.ready.record(stream)
            return db
        self._require_cuda()
        dev = self.device
        nbytes = 0

        def up(t):
            nonlocal nbytes
            if t is None:
                return None
            if t.device.type == "cuda":
                return t
            nbytes += t.numel() * t.element_size()
            return t.to(dev, non_blocking=True)

        dctx = []
        for c in hb.contexts:
            x = tuple(up(v) for v in c.x) if c.gridded else up(c.x)
            dctx.append(DeviceContext(c.gridded, x, up(c.y), up(c.mask), c.mono, c.x_batched, c.x_host, c.band_cache,
                                      c.y_batched))
        if isinstance(hb.xt, tuple):
            xt = tuple(up(v) for v in hb.xt)
            nt = int(xt[0].shape[-1]) * int(xt[1].shape[-1])
        else:
            xt = up(hb.xt)
            nt = int(xt.shape[-1])
        yt, aux = up(hb.yt), up(hb.aux_t)
        return DeviceBatch(dctx, xt, yt, aux, hb.grid, hb.B, nt, nbytes, xt_host=hb.xt_host)

    @staticmethod
    def _batch_tensors(batch: DeviceBatch):
        for c in batch.contexts:
            for t in (c.x if isinstance(c.x, tuple) else (c.x,)) + (c.y, c.mask):
                if t is not None:
                    yield t
        for t in (batch.xt if isinstance(batch.xt, tuple) else (batch.xt,)) + (batch.yt, batch.aux_t):
            if t is not None:
                yield t

    def prepare(self, contexts, xt, yt, aux_t, pinned: bool = False) -> DeviceBatch:
        return self.upload(self.stage_host(contexts, xt, yt, aux_t, pinned=pinned))

    # ------------------------------------------------------------------------------------------
    # (1) encoder
    # ------------------------------------------------------------------------------------------
    def encode(self, batch: DeviceBatch, broadcast: bool = True) -> torch.Tensor:
        """SetConv encoders of all context sets into disjoint channel ranges of one [B, Cin, n1, n2] tensor.  The sets
        are independent and individually too small to fill the GPU (a static field is encoded once, B = 1), so each
        runs on its own side stream (fork / join on events; serial when profiling or with CNP_NO_MULTISTREAM)."""
        cfg, g, B = self.cfg, batch.grid, batch.B
        enc = self._buf("enc", (B, cfg.in_channels, g.n1, g.n2))
        # broadcast=False: channels of context sets shared by the whole batch are written for task 0 only and flagged in
        # self._enc_shared_mask; the consumer (the blocked conversion of the folded first layer) broadcasts on the fly
        self._enc_shared_mask = 0
        self._enc_broadcast = broadcast
        plan = self._fused_plan(batch)
        if plan is not None and self._encode_fused(batch, plan, enc=enc):
            return enc
        multi = len(batch.contexts) > 1 and self._prof is None and not os.environ.get("CNP_NO_MULTISTREAM")
        main = torch.cuda.current_stream()
        if multi:
            if len(self._side_streams) < 4:
                self._side_streams = [torch.cuda.Stream() for _ in range(4)]
            fork = torch.cuda.Event()
            fork.record(main)
        ch = 0
        for k, c in enumerate(batch.contexts):
            Ck = cfg.dim_yc[k]
            if c.y.shape[1] != Ck:
                raise ValueError(f"context set {k}: expected {Ck} channels, got {c.y.shape[1]}")
            if multi:
                ss = self._side_streams[k % 4]
                ss.wait_event(fork)
                with torch.cuda.stream(ss):
                    self._encode_set(batch, k, c, enc, ch)
            else:
                self._encode_set(batch, k, c, enc, ch)
            ch += Ck + 1
        if multi:
            for ss in self._side_streams[:min(4, len(batch.contexts))]:
                main.wait_stream(ss)
        return enc

    # ---- fused encoder (enc_fused.cu): all context sets of a task in one launch ----
    def _fused_plan(self, batch: DeviceBatch):
        """Which context sets the fused encoder takes per task and which are encoded once per batch first (sets every
        task shares); None when a set needs the generic kernels (unsorted or per-task coordinates, bands wider than 32
        inputs, more than 8 channels)."""
        if os.environ.get("CNP_NO_ENC_FUSED") or len(batch.contexts) > 8:
            return None
        cfg, g, B = self.cfg, batch.grid, batch.B
        per_task, static = [], []
        ch = 0
        for k, c in enumerate(batch.contexts):
            Ck = cfg.dim_yc[k]
            if c.y.shape[1] != Ck:
                raise ValueError(f"context set {k}: expected {Ck} channels, got {c.y.shape[1]}")
            if Ck > 8 or Ck < 1:
                return None
            s2 = self._scale2(self.module.encoder.set_convs[k].log_scale)
            if c.gridded:
                if c.x_batched or c.mono[0] == 0 or c.mono[1] == 0 or c.x_host is None or not self._band_hint(c, g, s2):
                    return None
                (static if (not c.y_batched and B > 1) else per_task).append((k, c, ch, s2))
            else:
                per_task.append((k, c, ch, s2))
            ch += Ck + 1
        return dict(per_task=per_task, static=static)

    def _enc_tables(self, c: DeviceContext, g: GridSpec, scale2: float):
        """Band tables of a gridded set (band starts, lengths and SetConv weights per internal-grid row / column).  They
        depend only on (coordinates, grid, length scale) -- all fixed during training -- so they are built once
        (cnp_encode_tables) and kept on the device, keyed by the coordinate bytes."""
        key = (c.x_host[0].tobytes(), c.x_host[1].tobytes(), g, scale2)
        ent = self._enc_tabs.get(key)
        if ent is None:
            if len(self._enc_tabs) > 64:
                self._enc_tabs.clear()
            band = self._band_hint(c, g, scale2)
            ti = torch.empty(2 * (g.n1 + g.n2), dtype=torch.int32, device=self.device)
            tw = torch.empty(band * (g.n1 + g.n2), dtype=torch.float32, device=self.device)
            x1, x2 = c.x
            self._call("cnp_encode_tables", _ptr(x1), _ptr(x2), int(x1.shape[-1]), int(x2.shape[-1]), c.mono[0], c.mono[1],
                       g.start1, g.n1, g.start2, g.n2, g.res, scale2, band, _ptr(ti), _ptr(tw), _stream())
            ent = (band, ti, tw)
            if not torch.cuda.is_current_stream_capturing():     # (tensors made during a capture live in the graph's pool)
                self._enc_tabs[key] = ent
        return ent

    def _enc_sets(self, entries, grid: GridSpec, B: int, enc: Optional[torch.Tensor]):
        """The C-ABI description of the context sets.  ``enc`` given (fp32 mode): the V planes of the gridded sets are the
        channels of ``enc`` themselves; else they live in a workspace the assembly kernel gathers from."""
        K = _cabi
        sets = K.CnpEncSets()
        t_bytes = v_bytes = 0
        plane = grid.n1 * grid.n2
        for n, (k, c, ch, s2) in enumerate(entries):
            e = sets.s[n]
            Ck = self.cfg.dim_yc[k]
            e.kind, e.C, e.ch_off, e.scale2 = (1 if c.gridded else 0), Ck, ch, s2
            if c.gridded:
                e.batched = int(c.y_batched and B > 1)
                nb = B if e.batched else 1
                e.N1, e.N2 = int(c.x[0].shape[-1]), int(c.x[1].shape[-1])
                e.mono1, e.mono2 = c.mono
                e.KB, ti, tw = self._enc_tables(c, grid, s2)
                e.tab_i, e.tab_w = ti.data_ptr(), tw.data_ptr()
                T = self._buf(f"enc_T{k}", (nb, Ck + 1, e.N1, grid.n2))
                e.T = T.data_ptr()
                if enc is not None:
                    e.V, e.V_bs = enc.data_ptr() + 4 * ch * plane, enc.stride(0)
                else:
                    V = self._buf(f"enc_V{k}", (nb, Ck + 1, grid.n1, grid.n2))
                    e.V, e.V_bs = V.data_ptr(), V.stride(0)
                t_bytes += 4 * T.numel()
                v_bytes += 4 * nb * (Ck + 1) * plane
            else:
                e.batched = 1
                e.x1, e.N1 = _ptr(c.x), int(c.x.shape[-1])
            e.y = _ptr(c.y)
            e.mask = _ptr(c.mask)
        sets.n_sets = len(entries)
        return sets, t_bytes, v_bytes

    def _encode_fused(self, batch: DeviceBatch, plan: dict, enc: Optional[torch.Tensor] = None,
                      blk: Optional["_Blk"] = None) -> bool:
        """Launch the fused encoder (enc_fused.cu) into ``enc`` (fp32 NCHW) or ``blk`` (blocked bf16 + constant-1 channel):
        horizontal pass and vertical pass of every gridded set, then the off-grid sets + assembly per tile."""
        cfg, g, B = self.cfg, batch.grid, batch.B
        Cin = cfg.in_channels
        n_in = lambda c: 4.0 * (c.y.numel() + (c.mask.numel() if c.mask is not None else 0))
        S = _stream()
        entries = sorted(plan["per_task"] + plan["static"], key=lambda t: t[0])
        sets, t_bytes, v_bytes = self._enc_sets(entries, g, B, enc)
        by_grid = sum(n_in(c) for _, c, _, _ in entries if c.gridded)
        by_off = sum(n_in(c) for _, c, _, _ in entries if not c.gridded)
        if any(c.gridded for _, c, _, _ in entries):
            self._call("cnp_encode_hpass", C.byref(sets), B, g.n1, g.n2, S, work=(0.0, by_grid + t_bytes))
            self._call("cnp_encode_vpass", C.byref(sets), B, g.n1, g.n2, cfg.epsilon, S, work=(0.0, t_bytes + v_bytes))
        if blk is not None:
            bv = blk.view()
            self._call("cnp_encode_fused", C.byref(sets), B, g.start1, g.n1, g.start2, g.n2, g.res, cfg.epsilon, 1, None, 0,
                       Cin, C.byref(bv), blk.CB, S, work=(0.0, by_off + v_bytes + 2.0 * B * blk.CB * 8 * g.n1 * g.n2))
        else:
            n_og = sum(cfg.dim_yc[k] + 1 for k, c, _, _ in entries if not c.gridded)
            self._call("cnp_encode_fused", C.byref(sets), B, g.start1, g.n1, g.start2, g.n2, g.res, cfg.epsilon, 0,
                       _ptr(enc), enc.stride(0), Cin, None, 0, S, work=(0.0, by_off + 4.0 * B * n_og * g.n1 * g.n2))
            for k, c, ch, _ in plan["static"]:      # a field every task shares was encoded for task 0: broadcast
                Ck = cfg.dim_yc[k]
                enc[1:, ch:ch + Ck + 1].copy_(enc[:1, ch:ch + Ck + 1].expand(B - 1, -1, -1, -1))
        return True

    def encode_blocked(self, batch: DeviceBatch) -> Optional["_Blk"]:
        """bf16 UNet with the folded first layer: the encoder writes the UNet's blocked bf16 input (with the constant-1
        channel) directly.  None when the batch needs the per-set kernels (then ``encode`` + the layout conversion run)."""
        plan = self._fused_plan(batch)
        if plan is None:
            return None
        cfg, g = self.cfg, batch.grid
        cb0 = 2 * ((cfg.in_channels + 1 + 15) // 16)
        blk = self._blk("x_aug", batch.B, cb0, g.n1, g.n2)
        return blk if self._encode_fused(batch, plan, blk=blk) else None

    def _encode_set(self, batch: DeviceBatch, k: int, c: DeviceContext, enc: torch.Tensor, ch: int) -> None:
        cfg, g, B = self.cfg, batch.grid, batch.B
        s2 = self._scale2(self.module.encoder.set_convs[k].log_scale)
        Ck = cfg.dim_yc[k]
        if c.gridded:
            x1, x2 = c.x
            Be = B if c.y_batched else 1     # a field shared by the whole batch is encoded once ...
            by = 4.0 * (c.y.numel() + (c.mask.numel() if c.mask is not None else 0) + Be * (Ck + 1) * g.n1 * g.n2)
            N1, N2 = int(x1.shape[-1]), int(x2.shape[-1])
            band = self._band_hint(c, g, s2)
            ws, ws_bytes = None, 0
            if band:
                ws_bytes = _cabi.lib().cnp_setconv_enc_grid_workspace_bytes(Be, Ck, N1, g.n1, g.n2, band)
                ws = self._buf(f"enc_ws{k}", ((ws_bytes + 3) // 4,))
            self._call("cnp_setconv_enc_grid_fwd", _ptr(x1), _ptr(x2), int(c.x_batched), _ptr(c.y), _ptr(c.mask),
                       Be, Ck, N1, N2, c.mono[0], c.mono[1],
                       g.start1, g.n1, g.start2, g.n2, g.res, s2, cfg.epsilon, _ptr(enc), ch, cfg.in_channels,
                       band, _ptr(ws), ws_bytes, _stream(), work=(0.0, by))
            if Be < B:                        # ... and its channels broadcast to the other tasks
                if self._enc_broadcast:       # (plain D2D copy, or on the fly by the consumer)
                    enc[1:, ch:ch + Ck + 1].copy_(enc[:1, ch:ch + Ck + 1].expand(B - 1, -1, -1, -1))
                else:
                    self._enc_shared_mask |= ((1 << (Ck + 1)) - 1) << ch
        else:
            by = 4.0 * (c.x.numel() + c.y.numel() + B * (Ck + 1) * g.n1 * g.n2)
            self._call("cnp_setconv_enc_offgrid_fwd", _ptr(c.x), _ptr(c.y), _ptr(c.mask), B, Ck,
                       int(c.x.shape[-1]), g.start1, g.n1, g.start2, g.n2, g.res, s2, cfg.epsilon, _ptr(enc), ch,
                       cfg.in_channels, _stream(), work=(0.0, by))

    @staticmethod
    def _band_hint(c: DeviceContext, g: GridSpec, scale2: float) -> int:
        """Max number of inputs inside the truncation radius of any internal-grid point (both dims),
        computed on the host from the coordinates; 0 = use the generic kernel."""
        if c.x_host is None or c.mono[0] == 0 or c.mono[1] == 0:
            return 0
        key = (g, scale2)
        if key in c.band_cache:
            return c.band_cache[key]
        R = float(np.sqrt(np.float32(2.0 * 104.0) * np.float32(scale2))) * (1.0 + 1e-6)
        best = 0
        for d, xs in enumerate(c.x_host):
            gp = g.points(d).astype(np.float64)
            for row in xs:
                r = np.sort(row.astype(np.float64))
                cnt = np.searchsorted(r, gp + R, side="right") - np.searchsorted(r, gp - R, side="left")
                best = max(best, int(cnt.max()))
        best += 1
        c.band_cache[key] = best if best <= 32 else 0
        return c.band_cache[key]

    # ------------------------------------------------------------------------------------------
    # (2) UNet, fp32 path
    # ------------------------------------------------------------------------------------------
    def _levels(self, n1, n2):
        res = []
        h, w = n1, n2
        for s in self.cfg.unet_strides:
            h, w = h // s, w // s
            res.append((h, w))
        return res

    def _conv_f32(self, x, w, b, y, k, stride, relu):
        Bn, Cin, H, W = x.shape
        self._call("cnp_conv2d_fwd_f32", _ptr(x), x.stride(0), _ptr(w), _ptr(b), _ptr(y), y.stride(0), Bn, Cin, H, W,
                   y.shape[1], k, stride, int(relu), _stream())

    def _unet_fwd_f32(self, enc: torch.Tensor, B: int, n1: int, n2: int) -> Tuple[torch.Tensor, dict]:
        cfg, u = self.cfg, self.module.decoder.unet
        ch, st = cfg.unet_channels, cfg.unet_strides
        L = len(ch)
        res = self._levels(n1, n2)
        A = {}
        h_init = self._buf("h_init", (B, ch[0], n1, n2))
        self._conv_f32(enc, u.initial_linear.weight, u.initial_linear.bias, h_init, 1, 1, False)
        A["h_init"] = h_init
        # skip buffers: cat[i] holds [hs[i] ; h from level i+1] for i < L-1, hs[L-1] alone
        cat = []
        for i in range(L):
            c = ch[i] if i == L - 1 else 2 * ch[i]
            cat.append(self._buf(f"cat{i}", (B, c, res[i][0], res[i][1])))
        A["cat"] = cat
        x = h_init
        for i in range(L):
            lyr = u.before_turn_layers[i]
            y = cat[i][:, :ch[i]]
            self._conv_f32(x, lyr.weight, lyr.bias, y, 5, st[i], True)
            x = y
        ups = [None] * L
        h = None
        for i in range(L - 1, -1, -1):
            inp = cat[i]
            if st[i] == 2:
                up = self._buf(f"up{i}", (B, inp.shape[1], 2 * res[i][0], 2 * res[i][1]))
                self._call("cnp_upsample2x_fwd_f32", _ptr(inp), inp.stride(0), _ptr(up), up.stride(0), B, inp.shape[1],
                           res[i][0], res[i][1], _stream())
                inp = up
            ups[i] = inp
            lyr = u.after_turn_layers[i]
            if i > 0:
                y = cat[i - 1][:, ch[i - 1]:]
            else:
                y = self._buf("h_last", (B, ch[0], n1, n2))
            self._conv_f32(inp, lyr.weight, lyr.bias, y, 5, 1, True)
            h = y
        A["ups"], A["h_last"] = ups, h
        z = self._buf("z", (B, cfg.unet_out_channels, n1, n2))
        self._conv_f32(h, u.final_linear.weight, u.final_linear.bias, z, 1, 1, False)
        return z, A

    def _unet_bwd_f32(self, dz: torch.Tensor, enc: torch.Tensor, A: dict, grads: Dict[str, torch.Tensor], B, n1, n2,
                      up_done=None):
        cfg, u = self.cfg, self.module.decoder.unet
        ch, st = cfg.unet_channels, cfg.unet_strides
        L = len(ch)
        res = self._levels(n1, n2)
        cat, ups = A["cat"], A["ups"]
        S = _stream()

        def wgrad(x, dy, name, k, stride):
            Bn, Cin, H, W = x.shape
            self._call("cnp_conv2d_wgrad_f32", _ptr(x), x.stride(0), _ptr(dy), dy.stride(0),
                       _ptr(grads[name + ".weight"]), _ptr(grads[name + ".bias"]), Bn, Cin, H, W, dy.shape[1], k,
                       stride, S)

        def dgrad(dy, w, dx, k, stride, accumulate):
            Bn, Cin, H, W = dx.shape
            self._call("cnp_conv2d_dgrad_f32", _ptr(dy), dy.stride(0), _ptr(w), _ptr(dx), dx.stride(0), Bn, Cin, H, W,
                       dy.shape[1], k, stride, int(accumulate), S)

        def relu_bwd(d, y):
            self._call("cnp_relu_bwd_f32", _ptr(d), d.stride(0), _ptr(y), y.stride(0), B,
                       d.shape[1] * d.shape[2] * d.shape[3], S)

        P = "decoder.unet."
        h_last = A["h_last"]
        wgrad(h_last, dz, P + "final_linear", 1, 1)
        d_h = self._buf("d_h_last", h_last.shape)
        dgrad(dz, u.final_linear.weight, d_h, 1, 1, False)
        relu_bwd(d_h, h_last)
        d_cat = [self._buf(f"d_cat{i}", cat[i].shape) for i in range(L)]
        dy = d_h
        for i in range(0, L):
            name = P + f"after_turn_layers.{i}"
            lyr = u.after_turn_layers[i]
            x_in = ups[i]
            wgrad(x_in, dy, name, 5, 1)
            if st[i] == 2:
                d_up = self._buf(f"d_up{i}", x_in.shape)
                dgrad(dy, lyr.weight, d_up, 5, 1, False)
                self._call("cnp_upsample2x_bwd_f32", _ptr(d_up), d_up.stride(0), _ptr(d_cat[i]), d_cat[i].stride(0), B,
                           d_cat[i].shape[1], res[i][0], res[i][1], 0, S)
            else:
                dgrad(dy, lyr.weight, d_cat[i], 5, 1, False)
            relu_bwd(d_cat[i], cat[i])
            if i < L - 1:
                dy = d_cat[i][:, ch[i]:]
        if up_done is not None:
            up_done()
        # down path, deepest first; gradients accumulate into the skip halves
        for i in range(L - 1, -1, -1):
            name = P + f"before_turn_layers.{i}"
            lyr = u.before_turn_layers[i]
            dy = d_cat[i][:, :ch[i]]
            x_in = cat[i - 1][:, :ch[i - 1]] if i > 0 else A["h_init"]
            wgrad(x_in, dy, name, 5, st[i])
            if i > 0:
                dx = d_cat[i - 1][:, :ch[i - 1]]
                tmp = self._buf(f"d_tmp{i}", dx.shape)
                dgrad(dy, lyr.weight, tmp, 5, st[i], False)
                relu_bwd(tmp, x_in)
                dx.add_(tmp)
            else:
                d_init = self._buf("d_h_init", x_in.shape)
                dgrad(dy, lyr.weight, d_init, 5, st[i], False)
                wgrad(enc, d_init, P + "initial_linear", 1, 1)

    # ------------------------------------------------------------------------------------------
    # (2) UNet, bf16 tensor-core path
    # ------------------------------------------------------------------------------------------
    def _prepack_all(self):
        """Re-pack every weight whose version changed (optimiser step) on a side stream at the start of the step, so
        the ~25 small packing launches overlap the encoder instead of sitting between the convolutions."""
        if self.force_pack or self.weights_dirty:
            self.packs_recorded = 0
            self._forced = set()
            self._pack_epoch += 1
        if os.environ.get("CNP_NO_PREPACK"):
            return
        force = self.force_pack or self.weights_dirty
        stale = [(k, r) for k, r in self._pack_reqs.items()
                 if self._packed.get(k) is not None and (force or self._packed[k][0] != self._pack_version(r))]
        if force:
            self._forced = set(k for k, _ in stale)
        if not stale:
            return
        if self._pack_stream is None:
            self._pack_stream = torch.cuda.Stream()
        main = torch.cuda.current_stream()
        self._pack_stream.wait_stream(main)
        # forward packings first, with their own event: the first convolution then waits for ~8 small launches (hidden
        # behind the encoder) instead of all ~18; the input-gradient packings are needed a millisecond later
        fwd = [(k, r) for k, r in stale if ".dg." not in k]
        bwd = [(k, r) for k, r in stale if ".dg." in k]
        with torch.cuda.stream(self._pack_stream):
            for k, r in fwd:
                self._do_pack(k, r, self._packed[k][1])
            self._pack_event = torch.cuda.Event()
            self._pack_event.record(self._pack_stream)
            for k, r in bwd:
                self._do_pack(k, r, self._packed[k][1])
            self._pack_event_bwd = torch.cuda.Event()
            self._pack_event_bwd.record(self._pack_stream)

    @staticmethod
    def _pack_version(req) -> tuple:
        w, fold = req[0], req[7]
        return (w._version,) if fold is None else (w._version, fold[0]._version, fold[1]._version)

    def _do_pack(self, key: str, req, buf: torch.Tensor) -> None:
        """Pack (and, for the first layer, fold the initial 1x1 into) one weight tensor on the current stream."""
        w, kind, n_chunks, py, px, co_off, n_out, fold, pre = req
        ver = self._pack_version(req)
        Cout, Cin, k, _ = w.shape
        src = w
        if pre == "tswap":        # column strips of the polyphase resize-convolution: the same layer, taps transposed
            src = self._buf(f"tsw.{key}", tuple(w.shape))
            src.copy_(w.detach().transpose(2, 3))
        elif pre == "phase":      # 4x4 phase weights [2][2][Cout][Cin][4][4] of Upsample(x2, bilinear) + Conv 5x5
            lk = key.split(".")[0]                       # the three packings of a level share one phase tensor
            wp = self._buf(f"wp.{lk}", (2, 2, Cout, Cin, 4, 4))
            if self._wp_ver.get(lk) != (ver, self._pack_epoch):
                self._call("cnp_up_phase_weights", _ptr(w), Cout, Cin, _ptr(wp), _stream())
                self._wp_ver[lk] = (ver, self._pack_epoch)
            src, k = (wp[py] if kind == _cabi.KIND_UP_PHASE else wp), 4
        if fold is not None:
            w1, b1 = fold
            Cp = n_chunks * 8
            src = self._buf(f"fold.{key}", (Cout, Cp, k, k))
            self._call("cnp_fold_in_fwd", _ptr(w), _ptr(w1), _ptr(b1), Cout, Cin, w1.shape[1], Cp, k, _ptr(src), _stream())
            Cin = Cp
        self._call("cnp_conv_tc2_pack", _ptr(src), Cout, Cin, k, kind, n_chunks, py, px, co_off, n_out, _ptr(buf),
                   _stream())
        self.packs_recorded += 1
        self._packed[key] = (ver, buf)

    def _packed_weights(self, key: str, w: torch.Tensor, kind: int, n_chunks: int, py=0, px=0, co_off=0,
                        n_out=64, fold=None, pre=None) -> torch.Tensor:
        """``fold=(W1, b1)``: ``w`` is the first 5x5 and the packed tensor is the folded 5x5 over [x ; 1] (fold_in.cu).
        ``pre``: "tswap" packs the tap-transposed weights, "phase" the polyphase weights derived from ``w``."""
        req = (w, kind, n_chunks, py, px, co_off, n_out, fold, pre)
        self._pack_reqs[key] = req
        if self._pack_event is not None:          # first consumer of the step: order after the side-stream packing
            torch.cuda.current_stream().wait_event(self._pack_event)
            self._pack_event = None
        if ".dg." in key and getattr(self, "_pack_event_bwd", None) is not None:
            torch.cuda.current_stream().wait_event(self._pack_event_bwd)
            self._pack_event_bwd = None
        ent = self._packed.get(key)
        force = self.force_pack or self.weights_dirty
        fresh = not force or key in getattr(self, "_forced", ())      # forced: already re-packed in this forward
        if ent is not None and ent[0] == self._pack_version(req) and ent[1].device == w.device and fresh:
            return ent[1]
        if force:
            self._forced.add(key)
        nbytes = _cabi.lib().cnp_conv_tc2_packed_bytes(kind, n_chunks, n_out)
        buf = ent[1] if ent is not None else torch.empty(nbytes // 2, dtype=torch.bfloat16, device=w.device)
        self._do_pack(key, req, buf)
        return buf

    def _conv_tc(self, x: CnpBlk, n_chunks, wpk, kind, out: CnpConvOut, B, py=0, px=0, n_out=64, wpk2=None, w2_from_b=0):
        K = _cabi
        if kind in (K.KIND_K5S1, K.KIND_K5S1_DGRAD):
            kdim = n_chunks * 8 * 25
        elif kind in (K.KIND_K1, K.KIND_K1_DGRAD):
            kdim = n_chunks * 8
        elif kind == K.KIND_K5S2:
            kdim = 64 * 25
        elif kind == K.KIND_UP_PHASE:            # both x-phases: 2 x 64 outputs per low-res pixel, 16 taps each
            kdim = 2 * n_chunks * 8 * 16
        elif kind == K.KIND_UP_PHASE_DGRAD:      # 4 phases x 64 channels x 16 taps
            kdim = 4 * 64 * 16
        else:
            kdim = 64 * (3 if py == 0 else 2) * (5 if px == 2 else (3 if px == 0 else 2))   # px = 2: both x-phases
        fl = 2.0 * B * x.H * x.W * n_out * kdim
        if wpk2 is not None:
            self._call("cnp_conv_tc2_w2", C.byref(x), n_chunks, _ptr(wpk), _ptr(wpk2), w2_from_b, kind, py, px, n_out,
                       C.byref(out), B, _stream(), work=(fl, 0.0))
            return
        self._call("cnp_conv_tc2", C.byref(x), n_chunks, _ptr(wpk), kind, py, px, n_out, C.byref(out), B, _stream(),
                   work=(fl, 0.0))

    @staticmethod
    def _out_blk(blk: CnpBlk, bias=None, relu=False, mask: Optional[CnpBlk] = None, accumulate=False,
                 scatter=(1, 0, 1, 0)) -> CnpConvOut:
        o = CnpConvOut()
        o.mode = 0
        o.blk = blk
        o.sy, o.ay, o.sx, o.ax = scatter
        o.bias = _ptr(bias)
        o.relu = int(relu)
        o.mask = C.pointer(mask) if mask is not None else None
        o.accumulate = int(accumulate)
        o._keep = (bias, mask)
        return o

    # ---- polyphase resize-convolution (up_poly.cu) -------------------------------------------
    @staticmethod
    def _up_poly_ok(x: "_Blk") -> bool:
        """128-channel decoder levels; CNP_NO_POLYPHASE=1 keeps Upsample + Conv.  The band costs ~0.4 ms of small launches
        per level whatever its size (a strip tile streams the whole packed weight tensor), the interior saves 36 % of a
        cost that grows with the pixel count: measured break-even near 110 x 110 low-res pixels at B = 16
        (152 -> 304: 0.25 ms saved, 76 -> 152: 0.3 ms lost).  CNP_POLYPHASE_MIN_PIXELS overrides the threshold (tests)."""
        if os.environ.get("CNP_NO_POLYPHASE") or x.CB != 16 or x.H < 12 or x.W < 12:
            return False
        return x.B * x.H * x.W >= int(os.environ.get("CNP_POLYPHASE_MIN_PIXELS", 16 * 110 * 110))

    def _up_strip_blks(self, key: str, B: int, CB: int, H: int, W: int):
        """Row / column strip tensors of a level whose LOW-res size is H x W (up_poly.cu layout) as a pair of
        (buffer, first image).  A square level keeps both in ONE tensor of 4B images -- rows first -- so that a single
        launch with two weight tensors (cnp_conv_tc2_w2) serves them."""
        if H == W:
            t = self._blk(f"{key}.rc", 4 * B, CB, 6, 2 * W)
            return (t, 0), (t, 2 * B)
        return (self._blk(f"{key}.rows", 2 * B, CB, 6, 2 * W), 0), (self._blk(f"{key}.cols", 2 * B, CB, 6, 2 * H), 0)

    def _strip_conv(self, src, dst, n_chunks, w, key, kind, n_out, B, bias=None, relu=False):
        """The standard 5x5 kernel (``kind``) on the row strips and, tap-transposed, on the column strips."""
        (sr, sr0), (sc, sc0) = src
        (dr, dr0), (dc, dc0) = dst
        dg = ".dg" if kind == _cabi.KIND_K5S1_DGRAD else ""
        wr = self._packed_weights(f"{key}{dg}.00" if dg else key, w, kind, n_chunks, 0, 0, 0, n_out)
        wc = self._packed_weights(f"{key}{dg}.t", w, kind, n_chunks, 0, 0, 0, n_out, pre="tswap")
        if sr is sc and dr is dc:
            self._conv_tc(sr.view(0), n_chunks, wr, kind, self._out_blk(dr.view(0), bias=bias, relu=relu), 4 * B,
                          n_out=n_out, wpk2=wc, w2_from_b=2 * B)
            return
        self._conv_tc(sr.view(0, sr0), n_chunks, wr, kind, self._out_blk(dr.view(0, dr0), bias=bias, relu=relu), 2 * B,
                      n_out=n_out)
        self._conv_tc(sc.view(0, sc0), n_chunks, wc, kind, self._out_blk(dc.view(0, dc0), bias=bias, relu=relu), 2 * B,
                      n_out=n_out)

    def _up_poly_fwd(self, key: str, x: "_Blk", w: torch.Tensor, bias: torch.Tensor, dst: CnpBlk, B: int) -> dict:
        """dst (8 chunks at 2H x 2W) = relu(conv5x5(bilinear_up2x(x)) + bias) without the upsampled tensor: two phase
        launches for the interior, the standard kernel on four strips for the 4-pixel band.  Returns the strips of the
        upsampled tensor (the band's weight gradient needs them)."""
        K = _cabi
        ncb, H, W = x.CB, x.H, x.W
        sv = lambda p: C.byref(p[0].view(0, p[1]))
        u = self._up_strip_blks(f"{key}.u", B, ncb, H, W)
        o = self._up_strip_blks(f"{key}.o", B, 8, H, W)
        wpk = [self._packed_weights(f"{key}.ph{a}", w, K.KIND_UP_PHASE, ncb, py=a, pre="phase") for a in (0, 1)]

        def phase(a):
            self._conv_tc(x.view(0), ncb, wpk[a], K.KIND_UP_PHASE,
                          self._out_blk(dst, bias=bias, relu=True, scatter=(2, a, 2, 0)), B, py=a)

        def strips():
            self._call("cnp_up_strips_fwd", C.byref(x.view(0)), ncb, sv(u[0]), sv(u[1]), B, _stream())
            self._strip_conv(u, o, ncb, w, key, K.KIND_K5S1, 64, B, bias=bias, relu=True)

        # the two row phases and the strips are independent until the scatter: three streams, so that the tail of one
        # persistent launch (816 tiles on 148 SMs: a half-empty last round) is filled by the next
        self._fork_join([lambda: phase(0), lambda: phase(1), strips])
        self._call("cnp_up_strips_scatter", sv(o[0]), sv(o[1]), C.byref(dst), B, _stream())
        return {"u": u}

    def _fork_join(self, fns) -> None:
        """Run ``fns[0]`` on the current stream and the others on side streams forked from it; join before returning.
        Single stream under the per-call profiler or CNP_NO_MULTISTREAM=1."""
        if self._prof is not None or os.environ.get("CNP_NO_MULTISTREAM") or len(fns) == 1:
            for fn in fns:
                fn()
            return
        if len(self._side_streams) < 4:
            self._side_streams = [torch.cuda.Stream() for _ in range(4)]
        main = torch.cuda.current_stream()
        fork = torch.cuda.Event()
        fork.record(main)
        used = []
        for k, fn in enumerate(fns[1:]):
            ss = self._side_streams[k % 4]
            ss.wait_event(fork)
            with torch.cuda.stream(ss):
                fn()
            used.append(ss)
        fns[0]()
        for ss in used:
            main.wait_stream(ss)

    def _up_poly_bwd(self, key: str, x: "_Blk", w: torch.Tensor, dy: CnpBlk, dx: "_Blk", saved: dict,
                     gw: torch.Tensor, gb: torch.Tensor, B: int, mask: Optional["_Blk"] = None,
                     on_wgrad_stream=None) -> None:
        """Backward of ``_up_poly_fwd``: gw += dL/dw, gb += dL/dbias, dx (all chunks of x) = mask * dL/dx; dy is the
        gradient w.r.t. the layer's pre-activation (8 chunks at 2H x 2W).  ``on_wgrad_stream(fn)``: runner for the
        launches that only produce parameter gradients (the UNet backward's side stream)."""
        K = _cabi
        S = _stream()
        if on_wgrad_stream is None:
            on_wgrad_stream = lambda fn: fn()
        ncb, H, W = x.CB, x.H, x.W
        Cin = ncb * 8
        sv = lambda p: C.byref(p[0].view(0, p[1]))
        s2d = self._blk(f"{key}.dys2d", B, 32, H, W)
        dys = self._up_strip_blks(f"{key}.dy", B, 8, H, W)
        # the producer's epilogue may already have written the space-to-depth copy (``saved["s2d_done"]``)
        self._call("cnp_up_dy_split", C.byref(dy), None if saved.get("s2d_done") else C.byref(s2d.view()), sv(dys[0]),
                   sv(dys[1]), B, S)

        def weight_gradient():
            S = _stream()
            wsb = _cabi.lib().cnp_conv_tc_wgrad_workspace_bytes()
            ws = self._buf("wgrad_ws", (wsb // 4,))
            # weight gradient: phase gradients at low resolution + the strips' 5x5 gradients, folded into gw
            dwp = self._buf(f"{key}.dwp", (2, 2, 64, Cin, 4, 4))
            dwp.zero_()
            self._call("cnp_conv_tc_wgrad", C.byref(x.view(0)), ncb, C.byref(s2d.view(0)), K.WG_UP_PHASE, _ptr(dwp), _ptr(gb),
                       Cin, B, _ptr(ws), wsb, S, work=(2.0 * B * H * W * 4 * 64 * Cin * 16, 0.0))
            u = saved["u"]
            # column strips are transposed images convolved with the tap-transposed weights: their gradient, transposed
            # back, is the gradient w.r.t. w itself
            if u[0][0] is u[1][0]:
                # square level: both strip groups in one launch (rows -> gw, columns -> dwt), the fold transposes dwt
                dwt = self._buf(f"{key}.dwt", (64, Cin, 5, 5))
                dwt.zero_()
                self._call("cnp_conv_tc_wgrad_pair", C.byref(u[0][0].view(0)), ncb, C.byref(dys[0][0].view(0)), _ptr(gw),
                           _ptr(dwt), 2 * B, _ptr(gb), Cin, 4 * B, _ptr(ws), wsb, S,
                           work=(2.0 * 4 * B * 6 * 2 * W * 64 * Cin * 25, 0.0))
                self._call("cnp_up_wgrad_fold", _ptr(dwp), _ptr(dwt), 64, Cin, _ptr(gw), S)
            else:
                self._call("cnp_conv_tc_wgrad", sv(u[0]), ncb, sv(dys[0]), K.WG_K5S1, _ptr(gw),
                           _ptr(gb), Cin, 2 * B, _ptr(ws), wsb, S, work=(2.0 * 2 * B * 6 * 2 * W * 64 * Cin * 25, 0.0))
                self._call("cnp_conv_tc_wgrad", sv(u[1]), ncb, sv(dys[1]), K.WG_K5S1_T, _ptr(gw),
                           _ptr(gb), Cin, 2 * B, _ptr(ws), wsb, S, work=(2.0 * 2 * B * 6 * 2 * H * 64 * Cin * 25, 0.0))
                self._call("cnp_up_wgrad_fold", _ptr(dwp), None, 64, Cin, _ptr(gw), S)
        on_wgrad_stream(weight_gradient)
        # input gradient: one low-res launch over the four dY phases, then the band through the strips
        wpk = self._packed_weights(f"{key}.dg.ph", w, K.KIND_UP_PHASE_DGRAD, 32, n_out=128, pre="phase")
        mk = mask.view(0) if mask is not None else None
        du = self._up_strip_blks(f"{key}.du", B, ncb, H, W)
        # pack on the current stream first (a side stream must not be the first consumer of the packing event)
        self._packed_weights(f"{key}.dg.00", w, K.KIND_K5S1_DGRAD, 8, 0, 0, 0, 128)
        self._packed_weights(f"{key}.dg.t", w, K.KIND_K5S1_DGRAD, 8, 0, 0, 0, 128, pre="tswap")
        self._fork_join([
            lambda: self._conv_tc(s2d.view(0), 32, wpk, K.KIND_UP_PHASE_DGRAD, self._out_blk(dx.view(0), mask=mk), B,
                                  n_out=128),
            lambda: self._strip_conv(dys, du, 8, w, key, K.KIND_K5S1_DGRAD, 128, B)])
        self._call("cnp_up_strips_bwd_fold", sv(du[0]), sv(du[1]), C.byref(dx.view()),
                   C.byref(mk) if mk is not None else None, ncb, B, _stream())

    def _unet_fwd_bf16(self, enc: Optional[torch.Tensor], B: int, n1: int, n2: int, need_z: bool = True,
                       x_aug: Optional["_Blk"] = None) -> Tuple[Optional[torch.Tensor], dict]:
        K = _cabi
        cfg, u = self.cfg, self.module.decoder.unet
        st = cfg.unet_strides
        L = len(st)
        res = self._levels(n1, n2)
        S = _stream()
        A = {}
        # the initial 1x1 is folded into the first 5x5 (fold_in.cu) when that layer has stride 1: its input is then
        # the encoder output itself, in 2 (4, 6) bf16 chunks with a constant-1 channel carrying the 1x1's bias
        fold_in = st[0] == 1 and cfg.in_channels + 1 <= 64 and not os.environ.get("CNP_NO_FOLD_IN")
        if fold_in and x_aug is not None:        # written by the fused encoder (enc_fused.cu)
            cb0, h_init = x_aug.CB, x_aug
        elif fold_in:
            cb0 = 2 * ((cfg.in_channels + 1 + 15) // 16)
            h_init = self._blk("x_aug", B, cb0, n1, n2)
            self._call("cnp_blk_from_nchw_f32_ones", _ptr(enc), enc.stride(0), B, cfg.in_channels, n1, n2,
                       C.byref(h_init.view()), cb0, int(getattr(self, "_enc_shared_mask", 0)), S)
        else:
            cb0 = 8
            h_init = self._blk("h_init", B, 8, n1, n2)
            self._call("cnp_conv1x1_in_bf16", _ptr(enc), enc.stride(0), _ptr(u.initial_linear.weight),
                       _ptr(u.initial_linear.bias), B, cfg.in_channels, 64, C.byref(h_init.view()), S)
        cat = [self._blk(f"cat{i}", B, 8 if i == L - 1 else 16, res[i][0], res[i][1]) for i in range(L)]
        A["h_init"], A["cat"], A["fold_in"] = h_init, cat, fold_in
        phases = [None] * L
        x = h_init
        # layer i's epilogue also writes the space-to-depth copy that the stride-2 layer i+1 reads
        for i in range(L):
            if st[i] == 2 and phases[i] is None:     # producer could not fuse it (first level): separate kernel
                ph = self._blk(f"phase{i}", B, 32, res[i][0], res[i][1])
                self._call("cnp_blk_space_to_depth", C.byref(x.view(0)), 8, C.byref(ph.view()), B, S)
                phases[i] = ph
            lyr = u.before_turn_layers[i]
            out = self._out_blk(cat[i].view(0), bias=lyr.bias, relu=True)
            if i + 1 < L and st[i + 1] == 2 and res[i][0] % 2 == 0 and res[i][1] % 2 == 0:
                phases[i + 1] = self._blk(f"phase{i + 1}", B, 32, res[i + 1][0], res[i + 1][1])
                out._s2d_view = phases[i + 1].view()
                out.s2d = C.pointer(out._s2d_view)
            if st[i] == 1 and i == 0 and fold_in:
                wpk = self._packed_weights("before0.folded", lyr.weight, K.KIND_K5S1, cb0,
                                           fold=(u.initial_linear.weight, u.initial_linear.bias))
                self._conv_tc(x.view(0), cb0, wpk, K.KIND_K5S1, out, B)
            elif st[i] == 1:
                wpk = self._packed_weights(f"before{i}", lyr.weight, K.KIND_K5S1, 8)
                self._conv_tc(x.view(0), 8, wpk, K.KIND_K5S1, out, B)
            else:
                wpk = self._packed_weights(f"before{i}", lyr.weight, K.KIND_K5S2, 32)
                self._conv_tc(phases[i].view(0), 32, wpk, K.KIND_K5S2, out, B)
            x = cat[i]
        A["phases"] = phases
        ups = [None] * L
        polys = [None] * L
        h_last = self._blk("h_last", B, 8, n1, n2)
        for i in range(L - 1, -1, -1):
            inp, ncb = cat[i], cat[i].CB
            if st[i] == 2 and self._up_poly_ok(cat[i]):
                # resize-convolution without the upsampled tensor (polyphase interior + strips for the band)
                lyr = u.after_turn_layers[i]
                dst = cat[i - 1].view(8) if i > 0 else h_last.view(0)
                polys[i] = self._up_poly_fwd(f"after{i}", cat[i], lyr.weight, lyr.bias, dst, B)
                continue
            if st[i] == 2:
                up = self._blk(f"up{i}", B, ncb, 2 * res[i][0], 2 * res[i][1])
                self._call("cnp_blk_upsample2x_fwd", C.byref(inp.view()), ncb, C.byref(up.view()), B, S)
                inp = up
            ups[i] = inp
            lyr = u.after_turn_layers[i]
            dst = cat[i - 1].view(8) if i > 0 else h_last.view(0)
            wpk = self._packed_weights(f"after{i}", lyr.weight, K.KIND_K5S1, ncb)
            self._conv_tc(inp.view(0), ncb, wpk, K.KIND_K5S1, self._out_blk(dst, bias=lyr.bias, relu=True), B)
        A["ups"], A["h_last"], A["polys"] = ups, h_last, polys
        if not need_z:   # training: the decoder runs on h_last and the final 1x1 moves behind it (dec_blk.cu)
            return None, A
        z = self._buf("z", (B, 64, n1, n2))
        o = CnpConvOut()
        o.mode = 1
        o.f32, o.f32_bstride, o.f32_ch_off = _ptr(z), z.stride(0), 0
        o.sy, o.ay, o.sx, o.ax = 1, 0, 1, 0
        o.bias = _ptr(u.final_linear.bias)
        wpk = self._packed_weights("final", u.final_linear.weight, K.KIND_K1, 8)
        self._conv_tc(h_last.view(0), 8, wpk, K.KIND_K1, o, B)
        return z, A

    def _unet_bwd_bf16(self, dz: torch.Tensor, enc: torch.Tensor, A: dict, grads, B, n1, n2, up_done=None):
        """Backward of the bf16 UNet: tcgen05 dgrad and wgrad on the blocked bf16 activations."""
        K = _cabi
        cfg, u = self.cfg, self.module.decoder.unet
        st = cfg.unet_strides
        L = len(st)
        res = self._levels(n1, n2)
        cat, ups, h_last, h_init = A["cat"], A["ups"], A["h_last"], A["h_init"]
        S = _stream()
        P = "decoder.unet."

        def to_f32(blk: _Blk, cb_off, nch, key):
            t = self._buf(key, (B, nch, blk.H, blk.W))
            self._call("cnp_blk_to_nchw_f32", C.byref(blk.view(cb_off)), B, nch, _ptr(t), t.stride(0), S)
            return t

        def wgrad_f32(x, dy, name, k, stride):
            Bn, Cin, H, W = x.shape
            self._call("cnp_conv2d_wgrad_f32", _ptr(x), x.stride(0), _ptr(dy), dy.stride(0),
                       _ptr(grads[name + ".weight"]), _ptr(grads[name + ".bias"]), Bn, Cin, H, W, dy.shape[1], k,
                       stride, S)

        # Weight gradients run on their own stream: a layer's wgrad and dgrad both consume dY and nothing downstream of
        # the backward reads a weight gradient before the optimiser, so the wgrad chain only has to wait for "dY ready"
        # and the two chains fill each other's tails (every big kernel here is a persistent launch whose last round
        # leaves SMs idle).  One workspace, used in stream order.  CNP_NO_WGRAD_STREAM=1 keeps a single stream.
        main_stream = torch.cuda.current_stream()
        wg_stream = None
        if not (os.environ.get("CNP_NO_WGRAD_STREAM") or os.environ.get("CNP_NO_MULTISTREAM")) and self._prof is None:
            if self._wgrad_stream is None:
                self._wgrad_stream = torch.cuda.Stream()
            wg_stream = self._wgrad_stream

        def on_wgrad_stream(fn):
            """Run ``fn`` (launches that only produce weight / bias gradients) after everything issued so far."""
            if wg_stream is None:
                fn()
                return
            ev = torch.cuda.Event()
            ev.record(main_stream)
            wg_stream.wait_event(ev)
            with torch.cuda.stream(wg_stream):
                fn()

        def join_wgrad_stream():
            if wg_stream is not None:
                main_stream.wait_stream(wg_stream)

        def wgrad_tc(x: CnpBlk, n_chunks, dy: CnpBlk, kind, name, Cin):
            kk = 1 if kind == K.WG_K1 else 25
            wsb = _cabi.lib().cnp_conv_tc_wgrad_workspace_bytes()
            ws = self._buf("wgrad_ws", (wsb // 4,))
            on_wgrad_stream(lambda: self._call(
                "cnp_conv_tc_wgrad", C.byref(x), n_chunks, C.byref(dy), kind, _ptr(grads[name + ".weight"]),
                _ptr(grads[name + ".bias"]), Cin, B, _ptr(ws), wsb, _stream(),
                work=(2.0 * B * dy.H * dy.W * 64 * Cin * kk, 0.0)))

        def dgrad_tc(dy: CnpBlk, w, key, kind, n_out_ch, dst: _Blk, dst_cb, mask: Optional[_Blk], mask_cb,
                     accumulate=False, phase=None, s2d: Optional[_Blk] = None):
            # 128 input channels (the skip concatenations): one WIDE launch; 64: one PAIR launch
            py, px = phase if phase is not None else (0, 0)
            wpk = self._packed_weights(f"{key}.dg.{py}{px}", w, kind, 8, py, px, 0, n_out_ch)
            mk = mask.view(mask_cb) if mask is not None else None
            sc = (2, py, 2, px if px < 2 else 0) if phase is not None else (1, 0, 1, 0)
            o = self._out_blk(dst.view(dst_cb), mask=mk, accumulate=accumulate, scatter=sc)
            if s2d is not None:
                # chunks 8..15 of this gradient are dY of the next (polyphase) decoder level: the epilogue also writes
                # their space-to-depth copy, band left zero (what cnp_up_dy_split would otherwise produce in a pass of its own)
                o._s2d_view = s2d.view()
                o.s2d, o.s2d_c0, o.s2d_band = C.pointer(o._s2d_view), 8, 2
            self._conv_tc(dy, 8, wpk, kind, o, B, py, px, n_out_ch)

        # final 1x1 (already folded into the decoder when dz is the blocked d_h_last)
        if isinstance(dz, _Blk):
            d_hl = dz
        else:
            dz_blk = self._blk("dz_blk", B, 8, n1, n2)
            self._call("cnp_blk_from_nchw_f32", _ptr(dz), dz.stride(0), B, 64, n1, n2, C.byref(dz_blk.view()), S)
            wgrad_tc(h_last.view(0), 8, dz_blk.view(0), K.WG_K1, P + "final_linear", 64)
            d_hl = self._blk("d_h_last", B, 8, n1, n2)
            dgrad_tc(dz_blk.view(0), u.final_linear.weight, "final", K.KIND_K1_DGRAD, 64, d_hl, 0, h_last, 0)
        d_cat = [self._blk(f"d_cat{i}", B, cat[i].CB, res[i][0], res[i][1]) for i in range(L)]
        dy_blk, dy_cb = d_hl, 0
        for i in range(0, L):
            name = P + f"after_turn_layers.{i}"
            lyr = u.after_turn_layers[i]
            if A["polys"][i] is not None:
                self._up_poly_bwd(f"after{i}", cat[i], lyr.weight, dy_blk.view(dy_cb), d_cat[i], A["polys"][i],
                                  grads[name + ".weight"], grads[name + ".bias"], B, mask=cat[i],
                                  on_wgrad_stream=on_wgrad_stream)
                if i < L - 1:
                    dy_blk, dy_cb = d_cat[i], 8
                continue
            x_in = ups[i]
            nch = x_in.CB * 8
            wgrad_tc(x_in.view(0), x_in.CB, dy_blk.view(dy_cb), K.WG_K5S1, name, nch)
            if st[i] == 2:
                d_up = self._blk(f"d_up{i}", B, x_in.CB, x_in.H, x_in.W)
                dgrad_tc(dy_blk.view(dy_cb), lyr.weight, f"after{i}", K.KIND_K5S1_DGRAD, nch, d_up, 0, None, 0)
                self._call("cnp_blk_upsample2x_bwd", C.byref(d_up.view()), x_in.CB, C.byref(d_cat[i].view()),
                           C.byref(cat[i].view()), 0, B, S)
            else:
                nxt = A["polys"][i + 1] if i + 1 < L else None
                s2d = None
                if nxt is not None and nch == 128 and not os.environ.get("CNP_NO_S2D_EPILOGUE"):
                    s2d = self._blk(f"after{i + 1}.dys2d", B, 32, cat[i + 1].H, cat[i + 1].W)
                    nxt["s2d_done"] = True
                dgrad_tc(dy_blk.view(dy_cb), lyr.weight, f"after{i}", K.KIND_K5S1_DGRAD, nch, d_cat[i], 0, cat[i], 0,
                         s2d=s2d)
            if i < L - 1:
                dy_blk, dy_cb = d_cat[i], 8
        if up_done is not None:          # gradients of the head, the final 1x1 and the up path are complete
            join_wgrad_stream()
            up_done()
        for i in range(L - 1, -1, -1):
            name = P + f"before_turn_layers.{i}"
            lyr = u.before_turn_layers[i]
            x_src = cat[i - 1] if i > 0 else h_init
            if i == 0 and A["fold_in"]:
                # folded first layer: tensor-core wgrad w.r.t. the folded weights (few input channels: the M rows
                # hold kernel rows instead), then the chain rule to W5 / W1 / b1; no dgrad (the encoder is not trained)
                il = u.initial_linear
                cp = x_src.CB * 8
                dwf = self._buf("dwf0", (64, cp, 5, 5))
                wsb = _cabi.lib().cnp_conv_tc_wgrad_workspace_bytes()
                ws = self._buf("wgrad_ws", (wsb // 4,))

                def first_layer_wgrad():
                    dwf.zero_()
                    self._call("cnp_conv_tc_wgrad", C.byref(x_src.view(0)), x_src.CB, C.byref(d_cat[0].view(0)),
                               K.WG_K5S1_NARROW, _ptr(dwf), _ptr(grads[name + ".bias"]), cp, B, _ptr(ws), wsb, _stream(),
                               work=(2.0 * B * n1 * n2 * 64 * cp * 25, 0.0))
                    self._call("cnp_fold_in_bwd", _ptr(dwf), _ptr(lyr.weight), _ptr(il.weight), _ptr(il.bias), 64, 64,
                               cfg.in_channels, cp, 5, _ptr(grads[name + ".weight"]),
                               _ptr(grads[P + "initial_linear.weight"]), _ptr(grads[P + "initial_linear.bias"]), _stream())
                on_wgrad_stream(first_layer_wgrad)
                continue
            if st[i] == 2:
                wgrad_tc(A["phases"][i].view(0), 32, d_cat[i].view(0), K.WG_K5S2, name, 64)
            else:
                wgrad_tc(x_src.view(0), 8, d_cat[i].view(0), K.WG_K5S1, name, 64)
            if i > 0:
                # Stride-2 input gradient by output phases.  Default: TWO launches (one per row phase), each producing both
                # x-phases in its two lane groups, so that the scattered read-modify-write of the epilogue touches whole
                # 32 B sectors (four single-phase launches each moved ~300 MB for 47 MB of output);
                # CNP_S2_DGRAD_4PHASE=1 keeps the four-launch form.
                if st[i] == 2 and os.environ.get("CNP_S2_DGRAD_4PHASE"):
                    phases = ((0, 0), (0, 1), (1, 0), (1, 1))
                else:
                    phases = ((0, 2), (1, 2))
                if st[i] == 2 and os.environ.get("CNP_NO_MULTISTREAM"):
                    for py, px in phases:
                        dgrad_tc(d_cat[i].view(0), lyr.weight, f"before{i}", K.KIND_K5S2_DGRAD, 64, d_cat[i - 1], 0,
                                 cat[i - 1], 0, accumulate=True, phase=(py, px))
                elif st[i] == 2:
                    # the phases write disjoint pixels: run them on separate streams so the small launches
                    # share the SMs instead of queueing behind each other
                    if len(self._side_streams) < 4:
                        self._side_streams = [torch.cuda.Stream() for _ in range(4)]
                    main = torch.cuda.current_stream()
                    fork = torch.cuda.Event()
                    fork.record(main)
                    for si, (py, px) in enumerate(phases):
                        ss = self._side_streams[si]
                        ss.wait_event(fork)
                        with torch.cuda.stream(ss):
                            dgrad_tc(d_cat[i].view(0), lyr.weight, f"before{i}", K.KIND_K5S2_DGRAD, 64, d_cat[i - 1], 0,
                                     cat[i - 1], 0, accumulate=True, phase=(py, px))
                        main.wait_stream(ss)
                else:
                    dgrad_tc(d_cat[i].view(0), lyr.weight, f"before{i}", K.KIND_K5S1_DGRAD, 64, d_cat[i - 1], 0,
                             cat[i - 1], 0, accumulate=True)
            else:
                d_init = self._blk("d_h_init", B, 8, n1, n2)
                if st[i] == 2:
                    raise NotImplementedError("first UNet level with stride 2")
                dgrad_tc(d_cat[0].view(0), lyr.weight, "before0", K.KIND_K5S1_DGRAD, 64, d_init, 0, None, 0)
                self._call("cnp_conv1x1_in_wgrad", _ptr(enc), enc.stride(0), cfg.in_channels, C.byref(d_init.view(0)), B,
                           _ptr(grads[P + "initial_linear.weight"]), _ptr(grads[P + "initial_linear.bias"]), S)
        join_wgrad_stream()

    # ------------------------------------------------------------------------------------------
    # (3)+(4) decoder, head
    # ------------------------------------------------------------------------------------------
    def _mlp_params(self, grads: Optional[Dict[str, torch.Tensor]] = None) -> CnpMlpParams:
        p = CnpMlpParams()
        layers = self.module.decoder.mlp.layers
        p.n_layers = len(layers)
        p.likelihood = _cabi.LIKELIHOODS[self.cfg.likelihood]
        dims = self.module.mlp_dims()
        for i, d in enumerate(dims):
            p.dims[i] = d
        for i, l in enumerate(layers):
            p.W[i], p.b[i] = l.weight.data_ptr(), l.bias.data_ptr()
            if grads is not None:
                p.dW[i] = grads[f"decoder.mlp.layers.{i}.weight"].data_ptr()
                p.db[i] = grads[f"decoder.mlp.layers.{i}.bias"].data_ptr()
        return p

    def forward(self, batch: DeviceBatch, with_loss: bool = True):
        """Returns dict(mean [B,Nt], var [B,Nt], logp [B] f64, count [B] i32, ctx)."""
        self._require_cuda()
        cfg, g, B, Nt = self.cfg, batch.grid, batch.B, batch.Nt
        self.generation += 1
        self._touch_signature((B, Nt, g.n1, g.n2, tuple(tuple(c.y.shape[1:]) for c in batch.contexts)))
        if self.precision == "bf16":
            self._prepack_all()
        if batch.ready is not None:   # uploaded on a copy stream: order after the copy, keep the allocator informed
            cur = torch.cuda.current_stream()
            cur.wait_event(batch.ready)
            for t in self._batch_tensors(batch):
                t.record_stream(cur)
            batch.ready = None
        # with the folded first layer the only reader of the encoder output is the blocked conversion, which broadcasts
        # the channels of batch-shared context sets itself
        folded = (self.precision != "fp32" and cfg.unet_strides[0] == 1 and cfg.in_channels + 1 <= 64
                  and not os.environ.get("CNP_NO_FOLD_IN"))
        x_aug = self.encode_blocked(batch) if folded else None
        enc = self.encode(batch, broadcast=not folded) if x_aug is None else None
        on_grid = isinstance(batch.xt, tuple)
        if self.precision == "fp32":
            z, A = self._unet_fwd_f32(enc, B, g.n1, g.n2)
        else:
            # the fused on-grid decoders carry the Gaussian head only: other likelihoods decode from z (unfused)
            z, A = self._unet_fwd_bf16(enc, B, g.n1, g.n2, need_z=(on_grid and cfg.likelihood not in ("cnp", "het")),
                                       x_aug=x_aug)
            if not os.environ.get("CNP_NO_PREPACK"):
                self.weights_dirty = False     # every known packing (forward and dgrad) was refreshed by _prepack_all
            if self._pack_event_bwd is not None and not torch.is_grad_enabled():
                # forward-only call: no backward will join the packing stream's tail, do it here (stream capture needs it)
                torch.cuda.current_stream().wait_event(self._pack_event_bwd)
                self._pack_event_bwd = None
        Cz = cfg.unet_out_channels
        s2 = self._scale2(self.module.decoder.set_conv.log_scale)
        if on_grid:
            if z is None:
                return self._decode_grid_fused(batch, A["h_last"], s2)
            return self._decode_grid(batch, z, s2)
        f = self._buf("f", (B, Cz, Nt))
        if z is None:
            fin = self.module.decoder.unet.final_linear
            gbuf, swbuf = self._buf("dec_g", (B, Nt, 64)), self._buf("dec_sw", (B, Nt))
            A["dec_g"], A["dec_sw"] = gbuf, swbuf
            self._call("cnp_dec_blk_fwd", C.byref(A["h_last"].view(0)), _ptr(batch.xt), B, Nt, g.start1, g.start2, g.res,
                       s2, _ptr(fin.weight), _ptr(fin.bias), Cz, _ptr(gbuf), _ptr(swbuf), _ptr(f), _stream())
        else:
            self._call("cnp_setconv_dec_offgrid_fwd", _ptr(z), z.stride(0), _ptr(batch.xt), B, Cz, Nt, g.start1, g.n1,
                       g.start2, g.n2, g.res, s2, _ptr(f), Cz, _stream())
        mean = torch.empty((B, Nt), dtype=torch.float32, device=self.device)
        var = torch.empty((B, Nt), dtype=torch.float32, device=self.device)
        logp = count = None
        if with_loss:
            logp = torch.zeros(B, dtype=torch.float64, device=self.device)
            count = torch.zeros(B, dtype=torch.int32, device=self.device)
        p = self._mlp_params()
        Ca = cfg.dim_aux_t
        if batch.aux_t is None or batch.aux_t.shape[1] != Ca:
            raise ValueError(f"task needs Y_t_aux with {Ca} channels")
        zraw = None if p.likelihood == 0 else self._buf("head_zraw", (B, cfg.likelihood_channels, Nt))
        self._call("cnp_mlp_head_fwd", C.byref(p), _ptr(f), Cz, Cz, _ptr(batch.aux_t), Ca,
                   _ptr(batch.yt) if with_loss else None, B, Nt, _ptr(mean), _ptr(var), _ptr(zraw), _ptr(logp),
                   _ptr(count), _stream())
        return dict(mean=mean, var=var, logp=logp, count=count, ctx=dict(enc=enc, z=z, A=A, f=f, generation=self.generation))

    def _decode_grid(self, batch: DeviceBatch, z: torch.Tensor, s2: float):
        """On-grid targets (predict): separable truncated SetConv + per-point MLP head -> mean/std [B,P,Q]."""
        cfg, g, B = self.cfg, batch.grid, batch.B
        x1t, x2t = batch.xt
        P, Q, Cz, Ca = int(x1t.shape[-1]), int(x2t.shape[-1]), z.shape[1], cfg.dim_aux_t
        aux = batch.aux_t
        if aux is None or aux.shape[-3] != Ca or tuple(aux.shape[-2:]) != (P, Q):
            raise ValueError(f"on-grid prediction needs Y_t_aux of shape [{Ca},{P},{Q}]")
        aux = aux.reshape(-1, Ca, P, Q)
        aux_bs = 0 if aux.shape[0] == 1 else aux.stride(0)
        wsb = _cabi.lib().cnp_setconv_dec_grid_workspace_bytes(B, Cz, g.n1, P, Q)
        ws = self._buf("dec_grid_ws", ((wsb + 3) // 4,))
        f = self._buf("f_grid", (B, Cz, P, Q))
        self._call("cnp_setconv_dec_grid_fwd", _ptr(z), z.stride(0), _ptr(x1t), _ptr(x2t), B, Cz, P, Q, g.start1, g.n1,
                   g.start2, g.n2, g.res, s2, _ptr(f), f.stride(0), _ptr(ws), wsb, _stream(),
                   work=(0.0, 4.0 * (z.numel() + f.numel())))
        mean = torch.empty((B, P, Q), dtype=torch.float32, device=self.device)
        std = torch.empty((B, P, Q), dtype=torch.float32, device=self.device)
        p = self._mlp_params()
        if p.likelihood != 0:
            # Bernoulli-Gamma / spikes-Beta: the general head kernel over the flattened target grid ([B,C,P,Q] is
            # [B,C,P*Q]); mean and std of the spikes-and-slab distribution
            auxb = aux if aux.shape[0] == B else aux.expand(B, -1, -1, -1).contiguous()
            zraw = self._buf("head_zraw_grid", (B, cfg.likelihood_channels, P * Q))
            self._call("cnp_mlp_head_fwd", C.byref(p), _ptr(f), Cz, Cz, _ptr(auxb), Ca, None, B, P * Q, _ptr(mean), _ptr(std),
                       _ptr(zraw), None, None, _stream())
            std.sqrt_()
            return dict(mean=mean, std=std, var=None, logp=None, count=None, ctx=None)
        self._call("cnp_mlp_head_points_fwd", C.byref(p), _ptr(f), f.stride(0), Cz, _ptr(aux), aux_bs, Ca, B, P * Q,
                   _ptr(mean), _ptr(std), _stream(), work=(2.0 * B * P * Q * sum(
                       a * b for a, b in zip(self.module.mlp_dims()[:-1], self.module.mlp_dims()[1:])), 0.0))
        return dict(mean=mean, std=std, var=None, logp=None, count=None, ctx=None)

    def _decode_grid_fused(self, batch: DeviceBatch, h_last: _Blk, s2: float):
        """bf16 mode, on-grid targets: SetConv of the last hidden activation + final 1x1 folded into the MLP + head in
        two kernels (decode_grid.cu); the 64 x P x Q decoder output is never materialised."""
        cfg, g, B = self.cfg, batch.grid, batch.B
        x1t, x2t = batch.xt
        P, Q, Ca = int(x1t.shape[-1]), int(x2t.shape[-1]), cfg.dim_aux_t
        aux = batch.aux_t
        if aux is None or aux.shape[-3] != Ca or tuple(aux.shape[-2:]) != (P, Q):
            raise ValueError(f"on-grid prediction needs Y_t_aux of shape [{Ca},{P},{Q}]")
        aux = aux.reshape(-1, Ca, P, Q)
        aux_bs = 0 if aux.shape[0] == 1 else aux.stride(0)
        mean = torch.empty((B, P, Q), dtype=torch.float32, device=self.device)
        std = torch.empty((B, P, Q), dtype=torch.float32, device=self.device)
        fin = self.module.decoder.unet.final_linear
        p = self._mlp_params()
        dims = self.module.mlp_dims()
        flops = 2.0 * B * P * Q * (sum(a * b for a, b in zip(dims[:-1], dims[1:])) + 64 * 30)
        work = (flops, 4.0 * B * P * Q * (Ca + 2) + 2.0 * B * 64 * g.n1 * g.n2)
        # tensor-core tail when the MLP is the standard 3 x 64 one and the target grid is fine enough that any 128
        # consecutive target columns (plus the SetConv band) fit in 64 internal-grid columns
        use_tc = (os.environ.get("CNP_DECODE_TC", "1") != "0" and len(dims) == 5 and all(d == 64 for d in dims[1:4])
                  and Ca <= 6 and g.n2 >= 64 and batch.xt_host is not None)
        if use_tc:
            x2 = batch.xt_host[1].astype(np.float64)
            band = 2.0 * math.sqrt(2.0 * 104.0 * s2) / g.res + 4.0
            span = max((abs(x2[min(i + 127, Q - 1)] - x2[i]) for i in range(0, Q, 128)), default=0.0) / g.res
            use_tc = span + band <= 64.0
        name = "cnp_decode_grid_tc" if use_tc else "cnp_decode_grid_fused"
        wsb = getattr(_cabi.lib(), name + "_workspace_bytes")(B, g.n2 if use_tc else g.n1, P, Q)
        ws = self._buf("dec_fused_ws", ((wsb + 3) // 4,))
        self._call(name + "_fwd", C.byref(h_last.view(0)), _ptr(x1t), _ptr(x2t), B, P, Q, g.start1, g.start2,
                   g.res, s2, _ptr(fin.weight), _ptr(fin.bias), C.byref(p), _ptr(aux), aux_bs, Ca, _ptr(mean), _ptr(std),
                   _ptr(ws), wsb, _stream(), work=work)
        return dict(mean=mean, std=std, var=None, logp=None, count=None, ctx=None)

    def backward(self, batch: DeviceBatch, ctx: dict, dlogp: torch.Tensor) -> Dict[str, torch.Tensor]:
        """dlogp [B] fp32 = d loss / d logp_b.  Returns {param name: grad} (views of one flat buffer)."""
        cfg, g, B, Nt = self.cfg, batch.grid, batch.B, batch.Nt
        if ctx.get("generation") != self.generation:
            # the activations a backward needs live in engine-wide workspaces keyed by shape: a later forward
            # (a second loss_fn before .backward(), a validation pass with grad enabled) has overwritten them
            raise _cabi.CnpError("backward() of a loss whose forward is no longer the engine's latest one: call "
                                 ".backward() before the next loss_fn / model(task) (one loss in flight per model)")
        named = [(n, p) for n, p in self.module.named_parameters() if n.startswith("decoder.") and p.dim() > 0]
        # Flat gradient buffer in TWO buckets ordered by when the backward finishes them: [head MLP, final 1x1, up path]
        # first, [down path, initial 1x1] second.  Data-parallel training all-reduces the first bucket while the
        # down-path backward still runs (dist.enable_data_parallel; SURVEY 8(e): one 4.58 MB exchange per step).
        first = lambda n: n.startswith(("decoder.mlp.", "decoder.unet.final_linear.", "decoder.unet.after_turn_layers."))
        named = [t for t in named if first(t[0])] + [t for t in named if not first(t[0])]
        total = sum(p.numel() for _, p in named)
        n_first = sum(p.numel() for n, p in named if first(n))
        flat = torch.zeros(total, dtype=torch.float32, device=self.device)
        grads, off = {}, 0
        for n, p in named:
            grads[n] = flat[off:off + p.numel()].view_as(p)
            off += p.numel()
        dp = self.allreduce_group is not None and self.world_size > 1
        pending = []

        def reduce_bucket(t):
            import torch.distributed as dist
            if dist.get_backend(self.allreduce_group) == "nccl":      # mean over ranks inside the collective
                pending.append(dist.all_reduce(t, op=dist.ReduceOp.AVG, group=self.allreduce_group, async_op=True))
            else:
                dist.all_reduce(t, group=self.allreduce_group)
                t.mul_(1.0 / self.world_size)

        up_done = (lambda: reduce_bucket(flat[:n_first])) if dp else None
        f, z = ctx["f"], ctx["z"]
        Cz = cfg.unet_out_channels
        df = self._buf("df", (B, Cz, Nt))
        p = self._mlp_params(grads)
        hwsb = _cabi.lib().cnp_mlp_head_bwd_workspace_bytes(C.byref(p), B, Nt)
        hws = self._buf("head_bwd_ws", (max(1, hwsb // 4),))
        self._call("cnp_mlp_head_bwd", C.byref(p), _ptr(f), Cz, Cz, _ptr(batch.aux_t), cfg.dim_aux_t, _ptr(batch.yt), B,
                   Nt, _ptr(dlogp), _ptr(df), _ptr(hws), hwsb, _stream())
        s2 = self._scale2(self.module.decoder.set_conv.log_scale)
        if z is None:
            A = ctx["A"]
            fin = "decoder.unet.final_linear"
            dg = self._buf("dec_dg", (B, Nt, 64))
            self._call("cnp_dec_blk_bwd_params", _ptr(df), _ptr(A["dec_g"]), _ptr(A["dec_sw"]),
                       _ptr(self.module.decoder.unet.final_linear.weight), B, Nt, Cz, _ptr(dg),
                       _ptr(grads[fin + ".weight"]), _ptr(grads[fin + ".bias"]), _stream())
            dz = self._blk("d_h_last", B, 8, g.n1, g.n2)
            self._call("cnp_dec_blk_bwd_data", _ptr(dg), _ptr(batch.xt), B, Nt, g.start1, g.start2, g.res, s2,
                       C.byref(A["h_last"].view(0)), C.byref(dz.view(0)), _stream(),
                       work=(0.0, 2.0 * B * 64 * g.n1 * g.n2))
        else:
            dz = self._buf("dz", z.shape)
            self._call("cnp_setconv_dec_offgrid_bwd", _ptr(df), Cz, _ptr(batch.xt), B, Cz, Nt, g.start1, g.n1, g.start2,
                       g.n2, g.res, s2, _ptr(dz), dz.stride(0), _stream(), work=(0.0, 4.0 * (dz.numel() + df.numel())))
        if self.precision == "fp32":
            self._unet_bwd_f32(dz, ctx["enc"], ctx["A"], grads, B, g.n1, g.n2, up_done=up_done)
        else:
            self._unet_bwd_bf16(dz, ctx["enc"], ctx["A"], grads, B, g.n1, g.n2, up_done=up_done)
        if dp:
            reduce_bucket(flat[n_first:])
            for w in pending:
                w.wait()              # orders the current stream after the collectives (no host block with NCCL)
        self._flat_grad = flat
        self.weights_dirty = True     # an optimiser step follows a backward: the next forward re-packs the bf16 weights
        return grads
