"""``ConvNP`` -- drop-in for ``deepsensor.model.convnp.ConvNP`` on the path nzdownscale drives.

Call forms kept from the reference:
  ConvNP(data_processor, task_loader, **kw)      nzdownscale/downscaler/train.py:238-241,
                                                 validate_ERA.py:103-105, validate_WRF.py:314-316
  model.model  (nn.Module: parameters / state_dict / encoder)          train.py:249-262,354,413
  model.loss_fn(task, normalise=True)                                  train.py:370
  model(task)                                                          experiments/deepsensor/nz_downscaling.py:343
  model.predict(tasks, X_t=..., progress_bar=..., transform_params=..) validate_ERA.py:88-92
  model.save(dir)                                                      train_downscaling.py:208

Underneath, every array operation of the forward/backward is a kernel of libconvnp_b200.so
(``engine.Engine``); there is no torch-op or CPU fallback.
"""
from __future__ import annotations

import json
import os
from typing import Dict, List, Optional, Sequence, Union

import numpy as np
import torch

from . import _cabi
from .engine import DeviceBatch, Engine, HostBatch
from .model import ConvNPConfig, ConvNPModule
from .task import Masked, Task, convert_task_to_nps_args


class _LossFn(torch.autograd.Function):
    """loss = -mean_b(logp_b [/ N_b]); backward = the hand-written kernels."""

    @staticmethod
    def forward(ctx, engine: Engine, batch: DeviceBatch, normalise: bool, *params):
        out = engine.forward(batch, with_loss=True)
        logp, count = out["logp"], out["count"]
        # one launch for the scalar algebra: loss = -mean_b(logp_b / N_b) in float64 and the backward's seed d loss / d logp
        loss = torch.empty((), dtype=torch.float64, device=logp.device)
        coef = torch.empty(logp.shape[0], dtype=torch.float32, device=logp.device)
        engine._call("cnp_loss_mean", logp.data_ptr(), count.data_ptr(), int(logp.shape[0]), int(normalise), loss.data_ptr(),
                     coef.data_ptr(), torch.cuda.current_stream().cuda_stream)
        ctx.engine, ctx.batch, ctx.fctx = engine, batch, out["ctx"]
        ctx.coef = coef
        return loss

    @staticmethod
    def backward(ctx, grad_out):
        eng: Engine = ctx.engine
        dlogp = ctx.coef * grad_out.to(torch.float32)
        grads = eng.backward(ctx.batch, ctx.fctx, dlogp)
        outs = []
        for n, p in eng.module.named_parameters():
            outs.append(grads.get(n) if p.requires_grad else None)
        return (None, None, None, *outs)


class GaussianPrediction(dict):
    """What ``model(task)`` returns: mean / var / std tensors shaped like upstream ([B,1,Nt])."""

    @property
    def mean(self):
        return self["mean"]

    @property
    def var(self):
        return self["var"]

    @property
    def std(self):
        return self["std"]


class ConvNP:
    def __init__(self, *args, config: Optional[ConvNPConfig] = None, precision: Optional[str] = None,
                 verbose: bool = True, **kwargs):
        self.data_processor = None
        self.task_loader = None
        if len(args) >= 1:
            self.data_processor = args[0]
        if len(args) >= 2:
            self.task_loader = args[1]
        if len(args) >= 3 and isinstance(args[2], str):
            kwargs["model_ID"] = args[2]
        model_ID = kwargs.pop("model_ID", None)
        if model_ID is not None:
            config, state = self._read(model_ID)
        else:
            state = None
        if config is None:
            config = self._infer_config(self.task_loader, kwargs, verbose)
        self.config = config
        self.precision = precision or os.environ.get("CONVNP_B200_PRECISION", "fp32")
        self.model = ConvNPModule(config)
        if torch.cuda.is_available():
            self.model.to(torch.device("cuda", torch.cuda.current_device()))
        if state is not None:
            self.model.load_state_dict(state)
        self.engine = Engine(self.model, self.precision)
        self.N_mixture_components = 1

    # ------------------------------------------------------------------------------------------
    # construction helpers (deepsensor.model.defaults / convnp.__init__)
    # ------------------------------------------------------------------------------------------
    @staticmethod
    def _infer_config(task_loader, kw: dict, verbose: bool) -> ConvNPConfig:
        kw = dict(kw)
        say = print if verbose else (lambda *a, **k: None)
        if "dim_yc" not in kw:
            if task_loader is None:
                raise ValueError("ConvNP needs a TaskLoader or explicit dim_yc / dim_yt / dim_aux_t")
            kw["dim_yc"] = tuple(task_loader.context_dims)
            say(f"dim_yc inferred from TaskLoader: {kw['dim_yc']}")
        if "dim_yt" not in kw:
            kw["dim_yt"] = int(sum(task_loader.target_dims)) if task_loader is not None else 1
            say(f"dim_yt inferred from TaskLoader: {kw['dim_yt']}")
        if "dim_aux_t" not in kw:
            kw["dim_aux_t"] = int(task_loader.aux_at_target_dims) if task_loader is not None else 0
            say(f"dim_aux_t inferred from TaskLoader: {kw['dim_aux_t']}")
        ppu = kw.pop("internal_density", None) or kw.pop("points_per_unit", None)
        if ppu is None:
            if task_loader is None:
                raise ValueError("internal_density is required without a TaskLoader")
            ppu = task_loader.gen_ppu()
            say(f"internal_density inferred from TaskLoader: {ppu}")
        kw["points_per_unit"] = float(ppu)
        if "encoder_scales" not in kw:
            if task_loader is not None:
                kw["encoder_scales"] = tuple(task_loader.gen_encoder_scales(ppu))
            else:
                kw["encoder_scales"] = tuple(0.5 / ppu for _ in kw["dim_yc"])
            say(f"encoder_scales inferred from TaskLoader: {list(kw['encoder_scales'])}")
        elif not isinstance(kw["encoder_scales"], (list, tuple)):
            kw["encoder_scales"] = tuple(float(kw["encoder_scales"]) for _ in kw["dim_yc"])
        if "decoder_scale" not in kw:
            kw["decoder_scale"] = 1.0 / ppu
            say(f"decoder_scale inferred from TaskLoader: {kw['decoder_scale']}")
        if kw.get("dim_aux_t", 0) > 0 and "aux_t_mlp_layers" not in kw:
            kw["aux_t_mlp_layers"] = (64, 64, 64)
        for drop in ("unet_resize_convs", "unet_resize_conv_interp_method", "encoder_scales_learnable",
                     "decoder_scale_learnable", "dim_x", "verbose"):
            kw.pop(drop, None)
        return ConvNPConfig(**kw)

    @staticmethod
    def _read(model_dir: str):
        with open(os.path.join(model_dir, "model_config.json")) as f:
            cfg = json.load(f)
        state = torch.load(os.path.join(model_dir, "model.pt"), map_location="cpu")
        return ConvNPConfig(**cfg), state

    def save(self, model_ID: str):
        """Write ``model.pt`` + ``model_config.json`` like upstream ``ConvNP.save``."""
        os.makedirs(model_ID, exist_ok=True)
        torch.save({k: v.detach().cpu() for k, v in self.model.state_dict().items()},
                   os.path.join(model_ID, "model.pt"))
        with open(os.path.join(model_ID, "model_config.json"), "w") as f:
            json.dump(self.config.to_json(), f, indent=2)

    # ------------------------------------------------------------------------------------------
    # task -> device
    # ------------------------------------------------------------------------------------------
    @classmethod
    def modify_task(cls, task: Task) -> Task:
        """Host-side mirror of upstream ``ConvNP.modify_task`` (numpy only; tensors are made by the engine)."""
        if "batch_dim" not in task["ops"]:
            task = task.add_batch_dim()
        if "float32" not in task["ops"]:
            task = task.cast_to_float32()
        if "numpy_mask" not in task["ops"]:
            task = task.mask_nans_numpy()
        if "nps_mask" not in task["ops"]:
            task = task.mask_nans_nps()
        return task

    def stage_task(self, task: Task, pinned: bool = True, ctx_cache: Optional[dict] = None) -> HostBatch:
        """Stage a task once in page-locked host memory; ``loss_fn`` / ``__call__`` accept the result and
        then only pay the asynchronous H2D copy per call."""
        return self._to_device(task, pinned=pinned, upload=False, ctx_cache=ctx_cache)

    def _to_device(self, task, pinned: bool = False, upload: bool = True, ctx_cache: Optional[dict] = None):
        """Upload a task.  Raw tasks skip the host NaN scans: NaNs travel to the GPU and the encoder
        kernels derive the masks there (identical result to the Masked path, see tests)."""
        if isinstance(task, DeviceBatch):
            return task
        if isinstance(task, HostBatch):
            return self.engine.upload(task)
        if "nps_mask" in task["ops"] or "numpy_mask" in task["ops"]:
            task = self.modify_task(task)
        else:
            if "batch_dim" not in task["ops"]:
                task = task.add_batch_dim()
            if "float32" not in task["ops"]:
                task = task.cast_to_float32()
        ctx_data, xt, yt, kw = convert_task_to_nps_args(task)
        contexts = []
        for x, y in ctx_data:
            if isinstance(y, Masked):
                contexts.append((x, y.y, y.mask))
            elif isinstance(y, np.ma.MaskedArray):
                contexts.append((x, y.filled(np.nan), None))
            else:
                contexts.append((x, y, None))
        if isinstance(yt, np.ma.MaskedArray):
            yt = yt.filled(np.nan)
        hb = self.engine.stage_host(contexts, xt, yt, kw.get("aux_t"), pinned=pinned, ctx_cache=ctx_cache)
        return self.engine.upload(hb) if upload else hb

    # ------------------------------------------------------------------------------------------
    # public API
    # ------------------------------------------------------------------------------------------
    def loss_fn(self, task: Union[Task, DeviceBatch], fix_noise=None, num_lv_samples: int = 8,
                normalise: bool = False) -> torch.Tensor:
        batch = self._to_device(task)
        if batch.yt is None:
            raise ValueError("loss_fn needs target observations (Y_t)")
        params = [p for _, p in self.model.named_parameters()]
        if torch.is_grad_enabled() and any(p.requires_grad for p in params):
            return _LossFn.apply(self.engine, batch, bool(normalise), *params)
        out = self.engine.forward(batch, with_loss=True)
        denom = out["count"].clamp(min=1).to(torch.float64) if normalise else torch.ones_like(out["logp"])
        return -(out["logp"] / denom).mean()

    def __call__(self, task: Union[Task, DeviceBatch], n_samples: int = 10, requires_grad: bool = False):
        batch = self._to_device(task)
        with torch.no_grad():
            out = self.engine.forward(batch, with_loss=False)
        if out["var"] is None:  # on-grid targets: [B,1,P,Q]
            mean, std = out["mean"].unsqueeze(1), out["std"].unsqueeze(1)
            return GaussianPrediction(mean=mean, var=std * std, std=std)
        mean, var = out["mean"].unsqueeze(1), out["var"].unsqueeze(1)
        return GaussianPrediction(mean=mean, var=var, std=var.sqrt())

    def mean(self, task):
        return self(task)["mean"][0].cpu().numpy()

    def variance(self, task):
        return self(task)["var"][0].cpu().numpy()

    def std(self, task):
        return self(task)["std"][0].cpu().numpy()

    def stddev(self, task):
        return self.std(task)

    def logpdf(self, task):
        batch = self._to_device(task)
        with torch.no_grad():
            out = self.engine.forward(batch, with_loss=True)
        return float(out["logp"].sum().cpu())

    def predict(self, tasks, X_t, X_t_mask=None, X_t_is_normalised: bool = False, aux_at_targets_override=None,
                aux_at_targets_override_is_normalised: bool = False, resolution_factor: int = 1,
                pred_params=("mean", "std"), n_samples: int = 0, ar_sample: bool = False, unnormalise: bool = True,
                seed: int = 0, append_indexes=None, progress_bar: int = 0, verbose: bool = False,
                transform_params=None):
        from .predict import predict as _predict
        return _predict(self, tasks, X_t, X_t_mask=X_t_mask, X_t_is_normalised=X_t_is_normalised,
                        aux_at_targets_override=aux_at_targets_override, resolution_factor=resolution_factor,
                        pred_params=pred_params, unnormalise=unnormalise, progress_bar=progress_bar,
                        transform_params=transform_params)
