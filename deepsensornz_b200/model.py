"""Parameter container of the ConvNP (``ConvNP.model`` in DeepSensor terms).

The reference only relies on this object being an ``nn.Module`` with ``.parameters()``,
``.state_dict()``, ``.load_state_dict()`` and an ``.encoder`` sub-module whose parameters can be
frozen (nzdownscale/downscaler/train.py:249-251,257,262,354,413; validate_ERA.py:109).  Layer
structure restates upstream ``neuralprocesses.construct_convgnp`` as configured by DeepSensor's
``ConvNP``: frozen set-conv log-scales, a UNet with
5x5 convolutions / strides (1,2,2,2) / bilinear resize-convs, a decoder set-conv and the
aux-at-target MLP.  Arithmetic is NOT done here: ``forward`` belongs to ``engine.Engine`` which
drives libconvnp_b200.so.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field, asdict
from typing import List, Sequence, Tuple

import torch
import torch.nn as nn


# head inputs per target variable (upstream neuralprocesses construct_likelihood: mean + pre-softplus variance;
# k~, scale~, 2 log-probabilities; alpha~, beta~, 3 log-probabilities)
LIKELIHOOD_CHANNELS = {"cnp": 2, "het": 2, "bernoulli-gamma": 4, "cnp-spikes-beta": 5, "spikes-beta": 5}


@dataclass
class ConvNPConfig:
    dim_yc: Tuple[int, ...]
    dim_yt: int = 1
    dim_aux_t: int = 0
    points_per_unit: float = 250.0
    encoder_scales: Tuple[float, ...] = ()
    decoder_scale: float = 0.004
    unet_channels: Tuple[int, ...] = (64, 64, 64, 64)
    unet_kernels: int = 5
    unet_strides: Tuple[int, ...] = (1, 2, 2, 2)
    aux_t_mlp_layers: Tuple[int, ...] = (64, 64, 64)
    likelihood: str = "cnp"
    margin: float = 0.1
    epsilon: float = 1e-2

    def __post_init__(self):
        self.dim_yc = tuple(int(c) for c in self.dim_yc)
        self.unet_channels = tuple(int(c) for c in self.unet_channels)
        self.unet_strides = tuple(int(s) for s in self.unet_strides)
        self.aux_t_mlp_layers = tuple(int(c) for c in self.aux_t_mlp_layers)
        self.encoder_scales = tuple(float(s) for s in self.encoder_scales)
        if len(self.encoder_scales) != len(self.dim_yc):
            raise ValueError("need one encoder scale per context set")
        if len(self.unet_strides) != len(self.unet_channels):
            raise ValueError("need one stride per UNet level")
        if self.unet_kernels != 5:
            raise NotImplementedError("the hot path implements the 5x5 UNet DeepSensor configures")
        if any(s not in (1, 2) for s in self.unet_strides):
            raise NotImplementedError("UNet strides must be 1 or 2")
        if self.likelihood not in LIKELIHOOD_CHANNELS:
            raise NotImplementedError(
                f"likelihood '{self.likelihood}': the hot path has the heads nzdownscale selects per variable "
                "(nzdownscale/dataprocess/config.py:162-169): 'cnp' / 'het' (Gaussian), 'bernoulli-gamma', "
                "'cnp-spikes-beta'; the low-rank 'gnp' head is not built")
        if self.dim_yt != 1:
            raise NotImplementedError("single target variable (dim_yt=1) only")

    @property
    def in_channels(self) -> int:
        return sum(c + 1 for c in self.dim_yc)

    @property
    def grid_multiple(self) -> int:
        m = 1
        for s in self.unet_strides:
            m *= s
        return m

    @property
    def likelihood_channels(self) -> int:
        """Inputs of the likelihood per target point: 2 (Gaussian), 4 (Bernoulli-Gamma), 5 (spikes-Beta), times dim_yt."""
        return LIKELIHOOD_CHANNELS[self.likelihood] * self.dim_yt

    @property
    def unet_out_channels(self) -> int:
        return self.unet_channels[0] if self.dim_aux_t > 0 else self.likelihood_channels

    def to_json(self) -> dict:
        return asdict(self)


class SetConvScale(nn.Module):
    """A (frozen) set-conv length scale, stored as its log like upstream."""

    def __init__(self, scale: float):
        super().__init__()
        self.log_scale = nn.Parameter(torch.tensor(math.log(scale), dtype=torch.float32), requires_grad=False)


class Encoder(nn.Module):
    def __init__(self, scales: Sequence[float]):
        super().__init__()
        self.set_convs = nn.ModuleList([SetConvScale(s) for s in scales])


class UNetParams(nn.Module):
    def __init__(self, cfg: ConvNPConfig):
        super().__init__()
        ch, st = cfg.unet_channels, cfg.unet_strides
        L = len(ch)
        prev = (ch[0],) + tuple(ch)
        self.initial_linear = nn.Conv2d(cfg.in_channels, ch[0], 1)
        self.before_turn_layers = nn.ModuleList(
            [nn.Conv2d(prev[i], ch[i], 5, stride=st[i], padding=2) for i in range(L)])
        self.after_turn_layers = nn.ModuleList(
            [nn.Conv2d(ch[i] if i == L - 1 else 2 * ch[i], prev[i], 5, padding=2) for i in range(L)])
        self.final_linear = nn.Conv2d(ch[0], cfg.unet_out_channels, 1)


class MLPParams(nn.Module):
    def __init__(self, dims: Sequence[int]):
        super().__init__()
        self.layers = nn.ModuleList([nn.Linear(dims[i], dims[i + 1]) for i in range(len(dims) - 1)])


class Decoder(nn.Module):
    def __init__(self, cfg: ConvNPConfig):
        super().__init__()
        self.unet = UNetParams(cfg)
        self.set_conv = SetConvScale(cfg.decoder_scale)
        if cfg.dim_aux_t > 0:
            dims = (cfg.unet_channels[0] + cfg.dim_aux_t,) + tuple(cfg.aux_t_mlp_layers) + (cfg.likelihood_channels,)
            self.mlp = MLPParams(dims)
        else:
            self.mlp = MLPParams(())


class ConvNPModule(nn.Module):
    """``nps.Model(encoder, decoder)`` stand-in: parameters only."""

    def __init__(self, cfg: ConvNPConfig):
        super().__init__()
        self.cfg = cfg
        self.encoder = Encoder(cfg.encoder_scales)
        self.decoder = Decoder(cfg)

    def forward(self, *a, **k):  # pragma: no cover - arithmetic lives in the engine
        raise RuntimeError("ConvNPModule holds parameters only; use deepsensornz_b200.ConvNP (CUDA engine)")

    def load_state_dict(self, state_dict, strict: bool = True, assign: bool = False):
        """Loads this module's own checkpoints, and -- SURVEY 8(f)4 -- checkpoints written by upstream DeepSensor
        (``model.model.load_state_dict(torch.load(path))``: nzdownscale/downscaler/train.py:243-251 fine-tuning,
        validate_ERA.py:100-111), whose ``nps.Model`` key names differ from the ones here."""
        if set(state_dict.keys()) == set(self.state_dict().keys()):
            return super().load_state_dict(state_dict, strict=strict, assign=assign)
        mapped = map_upstream_state_dict(state_dict, self.state_dict())
        return super().load_state_dict(mapped, strict=True, assign=assign)

    def mlp_dims(self) -> List[int]:
        ls = self.decoder.mlp.layers
        return [ls[0].in_features] + [l.out_features for l in ls] if len(ls) else []


def num_params(module: nn.Module) -> int:
    """``deepsensor.backend.nps.num_params`` (nzdownscale/downscaler/train.py:262)."""
    return sum(int(p.numel()) for p in module.parameters())



def map_upstream_state_dict(upstream: dict, ours: dict, verbose: bool = False) -> dict:
    """Map a checkpoint with foreign key names (upstream ``neuralprocesses`` ``Model``) onto this module's keys.

    The upstream package cannot be imported here, so its key NAMES are unknown; what is known is the
    architecture, hence every tensor SHAPE and the order of layers inside each list (UNet levels 0..L-1 in
    ``before_turn_layers`` / ``after_turn_layers``, MLP layers in order, one length scale per context set in context
    order, then the decoder's).  The mapping is therefore by shape sequence: tensors are grouped by shape, in the order
    the checkpoint lists them, and the i-th upstream tensor of a shape goes to the i-th tensor of that shape here
    (scalars of shape () and (1,) count as one group).  It refuses -- rather than guess -- unless both sides hold
    exactly the same multiset of shapes; ``verbose`` prints the resulting table so it can be audited.
    """
    def key_of(t):
        shp = tuple(t.shape)
        return () if int(torch.as_tensor(t).numel()) == 1 else shp

    def groups(d):
        g = {}
        for k, v in d.items():
            if not torch.is_tensor(v):
                continue
            g.setdefault(key_of(v), []).append(k)
        return g

    gu, go = groups(upstream), groups(ours)
    if {k: len(v) for k, v in gu.items()} != {k: len(v) for k, v in go.items()}:
        only_u = {k: len(v) for k, v in gu.items() if len(v) != len(go.get(k, []))}
        only_o = {k: len(v) for k, v in go.items() if len(v) != len(gu.get(k, []))}
        raise RuntimeError("checkpoint does not match this ConvNP architecture: tensors per shape differ -- checkpoint "
                           f"{only_u} vs model {only_o} (unet_channels / dim_yc / likelihood / aux MLP must equal the "
                           "ones the checkpoint was trained with)")
    out = {}
    for shp, names in go.items():
        for mine, theirs in zip(names, gu[shp]):
            v = upstream[theirs]
            out[mine] = v.reshape(ours[mine].shape).to(ours[mine].dtype)
            if verbose:
                print(f"  {theirs:70s} -> {mine}  {tuple(ours[mine].shape)}")
    n_u = sum(int(v.numel()) for v in upstream.values() if torch.is_tensor(v))
    n_o = sum(int(v.numel()) for v in ours.values())
    if n_u != n_o:
        raise RuntimeError(f"checkpoint holds {n_u} values, the model {n_o}")
    return out
