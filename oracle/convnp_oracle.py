"""CPU oracle for the ConvNP hot path -- TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED.  The arithmetic of this path lives in two third-party packages that the
reference pins but does not vendor (``deepsensor==0.3.6`` /root/reference/environment.yml:278,
``neuralprocesses==0.2.6`` environment.yml:300).  Neither is importable in the build container
and the reference holds no tests / golden vectors for the path, so
this file is a *restatement of the published algorithm* (SURVEY.md Appendix A), anchored on the
reference's own call sites:

  * ``ConvNP(data_processor, task_loader, **kw)``        nzdownscale/downscaler/train.py:238-241
  * ``model.loss_fn(task, normalise=True)``               nzdownscale/downscaler/train.py:370
  * ``train_epoch(model, tasks, batch_size=..)``          nzdownscale/downscaler/train.py:388-394
  * ``model.predict(task, X_t=...)``                      nzdownscale/downscaler/validate_ERA.py:88-92
  * model hyper-parameters                                nzdownscale/dataprocess/config.py:2685-2689
  * printed dims / scales of the saved models             experiments/deepsensor/train/validation_precip.ipynb:181-186

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import this module; the product (``deepsensornz_b200``) never does.

Everything is plain torch on the CPU in float32 (float64 for the log-pdf), written for clarity:
dense, untruncated set-convs exactly as the upstream einsums, ``F.conv2d`` for the UNet.

Upstream modules restated (file names are upstream's, for when a copy becomes available):
  neuralprocesses/disc.py                 -> discretise_1d / discretise
  neuralprocesses/coders/setconv/*.py     -> setconv_weights / encode_set / encoder
  neuralprocesses/coders/nn.py (UNet,MLP) -> unet / mlp
  neuralprocesses/likelihood.py, dist/normal.py, model/loglik.py -> het_gaussian / loglik
  deepsensor/model/convnp.py (loss_fn)    -> loss_fn
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple, Union

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor
GridX = Tuple[Tensor, Tensor]  # (x1 [B,1,N1], x2 [B,1,N2])


# --------------------------------------------------------------------------------------
# A.2  Discretisation  (neuralprocesses/disc.py, variant 1 of SURVEY Appendix A.2)
# --------------------------------------------------------------------------------------
def discretise_1d(lo: float, hi: float, ppu: float, margin: float = 0.1, multiple: int = 8):
    """Internal grid along one dimension from the global min / max of all inputs.

    Evaluated in float64 on the host from the float32 extrema (float32 evaluation near
    multiples of ``multiple`` is a bit-exactness hazard; both oracle and product use float64).
    Returns (start, n, res); grid point i is ``float32(start + i * res)``.
    """
    res = 1.0 / float(ppu)
    g_lo = float(lo) - margin - res
    g_hi = float(hi) + margin + res
    n_raw = (g_hi - g_lo) / res + 1.0
    n = math.ceil(n_raw / multiple - 1e-9) * multiple
    start = g_lo - (n - n_raw) * res / 2.0
    start = round(start / res) * res
    return start, int(n), res


def _minmax_dim(xs: Sequence[Union[Tensor, GridX]], d: int):
    lo, hi = math.inf, -math.inf
    for x in xs:
        v = x[d] if isinstance(x, tuple) else x[:, d, :]
        if v.numel() == 0:
            continue
        lo = min(lo, float(v.min()))
        hi = max(hi, float(v.max()))
    return lo, hi


def discretise(xs, ppu, margin=0.1, multiple=8):
    """Grid spec ((start1,n1),(start2,n2),res) over every context input and the target input."""
    (lo1, hi1), (lo2, hi2) = _minmax_dim(xs, 0), _minmax_dim(xs, 1)
    s1, n1, res = discretise_1d(lo1, hi1, ppu, margin, multiple)
    s2, n2, _ = discretise_1d(lo2, hi2, ppu, margin, multiple)
    return (s1, n1), (s2, n2), res


def grid_points(start: float, n: int, res: float) -> Tensor:
    return torch.from_numpy((start + np.arange(n, dtype=np.float64) * res).astype(np.float32))


# --------------------------------------------------------------------------------------
# A.3  SetConv encoder (PrependDensityChannel + SetConv + DivideByFirstChannel + concat)
# --------------------------------------------------------------------------------------
def setconv_weights(x: Tensor, g: Tensor, log_scale: Tensor) -> Tensor:
    """w[b,n,i] = exp(-0.5 (x[b,n]-g[i])^2 / exp(2 log_scale)); x [B,N], g [G] -> [B,N,G]."""
    d2 = (x[:, :, None] - g[None, None, :]) ** 2
    return torch.exp(-0.5 * d2 / torch.exp(2.0 * log_scale))


def encode_set(x, y: Tensor, mask: Optional[Tensor], g1: Tensor, g2: Tensor, log_scale: Tensor,
               eps: float = 1e-2) -> Tensor:
    """One context set -> [B, 1+C, G1, G2] (density first, data divided by density+eps)."""
    if mask is None:
        dens = torch.ones_like(y[:, :1])
        data = y
    else:
        dens = mask
        data = y * mask
    yt = torch.cat([dens, data], dim=1)
    if isinstance(x, tuple):  # gridded: y [B,C,N1,N2]
        w1 = setconv_weights(x[0][:, 0, :], g1, log_scale)  # [B,N1,G1]
        w2 = setconv_weights(x[1][:, 0, :], g2, log_scale)  # [B,N2,G2]
        h = torch.einsum("bcpq,bpi,bqj->bcij", yt, w1, w2)
    else:  # off-grid: x [B,2,N], y [B,C,N]
        w1 = setconv_weights(x[:, 0, :], g1, log_scale)
        w2 = setconv_weights(x[:, 1, :], g2, log_scale)
        h = torch.einsum("bcn,bni,bnj->bcij", yt, w1, w2)
    return torch.cat([h[:, :1], h[:, 1:] / (h[:, :1] + eps)], dim=1)


def encoder(params, contexts, g1, g2, eps=1e-2) -> Tensor:
    outs = []
    for k, (x, y, m) in enumerate(contexts):
        outs.append(encode_set(x, y, m, g1, g2, params[f"encoder.set_convs.{k}.log_scale"], eps))
    return torch.cat(outs, dim=1)


# --------------------------------------------------------------------------------------
# A.4  UNet (k=5, strides (1,2,2,2), bilinear resize-convs, ReLU)
# --------------------------------------------------------------------------------------
def unet(params, x: Tensor, strides=(1, 2, 2, 2)) -> Tensor:
    p = lambda n: params["decoder.unet." + n]
    h = F.conv2d(x, p("initial_linear.weight"), p("initial_linear.bias"))
    hs = []
    for i, s in enumerate(strides):
        h = F.relu(F.conv2d(h, p(f"before_turn_layers.{i}.weight"), p(f"before_turn_layers.{i}.bias"),
                            stride=s, padding=2))
        hs.append(h)
    L = len(strides)

    def up(t, s):
        return t if s == 1 else F.interpolate(t, scale_factor=s, mode="bilinear", align_corners=False)

    h = F.relu(F.conv2d(up(hs[-1], strides[-1]), p(f"after_turn_layers.{L-1}.weight"),
                        p(f"after_turn_layers.{L-1}.bias"), padding=2))
    for i in range(L - 2, -1, -1):
        h = torch.cat([hs[i], h], dim=1)
        h = F.relu(F.conv2d(up(h, strides[i]), p(f"after_turn_layers.{i}.weight"),
                            p(f"after_turn_layers.{i}.bias"), padding=2))
    return F.conv2d(h, p("final_linear.weight"), p("final_linear.bias"))


# --------------------------------------------------------------------------------------
# A.5  SetConv decoder (grid -> targets)
# --------------------------------------------------------------------------------------
def decode(params, z: Tensor, g1: Tensor, g2: Tensor, xt) -> Tensor:
    ls = params["decoder.set_conv.log_scale"]
    if isinstance(xt, tuple):
        w1 = setconv_weights(xt[0][:, 0, :], g1, ls)  # [B,P,G1]
        w2 = setconv_weights(xt[1][:, 0, :], g2, ls)
        return torch.einsum("bcij,bpi,bqj->bcpq", z, w1, w2)
    w1 = setconv_weights(xt[:, 0, :], g1, ls)  # [B,T,G1]
    w2 = setconv_weights(xt[:, 1, :], g2, ls)
    return torch.einsum("bcij,bti,btj->bct", z, w1, w2)


# --------------------------------------------------------------------------------------
# A.6  Augment(aux_t) + MLP + heterogeneous Gaussian head
# --------------------------------------------------------------------------------------
def mlp(params, f: Tensor, n_layers: int) -> Tensor:
    """f [B,C,...] -> [B,2,...]; Linear over the channel axis, ReLU between layers."""
    shp = f.shape
    h = f.reshape(shp[0], shp[1], -1).transpose(1, 2)  # [B,P,C]
    for i in range(n_layers):
        h = F.linear(h, params[f"decoder.mlp.layers.{i}.weight"], params[f"decoder.mlp.layers.{i}.bias"])
        if i < n_layers - 1:
            h = F.relu(h)
    return h.transpose(1, 2).reshape(shp[0], -1, *shp[2:])


def het_gaussian(o: Tensor):
    mean = o[:, 0:1]
    var = 1e-6 + F.softplus(o[:, 1:2])
    return mean, var


def n_mlp_layers(params) -> int:
    n = 0
    while f"decoder.mlp.layers.{n}.weight" in params:
        n += 1
    return n


# --------------------------------------------------------------------------------------
# Spikes-and-slab likelihoods (neuralprocesses/likelihood.py BernoulliGammaLikelihood / SpikesBetaLikelihood over
# dist/spikeslab.py SpikesSlab -- [U], restated from memory like the rest of this file: spikes first, slab last).
# Selected per variable by /root/reference/nzdownscale/dataprocess/config.py:162-169:
#   'bernoulli-gamma'  (precipitation): o = (k~, scale~, l_zero, l_slab), spike at 0, slab Gamma(k, scale)
#   'cnp-spikes-beta'  (humidity):      o = (alpha~, beta~, l_0, l_1, l_slab), spikes at 0 and 1, slab Beta(alpha, beta)
# with a = 1e-6 + softplus(a~) (fp32, like the Gaussian head's variance), log-probabilities = log_softmax(l) and every
# log-pdf evaluated in float64.
# --------------------------------------------------------------------------------------
LIK_EPS = 1e-6
SPIKES = {"bernoulli-gamma": (0.0,), "cnp-spikes-beta": (0.0, 1.0), "spikes-beta": (0.0, 1.0)}


def spike_slab_params(o: Tensor, kind: str):
    """o [B, 2 + n_spikes + 1, ...] -> (a, b, log-probabilities [B, n_spikes + 1, ...]) in float64."""
    a = (LIK_EPS + F.softplus(o[:, 0])).double()
    b = (LIK_EPS + F.softplus(o[:, 1])).double()
    lp = torch.log_softmax(o[:, 2:].double(), dim=1)
    return a, b, lp


def slab_logpdf(a: Tensor, b: Tensor, y: Tensor, kind: str) -> Tensor:
    if kind == "bernoulli-gamma":
        return (a - 1.0) * torch.log(y) - y / b - torch.lgamma(a) - a * torch.log(b)
    x = y.clamp(LIK_EPS, 1.0 - LIK_EPS)
    return (a - 1.0) * torch.log(x) + (b - 1.0) * torch.log1p(-x) - (torch.lgamma(a) + torch.lgamma(b) - torch.lgamma(a + b))


def spike_slab_logpdf(o: Tensor, yt: Tensor, kind: str) -> Tensor:
    """Per-target log-pdf [B, ...] (NaN where the target is NaN)."""
    a, b, lp = spike_slab_params(o, kind)
    y = yt[:, 0].double()
    spikes = SPIKES[kind]
    ok = ~torch.isnan(y)
    y0 = torch.where(ok, y, torch.full_like(y, 0.5))
    is_spike = [y0 == s for s in spikes]
    any_spike = torch.zeros_like(ok)
    for m in is_spike:
        any_spike = any_spike | m
    y_safe = torch.where(any_spike, torch.full_like(y0, 0.5), y0)       # the slab is only evaluated off the spikes
    out = lp[:, len(spikes)] + slab_logpdf(a, b, y_safe, kind)
    for i, m in enumerate(is_spike):
        out = torch.where(m, lp[:, i], out)
    return torch.where(ok, out, torch.full_like(out, float("nan")))


def spike_slab_moments(o: Tensor, kind: str):
    """Mean and variance of the spikes-and-slab distribution, [B, 1, ...] float32 each."""
    a, b, lp = spike_slab_params(o, kind)
    p = lp.exp()
    if kind == "bernoulli-gamma":
        m1, m2 = a * b, a * (a + 1.0) * b * b
        mean, e2 = p[:, 1] * m1, p[:, 1] * m2
    else:
        t = a + b
        m1 = a / t
        m2 = a * b / (t * t * (t + 1.0)) + m1 * m1
        mean, e2 = p[:, 1] + p[:, 2] * m1, p[:, 1] + p[:, 2] * m2
    var = (e2 - mean * mean).clamp(min=0.0)
    return mean.float().unsqueeze(1), var.float().unsqueeze(1)


# --------------------------------------------------------------------------------------
# Full forward / loss
# --------------------------------------------------------------------------------------
def forward(params: Dict[str, Tensor], contexts, xt, aux_t: Optional[Tensor], ppu: float,
            margin: float = 0.1, eps: float = 1e-2, strides=(1, 2, 2, 2), return_internal: bool = False,
            likelihood: str = "cnp"):
    """contexts: list of (x, y, mask|None); off-grid x [B,2,N] y [B,C,N] mask [B,1,N];
    gridded x (x1 [B,1,N1], x2 [B,1,N2]) y [B,C,N1,N2] mask [B,1,N1,N2].  Returns the mean and variance of the
    predictive distribution (``o`` -- the raw likelihood inputs -- in the internals)."""
    mult = 1
    for s in strides:
        mult *= s
    (s1, n1), (s2, n2), res = discretise([c[0] for c in contexts] + [xt], ppu, margin, mult)
    g1, g2 = grid_points(s1, n1, res), grid_points(s2, n2, res)
    enc = encoder(params, contexts, g1, g2, eps)
    z = unet(params, enc, strides)
    f = decode(params, z, g1, g2, xt)
    if aux_t is not None:
        f = torch.cat([f, aux_t], dim=1)
    o = mlp(params, f, n_mlp_layers(params))
    if likelihood in ("cnp", "het"):
        mean, var = het_gaussian(o)
    else:
        mean, var = spike_slab_moments(o, likelihood)
    if return_internal:
        return mean, var, dict(enc=enc, z=z, f=f, o=o, grid=((s1, n1), (s2, n2), res))
    return mean, var


def loglik(mean: Tensor, var: Tensor, yt: Tensor, normalise: bool = True) -> Tensor:
    """Per-task diagonal-Normal log-pdf in float64; NaN targets are skipped; /N if normalise."""
    m, v, y = mean.double(), var.double(), yt.double()
    ok = ~torch.isnan(y)
    y0 = torch.where(ok, y, torch.zeros_like(y))
    lp = -0.5 * (math.log(2.0 * math.pi) + torch.log(v) + (y0 - m) ** 2 / v)
    lp = torch.where(ok, lp, torch.zeros_like(lp))
    B = lp.shape[0]
    lp = lp.reshape(B, -1).sum(dim=1)
    if normalise:
        lp = lp / ok.reshape(B, -1).sum(dim=1).clamp(min=1).double()
    return lp


def loglik_spike_slab(o: Tensor, yt: Tensor, kind: str, normalise: bool = True) -> Tensor:
    lp = spike_slab_logpdf(o, yt, kind)
    ok = ~torch.isnan(lp)
    lp = torch.where(ok, lp, torch.zeros_like(lp))
    B = lp.shape[0]
    s = lp.reshape(B, -1).sum(dim=1)
    if normalise:
        s = s / ok.reshape(B, -1).sum(dim=1).clamp(min=1).double()
    return s


def loss_fn(params, contexts, xt, yt, aux_t, ppu, normalise=True, likelihood: str = "cnp", **kw) -> Tensor:
    mean, var, info = forward(params, contexts, xt, aux_t, ppu, return_internal=True, likelihood=likelihood, **kw)
    if likelihood in ("cnp", "het"):
        return -loglik(mean, var, yt, normalise).mean()
    return -loglik_spike_slab(info["o"], yt, likelihood, normalise).mean()
