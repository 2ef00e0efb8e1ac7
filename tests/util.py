"""Shared test helpers: numpy Task -> oracle inputs, small model factory."""
import math

import numpy as np
import torch

from deepsensornz_b200 import ConvNP, Masked, Task
from deepsensornz_b200.task import convert_task_to_nps_args


def oracle_inputs(task: Task):
    """Apply the reference's host-side task ops and hand torch CPU tensors to the oracle."""
    t = ConvNP.modify_task(task)
    ctx, xt, yt, kw = convert_task_to_nps_args(t)
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a))
    contexts = []
    for x, y in ctx:
        xx = tuple(T(v) for v in x) if isinstance(x, tuple) else T(x)
        if isinstance(y, Masked):
            contexts.append((xx, T(y.y), T(y.mask)))
        else:
            contexts.append((xx, T(y), None))
    xt = tuple(T(v) for v in xt) if isinstance(xt, tuple) else T(xt)
    yt = T(np.asarray(yt)) if yt is not None else None
    aux = T(kw["aux_t"]) if "aux_t" in kw else None
    return contexts, xt, yt, aux


def small_model(precision="fp32", ppu=50, dim_yc=(3, 6, 1, 1), seed=0, **kw):
    torch.manual_seed(seed)
    n_lo = kw.pop("n_lo", 140)
    scales = (0.5 / (n_lo - 1) * 0.99286, 0.5 / (n_lo - 1) * 0.99286, 0.5 / 199.0, 0.5 / ppu)
    return ConvNP(dim_yc=dim_yc, dim_yt=1, dim_aux_t=5, internal_density=ppu, encoder_scales=scales,
                  decoder_scale=1.0 / ppu, precision=precision, verbose=False, **kw)


def cpu_params(model):
    return {k: v.detach().cpu().clone() for k, v in model.model.state_dict().items()}


def rel_err(a, b):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp(min=1e-30))


# ---------------------------------------------------------------------------------------------
# BASELINE-shaped models: internal_density 250 -> 304 x 304 internal grid,
# 1400 x 1400 land-mask context, with weights scaled so that the UNet matters to the loss
# ---------------------------------------------------------------------------------------------
def baseline_kwargs(ppu=250, dim_yc=(3, 6, 1, 1)):
    s_lo = 0.5 * (0.99643 - 0.00357) / 139.0
    return dict(dim_yc=dim_yc, dim_yt=1, dim_aux_t=5, internal_density=ppu,
                encoder_scales=(s_lo, s_lo, 0.5 / 1399.0, 0.5 / ppu), decoder_scale=1.0 / ppu,
                unet_channels=(64,) * 4, verbose=False)


def sensitive_state(state: dict, seed: int = 1, gain: float = 1.0) -> dict:
    """He-normal weights (std = gain * sqrt(2 / fan_in)) and N(0, 0.1^2) biases from a seeded CPU generator: with
    torch's default init the activations shrink layer by layer and the loss barely depends on the convolutions
    (VERDICT r01 weak #2); with these the loss moves by tens of percent when one UNet layer is zeroed
    (tests/test_oracle_cpu.py::test_sensitive_weights_make_the_unet_matter)."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    last = max(k for k in state if k.startswith("decoder.mlp.layers.") and k.endswith(".weight"))
    for k in sorted(state):
        v = state[k]
        if v.dim() in (2, 4):
            out[k] = torch.randn(v.shape, generator=g) * gain * math.sqrt(2.0 / v[0].numel())
            if k == last:          # keep the head's pre-activations O(1): a variance of 1e-6 would turn the NLL
                out[k] *= 0.1      # into a test of softplus underflow instead of a test of the path
        elif v.dim() == 1:
            out[k] = torch.randn(v.shape, generator=g) * 0.1
        else:
            out[k] = v.detach().cpu().clone()
    return out


def baseline_model(precision="fp32", dim_yc=(3, 6, 1, 1), ppu=250, seed=1, sensitive=True):
    torch.manual_seed(0)
    m = ConvNP(precision=precision, **baseline_kwargs(ppu, dim_yc))
    if sensitive:
        m.model.load_state_dict(sensitive_state({k: v.detach().cpu() for k, v in m.model.state_dict().items()}, seed))
    return m


def grad_probes(name: str, numel: int, n: int = 8) -> torch.Tensor:
    """[n, numel] fixed +-1 probe vectors per parameter name: <grad, probe_j> pins the direction of a gradient
    without storing the whole tensor in the golden file."""
    g = torch.Generator().manual_seed(abs(hash_name(name)) % (2 ** 31))
    return (torch.randint(0, 2, (n, numel), generator=g, dtype=torch.int8).to(torch.float64) * 2.0 - 1.0)


def hash_name(name: str) -> int:
    h = 2166136261
    for ch in name.encode():
        h = ((h ^ ch) * 16777619) & 0xFFFFFFFF
    return h
